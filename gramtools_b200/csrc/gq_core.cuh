// gq_core.cuh — flat index view + the vBWT search-state machine, shared by the sm_100a kernels
// (device) and the host-side k-mer index builder (same code, compiled for the host).
//
// What this replaces in the reference (libgramtools/):
//   * rank on the DNA BWT masks         src/genotype/quasimap/search/BWT_search.cpp:8-22,45-76
//   * marker scan of an SA interval      src/genotype/quasimap/search/vBWT_jump.cpp:94-117
//   * site entry / exit / direct deletion / double entry / double exit jumps
//                                        src/genotype/quasimap/search/vBWT_jump.cpp:3-92,134-265
//   * the per-base loop                  src/genotype/quasimap/quasimap.cpp:243-268
// The reference keeps a std::list of SearchStates and walks it breadth-first per base. States never
// interact, so here one thread walks the same tree depth-first over a private stack of
// variable-length entries; the multiset of final states is identical.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GQ_HD __host__ __device__ __forceinline__
#else
#define GQ_HD inline
#endif

// Byte model of the algorithm (tools/byte_model.py): the TEST-ONLY host emulation (tests/emu, built with
// -DGQ_EMU_COUNTERS) records which 32-byte sectors every strand touches in which structure. On the device, and in
// any build without that define, these macros are the plain accesses.
#if !defined(__CUDA_ARCH__) && defined(GQ_EMU_COUNTERS)
namespace gq {
void gq_emu_touch(const void* p, unsigned bytes);
void gq_emu_phase(int kernel);  // which kernel's work the following touches belong to
}
#define GQ_PHASE(k) ::gq::gq_emu_phase(k)
#define GQ_TOUCH(p, bytes) ::gq::gq_emu_touch((const void*)(p), (unsigned)(bytes))
namespace gq {
template <class T>
inline T& gq_emu_at(T* arr, size_t i) {  // the index expression is evaluated once (it may have side effects)
  gq_emu_touch(arr + i, sizeof(T));
  return arr[i];
}
}
#define GQ_AT(arr, i) ::gq::gq_emu_at((arr), (size_t)(i))
#else
#define GQ_PHASE(k) ((void)0)
#define GQ_TOUCH(p, bytes) ((void)0)
#define GQ_AT(arr, i) ((arr)[i])
#endif

namespace gq {

constexpr uint32_t kBlkShift = 6;     // 64 BWT positions per rank block
constexpr uint32_t kSuperShift = 15;  // 32768 positions per superblock (block counts fit u16)
constexpr uint32_t kTextSuperShift = 15;  // 32768 text positions per marker-rank superblock (relative ranks fit u16)
constexpr uint32_t kNoAllele = 0xFFFFFFFFu;
constexpr uint32_t kNotSnp = 0xFFFFFFFEu;

// 32-byte rank block = one DRAM/L2 sector, fetched with a single 256-bit load.
// Symbol at position j of the block: code = p0 | p1<<1 for A,C,G,T (p2 = 0); p2 = 1 marks a
// non-nucleotide: p0 = 1 variant marker (BWT > 4), p0 = 0 the sentinel.
struct alignas(32) RankBlk {
  uint64_t cnt;  // 4 x u16: A,C,G,T occurrences in [superblock start, block start)
  uint64_t p0, p1, p2;
};

// 8 bytes per 16 PRG text positions: what a width-1 search state needs to walk the text itself instead of
// the BWT (text mode, gq_device.cuh lane_text_step). `codes`: 2-bit base code of position 16g+i at bits
// [2i,2i+2) — the layout of a packed read word — 0 at markers; `info`: bit i (i < 16) = position 16g+i
// holds a variant marker; bits 16..31 = markers in [superblock start, 16g).
struct TextGrp {
  uint32_t codes;
  uint32_t info;
};

struct Node {        // flat coverage-graph node (reference: include/prg/coverage_graph.hpp:40-124)
  uint32_t site;     // 0 outside any site
  int32_t allele;    // -1 for bubble start/end nodes and outside sites
  uint32_t start;    // PRG position of the first base (sequence nodes)
  uint32_t len;      // number of bases (0 for bubble start/end/root/sink)
  uint32_t edge_off; // into edges[]
  uint32_t n_edges;
  uint32_t cov_off;  // offset into the flat per-base counters; kNoAllele if the node holds none
  uint32_t next0;    // edges[edge_off] (target of the first edge), inline: most nodes have exactly one edge, and
                     // the per-base traversal then needs one load per hop instead of two
};

struct KmerState {   // one seed SearchState (reference: kmer_index_types.hpp:24-27)
  uint32_t lo, hi;
  uint32_t path_off; // into kmer_paths: nt (site,allele) pairs then ng site ids
  uint32_t counts;   // nt | ng << 16
};

// Seed-pass view of the k-mer index: per k-mer (seed_off, 4^k + 1) a run of 8-byte entries, one per SUFFIX of
// every seed state of at most kSplitWidth suffixes, one per wider state. A suffix entry carries its text position
// and its LEFT CONTEXT — the up to 12 PRG bases left of the occurrence, up to the first marker or the text start —
// so the other occurrences of the seeding k-mer are rejected against the read without touching SA or text:
//   key = text position SA[i];  aux = 1<<31 | n_ctx << 24 | ctx, base at p-1 in bits 23:22, p-2 in 21:20, ...
// A wide state: key = lo, aux = hi (bit 31 clear: SA indices are below 2^31). seed_state[e] = index of the
// entry's state in kmer_states (its path, needed only by candidates that finish).
// The entries of a k-mer are bucketed by the first d bases of their left context (seed_bucket_bases(k): 2 for
// k <= 11): seed_off has 4^d + 1 offsets per k-mer — bucket 0 = wide states and suffixes with fewer than d context
// bases (a marker or the text start right there), bucket 1 + prefix = the rest — so a strand examines the bucket of
// its own next d bases and bucket 0 instead of every occurrence of its k-mer (a (k+d)-mer seed where the PRG allows
// it: 16x fewer entries at d = 2).
struct KmerSeed {
  uint32_t key, aux;
};
constexpr uint32_t kSplitWidth = 256;  // states of up to this many suffixes are enumerated in the index
constexpr uint32_t kSeedCtxBases = 12;
GQ_HD uint32_t seed_bucket_bases(uint32_t k) { return k <= 11 ? 2u : (k == 12 ? 1u : 0u); }
GQ_HD uint32_t seed_buckets(uint32_t k) { return (1u << (2 * seed_bucket_bases(k))) + 1u; }

struct IndexView {
  uint32_t n;  // SA size = |prg| + 1
  const RankBlk* rank_blk;
  const uint32_t* super_cnt;   // 4 per superblock: C[c] + occurrences of A,C,G,T before the superblock
  const uint32_t* mrank_blk;   // markers in BWT[0, block start)
  const uint32_t* marker_hit;  // 8 per BWT marker occurrence (one 32 B sector): (marker', allele, lo, hi, snp, site_sa,
                               // p_jump, p_site): the jump target; when no marker is adjacent on the far side the SA
                               // interval after the jump (else lo = ~0); for such entries the site's SNP table and
                               // C[site]; p_jump = SA[lo] when lo == hi, p_site = SA[C[site]] (text positions, so a
                               // width-1 state lands in text mode without an SA lookup)
  // text mode (width-1 states): the PRG itself, 2 bits per base + marker flags, and the same jump records
  // in TEXT order (index = rank of the marker among the markers of the text)
  const TextGrp* text_grp;     // one per 16 text positions
  const uint32_t* text_super;  // markers before each superblock of 2^kTextSuperShift positions
  const uint32_t* tmarker_hit; // marker_hit records, text order
  const uint32_t* isa;         // inverse suffix array (final SA index of a text-mode state)
  uint32_t c_base[4];          // first SA index of suffixes starting with A,C,G,T
  // per site slot s = (site_id - 5) / 2
  uint32_t n_slots;
  uint32_t any_nested;         // 1 if some site has a parent (coverage_graph.is_nested)
  const uint32_t* site_sa;     // SA index of the suffix starting with the odd (site entry) marker
  const uint32_t* allele_iv;   // 2 per slot: SA interval [lo,hi] of the even marker
  const uint32_t* par;         // 2 per slot: (parent site id or 0, parent allele)
  const uint32_t* tm_odd;      // target_map[odd marker]: id of the marker just left of the site, or 0
  const uint32_t* tm_even_off; // CSR over target_map[even marker]
  const uint32_t* tm_even;     // 2 per entry: (marker id, direct deletion allele)
  const uint32_t* entry_next;  // 8 per slot: SA interval (lo,hi) after entering the site AND consuming base c
                               // (c = 0..3; lo > hi = no allele ends in c); valid for 'simple' entries only
  const uint32_t* site_snp;    // per slot: kNotSnp, or one byte per base c: the allele (0xFF none) whose single
                               // base is c — a site of distinct 1-base alleles with nothing adjacent
  // end-of-read lookups
  const uint32_t* sa;
  const uint32_t* pos2node;
  const Node* nodes;
  const uint32_t* edges;
  // per slot {apos offset = allele_off[slot] + slot, per-base offset of the site's first base, text position of
  // the first allele's first symbol, text position after the site-end marker}, and the text position of every
  // allele's first symbol (n_alleles + 1 entries per site): coverage of a walk from text positions alone
  const uint32_t* site_rec;
  const uint32_t* apos;
  // k-mer index
  uint32_t k;
  const uint32_t* kmer_bits;   // 4^k bits: k-mer has >= 1 state
  const uint32_t* kmer_off;    // 4^k + 1
  const KmerState* kmer_states;
  const uint32_t* seed_off;    // 4^k * seed_buckets(k) + 1
  const KmerSeed* seed_ent;
  const uint32_t* seed_state;
  const uint32_t* kmer_paths;
};

GQ_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

GQ_HD RankBlk load_blk(const RankBlk* p) {
#if defined(__CUDA_ARCH__)
  RankBlk b;
  // one 256-bit read-only load (LDG.E.ENL2.256 on sm_100a) = exactly one 32 B sector
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.cnt), "=l"(b.p0), "=l"(b.p1), "=l"(b.p2) : "l"(p));
  return b;
#else
  GQ_TOUCH(p, 32);
  return *p;
#endif
}

// occurrences of base code c (0..3) in BWT[0, i), given the block holding position i
GQ_HD uint32_t rank_in_blk(const RankBlk& b, const uint32_t* super4, uint32_t c, uint32_t i) {
  uint64_t m = ~b.p2 & ((c & 1) ? b.p0 : ~b.p0) & ((c & 2) ? b.p1 : ~b.p1);
  uint32_t r = i & 63u;
  uint64_t below = r ? (m & (~0ull >> (64 - r))) : 0ull;
  return GQ_AT(super4, c) + ((uint32_t)(b.cnt >> (16 * c)) & 0xFFFFu) + (uint32_t)popc64(below);
}

// ------------------------------------------------------------------------------------------
// Per-thread work stack. Entry = 5 header words + 2*nt words of (site, allele) + ng site ids.
//   w0: pos (bits 0..27: index of the next read base to consume + 1, so 0 means "read finished")
//       | kind << 28
//   w1: lo   (JUMP: marker id)      w2: hi   (JUMP: allele)
//   w3: nt | ng << 16               w4: start offset of the entry below (kNoAllele at the bottom)
// ------------------------------------------------------------------------------------------
enum : uint32_t { K_SCAN = 0, K_READY = 1, K_JUMP = 2 };
constexpr uint32_t kHdr = 5;

struct Stack {
  uint32_t* mem;   // arena base
  uint32_t top;    // start of the top entry, kNoAllele when empty
  uint32_t limit;  // first word NOT usable by the stack (staging area of final states starts here)
  bool overflow;
};

GQ_HD uint32_t entry_words(uint32_t counts) { return kHdr + 2 * (counts & 0xFFFFu) + (counts >> 16); }

GQ_HD bool stack_empty(const Stack& s) { return s.top == kNoAllele; }

// push a copy of the top entry's paths with a new header; returns false on overflow
GQ_HD bool push_copy_of_top(Stack& s, uint32_t w0, uint32_t w1, uint32_t w2) {
  uint32_t* t = s.mem + s.top;
  uint32_t counts = t[3];
  uint32_t words = entry_words(counts);
  uint32_t ns = s.top + words;
  if (ns + words + 3 > s.limit) {  // +3: room for the in-place growth of one exit on the new top
    s.overflow = true;
    return false;
  }
  uint32_t* d = s.mem + ns;
  d[0] = w0;
  d[1] = w1;
  d[2] = w2;
  d[3] = counts;
  d[4] = s.top;
  for (uint32_t j = kHdr; j < words; ++j) d[j] = t[j];
  s.top = ns;
  return true;
}

GQ_HD void pop(Stack& s) { s.top = s.mem[s.top + 4]; }

// ---- jumps, applied in place to the top entry ---------------------------------------------
// Site exit (vBWT_jump.cpp:51-92): record the allele on the innermost entered site (or open a new
// locus if the read started inside the site) and move to the site-entry marker's SA position.
GQ_HD bool exit_site_in_place(Stack& s, const IndexView& v, uint32_t site, uint32_t allele) {
  uint32_t* t = s.mem + s.top;
  uint32_t nt = t[3] & 0xFFFFu, ng = t[3] >> 16;
  if (s.top + kHdr + 2 * nt + ng + 2 > s.limit) {
    s.overflow = true;
    return false;
  }
  uint32_t* T = t + kHdr;
  uint32_t* G = T + 2 * nt;
  if (ng > 0) {
    --ng;  // G[ng] is the site being left (asserted equal in the reference, :62-64)
    for (uint32_t j = ng; j-- > 0;) G[j + 2] = G[j];
  }
  T[2 * nt] = site;
  T[2 * nt + 1] = allele;
  ++nt;
  t[3] = nt | (ng << 16);
  uint32_t slot = (site - 5) >> 1;
  t[1] = t[2] = GQ_AT(v.site_sa, slot);
  return true;
}

// Process a pending locus sitting on top of the stack (kind K_JUMP). Mirrors the LIFO worklist of
// search_state_vBWT_jumps (vBWT_jump.cpp:134-183) with extend_targets_site_exit (:185-228) and
// extend_targets_site_entry (:230-265).
GQ_HD void process_jump(Stack& s, const IndexView& v) {
  uint32_t* t = s.mem + s.top;
  uint32_t pos = t[0] & 0x0FFFFFFFu;
  uint32_t marker = t[1], allele = t[2];
  if (marker & 1u) {  // odd: leave a site leftwards
    uint32_t site = marker;
    if (!exit_site_in_place(s, v, site, allele)) return;
    while (true) {
      uint32_t nxt = GQ_AT(v.tm_odd, (site - 5) >> 1);
      if (nxt == 0) {  // plain exit: commit, nothing adjacent
        t[0] = pos | (K_READY << 28);
        return;
      }
      if ((nxt & 1u) == 0) {  // exit followed by an entry: not committed, the entry is processed next
        t[0] = pos | (K_JUMP << 28);
        t[1] = nxt;
        t[2] = 0;
        return;
      }
      // double exit: parent's allele from par_map
      uint32_t slot = (site - 5) >> 1;
      uint32_t pal = GQ_AT(v.par, 2 * slot + 1);
      if (!exit_site_in_place(s, v, nxt, pal)) return;
      site = nxt;
    }
  } else {  // even: enter a site from its right end
    uint32_t site = marker - 1;
    uint32_t slot = (site - 5) >> 1;
    uint32_t nt = t[3] & 0xFFFFu, ng = t[3] >> 16;
    if (s.top + kHdr + 2 * nt + ng + 1 > s.limit) {
      s.overflow = true;
      return;
    }
    t[kHdr + 2 * nt + ng] = site;
    ++ng;
    t[3] = nt | (ng << 16);
    t[0] = pos | (K_READY << 28);
    t[1] = GQ_AT(v.allele_iv, 2 * slot);
    t[2] = GQ_AT(v.allele_iv, 2 * slot + 1);
    // direct deletions and double entries hang off the entered state
    uint32_t b = GQ_AT(v.tm_even_off, slot), e = GQ_AT(v.tm_even_off, slot + 1);
    uint32_t base_top = s.top;
    for (uint32_t j = b; j < e; ++j) {
      uint32_t id = GQ_AT(v.tm_even, 2 * j), del = GQ_AT(v.tm_even, 2 * j + 1);
      // copies must be taken from the entered state, which is no longer the top after the 1st push
      uint32_t save_top = s.top;
      uint32_t* src = s.mem + base_top;
      uint32_t counts = src[3];
      uint32_t words = entry_words(counts);
      uint32_t ns = save_top + entry_words(s.mem[save_top + 3]);
      if (ns + words + 3 > s.limit) {
        s.overflow = true;
        return;
      }
      uint32_t* d = s.mem + ns;
      d[0] = pos | (K_JUMP << 28);
      d[1] = id;
      d[2] = (id & 1u) ? del : kNoAllele;
      d[3] = counts;
      d[4] = save_top;
      for (uint32_t w = kHdr; w < words; ++w) d[w] = src[w];
      s.top = ns;
    }
  }
}

// Marker scan of the top entry's SA interval (left_markers_search, vBWT_jump.cpp:94-117): every
// BWT marker in [lo,hi] yields one pending locus pushed as a K_JUMP copy. `blo`/`bhi` are the
// already-fetched blocks of lo and hi+1 when they cover the interval, so the common narrow interval
// costs no extra loads.
GQ_HD void scan_markers(Stack& s, const IndexView& v, uint32_t pos, uint32_t lo, uint32_t hi) {
  uint32_t base_top = s.top;
  for (uint32_t blk = lo >> kBlkShift; blk <= (hi >> kBlkShift); ++blk) {
    RankBlk b = load_blk(v.rank_blk + blk);
    uint64_t m = b.p2 & b.p0;
    uint32_t first = blk << kBlkShift;
    if (first < lo) m &= ~0ull << (lo - first);
    if (hi - first < 63u) m &= ~0ull >> (63u - (hi - first));
    if (!m) continue;
    uint64_t all = b.p2 & b.p0;
    uint32_t mr0 = GQ_AT(v.mrank_blk, blk);
    while (m) {
#if defined(__CUDA_ARCH__)
      uint32_t bit = __ffsll((long long)m) - 1;
#else
      uint32_t bit = (uint32_t)__builtin_ctzll(m);
#endif
      m &= m - 1;
      uint32_t mr = mr0 + (bit ? (uint32_t)popc64(all & (~0ull >> (64 - bit))) : 0u);
      uint32_t marker = GQ_AT(v.marker_hit, 8 * mr), allele = GQ_AT(v.marker_hit, 8 * mr + 1);
      if (marker == 0) continue;
      // copy of the scanned state (always at base_top) with the locus in the header
      uint32_t* src = s.mem + base_top;
      uint32_t counts = src[3];
      uint32_t words = entry_words(counts);
      uint32_t ns = s.top + entry_words(s.mem[s.top + 3]);
      if (ns + words + 3 > s.limit) {
        s.overflow = true;
        return;
      }
      uint32_t* d = s.mem + ns;
      d[0] = pos | (K_JUMP << 28);
      d[1] = marker;
      d[2] = allele;
      d[3] = counts;
      d[4] = s.top;
      for (uint32_t w = kHdr; w < words; ++w) d[w] = src[w];
      s.top = ns;
    }
  }
}

GQ_HD uint64_t marker_bits_in(const RankBlk& b, uint32_t first, uint32_t lo, uint32_t hi) {
  // marker bits of block [first, first+64) restricted to [lo,hi]; caller guarantees overlap
  uint64_t m = b.p2 & b.p0;
  if (first < lo) m &= ~0ull << (lo - first);
  if (hi - first < 63u) m &= ~0ull >> (63u - (hi - first));
  return m;
}

// Does BWT[lo..hi] hold any variant marker? (bwt_markers_mask test of vBWT_jump.cpp:101)
GQ_HD bool interval_has_marker(const IndexView& v, uint32_t lo, uint32_t hi) {
  for (uint32_t blk = lo >> kBlkShift; blk <= (hi >> kBlkShift); ++blk) {
    RankBlk b = load_blk(v.rank_blk + blk);
    if (marker_bits_in(b, blk << kBlkShift, lo, hi)) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// The depth-first driver. `read(i)` returns the base code (0..3) at index i of the strand being
// mapped; `emit(entry)` receives a finished state (all bases consumed). One iteration = one
// transition of the top entry:
//   K_JUMP  : apply the pending locus (may push further entries)
//   K_SCAN  : marker-scan the interval; markers found -> entry becomes K_READY and the loci are
//             pushed above it; none -> consume the next base right away (the hot path: registers
//             only, one or two 32 B sectors per base)
//   K_READY : consume the next base without scanning (committed jump states and already-scanned
//             states; process_read_char_search_states, quasimap.cpp:258-268)
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class ReadFn, class EmitFn>
GQ_HD void run_stack(Stack& s, const IndexView& v, const uint32_t* super_cnt, ReadFn& read, EmitFn& emit) {
  while (!stack_empty(s) && !s.overflow) {
    uint32_t* t = s.mem + s.top;
    uint32_t w0 = t[0];
    uint32_t kind = w0 >> 28, pos = w0 & 0x0FFFFFFFu;
    if (kind == K_JUMP) {
      process_jump(s, v);
      continue;
    }
    if (pos == 0) {
      emit(t);
      pop(s);
      continue;
    }
    uint32_t lo = t[1], hi = t[2];
    bool alive = true;
    while (true) {
      uint32_t b0 = lo >> kBlkShift, bh = hi >> kBlkShift, b1 = (hi + 1) >> kBlkShift;
      RankBlk B0 = load_blk(v.rank_blk + b0);
      RankBlk B1 = (b1 == b0) ? B0 : load_blk(v.rank_blk + b1);
      if (kind == K_SCAN) {
        bool mk;
        if (bh == b0) mk = marker_bits_in(B0, b0 << kBlkShift, lo, hi) != 0;
        else if (bh == b1 && b1 == b0 + 1)
          mk = (marker_bits_in(B0, b0 << kBlkShift, lo, hi) | marker_bits_in(B1, b1 << kBlkShift, lo, hi)) != 0;
        else mk = interval_has_marker(v, lo, hi);
        if (mk) {
          t[0] = pos | (K_READY << 28);
          t[1] = lo;
          t[2] = hi;
          scan_markers(s, v, pos, lo, hi);
          break;
        }
      }
      uint32_t c = read(pos - 1);
      uint32_t r0 = rank_in_blk(B0, super_cnt + 4 * (b0 >> (kSuperShift - kBlkShift)), c, lo);
      uint32_t r1 = rank_in_blk(B1, super_cnt + 4 * (b1 >> (kSuperShift - kBlkShift)), c, hi + 1);
      if (r1 <= r0) {
        alive = false;
        break;
      }
      lo = r0;  // C[c] is folded into the superblock counters
      hi = r1 - 1;
      --pos;
      kind = K_SCAN;
      if (pos == 0) {
        t[0] = pos | (K_SCAN << 28);
        t[1] = lo;
        t[2] = hi;
        break;
      }
    }
    if (!alive) pop(s);
  }
}

// One-base steps for ALL four bases at once (k-mer index construction, index_build.cpp): the marker processing of a
// state — scan, jumps, the states they spawn — does not depend on the base that is consumed next, so it runs once;
// every entry that reaches its extension point (pos == 1) is handed to `ready(entry, lo, hi)` instead of being extended,
// in the order run_stack would have extended (and then emitted) it. The caller extends each ready state by the
// four bases with rank queries.
#if defined(__CUDACC__)
#pragma nv_exec_check_disable
#endif
template <class ReadyFn>
GQ_HD void run_stack_ready(Stack& s, const IndexView& v, ReadyFn& ready) {
  while (!stack_empty(s) && !s.overflow) {
    uint32_t* t = s.mem + s.top;
    uint32_t w0 = t[0];
    uint32_t kind = w0 >> 28, pos = w0 & 0x0FFFFFFFu;
    if (kind == K_JUMP) {
      process_jump(s, v);
      continue;
    }
    uint32_t lo = t[1], hi = t[2];
    if (kind == K_SCAN && interval_has_marker(v, lo, hi)) {
      t[0] = pos | (K_READY << 28);
      scan_markers(s, v, pos, lo, hi);
      continue;
    }
    ready(t, lo, hi);
    pop(s);
  }
}

}  // namespace gq
