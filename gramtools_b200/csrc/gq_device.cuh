// gq_device.cuh — per-strand device functions of the quasimap kernels (search + coverage).
//
// Written once with GQ_DEV so that tests/emu can compile the very same functions for the host and
// run them single-threaded for debugging (tests only — libgq.so contains the CUDA build alone and has
// no CPU path). Reference citations are next to each function.
#pragma once
#include "gq_core.cuh"
#include "kernels.cuh"

#if defined(__CUDA_ARCH__)
#define GQ_LDG(p) __ldg(p)
#else
#define GQ_LDG(p) (*(GQ_TOUCH((p), sizeof(*(p))), (p)))
#endif
#if defined(__CUDACC__)
#define GQ_DEV __host__ __device__
#else
#define GQ_DEV
#endif

#ifndef GQ_COV_CHUNK
#define GQ_COV_CHUNK 2  // per-site records fetched together by the coverage table route
#endif

namespace gq {

// path statistics of the host emulation (tests/emu): which route strands take. No-op on the device.
#if !defined(__CUDA_ARCH__) && defined(GQ_EMU_COUNTERS)
extern unsigned long long gq_emu_counters[32];
#define GQ_COUNT(i) (++gq_emu_counters[i])
#else
#define GQ_COUNT(i) ((void)0)
#endif

GQ_DEV inline uint32_t gq_atomic_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  GQ_TOUCH(p, 4);
  uint32_t o = *p;
  *p += v;
  return o;
#endif
}
// fire-and-forget add (RED.ADD: no return value travels back from L2)
GQ_DEV inline void gq_red_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  GQ_TOUCH(p, 4);
  *p += v;
#endif
}
GQ_DEV inline uint32_t gq_atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicCAS(p, cmp, v);
#else
  uint32_t o = *p;
  if (o == cmp) *p = v;
  return o;
#endif
}
GQ_DEV inline void gq_atomic_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  atomicOr(p, v);
#else
  *p |= v;
#endif
}
// append-style counter bump; on the device the lanes that arrive together issue ONE atomic
GQ_DEV inline uint32_t gq_atomic_inc_aggregated(uint32_t* p) {
#if defined(__CUDA_ARCH__)
  const unsigned m = __activemask();
  const int leader = __ffs(m) - 1;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(p, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
#else
  GQ_TOUCH(p, 4);
  return (*p)++;
#endif
}
GQ_DEV inline uint32_t gq_funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // bits [sh, sh+32) of hi:lo
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
GQ_DEV inline uint32_t gq_clz(uint32_t x) {  // 32 for x == 0
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)x);
#else
  return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
GQ_DEV inline uint32_t gq_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}
GQ_DEV inline void gq_threadfence() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}

GQ_DEV inline uint32_t pair_reverse32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  x = __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  x = (x >> 16) | (x << 16);
#endif
  return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);  // bit reversal, then un-swap inside each pair
}

// Word `m` of the reverse complement of a packed read of L bases (reverse_complement_read,
// quasimap.cpp:273-298): base j of the reverse strand = complement of base L-1-j, stored where a read keeps its
// base j. revcomp_kernel builds the reverse strands of a batch slice once, so every later pass treats an odd
// strand exactly like an even one — no per-word complement/reversal in the walks.
GQ_DEV inline uint32_t revcomp_word(const uint32_t* w, uint32_t L, uint32_t m) {
  const uint32_t s_hi = L - 1u - 16u * m;  // source base of the word's first base
  const uint32_t wi = s_hi >> 4, q = s_hi & 15u;
  // sources s_hi, s_hi-1, ... from the top bit pair down
  const uint32_t hi = GQ_LDG(w + wi), lo = (wi && q < 15u) ? GQ_LDG(w + wi - 1) : 0u;
  const uint32_t x = q == 15u ? hi : ((hi << (2 * (15u - q))) | (lo >> (2 * (q + 1u))));
  const uint32_t c = L - 16u * m;  // bases in this word (>= 1)
  const uint32_t y = pair_reverse32(~x);
  return c >= 16u ? y : (y & ((1u << (2 * c)) - 1u));
}

// Strand cursor over the 2-bit packed strand (the read, or its prepared reverse complement): the bases are
// consumed right to left (backward search). `cw` is a 64-bit shift register holding the next `left` bases, the
// next one in its top two bits. The register is topped up with the preceding word whenever it holds 16 bases or
// fewer, so (while the strand has that many left) at least 16 bases are always available to a text step.
struct ReadCursor {
  const uint32_t* w;
  uint32_t L;
  uint64_t cw;
  uint32_t left, wi;
  GQ_DEV inline void refill() {
    if (left <= 16 && wi != 0) {
      --wi;
      cw |= (uint64_t)GQ_LDG(w + wi) << (32 - 2 * left);  // left <= 16: the new 16 bases follow the valid ones
      left += 16;
    }
  }
  // position the cursor so that peek() returns the base at index pos-1 (pos >= 1)
  GQ_DEV inline void seek(uint32_t pos) {
    const uint32_t i = pos - 1;
    wi = i >> 4;
    const uint32_t q = i & 15u;
    cw = (uint64_t)(GQ_LDG(w + wi) << (2 * (15 - q))) << 32;
    left = q + 1;
    refill();
  }
  GQ_DEV inline uint32_t peek() const { return (uint32_t)(cw >> 62); }
  GQ_DEV inline uint32_t top32() const { return (uint32_t)(cw >> 32); }
  GQ_DEV inline void advance() {
    cw <<= 2;
    --left;
    refill();
  }
  // consume r <= min(left, 16) bases at once (a text step compares up to a whole word)
  GQ_DEV inline void advance_n(uint32_t r) {
    cw <<= 2 * r;
    left -= r;
    refill();
  }
};

// Finished states are staged at the top of the thread's arena (growing down) as
// [lo, hi, nt, ng, (site,allele)*nt, (site,-1)*ng]. Path-less states are split per SA index by the
// site/allele of the node they start in (handle_allele_encapsulated_state,
// encapsulated_search.cpp:30-88).
struct EmitStage {
  Stack* s;
  const IndexView* v;
  uint32_t n_states;

  GQ_DEV inline bool reserve(uint32_t words, uint32_t*& dst) {
    uint32_t top_end = s->top + entry_words(s->mem[s->top + 3]);
    if (top_end + words > s->limit) {
      s->overflow = true;
      return false;
    }
    s->limit -= words;
    dst = s->mem + s->limit;
    return true;
  }
  GQ_DEV inline void put(uint32_t lo, uint32_t hi, uint32_t site, uint32_t allele, bool with_path) {
    uint32_t* d;
    if (!reserve(with_path ? 6 : 4, d)) return;
    d[0] = lo;
    d[1] = hi;
    d[2] = with_path ? 1u : 0u;
    d[3] = 0;
    if (with_path) {
      d[4] = site;
      d[5] = allele;
    }
    ++n_states;
  }
  GQ_DEV void operator()(const uint32_t* t) {
    uint32_t lo = t[1], hi = t[2], nt = t[3] & 0xFFFFu, ng = t[3] >> 16;
    if (nt | ng) {
      uint32_t* d;
      if (!reserve(4 + 2 * nt + 2 * ng, d)) return;
      d[0] = lo;
      d[1] = hi;
      d[2] = nt;
      d[3] = ng;
      const uint32_t* T = t + kHdr;
      for (uint32_t j = 0; j < 2 * nt; ++j) d[4 + j] = T[j];
      const uint32_t* G = T + 2 * nt;
      for (uint32_t j = 0; j < ng; ++j) {
        d[4 + 2 * nt + 2 * j] = G[j];
        d[4 + 2 * nt + 2 * j + 1] = kNoAllele;
      }
      ++n_states;
      return;
    }
    bool cached = false;
    uint32_t c_lo = 0, c_hi = 0, c_site = 0, c_al = 0;
    for (uint32_t i = lo;; ++i) {
      uint32_t p = GQ_LDG(v->sa + i);
      const Node& nd = GQ_AT(v->nodes, GQ_LDG(v->pos2node + p));
      uint32_t site = nd.site, al = (uint32_t)nd.allele;
      if (site == 0) {
        if (cached) put(c_lo, c_hi, c_site, c_al, true);
        cached = false;
        put(i, i, 0, 0, false);
      } else if (!cached) {
        cached = true;
        c_lo = c_hi = i;
        c_site = site;
        c_al = al;
      } else if (site == c_site && al == c_al) {
        c_hi = i;
      } else {
        put(c_lo, c_hi, c_site, c_al, true);
        c_lo = c_hi = i;
        c_site = site;
        c_al = al;
      }
      if (s->overflow || i == hi) break;
    }
    if (cached) put(c_lo, c_hi, c_site, c_al, true);
  }
};

// ------------------------------------------------------------------------------------------------
// Lane state machine of the search kernel. One lane maps one strand at a time; the warp runs ONE flat
// loop in which every lane executes the same unit operation per iteration, so lanes re-converge every
// iteration instead of drifting apart inside nested per-read loops:
//   lane_refill    : idle lane takes the next strand, seeds its stack from the k-mer index
//   lane_to_text / lane_text_step : width-1 interval — its text position, then up to 16 bases per step in the
//                    packed PRG text (see "Text mode" below)
//   lane_step_wide : wider interval — one read base: marker test + two rank queries
//   lane_event     : everything else — marker scan / jumps, emitting finished states, pops, strand end
// Order of work differs from quasimap_read (quasimap.cpp:159-194) without changing results: the search
// runs first and the k-mer filter (all_read_kmers_occur_in_index, :212-225) is evaluated afterwards only
// for strands that produced no state (classify_strand) — a strand that maps end to end necessarily has
// all of its k-mers in the index, which holds every k-mer with >= 1 search state.
// ------------------------------------------------------------------------------------------------
enum LaneState : uint32_t {
  LS_IDLE = 0, LS_RUN = 1 /* width-1 interval */, LS_EV_SCAN = 2, LS_EV_POP = 3, LS_EV_TOP = 4, LS_EV_WIDE = 5,
  LS_RUNW = 6 /* wider interval */,
  LS_TEXT = 7 /* width-1 interval followed in the PRG text (lane_text_step): `p` is authoritative, lo/hi are not */,
  LS_EV_TSCAN = 8 /* text mode reached a marker: `mr` is its rank among the markers of the text */
};

struct Lane {
  uint32_t state;
  uint32_t strand;
  ReadCursor rd;
  Stack s;
  uint32_t n_states;
  uint32_t arena_words;
  // top entry cached in registers while state == LS_RUN
  uint32_t pos, lo, hi, kind;
  uint32_t mr;  // LS_EV_SCAN of a width-1 interval: rank of the marker at BWT[lo] among all BWT markers
  uint32_t p;   // LS_TEXT: text position SA[lo] of the state's single suffix
};

GQ_DEV inline void lane_writeback(Lane& ln, uint32_t kind) {
  uint32_t* t = ln.s.mem + ln.s.top;
  t[0] = ln.pos | (kind << 28);
  t[1] = ln.lo;
  t[2] = ln.hi;
}

// decide what the (memory) top of the stack needs next
GQ_DEV inline void lane_load_top(Lane& ln) {
  const uint32_t* t = ln.s.mem + ln.s.top;
  uint32_t w0 = t[0];
  ln.kind = w0 >> 28;
  ln.pos = w0 & 0x0FFFFFFFu;
  ln.lo = t[1];
  ln.hi = t[2];
  ln.state = (ln.kind == K_JUMP || ln.pos == 0) ? LS_EV_TOP : (ln.lo == ln.hi ? LS_RUN : LS_RUNW);
  if (ln.state != LS_EV_TOP) ln.rd.seek(ln.pos);  // stack entries resume at their own read position
}

constexpr uint32_t kSurvGeneral = 0x10000u;  // SeedOut::surv_cnt flag: the strand is on the general kernel's list
constexpr uint32_t kSurvListed = 0x20000u;   //  ... the strand is on mapped_list (the coverage work list)

// append a mapped strand to the coverage work list, once (the text kernel may have listed it already)
GQ_DEV inline void list_mapped(const SearchOut& o, uint32_t strand) {
  if (o.listed) {
#if defined(__CUDA_ARCH__)
    const uint32_t old = atomicOr(o.listed + strand, kSurvListed);
#else
    const uint32_t old = GQ_AT(o.listed, strand);
    GQ_AT(o.listed, strand) |= kSurvListed;
#endif
    if (old & kSurvListed) return;
  }
  GQ_AT(o.mapped_list, gq_atomic_inc_aggregated(o.n_mapped)) = strand;
}

GQ_DEV inline void lane_finish_strand(Lane& ln, const SearchOut& o) {
  uint32_t words = ln.arena_words - ln.s.limit;
  if (!ln.s.overflow && words) {
    uint32_t off = gq_atomic_add(o.pool_used, words);
    if (off + words > o.pool_cap) ln.s.overflow = true;
    else {
      for (uint32_t w = 0; w < words; ++w) GQ_AT(o.pool, off + w) = ln.s.mem[ln.s.limit + w];
      GQ_AT(o.st_off, ln.strand) = off;
      GQ_AT(o.st_words, ln.strand) = words;
      GQ_AT(o.st_count, ln.strand) = ln.n_states;
    }
  }
  if (ln.s.overflow) {
    GQ_AT(o.status, ln.strand) = ST_OVERFLOW;
    GQ_AT(o.overflow_list, gq_atomic_add(o.n_overflow, 1u)) = ln.strand;
  } else if (ln.n_states) {
    GQ_AT(o.status, ln.strand) = ST_MAPPED;
    list_mapped(o, ln.strand);
  } else
    GQ_AT(o.status, ln.strand) = ST_UNCLASSIFIED;
  ln.state = LS_IDLE;
}

// Start `strand` on this lane: seed with the index entry of its last k-mer (quasimap.cpp:178,235-241).
// code of the last k bases of a packed strand (L >= k)
GQ_DEV inline uint32_t seeding_kmer_code(const uint32_t* w, uint32_t L, uint32_t k) {
  const uint32_t j0 = L - k, wi = j0 >> 4, n_words = (L + 15) >> 4;
  const uint32_t wlo = GQ_LDG(w + wi), whi = (wi + 1 < n_words) ? GQ_LDG(w + wi + 1) : 0u;
  return gq_funnelshift_r(wlo, whi, 2 * (j0 & 15u)) & ((k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u));
}

GQ_DEV inline void lane_refill(Lane& ln, const IndexView& v, const BatchView& b, const SearchOut& o, uint32_t strand,
                               uint32_t* arena, uint32_t arena_words) {
  const uint32_t r = strand >> 1;
  const uint32_t L = GQ_AT(b.len, r);
  const uint32_t k = v.k;
  ln.strand = strand;
  ln.state = LS_IDLE;
  if (L == 0) {  // non-ACGT read, emptied by the encoder: skipped (quasimap.cpp:108-113)
    GQ_AT(o.status, strand) = ST_SKIPPED;
    return;
  }
  if (L < k) {  // cannot be seeded (UB in the reference, quasimap.cpp:206-210): counted as missing k-mer
    GQ_AT(o.status, strand) = ST_MISSING_KMER;
    return;
  }
  ln.rd = ReadCursor{b.strand_words(strand, GQ_AT(b.word_off, r)), L, 0, 0, 0};
  // seeding k-mer = last k bases of the strand; its code (base j at bits [2j,2j+2)) is a bit-field of the
  // packed strand
  const uint32_t code = seeding_kmer_code(ln.rd.w, L, k);
  uint32_t sb = GQ_LDG(v.kmer_off + code), se = GQ_LDG(v.kmer_off + code + 1);
  if (sb == se) {  // the seeding k-mer itself is not indexed: the k-mer filter fails
    GQ_AT(o.status, strand) = ST_MISSING_KMER;
    return;
  }
  ln.s.mem = arena;
  ln.s.limit = arena_words;
  ln.s.overflow = false;
  ln.s.top = kNoAllele;
  ln.n_states = 0;
  ln.arena_words = arena_words;
  uint32_t sp = 0;
  for (uint32_t j = sb; j < se; ++j) {
    KmerState ks = GQ_AT(v.kmer_states, j);
    uint32_t words = entry_words(ks.counts);
    if (sp + words + 3 > ln.s.limit) {
      ln.s.overflow = true;
      break;
    }
    uint32_t* t = ln.s.mem + sp;
    t[0] = (L - k) | (K_SCAN << 28);
    t[1] = ks.lo;
    t[2] = ks.hi;
    t[3] = ks.counts;
    t[4] = ln.s.top;
    for (uint32_t w = kHdr; w < words; ++w) t[w] = GQ_LDG(v.kmer_paths + ks.path_off + (w - kHdr));
    ln.s.top = sp;
    sp += words;
  }
  if (ln.s.overflow) {
    lane_finish_strand(ln, o);
    return;
  }
  lane_load_top(ln);
}

// One base for a lane in LS_RUNW: SA interval wider than one suffix (the first bases after seeding, and
// states that just entered a site). Two rank queries (BWT_search.cpp:45-76) + the marker test of the
// interval (vBWT_jump.cpp:100-101).
template <class SuperPtr>
GQ_DEV inline void lane_step_wide(Lane& ln, const IndexView& v, SuperPtr super_c) {
  const uint32_t lo = ln.lo, hi = ln.hi;
  const uint32_t b0 = lo >> kBlkShift, bh = hi >> kBlkShift, b1 = (hi + 1) >> kBlkShift;
  const RankBlk B0 = load_blk(v.rank_blk + b0);
  const RankBlk B1 = (b1 == b0) ? B0 : load_blk(v.rank_blk + b1);
  if (ln.kind == K_SCAN) {
    uint64_t mk;
    if (bh == b0) mk = marker_bits_in(B0, b0 << kBlkShift, lo, hi);
    else if (bh == b1 && b1 == b0 + 1)
      mk = marker_bits_in(B0, b0 << kBlkShift, lo, hi) | marker_bits_in(B1, b1 << kBlkShift, lo, hi);
    else {  // interval wider than the two fetched blocks: rare, resolved in the event path
      ln.state = LS_EV_WIDE;
      return;
    }
    if (mk) {
      ln.mr = kNoAllele;  // interval state: the event path computes marker ranks itself
      ln.state = LS_EV_SCAN;
      return;
    }
  }
  const uint32_t c = ln.rd.peek();
  const uint64_t x0 = (c & 1u) ? 0ull : ~0ull, x1 = (c & 2u) ? 0ull : ~0ull;
  const uint64_t m0 = ~B0.p2 & (B0.p0 ^ x0) & (B0.p1 ^ x1);
  const uint64_t m1 = ~B1.p2 & (B1.p0 ^ x0) & (B1.p1 ^ x1);
  const uint32_t ra = lo & 63u, rb = (hi + 1) & 63u;
  const uint32_t r0 = GQ_AT(super_c, 4 * (b0 >> (kSuperShift - kBlkShift)) + c) + ((uint32_t)(B0.cnt >> (16 * c)) & 0xFFFFu) +
                      (uint32_t)popc64(m0 & ((1ull << ra) - 1));
  const uint32_t r1 = GQ_AT(super_c, 4 * (b1 >> (kSuperShift - kBlkShift)) + c) + ((uint32_t)(B1.cnt >> (16 * c)) & 0xFFFFu) +
                      (uint32_t)popc64(m1 & ((1ull << rb) - 1));
  if (r1 <= r0) {
    ln.state = LS_EV_POP;
    return;
  }
  ln.lo = r0;
  ln.hi = r1 - 1;
  ln.kind = K_SCAN;
  ln.state = (ln.lo == ln.hi) ? LS_RUN : LS_RUNW;
  ln.rd.advance();
  if (--ln.pos == 0) ln.state = LS_EV_TOP;
}

// ------------------------------------------------------------------------------------------------
// Text mode. A width-1 SA interval [lo,lo] is ONE suffix of the PRG, at text position p = SA[lo]. Extending
// it backwards by base c (BWT_search.cpp:45-76) succeeds iff BWT[lo] = prg[p-1] = c, and the new interval
// is the single suffix at p-1 (LF(lo) = ISA[p-1]); a marker at prg[p-1] is the marker the reference finds
// in BWT[lo] (vBWT_jump.cpp:100-114), and p = 0 puts the sentinel there. So while the interval stays one
// suffix wide the whole backward search is a comparison of the packed read with the packed PRG, up to 16
// bases per step instead of one rank query per base — same states, same jumps, same final SA index
// (ISA[p], read once when the read is exhausted).
// ------------------------------------------------------------------------------------------------
GQ_DEV inline void lane_to_text(Lane& ln, const IndexView& v) {  // LS_RUN -> LS_TEXT
  ln.p = GQ_LDG(v.sa + ln.lo);
  ln.state = LS_TEXT;
}

GQ_DEV inline void lane_text_step(Lane& ln, const IndexView& v) {
  const uint32_t p = ln.p;
  if (p == 0) {  // BWT[lo] is the sentinel
    ln.state = LS_EV_POP;
    return;
  }
  // the positions [16g, p-1] of the group holding p-1, aligned so that p-1 sits where the cursor keeps
  // the next read base (top bit pair / top flag bit)
  const uint32_t q = p - 1, g = q >> 4, j = (q & 15u) + 1, sh = 16 - j;
#if defined(__CUDA_ARCH__)
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(v.text_grp) + g);
  const uint32_t codes = raw.x, info = raw.y;
#else
  const uint32_t codes = GQ_AT(v.text_grp, g).codes, info = GQ_AT(v.text_grp, g).info;
#endif
  const uint32_t run_mis = gq_clz(ln.rd.top32() ^ (codes << (2 * sh))) >> 1;  // equal bases from the top (<= 16)
  const uint32_t run_mark = gq_clz(info << (16 + sh));                   // marker-free positions from the top
  uint32_t limit = j < ln.rd.left ? j : ln.rd.left;
  limit = limit < ln.pos ? limit : ln.pos;
  if (run_mark < limit && run_mark <= run_mis) {
    // a marker after run_mark further bases: committed jump states do not scan again (K_READY), any
    // other state jumps (lane_event_scan, text flavour)
    if (run_mark == 0 && ln.kind != K_SCAN) {
      ln.state = LS_EV_POP;
      return;
    }
    ln.p = p - run_mark;
    ln.pos -= run_mark;
    ln.rd.advance_n(run_mark);
    ln.kind = K_SCAN;
    const uint32_t below = (1u << ((q - run_mark) & 15u)) - 1u;
    ln.mr = GQ_LDG(v.text_super + (g >> (kTextSuperShift - 4))) + (info >> 16) + gq_popc(info & below);
    ln.state = LS_EV_TSCAN;
    return;
  }
  if (run_mis < limit) {  // the PRG continues with another base
    ln.state = LS_EV_POP;
    return;
  }
  ln.p = p - limit;
  ln.pos -= limit;
  ln.rd.advance_n(limit);
  ln.kind = K_SCAN;
  if (ln.pos == 0) {  // finished: the state's SA index, for the record lane_event_top writes
    ln.lo = ln.hi = GQ_LDG(v.isa + ln.p);
    ln.state = LS_EV_TOP;
  }
}

// after a transition of the rare path: finish the strand, or cache the new top
GQ_DEV inline void lane_after_event(Lane& ln, const SearchOut& o) {
  if (ln.s.overflow || stack_empty(ln.s)) {
    lane_finish_strand(ln, o);
    return;
  }
  lane_load_top(ln);
}

// LS_EV_SCAN / LS_EV_WIDE: markers in the interval (left_markers_search, vBWT_jump.cpp:94-117)
GQ_DEV inline void lane_event_scan(Lane& ln, const IndexView& v, const SearchOut& o) {
  if (ln.state == LS_EV_WIDE && !interval_has_marker(v, ln.lo, ln.hi)) {
    ln.kind = K_READY;  // scanned, nothing found: extend without re-scanning
    ln.state = LS_RUNW;
    return;
  }
  const bool text = ln.state == LS_EV_TSCAN;
  if (text || ln.lo == ln.hi) {
    // Single suffix preceded by a marker: the un-jumped state cannot be extended by any base (its only
    // BWT symbol is the marker), so the jump replaces it in place instead of being pushed above it.
    uint32_t mr = ln.mr;
    if (!text && mr == kNoAllele) {  // width-1 interval reached through the wide step
      const uint32_t blk = ln.lo >> kBlkShift, bit = ln.lo & 63u;
      const RankBlk B = load_blk(v.rank_blk + blk);
      mr = GQ_LDG(v.mrank_blk + blk) + (uint32_t)popc64(B.p2 & B.p0 & ((1ull << bit) - 1));
    }
    // one 32 B sector: jump target, post-jump interval, SNP table + C[site] of a simple entry, and the text
    // positions behind the post-jump SA indices (the state continues in text mode without an SA lookup)
    const uint32_t* jr = (text ? v.tmarker_hit : v.marker_hit) + 8 * (size_t)mr;
    const uint32_t marker = GQ_LDG(jr), allele = GQ_LDG(jr + 1), jlo = GQ_LDG(jr + 2), jhi = GQ_LDG(jr + 3);
    const uint32_t snp = GQ_LDG(jr + 4), entered_site_sa = GQ_LDG(jr + 5);
    const uint32_t p_jump = GQ_LDG(jr + 6), p_site = GQ_LDG(jr + 7);
    if (marker == 0) {
      ln.state = LS_EV_POP;
      return;
    }
    uint32_t* t = ln.s.mem + ln.s.top;
    if (jlo != kNoAllele) {
      // pre-resolved jump (no adjacent marker on the other side): the whole vBWT jump is a path update
      // + the SA interval stored with the marker occurrence; registers stay authoritative for lo/hi/pos
      uint32_t nt = t[3] & 0xFFFFu, ng = t[3] >> 16;
      if (ln.s.top + kHdr + 2 * nt + ng + 2 > ln.s.limit) {
        ln.s.overflow = true;
        lane_finish_strand(ln, o);
        return;
      }
      uint32_t* T = t + kHdr;
      if (marker & 1u) {  // exit through `allele` (exit_site_in_place)
        uint32_t* G = T + 2 * nt;
        if (ng > 0) {
          --ng;
          for (uint32_t j = ng; j-- > 0;) G[j + 2] = G[j];
        }
        T[2 * nt] = marker;
        T[2 * nt + 1] = allele;
        ++nt;
      } else {
        // Enter the site whose end marker this is, and consume the next base right away: the entered
        // state is a committed jump state (extended without another scan), so entry + extension is a
        // table lookup. For a site of distinct single-base alleles the exit that follows is folded in
        // as well (entry, allele base, exit = one event) unless the read ends inside the site.
        const uint32_t slot = (marker - 6) >> 1;
        const uint32_t c = ln.rd.peek();
        if (snp != kNotSnp && ln.pos >= 2) {
          const uint32_t a = (snp >> (8 * c)) & 0xFFu;
          if (a == 0xFFu) {
            ln.state = LS_EV_POP;
            return;
          }
          for (uint32_t j = ng; j-- > 0;) T[2 * nt + 2 + j] = T[2 * nt + j];  // open sites stay behind T (room checked above)
          T[2 * nt] = marker - 1;
          T[2 * nt + 1] = a;
          t[3] = (nt + 1) | (ng << 16);
          ln.lo = ln.hi = entered_site_sa;
          ln.p = p_site;
          ln.pos -= 1;
          ln.rd.advance();
          ln.kind = K_READY;
          ln.state = LS_TEXT;
          return;
        }
        T[2 * nt + ng] = marker - 1;
        ++ng;
        t[3] = nt | (ng << 16);
        const uint32_t nlo = GQ_LDG(v.entry_next + 8 * slot + 2 * c), nhi = GQ_LDG(v.entry_next + 8 * slot + 2 * c + 1);
        if (nhi + 1 <= nlo) {  // no allele ends in c
          ln.state = LS_EV_POP;
          return;
        }
        ln.lo = nlo;
        ln.hi = nhi;
        ln.kind = K_SCAN;
        ln.state = (nlo == nhi) ? LS_RUN : LS_RUNW;
        ln.rd.advance();
        if (--ln.pos == 0) ln.state = LS_EV_TOP;
        return;
      }
      t[3] = nt | (ng << 16);
      ln.lo = jlo;
      ln.hi = jhi;
      ln.p = p_jump;
      ln.kind = K_READY;
      ln.state = (jlo == jhi) ? LS_TEXT : LS_RUNW;
      return;
    }
    t[0] = ln.pos | (K_JUMP << 28);
    t[1] = marker;
    t[2] = allele;
    process_jump(ln.s, v);
  } else {
    lane_writeback(ln, K_READY);
    scan_markers(ln.s, v, ln.pos, ln.lo, ln.hi);
  }
  lane_after_event(ln, o);
}

// LS_EV_POP: the top state died
GQ_DEV inline void lane_event_pop(Lane& ln, const SearchOut& o) {
  pop(ln.s);
  lane_after_event(ln, o);
}

// LS_EV_TOP: a pending locus (K_JUMP) or a finished state (pos == 0) sits on top
GQ_DEV inline void lane_event_top(Lane& ln, const IndexView& v, const SearchOut& o) {
  uint32_t* t = ln.s.mem + ln.s.top;
  // a state that finished in a step is still only in registers (equal to memory if it was loaded)
  if (ln.pos == 0 && ln.kind != K_JUMP) lane_writeback(ln, K_SCAN);
  if ((t[0] >> 28) == K_JUMP) process_jump(ln.s, v);
  else if (t[4] == kNoAllele && ln.n_states == 0 && t[3] != 0 && ln.s.limit == ln.arena_words) {
    // the strand's only state, with a path: write its record straight into the pool
    const uint32_t nt = t[3] & 0xFFFFu, ng = t[3] >> 16, words = 4 + 2 * nt + 2 * ng;
    const uint32_t off = gq_atomic_add(o.pool_used, words);
    if (off + words > o.pool_cap) {
      ln.s.overflow = true;
      lane_finish_strand(ln, o);
      return;
    }
    uint32_t* d = o.pool + off;
    GQ_TOUCH(d, 4 * words);
    d[0] = t[1];
    d[1] = t[2];
    d[2] = nt;
    d[3] = ng;
    const uint32_t* T = t + kHdr;
    for (uint32_t j = 0; j < 2 * nt; ++j) d[4 + j] = T[j];
    for (uint32_t j = 0; j < ng; ++j) {
      d[4 + 2 * nt + 2 * j] = T[2 * nt + j];
      d[4 + 2 * nt + 2 * j + 1] = kNoAllele;
    }
    GQ_AT(o.st_off, ln.strand) = off;
    GQ_AT(o.st_words, ln.strand) = words;
    GQ_AT(o.st_count, ln.strand) = 1;
    GQ_AT(o.status, ln.strand) = ST_MAPPED;
    list_mapped(o, ln.strand);
    ln.state = LS_IDLE;
    return;
  } else {
    EmitStage emit{&ln.s, &v, ln.n_states};
    emit(t);
    ln.n_states = emit.n_states;
    pop(ln.s);
  }
  lane_after_event(ln, o);
}

GQ_DEV inline void lane_event(Lane& ln, const IndexView& v, const SearchOut& o) {
  if (ln.state == LS_EV_SCAN || ln.state == LS_EV_WIDE || ln.state == LS_EV_TSCAN) lane_event_scan(ln, v, o);
  else if (ln.state == LS_EV_POP) lane_event_pop(ln, o);
  else lane_event_top(ln, v, o);
}

// ------------------------------------------------------------------------------------------------
// Seed pass. For one strand: look up the seeding k-mer (quasimap.cpp:178,235-241); each of its seed states is
// SPLIT into its suffixes — in the index itself (seed view, KmerSeed) for states of up to kSplitWidth suffixes,
// after narrowing by rank steps for wider ones: every occurrence becomes a width-1 state of its own — a
// candidate — that the text kernel walks through the PRG text. Splitting is exact: a SearchState's interval is
// a set of suffixes that the reference advances in lock-step (one LF step per suffix, one jump per
// marker-preceded suffix, vBWT_jump.cpp:94-117); walking them one by one visits the same (suffix, path) pairs.
// Only the grouping of the final states can differ — two suffixes of one state that both survive to the end
// of the read stay ONE state in the reference — so a strand with more than one finished candidate is handed
// to the general search kernel, which redoes it with interval states. False candidates (a 10-mer has several
// occurrences, one of them real) are rejected against their left context here, or within a step or two of
// the text walk (verify pass).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kPreSteps = 8;     // rank steps at most, while the interval is wide
constexpr uint32_t kNarrowWidth = 4;  // stop narrowing at this many suffixes
constexpr uint32_t kMaxSplit = 32;    // wider than this after narrowing: general kernel

// part 1: k-mer lookup. The strand's seed entries are two runs of the seed view: bucket 0 of its k-mer (wide states,
// suffixes with a marker right to their left) and the bucket of the strand's own next d bases (KmerSeed); entry t
// of the strand is s0 + t for t < n0, s1 + (t - n0) beyond. Returns n0 + n1 (0: the strand is already classified).
struct SeedLookup {
  bool classified;  // the strand's status is settled (skipped / too short / k-mer not indexed): nothing to seed
  uint32_t s0, n0, s1, n1;
  uint32_t pos0;  // bases left of the seeding k-mer (L - k)
  uint32_t ctx;   // the next 12 of them, first one in bits 23:22 (valid where pos0 allows)
  GQ_DEV inline uint32_t entry(uint32_t t) const { return t < n0 ? s0 + t : s1 + (t - n0); }
};

GQ_DEV inline uint32_t preseed_lookup(const IndexView& v, const BatchView& b, const SearchOut& o, uint32_t strand,
                                      SeedLookup& out) {
  const uint32_t r = strand >> 1;
  const uint32_t L = GQ_AT(b.len, r);
  const uint32_t k = v.k;
  out.s0 = out.n0 = out.s1 = out.n1 = out.pos0 = out.ctx = 0;
  out.classified = true;
  if (L == 0) {  // non-ACGT read, emptied by the encoder: skipped (quasimap.cpp:108-113)
    GQ_AT(o.status, strand) = ST_SKIPPED;
    return 0;
  }
  if (L < k) {  // cannot be seeded (UB in the reference, quasimap.cpp:206-210): counted as missing k-mer
    GQ_AT(o.status, strand) = ST_MISSING_KMER;
    return 0;
  }
  const uint32_t* w = b.strand_words(strand, GQ_AT(b.word_off, r));
  const uint32_t code = seeding_kmer_code(w, L, k);
  out.pos0 = L - k;
  if (out.pos0) {
    ReadCursor rd{w, L, 0, 0, 0};
    rd.seek(out.pos0);
    out.ctx = rd.top32() >> 8;
  }
  const uint32_t d = seed_bucket_bases(k), B = seed_buckets(k);
  const uint32_t* off = v.seed_off + (size_t)code * B;
  const uint32_t run_b = GQ_LDG(off), run_e = GQ_LDG(off + B);
  if (run_b == run_e) {  // the seeding k-mer itself is not indexed: the k-mer filter fails
    GQ_AT(o.status, strand) = ST_MISSING_KMER;
    return 0;
  }
  out.classified = false;  // seeded: possibly with no entry at all in the strand's buckets (then nothing maps)
  if (out.pos0 < d) {  // fewer bases left than the buckets are keyed on: every entry of the k-mer
    out.s0 = run_b;
    out.n0 = run_e - run_b;
    return out.n0;
  }
  const uint32_t q = 1u + (d ? out.ctx >> (24 - 2 * d) : 0u);
  out.s0 = run_b;
  out.n0 = GQ_LDG(off + 1) - run_b;
  out.s1 = GQ_LDG(off + q);
  out.n1 = GQ_LDG(off + q + 1) - out.s1;
  return out.n0 + out.n1;
}

// the strand goes to the general search kernel (once: whoever sets the flag appends it)
GQ_DEV inline void send_to_general(const SeedOut& pre, uint32_t strand) {
#if defined(__CUDA_ARCH__)
  const uint32_t old = atomicOr(pre.surv_cnt + strand, kSurvGeneral);
#else
  const uint32_t old = GQ_AT(pre.surv_cnt, strand);
  GQ_AT(pre.surv_cnt, strand) |= kSurvGeneral;
#endif
  if (!(old & kSurvGeneral)) GQ_AT(pre.gen_list, gq_atomic_add(pre.n_gen, 1u)) = strand;
}

// part 2: the candidates of ONE seed state, as a plan of entries {first SA index, number of suffixes,
// pos | kind << 28}. (Only states wider than kSplitWidth get here: the suffixes of the others are entries of
// the index.) A seed state narrower than kNarrowWidth suffixes is one entry as it stands. A wider one is
// narrowed first, the way the reference advances it (quasimap.cpp:258-268): every marker-preceded suffix of
// the interval becomes an entry of its own (its walk starts with that jump — the state left_markers_search
// would spawn, vBWT_jump.cpp:94-117), then the interval consumes the next read base with two rank queries
// (BWT_search.cpp:45-76), until fewer than kNarrowWidth suffixes are left. Returns the number of candidates,
// or kNoAllele when the strand needs the general kernel.
constexpr uint32_t kMaxPlan = 24;
struct SeedPlan {
  uint32_t n;
  uint32_t lo[kMaxPlan], w[kMaxPlan], w0[kMaxPlan];
};

template <class SuperPtr>
GQ_DEV inline uint32_t seed_state_plan(const IndexView& v, SuperPtr super_c, const uint32_t* w, uint32_t L,
                                       uint32_t lo, uint32_t hi, SeedPlan& plan) {
  plan.n = 0;
  const uint32_t pos0 = L - v.k;
  uint32_t w0 = pos0 | (K_SCAN << 28), total = 0;
  if (hi - lo >= kNarrowWidth) {
    Lane ln;
    ln.rd = ReadCursor{w, L, 0, 0, 0};
    ln.pos = pos0;
    ln.lo = lo;
    ln.hi = hi;
    ln.p = ln.mr = 0;
    ln.kind = K_SCAN;
    ln.rd.seek(ln.pos);
    ln.state = LS_RUNW;
    for (uint32_t s = 0; s < kPreSteps && ln.state == LS_RUNW && ln.hi - ln.lo >= kNarrowWidth && ln.pos > 1; ++s) {
      if (ln.kind == K_SCAN) {  // marker-preceded suffixes leave the interval as candidates of their own
        for (uint32_t blk = ln.lo >> kBlkShift; blk <= (ln.hi >> kBlkShift); ++blk) {
          uint64_t m = marker_bits_in(load_blk(v.rank_blk + blk), blk << kBlkShift, ln.lo, ln.hi);
          while (m) {
#if defined(__CUDA_ARCH__)
            const uint32_t bit = __ffsll((long long)m) - 1;
#else
            const uint32_t bit = (uint32_t)__builtin_ctzll(m);
#endif
            m &= m - 1;
            if (plan.n + 1 >= kMaxPlan) {  // keep one entry for the interval itself
              GQ_COUNT(1);
              return kNoAllele;
            }
            plan.lo[plan.n] = (blk << kBlkShift) + bit;
            plan.w[plan.n] = 1;
            plan.w0[plan.n] = ln.pos | (K_SCAN << 28);
            ++plan.n;
            ++total;
          }
        }
        ln.kind = K_READY;  // scanned
      }
      lane_step_wide(ln, v, super_c);
    }
    if (ln.state != LS_EV_POP && (ln.state == LS_EV_WIDE || ln.hi - ln.lo >= kMaxSplit)) {
      GQ_COUNT(1);
      return kNoAllele;
    }
    // LS_RUN / LS_RUNW / LS_EV_SCAN (a marker inside the narrow interval: every suffix checks its own symbol)
    lo = ln.lo;
    hi = ln.state == LS_EV_POP ? lo - 1 : ln.hi;
    w0 = ln.pos | (ln.kind << 28);
  }
  if (hi + 1 != lo) {
    plan.lo[plan.n] = lo;
    plan.w[plan.n] = hi + 1 - lo;
    plan.w0[plan.n] = w0;
    ++plan.n;
    total += hi + 1 - lo;
  }
  GQ_COUNT(2);
  return total;
}

// part 3: the suffixes of the plan become candidates {text position, pos | kind << 28} — unless their very first
// text step already fails (wrong base left of the suffix: most of the other occurrences of the seeding k-mer
// end here, before they cost a record). Returns the number kept, or kNoAllele when they do not fit.
constexpr uint32_t kMaxCand = 24;
struct SeedCands {
  uint32_t p[kMaxCand], w0[kMaxCand];
};

GQ_DEV inline uint32_t seed_filter(const IndexView& v, const SeedPlan& plan, const uint32_t* w, uint32_t L,
                                   SeedCands& out) {
  uint32_t n = 0, cur_pos = 0;
  ReadCursor rd{w, L, 0, 0, 0};
  for (uint32_t e = 0; e < plan.n; ++e) {
    const uint32_t w0 = plan.w0[e], pos = w0 & 0x0FFFFFFFu;
    if (pos != cur_pos) {  // entries of one seed state share a few read positions
      rd.seek(pos);
      cur_pos = pos;
    }
    for (uint32_t i = 0; i < plan.w[e]; ++i) {
      const uint32_t p = GQ_LDG(v.sa + plan.lo[e] + i);
      Lane ln;
      ln.rd = rd;
      ln.pos = pos;
      ln.kind = w0 >> 28;
      ln.p = p;
      ln.lo = ln.hi = ln.mr = 0;
      ln.state = LS_TEXT;
      lane_text_step(ln, v);
      if (ln.state == LS_EV_POP) continue;
      if (n == kMaxCand) return kNoAllele;
      out.p[n] = p;
      out.w0[n] = w0;
      ++n;
    }
  }
  return n;
}

// parts 2 + 3 for seed entry j. The common case — a suffix of a narrow state — is decided from its 8-byte
// entry alone: text position + left context (KmerSeed), compared with the next read bases in registers.
template <class SuperPtr>
GQ_DEV inline uint32_t seed_state_cands(const IndexView& v, SuperPtr super_c, const uint32_t* w, uint32_t L,
                                        uint32_t j, SeedCands& out) {
  const uint32_t pos0 = L - v.k;
  if (pos0 == 0) {  // the seed states are the final states
    GQ_COUNT(0);
    return kNoAllele;
  }
#if defined(__CUDA_ARCH__)
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(v.seed_ent) + j);
  const uint32_t key = raw.x, aux = raw.y;
#else
  const uint32_t key = GQ_AT(v.seed_ent, j).key, aux = GQ_AT(v.seed_ent, j).aux;
#endif
  if (aux & 0x80000000u) {
    uint32_t m = (aux >> 24) & 0x7Fu;  // context bases, up to the first marker / the text start
    m = m < pos0 ? m : pos0;
    if (m) {
      ReadCursor rd{w, L, 0, 0, 0};
      rd.seek(pos0);
      const uint32_t x = ((rd.top32() >> 8) ^ aux) & ((0xFFFFFFFFu << (24 - 2 * m)) & 0xFFFFFFu);
      if (x) return 0;  // another occurrence of the k-mer: the read continues differently
    }
    GQ_COUNT(2);
    out.p[0] = key;
    out.w0[0] = pos0 | (K_SCAN << 28);
    return 1;
  }
  SeedPlan plan;
  const uint32_t cnt = seed_state_plan(v, super_c, w, L, key, aux, plan);
  return cnt == kNoAllele ? cnt : seed_filter(v, plan, w, L, out);
}

// candidate record: 4 words {strand, k-mer state index, text position, pos | kind << 28}
GQ_DEV inline void seed_write(const SeedCands& c, uint32_t n, const SeedOut& pre, uint32_t strand, uint32_t j,
                              uint32_t base) {
  uint32_t* d = pre.rec + 4 * (size_t)base;
  GQ_TOUCH(d, 16 * n);
  for (uint32_t i = 0; i < n; ++i, d += 4) {
    d[0] = strand;
    d[1] = j;
    d[2] = c.p[i];
    d[3] = c.w0[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Text kernel: every candidate is followed by one thread, entirely in text mode: registers + a short local
// path, no stack, no arena. Jumps are taken from the pre-resolved text-order records (exit of a site,
// whole-SNP crossing, plain entry); anything else — a jump that needs the general machinery, an interval
// that widens, a path longer than the local buffers — sends the strand to the general search kernel, which
// redoes it from the k-mer index (same results: both follow quasimap.cpp:227-268 / vBWT_jump.cpp state by
// state). The path of the seed state itself (k-mer index) is only fetched when the candidate finishes.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kFastT = 24;  // (site, allele) pairs added during the walk
constexpr uint32_t kFastG = 4;   // sites opened during the walk
enum FastResult : uint32_t { FAST_NONE = 0, FAST_DEAD = 1, FAST_MAPPED = 2, FAST_BAIL = 3 };

struct FastLane {
  Lane ln;
  uint32_t nt, ng;       // local path: pairs in T, open sites in G
  uint32_t nt0, ng0;     // what is left of the seed state's own path (its open sites can be closed by exits)
  uint32_t path_off;     // ... in kmer_paths
  uint32_t T[2 * kFastT], G[kFastG];
  uint32_t result;  // FAST_NONE while running
  bool p_valid;     // ln.p is the text position of the current suffix
};

GQ_DEV inline bool fast_running(const FastLane& f) {
  return f.result == FAST_NONE && (f.ln.state == LS_TEXT || f.ln.state == LS_EV_TSCAN);
}

// candidate record -> lane (TRACK: the seed state's path counts are needed, from the k-mer index)
template <bool TRACK>
GQ_DEV inline void fast_begin(FastLane& f, const IndexView& v, const BatchView& b, const SeedOut& pre, uint32_t idx) {
#if defined(__CUDA_ARCH__)
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(pre.rec) + idx);
  const uint32_t strand = c.x, j = c.y, p = c.z, w0 = c.w;
#else
  const uint32_t* c = pre.rec + 4 * (size_t)idx;
  GQ_TOUCH(c, 16);
  const uint32_t strand = c[0], j = c[1], p = c[2], w0 = c[3];
#endif
  const uint32_t r = strand >> 1;
  const uint32_t L = GQ_AT(b.len, r), woff = GQ_AT(b.word_off, r);
  f.ln.strand = strand;
  f.result = FAST_NONE;
  f.p_valid = true;
  f.nt = f.ng = 0;
  f.nt0 = f.ng0 = f.path_off = 0;
  if (TRACK) {
    const KmerState ks = GQ_AT(v.kmer_states, j);
    f.nt0 = ks.counts & 0xFFFFu;
    f.ng0 = ks.counts >> 16;
    f.path_off = ks.path_off;
  }
  f.ln.rd = ReadCursor{b.strand_words(strand, woff), L, 0, 0, 0};
  f.ln.pos = w0 & 0x0FFFFFFFu;
  f.ln.kind = w0 >> 28;
  f.ln.p = p;
  f.ln.lo = f.ln.hi = 0;
  f.ln.mr = 0;
  f.ln.rd.seek(f.ln.pos);  // pos >= 1: seed states of reads with L == k never become candidates
  f.ln.state = LS_TEXT;
}

// LS_EV_TSCAN: the marker left of the suffix (lane_event_scan's pre-resolved cases, on the local path).
// TRACK = false (verify pass): same walk, path not recorded.
template <bool TRACK>
GQ_DEV inline void fast_event(FastLane& f, const IndexView& v) {
  Lane& ln = f.ln;
  const uint32_t* jr = v.tmarker_hit + 8 * (size_t)ln.mr;
#if defined(__CUDA_ARCH__)
  const uint4 ja = __ldg(reinterpret_cast<const uint4*>(jr)), jb = __ldg(reinterpret_cast<const uint4*>(jr) + 1);
  const uint32_t marker = ja.x, allele = ja.y, jlo = ja.z, jhi = ja.w, snp = jb.x, p_jump = jb.z, p_site = jb.w;
#else
  GQ_TOUCH(jr, 32);
  const uint32_t marker = jr[0], allele = jr[1], jlo = jr[2], jhi = jr[3], snp = jr[4], p_jump = jr[6], p_site = jr[7];
#endif
  if (marker == 0) {
    f.result = FAST_DEAD;
    return;
  }
  if (jlo == kNoAllele) {  // adjacent markers: general jump machinery
    GQ_COUNT(5);
    f.result = FAST_BAIL;
    return;
  }
  if (marker & 1u) {  // leave the site through `allele` (exit_site_in_place)
    if (jlo != jhi || (TRACK && f.nt == kFastT)) {
      f.result = FAST_BAIL;
      return;
    }
    if (TRACK) {
      if (f.ng > 0) --f.ng;  // the site being left is the innermost open one
      else if (f.ng0 > 0) --f.ng0;
      f.T[2 * f.nt] = marker;
      f.T[2 * f.nt + 1] = allele;
      ++f.nt;
    }
    ln.p = p_jump;
    ln.kind = K_READY;
    ln.state = LS_TEXT;
    return;
  }
  // enter the site from its right end and consume the next base
  const uint32_t c = ln.rd.peek();
  if (snp != kNotSnp && ln.pos >= 2) {  // site of distinct single-base alleles: entry, base, exit
    const uint32_t a = (snp >> (8 * c)) & 0xFFu;
    if (a == 0xFFu) {
      f.result = FAST_DEAD;
      return;
    }
    if (TRACK) {
      if (f.nt == kFastT) {
        f.result = FAST_BAIL;
        return;
      }
      f.T[2 * f.nt] = marker - 1;
      f.T[2 * f.nt + 1] = a;
      ++f.nt;
    }
    ln.p = p_site;
    ln.pos -= 1;
    ln.rd.advance();
    ln.kind = K_READY;
    ln.state = LS_TEXT;
    return;
  }
  if (TRACK) {
    if (f.ng == kFastG) {
      f.result = FAST_BAIL;
      return;
    }
    f.G[f.ng++] = marker - 1;
  }
  const uint32_t slot = (marker - 6) >> 1;
  const uint32_t nlo = GQ_LDG(v.entry_next + 8 * slot + 2 * c), nhi = GQ_LDG(v.entry_next + 8 * slot + 2 * c + 1);
  if (nhi + 1 <= nlo) {  // no allele ends in c
    f.result = FAST_DEAD;
    return;
  }
  if (nlo != nhi) {  // several alleles end in c
    GQ_COUNT(6);
    f.result = FAST_BAIL;
    return;
  }
  ln.rd.advance();
  ln.kind = K_SCAN;
  if (--ln.pos == 0) {
    ln.lo = ln.hi = nlo;
    f.p_valid = false;
    ln.state = LS_EV_TOP;
  } else {
    ln.p = GQ_LDG(v.sa + nlo);
    ln.state = LS_TEXT;
  }
}

// ------------------------------------------------------------------------------------------------
// Branching walks (nested PRGs). A jump that is not pre-resolved — a marker with another marker on its far side, a
// site whose alleles can be empty or end in a nested site, several alleles ending in the read's next base — makes
// the reference push several SearchStates (search_state_vBWT_jumps, vBWT_jump.cpp:134-265). The text walker follows
// them depth-first: the path so far stays where it is (the branches share it as a prefix), and every alternative
// becomes a small FORK record — a text position to continue from, or a pending (marker, allele) jump — on a short
// private stack. What comes out is the same set of (suffix, path) pairs the reference's states hold (DESIGN §3,
// observation 2); as before, a strand is only finished here when exactly ONE branch of ONE candidate reaches the
// end of the read — a second finisher (two suffixes of one reference state, or two states) hands the strand to the
// general kernel, which keeps intervals together.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kFastForks = 10;
constexpr uint32_t kForkJump = 0xFu;  // `kind` field of a fork that is a pending jump

struct FastForks {
  uint32_t n;
  uint32_t w0[kFastForks];   // pos | kind << 28
  uint32_t a[kFastForks];    // text position to continue from / marker of the pending jump
  uint32_t b[kFastForks];    // allele of the pending jump
  uint32_t cnt[kFastForks];  // nt | ng << 8 | ng0 << 16 at the fork
  uint32_t G[kFastForks][kFastG];
  bool active;               // the candidate still has a branch walking or waiting
  bool have_fin;             // one branch has reached the end of the read: kept here while the others are tried
  bool fin_p_valid;
  uint32_t fin_lo, fin_p, fin_cnt;
  uint32_t fin_from;  // first path pair kept in fin_T: the pairs below it are shared with every pending fork and stay in f.T
  uint32_t fin_T[2 * kFastT], fin_G[kFastG];
};

GQ_DEV inline void forks_init(FastForks& fk, bool active) {
  fk.n = 0;
  fk.active = active;
  fk.have_fin = false;
}

GQ_DEV inline bool fork_push(FastLane& f, FastForks& fk, uint32_t w0, uint32_t a, uint32_t b) {
  if (fk.n == kFastForks) {
    f.result = FAST_BAIL;
    return false;
  }
  const uint32_t i = fk.n++;
  fk.w0[i] = w0;
  fk.a[i] = a;
  fk.b[i] = b;
  fk.cnt[i] = f.nt | (f.ng << 8) | (f.ng0 << 16);
  for (uint32_t j = 0; j < kFastG; ++j) fk.G[i][j] = f.G[j];
  return true;
}

// Apply the jump (marker, allele) to the walking branch: search_state_vBWT_jumps' worklist for ONE locus, with
// extend_targets_site_exit (:185-228) and extend_targets_site_entry (:230-265), on the local path.
GQ_DEV inline void fast_jump(FastLane& f, FastForks& fk, const IndexView& v, uint32_t marker, uint32_t allele) {
  Lane& ln = f.ln;
  while (true) {
    if (marker & 1u) {  // leave site `marker` through `allele`; then whatever sits directly left of the site
      if (f.nt == kFastT) {
        f.result = FAST_BAIL;
        return;
      }
      if (f.ng > 0) --f.ng;  // the site being left is the innermost open one
      else if (f.ng0 > 0) --f.ng0;
      f.T[2 * f.nt] = marker;
      f.T[2 * f.nt + 1] = allele;
      ++f.nt;
      const uint32_t slot = (marker - 5) >> 1;
      const uint32_t nxt = GQ_LDG(v.tm_odd + slot);
      if (nxt == 0) {  // plain exit: the suffix that starts with the site-entry marker
        ln.p = GQ_LDG(v.site_rec + 4 * (size_t)slot + 2) - 1;
        ln.kind = K_READY;
        ln.state = LS_TEXT;
        return;
      }
      if (nxt & 1u) {  // double exit: the parent's allele from par_map
        allele = GQ_LDG(v.par + 2 * slot + 1);
        marker = nxt;
      } else {  // exit followed by an entry
        marker = nxt;
        allele = 0;
      }
      continue;
    }
    // enter site marker - 1 from its right end
    const uint32_t site = marker - 1, slot = (site - 5) >> 1;
    if (f.ng == kFastG) {
      f.result = FAST_BAIL;
      return;
    }
    f.G[f.ng++] = site;
    // direct deletions and double entries hang off the entered state: pending jumps, no base consumed yet
    const uint32_t tb = GQ_LDG(v.tm_even_off + slot), te = GQ_LDG(v.tm_even_off + slot + 1);
    for (uint32_t j = tb; j < te; ++j) {
      const uint32_t id = GQ_LDG(v.tm_even + 2 * j), del = GQ_LDG(v.tm_even + 2 * j + 1);
      if (!fork_push(f, fk, ln.pos | (kForkJump << 28), id, (id & 1u) ? del : kNoAllele)) return;
    }
    // the alleles that end in the read's next base: one suffix each
    const uint32_t c = ln.rd.peek();
    const uint32_t nlo = GQ_LDG(v.entry_next + 8 * slot + 2 * c), nhi = GQ_LDG(v.entry_next + 8 * slot + 2 * c + 1);
    if (nhi + 1 <= nlo) {  // none: this branch ends here (its forks live on)
      ln.state = LS_EV_POP;
      return;
    }
    if (ln.pos == 1 && nhi != nlo) {  // the read ends on that base: several suffixes of one state finish together
      GQ_COUNT(6);
      f.result = FAST_BAIL;
      return;
    }
    for (uint32_t i = nlo + 1; i <= nhi; ++i)
      if (!fork_push(f, fk, (ln.pos - 1) | (K_SCAN << 28), GQ_LDG(v.sa + i), 0)) return;
    ln.rd.advance();
    ln.kind = K_SCAN;
    if (--ln.pos == 0) {
      ln.lo = ln.hi = nlo;
      f.p_valid = false;
      ln.state = LS_EV_TOP;
    } else {
      ln.p = GQ_LDG(v.sa + nlo);
      ln.state = LS_TEXT;
    }
    return;
  }
}

// LS_EV_TSCAN of a branching walk: the pre-resolved shortcuts of fast_event (plain exit, whole-SNP crossing), the
// jump machinery for everything else
GQ_DEV inline void fast_event_dfs(FastLane& f, FastForks& fk, const IndexView& v) {
  Lane& ln = f.ln;
  const uint32_t* jr = v.tmarker_hit + 8 * (size_t)ln.mr;
#if defined(__CUDA_ARCH__)
  const uint4 ja = __ldg(reinterpret_cast<const uint4*>(jr)), jb = __ldg(reinterpret_cast<const uint4*>(jr) + 1);
  const uint32_t marker = ja.x, allele = ja.y, jlo = ja.z, snp = jb.x, p_jump = jb.z, p_site = jb.w;
#else
  GQ_TOUCH(jr, 32);
  const uint32_t marker = jr[0], allele = jr[1], jlo = jr[2], snp = jr[4], p_jump = jr[6], p_site = jr[7];
#endif
  if (marker == 0) {
    f.result = FAST_DEAD;
    return;
  }
  if (jlo != kNoAllele && (marker & 1u)) {  // exit with nothing adjacent
    if (f.nt == kFastT) {
      f.result = FAST_BAIL;
      return;
    }
    if (f.ng > 0) --f.ng;
    else if (f.ng0 > 0) --f.ng0;
    f.T[2 * f.nt] = marker;
    f.T[2 * f.nt + 1] = allele;
    ++f.nt;
    ln.p = p_jump;
    ln.kind = K_READY;
    ln.state = LS_TEXT;
    return;
  }
  if (jlo != kNoAllele && snp != kNotSnp && ln.pos >= 2) {  // site of distinct single-base alleles: entry, base, exit
    const uint32_t a = (snp >> (8 * ln.rd.peek())) & 0xFFu;
    if (a == 0xFFu) {
      f.result = FAST_DEAD;
      return;
    }
    if (f.nt == kFastT) {
      f.result = FAST_BAIL;
      return;
    }
    f.T[2 * f.nt] = marker - 1;
    f.T[2 * f.nt + 1] = a;
    ++f.nt;
    ln.p = p_site;
    ln.pos -= 1;
    ln.rd.advance();
    ln.kind = K_READY;
    ln.state = LS_TEXT;
    return;
  }
  if (jlo == kNoAllele) GQ_COUNT(5);
  fast_jump(f, fk, v, marker, allele);
}

// keep the finished branch aside while the pending forks are tried: only the path pairs a fork can overwrite (those
// from the smallest fork point on) need a copy
GQ_DEV inline void fast_keep_finisher(const FastLane& f, FastForks& fk) {
  fk.have_fin = true;
  fk.fin_p_valid = f.p_valid;
  fk.fin_lo = f.ln.lo;
  fk.fin_p = f.ln.p;
  fk.fin_cnt = f.nt | (f.ng << 8) | (f.ng0 << 16);
  uint32_t from = f.nt;
  for (uint32_t i = 0; i < fk.n; ++i) {
    const uint32_t nt_i = fk.cnt[i] & 0xFFu;
    from = nt_i < from ? nt_i : from;
  }
  fk.fin_from = from;
  for (uint32_t j = 2 * from; j < 2 * f.nt; ++j) fk.fin_T[j] = f.T[j];
  for (uint32_t j = 0; j < kFastG; ++j) fk.fin_G[j] = f.G[j];
}

// The walking branch has ended (dead, finished, or it needs the general kernel): remember a finisher, take the
// next fork, or close the candidate. Afterwards either a branch is walking again, or fk.active is false and the lane
// holds the candidate's outcome for fast_outcome (state LS_EV_TOP = its single finished branch).
GQ_DEV inline void fast_branch_end(FastLane& f, FastForks& fk, const IndexView& v) {
  Lane& ln = f.ln;
  if (f.result == FAST_BAIL) {
    fk.active = false;
    return;
  }
  if (f.result == FAST_NONE && ln.state == LS_EV_TOP) {  // this branch consumed the whole read
    if (fk.have_fin) {  // a second one: two suffixes / states finish — the general kernel redoes the strand
      GQ_COUNT(4);
      f.result = FAST_BAIL;
      fk.active = false;
      return;
    }
    if (fk.n == 0) {  // nothing else to try (the usual case): the lane already holds the candidate's outcome
      fk.active = false;
      return;
    }
    fast_keep_finisher(f, fk);
  }
  f.result = FAST_NONE;
  while (true) {
    if (fk.n == 0) {  // nothing left to try: the candidate's outcome
      fk.active = false;
      if (fk.have_fin) {
        f.nt = fk.fin_cnt & 0xFFu;
        f.ng = (fk.fin_cnt >> 8) & 0xFFu;
        f.ng0 = fk.fin_cnt >> 16;
        for (uint32_t j = 2 * fk.fin_from; j < 2 * f.nt; ++j) f.T[j] = fk.fin_T[j];
        for (uint32_t j = 0; j < kFastG; ++j) f.G[j] = fk.fin_G[j];
        f.p_valid = fk.fin_p_valid;
        ln.p = fk.fin_p;
        ln.lo = ln.hi = fk.fin_lo;
        ln.pos = 0;
        ln.state = LS_EV_TOP;
      } else
        f.result = FAST_DEAD;
      return;
    }
    const uint32_t i = --fk.n;
    f.nt = fk.cnt[i] & 0xFFu;
    f.ng = (fk.cnt[i] >> 8) & 0xFFu;
    f.ng0 = fk.cnt[i] >> 16;
    for (uint32_t j = 0; j < kFastG; ++j) f.G[j] = fk.G[i][j];
    ln.pos = fk.w0[i] & 0x0FFFFFFFu;
    ln.rd.seek(ln.pos);  // pos >= 1: forks are only made while bases are left
    f.p_valid = true;
    if ((fk.w0[i] >> 28) == kForkJump) {
      fast_jump(f, fk, v, fk.a[i], fk.b[i]);
      if (f.result == FAST_BAIL) {
        fk.active = false;
        return;
      }
      if (f.result == FAST_NONE && (ln.state == LS_TEXT || ln.state == LS_EV_TSCAN)) return;  // walking again
      if (f.result == FAST_NONE && ln.state == LS_EV_TOP) {  // finished inside the jump (read ends on the entered base)
        if (fk.have_fin) {
          GQ_COUNT(4);
          f.result = FAST_BAIL;
          fk.active = false;
          return;
        }
        fast_keep_finisher(f, fk);
      }
      f.result = FAST_NONE;  // dead or recorded: next fork
      continue;
    }
    ln.p = fk.a[i];
    ln.kind = fk.w0[i] >> 28;
    ln.state = LS_TEXT;
    return;
  }
}

// Verify pass: does the candidate agree with the PRG on kVerifyBases further bases (or to the end of the read,
// or up to something only the full walk can decide)? False candidates — the other occurrences of the
// seeding k-mer — end here, so the full walk only sees about one candidate per mappable strand.
constexpr uint32_t kVerifyBases = 8;
constexpr uint32_t kVerifyIters = 6;
GQ_DEV inline bool fast_verified(const FastLane& f, uint32_t pos0) {
  return !fast_running(f) || pos0 - f.ln.pos >= kVerifyBases;
}
GQ_DEV inline bool fast_alive(const FastLane& f) {
  return !(f.result == FAST_DEAD || (f.result == FAST_NONE && f.ln.state == LS_EV_POP));
}

// after the walk: classify the outcome; for a finished state, settle its record (encapsulated_search.cpp:30-88
// for a path-less one) and return the pool words it needs (0 otherwise)
GQ_DEV inline uint32_t fast_outcome(FastLane& f, const IndexView& v) {
  if (f.result != FAST_NONE) return 0;
  if (f.ln.state != LS_EV_TOP) {
    f.result = FAST_DEAD;
    return 0;
  }
  f.result = FAST_MAPPED;
  GQ_COUNT(7);  // candidates finished by the fast path
  if (!(f.nt | f.ng | f.nt0 | f.ng0)) {
    const uint32_t pf = f.p_valid ? f.ln.p : GQ_LDG(v.sa + f.ln.lo);
    const Node& nd = GQ_AT(v.nodes, GQ_LDG(v.pos2node + pf));
    if (nd.site != 0) {
      f.T[0] = nd.site;
      f.T[1] = (uint32_t)nd.allele;
      f.nt = 1;
    }
  }
  return 4 + 2 * (f.nt0 + f.nt) + 2 * (f.ng0 + f.ng);
}

// write the single final state of the strand at pool offset `off`: the seed state's path, then the local one
GQ_DEV inline void fast_emit(const FastLane& f, const IndexView& v, const SearchOut& o, uint32_t off) {
  uint32_t* d = o.pool + off;
  const uint32_t nt = f.nt0 + f.nt, ng = f.ng0 + f.ng;
  GQ_TOUCH(d, 4 * (4 + 2 * nt + 2 * ng));
  d[0] = f.ln.lo;
  d[1] = f.ln.hi;
  d[2] = nt;
  d[3] = ng;
  d += 4;
  for (uint32_t j = 0; j < 2 * f.nt0; ++j) *d++ = GQ_LDG(v.kmer_paths + f.path_off + j);
  for (uint32_t j = 0; j < 2 * f.nt; ++j) *d++ = f.T[j];
  for (uint32_t j = 0; j < f.ng0; ++j) {
    *d++ = GQ_LDG(v.kmer_paths + f.path_off + 2 * f.nt0 + j);
    *d++ = kNoAllele;
  }
  for (uint32_t j = 0; j < f.ng; ++j) {
    *d++ = f.G[j];
    *d++ = kNoAllele;
  }
  const uint32_t strand = f.ln.strand;
  GQ_AT(o.st_off, strand) = off;
  GQ_AT(o.st_words, strand) = 4 + 2 * nt + 2 * ng;
  GQ_AT(o.st_count, strand) = 1;
}

// a finished candidate claims its strand: 0 = first one (emit), otherwise the strand is (or becomes) the
// general kernel's
GQ_DEV inline bool fast_claim(const SeedOut& pre, uint32_t strand) {
  const uint32_t old = gq_atomic_add(pre.surv_cnt + strand, 1u);
  if (old & kSurvGeneral) return false;
  if (old & 0xFFFFu) {
    GQ_COUNT(4);  // several finished candidates
    send_to_general(pre, strand);
    return false;
  }
  return true;
}

// all_read_kmers_occur_in_index (quasimap.cpp:212-225) for a strand whose search found nothing:
// any k-mer absent from the index -> missing_kmer, else no_extension (quasimap.cpp:170-186).
// The k-mer code convention (base j at bits [2j,2j+2)) makes the code of the window starting at base i
// a plain bit-field of the 2-bit packed read, so the loop slides a 64-bit window by one base per
// iteration ("any k-mer missing" does not depend on the order in which windows are visited).
GQ_DEV inline void classify_strand(const IndexView& v, const BatchView& b, const SearchOut& o, uint32_t strand) {
  const uint32_t r = strand >> 1;
  const uint32_t L = GQ_AT(b.len, r), k = v.k;
  const uint32_t* w = b.strand_words(strand, GQ_AT(b.word_off, r));
  const uint32_t mask = (k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
  const uint32_t n_words = (L + 15) >> 4;
  uint64_t win = GQ_LDG(w);
  if (n_words > 1) win |= (uint64_t)GQ_LDG(w + 1) << 32;
  bool missing = false;
  for (uint32_t i = 0; i + k <= L; ++i) {
    const uint32_t code = (uint32_t)win & mask;
    if (!((GQ_LDG(v.kmer_bits + (code >> 5)) >> (code & 31u)) & 1u)) {
      missing = true;
      break;
    }
    win >>= 2;
    if ((i & 15u) == 15u) {  // 16 bases consumed: bring in the next packed word
      uint32_t nw = (i >> 4) + 2;
      if (nw < n_words) win |= (uint64_t)GQ_LDG(w + nw) << 32;
    }
  }
  GQ_AT(o.status, strand) = missing ? ST_MISSING_KMER : ST_NO_EXTENSION;
}

// Single-lane driver (host emulation and a reference for the kernels): seed pass, text walk of every
// candidate, the general lane state machine (seeded from the k-mer index) where the fast path does not
// apply; then the k-mer filter if nothing mapped.
GQ_DEV inline void map_strand(const IndexView& v, const uint32_t* super_cnt, const BatchView& b, const SearchOut& o,
                              const SeedOut& pre, uint32_t strand, uint32_t* arena, uint32_t arena_words) {
  bool general = GQ_AT(o.status, strand) == ST_OVERFLOW;  // overflow re-runs always use the general machinery
  if (!general) {
    GQ_PHASE(0);  // seed_kernel
    SeedLookup lk;
    const uint32_t ns = preseed_lookup(v, b, o, strand, lk);
    if (lk.classified) return;  // skipped / too short / k-mer not indexed
    GQ_AT(o.status, strand) = ST_UNCLASSIFIED;
    GQ_AT(pre.surv_cnt, strand) = 0;
    // the emulation reuses the candidate pool strand by strand
    uint32_t total = 0;
    const uint32_t r = strand >> 1;
    for (uint32_t t = 0; t < ns && !general; ++t) {
      SeedCands cands;
      const uint32_t cnt = seed_state_cands(v, super_cnt, b.strand_words(strand, GQ_AT(b.word_off, r)), GQ_AT(b.len, r), lk.entry(t), cands);
      if (cnt == kNoAllele || total + cnt > pre.cap) general = true;
      else {
        seed_write(cands, cnt, pre, strand, GQ_AT(v.seed_state, lk.entry(t)), total);
        total += cnt;
      }
    }
    if (!general) {
      // verify pass over all candidates of the strand first (verify_kernel), then the full walks of the survivors from
      // the start (text_kernel)
      for (uint32_t i = 0; i < total; ++i) {
        FastLane f;
        GQ_PHASE(1);  // verify_kernel
        fast_begin<false>(f, v, b, pre, i);
        const uint32_t pos0 = f.ln.pos;
        for (uint32_t it = 0; it < kVerifyIters && !fast_verified(f, pos0); ++it) {
          if (f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
          if (fast_running(f) && f.ln.state == LS_EV_TSCAN) fast_event<false>(f, v);
        }
        if (!fast_alive(f)) {
          pre.rec[4 * (size_t)i] = kNoAllele;  // (the kernel compacts the survivors instead)
          continue;
        }
        GQ_TOUCH(pre.rec + 4 * (size_t)i, 16);  // the survivor is copied to the text kernel's list
      }
      for (uint32_t i = 0; i < total; ++i) {
        if (pre.rec[4 * (size_t)i] == kNoAllele) continue;
        FastLane f;
        GQ_PHASE(2);  // text_kernel
        fast_begin<true>(f, v, b, pre, i);
        if (v.any_nested) {  // text_kernel<true>: branching walks
          FastForks fk;
          forks_init(fk, true);
          while (fk.active) {
            if (fast_running(f) && f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
            if (fast_running(f) && f.ln.state == LS_EV_TSCAN) fast_event_dfs(f, fk, v);
            if (!fast_running(f)) fast_branch_end(f, fk, v);
          }
        } else {  // text_kernel<false>
          while (fast_running(f)) {
            if (f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
            if (f.ln.state == LS_EV_TSCAN) fast_event<true>(f, v);
          }
        }
        const uint32_t words = fast_outcome(f, v);
        if (f.result == FAST_MAPPED) {
          if (!fast_claim(pre, strand)) continue;
          const uint32_t off = gq_atomic_add(o.pool_used, words);
          if (off + words > o.pool_cap) {
            GQ_AT(o.status, strand) = ST_OVERFLOW;
            GQ_AT(o.overflow_list, gq_atomic_add(o.n_overflow, 1u)) = strand;
            return;
          }
          fast_emit(f, v, o, off);
          GQ_AT(o.status, strand) = ST_MAPPED;
          list_mapped(o, strand);
        } else if (f.result == FAST_BAIL)
          send_to_general(pre, strand);
      }
      general = (GQ_AT(pre.surv_cnt, strand) & kSurvGeneral) != 0;
    }
  }
  if (general) {
    GQ_COUNT(8);  // strands through the general machinery
    GQ_PHASE(3);  // search_kernel
    Lane ln;
    ln.state = LS_IDLE;
    lane_refill(ln, v, b, o, strand, arena, arena_words);
    while (ln.state != LS_IDLE) {
      if (ln.state == LS_RUN) lane_to_text(ln, v);
      else if (ln.state == LS_TEXT) lane_text_step(ln, v);
      else if (ln.state == LS_RUNW) lane_step_wide(ln, v, super_cnt);
      else lane_event(ln, v, o);
    }
  }
  if (GQ_AT(o.status, strand) == ST_UNCLASSIFIED) {
    GQ_PHASE(4);  // classify_kernel
    classify_strand(v, b, o, strand);
  }
}

struct Scratch {
  uint32_t* mem;
  uint32_t used, cap;
  bool overflow;
  GQ_DEV inline uint32_t* alloc(uint32_t words) {
    if (used + words > cap) {
      overflow = true;
      return nullptr;
    }
    uint32_t* p = mem + used;
    used += words;
    return p;
  }
};

// std::mt19937 seeded with `seed`: j-th raw output (j < 227), from the init LCG alone
// (reference: RandomInclusiveInt, src/common/random.cpp:4-19).
GQ_DEV uint32_t mt19937_output(uint32_t seed, uint32_t j) {
  uint32_t x = seed, mj = 0, mj1 = 0, mj397 = 0;
  for (uint32_t i = 0;; ++i) {
    if (i == j) mj = x;
    if (i == j + 1) mj1 = x;
    if (i == j + 397) {
      mj397 = x;
      break;
    }
    x = 1812433253u * (x ^ (x >> 30)) + (i + 1);
  }
  uint32_t y = (mj & 0x80000000u) | (mj1 & 0x7FFFFFFFu);
  uint32_t z = mj397 ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
  z ^= z >> 11;
  z ^= (z << 7) & 0x9D2C5680u;
  z ^= (z << 15) & 0xEFC60000u;
  z ^= z >> 18;
  return z;
}

// std::uniform_int_distribution<uint32_t>(1, range)(mt19937(seed)) as libstdc++ 13 computes it
// (Lemire's multiply-shift with rejection, bits/uniform_int_dist.h `_S_nd`).
GQ_DEV uint32_t uniform_1_to(uint32_t seed, uint32_t range) {
  uint32_t j = 0;
  uint64_t product = (uint64_t)mt19937_output(seed, j) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (0u - range) % range;
    while (low < threshold && j < 200) {
      ++j;
      product = (uint64_t)mt19937_output(seed, j) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32) + 1u;
}

struct StateRec {
  uint32_t lo, hi, nt, ng;
  const uint32_t* T;  // nt pairs
  const uint32_t* G;  // ng pairs (site, -1)
  GQ_DEV inline uint32_t words() const { return 4 + 2 * nt + 2 * ng; }
};
GQ_DEV inline StateRec parse_rec(const uint32_t* p) {
  StateRec r{p[0], p[1], p[2], p[3], p + 4, p + 4 + 2 * p[2]};
  return r;
}

// LocusFinder (coverage_common.cpp:10-83). Appends to `loci` (pairs, deduplicated) and `base`
// (level-0 sites, deduplicated); `used` is per-state scratch.
struct LocusLists {
  uint32_t *loci, *base, *used;
  uint32_t n_loci, n_base, n_used, cap;
  bool overflow;
};
GQ_DEV void add_locus(LocusLists& l, uint32_t site, uint32_t allele) {
  for (uint32_t i = 0; i < l.n_loci; ++i)
    if (l.loci[2 * i] == site && l.loci[2 * i + 1] == allele) return;
  if (l.n_loci >= l.cap) {
    l.overflow = true;
    return;
  }
  l.loci[2 * l.n_loci] = site;
  l.loci[2 * l.n_loci + 1] = allele;
  ++l.n_loci;
}
GQ_DEV void assign_nested(const IndexView& v, LocusLists& l, uint32_t site, uint32_t allele) {
  while (true) {
    for (uint32_t i = 0; i < l.n_used; ++i)
      if (l.used[i] == site) return;
    if (l.n_used >= l.cap || l.n_base >= l.cap) {
      l.overflow = true;
      return;
    }
    l.used[l.n_used++] = site;
    add_locus(l, site, allele);
    uint32_t slot = (site - 5) >> 1;
    uint32_t ps = GQ_AT(v.par, 2 * slot);
    if (ps == 0) {
      bool seen = false;
      for (uint32_t i = 0; i < l.n_base; ++i) seen |= l.base[i] == site;
      if (!seen) l.base[l.n_base++] = site;
      return;
    }
    allele = GQ_AT(v.par, 2 * slot + 1);
    site = ps;
  }
}
GQ_DEV void locus_finder(const IndexView& v, const StateRec& st, LocusLists& l) {
  l.n_used = 0;
  if (st.ng > 0) {  // assign_traversing_loci :52-74
    uint32_t seed_site = st.G[2 * (st.ng - 1)];
    uint32_t last_al = 0;
    for (uint32_t i = st.lo;; ++i) {
      uint32_t p = GQ_LDG(v.sa + i);
      last_al = (uint32_t)GQ_AT(v.nodes, GQ_LDG(v.pos2node + p)).allele;
      add_locus(l, seed_site, last_al);
      if (i == st.hi) break;
    }
    assign_nested(v, l, seed_site, last_al);
  }
  for (uint32_t j = 0; j < st.nt; ++j) assign_nested(v, l, st.T[2 * j], st.T[2 * j + 1]);  // :76-83
}

GQ_DEV void sort_u32(uint32_t* a, uint32_t n) {
  for (uint32_t i = 1; i < n; ++i) {
    uint32_t x = a[i], j = i;
    while (j > 0 && a[j - 1] > x) {
      a[j] = a[j - 1];
      --j;
    }
    a[j] = x;
  }
}
// lexicographic compare of two sorted site sets (std::set<Marker> ordering inside std::map)
GQ_DEV int cmp_key(const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb) {
  uint32_t n = na < nb ? na : nb;
  for (uint32_t i = 0; i < n; ++i) {
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  }
  return na == nb ? 0 : (na < nb ? -1 : 1);
}

// Traverser (allele_base.cpp:137-219) over the flat graph.
struct Trav {
  const IndexView* v;
  uint32_t cur;  // node id, kNoAllele = null
  uint32_t remaining;
  const uint32_t* T;
  uint32_t ti;
  bool first;
  uint32_t start_pos, end_pos;
  bool bad;
  GQ_DEV inline const Node& node() const { return GQ_AT(v->nodes, cur); }
  GQ_DEV void update_coordinates() {
    const Node& nd = node();
    end_pos = 0;
    if (nd.len > 0) {
      uint32_t a = nd.len - 1, bnd = start_pos + remaining - 1;
      end_pos = a < bnd ? a : bnd;
      remaining -= (end_pos - start_pos + 1);
    }
  }
  GQ_DEV void go_to_next_site() {
    start_pos = 0;
    while (node().n_edges == 1) {
      if (remaining == 0) {
        cur = kNoAllele;
        return;
      }
      cur = node().next0;
      update_coordinates();
      const Node& nd = node();
      if (nd.allele != -1 && nd.site != 0) return;
    }
    if (ti == 0 || node().n_edges == 0) {
      bad = true;
      cur = kNoAllele;
      return;
    }
    --ti;
    uint32_t allele = T[2 * ti + 1];
    if (allele >= node().n_edges) {
      bad = true;
      cur = kNoAllele;
      return;
    }
    cur = GQ_AT(v->edges, node().edge_off + allele);
    update_coordinates();
  }
  // returns false when the traversal is over
  GQ_DEV bool next() {
    if (first) {
      first = false;
      update_coordinates();
      const Node& nd = node();
      if (!(nd.allele != -1 && nd.site != 0)) go_to_next_site();
      return cur != kNoAllele;
    }
    if (remaining == 0) return false;
    go_to_next_site();
    return cur != kNoAllele;
  }
};

// Per-node hulls of one read (PbCovRecorder, allele_base.cpp:221-296): open-addressing table keyed by node id
// (a strand with dozens of final states visits thousands of nodes; the first version scanned a list per insertion,
// quadratic in exactly the strands that are slowest anyway).
struct Hull {
  uint32_t* e;   // cap triples (node, start, end); node == kNoAllele: empty slot
  uint32_t n, cap, shift;  // cap = 1 << (32 - shift)
  bool overflow;
  bool at_limit;  // no larger table fits the memory it was given
};
// A table of at most 2^want_lg slots (fewer if `words` cannot hold them). The table is sized for what a walk
// touches — a 150-base read visits a few dozen nodes — not for the arena: clearing and, at the end, scanning a table
// that filled the whole arena cost every strand 512 stores + 512 loads (65536 each in the large-arena re-runs).
// record_general starts small and rebuilds the table four times larger when it fills up.
GQ_DEV inline void hull_init(Hull& h, uint32_t* mem, uint32_t words, uint32_t want_lg) {
  uint32_t lg = 2;
  while (lg < 20 && lg < want_lg && 3u * (2u << lg) <= words) ++lg;
  h.e = mem;
  h.cap = 1u << lg;
  h.shift = 32 - lg;
  h.n = 0;
  h.overflow = 3u * h.cap > words;
  h.at_limit = lg < want_lg;  // the arena holds no larger table
  if (!h.overflow)
    for (uint32_t i = 0; i < h.cap; ++i) mem[3 * i] = kNoAllele;
}
GQ_DEV void hull_add(const IndexView& v, Hull& h, uint32_t node, uint32_t s, uint32_t e) {
  if (GQ_AT(v.nodes, node).len == 0) return;  // process_Node :282-287
  for (uint32_t i = (node * 2654435761u) >> h.shift;; i = (i + 1) & (h.cap - 1)) {
    uint32_t* t = h.e + 3 * i;
    if (t[0] == node) {
      if (s < t[1]) t[1] = s;
      if (e > t[2]) t[2] = e;
      return;
    }
    if (t[0] == kNoAllele) {
      if (2 * (h.n + 1) > h.cap) {  // more than half full: the strand is re-run with a larger arena
        h.overflow = true;
        return;
      }
      t[0] = node;
      t[1] = s;
      t[2] = e;
      ++h.n;
      return;
    }
  }
}

// `alleles`: n allele ids, `stride` words apart (2 inside a sorted locus list, 1 in a group record)
GQ_DEV uint32_t hash_group(uint32_t slot, const uint32_t* alleles, uint32_t n, uint32_t stride) {
  uint32_t h = 2166136261u ^ slot;
  h *= 16777619u;
  for (uint32_t i = 0; i < n; ++i) {
    h ^= alleles[stride * i];
    h *= 16777619u;
  }
  h ^= h >> 15;
  return h;
}

// multi-allele group of one site: find-or-insert in the open-addressing table. Returns the table slot (whose
// counter the caller bumps when it commits), or kNoAllele when the table or the record pool is full — the caller
// then gives the strand back uncommitted and the host grows the table (capi.cu: grow_groups). An inserted key
// that is never counted is harmless: readers skip zero counters.
GQ_DEV uint32_t grouped_find_or_insert(const CoverageView& c, uint32_t slot, const uint32_t* alleles, uint32_t n,
                                       uint32_t stride) {
  uint32_t maskc = c.gtab_cap - 1;
  uint32_t h = hash_group(slot, alleles, n, stride) & maskc;
  uint32_t mine = 0;  // offset + 1 of a record this thread allocated (lazily)
  uint32_t found = kNoAllele;
  // at most half the table is probed: a table that full is grown rather than searched
  for (uint32_t probe = 0; probe < (c.gtab_cap >> 1) + 1; ++probe, h = (h + 1) & maskc) {
    uint32_t cur = gq_atomic_add(c.gtab + h, 0u);
    if (cur == 0) {
      if (!mine) {
        uint32_t off = gq_atomic_add(c.gpool_used, n + 2);
        if (off + n + 2 > c.gpool_cap) {
          gq_atomic_or(c.error_flags, 1u);
          return kNoAllele;
        }
        c.gpool[off] = slot;
        c.gpool[off + 1] = n;
        for (uint32_t i = 0; i < n; ++i) c.gpool[off + 2 + i] = alleles[stride * i];
        gq_threadfence();
        mine = off + 1;
      }
      cur = gq_atomic_cas(c.gtab + h, 0u, mine);
      if (cur == 0) return h;
    }
    // occupied: same key?
    const volatile uint32_t* rec = c.gpool + (cur - 1);
    bool same = rec[0] == slot && rec[1] == n;
    for (uint32_t i = 0; same && i < n; ++i) same = rec[2 + i] == alleles[stride * i];
    if (same) {
      found = h;
      break;
    }
  }
  // a record allocated for a slot somebody else won: hand it back if it is still the newest allocation (many
  // threads meeting the same new group at kernel start would otherwise each leak one; what cannot be handed back
  // is dropped when the table is next rebuilt)
  if (mine) gq_atomic_cas(c.gpool_used, mine - 1 + n + 2, mine - 1);
  if (found == kNoAllele) gq_atomic_or(c.error_flags, 1u);
  return found;
}

// The general route of record_strand (several states, several occurrences, nested sites): MappingInstanceSelector,
// LocusFinder and PbCovRecorder restated over flat scratch. Nothing is committed before every scratch structure has
// proved large enough, so the caller can run it again (lists in the arena, or a larger arena) without double counting.
enum RecResult : uint32_t { REC_OK = 0, REC_LISTS = 1 /* LocusFinder's sets overflowed */, REC_ARENA = 2 };
constexpr uint32_t kLocalLoci = 96;

GQ_DEV uint32_t record_general(const IndexView& v, const BatchView& b, const CoverageView& c, uint32_t strand,
                               const uint32_t* recs, uint32_t ns, uint32_t L, uint32_t nonvar, uint32_t* arena,
                               uint32_t arena_words, uint32_t* local_lists) {
  Scratch sc{arena, 0, arena_words, false};
  uint32_t* key_off = sc.alloc(2 * ns);  // (offset, len) per state; len = 0xFFFFFFFF for path-less
  uint32_t* rep = sc.alloc(ns);          // class representative per state
  if (!key_off || !rep) return REC_ARENA;
  LocusLists ll;
  ll.overflow = false;
  if (local_lists) {  // the small sets of LocusFinder in thread-local memory (L1-resident) ...
    ll.cap = kLocalLoci;
    ll.used = local_lists;
    ll.base = local_lists + kLocalLoci;
    ll.loci = local_lists + 2 * kLocalLoci;
  } else {            // ... or, when a strand's sets do not fit there, in the arena
    const uint32_t cap = (arena_words - sc.used) / 16;
    ll.cap = cap;
    ll.used = sc.alloc(cap);
    ll.base = sc.alloc(cap);
    ll.loci = sc.alloc(2 * cap);
    if (sc.overflow) return REC_ARENA;
  }
  {
    const uint32_t* p = recs;
    for (uint32_t j = 0; j < ns; ++j) {
      StateRec st = parse_rec(p);
      p += st.words();
      if (!(st.nt | st.ng)) {
        key_off[2 * j] = 0;
        key_off[2 * j + 1] = 0xFFFFFFFFu;
        continue;
      }
      ll.n_loci = ll.n_base = 0;
      locus_finder(v, st, ll);
      if (ll.overflow) return REC_LISTS;
      sort_u32(ll.base, ll.n_base);
      uint32_t* k = sc.alloc(ll.n_base);
      if (!k && ll.n_base) return REC_ARENA;
      for (uint32_t i = 0; i < ll.n_base; ++i) k[i] = ll.base[i];
      key_off[2 * j] = (uint32_t)(k - arena);
      key_off[2 * j + 1] = ll.n_base;
    }
  }
  // class representative of every state with a path = the first state with the same key (compared with the
  // representatives found so far only), and the number of distinct classes
  uint32_t ncls = 0;
  for (uint32_t j = 0; j < ns; ++j) {
    rep[j] = 0xFFFFFFFFu;
    if (key_off[2 * j + 1] == 0xFFFFFFFFu) continue;
    rep[j] = j;
    for (uint32_t i = 0; i < j; ++i)
      if (rep[i] == i && cmp_key(arena + key_off[2 * i], key_off[2 * i + 1], arena + key_off[2 * j], key_off[2 * j + 1]) == 0) {
        rep[j] = i;
        break;
      }
    if (rep[j] == j) ++ncls;
  }
  // random_select_entry :97-107
  uint32_t total = nonvar + ncls;
  uint32_t pick = total == 1 ? 1u : uniform_1_to(GQ_AT(b.seeds, strand >> 1), total);
  if (pick <= nonvar) return REC_OK;
  // the (pick - nonvar - 1)-th class in std::map order: the representative with that many smaller ones
  uint32_t want = pick - nonvar - 1;
  int chosen = -1;
  for (uint32_t j = 0; j < ns && chosen < 0; ++j) {
    if (rep[j] != j) continue;
    uint32_t less = 0;
    for (uint32_t i = 0; i < ns; ++i)
      if (i != j && rep[i] == i &&
          cmp_key(arena + key_off[2 * i], key_off[2 * i + 1], arena + key_off[2 * j], key_off[2 * j + 1]) < 0)
        ++less;
    if (less == want) chosen = (int)j;
  }
  if (chosen < 0) {
    gq_atomic_or(c.error_flags, 2u);
    return REC_OK;
  }
  // ---- pass 2: loci of the chosen class + per-node hulls (PbCovRecorder :221-296) ----
  Hull hull;
  for (uint32_t hull_lg = 6;; hull_lg += 2) {  // 64 slots first; a table that fills up is rebuilt four times larger
    ll.n_loci = 0;
    hull_init(hull, arena + sc.used, arena_words - sc.used, hull_lg);
    if (hull.overflow) return REC_ARENA;
    const uint32_t* p = recs;
    for (uint32_t j = 0; j < ns && !hull.overflow; ++j) {
      StateRec st = parse_rec(p);
      p += st.words();
      if (rep[j] != (uint32_t)chosen) continue;  // not a state of the chosen class (or path-less)
      ll.n_base = 0;
      locus_finder(v, st, ll);  // loci accumulate across the class (set union)
      if (ll.overflow) return REC_LISTS;
      bool first = true;
      for (uint32_t occ = st.lo;; ++occ) {
        uint32_t pos = GQ_LDG(v.sa + occ);
        uint32_t nid = GQ_LDG(v.pos2node + pos);
        Trav t;
        t.v = &v;
        t.cur = nid;
        t.remaining = L;
        t.T = st.T;
        t.ti = st.nt;
        t.first = true;
        const Node& nd = GQ_AT(v.nodes, nid);
        t.start_pos = nd.len > 1 ? pos - nd.start : 0;  // setup_random_access, coverage_graph.cpp:131-144
        t.end_pos = 0;
        t.bad = false;
        if (first) {
          first = false;
          while (!hull.overflow && t.next()) hull_add(v, hull, t.cur, t.start_pos, t.end_pos);
        } else if (t.next())
          hull_add(v, hull, t.cur, t.start_pos, t.end_pos);
        if (t.bad) gq_atomic_or(c.error_flags, 2u);
        if (hull.overflow || occ == st.hi) break;
      }
    }
    if (!hull.overflow) break;
    if (hull.at_limit) return REC_ARENA;  // the arena holds no larger table: re-run with a larger arena
  }
  // sort loci by (site, allele) — std::set<VariantLocus> order
  for (uint32_t i = 1; i < ll.n_loci; ++i) {
    uint32_t s0 = ll.loci[2 * i], a0 = ll.loci[2 * i + 1], j = i;
    while (j > 0 && (ll.loci[2 * (j - 1)] > s0 ||
                     (ll.loci[2 * (j - 1)] == s0 && (int32_t)ll.loci[2 * (j - 1) + 1] > (int32_t)a0))) {
      ll.loci[2 * j] = ll.loci[2 * (j - 1)];
      ll.loci[2 * j + 1] = ll.loci[2 * (j - 1) + 1];
      --j;
    }
    ll.loci[2 * j] = s0;
    ll.loci[2 * j + 1] = a0;
  }
  // multi-allele groups first find (or create) their table slots — kept in `used`, no longer needed by the locus
  // finder; a full table gives the strand back before any counter is touched (grouped_allele_counts.cpp:17-49)
  {
    uint32_t ng = 0;
    for (uint32_t i = 0; i < ll.n_loci;) {
      uint32_t e = i;
      while (e < ll.n_loci && ll.loci[2 * e] == ll.loci[2 * i]) ++e;
      if (e - i > 1) {
        if (ng >= ll.cap) return REC_LISTS;
        const uint32_t h = grouped_find_or_insert(c, (ll.loci[2 * i] - 5) >> 1, ll.loci + 2 * i + 1, e - i, 2);
        if (h == kNoAllele) return REC_ARENA;
        ll.used[ng++] = h;
      }
      i = e;
    }
  }
  // ---- commit (nothing above touched the counters, so an overflow re-run cannot double count) ----
  for (uint32_t i = 0, ng = 0; i < ll.n_loci;) {
    uint32_t site = ll.loci[2 * i], slot = (site - 5) >> 1;
    uint32_t e = i;
    while (e < ll.n_loci && ll.loci[2 * e] == site) {
      gq_red_add(c.allele_sum + GQ_AT(c.allele_off, slot) + ll.loci[2 * e + 1], 1u);  // allele_sum.cpp:31-43
      ++e;
    }
    if (e - i == 1) gq_red_add(c.grouped_single + GQ_AT(c.allele_off, slot) + ll.loci[2 * i + 1], 1u);
    else gq_red_add(c.gcount + ll.used[ng++], 1u);
    i = e;
  }
  for (uint32_t i = 0; i < hull.cap; ++i) {
    if (hull.e[3 * i] == kNoAllele) continue;
    const Node& nd = GQ_AT(v.nodes, hull.e[3 * i]);
    if (nd.cov_off == kNoAllele) continue;
    for (uint32_t x = hull.e[3 * i + 1]; x <= hull.e[3 * i + 2]; ++x) gq_red_add(c.per_base + nd.cov_off + x, 1u);
  }
  return REC_OK;
}

GQ_DEV bool record_strand(const IndexView& v, const BatchView& b, const SearchOut& o, const CoverageView& c,
                              uint32_t strand, uint32_t* arena, uint32_t arena_words) {
  GQ_PHASE(5);  // coverage_kernel
  const uint32_t ns = GQ_AT(o.st_count, strand);
  const uint32_t* recs = o.pool + GQ_AT(o.st_off, strand);
  GQ_TOUCH(recs, 4 * GQ_AT(o.st_words, strand));
  const uint32_t L = GQ_AT(b.len, strand >> 1);

  // ---- pass 1: non-variant mapping count, per-state class keys (MappingInstanceSelector) ----
  uint32_t nonvar = 0, npath = 0;
  {
    const uint32_t* p = recs;
    for (uint32_t j = 0; j < ns; ++j) {
      StateRec st = parse_rec(p);
      p += st.words();
      if (st.nt | st.ng) ++npath;
      else nonvar += st.hi - st.lo + 1;  // count_nonvar_search_states :137-148
    }
  }
  if (npath == 0) return true;
  // Table route (nearly every strand of a non-nested PRG): exactly one state, one occurrence. One class and no
  // non-variant mapping -> generate(1,1) == 1 selects it (coverage_common.cpp:97-107); every locus is its own
  // level-0 site, each allele is visited once, so no sets / hulls are needed — and no graph either: the forward
  // traversal of PbCovRecorder / Traverser (allele_base.cpp:137-296) from the read's first base, choosing the
  // path's alleles back to front, is arithmetic on text positions: from position p the next site of the path
  // starts `site start - p` bases further on, the chosen allele a spans [apos[a], apos[a+1] - 1), its bases sit at
  // per-base offset `site base + (apos[a] - apos[0]) - a`, and the walk continues after the site-end marker. The
  // per-site records of the path's sites are independent loads (fetched four at a time), where the graph walk
  // was a chain of ~14 dependent node loads per read.
  if (ns == 1 && !v.any_nested) {
    StateRec st = parse_rec(recs);
    if (st.lo == st.hi && st.ng <= 1) {
      uint32_t p = GQ_LDG(v.sa + st.lo);
      uint32_t remaining = L;
      const uint32_t n_el = st.nt + st.ng;  // elements in read order: the open site (read starts inside it), then
                                            // the path back to front
      constexpr uint32_t kChunk = GQ_COV_CHUNK;
      for (uint32_t e0 = 0; e0 < n_el; e0 += kChunk) {
        uint32_t site[kChunk], al[kChunk], r_a2[kChunk], r_cov[kChunk], r_first[kChunk], r_after[kChunk];
        uint32_t a_lo[kChunk], a_hi[kChunk];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t q = 0; q < kChunk; ++q) {
          const uint32_t e = e0 + q;
          if (e >= n_el) break;
          if (st.ng && e == 0) {
            site[q] = st.G[0];
            al[q] = kNoAllele;
          } else {
            const uint32_t j = st.nt - 1 - (e - st.ng);
            site[q] = st.T[2 * j];
            al[q] = st.T[2 * j + 1];
          }
          const uint32_t slot = (site[q] - 5) >> 1;
#if defined(__CUDA_ARCH__)
          const uint4 r = __ldg(reinterpret_cast<const uint4*>(v.site_rec) + slot);
          r_a2[q] = r.x, r_cov[q] = r.y, r_first[q] = r.z, r_after[q] = r.w;
#else
          GQ_TOUCH(v.site_rec + 4 * (size_t)slot, 16);
          r_a2[q] = v.site_rec[4 * (size_t)slot], r_cov[q] = v.site_rec[4 * (size_t)slot + 1];
          r_first[q] = v.site_rec[4 * (size_t)slot + 2], r_after[q] = v.site_rec[4 * (size_t)slot + 3];
#endif
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t q = 0; q < kChunk; ++q) {
          if (e0 + q >= n_el) break;
          if (al[q] == kNoAllele) {  // the allele the read starts in: the one whose span holds p
            uint32_t a = 0;
            while (p >= GQ_LDG(v.apos + r_a2[q] + a + 1)) ++a;
            al[q] = a;
          }
          a_lo[q] = GQ_LDG(v.apos + r_a2[q] + al[q]);
          a_hi[q] = GQ_LDG(v.apos + r_a2[q] + al[q] + 1);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (uint32_t q = 0; q < kChunk; ++q) {
          const uint32_t e = e0 + q;
          if (e >= n_el) break;
          const uint32_t slot = (site[q] - 5) >> 1, ai = r_a2[q] - slot + al[q];  // allele_off[slot] + allele
          gq_red_add(c.allele_sum + ai, 1u);       // allele_sum.cpp:31-43
          gq_red_add(c.grouped_single + ai, 1u);   // grouped_allele_counts.cpp:17-49, a one-allele group
          if (remaining == 0) continue;
          uint32_t off = 0;
          if (e == 0 && p >= r_first[q]) off = p - a_lo[q];  // the read starts inside this allele
          else {
            const uint32_t gap = r_first[q] - 1 - p;  // bases before the site-entry marker
            if (remaining <= gap) {
              remaining = 0;
              continue;
            }
            remaining -= gap;
          }
          const uint32_t len = a_hi[q] - 1 - a_lo[q];
          const uint32_t cnt = len - off < remaining ? len - off : remaining;
          uint32_t* pb = c.per_base + r_cov[q] + (a_lo[q] - r_first[q]) - al[q] + off;
          for (uint32_t x = 0; x < cnt; ++x) gq_red_add(pb + x, 1u);  // allele_base.cpp:221-296
          remaining -= cnt;
          p = r_after[q];
        }
      }
      return true;
    }
  }
  // Nested PRG, one state, one occurrence (about half of the mapped strands of a nested PRG): still a single class
  // that generate(1,1) selects, and one forward walk that visits every node once — no keys, no hulls. The loci are
  // the path's sites and their parents (LocusFinder), at most one allele per site, so every group has one allele.
  if (ns == 1 && v.any_nested) {
    StateRec st = parse_rec(recs);
    if (st.lo == st.hi) {
      constexpr uint32_t kLoc = 40;
      uint32_t used_l[kLoc], base_l[kLoc], loci_l[2 * kLoc];
      LocusLists l1;
      l1.loci = loci_l, l1.base = base_l, l1.used = used_l;
      l1.n_loci = l1.n_base = l1.n_used = 0;
      l1.cap = kLoc;
      l1.overflow = false;
      locus_finder(v, st, l1);
      if (!l1.overflow) {
        for (uint32_t i = 0; i < l1.n_loci; ++i) {
          const uint32_t ai = GQ_AT(c.allele_off, (l1.loci[2 * i] - 5) >> 1) + l1.loci[2 * i + 1];
          gq_red_add(c.allele_sum + ai, 1u);
          gq_red_add(c.grouped_single + ai, 1u);
        }
        const uint32_t pos0 = GQ_LDG(v.sa + st.lo);
        const uint32_t nid0 = GQ_LDG(v.pos2node + pos0);
        Trav t;
        t.v = &v;
        t.cur = nid0;
        t.remaining = L;
        t.T = st.T;
        t.ti = st.nt;
        t.first = true;
        const Node& nd0 = GQ_AT(v.nodes, nid0);
        t.start_pos = nd0.len > 1 ? pos0 - nd0.start : 0;
        t.end_pos = 0;
        t.bad = false;
        while (t.next()) {
          const Node& nd = GQ_AT(v.nodes, t.cur);
          if (nd.len == 0 || nd.cov_off == kNoAllele) continue;
          for (uint32_t x = t.start_pos; x <= t.end_pos; ++x) gq_red_add(c.per_base + nd.cov_off + x, 1u);
        }
        if (t.bad) gq_atomic_or(c.error_flags, 2u);
        return true;
      }
    }
  }
  // the general route: LocusFinder's sets first in thread-local memory, in the arena if a strand outgrows that
  uint32_t local_lists[4 * kLocalLoci];
  uint32_t r = record_general(v, b, c, strand, recs, ns, L, nonvar, arena, arena_words, local_lists);
  if (r == REC_LISTS) r = record_general(v, b, c, strand, recs, ns, L, nonvar, arena, arena_words, nullptr);
  return r == REC_OK;
}

}  // namespace gq
