// emu_capi.cpp — TEST-ONLY host emulation of the device functions in gramtools_b200/csrc/gq_device.cuh.
//
// Compiles the exact per-strand code of the CUDA kernels (map_strand, record_strand) for the host and
// runs it one strand at a time, so that the flat index + search/coverage logic can be checked against
// the oracle in the GPU-less container (`pytest -m "not gpu"`). This is debugging infrastructure: it is
// built into tests/_build/libgq_emu.so, never into libgq.so, and is not a supported CPU path.
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#define GQ_EMU_COUNTERS 1
#include "../../gramtools_b200/csrc/gq_device.cuh"
#include "../../gramtools_b200/csrc/index_build.hpp"
#include "../../gramtools_b200/csrc/sais.hpp"

#include <set>
#include <unordered_set>

#include "../../gramtools_b200/csrc/read_file.hpp"
#include "old_read_file.hpp"

using namespace gq;
namespace gq {
unsigned long long gq_emu_counters[32];

// ---- byte model (tools/byte_model.py): distinct 32 B sectors touched per THREAD-SIZED unit of work ------------
// (one strand in seed / general / classify / coverage, one candidate in verify / text), per kernel and structure.
// Off unless emu_track(1) was called: the index builder runs through the same accessors.
constexpr int kKernels = 6, kMaxRanges = 64;
struct Range {
  const char* name;
  uintptr_t lo, hi;
  unsigned granule;  // 32: random access, a touch costs its whole sector; 4 / 1: per-strand / per-candidate arrays
                     // that consecutive threads read or write side by side (coalesced): a touch costs its bytes
};
static bool g_track = false;
static int g_kernel = 0, g_n_ranges = 0;
static Range g_ranges[kMaxRanges];
static std::unordered_set<uint64_t> g_sectors;                  // (range id << 48) | sector index within the range
static unsigned long long g_bytes[kKernels][kMaxRanges + 1];  // [kernel][range] (last = unregistered memory)
static unsigned long long g_units[kKernels];

static void flush_unit() {
  if (g_sectors.empty()) return;
  for (uint64_t key : g_sectors) g_bytes[g_kernel][key >> 48] += g_ranges[key >> 48].granule;
  g_units[g_kernel]++;
  g_sectors.clear();
}
void gq_emu_phase(int kernel) {
  if (!g_track) return;
  flush_unit();
  g_kernel = kernel;
}
void gq_emu_touch(const void* p, unsigned bytes) {
  if (!g_track || bytes == 0) return;
  const uintptr_t a = (uintptr_t)p;
  int id = kMaxRanges;
  for (int i = 0; i < g_n_ranges; ++i)
    if (a >= g_ranges[i].lo && a < g_ranges[i].hi) {
      id = i;
      break;
    }
  if (id == kMaxRanges) return;  // host-only scratch (stacks, local arrays): not device-memory traffic
  // device allocations start on a sector boundary: the offset from the start of the array decides the sector
  const uintptr_t base = g_ranges[id].lo, g = g_ranges[id].granule;
  for (uintptr_t s = (a - base) / g; s <= (a + bytes - 1 - base) / g; ++s) g_sectors.insert(((uint64_t)id << 48) | s);
}
static void add_range(const char* name, const void* p, size_t bytes, unsigned granule) {
  if (!bytes || g_n_ranges == kMaxRanges) return;
  for (int i = 0; i < g_n_ranges; ++i)
    if (std::string(g_ranges[i].name) == name) {
      g_ranges[i] = {name, (uintptr_t)p, (uintptr_t)p + bytes, granule};
      return;
    }
  g_ranges[g_n_ranges++] = {name, (uintptr_t)p, (uintptr_t)p + bytes, granule};
}
}  // namespace gq

struct Emu {
  HostIndex h;
  std::vector<uint32_t> counters, gtab, gcount, gpool, gsmall;
  unsigned long long stats[5] = {0, 0, 0, 0, 0};
  // last batch
  std::vector<uint8_t> status;
  std::vector<uint32_t> st_off, st_words, st_count, pool;
  uint32_t n_reads = 0;
  uint64_t reruns = 0;
};
static thread_local std::string g_err;
static int g_force_general = 0;

extern "C" {
const char* emu_last_error() { return g_err.c_str(); }

struct Emu;
void* emu_finish(Emu* e);
void* emu_new_from(const uint32_t* prg, uint64_t n, uint32_t k, const char* kmer_index_dir);
void* emu_new(const uint32_t* prg, uint64_t n, uint32_t k) { return emu_new_from(prg, n, k, nullptr); }
// kmer_index_dir != NULL: the k-mer index comes from the sdsl files of that gram_dir (kmer_index::load)
void* emu_new_from(const uint32_t* prg, uint64_t n, uint32_t k, const char* kmer_index_dir) {
  try {
    auto* e = new Emu();
    build_host_index(prg, n, k, e->h, nullptr, nullptr, kmer_index_dir);
    return emu_finish(e);
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
// coverage accumulators of a freshly built / loaded index
void* emu_finish(Emu* e) {
  uint64_t na = e->h.allele_off.back();
  e->counters.assign(2 * na + e->h.n_per_base + 1, 0);
  uint32_t cap = 1024;
  while (cap < 4 * na) cap <<= 1;
  e->gtab.assign(cap, 0);
  e->gcount.assign(cap, 0);
  e->gpool.assign((size_t)cap * 4, 0);
  e->gsmall.assign(4, 0);
  return e;
}
void emu_free(void* e) { delete (Emu*)e; }
// Reader equivalence (tests/test_read_file.py). mode 0: the old line-by-line reader; 1: ReadFile::next; 2:
// ReadFile::next_seq (sequence appended to one text, qualities only counted). out = {records, bases, digest of the
// sequences (length-prefixed), digest of the qualities (modes 0 and 1)}; buffer_bytes = inflate buffer of the new reader
// (tiny values force lines to straddle refills). Returns -1 when the file cannot be opened.
int emu_read_file(const char* path, int mode, uint64_t buffer_bytes, uint64_t out[4]) {
  auto fnv = [](uint64_t& d, const char* p, size_t n) {
    d = (d ^ n) * 1099511628211ull;
    for (size_t i = 0; i < n; ++i) d = (d ^ (uint8_t)p[i]) * 1099511628211ull;
  };
  out[0] = out[1] = 0;
  out[2] = out[3] = 1469598103934665603ull;
  try {
    std::string seq, qual;
    if (mode == 0) {
      OldReadFile rf(path);
      while (rf.next(seq, qual)) {
        ++out[0];
        out[1] += seq.size();
        fnv(out[2], seq.data(), seq.size());
        fnv(out[3], qual.data(), qual.size());
      }
    } else if (mode == 1) {
      gq::ReadFile rf(path, buffer_bytes ? buffer_bytes : (size_t(4) << 20));
      while (rf.next(seq, qual)) {
        ++out[0];
        out[1] += seq.size();
        fnv(out[2], seq.data(), seq.size());
        fnv(out[3], qual.data(), qual.size());
      }
    } else {
      gq::ReadFile rf(path, buffer_bytes ? buffer_bytes : (size_t(4) << 20));
      std::string text;
      size_t len = 0, at = 0;
      while (rf.next_seq(text, len)) {
        ++out[0];
        out[1] += len;
        if (text.size() != at + len) return -2;
        fnv(out[2], text.data() + at, len);
        at += len;
        if (text.size() > (1u << 20)) text.clear(), at = 0;
      }
      if (text.size() != at) return -3;  // a rejected record must leave nothing behind
    }
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
int emu_index_save(void* ev, const char* path) {
  try {
    host_index_save(((Emu*)ev)->h, path);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
void* emu_load(const char* path) {
  try {
    auto* e = new Emu();
    host_index_load(e->h, path);
    return emu_finish(e);
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}
int emu_kmer_index_dump(void* ev, const char* dir) {
  try {
    kmer_index_dump(((Emu*)ev)->h, dir);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
// sdsl::int_vector serialisation as restated in index_build.cpp (width 0 on read = run-time width from the file)
int emu_write_int_vector(const char* path, const uint64_t* values, uint64_t n, uint32_t width, int fixed_width) {
  try {
    write_int_vector(path, std::vector<uint64_t>(values, values + n), width, fixed_width != 0);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
int64_t emu_read_int_vector(const char* path, uint32_t fixed_width, uint64_t* out, uint64_t cap, uint32_t* width) {
  try {
    std::vector<uint64_t> v = read_int_vector(path, fixed_width, width);
    for (uint64_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return (int64_t)v.size();
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
// test hook: shrink the multi-allele group table (power of two) so that its growth path runs
void emu_set_gtab_cap(void* ev, uint32_t cap) {
  auto* e = (Emu*)ev;
  e->gtab.assign(cap, 0);
  e->gcount.assign(cap, 0);
  e->gpool.assign((size_t)cap * 4, 0);
  e->gsmall.assign(4, 0);
}

void emu_sizes(void* ev, uint64_t out[6]) {
  auto* e = (Emu*)ev;
  out[0] = e->h.n_sites;
  out[1] = e->h.allele_off.back();
  out[2] = e->h.n_per_base;
  out[3] = e->h.is_nested;
  out[4] = e->h.n;
  out[5] = e->h.kmer_off.back();
}

// k-mer index dump for parity with the oracle's index: per k-mer code the canonical record list
uint64_t emu_kmer_states(void* ev, uint32_t* words) {
  auto* e = (Emu*)ev;
  const HostIndex& h = e->h;
  uint64_t t = 0;
  uint64_t nk = 1ull << (2 * h.k);
  for (uint64_t c = 0; c < nk; ++c) {
    for (uint32_t j = h.kmer_off[c]; j < h.kmer_off[c + 1]; ++j) {
      const KmerState& ks = h.kmer_states[j];
      uint32_t nt = ks.counts & 0xFFFF, ng = ks.counts >> 16;
      if (words) {
        words[t] = (uint32_t)c;
        words[t + 1] = ks.lo;
        words[t + 2] = ks.hi;
        words[t + 3] = nt;
        words[t + 4] = ng;
        for (uint32_t i = 0; i < 2 * nt; ++i) words[t + 5 + i] = h.kmer_paths[ks.path_off + i];
        for (uint32_t i = 0; i < ng; ++i) {
          words[t + 5 + 2 * nt + 2 * i] = h.kmer_paths[ks.path_off + 2 * nt + i];
          words[t + 5 + 2 * nt + 2 * i + 1] = kNoAllele;
        }
      }
      t += 5 + 2 * nt + 2 * ng;
    }
  }
  return t;
}
// Replace the k-mer index by hand-written states (tests of the gram_dir files against the literals of the reference's
// test_dump_and_load.cpp): words = records [code, lo, hi, nt, ng, (site, allele) * nt, (site, ignored) * ng], any order
// of k-mers, states of one k-mer in the order given. The seed view is not rebuilt (the states need not exist in the PRG).
int emu_set_kmer_index(void* ev, const uint32_t* words, uint64_t n_words) {
  auto* e = (Emu*)ev;
  HostIndex& h = e->h;
  const uint64_t nk = 1ull << (2 * h.k);
  std::vector<std::vector<const uint32_t*>> per_code(nk);
  for (uint64_t t = 0; t < n_words;) {
    const uint32_t* r = words + t;
    if (r[0] >= nk) return -1;
    per_code[r[0]].push_back(r);
    t += 5 + 2 * r[3] + 2 * r[4];
  }
  h.kmer_off.assign(nk + 1, 0);
  h.kmer_states.clear();
  h.kmer_paths.clear();
  std::fill(h.kmer_bits.begin(), h.kmer_bits.end(), 0u);
  for (uint64_t c = 0; c < nk; ++c) {
    for (const uint32_t* r : per_code[c]) {
      const uint32_t nt = r[3], ng = r[4];
      h.kmer_states.push_back(KmerState{r[1], r[2], (uint32_t)h.kmer_paths.size(), nt | (ng << 16)});
      for (uint32_t i = 0; i < 2 * nt; ++i) h.kmer_paths.push_back(r[5 + i]);
      for (uint32_t i = 0; i < ng; ++i) h.kmer_paths.push_back(r[5 + 2 * nt + 2 * i]);
    }
    h.kmer_off[c + 1] = (uint32_t)h.kmer_states.size();
    if (!per_code[c].empty()) h.kmer_bits[c >> 5] |= 1u << (c & 31);
  }
  if (h.kmer_paths.empty()) h.kmer_paths.push_back(0);
  if (h.kmer_states.empty()) h.kmer_states.push_back(KmerState{});
  return 0;
}
// SA-IS with 64-bit indices (texts beyond 2^31 symbols) against the 32-bit instantiation on the same text
int emu_sais64_agrees(const int32_t* text, uint64_t n, int32_t sigma) {
  std::vector<int32_t> t(text, text + n);
  std::vector<int32_t> a = suffix_array(t, sigma);
  std::vector<int64_t> b = suffix_array64(t, sigma);
  std::vector<uint32_t> c = suffix_array_u32(t, sigma, true);
  for (uint64_t i = 0; i < n; ++i)
    if ((int64_t)a[i] != b[i] || (uint32_t)a[i] != c[i]) return 0;
  return 1;
}
void emu_sa(void* ev, uint32_t* out) {
  auto* e = (Emu*)ev;
  std::memcpy(out, e->h.sa.data(), e->h.sa.size() * 4);
}

int emu_map(void* ev, const uint8_t* bases, const uint64_t* off, uint64_t n_reads, const uint32_t* seeds,
            uint32_t arena_words) {
  auto* e = (Emu*)ev;
  try {
    const HostIndex& h = e->h;
    IndexView v = h.view();
    // pack as pack_kernel does
    std::vector<uint32_t> word_off(n_reads + 1), len(n_reads), packed;
    for (uint64_t r = 0; r < n_reads; ++r) {
      word_off[r] = (uint32_t)packed.size();
      uint32_t L = (uint32_t)(off[r + 1] - off[r]);
      len[r] = L;
      for (uint32_t w = 0; w < (L + 15) / 16; ++w) {
        uint32_t x = 0;
        for (uint32_t j = 0; j < 16 && w * 16 + j < L; ++j) x |= ((uint32_t)(bases[off[r] + w * 16 + j] - 1) & 3u) << (2 * j);
        packed.push_back(x);
      }
    }
    word_off[n_reads] = (uint32_t)packed.size();
    packed.push_back(0);
    // reverse strands as revcomp_kernel builds them — and checked here base by base against the definition
    std::vector<uint32_t> packed_rc(packed.size(), 0);
    for (uint64_t r = 0; r < n_reads; ++r)
      for (uint32_t m = 0; m < (len[r] + 15) / 16; ++m) {
        packed_rc[word_off[r] + m] = revcomp_word(packed.data() + word_off[r], len[r], m);
        uint32_t want = 0;
        for (uint32_t j = 16 * m; j < 16 * m + 16 && j < len[r]; ++j)
          want |= (3u - (((uint32_t)bases[off[r] + len[r] - 1 - j] - 1) & 3u)) << (2 * (j & 15u));
        if (packed_rc[word_off[r] + m] != want) throw std::runtime_error("revcomp_word disagrees with the definition");
      }
    e->n_reads = (uint32_t)n_reads;
    e->status.assign(2 * n_reads, 0);
    e->st_off.assign(2 * n_reads, 0);
    e->st_words.assign(2 * n_reads, 0);
    e->st_count.assign(2 * n_reads, 0);
    e->pool.assign(std::max<size_t>(1 << 16, n_reads * 256), 0);
    std::vector<uint32_t> small(8, 0), ovf(2 * n_reads + 1), cov_ovf(2 * n_reads + 1), mapped(4 * n_reads + 1);
    BatchView b{packed.data(), packed_rc.data(), word_off.data(), len.data(), seeds, (uint32_t)n_reads, 0, (uint32_t)n_reads};
    std::vector<uint32_t> seed_rec(4 * 4096), surv_cnt(2 * n_reads + 1, 0), gen(2 * n_reads + 1), pre_small(2, 0);
    SearchOut o{e->status.data(), e->st_off.data(), e->st_words.data(), e->st_count.data(), e->pool.data(),
                (uint32_t)e->pool.size(), &small[0], ovf.data(), &small[1], mapped.data(), &small[3], &small[4],
                surv_cnt.data()};
    CoverageView c{};
    uint64_t na = h.allele_off.back();
    c.allele_sum = e->counters.data();
    c.grouped_single = e->counters.data() + na;
    c.per_base = e->counters.data() + 2 * na;
    c.gtab = e->gtab.data();
    c.gcount = e->gcount.data();
    c.gtab_cap = (uint32_t)e->gtab.size();
    c.gpool = e->gpool.data();
    c.gpool_cap = (uint32_t)e->gpool.size();
    c.gpool_used = &e->gsmall[0];
    c.error_flags = &e->gsmall[1];
    c.stats = e->stats;
    c.allele_off = h.allele_off.data();
    SeedOut pre{seed_rec.data(), (uint32_t)(seed_rec.size() / 4), &pre_small[0], surv_cnt.data(), gen.data(),
                &pre_small[1]};
    std::vector<uint32_t> arena(arena_words), big;
    if (g_track) {
#define REG(name, vec) add_range(name, (vec).data(), (vec).size() * sizeof((vec)[0]), 32)
#define REGS(name, vec, g) add_range(name, (vec).data(), (vec).size() * sizeof((vec)[0]), g)
      REGS("packed reads", packed, 4); REGS("packed reverse strands", packed_rc, 4); REGS("read len", len, 4); REGS("read word_off", word_off, 4);
      add_range("read seeds", seeds, n_reads * 4, 4);
      REGS("strand status", e->status, 1); REGS("strand st_off", e->st_off, 4); REGS("strand st_words", e->st_words, 4);
      REGS("strand st_count", e->st_count, 4); REGS("final-state pool", e->pool, 4); REGS("mapped list", mapped, 4);
      REGS("candidate records", seed_rec, 4); REGS("strand flags (surv_cnt)", surv_cnt, 4); REGS("general list", gen, 4);
      REG("coverage counters", e->counters); REG("group table", e->gtab); REG("group counts", e->gcount);
      REG("group pool", e->gpool);
      REG("rank_blk", h.rank_blk); REG("super_cnt", h.super_cnt); REG("mrank_blk", h.mrank_blk);
      REG("marker_hit", h.marker_hit); REG("text_grp", h.text_grp); REG("text_super", h.text_super);
      REG("tmarker_hit", h.tmarker_hit); REG("isa", h.isa); REG("site_sa", h.site_sa); REG("allele_iv", h.allele_iv);
      REG("par", h.par); REG("tm_odd", h.tm_odd); REG("tm_even_off", h.tm_even_off); REG("tm_even", h.tm_even);
      REG("entry_next", h.entry_next); REG("site_snp", h.site_snp); REG("sa", h.sa); REG("pos2node", h.pos2node);
      REG("nodes", h.nodes); REG("edges", h.edges); REG("kmer_bits (smem copy on the device)", h.kmer_bits);
      REG("kmer_off", h.kmer_off); REG("kmer_states", h.kmer_states);
      REG("seed_off", h.seed_off); REG("seed_ent", h.seed_ent); REG("seed_state", h.seed_state);
      REG("kmer_paths", h.kmer_paths); REG("allele_off", h.allele_off);
#undef REGS
#undef REG
    }
    for (uint32_t s = 0; s < 2 * n_reads; ++s) {
      uint32_t aw = arena_words;
      uint32_t* a = arena.data();
      while (true) {  // same policy as the library: re-run the strand with a 4x larger arena
        small[1] = 0;
        small[3] = 0;
        if (g_force_general) e->status[s] = ST_OVERFLOW;  // map_strand's re-run route = general machinery
        map_strand(v, v.super_cnt, b, o, pre, s, a, aw);
        if (e->status[s] != ST_OVERFLOW) break;
        e->reruns++;
        if (small[0] > e->pool.size()) {  // the final-state pool ran out (libgq grows it the same way): redo the strand
          const size_t used = std::min<size_t>(e->pool.size(), small[0]);
          e->pool.resize(std::max<size_t>(2 * e->pool.size(), 2 * (size_t)small[0]), 0);
          small[0] = (uint32_t)used;
          o.pool = e->pool.data();
          o.pool_cap = (uint32_t)e->pool.size();
          continue;
        }
        aw *= 4;
        if (aw > (1u << 28)) throw std::runtime_error("emu: arena overflow persists");
        big.assign(aw, 0);
        a = big.data();
      }
      if (e->status[s] != ST_MAPPED) continue;
      while (!record_strand(v, b, o, c, s, a, aw)) {
        e->reruns++;
        if (e->gsmall[1] & 1u) {  // group table / pool full: rebuild 4x larger (libgq: grow_groups), strand again
          std::vector<uint32_t> ot = e->gtab, oc = e->gcount, op = e->gpool;
          const size_t cap = ot.size() * 4;
          e->gtab.assign(cap, 0);
          e->gcount.assign(cap, 0);
          e->gpool.assign(std::max(op.size(), cap * 4), 0);
          uint32_t used = 0;
          for (size_t i = 0; i < ot.size(); ++i) {
            if (!ot[i] || !oc[i]) continue;
            const uint32_t* rec = op.data() + (ot[i] - 1);
            uint32_t hsh = 2166136261u ^ rec[0];
            hsh *= 16777619u;
            for (uint32_t q = 0; q < rec[1]; ++q) {
              hsh ^= rec[2 + q];
              hsh *= 16777619u;
            }
            hsh ^= hsh >> 15;
            hsh &= (uint32_t)cap - 1;
            while (e->gtab[hsh]) hsh = (hsh + 1) & ((uint32_t)cap - 1);
            e->gtab[hsh] = used + 1;
            e->gcount[hsh] = oc[i];
            std::memcpy(e->gpool.data() + used, rec, (2 + rec[1]) * 4);
            used += 2 + rec[1];
          }
          e->gsmall[0] = used;
          e->gsmall[1] &= ~1u;
          c.gtab = e->gtab.data();
          c.gcount = e->gcount.data();
          c.gtab_cap = (uint32_t)cap;
          c.gpool = e->gpool.data();
          c.gpool_cap = (uint32_t)e->gpool.size();
          continue;
        }
        aw *= 4;
        if (aw > (1u << 28)) throw std::runtime_error("emu: coverage scratch overflow persists");
        big.assign(aw, 0);
        a = big.data();
      }
    }
    if (g_track) gq_emu_phase(0);  // flush the last unit
    for (uint64_t r = 0; r < n_reads; ++r) {
      e->stats[0] += 2;
      for (int s = 0; s < 2; ++s) {
        uint8_t st = e->status[2 * r + s];
        if (st == ST_SKIPPED) e->stats[1]++;
        else if (st == ST_MISSING_KMER) e->stats[2]++;
        else if (st == ST_NO_EXTENSION) e->stats[3]++;
        else if (st == ST_MAPPED) e->stats[4]++;
      }
    }
    if (e->gsmall[1]) throw std::runtime_error("emu: coverage error flags " + std::to_string(e->gsmall[1]));
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}
// byte model: switch tracking on/off (resets the totals), and read them: per kernel and structure
void emu_track(int on) {
  g_track = on != 0;
  g_sectors.clear();
  std::memset(g_bytes, 0, sizeof g_bytes);
  std::memset(g_units, 0, sizeof g_units);
  g_kernel = 0;
}
int emu_track_ranges() { return g_n_ranges; }
const char* emu_track_range_name(int i) { return i < g_n_ranges ? g_ranges[i].name : ""; }
void emu_track_read(unsigned long long* bytes /* kKernels x n_ranges */, unsigned long long* units /* kKernels */) {
  for (int k = 0; k < kKernels; ++k) {
    units[k] = g_units[k];
    for (int i = 0; i < g_n_ranges; ++i) bytes[k * g_n_ranges + i] = g_bytes[k][i];
  }
}
// 1: every strand through the general lane machine (as libgq does with the seed_pass option off)
void emu_force_general(int on) { g_force_general = on; }

void emu_path_counters(uint64_t* out32, int reset) {
  for (int i = 0; i < 32; ++i) {
    out32[i] = gq_emu_counters[i];
    if (reset) gq_emu_counters[i] = 0;
  }
}
// Structural invariants of the flat index (what the kernels assume about text mode, the seed view, ...),
// checked directly on the host copy. Returns 0, or -1 with the first violation in emu_last_error().
// FNV-1a over every array of the flat index: two builds of one PRG must give the same value whatever the thread
// count, the suffix-array builder or the build order inside the parallel passes
// layout_free != 0: the k-mer states are mixed in as (lo, hi, counts, path words) instead of (lo, hi, path_off, counts) +
// the path pool, so that two indexes that differ only in where the paths sit in the pool give the same value
uint64_t emu_index_digest2(void* ev, int layout_free);
uint64_t emu_index_digest(void* ev) { return emu_index_digest2(ev, 0); }
uint64_t emu_index_digest2(void* ev, int layout_free) {
  const HostIndex& h = ((Emu*)ev)->h;
  uint64_t d = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t bytes) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < bytes; ++i) d = (d ^ b[i]) * 1099511628211ull;
    d = (d ^ bytes) * 1099511628211ull;
  };
#define MIX(vec) mix((vec).data(), (vec).size() * sizeof((vec)[0]))
  MIX(h.prg); MIX(h.sa); MIX(h.isa); MIX(h.rank_blk); MIX(h.super_cnt); MIX(h.mrank_blk); MIX(h.marker_hit);
  MIX(h.tmarker_hit); MIX(h.text_grp); MIX(h.text_super); MIX(h.pos2node); MIX(h.nodes); MIX(h.edges);
  MIX(h.site_sa); MIX(h.allele_iv); MIX(h.entry_next); MIX(h.site_snp); MIX(h.allele_off); MIX(h.n_alleles);
  MIX(h.kmer_bits); MIX(h.kmer_off); MIX(h.seed_off); MIX(h.seed_ent);
  if (!layout_free) {
    MIX(h.kmer_states); MIX(h.kmer_paths);
  } else {
    const uint64_t n_states = h.kmer_off.empty() ? 0 : h.kmer_off.back();
    for (uint64_t j = 0; j < n_states; ++j) {
      const KmerState& ks = h.kmer_states[j];
      const uint32_t head[3] = {ks.lo, ks.hi, ks.counts};
      mix(head, sizeof(head));
      mix(h.kmer_paths.data() + ks.path_off, 4 * (size_t)(2 * (ks.counts & 0xFFFFu) + (ks.counts >> 16)));
    }
  }
  MIX(h.seed_state); MIX(h.site_rec); MIX(h.apos);
#undef MIX
  mix(h.c_base, sizeof(h.c_base));
  return d;
}

int emu_index_check(void* ev) {
  auto* e = (Emu*)ev;
  const HostIndex& h = e->h;
  try {
    auto fail = [](const std::string& m) { throw std::runtime_error("index invariant: " + m); };
    const uint32_t L = (uint32_t)h.prg.size(), n = h.n;
    if (n != L + 1 || h.sa.size() != n || h.isa.size() != n) fail("sizes");
    for (uint32_t i = 0; i < n; ++i)
      if (h.isa[h.sa[i]] != i) fail("isa[sa[i]] != i");
    // text groups: codes, marker flags, relative marker ranks
    uint32_t tm = 0, tsup = 0;
    std::vector<uint32_t> text_rank(L, 0);
    for (uint32_t q = 0; q < L; ++q) {
      if ((q & ((1u << kTextSuperShift) - 1)) == 0) {
        tsup = tm;
        if (h.text_super[q >> kTextSuperShift] != tm) fail("text_super");
      }
      const TextGrp& g = h.text_grp[q >> 4];
      if ((q & 15u) == 0 && (g.info >> 16) != tm - tsup) fail("relative marker rank");
      const bool mk = h.prg[q] > 4;
      if (((g.info >> (q & 15u)) & 1u) != (mk ? 1u : 0u)) fail("marker flag");
      if (!mk && ((g.codes >> (2 * (q & 15u))) & 3u) != h.prg[q] - 1) fail("base code");
      if (mk) text_rank[q] = tm++;
    }
    // jump records: BWT order and text order hold the same record for the same marker
    uint32_t mr = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t p = h.sa[i];
      if (p == 0 || h.prg[p - 1] <= 4) continue;
      for (int w = 0; w < 8; ++w)
        if (h.marker_hit[8 * (size_t)mr + w] != h.tmarker_hit[8 * (size_t)text_rank[p - 1] + w]) fail("tmarker_hit");
      const uint32_t jlo = h.marker_hit[8 * (size_t)mr + 2], jhi = h.marker_hit[8 * (size_t)mr + 3];
      const uint32_t pj = h.marker_hit[8 * (size_t)mr + 6];
      if (jlo != kNoAllele && jlo == jhi ? pj != h.sa[jlo] : pj != kNoAllele) fail("p_jump");
      ++mr;
    }
    if (mr != tm) fail("marker count");
    // seed view of the k-mer index
    const uint64_t nk = 1ull << (2 * h.k);
    const uint32_t sd_d = seed_bucket_bases(h.k), sd_B = seed_buckets(h.k);
    if (h.seed_off.size() != nk * sd_B + 1) fail("seed_off size");
    for (uint64_t c = 0; c < nk; ++c) {
      // expected entries of the k-mer (state order), then the stored ones bucket by bucket: same multiset, every
      // entry in the bucket of its first d context bases (bucket 0: wide states / shorter contexts)
      std::multiset<std::vector<uint32_t>> expect, got;
      for (uint32_t j = h.kmer_off[c]; j < h.kmer_off[c + 1]; ++j) {
        const KmerState& ks = h.kmer_states[j];
        if (ks.hi - ks.lo + 1 > kSplitWidth) {
          expect.insert({ks.lo, ks.hi, j, 0});
          continue;
        }
        for (uint32_t i = ks.lo; i <= ks.hi; ++i) {
          const uint32_t p = h.sa[i];
          uint32_t nctx = 0, ctx = 0;
          while (nctx < kSeedCtxBases && nctx < p && h.prg[p - 1 - nctx] <= 4) {
            ctx |= (h.prg[p - 1 - nctx] - 1) << (22 - 2 * nctx);
            ++nctx;
          }
          const uint32_t bucket = nctx >= sd_d ? 1u + (sd_d ? ctx >> (24 - 2 * sd_d) : 0u) : 0u;
          expect.insert({p, 0x80000000u | (nctx << 24) | ctx, j, bucket});
        }
      }
      for (uint32_t q = 0; q < sd_B; ++q)
        for (uint32_t ent = h.seed_off[c * sd_B + q]; ent < h.seed_off[c * sd_B + q + 1]; ++ent)
          got.insert({h.seed_ent[ent].key, h.seed_ent[ent].aux, h.seed_state[ent], q});
      if (expect != got) fail("seed view entries of a k-mer");
      // presence set
      const bool present = h.kmer_off[c + 1] > h.kmer_off[c];
      if ((((h.kmer_bits[c >> 5] >> (c & 31)) & 1u) != 0) != present) fail("kmer_bits");
    }
    for (const Node& nd : h.nodes)
      if (nd.n_edges && nd.next0 != h.edges[nd.edge_off]) fail("next0");
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

uint64_t emu_reruns(void* ev) { return ((Emu*)ev)->reruns; }
void emu_status(void* ev, uint8_t* out) {
  auto* e = (Emu*)ev;
  std::memcpy(out, e->status.data(), e->status.size());
}
uint64_t emu_states_size(void* ev) {
  auto* e = (Emu*)ev;
  uint64_t t = 0;
  for (size_t i = 0; i < e->st_words.size(); ++i)
    if (e->status[i] == ST_MAPPED) t += e->st_words[i];
  return t;
}
void emu_states(void* ev, uint64_t* off, uint32_t* count, uint32_t* words) {
  auto* e = (Emu*)ev;
  uint64_t t = 0;
  for (size_t i = 0; i < e->st_words.size(); ++i) {
    off[i] = t;
    bool m = e->status[i] == ST_MAPPED;
    count[i] = m ? e->st_count[i] : 0;
    if (m) {
      std::memcpy(words + t, e->pool.data() + e->st_off[i], (size_t)e->st_words[i] * 4);
      t += e->st_words[i];
    }
  }
  off[e->st_words.size()] = t;
}
void emu_allele_sum(void* ev, uint16_t* out) {
  auto* e = (Emu*)ev;
  uint64_t na = e->h.allele_off.back();
  for (uint64_t i = 0; i < na; ++i) out[i] = (uint16_t)(e->counters[i] & 0xFFFF);
}
void emu_per_base(void* ev, uint16_t* out) {
  auto* e = (Emu*)ev;
  uint64_t na = e->h.allele_off.back();
  for (uint64_t i = 0; i < e->h.n_per_base; ++i) out[i] = (uint16_t)std::min<uint32_t>(e->counters[2 * na + i], 65535u);
}
uint64_t emu_grouped(void* ev, uint32_t* words) {
  auto* e = (Emu*)ev;
  const HostIndex& h = e->h;
  uint64_t na = h.allele_off.back();
  std::map<std::vector<uint32_t>, uint64_t> g;
  for (uint32_t s = 0; s < h.n_slots; ++s)
    for (uint32_t a = 0; a < h.n_alleles[s]; ++a)
      if (e->counters[na + h.allele_off[s] + a]) g[{s, a}] += e->counters[na + h.allele_off[s] + a];
  for (size_t i = 0; i < e->gtab.size(); ++i) {
    if (!e->gtab[i] || !e->gcount[i]) continue;
    const uint32_t* rec = e->gpool.data() + (e->gtab[i] - 1);
    std::vector<uint32_t> key{rec[0]};
    key.insert(key.end(), rec + 2, rec + 2 + rec[1]);
    g[key] += e->gcount[i];
  }
  uint64_t t = 0;
  for (auto& kv : g) {
    if (words) {
      words[t] = kv.first[0];
      words[t + 1] = (uint32_t)(kv.second & 0xFFFF);
      words[t + 2] = (uint32_t)kv.first.size() - 1;
      for (size_t i = 1; i < kv.first.size(); ++i) words[t + 2 + i] = kv.first[i];
    }
    t += 2 + kv.first.size();
  }
  return t;
}
// raw u32 accumulators [allele_sum | grouped_single | per_base] (what gq_coverage_device_ptrs exposes)
void emu_counters_raw(void* ev, uint32_t* out) {
  auto* e = (Emu*)ev;
  uint64_t na = e->h.allele_off.back();
  std::memcpy(out, e->counters.data(), (2 * na + e->h.n_per_base) * 4);
}
// multi-allele groups only, raw counts: [slot, count, n, alleles...] (gq_coverage_groups_export)
uint64_t emu_groups_raw(void* ev, uint32_t* words) {
  auto* e = (Emu*)ev;
  uint64_t t = 0;
  for (size_t i = 0; i < e->gtab.size(); ++i) {
    if (!e->gtab[i] || !e->gcount[i]) continue;
    const uint32_t* rec = e->gpool.data() + (e->gtab[i] - 1);
    if (words) {
      words[t] = rec[0];
      words[t + 1] = e->gcount[i];
      words[t + 2] = rec[1];
      for (uint32_t j = 0; j < rec[1]; ++j) words[t + 3 + j] = rec[2 + j];
    }
    t += 3 + rec[1];
  }
  return t;
}
void emu_stats(void* ev, uint64_t out[5]) {
  auto* e = (Emu*)ev;
  for (int i = 0; i < 5; ++i) out[i] = e->stats[i];
}
}
