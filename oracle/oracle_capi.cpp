// oracle_capi.cpp — C entry points over the CPU ORACLE (TEST INFRASTRUCTURE ONLY; see gq_oracle.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs. Never linked into libgq.so.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>

#include "gq_oracle.hpp"

using namespace gqo;

struct OracleHandle {
  Mapper m;
  std::vector<uint8_t> status;                 // 2 * n_reads of the last batch
  std::vector<std::vector<uint32_t>> states;   // canonical records per strand
  std::vector<uint32_t> state_count;
  double last_seconds = 0;
};

static thread_local std::string g_err;

extern "C" {

const char* gqo_last_error() { return g_err.c_str(); }

void* gqo_new(const uint32_t* prg, uint64_t n, uint32_t k) {
  try {
    auto* h = new OracleHandle();
    h->m = Mapper(std::vector<Marker>(prg, prg + n), k);
    return h;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void gqo_free(void* h) { delete (OracleHandle*)h; }

// i-th raw draws of std::mt19937(seed): the per-read selection seeds of handle_read_file
// (quasimap.cpp:136-137; 5000 drawn per batch, so read j of a file gets draw j while batches are full)
void gqo_master_seeds(uint32_t seed, uint64_t n, uint32_t* out) {
  std::mt19937 g;
  g.seed(seed);
  for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t)g();
}

void gqo_sizes(void* hv, uint64_t out[6]) {
  auto* h = (OracleHandle*)hv;
  uint64_t na = 0;
  for (auto& s : h->m.cov.allele_sum) na += s.size();
  out[0] = h->m.info.num_sites;
  out[1] = na;
  out[2] = h->m.per_base_flat().size();
  out[3] = h->m.info.graph.is_nested;
  out[4] = h->m.info.fm.size();
  uint64_t ks = 0;
  for (auto& e : h->m.kmers) ks += e.second.size();
  out[5] = ks;
}

// The k-mer index as flat records [code, lo, hi, nt, ng, (site, allele) * nt, (site, 0xFFFFFFFF) * ng], k-mers by
// ascending code (base j of the k-mer at bits [2j, 2j + 2)), the states of a k-mer in the index's own order
// (build.cpp's worklist order, pinned by test_build.cpp:404-429). words == NULL: only the size.
uint64_t gqo_kmer_states(void* hv, uint32_t* words) {
  auto* h = (OracleHandle*)hv;
  std::vector<std::pair<uint64_t, const SearchStates*>> by_code;
  for (auto& e : h->m.kmers) {
    uint64_t code = 0;
    for (size_t j = 0; j < e.first.size(); ++j) code |= (uint64_t)(e.first[j] - 1) << (2 * j);
    by_code.emplace_back(code, &e.second);
  }
  std::sort(by_code.begin(), by_code.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  uint64_t t = 0;
  for (auto& kv : by_code)
    for (const SearchState& st : *kv.second) {
      const uint32_t nt = (uint32_t)st.traversed.size(), ng = (uint32_t)st.traversing.size();
      if (words) {
        words[t] = (uint32_t)kv.first;
        words[t + 1] = (uint32_t)st.lo;
        words[t + 2] = (uint32_t)st.hi;
        words[t + 3] = nt;
        words[t + 4] = ng;
        uint64_t q = t + 5;
        for (auto& loc : st.traversed) {
          words[q++] = (uint32_t)loc.first;
          words[q++] = (uint32_t)loc.second;
        }
        for (auto& loc : st.traversing) {
          words[q++] = (uint32_t)loc.first;
          words[q++] = 0xFFFFFFFFu;
        }
      }
      t += 5 + 2 * nt + 2 * ng;
    }
  return t;
}

// handle_reads_buffer (quasimap.cpp:82-118) over one batch; threads = omp threads (1 = serial)
int gqo_map(void* hv, const uint8_t* bases, const uint64_t* off, uint64_t n_reads, const uint32_t* seeds,
            int threads, int want_states, int count_events) {
  auto* h = (OracleHandle*)hv;
  try {
    h->status.assign(2 * n_reads, 0);
    h->states.assign(want_states ? 2 * n_reads : 0, {});
    h->state_count.assign(2 * n_reads, 0);
    h->m.count_events = count_events != 0;
    if (count_events) threads = 1;
    std::string err;
    auto t0 = std::chrono::steady_clock::now();
    const bool atomic = threads > 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n_reads; ++i) {
      try {
        Sequence read(bases + off[i], bases + off[i + 1]);
        StrandStatus st[2];
        SearchStates f, r;
        h->m.quasimap_forward_reverse(read, seeds[i], atomic, st, want_states ? &f : nullptr, want_states ? &r : nullptr);
        h->status[2 * i] = (uint8_t)st[0];
        h->status[2 * i + 1] = (uint8_t)st[1];
        if (want_states) {
          h->states[2 * i] = canonical_states(f);
          h->states[2 * i + 1] = canonical_states(r);
          h->state_count[2 * i] = (uint32_t)f.size();
          h->state_count[2 * i + 1] = (uint32_t)r.size();
        }
      } catch (const std::exception& e) {
#pragma omp critical(gqo_err)
        err = e.what();
      }
    }
    h->last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!err.empty()) throw std::runtime_error(err);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// quasimap_read on the reads as given, no reverse complement: what the reference's test helper prg_setup::quasimap_reads
// does (tests/test_resources/test_resources.cpp:48-56); coverage accumulates in the handle. Used by the tests that
// transcribe reference cases built on that helper (read statistics, genotyping).
int gqo_map_forward(void* hv, const uint8_t* bases, const uint64_t* off, uint64_t n_reads, const uint32_t* seeds) {
  auto* h = (OracleHandle*)hv;
  try {
    h->status.assign(2 * n_reads, 0);
    h->states.assign(0, {});
    h->state_count.assign(2 * n_reads, 0);
    for (uint64_t i = 0; i < n_reads; ++i) {
      Sequence read(bases + off[i], bases + off[i + 1]);
      h->status[2 * i] = (uint8_t)h->m.quasimap_read(read, seeds[i], false, nullptr);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
double gqo_last_seconds(void* hv) { return ((OracleHandle*)hv)->last_seconds; }

void gqo_status(void* hv, uint8_t* out) {
  auto* h = (OracleHandle*)hv;
  std::memcpy(out, h->status.data(), h->status.size());
}
uint64_t gqo_states_size(void* hv) {
  auto* h = (OracleHandle*)hv;
  uint64_t t = 0;
  for (auto& s : h->states) t += s.size();
  return t;
}
void gqo_states(void* hv, uint64_t* off, uint32_t* count, uint32_t* words) {
  auto* h = (OracleHandle*)hv;
  uint64_t t = 0;
  for (size_t i = 0; i < h->states.size(); ++i) {
    off[i] = t;
    count[i] = h->state_count[i];
    std::memcpy(words + t, h->states[i].data(), h->states[i].size() * 4);
    t += h->states[i].size();
  }
  off[h->states.size()] = t;
}
void gqo_allele_sum(void* hv, uint16_t* out) {
  auto* h = (OracleHandle*)hv;
  size_t t = 0;
  for (auto& s : h->m.cov.allele_sum)
    for (auto c : s) out[t++] = c;
}
void gqo_per_base(void* hv, uint16_t* out) {
  auto* h = (OracleHandle*)hv;
  auto v = h->m.per_base_flat();
  std::memcpy(out, v.data(), v.size() * 2);
}
// records [site_index, count, n, allele ids...] sorted by (site, ids); words == NULL -> size only
uint64_t gqo_grouped(void* hv, uint32_t* words) {
  auto* h = (OracleHandle*)hv;
  uint64_t t = 0;
  for (size_t s = 0; s < h->m.cov.grouped.size(); ++s)
    for (auto& e : h->m.cov.grouped[s]) {
      if (words) {
        words[t] = (uint32_t)s;
        words[t + 1] = e.second;
        words[t + 2] = (uint32_t)e.first.size();
        for (size_t i = 0; i < e.first.size(); ++i) words[t + 3 + i] = (uint32_t)e.first[i];
      }
      t += 3 + e.first.size();
    }
  return t;
}
void gqo_stats(void* hv, uint64_t out[5]) {
  auto& s = ((OracleHandle*)hv)->m.stats;
  out[0] = s.all_reads;
  out[1] = s.skipped;
  out[2] = s.missing_kmer;
  out[3] = s.no_extension;
  out[4] = s.exact_mapped;
}
void gqo_events(void* hv, uint64_t out[7]) {
  auto& e = ((OracleHandle*)hv)->m.events;
  out[0] = e.q_rank;
  out[1] = e.w_marker;
  out[2] = e.q_sa;
  out[3] = e.q_node;
  out[4] = e.a_cov;
  out[5] = e.strands;
  out[6] = e.bases;
}
int gqo_max_threads() { return omp_get_max_threads(); }
}
