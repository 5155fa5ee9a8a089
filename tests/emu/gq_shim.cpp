// TEST-ONLY: libgq_shim.so, preloaded (LD_PRELOAD) under the `gram` executable so that its HOST side — argument
// handling, sequence-file reader, batch queue, packer hand-over, coverage dumps, read statistics, the genotyping step,
// `gram build`'s files — runs to the end on a machine without a GPU (tests/test_gram_cli_host.py). The device-bound
// entry points of include/gq.h are answered by the host emulation of the device functions (emu_capi.cpp, the same
// code tests/test_host_parity.py holds against the oracle); everything else (gq_pack_ascii, gq_packed_words,
// gq_level_genotype, gq_read_depth_stats_host …) still comes from the real libgq.so. Never shipped, never linked into
// the product: the product has no CPU path.
#include <dlfcn.h>

#include "../../include/gq.h"
#include "emu_capi.cpp"

namespace {
thread_local std::string shim_err;
int fail(const std::string& what) {
  shim_err = what;
  return -1;
}
Emu* E(const gq_index* idx) { return (Emu*)idx; }
template <typename F>
F real(const char* name) {
  return (F)dlsym(RTLD_NEXT, name);
}
}  // namespace

extern "C" {

const char* gq_last_error(void) {
  if (!shim_err.empty()) return shim_err.c_str();
  auto f = real<const char* (*)(void)>("gq_last_error");
  return f ? f() : "";
}
int gq_device_count(int* n) {
  *n = 1;
  return 0;
}
int gq_host_alloc(uint64_t bytes, void** out) {
  *out = std::malloc(bytes ? bytes : 1);
  return *out ? 0 : fail("out of memory");
}
int gq_host_free(void* p) {
  std::free(p);
  return 0;
}
int gq_index_build(const uint32_t* prg, uint64_t n, uint32_t k, int, gq_index** out) {
  *out = (gq_index*)emu_new(prg, n, k);
  return *out ? 0 : fail(emu_last_error());
}
int gq_index_build_from_gram_dir(const uint32_t* prg, uint64_t n, uint32_t k, int, const char* dir, gq_index** out) {
  *out = (gq_index*)emu_new_from(prg, n, k, dir);
  return *out ? 0 : fail(emu_last_error());
}
int gq_index_load(const char* path, int, gq_index** out) {
  *out = (gq_index*)emu_load(path);
  return *out ? 0 : fail(emu_last_error());
}
int gq_index_save(const gq_index* idx, const char* path) {
  return emu_index_save((void*)idx, path) == 0 ? 0 : fail(emu_last_error());
}
int gq_kmer_index_dump(const gq_index* idx, const char* dir) {
  return emu_kmer_index_dump((void*)idx, dir) == 0 ? 0 : fail(emu_last_error());
}
int gq_index_prg(const gq_index* idx, uint32_t* prg_out, uint64_t* n) {
  const auto& prg = E(idx)->h.prg;
  *n = prg.size();
  if (prg_out) std::copy(prg.begin(), prg.end(), prg_out);
  return 0;
}
int gq_index_destroy(gq_index* idx) {
  emu_free(idx);
  return 0;
}
int gq_index_describe(const gq_index* idx, gq_layout* out) {
  const HostIndex& h = E(idx)->h;
  std::memset(out, 0, sizeof *out);
  out->n_symbols = h.prg.size();
  out->sa_size = h.n;
  out->kmer_size = h.k;
  out->n_sites = h.n_sites;
  out->n_site_slots = h.n_slots;
  out->is_nested = h.is_nested;
  out->n_alleles = h.allele_off.back();
  out->n_per_base = h.n_per_base;
  out->n_kmer_states = h.kmer_off.back();
  return 0;
}
int gq_index_allele_offsets(const gq_index* idx, uint64_t* allele_off) {
  const HostIndex& h = E(idx)->h;
  for (size_t i = 0; i < h.allele_off.size(); ++i) allele_off[i] = h.allele_off[i];
  return 0;
}
int gq_index_per_base_layout(const gq_index* idx, uint64_t* off_len) {
  // non-nested PRGs: the bases of a site's alleles follow one another in the PRG and in the flat per-base vector
  const HostIndex& h = E(idx)->h;
  uint64_t pb = 0, a = 0, run = 0;
  bool in_site = false;
  for (uint32_t m : h.prg) {
    if (m <= 4) {
      if (in_site) ++run;
    } else if (m & 1u) {
      in_site = true;
      run = 0;
    } else {
      off_len[2 * a] = pb;
      off_len[2 * a + 1] = run;
      pb += run;
      run = 0;
      ++a;
    }
  }
  // a site-end marker closes its last allele and is followed by bases outside the site
  // (in_site is left set: the next even marker can only come after the next odd one)
  return 0;
}
int gq_map_batch_packed(gq_index* idx, const uint32_t* packed, const uint32_t* word_off, const uint32_t* len, uint64_t n,
                        const uint32_t* seeds) {
  std::vector<uint8_t> bases;
  std::vector<uint64_t> off{0};
  for (uint64_t r = 0; r < n; ++r) {
    for (uint32_t j = 0; j < len[r]; ++j) bases.push_back((uint8_t)(1 + ((packed[word_off[r] + (j >> 4)] >> (2 * (j & 15))) & 3u)));
    off.push_back(bases.size());
  }
  if (bases.empty()) bases.push_back(0);
  return emu_map(idx, bases.data(), off.data(), n, seeds, 256) == 0 ? 0 : fail(emu_last_error());
}
int gq_coverage_fetch(gq_index* idx, uint16_t* allele_sum, uint16_t* per_base, uint64_t stats[5]) {
  if (allele_sum) emu_allele_sum(idx, allele_sum);
  if (per_base) emu_per_base(idx, per_base);
  if (stats) emu_stats(idx, stats);
  return 0;
}
int gq_coverage_grouped(gq_index* idx, uint32_t* words, uint64_t* n_words) {
  *n_words = emu_grouped(idx, words);
  return 0;
}
int gq_read_depth_stats(gq_index* idx, double out[2], uint64_t counts[2]) {
  const HostIndex& h = E(idx)->h;
  std::vector<uint16_t> pb(h.n_per_base ? h.n_per_base : 1);
  emu_per_base(idx, pb.data());
  std::vector<uint32_t> g(emu_grouped(idx, nullptr) + 1);
  const uint64_t ng = emu_grouped(idx, g.data());
  auto f = real<int (*)(const uint32_t*, uint64_t, const uint16_t*, uint64_t, const uint32_t*, uint64_t, double*, uint64_t*)>(
      "gq_read_depth_stats_host");
  if (!f) return fail("libgq.so is not loaded behind the shim");
  shim_err.clear();
  return f(h.prg.data(), h.prg.size(), pb.data(), h.n_per_base, g.data(), ng, out, counts);
}

}  // extern "C"
