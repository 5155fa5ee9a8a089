// gq_handle.hpp — the handle behind `gq_index*` (include/gq.h), shared by capi.cu and comm.cu.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "index_build.hpp"
#include "kernels.cuh"

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    cap = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

struct gq_index {
  gq::HostIndex h;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<void*> index_allocs;
  size_t index_bytes = 0;
  gq::IndexView dv{};
  // coverage accumulators
  DevBuf<uint32_t> counters;  // allele_sum | grouped_single | per_base
  DevBuf<uint32_t> allele_off;
  DevBuf<uint32_t> gtab, gcount, gpool, gsmall;  // gsmall: [gpool_used, error_flags]
  DevBuf<unsigned long long> stats, stats_batch;  // totals; the counters of the batch being mapped
  uint32_t* post_host = nullptr;                  // pinned: [small 0..3 | gsmall 0..1] read back once per call
  uint64_t n_alleles = 0, n_per_base = 0;
  // batch
  DevBuf<uint8_t> bases;
  DevBuf<uint64_t> offsets;
  DevBuf<uint32_t> word_off, packed, packed_rc, len, seeds;
  uint32_t n_reads = 0;
  uint32_t total_words = 0;
  // search outputs
  DevBuf<uint8_t> status;
  DevBuf<uint32_t> st_off, st_words, st_count, pool, small;  // small: [pool_used, n_overflow, n_cov_overflow]
  DevBuf<uint32_t> overflow_list, cov_overflow_list, mapped_list, multi_list, heavy_list;
  DevBuf<uint32_t> seed_rec, surv_rec, surv_cnt, gen_list;  // seed pass (SeedOut): survivor records, per-strand counts, general list
  uint32_t seed_recs_per_read = 16;  // candidate records per read (set from the index: ~2.5 x mean suffixes per indexed k-mer, both strands); a full pool sends strands to the general kernel
  bool use_seed_pass = true;
  DevBuf<uint32_t> arena, big_arena;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t aux_stream = nullptr;  // second compute stream of the pipelined path
  cudaEvent_t aux_event = nullptr;
  cudaStream_t cls_stream = nullptr;  // early k-mer filter of the pipelined path (beside the later slices)
  cudaEvent_t cls_event[3] = {nullptr, nullptr, nullptr};
  uint32_t early_classify = 1;
  DevBuf<uint32_t> arena2;
  void* fetch_host = nullptr;  // pinned staging of gq_coverage_fetch
  size_t fetch_host_bytes = 0;
  DevBuf<uint16_t> fetch_dev;
  std::vector<cudaEvent_t> chunk_events, tl_copy;
  uint32_t chunk_reads = 1u << 18, tail_chunk_reads = 1u << 16;
  uint32_t resident_slices = 1;  // gq_map_resident: slices run on two streams
  bool overlap_classify = true;  // single-slice runs: classify_kernel beside coverage_kernel on a second stream  // slice size of the H2D / compute pipeline in gq_map_batch
  // options
  uint32_t arena_words = 512;
  uint32_t n_threads = 148 * 1280;      // search kernel lanes (5 CTAs of 256 per SM)
  uint32_t cov_threads = 148 * 1024;    // coverage kernel threads
  uint32_t big_arena_words = 1u << 16;  // overflow re-runs: 4096 lanes x 256 KB, then x4 words and 4x fewer lanes
  uint32_t big_threads = 4096;
  uint32_t pool_words_per_read = 48;
  bool super_in_smem = true;
  uint32_t rf_thresh = 8, ev_thresh = 8, leave_opt = 0, wait_opt = 0;
  // multi-GPU (comm.cu): ncclComm_t of this handle, its rank and the number of ranks
  void* comm = nullptr;
  int comm_rank = 0, comm_ranks = 1;
  DevBuf<uint32_t> x_words, x_off, x_cnt, x_cnt_all, x_all_words, x_all_off;  // exchange buffers, kept between calls
  uint32_t* x_counts_host = nullptr;                                          // pinned: 2 words per rank
  // run info
  double info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // per-kernel CUDA events of a single-slice run: before seed, after seed, after verify, after text, after the
  // general kernel, around classify (its own stream), after coverage
  cudaEvent_t kev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double kernel_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // seed, verify, text, general, classify, coverage
};


// capi.cu
gq::CoverageView cov_view(gq_index* ix);
void collect_groups(gq_index* ix, std::map<std::vector<uint32_t>, uint64_t>& out, bool multi_only);
void rebuild_groups(gq_index* ix, const std::map<std::vector<uint32_t>, uint64_t>& g, size_t min_cap);
void set_last_error(const std::string& what);
