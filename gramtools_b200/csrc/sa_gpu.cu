// sa_gpu.cu — suffix array of the encoded PRG on the GPU (set-up, not the mapping path).
//
// Replaces `sdsl::construct(fm_index, prg, cfg, 4)` (libgramtools/src/prg/make_data_structures.cpp:9-33) for the
// product path: the host SA-IS of sais.hpp stays as the builder of the GPU-less test emulation and as the
// known-answer check of this one (tests/test_gpu_parity.py::test_gpu_suffix_array).
//
// Prefix doubling. rank_h[i] = dense rank of the h-symbol prefix of suffix i. One round: key[i] = (rank_h[i],
// rank_h[i+h] + 1 or 0 past the end) packed into 2B bits of a 64-bit word (B = bits of n), one radix sort of the
// (key, i) pairs (cub::DeviceRadixSort over exactly those 2B bits), one flag + inclusive-sum pass for rank_2h, one
// scatter back to text order. The sentinel is the unique minimum, so ranks become distinct after at most
// ceil(log2(longest repeat)) + 1 rounds: 6 for the random-reference configurations, ~log2 n for repeat-rich PRGs.
// Memory: 28 bytes per symbol (two key buffers, two index buffers, the ranks) — 7.6 GB at config 4 (270 M symbols),
// 92 GB at config 5 (3.3e9 symbols): one B200 holds the whole-genome construction. All counts are 64-bit; positions
// are 32-bit unsigned words, as everywhere in the index (n < 2^32 - 1).
// The sort is library code (CUB, shipped with the toolkit) — index construction, never on the timed path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "index_build.hpp"

namespace gq {
namespace {

#define SA_OK(x)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("sa_gpu: ") + cudaGetErrorString(e_)); \
  } while (0)

__global__ void sa_keys_kernel(const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, uint32_t B,
                               uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r2 = (h && i + h < n) ? (uint64_t)rank[i + h] + 1 : 0;
    key[i] = h ? (((uint64_t)rank[i] << B) | r2) : (uint64_t)rank[i];
    idx[i] = (uint32_t)i;
  }
}

__global__ void sa_flags_kernel(const uint64_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ flag) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
    flag[j] = (j && key[j] != key[j - 1]) ? 1u : 0u;
}

__global__ void sa_scatter_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ rank_sorted, uint64_t n,
                                  uint32_t* __restrict__ rank) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
    rank[idx[j]] = rank_sorted[j];
}

struct Dev {
  void* p = nullptr;
  ~Dev() {
    if (p) cudaFree(p);
  }
  void alloc(size_t bytes) { SA_OK(cudaMalloc(&p, bytes ? bytes : 1)); }
};

}  // namespace

std::vector<uint32_t> gpu_suffix_array(const std::vector<int32_t>& text, int32_t sigma, int device, int* rounds_out) {
  const uint64_t n = text.size();
  if (n == 0) return {};
  if (n >= 0xFFFFFFFFull) throw std::runtime_error("sa_gpu: text positions are 32-bit words");
  SA_OK(cudaSetDevice(device));
  uint32_t B = 1;
  while ((1ull << B) <= n) ++B;  // rank + 1 <= n fits B bits
  uint32_t Bs = 1;
  while ((1ull << Bs) < (uint64_t)sigma) ++Bs;
  Dev key_a, key_b, idx_a, idx_b, rank;
  key_a.alloc(8 * n);
  key_b.alloc(8 * n);
  idx_a.alloc(4 * n);
  idx_b.alloc(4 * n);
  rank.alloc(4 * n);
  SA_OK(cudaMemcpy(rank.p, text.data(), 4 * n, cudaMemcpyHostToDevice));  // round 0: the symbols are the ranks
  size_t tmp_sort = 0, tmp_scan = 0;
  SA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)key_a.p, (uint64_t*)key_b.p,
                                        (const uint32_t*)idx_a.p, (uint32_t*)idx_b.p, n, 0, 64));
  SA_OK(cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (const uint32_t*)key_a.p, (uint32_t*)key_a.p, n));
  Dev tmp;
  tmp.alloc(tmp_sort > tmp_scan ? tmp_sort : tmp_scan);
  const uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 255) / 256, 148ull * 16);
  int rounds = 0;
  for (uint64_t h = 0;; h = h ? 2 * h : 1) {
    ++rounds;
    sa_keys_kernel<<<blocks, 256>>>((const uint32_t*)rank.p, n, h, B, (uint64_t*)key_a.p, (uint32_t*)idx_a.p);
    const int end_bit = h ? (int)(2 * B) : (int)Bs;
    size_t tb = tmp_sort;
    SA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, (const uint64_t*)key_a.p, (uint64_t*)key_b.p, (const uint32_t*)idx_a.p,
                                          (uint32_t*)idx_b.p, n, 0, end_bit));
    // key_a is free now: its first half takes the flags, its second half the scanned ranks
    uint32_t* flag = (uint32_t*)key_a.p;
    uint32_t* rs = flag + n;
    sa_flags_kernel<<<blocks, 256>>>((const uint64_t*)key_b.p, n, flag);
    tb = tmp_scan;
    SA_OK(cub::DeviceScan::InclusiveSum(tmp.p, tb, (const uint32_t*)flag, rs, n));
    sa_scatter_kernel<<<blocks, 256>>>((const uint32_t*)idx_b.p, rs, n, (uint32_t*)rank.p);
    uint32_t last = 0;
    SA_OK(cudaMemcpy(&last, rs + (n - 1), 4, cudaMemcpyDeviceToHost));
    if ((uint64_t)last + 1 == n) break;  // all ranks distinct: idx_b is the suffix array
    if (h >= n) throw std::runtime_error("sa_gpu: ranks did not separate (the text must end in a unique minimum)");
  }
  SA_OK(cudaGetLastError());
  std::vector<uint32_t> sa(n);
  SA_OK(cudaMemcpy(sa.data(), idx_b.p, 4 * n, cudaMemcpyDeviceToHost));
  if (rounds_out) *rounds_out = rounds;
  return sa;
}

}  // namespace gq
