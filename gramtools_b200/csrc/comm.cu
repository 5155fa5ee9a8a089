// comm.cu — the one exchange step of the path: reads are sharded over the GPUs, the index is
// replicated, and at the end the additive coverage counters are summed over all ranks.
//
//   * dense accumulators (allele_sum | grouped singles | per-base, one contiguous u32 allocation) and the five
//     read counters: ncclAllReduce(sum), in place, over NVLink / NVSwitch;
//   * sparse multi-allele groups: every rank exports its (site, allele set) -> count records on the device,
//     the records are all-gathered, and each rank merges the other ranks' records into its own table.
// The reference does this with OpenMP atomics / a critical section on one shared Coverage object
// (quasimap.cpp:90-118, grouped_allele_counts.cpp:17-49); its uint16 semantics are applied after the reduction
// (gq_coverage_fetch), which is exact because every increment is +1 and additions commute.
//
// NCCL is loaded at run time (dlopen) so that libgq.so itself has no link-time dependency on it: a process that
// already holds a libnccl.so.2 (PyTorch) shares that copy, a plain C++ caller (`gram`) gets the system one.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "../../include/gq.h"
#include "gq_handle.hpp"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    auto sym = [&](const char* n) { return dlsym(api.lib, n); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllReduce ||
      !api.AllGather || !api.GroupStart || !api.GroupEnd)
    throw std::runtime_error("NCCL (libnccl.so.2) could not be loaded: multi-GPU entry points are unavailable");
  return api;
}

#define NCCL_OK(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess)                                                                                 \
      throw std::runtime_error(std::string(#expr) + ": " +                                                 \
                               (nccl().GetErrorString ? nccl().GetErrorString(_r) : "NCCL error"));       \
  } while (0)

static_assert(sizeof(ncclUniqueId) == GQ_COMM_ID_BYTES, "gq.h: GQ_COMM_ID_BYTES must match ncclUniqueId");

// The reduction for a set of handles that live in THIS process (one per GPU; a single handle in the usual
// one-process-per-GPU set-up). Every NCCL call on several communicators of one process sits inside a group.
void allreduce(gq_index* const* hs, int n_local) {
  NcclApi& api = nccl();
  for (int i = 0; i < n_local; ++i)
    if (!hs[i] || !hs[i]->comm) throw std::runtime_error("gq_coverage_allreduce: no communicator on the handle (gq_comm_init)");
  const int n_ranks = hs[0]->comm_ranks;
  // ---- 1. sparse groups exported (table scan; a handful of microseconds when there are none), then ONE NCCL group:
  //         dense counters + stats summed in place, the ranks' record / word counts gathered ----
  for (int i = 0; i < n_local; ++i) {
    gq_index* ix = hs[i];
    CUDA_OK(cudaSetDevice(ix->device));
    // bounds from the pool's capacity: a record has one word more than its pool record, a multi-allele pool
    // record has at least 4 words
    ix->x_words.reserve(ix->gpool.cap + ix->gpool.cap / 4 + 4);
    ix->x_off.reserve(ix->gpool.cap / 4 + 4);
    ix->x_cnt.reserve(2);
    ix->x_cnt_all.reserve(2 * (size_t)n_ranks);
    if (!ix->x_counts_host) CUDA_OK(cudaHostAlloc((void**)&ix->x_counts_host, 8 * (size_t)n_ranks, cudaHostAllocDefault));
    CUDA_OK(cudaMemsetAsync(ix->x_cnt.p, 0, 8, ix->stream));
    gq::launch_groups_export(cov_view(ix), ix->x_words.p, (uint32_t)ix->x_words.cap, ix->x_off.p, (uint32_t)ix->x_off.cap,
                             ix->x_cnt.p, ix->stream);
  }
  NCCL_OK(api.GroupStart());
  for (int i = 0; i < n_local; ++i) {
    gq_index* ix = hs[i];
    const size_t n_cnt = 2 * ix->n_alleles + ix->n_per_base;
    if (n_cnt) NCCL_OK(api.AllReduce(ix->counters.p, ix->counters.p, n_cnt, ncclUint32, ncclSum, (ncclComm_t)ix->comm, ix->stream));
    NCCL_OK(api.AllReduce(ix->stats.p, ix->stats.p, 5, ncclUint64, ncclSum, (ncclComm_t)ix->comm, ix->stream));
    NCCL_OK(api.AllGather(ix->x_cnt.p, ix->x_cnt_all.p, 2, ncclUint32, (ncclComm_t)ix->comm, ix->stream));
  }
  NCCL_OK(api.GroupEnd());
  for (int i = 0; i < n_local; ++i) {
    CUDA_OK(cudaSetDevice(hs[i]->device));
    CUDA_OK(cudaMemcpyAsync(hs[i]->x_counts_host, hs[i]->x_cnt_all.p, 8 * (size_t)n_ranks, cudaMemcpyDeviceToHost, hs[i]->stream));
  }
  for (int i = 0; i < n_local; ++i) {
    CUDA_OK(cudaSetDevice(hs[i]->device));
    CUDA_OK(cudaStreamSynchronize(hs[i]->stream));
  }
  const uint32_t* counts = hs[0]->x_counts_host;  // identical on every rank
  uint32_t max_recs = 0, max_words = 0;
  uint64_t total_recs = 0, total_words = 0;
  for (int r = 0; r < n_ranks; ++r) {
    max_recs = std::max(max_recs, counts[2 * r]);
    max_words = std::max(max_words, counts[2 * r + 1]);
    total_recs += counts[2 * r];
    total_words += counts[2 * r + 1];
  }
  if (total_recs == 0) return;  // no multi-allele group anywhere (e.g. a SNP-only PRG): done
  // ---- 2. sparse groups: records gathered at a common stride, every rank merges the other ranks' into its table ----
  for (int i = 0; i < n_local; ++i) {
    gq_index* ix = hs[i];
    CUDA_OK(cudaSetDevice(ix->device));
    if (ix->x_words.cap < max_words || ix->x_off.cap < max_recs) {  // send buffers padded to the largest rank's size
      DevBuf<uint32_t> w, o;
      w.reserve(max_words);
      o.reserve(max_recs);
      const uint32_t my = (uint32_t)ix->comm_rank;
      if (counts[2 * my + 1]) CUDA_OK(cudaMemcpy(w.p, ix->x_words.p, (size_t)counts[2 * my + 1] * 4, cudaMemcpyDeviceToDevice));
      if (counts[2 * my]) CUDA_OK(cudaMemcpy(o.p, ix->x_off.p, (size_t)counts[2 * my] * 4, cudaMemcpyDeviceToDevice));
      ix->x_words.release();
      ix->x_off.release();
      ix->x_words = w;
      ix->x_off = o;
    }
    ix->x_all_words.reserve((size_t)max_words * n_ranks);
    ix->x_all_off.reserve((size_t)max_recs * n_ranks);
    // the merged table must hold every rank's groups without running full: pre-size it (keeps its content)
    if (ix->gtab.cap < 4 * total_recs || ix->gpool.cap < 2 * total_words + 16) {
      std::map<std::vector<uint32_t>, uint64_t> g;
      collect_groups(ix, g, true);
      size_t cap = ix->gtab.cap;
      while (cap < 4 * total_recs) cap <<= 1;
      rebuild_groups(ix, g, cap);
      if (ix->gpool.cap < 2 * total_words + 16) {  // rebuild sized the pool for the local groups only
        DevBuf<uint32_t> np;
        np.reserve(2 * total_words + 16);
        CUDA_OK(cudaMemcpy(np.p, ix->gpool.p, ix->gpool.cap * 4, cudaMemcpyDeviceToDevice));
        ix->gpool.release();
        ix->gpool = np;
      }
    }
  }
  NCCL_OK(api.GroupStart());
  for (int i = 0; i < n_local; ++i) {
    NCCL_OK(api.AllGather(hs[i]->x_words.p, hs[i]->x_all_words.p, max_words, ncclUint32, (ncclComm_t)hs[i]->comm, hs[i]->stream));
    NCCL_OK(api.AllGather(hs[i]->x_off.p, hs[i]->x_all_off.p, max_recs, ncclUint32, (ncclComm_t)hs[i]->comm, hs[i]->stream));
  }
  NCCL_OK(api.GroupEnd());
  for (int i = 0; i < n_local; ++i) {
    gq_index* ix = hs[i];
    CUDA_OK(cudaSetDevice(ix->device));
    gq::launch_groups_import(cov_view(ix), ix->x_all_words.p, max_words, ix->x_all_off.p, max_recs, ix->x_cnt_all.p,
                             (uint32_t)n_ranks, (uint32_t)ix->comm_rank, ix->stream);
  }
  for (int i = 0; i < n_local; ++i) {
    CUDA_OK(cudaSetDevice(hs[i]->device));
    CUDA_OK(cudaStreamSynchronize(hs[i]->stream));
    CUDA_OK(cudaGetLastError());
    uint32_t gs[2];
    CUDA_OK(cudaMemcpy(gs, hs[i]->gsmall.p, 8, cudaMemcpyDeviceToHost));
    if (gs[1] & 1u) throw std::runtime_error("gq_coverage_allreduce: group table overflow during the merge (internal error)");
  }
}

}  // namespace

#define GQ_TRY try {
#define GQ_CATCH                    \
  }                                 \
  catch (const std::exception& e) { \
    set_last_error(e.what());       \
    return -1;                      \
  }                                 \
  return 0;

extern "C" {

int gq_comm_unique_id(uint8_t id[GQ_COMM_ID_BYTES]) {
  GQ_TRY
  if (!id) throw std::runtime_error("null argument");
  ncclUniqueId u;
  NCCL_OK(nccl().GetUniqueId(&u));
  std::memcpy(id, &u, sizeof u);
  GQ_CATCH
}

int gq_comm_init(gq_index* ix, const uint8_t id[GQ_COMM_ID_BYTES], int rank, int n_ranks) {
  GQ_TRY
  if (!ix || !id) throw std::runtime_error("null argument");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::runtime_error("invalid rank");
  if (ix->comm) throw std::runtime_error("the handle already has a communicator");
  CUDA_OK(cudaSetDevice(ix->device));
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ncclComm_t c = nullptr;
  NCCL_OK(nccl().CommInitRank(&c, n_ranks, u, rank));
  ix->comm = c;
  ix->comm_rank = rank;
  ix->comm_ranks = n_ranks;
  GQ_CATCH
}

int gq_comm_init_all(gq_index** per_gpu, int n) {
  GQ_TRY
  if (!per_gpu || n < 1) throw std::runtime_error("null argument");
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!per_gpu[i]) throw std::runtime_error("null handle");
    if (per_gpu[i]->comm) throw std::runtime_error("a handle already has a communicator");
    devs[i] = per_gpu[i]->device;
  }
  std::vector<ncclComm_t> comms(n, nullptr);
  NCCL_OK(nccl().CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    per_gpu[i]->comm = comms[i];
    per_gpu[i]->comm_rank = i;
    per_gpu[i]->comm_ranks = n;
  }
  GQ_CATCH
}

int gq_comm_destroy(gq_index* ix) {
  GQ_TRY
  if (ix && ix->comm) {
    cudaSetDevice(ix->device);
    ix->x_words.release();
    ix->x_off.release();
    ix->x_cnt.release();
    ix->x_cnt_all.release();
    ix->x_all_words.release();
    ix->x_all_off.release();
    if (ix->x_counts_host) cudaFreeHost(ix->x_counts_host);
    ix->x_counts_host = nullptr;
    nccl().CommDestroy((ncclComm_t)ix->comm);
    ix->comm = nullptr;
    ix->comm_ranks = 1;
    ix->comm_rank = 0;
  }
  GQ_CATCH
}

int gq_coverage_allreduce(gq_index* ix) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  allreduce(&ix, 1);
  GQ_CATCH
}

int gq_coverage_allreduce_all(gq_index** per_gpu, int n) {
  GQ_TRY
  if (!per_gpu || n < 1) throw std::runtime_error("null argument");
  if (per_gpu[0] && per_gpu[0]->comm_ranks != n) throw std::runtime_error("the handles must be all the ranks of one gq_comm_init_all");
  allreduce(per_gpu, n);
  GQ_CATCH
}

int gq_comm_version(int* version) {
  GQ_TRY
  if (!version) throw std::runtime_error("null argument");
  *version = 0;
  if (nccl().GetVersion) NCCL_OK(nccl().GetVersion(version));
  GQ_CATCH
}

}  // extern "C"
