// capi.cu — the C ABI of libgq.so (include/gq.h) over the host index builder and the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gq.h"
#include "index_build.hpp"
#include "kernels.cuh"

namespace gq {
void debug_counters(unsigned long long* out32);
void launch_stats(const uint8_t* status, const uint32_t* len, uint32_t n_reads, unsigned long long* stats,
                  cudaStream_t st);
}

static thread_local std::string g_err;
void set_last_error(const std::string& what) { g_err = what; }

#include "gq_handle.hpp"

template <class T>
static const T* upload(gq_index* ix, const std::vector<T>& v) {
  T* d = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CUDA_OK(cudaMalloc(&d, bytes));
  if (!v.empty()) CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  ix->index_allocs.push_back(d);
  ix->index_bytes += bytes;
  return d;
}

static void upload_index(gq_index* ix) {
  const gq::HostIndex& h = ix->h;
  gq::IndexView& v = ix->dv;
  v.n = h.n;
  v.rank_blk = upload(ix, h.rank_blk);
  v.super_cnt = upload(ix, h.super_cnt);
  v.mrank_blk = upload(ix, h.mrank_blk);
  v.marker_hit = upload(ix, h.marker_hit);
  v.text_grp = upload(ix, h.text_grp);
  v.text_super = upload(ix, h.text_super);
  v.tmarker_hit = upload(ix, h.tmarker_hit);
  v.isa = upload(ix, h.isa);
  for (int i = 0; i < 4; ++i) v.c_base[i] = h.c_base[i];
  v.n_slots = h.n_slots;
  v.any_nested = h.is_nested ? 1u : 0u;
  v.site_sa = upload(ix, h.site_sa);
  v.allele_iv = upload(ix, h.allele_iv);
  v.par = upload(ix, h.par);
  v.tm_odd = upload(ix, h.tm_odd);
  v.tm_even_off = upload(ix, h.tm_even_off);
  v.tm_even = upload(ix, h.tm_even);
  v.entry_next = upload(ix, h.entry_next);
  v.site_snp = upload(ix, h.site_snp);
  v.sa = upload(ix, h.sa);
  v.pos2node = upload(ix, h.pos2node);
  v.nodes = upload(ix, h.nodes);
  v.edges = upload(ix, h.edges);
  v.site_rec = upload(ix, h.site_rec);
  v.apos = upload(ix, h.apos);
  v.k = h.k;
  v.kmer_bits = upload(ix, h.kmer_bits);
  v.kmer_off = upload(ix, h.kmer_off);
  v.kmer_states = upload(ix, h.kmer_states);
  v.seed_off = upload(ix, h.seed_off);
  v.seed_ent = upload(ix, h.seed_ent);
  v.seed_state = upload(ix, h.seed_state);
  v.kmer_paths = upload(ix, h.kmer_paths);
}

static uint32_t next_pow2(uint64_t x) {
  uint32_t p = 1;
  while (p < x && p < (1u << 30)) p <<= 1;
  return p;
}

static void alloc_coverage(gq_index* ix) {
  const gq::HostIndex& h = ix->h;
  ix->n_alleles = h.allele_off.back();
  ix->n_per_base = h.n_per_base;
  size_t nc = 2 * ix->n_alleles + ix->n_per_base;
  ix->counters.reserve(nc + 1);
  ix->allele_off.reserve(h.allele_off.size());
  CUDA_OK(cudaMemcpy(ix->allele_off.p, h.allele_off.data(), h.allele_off.size() * 4, cudaMemcpyHostToDevice));
  uint32_t cap = next_pow2(std::max<uint64_t>(1024, 4 * ix->n_alleles));
  ix->gtab.reserve(cap);
  ix->gcount.reserve(cap);
  ix->gpool.reserve((size_t)cap * 4);
  ix->gsmall.reserve(4);
  ix->stats.reserve(8);
  ix->stats_batch.reserve(8);
  if (!ix->post_host) CUDA_OK(cudaHostAlloc((void**)&ix->post_host, 64, cudaHostAllocDefault));
}

static void reset_coverage(gq_index* ix) {
  CUDA_OK(cudaMemsetAsync(ix->counters.p, 0, ix->counters.bytes(), ix->stream));
  CUDA_OK(cudaMemsetAsync(ix->gtab.p, 0, ix->gtab.bytes(), ix->stream));
  CUDA_OK(cudaMemsetAsync(ix->gcount.p, 0, ix->gcount.bytes(), ix->stream));
  CUDA_OK(cudaMemsetAsync(ix->gsmall.p, 0, ix->gsmall.bytes(), ix->stream));
  CUDA_OK(cudaMemsetAsync(ix->stats.p, 0, ix->stats.bytes(), ix->stream));
  CUDA_OK(cudaStreamSynchronize(ix->stream));
}

gq::CoverageView cov_view(gq_index* ix) {
  gq::CoverageView c{};
  c.allele_sum = ix->counters.p;
  c.grouped_single = ix->counters.p + ix->n_alleles;
  c.per_base = ix->counters.p + 2 * ix->n_alleles;
  c.gtab = ix->gtab.p;
  c.gcount = ix->gcount.p;
  c.gtab_cap = (uint32_t)ix->gtab.cap;
  c.gpool = ix->gpool.p;
  c.gpool_cap = (uint32_t)ix->gpool.cap;
  c.gpool_used = ix->gsmall.p;
  c.error_flags = ix->gsmall.p + 1;
  c.stats = ix->stats.p;
  c.allele_off = ix->allele_off.p;
  return c;
}

static void reserve_batch(gq_index* ix, uint64_t n_reads, uint64_t nb) {
  if (n_reads >= (1ull << 30)) throw std::runtime_error("batch too large (max 2^30 reads per batch)");
  uint64_t max_words = (nb >> 4) + n_reads + 2;  // read r starts at word (offset >> 4) + r
  if (max_words >= (1ull << 32)) throw std::runtime_error("batch too large (packed words exceed 2^32)");
  ix->bases.reserve(nb + 16);
  ix->offsets.reserve(n_reads + 1);
  ix->word_off.reserve(n_reads + 1);
  ix->packed.reserve(max_words);
  ix->packed_rc.reserve(max_words);
  ix->len.reserve(n_reads);
  ix->seeds.reserve(n_reads);
}

struct Chunk {
  uint32_t r0, r1;
};
constexpr size_t kMaxChunks = 64;  // per-slice counters live in fixed slots of `small`

static void grow_groups(gq_index* ix);

static std::vector<Chunk> make_chunks(gq_index* ix, uint32_t n, bool pipelined) {
  std::vector<Chunk> ch;
  if (!pipelined) {  // resident batch: equal slices (alternating between the two compute streams)
    const uint32_t k = std::max<uint32_t>(1, std::min<uint32_t>(std::min<uint32_t>(ix->resident_slices, (uint32_t)kMaxChunks), (n + 65535) / 65536));
    const uint32_t per = (n + k - 1) / k;
    for (uint32_t r = 0; r < n; r += per) ch.push_back({r, std::min(n, r + per)});
    return ch;
  }
  // Slices of chunk_reads reads, then halving towards the end: the copy engine is the bottleneck of the
  // pipelined path, so what matters is how little compute is left when the last byte has arrived.
  const uint32_t big = std::max<uint32_t>(ix->chunk_reads, (n + 23) / 24), small = std::max<uint32_t>(ix->tail_chunk_reads, 1024);
  for (uint32_t r = 0; r < n;) {
    uint32_t left = n - r;
    uint32_t per = left <= small ? left : std::max(small, std::min(big, left / 2));
    if (ch.size() + 1 == kMaxChunks) per = left;  // (only with extreme chunk options) the last slot takes the rest
    ch.push_back({r, r + per});
    r += per;
  }
  return ch;
}

// The caller's host buffers of one batch: unpacked (one byte per base, packed on the device after the copy) or
// already 2-bit packed on the host (gq_pack_reads / gq_pack_ascii: a third of the PCIe bytes, no pack kernel).
struct HostBatch {
  const uint8_t* bases = nullptr;    // unpacked form
  const uint64_t* off = nullptr;
  const uint32_t* packed = nullptr;  // packed form
  const uint32_t* word_off = nullptr;
  const uint32_t* len = nullptr;
  const uint32_t* seeds = nullptr;
  bool is_packed() const { return packed != nullptr || word_off != nullptr; }
};

// H2D of one slice of the caller's buffers, all on one stream in slice order
static void copy_chunk(gq_index* ix, const HostBatch& hb, Chunk c, cudaStream_t st) {
  cudaStream_t st_small = st;
  const size_t nr = c.r1 - c.r0;
  if (hb.is_packed()) {
    const uint32_t w0 = hb.word_off[c.r0], w1 = hb.word_off[c.r1];
    if (w1 > w0) CUDA_OK(cudaMemcpyAsync(ix->packed.p + w0, hb.packed + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(ix->word_off.p + c.r0, hb.word_off + c.r0, (nr + 1) * 4, cudaMemcpyHostToDevice, st_small));
    CUDA_OK(cudaMemcpyAsync(ix->len.p + c.r0, hb.len + c.r0, nr * 4, cudaMemcpyHostToDevice, st_small));
    CUDA_OK(cudaMemcpyAsync(ix->seeds.p + c.r0, hb.seeds + c.r0, nr * 4, cudaMemcpyHostToDevice, st_small));
    return;
  }
  const uint8_t* bases = hb.bases;
  const uint64_t* off = hb.off;
  const uint32_t* seeds = hb.seeds;
  uint64_t b0 = off[c.r0], b1 = off[c.r1];
  if (b1 > b0) CUDA_OK(cudaMemcpyAsync(ix->bases.p + b0, bases + b0, b1 - b0, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(ix->offsets.p + c.r0, off + c.r0, (size_t)(c.r1 - c.r0 + 1) * 8, cudaMemcpyHostToDevice, st_small));
  CUDA_OK(cudaMemcpyAsync(ix->seeds.p + c.r0, seeds + c.r0, (size_t)(c.r1 - c.r0) * 4, cudaMemcpyHostToDevice, st_small));
}

// H2D + 2-bit packing of one slice on one stream
static void upload_chunk(gq_index* ix, const uint8_t* bases, const uint64_t* off, const uint32_t* seeds, Chunk c,
                         cudaStream_t st) {
  HostBatch hb;
  hb.bases = bases;
  hb.off = off;
  hb.seeds = seeds;
  copy_chunk(ix, hb, c, st);
  gq::launch_pack(ix->bases.p, ix->offsets.p, c.r0, c.r1, ix->word_off.p, ix->packed.p, ix->len.p, st);
}

static void do_upload(gq_index* ix, const uint8_t* bases, const uint64_t* off, uint64_t n_reads,
                      const uint32_t* seeds) {
  CUDA_OK(cudaSetDevice(ix->device));
  if (n_reads && off[0] != 0) throw std::runtime_error("read_offsets[0] must be 0");
  uint64_t nb = n_reads ? off[n_reads] : 0;
  reserve_batch(ix, n_reads, nb);
  ix->n_reads = (uint32_t)n_reads;
  if (n_reads == 0) return;
  upload_chunk(ix, bases, off, seeds, Chunk{0, (uint32_t)n_reads}, ix->stream);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(ix->stream));  // the caller's buffers may go away after this call
  ix->info[5] = (double)(nb + (n_reads + 1) * 8 + n_reads * 4);  // H2D bytes of this upload
}

// Map the batch. host pointers != nullptr: the reads come from the caller's HOST buffers and the H2D
// copy of slice i+1 (copy stream) overlaps the kernels of slice i (compute stream).
static void do_map(gq_index* ix, const HostBatch* hb = nullptr) {
  CUDA_OK(cudaSetDevice(ix->device));
  const bool pipelined = hb != nullptr;
  const bool host_packed = hb && hb->is_packed();
  const auto t_host0 = std::chrono::steady_clock::now();  // info[7]: host time spent enqueueing the call
  const uint32_t n = ix->n_reads;
  for (int i = 0; i < 8; ++i)
    if (i != 5) ix->info[i] = 0;
  if (n == 0) return;
  cudaStream_t st = ix->stream;
  ix->status.reserve(2 * (size_t)n);
  ix->st_off.reserve(2 * (size_t)n);
  ix->st_words.reserve(2 * (size_t)n);
  ix->st_count.reserve(2 * (size_t)n);
  ix->overflow_list.reserve(4 * (size_t)n + 16);  // a strand can be flagged by the text kernel and again by the general one
  ix->cov_overflow_list.reserve(2 * (size_t)n);
  // one entry per mapped strand, plus a second one for a strand whose final state overflowed the pool in the text
  // kernel (its slot stays in the list, the re-run appends it again)
  ix->mapped_list.reserve(4 * (size_t)n + 16);
  if (ix->h.is_nested) {  // strands with several final states (coverage, later passes)
    ix->multi_list.reserve(2 * (size_t)n);
    ix->heavy_list.reserve(2 * (size_t)n);
  }
  ix->small.reserve(8 + 8 * kMaxChunks);
  ix->surv_cnt.reserve(2 * (size_t)n);
  ix->gen_list.reserve(2 * (size_t)n);
  ix->seed_rec.reserve(4 * std::max<size_t>((size_t)n * ix->seed_recs_per_read, 1 << 16));
  ix->surv_rec.reserve(4 * std::max<size_t>((size_t)n * ix->seed_recs_per_read, 1 << 16));
  size_t pool_need = std::max<size_t>((size_t)n * ix->pool_words_per_read, 1 << 16);
  pool_need = std::min<size_t>(pool_need, 0xFFFFFFF0ull);
  ix->pool.reserve(pool_need);
  std::vector<Chunk> chunks = make_chunks(ix, n, pipelined);
  uint32_t max_chunk = 0;
  for (auto& c : chunks) max_chunk = std::max(max_chunk, c.r1 - c.r0);
  // persistent lanes: never more than one lane per 4 strands, so the refill/batching steady state exists
  uint32_t threads = std::min<uint32_t>(ix->n_threads, std::max<uint32_t>(256, ((2 * max_chunk / 4 + 255) / 256) * 256));
  uint32_t threads2 = std::min<uint32_t>(ix->cov_threads, ((2 * max_chunk + 255) / 256) * 256);
  ix->arena.reserve((size_t)std::max(threads, threads2) * ix->arena_words);
  // small: [0] pool_used [1] n_overflow [2] n_cov_overflow;
  // per chunk c: [8+4c] n_mapped [9+4c] work counter [10+4c] survivor records [11+4c] n_gen (general-kernel work list)
  CUDA_OK(cudaMemsetAsync(ix->small.p, 0, (8 + 8 * kMaxChunks) * 4, st));

  gq::BatchView b{ix->packed.p, ix->packed_rc.p, ix->word_off.p, ix->len.p, ix->seeds.p, n, 0, n};
  gq::SearchOut o{ix->status.p, ix->st_off.p, ix->st_words.p, ix->st_count.p, ix->pool.p, (uint32_t)ix->pool.cap,
                  ix->small.p,  ix->overflow_list.p, ix->small.p + 1, ix->mapped_list.p, ix->small.p + 8,
                  ix->small.p + 9, ix->use_seed_pass ? ix->surv_cnt.p : nullptr};
  gq::CoverageView c = cov_view(ix);
  int launches = 0;
  if (pipelined) {
    if (!ix->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
    while (ix->chunk_events.size() < chunks.size()) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ix->chunk_events.push_back(e);
    }
    // the copy streams must not start overwriting buffers the compute stream may still be reading
    CUDA_OK(cudaEventRecord(ix->ev[3], st));
    CUDA_OK(cudaStreamWaitEvent(ix->copy_stream, ix->ev[3], 0));
    // copies only on the copy streams, back to back; packing runs on the compute stream
    for (size_t i = 0; i < chunks.size(); ++i) {
      // one stream, in slice order: with the small per-read arrays on a second stream the copy engine served that
      // stream only after ALL the big copies (measured: the first slice's kernels started when the last byte of the
      // batch had arrived)
      copy_chunk(ix, *hb, chunks[i], ix->copy_stream);
      CUDA_OK(cudaEventRecord(ix->chunk_events[i], ix->copy_stream));
      if (getenv("GQ_TIMELINE")) {
        cudaEvent_t e;
        CUDA_OK(cudaEventCreate(&e));
        CUDA_OK(cudaEventRecord(e, ix->copy_stream));
        ix->tl_copy.push_back(e);
      }
    }
  }
  // developer aid (GQ_TIMELINE=1): when each slice's copy ended and its kernels started / ended, relative to ev[0]
  static const bool timeline = getenv("GQ_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl;
  if (timeline && pipelined)
    for (size_t i = 0; i < 4 * chunks.size(); ++i) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreate(&e));
      tl.push_back(e);
    }
  CUDA_OK(cudaEventRecord(ix->ev[0], st));
  // Slices alternate between two compute streams (each with its own arena), so the kernels of slice i+1
  // fill the GPU while slice i's are in their latency-bound tails; the caller's stream joins at the end.
  const bool two_streams = chunks.size() > 1;
  if (two_streams) {
    if (!ix->aux_stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&ix->aux_stream, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&ix->aux_event, cudaEventDisableTiming));
    }
    ix->arena2.reserve((size_t)std::max(threads, threads2) * ix->arena_words);
    if (!pipelined) CUDA_OK(cudaEventRecord(ix->ev[3], st));
    CUDA_OK(cudaStreamWaitEvent(ix->aux_stream, ix->ev[3], 0));  // after the counters' memset
  }
  // pipelined batches of four or more slices: the last slice but two ends the early k-mer filter's share
  const size_t early_upto = (pipelined && ix->early_classify && chunks.size() >= 4) ? chunks.size() - 3 : SIZE_MAX;
  for (size_t i = 0; i < chunks.size(); ++i) {
    cudaStream_t cs = (two_streams && (i & 1)) ? ix->aux_stream : st;
    uint32_t* arena = (two_streams && (i & 1)) ? ix->arena2.p : ix->arena.p;
    if (pipelined) {
      CUDA_OK(cudaStreamWaitEvent(cs, ix->chunk_events[i], 0));
      if (!tl.empty()) CUDA_OK(cudaEventRecord(tl[4 * i + 1], cs));
      if (!host_packed) {
        gq::launch_pack(ix->bases.p, ix->offsets.p, chunks[i].r0, chunks[i].r1, ix->word_off.p, ix->packed.p, ix->len.p, cs);
        ++launches;
      }
    }
    gq::BatchView bc = b;
    bc.read_begin = chunks[i].r0;
    bc.read_end = chunks[i].r1;
    gq::launch_revcomp(bc, ix->packed_rc.p, cs);  // the slice's reverse strands
    ++launches;
    gq::SearchOut oc = o;
    oc.mapped_list = ix->mapped_list.p + 2 * (size_t)chunks[i].r0;
    oc.n_mapped = ix->small.p + 8 + 4 * i;
    oc.work_counter = ix->small.p + 9 + 4 * i;
    if (ix->use_seed_pass) {
      // seed pass -> verify + text kernels (candidates walked through the PRG text) -> general kernel (the rest)
      uint32_t* gen_list = ix->gen_list.p + 2 * (size_t)chunks[i].r0;
      uint32_t* n_gen = ix->small.p + 11 + 4 * i;
      gq::SeedOut pre{ix->seed_rec.p + 4 * (size_t)ix->seed_recs_per_read * chunks[i].r0,
                      (uint32_t)std::min<uint64_t>((uint64_t)ix->seed_recs_per_read * (chunks[i].r1 - chunks[i].r0), 0x3FFFFFFFull),
                      ix->small.p + 10 + 4 * i, ix->surv_cnt.p, gen_list, n_gen};
      const bool timed = chunks.size() == 1;  // per-kernel events (single-slice runs)
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[0], cs));
      gq::launch_seed(ix->dv, bc, oc, pre, cs);
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[1], cs));
      gq::launch_text(ix->dv, bc, oc, pre, ix->surv_rec.p + 4 * (size_t)ix->seed_recs_per_read * chunks[i].r0,
                      ix->small.p + 8 + 4 * kMaxChunks + i, cs, timed ? ix->kev[2] : nullptr);
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[3], cs));
      ++launches;
      gq::launch_search(ix->dv, bc, oc, arena, ix->arena_words, threads, gen_list,
                        2 * (chunks[i].r1 - chunks[i].r0), ix->super_in_smem, ix->rf_thresh, ix->ev_thresh, cs,
                        ix->leave_opt, ix->wait_opt, n_gen);
      launches += 2;
    } else
      gq::launch_search(ix->dv, bc, oc, arena, ix->arena_words, threads, nullptr, 0, ix->super_in_smem,
                        ix->rf_thresh, ix->ev_thresh, cs, ix->leave_opt, ix->wait_opt);
    if (chunks.size() == 1) CUDA_OK(cudaEventRecord(ix->ev[1], st));
    if (!two_streams && ix->overlap_classify) {
      // classify (issue-bound) and coverage (latency-bound) are independent: run them side by side
      if (!ix->aux_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&ix->aux_stream, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&ix->aux_event, cudaEventDisableTiming));
      }
      CUDA_OK(cudaEventRecord(ix->aux_event, cs));
      CUDA_OK(cudaStreamWaitEvent(ix->aux_stream, ix->aux_event, 0));
      CUDA_OK(cudaEventRecord(ix->kev[4], ix->aux_stream));
      gq::launch_classify(ix->dv, bc, oc, nullptr, 0, ix->aux_stream);
      CUDA_OK(cudaEventRecord(ix->kev[5], ix->aux_stream));
      gq::launch_coverage(ix->dv, bc, oc, c, arena, ix->arena_words, threads2, nullptr, 0,
                          ix->cov_overflow_list.p, ix->small.p + 2, ix->small.p + 8 + 5 * kMaxChunks + i, cs,
                          ix->multi_list.p + 2 * (size_t)chunks[i].r0, ix->small.p + 8 + 6 * kMaxChunks + i,
                          ix->heavy_list.p + 2 * (size_t)chunks[i].r0, ix->small.p + 8 + 7 * kMaxChunks + i);
      CUDA_OK(cudaEventRecord(ix->kev[6], cs));
      CUDA_OK(cudaEventRecord(ix->aux_event, ix->aux_stream));
      CUDA_OK(cudaStreamWaitEvent(cs, ix->aux_event, 0));
    } else {
      const bool timed = chunks.size() == 1;
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[4], cs));
      // sliced batches: the k-mer filter runs ONCE over the whole batch after the last slice (it only reads the
      // statuses; per slice it paid its fixed cost — two passes, a 128 KB set copied per CTA — seven times)
      if (chunks.size() == 1) gq::launch_classify(ix->dv, bc, oc, nullptr, 0, cs);
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[5], cs));
      gq::launch_coverage(ix->dv, bc, oc, c, arena, ix->arena_words, threads2, nullptr, 0,
                          ix->cov_overflow_list.p, ix->small.p + 2, ix->small.p + 8 + 5 * kMaxChunks + i, cs,
                          ix->multi_list.p + 2 * (size_t)chunks[i].r0, ix->small.p + 8 + 6 * kMaxChunks + i,
                          ix->heavy_list.p + 2 * (size_t)chunks[i].r0, ix->small.p + 8 + 7 * kMaxChunks + i);
      if (timed) CUDA_OK(cudaEventRecord(ix->kev[6], cs));
    }
    launches += 3;
    if (!tl.empty()) CUDA_OK(cudaEventRecord(tl[4 * i + 2], cs));
    if (i == early_upto) {
      // Pipelined path: the k-mer filter of everything mapped so far runs NOW on a stream of its own, beside the
      // remaining (small) slices — the call is copy-bound, so the GPU has room — and only the last slices' strands
      // are left for the pass after the last slice (the filter over the whole batch was 0.13 ms of serial tail).
      if (!ix->cls_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&ix->cls_stream, cudaStreamNonBlocking));
        for (auto& e : ix->cls_event) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
      CUDA_OK(cudaEventRecord(ix->cls_event[0], st));
      CUDA_OK(cudaEventRecord(ix->cls_event[1], ix->aux_stream));
      CUDA_OK(cudaStreamWaitEvent(ix->cls_stream, ix->cls_event[0], 0));
      CUDA_OK(cudaStreamWaitEvent(ix->cls_stream, ix->cls_event[1], 0));
      gq::BatchView be = b;
      be.read_begin = 0;
      be.read_end = chunks[i].r1;
      gq::launch_classify(ix->dv, be, o, nullptr, 0, ix->cls_stream);
      CUDA_OK(cudaEventRecord(ix->cls_event[2], ix->cls_stream));
      ++launches;
    }
  }
  if (two_streams) {
    CUDA_OK(cudaEventRecord(ix->aux_event, ix->aux_stream));
    CUDA_OK(cudaStreamWaitEvent(st, ix->aux_event, 0));
  }
  if (chunks.size() > 1) {
    gq::BatchView bl = b;
    if (early_upto != SIZE_MAX) {
      bl.read_begin = chunks[early_upto].r1;
      CUDA_OK(cudaStreamWaitEvent(st, ix->cls_event[2], 0));
    }
    gq::launch_classify(ix->dv, bl, o, nullptr, 0, st);
  }
  // the batch's five counters, committed to the totals on the device unless a strand overflowed (then they are
  // counted again after the re-runs); everything the host needs comes back in ONE pinned copy, one synchronisation
  CUDA_OK(cudaMemsetAsync(ix->stats_batch.p, 0, 40, st));
  gq::launch_stats(ix->status.p, ix->len.p, n, ix->stats_batch.p, st);
  gq::launch_stats_commit(ix->stats_batch.p, ix->stats.p, ix->small.p, st);
  launches += 2;
  CUDA_OK(cudaEventRecord(ix->ev[2], st));
  ix->info[7] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
  uint32_t* small = ix->post_host;
  uint32_t* gs = ix->post_host + 4;
  CUDA_OK(cudaMemcpyAsync(small, ix->small.p, 16, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(gs, ix->gsmall.p, 8, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  const bool had_reruns = small[1] > 0 || small[2] > 0;
  {
    float ms_t = 0;
    cudaEventElapsedTime(&ms_t, ix->ev[0], ix->ev[2]);
    ix->info[6] = ms_t;  // all kernels of the call (CUDA events on the caller's stream)
  }
  if (!tl.empty()) {
    for (size_t i = 0; i < chunks.size(); ++i) {
      float a = 0, b2 = 0, c2 = 0;
      cudaEventElapsedTime(&a, ix->ev[3], ix->tl_copy[i]);
      cudaEventElapsedTime(&b2, ix->ev[3], tl[4 * i + 1]);
      cudaEventElapsedTime(&c2, ix->ev[3], tl[4 * i + 2]);
      fprintf(stderr, "slice %zu: %u reads, copy done %.3f ms, kernels %.3f .. %.3f ms (stream %d)\n", i,
              chunks[i].r1 - chunks[i].r0, a, b2, c2, (int)(i & 1));
    }
    for (auto e : tl) cudaEventDestroy(e);
    for (auto e : ix->tl_copy) cudaEventDestroy(e);
    ix->tl_copy.clear();
  }
  if (chunks.size() == 1) {
    float ms_s = 0, ms_c = 0;
    cudaEventElapsedTime(&ms_s, ix->ev[0], ix->ev[1]);
    cudaEventElapsedTime(&ms_c, ix->ev[1], ix->ev[2]);
    ix->info[2] = ms_s;
    ix->info[3] = ms_c;
    for (auto& m : ix->kernel_ms) m = 0;
    if (ix->use_seed_pass) {
      auto el = [&](int a, int b) {
        float ms = 0;
        return cudaEventElapsedTime(&ms, ix->kev[a], ix->kev[b]) == cudaSuccess ? (double)ms : 0.0;
      };
      ix->kernel_ms[0] = el(0, 1);  // seed
      ix->kernel_ms[1] = el(1, 2);  // verify
      ix->kernel_ms[2] = el(2, 3);  // text
      float ms_g = 0;               // general: from the end of the text kernel to the end of the search phase
      cudaEventElapsedTime(&ms_g, ix->kev[3], ix->ev[1]);
      ix->kernel_ms[3] = ms_g;
      ix->kernel_ms[4] = el(4, 5);  // classify (second stream, beside coverage)
      // coverage: beside classify (second stream) it starts with the search phase's end, else after classify
      float ms_cov = 0;
      cudaEventElapsedTime(&ms_cov, ix->overlap_classify ? ix->ev[1] : ix->kev[5], ix->kev[6]);
      ix->kernel_ms[5] = ms_cov;
      float ms_rc = 0;  // revcomp: from the start of the call's device work to the seed kernel
      cudaEventElapsedTime(&ms_rc, ix->ev[0], ix->kev[0]);
      ix->kernel_ms[6] = ms_rc;
    }
  }
  // list-mode re-runs below hand out work from counter slot [9] and append to the (already consumed)
  // mapped list of slice 0
  uint64_t rerun = 0;
  // ---- overflow re-runs: same kernels in list mode (strands seeded inside the search kernel) ----
  // Arena / state-pool overflows are re-run with fewer lanes and much larger per-lane arenas (x4 per
  // further retry). (A full survivor pool of the seed pass is not an overflow: those strands simply take
  // the general kernel.)
  uint32_t big_words = ix->big_arena_words, big_threads = ix->big_threads;
  const bool seed_pool_miss = false;
  int guard = 0;
  while (small[1] > 0) {
    uint32_t n_list = small[1];
    rerun += n_list;
    if (++guard > 12) throw std::runtime_error("search state arena overflow persists at the largest arena size");
    const bool normal_cfg = seed_pool_miss && guard == 1;
    if (small[0] > ix->pool.cap) {  // the pool ran out: grow it, keeping what was written
      size_t ncap = std::min<size_t>(std::max<size_t>((size_t)small[0] * 2, ix->pool.cap * 2), 0xFFFFFFF0ull);
      if (ncap <= ix->pool.cap) throw std::runtime_error("final-state pool exceeds 2^32 words; use smaller batches");
      uint32_t* np = nullptr;
      CUDA_OK(cudaMalloc(&np, ncap * 4));
      CUDA_OK(cudaMemcpyAsync(np, ix->pool.p, ix->pool.cap * 4, cudaMemcpyDeviceToDevice, st));
      CUDA_OK(cudaStreamSynchronize(st));
      uint32_t used = (uint32_t)ix->pool.cap;
      cudaFree(ix->pool.p);
      ix->pool.p = np;
      ix->pool.cap = ncap;
      CUDA_OK(cudaMemcpyAsync(ix->small.p, &used, 4, cudaMemcpyHostToDevice, st));
      o.pool = np;
      o.pool_cap = (uint32_t)ncap;
    }
    uint32_t bt, aw;
    uint32_t* ar;
    if (normal_cfg) {
      bt = std::min<uint32_t>(ix->n_threads, std::max<uint32_t>(256, ((n_list / 4 + 255) / 256) * 256));
      aw = ix->arena_words;
      ix->arena.reserve((size_t)std::max(bt, threads2) * aw);
      ar = ix->arena.p;
    } else {
      bt = std::min<uint32_t>(big_threads, ((n_list + 255) / 256) * 256);
      aw = big_words;
      ix->big_arena.reserve((size_t)bt * big_words);
      ar = ix->big_arena.p;
    }
    // The kernel appends to the same list while reading it: read from a copy. A strand can be on the list twice
    // (flagged by the text kernel when the pool filled up, and again by the general kernel that took it over):
    // the copy holds every strand once, so no strand is mapped by two lanes or recorded twice.
    std::vector<uint32_t> hl(n_list);
    CUDA_OK(cudaMemcpyAsync(hl.data(), ix->overflow_list.p, (size_t)n_list * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    std::sort(hl.begin(), hl.end());
    hl.erase(std::unique(hl.begin(), hl.end()), hl.end());
    n_list = (uint32_t)hl.size();
    DevBuf<uint32_t> list;
    list.reserve(n_list);
    CUDA_OK(cudaMemcpyAsync(list.p, hl.data(), (size_t)n_list * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(ix->small.p + 1, 0, 4, st));
    CUDA_OK(cudaMemsetAsync(ix->small.p + 9, 0, 4, st));  // work counter of the list run
    gq::launch_search(ix->dv, b, o, ar, aw, bt, list.p, n_list, ix->super_in_smem, ix->rf_thresh, ix->ev_thresh, st,
                      ix->leave_opt, ix->wait_opt);
    gq::launch_classify(ix->dv, b, o, list.p, n_list, st);
    ++launches;
    gq::launch_coverage(ix->dv, b, o, c, ar, aw, std::min<uint32_t>(bt, threads2), list.p, n_list,
                        ix->cov_overflow_list.p, ix->small.p + 2, ix->small.p + 8 + 5 * kMaxChunks, st);
    launches += 2;
    CUDA_OK(cudaMemcpyAsync(small, ix->small.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    list.release();
    if (small[1] > 0 && !normal_cfg) {
      big_words *= 4;
      big_threads = std::max<uint32_t>(256, big_threads / 4);
    }
  }
  guard = 0;
  while (small[2] > 0) {
    uint32_t n_list = small[2];
    rerun += n_list;
    if (++guard > 16) throw std::runtime_error("coverage scratch overflow persists at the largest arena size");
    // strands given back because the multi-allele group table (or its record pool) was full: grow it first
    uint32_t gs0[2];
    CUDA_OK(cudaMemcpyAsync(gs0, ix->gsmall.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    const bool grouped_full = (gs0[1] & 1u) != 0;
    if (grouped_full) {
      grow_groups(ix);  // clears the "full" flag, keeps the others
      c = cov_view(ix);
    }
    uint32_t bt = std::min<uint32_t>(big_threads, ((n_list + 255) / 256) * 256);
    ix->big_arena.reserve((size_t)bt * big_words);
    DevBuf<uint32_t> list;
    list.reserve(n_list);
    CUDA_OK(cudaMemcpyAsync(list.p, ix->cov_overflow_list.p, (size_t)n_list * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_OK(cudaMemsetAsync(ix->small.p + 2, 0, 4, st));
    gq::launch_coverage(ix->dv, b, o, c, ix->big_arena.p, big_words, bt, list.p, n_list, ix->cov_overflow_list.p,
                        ix->small.p + 2, ix->small.p + 8 + 5 * kMaxChunks, st);
    ++launches;
    CUDA_OK(cudaMemcpyAsync(small, ix->small.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
    list.release();
    if (small[2] > 0 && !grouped_full) {
      big_words *= 4;
      big_threads = std::max<uint32_t>(256, big_threads / 4);
    }
  }
  if (had_reruns) {  // statuses changed: count the batch again, commit, and look at the error flags once more
    CUDA_OK(cudaMemsetAsync(ix->stats_batch.p, 0, 40, st));
    gq::launch_stats(ix->status.p, ix->len.p, n, ix->stats_batch.p, st);
    gq::launch_stats_commit(ix->stats_batch.p, ix->stats.p, nullptr, st);
    launches += 2;
    CUDA_OK(cudaEventRecord(ix->ev[2], st));  // end of the call's device work, re-runs included
    CUDA_OK(cudaMemcpyAsync(gs, ix->gsmall.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    CUDA_OK(cudaGetLastError());
  }
  if (gs[1] & 1u) throw std::runtime_error("grouped allele count table is full (internal error: flagged strands were not re-run)");
  if (gs[1] & 2u) throw std::runtime_error("inconsistent traversal while recording per-base coverage");
  {
    float ms_t = 0;
    cudaEventElapsedTime(&ms_t, ix->ev[0], ix->ev[2]);
    ix->info[6] = ms_t;  // all kernels of the call, overflow re-runs included (CUDA events on the caller's stream)
  }
  ix->info[0] = launches;
  ix->info[1] = (double)rerun;
  ix->info[4] = small[0];
}

// device side of a handle whose host index (ix->h) is ready: streams, upload, candidate pool sizing, accumulators
static void finish_handle(gq_index* ix) {
  const int device = ix->device;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
  ix->stream = ix->own_stream;
  for (auto& e : ix->ev) CUDA_OK(cudaEventCreate(&e));
  for (auto& e : ix->kev) CUDA_OK(cudaEventCreate(&e));
  upload_index(ix);
  {  // candidate records per read: a forward strand's seeding k-mer is drawn by occurrence (size-biased
     // mean of the suffix counts), a reverse strand's is any k-mer (plain mean over all 4^k)
    double sum = 0, sum2 = 0;
    const auto& h = ix->h;
    const uint64_t nk = 1ull << (2 * h.k);
    for (uint64_t c = 0; c < nk; ++c) {
      double w = 0;
      for (uint32_t j = h.kmer_off[c]; j < h.kmer_off[c + 1]; ++j) w += (double)(h.kmer_states[j].hi - h.kmer_states[j].lo + 1);
      sum += w;
      sum2 += w * w;
    }
    const double per_read = (sum > 0 ? sum2 / sum : 1.0) + sum / (double)nk;
    // ... of which the left-context check of the seed pass keeps a few per read whatever the k-mer's frequency
    // (config 2: 1.8); a pool that fills up only sends strands to the general kernel
    ix->seed_recs_per_read = (uint32_t)std::min(64.0, 8.0 + per_read / 8.0);
  }
  alloc_coverage(ix);
  reset_coverage(ix);
  // nested PRGs: paths of a dozen loci and dozens of states per strand are common — larger per-lane arenas keep
  // most strands out of the overflow re-runs
  if (ix->h.is_nested) ix->arena_words = std::max<uint32_t>(ix->arena_words, 2048);
}

// -------------------------------------------------------------------------------------------------
#define GQ_TRY try {
#define GQ_CATCH                    \
  }                                 \
  catch (const std::exception& e) { \
    g_err = e.what();               \
    return -1;                      \
  }                                 \
  return 0;

extern "C" {

const char* gq_last_error(void) { return g_err.c_str(); }

// developer hook (not in gq.h): warp-loop statistics when built with -DGQ_DEBUG_COUNTERS
void gq_debug_counters(unsigned long long* out32) { gq::debug_counters(out32); }

static int index_build_impl(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device, const char* kmer_dir,
                            gq_index** out);

int gq_index_build(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device, gq_index** out) {
  return index_build_impl(prg, n_symbols, kmer_size, device, nullptr, out);
}

int gq_index_build_from_gram_dir(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device,
                                 const char* gram_dir, gq_index** out) {
  if (!gram_dir) {
    g_err = "null argument";
    return -1;
  }
  return index_build_impl(prg, n_symbols, kmer_size, device, gram_dir, out);
}

int gq_index_save(const gq_index* ix, const char* path) {
  GQ_TRY
  if (!ix || !path) throw std::runtime_error("null argument");
  gq::host_index_save(ix->h, path);
  GQ_CATCH
}

int gq_index_load(const char* path, int device, gq_index** out) {
  gq_index* ix = nullptr;
  GQ_TRY
  if (!path || !out) throw std::runtime_error("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw std::runtime_error("no CUDA device: libgq has no CPU fallback");
  if (device < 0 || device >= ndev) throw std::runtime_error("invalid device ordinal");
  ix = new gq_index();
  ix->device = device;
  gq::host_index_load(ix->h, path);
  finish_handle(ix);
  *out = ix;
  }
  catch (const std::exception& e) {
    g_err = e.what();
    if (ix) gq_index_destroy(ix);
    return -1;
  }
  return 0;
}

int gq_index_prg(const gq_index* ix, uint32_t* prg_out, uint64_t* n_symbols) {
  GQ_TRY
  if (!ix || !n_symbols) throw std::runtime_error("null argument");
  *n_symbols = ix->h.prg.size();
  if (prg_out) std::copy(ix->h.prg.begin(), ix->h.prg.end(), prg_out);
  GQ_CATCH
}

int gq_kmer_index_dump(const gq_index* ix, const char* gram_dir) {
  GQ_TRY
  if (!ix || !gram_dir) throw std::runtime_error("null argument");
  gq::kmer_index_dump(ix->h, gram_dir);
  GQ_CATCH
}

static int index_build_impl(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device, const char* kmer_dir,
                            gq_index** out) {
  gq_index* ix = nullptr;
  GQ_TRY
  if (!prg || !out) throw std::runtime_error("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw std::runtime_error("no CUDA device: libgq has no CPU fallback");
  if (device < 0 || device >= ndev) throw std::runtime_error("invalid device ordinal");
  ix = new gq_index();
  ix->device = device;
  // suffix array on the GPU (sa_gpu.cu) unless GQ_HOST_SA is set (developer switch: host SA-IS, for comparison)
  struct SaCtx {
    int device;
  } sctx{device};
  gq::SaBuilder sab = [](const std::vector<int32_t>& text, int32_t sigma, void* c) {
    return gq::gpu_suffix_array(text, sigma, ((SaCtx*)c)->device);
  };
  gq::build_host_index(prg, n_symbols, kmer_size, ix->h, getenv("GQ_HOST_SA") ? nullptr : sab, &sctx, kmer_dir);
  finish_handle(ix);
  *out = ix;
  }
  catch (const std::exception& e) {
    g_err = e.what();
    if (ix) gq_index_destroy(ix);
    return -1;
  }
  return 0;
}

int gq_suffix_array(const uint32_t* prg, uint64_t n_symbols, int device, uint32_t* sa_out, int* rounds) {
  GQ_TRY
  if (!prg || !sa_out || n_symbols == 0) throw std::runtime_error("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw std::runtime_error("no CUDA device: libgq has no CPU fallback");
  if (device < 0 || device >= ndev) throw std::runtime_error("invalid device ordinal");
  if (n_symbols >= (1ull << 32) - 3) throw std::runtime_error("PRG too long: text positions are 32-bit words");
  std::vector<uint32_t> present;
  std::vector<int32_t> text;
  const int32_t sigma = gq::compress_text(prg, n_symbols, present, text);
  std::vector<uint32_t> sa = gq::gpu_suffix_array(text, sigma, device, rounds);
  std::copy(sa.begin(), sa.end(), sa_out);
  GQ_CATCH
}

int gq_index_clone(const gq_index* src, int device, gq_index** out) {
  gq_index* ix = nullptr;
  GQ_TRY
  if (!src || !out) throw std::runtime_error("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) throw std::runtime_error("invalid device ordinal");
  ix = new gq_index();
  ix->device = device;
  ix->h = src->h;  // the host index is shared by value: built once, uploaded per GPU
  finish_handle(ix);
  *out = ix;
  }
  catch (const std::exception& e) {
    g_err = e.what();
    if (ix) gq_index_destroy(ix);
    return -1;
  }
  return 0;
}

int gq_index_destroy(gq_index* ix) {
  if (!ix) return 0;
  gq_comm_destroy(ix);
  cudaSetDevice(ix->device);
  for (void* p : ix->index_allocs) cudaFree(p);
  ix->counters.release();
  ix->allele_off.release();
  ix->gtab.release();
  ix->gcount.release();
  ix->gpool.release();
  ix->gsmall.release();
  ix->stats.release();
  ix->stats_batch.release();
  if (ix->post_host) cudaFreeHost(ix->post_host);
  ix->bases.release();
  ix->offsets.release();
  ix->word_off.release();
  ix->packed.release();
  ix->packed_rc.release();
  ix->len.release();
  ix->seeds.release();
  ix->status.release();
  ix->st_off.release();
  ix->st_words.release();
  ix->st_count.release();
  ix->pool.release();
  ix->small.release();
  ix->overflow_list.release();
  ix->cov_overflow_list.release();
  ix->mapped_list.release();
  ix->multi_list.release();
  ix->heavy_list.release();
  ix->seed_rec.release();
  ix->surv_rec.release();
  ix->surv_cnt.release();
  ix->gen_list.release();
  ix->arena.release();
  ix->big_arena.release();
  for (auto& e : ix->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ix->kev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ix->chunk_events) cudaEventDestroy(e);
  if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
  if (ix->aux_stream) cudaStreamDestroy(ix->aux_stream);
  if (ix->aux_event) cudaEventDestroy(ix->aux_event);
  if (ix->cls_stream) cudaStreamDestroy(ix->cls_stream);
  for (auto e : ix->cls_event)
    if (e) cudaEventDestroy(e);
  ix->arena2.release();
  if (ix->fetch_host) cudaFreeHost(ix->fetch_host);
  ix->fetch_dev.release();
  if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
  delete ix;
  return 0;
}

int gq_index_describe(const gq_index* ix, gq_layout* out) {
  GQ_TRY
  if (!ix || !out) throw std::runtime_error("null argument");
  out->n_symbols = ix->h.prg.size();
  out->sa_size = ix->h.n;
  out->kmer_size = ix->h.k;
  out->n_sites = ix->h.n_sites;
  out->n_site_slots = ix->h.n_slots;
  out->is_nested = ix->h.is_nested;
  out->n_alleles = ix->n_alleles;
  out->n_per_base = ix->n_per_base;
  out->n_kmer_states = ix->h.kmer_off.back();
  out->device_bytes = ix->index_bytes;
  GQ_CATCH
}

int gq_index_allele_offsets(const gq_index* ix, uint64_t* allele_off) {
  GQ_TRY
  if (!ix || !allele_off) throw std::runtime_error("null argument");
  for (size_t i = 0; i < ix->h.allele_off.size(); ++i) allele_off[i] = ix->h.allele_off[i];
  GQ_CATCH
}

int gq_index_per_base_layout(const gq_index* ix, uint64_t* off_len) {
  GQ_TRY
  if (!ix || !off_len) throw std::runtime_error("null argument");
  const gq::HostIndex& h = ix->h;
  for (uint32_t s = 0; s < h.n_slots; ++s) {
    if (h.n_alleles[s] == 0) continue;
    const gq::Node& sn = h.nodes[h.site_start_node[s]];
    for (uint32_t a = 0; a < h.n_alleles[s]; ++a) {
      const gq::Node& an = h.nodes[h.edges[sn.edge_off + a]];
      uint64_t* d = off_len + 2 * ((uint64_t)h.allele_off[s] + a);
      bool seq = an.len > 0 && an.cov_off != gq::kNoAllele && an.site == 5 + 2 * s;
      d[0] = seq ? an.cov_off : 0;
      d[1] = seq ? an.len : 0;
    }
  }
  GQ_CATCH
}

int gq_batch_upload(gq_index* ix, const uint8_t* bases, const uint64_t* off, uint64_t n_reads, const uint32_t* seeds) {
  GQ_TRY
  if (!ix || (n_reads && (!bases || !off || !seeds))) throw std::runtime_error("null argument");
  do_upload(ix, bases, off, n_reads, seeds);
  GQ_CATCH
}

int gq_map_resident(gq_index* ix) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  do_map(ix);
  GQ_CATCH
}

int gq_map_batch(gq_index* ix, const uint8_t* bases, const uint64_t* off, uint64_t n_reads, const uint32_t* seeds) {
  GQ_TRY
  if (!ix || (n_reads && (!bases || !off || !seeds))) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  if (n_reads && off[0] != 0) throw std::runtime_error("read_offsets[0] must be 0");
  uint64_t nb = n_reads ? off[n_reads] : 0;
  reserve_batch(ix, n_reads, nb);
  ix->n_reads = (uint32_t)n_reads;
  ix->info[5] = (double)(nb + (n_reads + 1) * 8 + n_reads * 4);
  HostBatch hb;
  hb.bases = bases;
  hb.off = off;
  hb.seeds = seeds;
  do_map(ix, &hb);
  GQ_CATCH
}

// pinned (page-locked) host memory for the buffers handed to gq_map_batch*: asynchronous H2D copies at full PCIe rate
int gq_host_alloc(uint64_t bytes, void** out) {
  GQ_TRY
  if (!out) throw std::runtime_error("null argument");
  *out = nullptr;
  CUDA_OK(cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable));
  GQ_CATCH
}

int gq_host_free(void* p) {
  GQ_TRY
  if (p) CUDA_OK(cudaFreeHost(p));
  GQ_CATCH
}

int gq_device_count(int* n) {
  GQ_TRY
  if (!n) throw std::runtime_error("null argument");
  *n = 0;
  if (cudaGetDeviceCount(n) != cudaSuccess) *n = 0;
  GQ_CATCH
}

int gq_packed_words(const uint64_t* off, uint64_t n_reads, uint64_t* n_words) {
  GQ_TRY
  if (!n_words || (n_reads && !off)) throw std::runtime_error("null argument");
  *n_words = (n_reads ? (off[n_reads] >> 4) : 0) + n_reads + 1;
  GQ_CATCH
}

// read r starts at word (off[r] >> 4) + r: word-aligned, non-overlapping, computable per read without a prefix
// sum (at most one word wasted per read) — the layout pack_kernel produces on the device
// ASCII -> 2-bit code (A/a 0, C/c 1, G/g 2, T/t 3), 0x80 for every other character (encode_char, utils.cpp:13-47)
struct AsciiLut {
  uint8_t t[256];
  AsciiLut() {
    for (int i = 0; i < 256; ++i) t[i] = 0x80;
    t['A'] = t['a'] = 0;
    t['C'] = t['c'] = 1;
    t['G'] = t['g'] = 2;
    t['T'] = t['t'] = 3;
  }
};
static const AsciiLut kAsciiLut;

int gq_pack_reads(const uint8_t* bases, const uint64_t* off, uint64_t n_reads, uint32_t* packed, uint32_t* word_off,
                  uint32_t* len, int n_threads) {
  GQ_TRY
  if (n_reads && (!bases || !off || !packed || !word_off || !len)) throw std::runtime_error("null argument");
  if (n_reads && ((off[n_reads] >> 4) + n_reads + 1 >= (1ull << 32))) throw std::runtime_error("batch too large (packed words exceed 2^32)");
  const int nt = n_threads > 0 ? n_threads : 1;
  (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
    const uint64_t b0 = off[r];
    const uint32_t L = (uint32_t)(off[r + 1] - b0), w0 = (uint32_t)(b0 >> 4) + (uint32_t)r;
    word_off[r] = w0;
    len[r] = L;
    const uint8_t* src = bases + b0;
    const uint32_t full = L >> 4;
    for (uint32_t w = 0; w < full; ++w) {  // whole words: 16 bases, unrolled
      const uint8_t* s16 = src + 16 * w;
      uint32_t x = 0;
      for (uint32_t j = 0; j < 16; ++j) x |= ((uint32_t)(s16[j] - 1) & 3u) << (2 * j);
      packed[w0 + w] = x;
    }
    if (L & 15u) {
      uint32_t x = 0;
      for (uint32_t j = 0; j < (L & 15u); ++j) x |= ((uint32_t)(src[16 * full + j] - 1) & 3u) << (2 * j);
      packed[w0 + full] = x;
    }
  }
  if (word_off) word_off[n_reads] = (uint32_t)((n_reads ? (off[n_reads] >> 4) : 0) + n_reads);
  GQ_CATCH
}

// the same from ASCII sequence text (FASTQ / FASTA sequence lines): upper or lower case ACGT; a read holding any
// other character becomes EMPTY, as encode_dna_bases does (utils.cpp:13-47,72-92) — len 0, counted as skipped.
// One table look-up per character, 16 per packed word, validity OR-ed over the whole read (the ingestion pipeline of
// `gram genotype` runs this on every read: 0.6 -> several M reads/s per thread against a compare chain per base).
int gq_pack_ascii(const char* text, const uint64_t* off, uint64_t n_reads, uint32_t* packed, uint32_t* word_off,
                  uint32_t* len, int n_threads) {
  GQ_TRY
  if (n_reads && (!text || !off || !packed || !word_off || !len)) throw std::runtime_error("null argument");
  if (n_reads && ((off[n_reads] >> 4) + n_reads + 1 >= (1ull << 32))) throw std::runtime_error("batch too large (packed words exceed 2^32)");
  const int nt = n_threads > 0 ? n_threads : 1;
  (void)nt;
  const uint8_t* lut = kAsciiLut.t;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t r = 0; r < (int64_t)n_reads; ++r) {
    const uint64_t b0 = off[r];
    const uint32_t L = (uint32_t)(off[r + 1] - b0), w0 = (uint32_t)(b0 >> 4) + (uint32_t)r;
    word_off[r] = w0;
    const unsigned char* src = (const unsigned char*)text + b0;
    const uint32_t full = L >> 4;
    uint32_t bad = 0;
    for (uint32_t w = 0; w < full; ++w) {
      const unsigned char* s16 = src + 16 * w;
      uint32_t x = 0;
      for (uint32_t j = 0; j < 16; ++j) {
        const uint32_t c = lut[s16[j]];
        bad |= c;
        x |= (c & 3u) << (2 * j);
      }
      packed[w0 + w] = x;
    }
    if (L & 15u) {
      uint32_t x = 0;
      for (uint32_t j = 0; j < (L & 15u); ++j) {
        const uint32_t c = lut[src[16 * full + j]];
        bad |= c;
        x |= (c & 3u) << (2 * j);
      }
      packed[w0 + full] = x;
    }
    len[r] = (bad & 0x80u) ? 0u : L;
  }
  if (word_off) word_off[n_reads] = (uint32_t)((n_reads ? (off[n_reads] >> 4) : 0) + n_reads);
  GQ_CATCH
}

int gq_map_batch_packed(gq_index* ix, const uint32_t* packed, const uint32_t* word_off, const uint32_t* len,
                        uint64_t n_reads, const uint32_t* seeds) {
  GQ_TRY
  if (!ix || (n_reads && (!packed || !word_off || !len || !seeds))) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  if (n_reads >= (1ull << 30)) throw std::runtime_error("batch too large (max 2^30 reads per batch)");
  const uint64_t total_words = n_reads ? word_off[n_reads] : 0;
  ix->word_off.reserve(n_reads + 1);
  ix->packed.reserve(total_words + 2);
  ix->packed_rc.reserve(total_words + 2);
  ix->len.reserve(n_reads);
  ix->seeds.reserve(n_reads);
  ix->n_reads = (uint32_t)n_reads;
  ix->info[5] = (double)(total_words * 4 + (n_reads + 1) * 4 + n_reads * 8);
  HostBatch hb;
  hb.packed = packed;
  hb.word_off = word_off;
  hb.len = len;
  hb.seeds = seeds;
  do_map(ix, &hb);
  GQ_CATCH
}

int gq_batch_status(gq_index* ix, uint8_t* status) {
  GQ_TRY
  if (!ix || !status) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  if (ix->n_reads) CUDA_OK(cudaMemcpy(status, ix->status.p, 2 * (size_t)ix->n_reads, cudaMemcpyDeviceToHost));
  GQ_CATCH
}

static void fetch_strand_tables(gq_index* ix, std::vector<uint32_t>& off, std::vector<uint32_t>& words,
                                std::vector<uint32_t>& count, std::vector<uint8_t>& status) {
  size_t m = 2 * (size_t)ix->n_reads;
  off.resize(m);
  words.resize(m);
  count.resize(m);
  status.resize(m);
  if (!m) return;
  CUDA_OK(cudaMemcpy(off.data(), ix->st_off.p, m * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(words.data(), ix->st_words.p, m * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(count.data(), ix->st_count.p, m * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(status.data(), ix->status.p, m, cudaMemcpyDeviceToHost));
}

int gq_batch_states_size(gq_index* ix, uint64_t* n_words) {
  GQ_TRY
  if (!ix || !n_words) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  std::vector<uint32_t> off, words, count;
  std::vector<uint8_t> status;
  fetch_strand_tables(ix, off, words, count, status);
  uint64_t t = 0;
  for (size_t i = 0; i < words.size(); ++i)
    if (status[i] == gq::ST_MAPPED) t += words[i];
  *n_words = t;
  GQ_CATCH
}

int gq_batch_states(gq_index* ix, uint64_t* strand_off, uint32_t* strand_count, uint32_t* out_words) {
  GQ_TRY
  if (!ix || !strand_off || !strand_count) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  std::vector<uint32_t> off, words, count;
  std::vector<uint8_t> status;
  fetch_strand_tables(ix, off, words, count, status);
  uint32_t used = 0;
  if (ix->n_reads) CUDA_OK(cudaMemcpy(&used, ix->small.p, 4, cudaMemcpyDeviceToHost));
  used = (uint32_t)std::min<size_t>(used, ix->pool.cap);
  std::vector<uint32_t> pool(used);
  if (used) CUDA_OK(cudaMemcpy(pool.data(), ix->pool.p, (size_t)used * 4, cudaMemcpyDeviceToHost));
  uint64_t t = 0;
  for (size_t i = 0; i < words.size(); ++i) {
    strand_off[i] = t;
    bool m = status[i] == gq::ST_MAPPED;
    strand_count[i] = m ? count[i] : 0;
    if (m) {
      if (out_words) std::memcpy(out_words + t, pool.data() + off[i], (size_t)words[i] * 4);
      t += words[i];
    }
  }
  strand_off[words.size()] = t;
  GQ_CATCH
}

int gq_coverage_fetch(gq_index* ix, uint16_t* allele_sum, uint16_t* per_base, uint64_t stats[5]) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  CUDA_OK(cudaStreamSynchronize(ix->stream));
  // uint16 semantics applied on the device; one D2H of the uint16 vectors + counters into pinned memory
  const size_t n16 = ix->n_alleles + ix->n_per_base;
  const size_t bytes = n16 * 2 + 64;
  if (ix->fetch_host_bytes < bytes) {
    if (ix->fetch_host) cudaFreeHost(ix->fetch_host);
    ix->fetch_host = nullptr;
    CUDA_OK(cudaMallocHost(&ix->fetch_host, bytes));
    ix->fetch_host_bytes = bytes;
  }
  ix->fetch_dev.reserve(n16 + 1);
  unsigned long long* hs = (unsigned long long*)ix->fetch_host;
  uint16_t* h16 = (uint16_t*)((char*)ix->fetch_host + 64);
  if ((allele_sum || per_base) && n16) {
    gq::launch_fetch(ix->counters.p, (uint32_t)ix->n_alleles, ix->counters.p + 2 * ix->n_alleles, (uint32_t)ix->n_per_base,
                     ix->fetch_dev.p, ix->stream);
    CUDA_OK(cudaMemcpyAsync(h16, ix->fetch_dev.p, n16 * 2, cudaMemcpyDeviceToHost, ix->stream));
  }
  if (stats) CUDA_OK(cudaMemcpyAsync(hs, ix->stats.p, 40, cudaMemcpyDeviceToHost, ix->stream));
  CUDA_OK(cudaStreamSynchronize(ix->stream));
  if (allele_sum && ix->n_alleles) std::memcpy(allele_sum, h16, ix->n_alleles * 2);
  if (per_base && ix->n_per_base) std::memcpy(per_base, h16 + ix->n_alleles, ix->n_per_base * 2);
  if (stats)
    for (int i = 0; i < 5; ++i) stats[i] = hs[i];
  GQ_CATCH
}

}  // extern "C"

// all groups of this handle as (slot, alleles) -> raw count; singles first from the dense array
void collect_groups(gq_index* ix, std::map<std::vector<uint32_t>, uint64_t>& out, bool multi_only) {
  CUDA_OK(cudaSetDevice(ix->device));
  CUDA_OK(cudaStreamSynchronize(ix->stream));
  const gq::HostIndex& h = ix->h;
  if (!multi_only && ix->n_alleles) {
    std::vector<uint32_t> single(ix->n_alleles);
    CUDA_OK(cudaMemcpy(single.data(), ix->counters.p + ix->n_alleles, ix->n_alleles * 4, cudaMemcpyDeviceToHost));
    for (uint32_t s = 0; s < h.n_slots; ++s)
      for (uint32_t a = 0; a < h.n_alleles[s]; ++a) {
        uint32_t cnt = single[h.allele_off[s] + a];
        if (cnt) out[{s, a}] += cnt;
      }
  }
  uint32_t gs[2];
  CUDA_OK(cudaMemcpy(gs, ix->gsmall.p, 8, cudaMemcpyDeviceToHost));
  if (gs[0] == 0) return;
  std::vector<uint32_t> tab(ix->gtab.cap), cnt(ix->gtab.cap), pool(std::min<size_t>(gs[0], ix->gpool.cap));
  CUDA_OK(cudaMemcpy(tab.data(), ix->gtab.p, tab.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(cnt.data(), ix->gcount.p, cnt.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(pool.data(), ix->gpool.p, pool.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < tab.size(); ++i) {
    if (!tab[i] || !cnt[i]) continue;
    const uint32_t* rec = pool.data() + (tab[i] - 1);
    std::vector<uint32_t> key{rec[0]};
    key.insert(key.end(), rec + 2, rec + 2 + rec[1]);
    out[key] += cnt[i];
  }
}

static int write_groups(const std::map<std::vector<uint32_t>, uint64_t>& g, uint32_t* words, uint64_t* n_words,
                        bool wrap16) {
  uint64_t t = 0;
  for (auto& e : g) {
    uint32_t cnt = wrap16 ? (uint32_t)(e.second & 0xFFFFu) : (uint32_t)e.second;
    if (wrap16 && cnt == 0) {
      // a key whose uint16 counter wrapped to exactly 0 still exists in the reference's map
    }
    if (words) {
      words[t] = e.first[0];
      words[t + 1] = cnt;
      words[t + 2] = (uint32_t)e.first.size() - 1;
      for (size_t i = 1; i < e.first.size(); ++i) words[t + 2 + i] = e.first[i];
    }
    t += 2 + e.first.size();
  }
  *n_words = t;
  return 0;
}

extern "C" {

int gq_coverage_grouped(gq_index* ix, uint32_t* words, uint64_t* n_words) {
  GQ_TRY
  if (!ix || !n_words) throw std::runtime_error("null argument");
  std::map<std::vector<uint32_t>, uint64_t> g;
  collect_groups(ix, g, false);
  write_groups(g, words, n_words, true);
  GQ_CATCH
}

int gq_coverage_groups_export(gq_index* ix, uint32_t* words, uint64_t* n_words) {
  GQ_TRY
  if (!ix || !n_words) throw std::runtime_error("null argument");
  std::map<std::vector<uint32_t>, uint64_t> g;
  collect_groups(ix, g, true);
  write_groups(g, words, n_words, false);
  GQ_CATCH
}

}  // extern "C"

// Rebuild the multi-allele group table + record pool from `g` on the host (with the device's hash) and upload
// them, growing both so that the table is at most a quarter full. Drops leaked / never-counted records.
static uint32_t hash_group_host(const std::vector<uint32_t>& key) {
  uint32_t hsh = 2166136261u ^ key[0];
  hsh *= 16777619u;
  for (size_t i = 1; i < key.size(); ++i) {
    hsh ^= key[i];
    hsh *= 16777619u;
  }
  hsh ^= hsh >> 15;
  return hsh;
}

void rebuild_groups(gq_index* ix, const std::map<std::vector<uint32_t>, uint64_t>& g, size_t min_cap) {
  size_t pool_words = 0;
  for (auto& e : g) pool_words += e.first.size() + 1;
  size_t cap = std::max<size_t>(ix->gtab.cap, min_cap);
  while (cap < 4 * g.size()) cap <<= 1;
  if (cap > (1u << 30)) throw std::runtime_error("grouped allele count table exceeds 2^30 slots");
  const size_t pool_cap = std::max<size_t>(std::max<size_t>(ix->gpool.cap, cap * 4), 2 * pool_words);
  if (pool_cap >= 0xFFFFFFF0ull) throw std::runtime_error("grouped allele count pool exceeds 2^32 words");
  ix->gtab.reserve(cap);
  ix->gcount.reserve(cap);
  ix->gpool.reserve(pool_cap);
  std::vector<uint32_t> tab(ix->gtab.cap, 0), cnt(ix->gtab.cap, 0), pool;
  pool.reserve(pool_words);
  const uint32_t maskc = (uint32_t)ix->gtab.cap - 1;
  for (auto& e : g) {
    uint32_t hsh = hash_group_host(e.first) & maskc;
    while (tab[hsh]) hsh = (hsh + 1) & maskc;  // terminates: the table is at most a quarter full
    tab[hsh] = (uint32_t)pool.size() + 1;
    cnt[hsh] = (uint32_t)e.second;
    pool.push_back(e.first[0]);
    pool.push_back((uint32_t)e.first.size() - 1);
    pool.insert(pool.end(), e.first.begin() + 1, e.first.end());
  }
  CUDA_OK(cudaMemcpy(ix->gtab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(ix->gcount.p, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice));
  if (!pool.empty()) CUDA_OK(cudaMemcpy(ix->gpool.p, pool.data(), pool.size() * 4, cudaMemcpyHostToDevice));
  uint32_t gs[2];
  CUDA_OK(cudaMemcpy(gs, ix->gsmall.p, 8, cudaMemcpyDeviceToHost));
  gs[0] = (uint32_t)pool.size();  // pool bump pointer
  gs[1] &= ~1u;                   // "table / pool full" is settled; other error flags stay
  CUDA_OK(cudaMemcpy(ix->gsmall.p, gs, 8, cudaMemcpyHostToDevice));
}

// the group table or its pool filled up during a batch (flagged strands were given back uncommitted): 4x larger
static void grow_groups(gq_index* ix) {
  std::map<std::vector<uint32_t>, uint64_t> g;
  collect_groups(ix, g, true);
  rebuild_groups(ix, g, ix->gtab.cap * 4);
}

extern "C" {

int gq_coverage_groups_import(gq_index* ix, const uint32_t* words, uint64_t n_words, int replace) {
  GQ_TRY
  if (!ix || (n_words && !words)) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  CUDA_OK(cudaStreamSynchronize(ix->stream));
  std::map<std::vector<uint32_t>, uint64_t> g;
  if (!replace) collect_groups(ix, g, true);
  for (uint64_t t = 0; t < n_words;) {
    if (t + 3 > n_words || t + 3 + words[t + 2] > n_words) throw std::runtime_error("truncated group record");
    uint32_t n = words[t + 2];
    std::vector<uint32_t> key{words[t]};
    key.insert(key.end(), words + t + 3, words + t + 3 + n);
    g[key] += words[t + 1];
    t += 3 + n;
  }
  rebuild_groups(ix, g, 0);
  GQ_CATCH
}

int gq_read_depth_stats(gq_index* ix, double out[2], uint64_t counts[2]) {
  GQ_TRY
  if (!ix || !out || !counts) throw std::runtime_error("null argument");
  const gq::HostIndex& h = ix->h;
  std::map<std::vector<uint32_t>, uint64_t> g;
  collect_groups(ix, g, false);
  std::vector<uint32_t> pb32(ix->n_per_base);
  if (ix->n_per_base)
    CUDA_OK(cudaMemcpy(pb32.data(), ix->counters.p + 2 * ix->n_alleles, ix->n_per_base * 4, cudaMemcpyDeviceToHost));
  // get_max_cov_haplogroup (read_stats.cpp:72-93): per site, allele with the largest summed group count
  // (uint16 counters in the reference: wrap each group count first); ties -> smallest allele id
  std::vector<std::map<uint32_t, uint32_t>> per_site(h.n_slots);
  for (auto& e : g) {
    uint16_t cnt = (uint16_t)(e.second & 0xFFFFu);
    for (size_t i = 1; i < e.first.size(); ++i) {
      auto& c = per_site[e.first[0]][e.first[i]];
      c = (uint16_t)(c + cnt);
    }
  }
  auto max_hap = [&](uint32_t slot) {
    std::pair<uint32_t, uint32_t> best{0, 0};
    bool first = true;
    for (auto& kv : per_site[slot])
      if (first || kv.second > best.second) best = kv, first = false;
    return best;
  };
  std::vector<double> covs;
  uint64_t nocov = 0;
  double total = 0;
  for (uint32_t s = 0; s < h.n_slots; ++s) {
    if (h.n_alleles[s] == 0 || h.par[2 * s] != 0) continue;  // nested sites are skipped (:131-132)
    // extract_max_coverage_allele (:95-117): walk from the bubble start to its end
    uint32_t cur = h.site_start_node[s];
    // the bubble end of site s is the node every allele path converges to: follow allele 0 greedily
    auto top = max_hap(s);
    uint64_t allele_cov = top.second;
    double sum = 0;
    uint64_t nb = 0;
    uint32_t depth = 0;  // open bubbles below the level-0 one
    while (true) {
      const gq::Node& nd = h.nodes[cur];
      bool is_start = nd.n_edges > 1 && nd.len == 0;
      if (is_start) {
        uint32_t slot = (nd.site - 5) / 2;
        auto m = max_hap(slot);
        if (m.first >= nd.n_edges) throw std::runtime_error("inconsistent grouped allele counts");
        ++depth;
        cur = h.edges[nd.edge_off + m.first];
        continue;
      }
      if (nd.len == 0 && nd.allele == -1 && nd.site != 0) {  // a bubble end
        if (--depth == 0) break;
      }
      if (nd.len > 0 && nd.cov_off != gq::kNoAllele) {
        for (uint32_t i = 0; i < nd.len; ++i) sum += std::min<uint32_t>(pb32[nd.cov_off + i], 65535u);
        nb += nd.len;
      }
      if (nd.n_edges == 0) break;
      cur = h.edges[nd.edge_off];
    }
    double site_cov = nb ? sum / (double)nb : (double)allele_cov;
    total += site_cov;
    covs.push_back(site_cov);
    if (allele_cov == 0) ++nocov;
  }
  double mean = covs.empty() ? 0.0 / 0.0 : total / covs.size();
  double var = 0;
  for (double c : covs) var += (c - mean) * (c - mean);
  var = covs.empty() ? 0.0 / 0.0 : var / covs.size();
  out[0] = mean;
  out[1] = var;
  counts[0] = nocov;
  counts[1] = covs.size();
  GQ_CATCH
}

int gq_coverage_reset(gq_index* ix) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  CUDA_OK(cudaSetDevice(ix->device));
  reset_coverage(ix);
  GQ_CATCH
}

int gq_coverage_device_ptrs(gq_index* ix, void** counters, uint64_t* n_counters, void** stats) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  if (counters) *counters = ix->counters.p;
  if (n_counters) *n_counters = 2 * ix->n_alleles + ix->n_per_base;
  if (stats) *stats = ix->stats.p;
  GQ_CATCH
}

int gq_set_stream(gq_index* ix, void* cuda_stream) {
  GQ_TRY
  if (!ix) throw std::runtime_error("null argument");
  ix->stream = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
  GQ_CATCH
}

int gq_set_option(gq_index* ix, const char* name, int64_t value) {
  GQ_TRY
  if (!ix || !name) throw std::runtime_error("null argument");
  std::string n(name);
  if (n == "arena_words") {
    if (value < 32) throw std::runtime_error("arena_words must be >= 32");
    ix->arena_words = (uint32_t)value;
    ix->arena.release();
  } else if (n == "threads") {
    if (value < 256) throw std::runtime_error("threads must be >= 256");
    ix->n_threads = (uint32_t)(value / 256 * 256);
    ix->arena.release();
  } else if (n == "big_arena_words") {
    ix->big_arena_words = (uint32_t)std::max<int64_t>(value, 64);
    ix->big_arena.release();
  } else if (n == "big_threads") {
    ix->big_threads = (uint32_t)std::max<int64_t>(value / 256 * 256, 256);
    ix->big_arena.release();
  } else if (n == "pool_words_per_read") {
    ix->pool_words_per_read = (uint32_t)std::max<int64_t>(value, 1);
    ix->pool.release();
  } else if (n == "gtab_cap") {  // (tests) shrink the multi-allele group table so that its growth path runs
    uint32_t cap = 4;
    while (cap < (uint64_t)std::max<int64_t>(value, 4)) cap <<= 1;
    CUDA_OK(cudaSetDevice(ix->device));
    CUDA_OK(cudaStreamSynchronize(ix->stream));
    ix->gtab.release();
    ix->gcount.release();
    ix->gpool.release();
    ix->gtab.reserve(cap);
    ix->gcount.reserve(cap);
    ix->gpool.reserve((size_t)cap * 4);
    reset_coverage(ix);
  } else if (n == "overlap_classify") {
    ix->overlap_classify = value != 0;
  } else if (n == "resident_slices") {
    ix->resident_slices = (uint32_t)std::max<int64_t>(value, 1);
  } else if (n == "early_classify") {
    ix->early_classify = value != 0;
  } else if (n == "tail_chunk_reads") {
    ix->tail_chunk_reads = (uint32_t)std::max<int64_t>(value, 1024);
  } else if (n == "chunk_reads") {
    ix->chunk_reads = (uint32_t)std::max<int64_t>(value, 1024);
  } else if (n == "seed_pass") {
    ix->use_seed_pass = value != 0;

  } else if (n == "seed_recs_per_read") {
    ix->seed_recs_per_read = (uint32_t)std::max<int64_t>(value, 1);
    ix->seed_rec.release();
  ix->surv_rec.release();
  } else if (n == "leave") {
    ix->leave_opt = (uint32_t)value;
  } else if (n == "wait_max") {
    ix->wait_opt = (uint32_t)value;
  } else if (n == "rf_thresh") {
    ix->rf_thresh = (uint32_t)value;
  } else if (n == "ev_thresh") {
    ix->ev_thresh = (uint32_t)value;
  } else if (n == "super_in_smem") {
    ix->super_in_smem = value != 0;
  } else
    throw std::runtime_error("unknown option: " + n);
  GQ_CATCH
}

int gq_last_kernel_ms(gq_index* ix, double ms[8]) {
  GQ_TRY
  if (!ix || !ms) throw std::runtime_error("null argument");
  for (int i = 0; i < 8; ++i) ms[i] = ix->kernel_ms[i];
  GQ_CATCH
}

int gq_last_run_info(gq_index* ix, double info[8]) {
  GQ_TRY
  if (!ix || !info) throw std::runtime_error("null argument");
  for (int i = 0; i < 8; ++i) info[i] = ix->info[i];
  GQ_CATCH
}

}  // extern "C"
