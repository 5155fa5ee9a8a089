// kernels.cu — hand-written sm_100a kernels for the quasimap hot path (per-strand logic in gq_device.cuh).
//
//   pack_kernel      uint8 bases -> 2-bit packed words (16 bases / word)
//   seed_kernel      seeding k-mer -> seed SearchStates -> candidates: one width-1 state per suffix
//                    (reference: quasimap.cpp:159-194,235-241; BWT_search.cpp for narrowing wide seeds)
//   verify_kernel    drops the candidates that disagree with the PRG within a few bases
//   text_kernel      walks the surviving candidates through the packed PRG text, pre-resolved jumps at
//                    variant markers, writes the strand's final SearchState
//                    (reference: quasimap.cpp:227-268, vBWT_jump.cpp, encapsulated_search.cpp)
//   search_kernel    general lane state machine for everything the text route does not take
//                    (interval states, marker scans, LIFO jump worklist; same reference functions)
//   classify_kernel  k-mer filter for strands that found nothing (quasimap.cpp:212-225)
//   coverage_kernel  equivalence classes, seeded selection, allele-sum / grouped / per-base recording
//                    (reference: coverage/coverage_common.cpp, allele_sum.cpp,
//                     grouped_allele_counts.cpp, allele_base.cpp)
//   stats_kernel     the five QuasimapReadsStats counters (quasimap.hpp:17-24)
//   fetch_kernel     uint16 wrap / saturate view of the accumulators
//
// Integer / bit arithmetic only; the kernels are bound by the latency of dependent 32 B sector loads (L2 when
// the index fits, HBM beyond), so there is no tensor-core work here. Rank superblock counters are staged
// into shared memory with one TMA bulk copy per CTA.
#include "kernels.cuh"
#include "gq_device.cuh"

#include <cstdio>

namespace gq {

// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ offsets, uint32_t r0,
                            uint32_t r1, uint32_t* __restrict__ word_off, uint32_t* __restrict__ packed,
                            uint32_t* __restrict__ len) {
  // one warp per read; lanes write consecutive words
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = r0 + warp; r < r1; r += nwarps) {
    uint64_t b0 = offsets[r];
    uint32_t L = (uint32_t)(offsets[r + 1] - b0);
    uint32_t w0 = (uint32_t)(b0 >> 4) + r, nw = (L + 15) >> 4;
    if (lane == 0) {
      len[r] = L;
      word_off[r] = w0;
    }
    for (uint32_t w = lane; w < nw; w += 32) {
      uint32_t x = 0;
      uint32_t base = w << 4;
      uint32_t cnt = min(16u, L - base);
      for (uint32_t j = 0; j < cnt; ++j) x |= ((uint32_t)(bases[b0 + base + j] - 1) & 3u) << (2 * j);
      packed[w0 + w] = x;
    }
  }
}

void launch_pack(const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1, uint32_t* word_off,
                 uint32_t* packed, uint32_t* len, cudaStream_t st) {
  if (r1 <= r0) return;
  uint32_t blocks = min((r1 - r0 + 7) / 8, 148u * 16u);
  pack_kernel<<<blocks, 256, 0, st>>>(bases, offsets, r0, r1, word_off, packed, len);
}

// Reverse strands of a slice, once per batch: a thread per read, eight output words per round with all their source
// loads issued together (the pass is a dependent chain len/word_off -> words -> store per read, so what it needs is
// loads in flight, not threads). Every later kernel walks an odd strand through packed_rc with the code of an
// even one.
__global__ void __launch_bounds__(256)
    revcomp_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ word_off,
                   const uint32_t* __restrict__ len, uint32_t r0, uint32_t r1, uint32_t* __restrict__ packed_rc) {
  const uint32_t r = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= r1) return;
  const uint32_t L = __ldg(len + r), woff = __ldg(word_off + r), nw = (L + 15) >> 4;
  const uint32_t* w = packed + woff;
  uint32_t* out = packed_rc + woff;
  for (uint32_t m0 = 0; m0 < nw; m0 += 8) {
    uint32_t y[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) y[j] = m0 + j < nw ? revcomp_word(w, L, m0 + j) : 0u;
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      if (m0 + j < nw) out[m0 + j] = y[j];
  }
}

void launch_revcomp(const BatchView& b, uint32_t* packed_rc, cudaStream_t st) {
  if (b.read_end <= b.read_begin) return;
  const uint32_t blocks = (b.read_end - b.read_begin + 255) / 256;
  revcomp_kernel<<<blocks, 256, 0, st>>>(b.packed, b.word_off, b.len, b.read_begin, b.read_end, packed_rc);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_stage_super(uint32_t* s_super, const uint32_t* g_super, uint32_t bytes,
                                                uint64_t* bar) {
  // one elected thread arms an mbarrier and issues a single TMA bulk copy (UBLKCP) global -> shared
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(s_super);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                 "l"(g_super), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar_a), "r"(0u)
        : "memory");
  }
}

constexpr int kSearchThreads = 256;
constexpr int kMaxSuperSmem = 2048;  // superblocks (x16 B = 32 KB) staged in shared memory

// Warp-synchronous driver of the lane state machine (gq_device.cuh). Each warp owns a contiguous chunk
// of strands; ballots decide, warp-uniformly, which unit operation the whole warp executes next:
//   * refill idle lanes                     when >= rf_thresh lanes are idle
//   * one class of rare-path transitions    when >= ev_thresh lanes wait in that class
//     (scan/jump, pop, top: lanes of one class run the same code together)
//   * everything pending                    when no lane can take a hot step, or too many lanes wait
//   * otherwise the text step (x kHotUnroll) for every lane holding a width-1 state
// Batching the rare paths keeps the hot step near full lane occupancy (v1 ran 3.5 lanes/instruction).
#ifndef GQ_HOT_UNROLL
#define GQ_HOT_UNROLL 2
#endif
constexpr int kHotUnroll = GQ_HOT_UNROLL;

#ifndef GQ_SEARCH_MIN_BLOCKS
#define GQ_SEARCH_MIN_BLOCKS 5  // 48 registers, 1280 resident lanes per SM
#endif
// Seed pass. Warp-convergent rounds of 32 strands. Phase A: every lane looks up the seed entries of its strand
// (preseed_lookup). Phase B: the entries of the round are examined (seed_state_cands) by all lanes —
//   * few entries per strand (config 2: ~4): the round's entries are spread evenly over the lanes, so lanes
//     run the same code instead of per-strand loops of very different lengths;
//   * many entries per strand (large genomes: tens to hundreds): strand by strand, 32 consecutive entries per
//     step — coalesced 8-byte loads, no search for the owner.
// Survivors are staged in shared memory and written through one warp-aggregated allocation when the stage
// fills. Superblock counters (only needed to narrow very wide seed states) come from shared memory when they
// fit, like in the search kernel.
constexpr uint32_t kSeedStage = 96;        // candidates staged per warp
constexpr uint32_t kSeedStrandMode = 384;  // entries per round of 32 strands from which mode 2 is used

struct SeedStage {
  uint32_t* rec;  // this warp's kSeedStage x 4 words in shared memory
  uint32_t n;     // staged candidates (warp-uniform)
};

// write the staged candidates to the pool; a full pool sends their strands to the general kernel
__device__ __forceinline__ void seed_flush(SeedStage& st, const SeedOut& pre, uint32_t lane) {
  const uint32_t full = 0xFFFFFFFFu;
  if (st.n == 0) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(pre.n_surv, st.n);
  base = __shfl_sync(full, base, 0);
  __syncwarp();
  const bool fits = base + st.n <= pre.cap;
  for (uint32_t q = lane; q < st.n; q += 32) {
    const uint4 c = reinterpret_cast<const uint4*>(st.rec)[q];
    if (fits) reinterpret_cast<uint4*>(pre.rec)[base + q] = c;
    else {
      send_to_general(pre, c.x);
      if (base + q < pre.cap) pre.rec[4 * (size_t)(base + q)] = kNoAllele;  // the slot stays dead
    }
  }
  __syncwarp();
  st.n = 0;
}

// stage up to `cnt` candidates per lane (cnt is 0 or 1 except for very wide seed states)
__device__ __forceinline__ void seed_push(SeedStage& st, const SeedOut& pre, const IndexView& v, const SeedCands& c,
                                          uint32_t cnt, uint32_t strand, uint32_t entry, uint32_t lane) {
  const uint32_t full = 0xFFFFFFFFu;
  uint32_t most = cnt;
#pragma unroll
  for (int d = 16; d; d >>= 1) most = max(most, __shfl_xor_sync(full, most, d));
  for (uint32_t r = 0; r < most; ++r) {
    const uint32_t mm = __ballot_sync(full, r < cnt);
    const uint32_t k = __popc(mm);
    if (st.n + k > kSeedStage) seed_flush(st, pre, lane);
    if (r < cnt) {
      uint32_t* d = st.rec + 4 * (st.n + __popc(mm & ((1u << lane) - 1u)));
      d[0] = strand;
      d[1] = __ldg(v.seed_state + entry);
      d[2] = c.p[r];
      d[3] = c.w0[r];
    }
    st.n += k;
  }
}

template <bool SUPER_SMEM>
__global__ void __launch_bounds__(256)
    seed_kernel(IndexView v, BatchView b, SearchOut o, SeedOut pre, uint32_t n_super_smem) {
  extern __shared__ __align__(128) uint32_t s_super[];
  __shared__ alignas(8) uint64_t s_bar;
  __shared__ alignas(16) uint32_t s_stage[8][kSeedStage * 4];
  if (SUPER_SMEM) tma_stage_super(s_super, v.super_cnt, n_super_smem * 16u, &s_bar);
  const uint32_t* super_c = SUPER_SMEM ? (const uint32_t*)s_super : v.super_cnt;
  const uint32_t n = 2 * (b.read_end - b.read_begin);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t full = 0xFFFFFFFFu;
  SeedStage stage{s_stage[threadIdx.x >> 5], 0};
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    const uint32_t strand = 2 * b.read_begin + i;
    SeedLookup lk;
    lk.classified = true;
    lk.s0 = lk.n0 = lk.s1 = lk.n1 = lk.pos0 = lk.ctx = 0;
    const uint32_t ns = i < n ? preseed_lookup(v, b, o, strand, lk) : 0;
    uint32_t my_L = 0, my_woff = 0;
    if (!lk.classified) {
      o.status[strand] = ST_UNCLASSIFIED;  // until a candidate finishes or the general kernel decides
      pre.surv_cnt[strand] = 0;
      my_L = b.len[strand >> 1];
      my_woff = b.word_off[strand >> 1];
    }
    // (the lookup also left the strand's next 12 bases left of the seeding k-mer in lk.ctx: every suffix entry of
    // the k-mer is decided by one XOR against its stored left context)
    const uint32_t my_pos0 = lk.pos0, my_ctx = lk.ctx;
    // one seed entry: a suffix entry is accepted or rejected from its 8 bytes and the owner strand's context; only
    // a wide state (more than kSplitWidth suffixes) takes the narrowing path
    auto examine = [&](uint32_t e, uint32_t o_pos0, uint32_t o_ctx, uint32_t o_woff, uint32_t o_L, uint32_t o_strand,
                       SeedCands& cands) -> uint32_t {
      if (o_pos0 == 0) return kNoAllele;  // L == k: the seed states are the final states (general kernel)
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(v.seed_ent) + e);
      if (!(raw.y & 0x80000000u)) return seed_state_cands(v, super_c, b.strand_words(o_strand, o_woff), o_L, e, cands);
      uint32_t m = (raw.y >> 24) & 0x7Fu;
      m = m < o_pos0 ? m : o_pos0;
      if (m && (((o_ctx ^ raw.y) & ((0xFFFFFFFFu << (24 - 2 * m)) & 0xFFFFFFu)) != 0)) return 0;
      cands.p[0] = raw.x;
      cands.w0[0] = o_pos0 | (K_SCAN << 28);
      return 1;
    };
    uint32_t incl = ns;  // inclusive warp scan of the entry counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(full, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t total = __shfl_sync(full, incl, 31);
    uint32_t general = 0;  // bit per strand of the round: needs the general kernel
    if (total < kSeedStrandMode) {
      for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        // owner of task t = first lane whose inclusive count exceeds t
        uint32_t lo_l = 0, hi_l = 31;
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const uint32_t mid = (lo_l + hi_l) >> 1;
          const uint32_t val = __shfl_sync(full, incl, mid);
          if (val > t) hi_l = mid;
          else lo_l = mid + 1;
        }
        const uint32_t owner = hi_l;
        const uint32_t o_incl = __shfl_sync(full, incl, owner), o_ns = __shfl_sync(full, ns, owner),
                       o_s0 = __shfl_sync(full, lk.s0, owner), o_n0 = __shfl_sync(full, lk.n0, owner),
                       o_s1 = __shfl_sync(full, lk.s1, owner), o_L = __shfl_sync(full, my_L, owner),
                       o_woff = __shfl_sync(full, my_woff, owner), o_pos0 = __shfl_sync(full, my_pos0, owner),
                       o_ctx = __shfl_sync(full, my_ctx, owner);
        const uint32_t o_strand = 2 * b.read_begin + i0 + owner;
        const uint32_t te = t - (o_incl - o_ns);  // entry te of the owner strand: bucket 0 first, then its own bucket
        const uint32_t e = te < o_n0 ? o_s0 + te : o_s1 + (te - o_n0);
        SeedCands cands;
        uint32_t cnt = 0;
        if (t < total) cnt = examine(e, o_pos0, o_ctx, o_woff, o_L, o_strand, cands);
        const bool bad = cnt == kNoAllele;
        general |= __reduce_or_sync(full, bad ? (1u << owner) : 0u);
        seed_push(stage, pre, v, cands, bad ? 0u : cnt, o_strand, e, lane);
      }
    } else {
      for (uint32_t owner = 0; owner < 32; ++owner) {
        const uint32_t o_ns = __shfl_sync(full, ns, owner);
        if (o_ns == 0) continue;
        const uint32_t o_s0 = __shfl_sync(full, lk.s0, owner), o_n0 = __shfl_sync(full, lk.n0, owner),
                       o_s1 = __shfl_sync(full, lk.s1, owner), o_L = __shfl_sync(full, my_L, owner),
                       o_woff = __shfl_sync(full, my_woff, owner), o_pos0 = __shfl_sync(full, my_pos0, owner),
                       o_ctx = __shfl_sync(full, my_ctx, owner);
        const uint32_t o_strand = 2 * b.read_begin + i0 + owner;
        for (uint32_t c0 = 0; c0 < o_ns; c0 += 32) {
          const uint32_t te = c0 + lane;
          const uint32_t e = te < o_n0 ? o_s0 + te : o_s1 + (te - o_n0);
          SeedCands cands;
          uint32_t cnt = 0;
          if (c0 + lane < o_ns) cnt = examine(e, o_pos0, o_ctx, o_woff, o_L, o_strand, cands);
          const bool bad = cnt == kNoAllele;
          if (__any_sync(full, bad)) general |= 1u << owner;
          seed_push(stage, pre, v, cands, bad ? 0u : cnt, o_strand, e, lane);
        }
      }
    }
    if ((general >> lane) & 1u) send_to_general(pre, strand);
  }
  seed_flush(stage, pre, lane);
}

void launch_seed(const IndexView& v, const BatchView& b, const SearchOut& o, const SeedOut& pre, cudaStream_t st) {
  uint32_t work = 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  static const int smode = getenv("GQ_SEED_GRID") ? atoi(getenv("GQ_SEED_GRID")) : 0;
  uint32_t blocks = smode ? (work + 255) / 256 : min((work + 255) / 256, 148u * 8u);
  uint32_t n_super = (v.n >> kSuperShift) + 1;
  if (n_super <= (uint32_t)kMaxSuperSmem)
    seed_kernel<true><<<blocks, 256, n_super * 16, st>>>(v, b, o, pre, n_super);
  else
    seed_kernel<false><<<blocks, 256, 0, st>>>(v, b, o, pre, 0);
}

// Verify pass: one thread per candidate; a few warp-convergent walk iterations decide whether the candidate
// is real (see fast_verified). Survivors are copied, densely, behind the candidate records (second half of
// the pool) for the text kernel.
__global__ void __launch_bounds__(256) verify_kernel(IndexView v, BatchView b, SeedOut pre, uint32_t* surv_rec,
                                                     uint32_t* n_verified) {
  const uint32_t n = min(*pre.n_surv, pre.cap);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t full = 0xFFFFFFFFu;
  // survivors of a round are written one round later, so the list allocation (one atomic per warp) is in
  // flight during the next round instead of stalling this one
  uint4 pend_rec = make_uint4(0, 0, 0, 0);
  uint32_t pend_base = 0, pend_rank = 0, pend_mask = 0;
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    FastLane f;
    f.result = FAST_DEAD;
    f.ln.state = LS_IDLE;
    f.ln.pos = 0;
    uint4 rec = make_uint4(kNoAllele, 0, 0, 0);
    if (i < n) rec = __ldg(reinterpret_cast<const uint4*>(pre.rec) + i);
    if (rec.x != kNoAllele) fast_begin<false>(f, v, b, pre, i);
    const uint32_t pos0 = f.ln.pos;
    for (uint32_t it = 0; it < kVerifyIters && __any_sync(full, !fast_verified(f, pos0)); ++it) {
      if (!fast_verified(f, pos0) && f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
      if (!fast_verified(f, pos0) && f.ln.state == LS_EV_TSCAN) fast_event<false>(f, v);
    }
    const bool alive = rec.x != kNoAllele && fast_alive(f);
    const uint32_t mm = __ballot_sync(full, alive);
    if (pend_mask) {  // flush the previous round
      const uint32_t base = __shfl_sync(full, pend_base, 0);
      if ((pend_mask >> lane) & 1u) reinterpret_cast<uint4*>(surv_rec)[base + pend_rank] = pend_rec;
    }
    if (mm && lane == 0) pend_base = atomicAdd(n_verified, (uint32_t)__popc(mm));
    pend_mask = mm;
    pend_rank = __popc(mm & ((1u << lane) - 1u));
    pend_rec = rec;
  }
  if (pend_mask) {
    const uint32_t base = __shfl_sync(full, pend_base, 0);
    if ((pend_mask >> lane) & 1u) reinterpret_cast<uint4*>(surv_rec)[base + pend_rank] = pend_rec;
  }
}

// Text kernel: one thread per verified candidate, 32 per warp round. The walk is ONE warp-convergent loop
// — text step for the lanes in text mode, then the jump for the lanes that reached a marker — so lanes meet
// again every iteration; the round ends with a convergent emission phase (one pool allocation and one
// mapped-list allocation per warp).
template <bool DFS>
__global__ void __launch_bounds__(256) text_kernel(IndexView v, BatchView b, SearchOut o, SeedOut pre) {
  const uint32_t n = min(*pre.n_surv, pre.cap);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t full = 0xFFFFFFFFu;
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    FastLane f;
    f.result = FAST_DEAD;
    f.ln.state = LS_IDLE;
    f.ln.strand = 0;
    if (i < n) fast_begin<true>(f, v, b, pre, i);
    if (DFS) {  // nested PRGs: branching walks (fast_jump), forks on a private stack
      FastForks fk;
      forks_init(fk, i < n);
      while (__any_sync(full, fk.active)) {
        if (fk.active && fast_running(f) && f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
        if (fk.active && fast_running(f) && f.ln.state == LS_EV_TSCAN) fast_event_dfs(f, fk, v);
        if (fk.active && !fast_running(f)) fast_branch_end(f, fk, v);  // next fork, or the candidate's outcome
      }
    } else {  // straight walks: a jump that is not pre-resolved hands the strand to the general kernel
      while (__any_sync(full, fast_running(f))) {
        if (fast_running(f) && f.ln.state == LS_TEXT) lane_text_step(f.ln, v);
        if (fast_running(f) && f.ln.state == LS_EV_TSCAN) fast_event<true>(f, v);
      }
    }
    uint32_t words = i < n ? fast_outcome(f, v) : 0;
    const uint32_t strand = f.ln.strand;
    if (f.result == FAST_BAIL) send_to_general(pre, strand);
    // A finished candidate claims its strand; a second one (reads in repeats) makes the strand the general
    // kernel's. Only claim winners take pool space and a mapped-list slot: a strand can have many finished
    // candidates, the list holds one entry per strand.
    bool emit = f.result == FAST_MAPPED && fast_claim(pre, strand);
    if (!emit) words = 0;
    uint32_t incl = words;  // pool space for the round: warp scan + one atomic
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(full, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t total = __shfl_sync(full, incl, 31);
    // pool space and mapped-list slots of the round: two atomics, issued together (one round trip, not two)
    const uint32_t mm = __ballot_sync(full, emit);
    uint32_t base = 0, mbase = 0;
    if (total && lane == 31) base = atomicAdd(o.pool_used, total);
    if (mm && lane == 0) mbase = atomicAdd(o.n_mapped, (uint32_t)__popc(mm));
    base = __shfl_sync(full, base, 31);
    mbase = __shfl_sync(full, mbase, 0);
    if (emit) {
      const uint32_t off = base + incl - words;
      if (off + words > o.pool_cap) {  // final-state pool full: re-run after the host has grown it
        o.status[strand] = ST_OVERFLOW;
        o.overflow_list[atomicAdd(o.n_overflow, 1u)] = strand;
      } else {
        fast_emit(f, v, o, off);
        o.status[strand] = ST_MAPPED;
        atomicOr(pre.surv_cnt + strand, kSurvListed);
      }
      // (the slot of an overflowed strand stays in the list: the coverage pass skips entries whose status is not
      // ST_MAPPED, and the re-run records the strand from its own list)
      o.mapped_list[mbase + __popc(mm & ((1u << lane) - 1u))] = strand;
    }
  }
}

void launch_text(const IndexView& v, const BatchView& b, const SearchOut& o, const SeedOut& pre, uint32_t* surv_rec,
                 uint32_t* n_verified, cudaStream_t st, cudaEvent_t between) {
  uint32_t work = 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  uint32_t blocks = min((work + 255) / 256, 148u * 8u);
  // one block per 256 (verify) / 128 (text) candidate slots instead of a grid-stride loop over a resident grid: the
  // block scheduler then hands out the walks dynamically (walk lengths vary; measured 0.073 -> 0.065 and
  // 0.282 -> 0.259 ms per 1 M reads at config 2); blocks beyond the device-side candidate count exit at once
  static const int vmode = getenv("GQ_VERIFY_GRID") ? atoi(getenv("GQ_VERIFY_GRID")) : 1;
  static const int tmode = getenv("GQ_TEXT_GRID") ? atoi(getenv("GQ_TEXT_GRID")) : 2;
  const uint32_t cand_cap = pre.cap;
  auto grid_of = [&](int mode, uint32_t& thr) {
    thr = mode == 2 ? 128u : (mode == 3 ? 64u : 256u);
    if (mode == 0) return blocks;
    const uint64_t items = min((uint64_t)cand_cap, (uint64_t)work * 2);  // candidates: a device-side count
    return (uint32_t)max((uint64_t)1, (items + thr - 1) / thr);
  };
  uint32_t vthr, tthr;
  const uint32_t vblocks = grid_of(vmode, vthr), tblocks = grid_of(tmode, tthr);
  verify_kernel<<<vblocks, vthr, 0, st>>>(v, b, pre, surv_rec, n_verified);
  if (between) cudaEventRecord(between, st);
  SeedOut ver = pre;  // the text kernel's candidates are the verified ones
  ver.rec = surv_rec;
  ver.n_surv = n_verified;
  if (v.any_nested) text_kernel<true><<<tblocks, tthr, 0, st>>>(v, b, o, ver);
  else text_kernel<false><<<tblocks, tthr, 0, st>>>(v, b, o, ver);
}

#ifdef GQ_DEBUG_COUNTERS
__device__ unsigned long long g_dbg[32];
#define DBG(i, v) do { if (lane == 0) atomicAdd(&g_dbg[i], (unsigned long long)(v)); } while (0)
#else
#define DBG(i, v) do { } while (0)
#endif

template <bool SUPER_SMEM>
__global__ void __launch_bounds__(kSearchThreads, GQ_SEARCH_MIN_BLOCKS)
    search_kernel(IndexView v, BatchView b, SearchOut o, uint32_t* arena, uint32_t arena_words,
                  const uint32_t* list, uint32_t n_list, uint32_t n_super_smem, uint32_t rf_thresh,
                  uint32_t ev_thresh, uint32_t wait_max, uint32_t leave, const uint32_t* n_list_dev) {
  extern __shared__ __align__(128) uint32_t s_super[];  // n_super_smem x 16 B (dynamic: keeps 5 CTAs/SM)
  __shared__ alignas(8) uint64_t s_bar;
  // work items: a list of strands (the seed pass's general list, overflow re-runs) or the whole slice
  const uint32_t* work_list = list;
  const uint32_t work = list ? (n_list_dev ? *n_list_dev : n_list) : 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  if (SUPER_SMEM) tma_stage_super(s_super, v.super_cnt, n_super_smem * 16u, &s_bar);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t strand0 = 2 * b.read_begin;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  bool work_left = work > 0;  // warp-uniform: strands are handed out by one global counter
  uint32_t* my_arena = arena + (size_t)tid * arena_words;
  Lane ln;
  ln.state = LS_IDLE;
  ln.strand = kNoAllele;
  const uint32_t full = 0xFFFFFFFFu;
  while (true) {
    // one REDUX.ADD gives all class populations: 6-bit fields idle | wide | scan | pop | top
    const uint32_t st = ln.state;
    const uint32_t field = st == LS_IDLE ? 1u
                           : st == LS_RUNW ? (1u << 6)
                           : (st == LS_EV_SCAN || st == LS_EV_WIDE || st == LS_EV_TSCAN) ? (1u << 12)
                           : st == LS_EV_POP ? (1u << 18)
                           : st == LS_EV_TOP ? (1u << 24) : 0u;
    const uint32_t votes = __reduce_add_sync(full, field);
    const uint32_t c_idle = votes & 63u, n_wide = (votes >> 6) & 63u, n_scan = (votes >> 12) & 63u,
                   n_pop = (votes >> 18) & 63u, n_top = (votes >> 24) & 63u;
    if (c_idle == 32 && !work_left) break;
    const uint32_t n_idle = work_left ? c_idle : 0;
    const uint32_t n_run = 32u - (c_idle + n_wide + n_scan + n_pop + n_top);
    // service a class when enough lanes wait in it; when nothing can step, or too many lanes wait in
    // total, service the most populated class
    const uint32_t waiting = n_idle + n_wide + n_scan + n_pop + n_top;
    const bool force = n_run == 0 || waiting >= wait_max;
    const uint32_t big = max(max(max(n_idle, n_wide), max(n_scan, n_pop)), n_top);
    DBG(0, 1); DBG(1, n_run); DBG(2, waiting);
    if (n_wide && (n_wide >= ev_thresh || (force && n_wide == big))) {
      DBG(3, 1); DBG(4, n_wide);
      if (ln.state == LS_RUNW) lane_step_wide(ln, v, SUPER_SMEM ? (const uint32_t*)s_super : v.super_cnt);
      continue;
    }
    if (n_scan && (n_scan >= ev_thresh || (force && n_scan == big))) {
      DBG(5, 1); DBG(6, n_scan);
      if (ln.state == LS_EV_SCAN || ln.state == LS_EV_WIDE || ln.state == LS_EV_TSCAN) lane_event_scan(ln, v, o);
      continue;
    }
    if (n_top && (n_top >= ev_thresh || (force && n_top == big))) {
      DBG(7, 1); DBG(8, n_top);
      if (ln.state == LS_EV_TOP) lane_event_top(ln, v, o);
      continue;
    }
    if (n_pop && (n_pop >= ev_thresh || (force && n_pop == big))) {
      DBG(9, 1); DBG(10, n_pop);
      if (ln.state == LS_EV_POP) lane_event_pop(ln, o);
      continue;
    }
    if (n_idle && (n_idle >= rf_thresh || (force && n_idle == big))) {
      DBG(11, 1); DBG(12, n_idle);
      const uint32_t idle = __ballot_sync(full, ln.state == LS_IDLE);
      // scarce work (the seed pass's general list is usually short) is spread over the warps: the kernel is
      // a chain of dependent loads per strand, so one strand per warp finishes sooner than 32
      const uint32_t take = min(c_idle, max(1u, work / n_warps));
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(o.work_counter, take);
      base = __shfl_sync(full, base, 0);
      if (ln.state == LS_IDLE) {
        const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
        uint32_t i = base + rank;
        if (rank < take && i < work) {
          const uint32_t strand = work_list ? work_list[i] : strand0 + i;
          lane_refill(ln, v, b, o, strand, my_arena, arena_words);
        }
      }
      work_left = base + take < work;
      continue;
    }
    // hot loop: keep stepping until `leave` lanes have dropped out of LS_RUN (or none is left); the
    // per-step control cost is one ballot + popc instead of the full vote above
    const uint32_t n_run0 = n_run;
    const uint32_t stay = n_run0 > leave ? n_run0 - leave : 0;
    while (true) {
#pragma unroll
      for (int u = 0; u < kHotUnroll; ++u) {
        // width-1 states walk the PRG text (up to 16 bases per step); a state that just became one suffix
        // wide first fetches its text position
        if (ln.state == LS_TEXT) lane_text_step(ln, v);
        else if (ln.state == LS_RUN) lane_to_text(ln, v);
      }
      const uint32_t n_run = __popc(__ballot_sync(full, ln.state == LS_RUN || ln.state == LS_TEXT));
      DBG(13, 1); DBG(14, n_run);
      if (n_run <= stay) break;
    }
  }
}

// k-mer filter for failed strands (all_read_kmers_occur_in_index, quasimap.cpp:212-225): one LANE per strand, in
// a flat warp loop with refill. A k-mer code is a bit-field of the packed read (classify_strand), so a lane
// slides a 64-bit window over its strand — shift, mask, one bit test per k-mer — and stops at the first absent
// k-mer (a random 150-mer misses the index after ~65 of its 141 10-mers at config 2); a lane that is done takes
// the next unclassified strand of the warp's current group of 32 statuses, so lanes stay busy whatever the
// strands' lengths. The probes are random 4-byte reads, so when the 4^k-bit set fits (k <= 10: 128 KB) each CTA
// keeps a copy in shared memory (one persistent 1024-thread CTA per SM); the reverse strand reverse-complements
// windows are probed, as they are, against the set indexed by reverse-complement codes: one launch per orientation.
// [v1 gave a whole warp to each strand: 181 warp instructions per strand, mostly per-strand set-up; this form
// needs ~25.]
constexpr uint32_t kClassifySmemBytes = 160 * 1024;

template <bool SMEM>
__global__ void __launch_bounds__(SMEM ? 1024 : 256)
    classify_kernel(IndexView v, BatchView b, SearchOut o, const uint32_t* list, uint32_t n_list, uint32_t bits_words,
                    const uint32_t* __restrict__ g_bits) {
  extern __shared__ __align__(16) uint32_t s_bits[];
  if (SMEM) {
    const uint4* src = reinterpret_cast<const uint4*>(g_bits);
    uint4* dst = reinterpret_cast<uint4*>(s_bits);
    for (uint32_t q = threadIdx.x; q < (bits_words + 3) / 4; q += blockDim.x) dst[q] = __ldg(src + q);
    __syncthreads();
  }
  const uint32_t n = list ? n_list : 2 * (b.read_end - b.read_begin);
  const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
  const uint32_t full = 0xFFFFFFFFu;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t k = v.k, mask = (k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
  // The warp's share: groups of 32 consecutive strands, round-robin over the warps. A group is fetched by the
  // whole warp (lane = position: status, length, word offset, first two packed words — coalesced) one group
  // AHEAD of its use, so handing a strand to a free lane is a few shuffles and never waits for memory.
  struct Group {
    uint32_t strand, L, woff, w0, w1;
    bool need;
  };
  uint32_t g_load = warp;  // next group to fetch
  auto fetch = [&](Group& G) {
    const uint32_t idx = g_load * 32u + lane;
    g_load += n_warps;
    G.strand = idx < n ? (list ? list[idx] : 2 * b.read_begin + idx) : 0u;
    G.need = idx < n && o.status[G.strand] == ST_UNCLASSIFIED;
    G.L = G.woff = G.w0 = G.w1 = 0;
    if (G.need) {
      G.L = b.len[G.strand >> 1];
      G.woff = b.word_off[G.strand >> 1];
      const uint32_t* gw = b.strand_words(G.strand, G.woff);
      G.w0 = __ldg(gw);
      if (G.L > 16) G.w1 = __ldg(gw + 1);
    }
  };
  Group cur, nxt;
  uint32_t cur_first = warp * 32u;  // index of the current group's first strand
  fetch(cur);
  fetch(nxt);
  uint32_t pend = __ballot_sync(full, cur.need);  // strands of the current group not yet handed out
  // lane state: the strand's k-mers i .. n_kmers-1 are still to be probed; `win` holds the packed bases from k-mer i on
  bool active = false;
  uint32_t strand = 0, i = 0, n_kmers = 0, n_words = 0;
  uint32_t wlo = 0, whi = 0;
  const uint32_t* w = nullptr;
  auto probe = [&](uint32_t code) -> uint32_t {  // bit `code` of the presence set, in bit 0 of the result
    const uint32_t word = SMEM ? s_bits[code >> 5] : __ldg(g_bits + (code >> 5));
    return word >> (code & 31u);
  };
  while (true) {
    uint32_t idle = __ballot_sync(full, !active);
    if (__popc(idle) >= 8 || idle == full) {
      while (idle) {  // hand unclassified strands to the free lanes
        if (!pend) {
          if (cur_first >= n) break;  // the warp's share is used up
          cur = nxt;
          cur_first += n_warps * 32u;
          pend = __ballot_sync(full, cur.need);
          fetch(nxt);
          continue;
        }
        const uint32_t n_take = min(__popc(idle), __popc(pend));
        const uint32_t r = __popc(idle & lt);
        const bool take = !active && r < n_take;
        const uint32_t src = take ? (uint32_t)__fns(pend, 0, r + 1) & 31u : 0u;  // r-th pending strand of the group
        const uint32_t s_new = __shfl_sync(full, cur.strand, src), L_new = __shfl_sync(full, cur.L, src),
                       woff_new = __shfl_sync(full, cur.woff, src), w0_new = __shfl_sync(full, cur.w0, src),
                       w1_new = __shfl_sync(full, cur.w1, src);
        if (take) {
          strand = s_new;
          w = b.strand_words(s_new, woff_new);
          n_words = (L_new + 15) >> 4;
          n_kmers = L_new - k + 1;  // unclassified strands have L >= k
          wlo = w0_new;
          whi = w1_new;
          i = 0;
          active = true;
        }
        for (uint32_t j = 0; j < n_take; ++j) {  // the n_take lowest pending strands / free lanes are served
          pend &= pend - 1;
          idle &= idle - 1;
        }
      }
      if (!__any_sync(full, active)) break;
    }
    // up to 8 k-mers per lane and iteration: four at a time without a test in between while at least four are left
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (active) {
        bool done;
        uint32_t present;
        if (n_kmers - i >= 4) {
          present = probe(wlo & mask) & probe(__funnelshift_r(wlo, whi, 2) & mask) &
                    probe(__funnelshift_r(wlo, whi, 4) & mask) & probe(__funnelshift_r(wlo, whi, 6) & mask);
          wlo = __funnelshift_r(wlo, whi, 8);
          whi >>= 8;
          i += 4;
          if ((i & 15u) == 0) {  // 16 bases consumed: the low word is the next packed word, bring in the one after
            const uint32_t nw = (i >> 4) + 1;
            whi = nw < n_words ? __ldg(w + nw) : 0u;
          }
          done = i == n_kmers;
        } else {  // the last one to three k-mers of the strand
          present = 1u;
          for (; i < n_kmers; ++i) {
            present &= probe(wlo & mask);
            wlo = __funnelshift_r(wlo, whi, 2);
            whi >>= 2;
          }
          done = true;
        }
        if (!(present & 1u) || done) {
          o.status[strand] = (present & 1u) ? ST_NO_EXTENSION : ST_MISSING_KMER;
          active = false;
        }
      }
    }
  }
}

void launch_classify(const IndexView& v, const BatchView& b, const SearchOut& o, const uint32_t* list, uint32_t n_list,
                     cudaStream_t st) {
  uint32_t work = list ? n_list : 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  const uint64_t bits_words = ((1ull << (2 * v.k)) + 31) / 32;
  const uint64_t bytes = ((bits_words + 3) / 4) * 16;
  // both strand orientations in one pass: the reverse strands exist as packed reads of their own (packed_rc)
  if (bytes <= kClassifySmemBytes && work >= 148u * 32u * 4u) {
    cudaFuncSetAttribute(classify_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClassifySmemBytes);
    classify_kernel<true><<<148, 1024, bytes, st>>>(v, b, o, list, n_list, (uint32_t)bits_words, v.kmer_bits);
  } else {
    uint32_t blocks = min((work + 255) / 256, 148u * 8u);
    classify_kernel<false><<<blocks, 256, 0, st>>>(v, b, o, list, n_list, (uint32_t)bits_words, v.kmer_bits);
  }
}

// uint16 view of the accumulators for gq_coverage_fetch: allele_sum wraps mod 65536 (allele_sum.cpp:41),
// per-base coverage saturates at 65535 (allele_base.cpp:239)
__global__ void fetch_kernel(const uint32_t* __restrict__ allele_sum, uint32_t n_alleles,
                             const uint32_t* __restrict__ per_base, uint32_t n_per_base, uint16_t* __restrict__ out) {
  const uint32_t n = n_alleles + n_per_base;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = i < n_alleles ? (uint16_t)(allele_sum[i] & 0xFFFFu) : (uint16_t)min(per_base[i - n_alleles], 65535u);
}

void launch_fetch(const uint32_t* allele_sum, uint32_t n_alleles, const uint32_t* per_base, uint32_t n_per_base,
                  uint16_t* out, cudaStream_t st) {
  const uint32_t n = n_alleles + n_per_base;
  if (n == 0) return;
  fetch_kernel<<<min((n + 255) / 256, 148u * 8u), 256, 0, st>>>(allele_sum, n_alleles, per_base, n_per_base, out);
}

// ------------------------------------------------------------------------------------------------
// Sparse multi-allele groups for the multi-GPU exchange (comm.cu): every occupied, counted table slot becomes a
// record [slot, count, n, alleles...] in `words` (+ its start in `rec_off`); n_out = {records, words}.
__global__ void groups_export_kernel(CoverageView c, uint32_t* __restrict__ words, uint32_t words_cap,
                                     uint32_t* __restrict__ rec_off, uint32_t rec_cap, uint32_t* __restrict__ n_out) {
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < c.gtab_cap; h += gridDim.x * blockDim.x) {
    const uint32_t cur = c.gtab[h], cnt = c.gcount[h];
    if (!cur || !cnt) continue;
    const uint32_t* rec = c.gpool + (cur - 1);
    const uint32_t n = rec[1];
    const uint32_t i = atomicAdd(n_out, 1u), w = atomicAdd(n_out + 1, 3u + n);
    if (i >= rec_cap || w + 3 + n > words_cap) continue;  // sized from a first pass: cannot happen
    rec_off[i] = w;
    words[w] = rec[0];
    words[w + 1] = cnt;
    words[w + 2] = n;
    for (uint32_t j = 0; j < n; ++j) words[w + 3 + j] = rec[2 + j];
  }
}

// add the records of every OTHER rank (gathered, `stride_*` apart per rank) into this GPU's table
__global__ void groups_import_kernel(CoverageView c, const uint32_t* __restrict__ words, uint32_t stride_words,
                                     const uint32_t* __restrict__ rec_off, uint32_t stride_recs,
                                     const uint32_t* __restrict__ counts /* 2 per rank */, uint32_t n_ranks,
                                     uint32_t my_rank) {
  for (uint32_t r = 0; r < n_ranks; ++r) {
    if (r == my_rank) continue;
    const uint32_t n_rec = counts[2 * r];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
      const uint32_t* rec = words + (size_t)r * stride_words + rec_off[(size_t)r * stride_recs + i];
      const uint32_t h = grouped_find_or_insert(c, rec[0], rec + 3, rec[2], 1);
      if (h != kNoAllele) atomicAdd(c.gcount + h, rec[1]);  // table pre-sized by the host: h is always valid
    }
  }
}

void launch_groups_export(const CoverageView& c, uint32_t* words, uint32_t words_cap, uint32_t* rec_off, uint32_t rec_cap,
                          uint32_t* n_out, cudaStream_t st) {
  groups_export_kernel<<<min((c.gtab_cap + 255) / 256, 148u * 8u), 256, 0, st>>>(c, words, words_cap, rec_off, rec_cap, n_out);
}

void launch_groups_import(const CoverageView& c, const uint32_t* words, uint32_t stride_words, const uint32_t* rec_off,
                          uint32_t stride_recs, const uint32_t* counts, uint32_t n_ranks, uint32_t my_rank, cudaStream_t st) {
  groups_import_kernel<<<148 * 4, 256, 0, st>>>(c, words, stride_words, rec_off, stride_recs, counts, n_ranks, my_rank);
}

int search_kernel_smem_limit_superblocks() { return kMaxSuperSmem; }

void debug_counters(unsigned long long* out32) {
#ifdef GQ_DEBUG_COUNTERS
  cudaMemcpyFromSymbol(out32, g_dbg, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(g_dbg, z, sizeof z);
#else
  for (int i = 0; i < 32; ++i) out32[i] = 0;
#endif
}

void launch_search(const IndexView& v, const BatchView& b, const SearchOut& o, uint32_t* arena,
                   uint32_t arena_words, uint32_t n_threads, const uint32_t* list, uint32_t n_list,
                   bool super_in_smem, uint32_t rf_thresh, uint32_t ev_thresh, cudaStream_t st, uint32_t leave_opt,
                   uint32_t wait_opt, const uint32_t* n_list_dev) {
  uint32_t work = list ? n_list : 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  uint32_t blocks = (min(work, n_threads) + kSearchThreads - 1) / kSearchThreads;
  uint32_t n_super = (v.n >> kSuperShift) + 1;
  uint32_t n_super_smem = (super_in_smem && n_super <= (uint32_t)kMaxSuperSmem) ? n_super : 0;
  rf_thresh = max(1u, min(32u, rf_thresh));
  ev_thresh = max(1u, min(32u, ev_thresh));
  uint32_t wait_max = wait_opt ? wait_opt : min(32u, rf_thresh + ev_thresh);
  uint32_t leave = leave_opt ? leave_opt : max(1u, ev_thresh / 2);
  if (n_super_smem)
    search_kernel<true><<<blocks, kSearchThreads, n_super_smem * 16, st>>>(v, b, o, arena, arena_words, list, n_list, n_super_smem,
                                                           rf_thresh, ev_thresh, wait_max, leave, n_list_dev);
  else
    search_kernel<false><<<blocks, kSearchThreads, 0, st>>>(v, b, o, arena, arena_words, list, n_list, 0, rf_thresh,
                                                            ev_thresh, wait_max, leave, n_list_dev);
}

// ------------------------------------------------------------------------------------------------
// Coverage recording, one thread per mapped strand.
// ------------------------------------------------------------------------------------------------
// Nested PRG, strand with ONE final state of ONE occurrence, for the 32 strands of a warp side by side (the route
// record_strand takes for such a strand, restated warp-convergently: lanes left to themselves in its data-dependent
// loops never met again — 1.6 active lanes per instruction). Every loop below advances all lanes together, one
// path element / one graph node / one base per iteration. Must be called by all 32 lanes; `strand` = kNoAllele for a
// lane without work. Returns true for a lane whose strand needs the general route (nothing recorded for it).
__device__ __forceinline__ bool record_single_nested_warp(const IndexView& v, const BatchView& b, const SearchOut& o,
                                                          const CoverageView& c, uint32_t strand) {
  const uint32_t full = 0xFFFFFFFFu;
  bool valid = strand != kNoAllele;
  bool general = false;
  StateRec st{0, 0, 0, 0, nullptr, nullptr};
  uint32_t L = 0;
  if (valid) {
    st = parse_rec(o.pool + o.st_off[strand]);
    L = b.len[strand >> 1];
    if (!(st.nt | st.ng)) valid = false;  // a path-less state records nothing (coverage_common.cpp:137-148)
    else if (st.lo != st.hi) valid = false, general = true;
  }
  constexpr uint32_t kLoc = 40;
  uint32_t used_l[kLoc], base_l[kLoc], loci_l[2 * kLoc];
  LocusLists l1;
  l1.loci = loci_l, l1.base = base_l, l1.used = used_l;
  l1.n_loci = l1.n_base = l1.n_used = 0;
  l1.cap = kLoc;
  l1.overflow = false;
  uint32_t pos0 = 0, nid0 = 0;
  if (valid) {  // where the read starts; a read that starts inside a site: its allele from the node (LocusFinder :52-74)
    pos0 = __ldg(v.sa + st.lo);
    nid0 = __ldg(v.pos2node + pos0);
    if (st.ng > 0) {
      const uint32_t seed_site = st.G[2 * (st.ng - 1)], al = (uint32_t)v.nodes[nid0].allele;
      add_locus(l1, seed_site, al);
      assign_nested(v, l1, seed_site, al);
    }
  }
  const uint32_t nt = valid ? st.nt : 0;
  for (uint32_t j = 0; __any_sync(full, j < nt); ++j)
    if (j < nt) assign_nested(v, l1, st.T[2 * j], st.T[2 * j + 1]);  // LocusFinder :76-83
  if (valid && l1.overflow) valid = false, general = true;
  const uint32_t n_loci = valid ? l1.n_loci : 0;
  for (uint32_t i = 0; __any_sync(full, i < n_loci); ++i)
    if (i < n_loci) {  // at most one allele per site: single-allele groups
      const uint32_t ai = c.allele_off[(loci_l[2 * i] - 5) >> 1] + loci_l[2 * i + 1];
      gq_red_add(c.allele_sum + ai, 1u);
      gq_red_add(c.grouped_single + ai, 1u);
    }
  Trav t;
  t.v = &v;
  t.cur = nid0;
  t.remaining = L;
  t.T = st.T;
  t.ti = st.nt;
  t.first = true;
  t.start_pos = 0;
  t.end_pos = 0;
  t.bad = false;
  if (valid) {
    const Node& nd0 = v.nodes[nid0];
    t.start_pos = nd0.len > 1 ? pos0 - nd0.start : 0;
  }
  bool alive = valid;
  while (__any_sync(full, alive)) {
    uint32_t cnt = 0;
    uint32_t* pb = nullptr;
    if (alive) {
      alive = t.next();
      if (alive) {
        const Node& nd = v.nodes[t.cur];
        if (nd.len != 0 && nd.cov_off != kNoAllele) {
          cnt = t.end_pos - t.start_pos + 1;
          pb = c.per_base + nd.cov_off + t.start_pos;
        }
      }
    }
    for (uint32_t x = 0; __any_sync(full, x < cnt); ++x)
      if (x < cnt) gq_red_add(pb + x, 1u);  // allele_base.cpp:221-296
  }
  if (valid && t.bad) atomicOr(c.error_flags, 2u);
  return general;
}

#ifndef GQ_HEAVY_STATES
#define GQ_HEAVY_STATES 4
#endif
constexpr uint32_t kHeavyStates = GQ_HEAVY_STATES;  // nested PRGs: strands with this many final states get a warp each

// mode 0: every strand of the work list, handed out one by one (non-nested PRGs: nearly all strands take the table
//         route; also the overflow re-runs).
// mode 1: nested PRGs, first pass — the single-state strands (nine in ten), a warp's 32 strands side by side on the
//         same route (LocusFinder + one forward walk each); strands with several states are only collected ...
// mode 2: ... and recorded in a second pass over that list, one by one: their cost ranges over orders of magnitude
//         (two states to hundreds), and lanes sharing a static batch would wait for the slowest.
__global__ void __launch_bounds__(256)
    coverage_kernel(IndexView v, BatchView b, SearchOut o, CoverageView c, uint32_t* arena, uint32_t arena_words,
                    const uint32_t* list, uint32_t n_list, const uint32_t* n_list_dev, uint32_t* overflow_list,
                    uint32_t* n_overflow, uint32_t* work_counter, uint32_t mode, uint32_t* multi_list,
                    uint32_t* n_multi, uint32_t* heavy_list, uint32_t* n_heavy) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t* my_arena = arena + (size_t)tid * arena_words;
  const uint32_t* work_list = list ? list : o.mapped_list;
  const uint32_t n = list ? (n_list_dev ? *n_list_dev : n_list) : *o.n_mapped;
  if (mode == 1) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i0 = tid & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
      const uint32_t i = i0 + lane;
      uint32_t strand = i < n ? work_list[i] : kNoAllele;
      if (strand != kNoAllele && o.status[strand] != ST_MAPPED) strand = kNoAllele;  // kNoAllele: slot of a lost claim
      const uint32_t n_st = strand != kNoAllele ? o.st_count[strand] : 0u;
      bool multi = n_st > 1;
      if (record_single_nested_warp(v, b, o, c, multi ? kNoAllele : strand)) multi = true;
      const bool heavy = multi && n_st >= kHeavyStates;  // a warp each (coverage_multi_kernel)
      multi = multi && !heavy;
      const uint32_t mm = __ballot_sync(0xFFFFFFFFu, multi), mh = __ballot_sync(0xFFFFFFFFu, heavy);
      if (mm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(n_multi, (uint32_t)__popc(mm));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (multi) multi_list[base + __popc(mm & ((1u << lane) - 1u))] = strand;
      }
      if (mh) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(n_heavy, (uint32_t)__popc(mh));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (heavy) heavy_list[base + __popc(mh & ((1u << lane) - 1u))] = strand;
      }
    }
    return;
  }
  for (uint32_t i = atomicAdd(work_counter, 1u); i < n; i = atomicAdd(work_counter, 1u)) {
    uint32_t strand = work_list[i];
    if (strand == kNoAllele || o.status[strand] != ST_MAPPED) continue;  // kNoAllele: slot of a lost claim
    if (!record_strand(v, b, o, c, strand, my_arena, arena_words)) overflow_list[atomicAdd(n_overflow, 1u)] = strand;
  }
}

// ------------------------------------------------------------------------------------------------
// Nested PRGs, strands with MANY final states (kHeavyStates or more; up to hundreds where alleles coincide): one
// WARP per strand. The general route of record_strand (record_general) runs such a strand on one thread — 10^4 to
// 10^5 dependent loads, and the kernel lasts as long as its heaviest strand. Here the states are spread over the
// lanes: class keys (LocusFinder per state) in parallel, class representatives and the seeded pick by comparison
// across lanes, then the states of the chosen class traverse the graph in parallel into per-warp shared-memory
// tables — the set of loci (64-bit keys, atomicCAS) and the per-node hulls (atomicCAS on the node id, atomicMin /
// atomicMax on the range: PbCovRecorder's min/max per node commute). Nothing is committed before every table has
// proved large enough; a strand that does not fit (more than kMultiStates states, a key of more than kMultiKey
// level-0 sites, full tables) goes to the overflow list and is recorded by the one-thread route.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMultiWarps = 8, kMultiHull = 512, kMultiLoci = 256, kMultiKey = 16, kMultiStates = 1024;
constexpr uint32_t kMultiSmemWordsPerWarp = 3 * kMultiHull + 2 * kMultiLoci;

__global__ void __launch_bounds__(32 * kMultiWarps)
    coverage_multi_kernel(IndexView v, BatchView b, SearchOut o, CoverageView c, uint32_t* scratch, uint32_t warp_words,
                          const uint32_t* list, const uint32_t* n_list_dev, uint32_t* overflow_list, uint32_t* n_overflow,
                          uint32_t* work_counter) {
  extern __shared__ __align__(16) uint32_t s_multi[];
  const uint32_t full = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t* s_hull = s_multi + wib * kMultiSmemWordsPerWarp;                       // triples (node, start, end)
  unsigned long long* s_loci = reinterpret_cast<unsigned long long*>(s_hull + 3 * kMultiHull);  // site << 32 | allele + 1
  uint32_t* w_scr = scratch + (size_t)warp * warp_words;
  const uint32_t n = *n_list_dev;
  while (true) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(work_counter, 1u);
    i = __shfl_sync(full, i, 0);
    if (i >= n) break;
    const uint32_t strand = list[i];
    if (o.status[strand] != ST_MAPPED) continue;
    const uint32_t ns = o.st_count[strand];
    const uint32_t* recs = o.pool + o.st_off[strand];
    const uint32_t L = b.len[strand >> 1];
    auto give_back = [&]() {  // the one-thread route takes it (re-run pass of the host)
      if (lane == 0) overflow_list[atomicAdd(n_overflow, 1u)] = strand;
    };
    if (ns > kMultiStates || warp_words < ns * (3 + kMultiKey) + 3 * kMultiLoci) {
      give_back();
      continue;
    }
    uint32_t* rec_off = w_scr;          // ns
    uint32_t* key_len = w_scr + ns;     // ns; 0xFFFFFFFF = path-less state
    uint32_t* rep = w_scr + 2 * ns;     // ns
    uint32_t* keys = w_scr + 3 * ns;    // ns x kMultiKey
    if (lane == 0) {
      uint32_t off = 0;
      for (uint32_t j = 0; j < ns; ++j) {
        rec_off[j] = off;
        off += 4 + 2 * recs[off + 2] + 2 * recs[off + 3];
      }
    }
    __syncwarp();
    constexpr uint32_t kLoc = 48;
    uint32_t used_l[kLoc], base_l[kLoc], loci_l[2 * kLoc];
    LocusLists ll;
    ll.loci = loci_l, ll.base = base_l, ll.used = used_l;
    ll.cap = kLoc;
    // ---- class keys, non-variant mappings (MappingInstanceSelector, coverage_common.cpp:85-164) ----
    bool bad = false;
    uint32_t nonvar = 0, npath = 0;
    for (uint32_t j = lane; j < ns; j += 32) {
      const StateRec st = parse_rec(recs + rec_off[j]);
      if (!(st.nt | st.ng)) {
        key_len[j] = 0xFFFFFFFFu;
        nonvar += st.hi - st.lo + 1;
        continue;
      }
      ++npath;
      ll.n_loci = ll.n_base = ll.n_used = 0;
      ll.overflow = false;
      locus_finder(v, st, ll);
      if (ll.overflow || ll.n_base > kMultiKey) {
        bad = true;
        key_len[j] = 0;
        continue;
      }
      sort_u32(ll.base, ll.n_base);
      for (uint32_t q = 0; q < ll.n_base; ++q) keys[j * kMultiKey + q] = ll.base[q];
      key_len[j] = ll.n_base;
    }
    __syncwarp();
    nonvar = __reduce_add_sync(full, nonvar);
    npath = __reduce_add_sync(full, npath);
    if (__any_sync(full, bad)) {
      give_back();
      continue;
    }
    if (npath == 0) continue;
    // ---- class representatives: the first state with the same key; number of classes ----
    uint32_t ncls = 0;
    for (uint32_t j = lane; j < ns; j += 32) {
      uint32_t r = 0xFFFFFFFFu;
      if (key_len[j] != 0xFFFFFFFFu) {
        r = j;
        for (uint32_t q = 0; q < j; ++q)
          if (key_len[q] != 0xFFFFFFFFu && cmp_key(keys + q * kMultiKey, key_len[q], keys + j * kMultiKey, key_len[j]) == 0) {
            r = q;
            break;
          }
        if (r == j) ++ncls;
      }
      rep[j] = r;
    }
    __syncwarp();
    ncls = __reduce_add_sync(full, ncls);
    // ---- random_select_entry (:97-107): seeded pick among non-variant mappings + classes ----
    const uint32_t total = nonvar + ncls;
    uint32_t pick = 1;
    if (total != 1) {
      if (lane == 0) pick = uniform_1_to(b.seeds[strand >> 1], total);
      pick = __shfl_sync(full, pick, 0);
    }
    if (pick <= nonvar) continue;
    const uint32_t want = pick - nonvar - 1;
    uint32_t chosen = 0xFFFFFFFFu;  // the representative with exactly `want` smaller ones (std::map order)
    for (uint32_t j = lane; j < ns; j += 32) {
      if (rep[j] != j) continue;
      uint32_t less = 0;
      for (uint32_t q = 0; q < ns; ++q)
        if (q != j && rep[q] == q && cmp_key(keys + q * kMultiKey, key_len[q], keys + j * kMultiKey, key_len[j]) < 0) ++less;
      if (less == want) chosen = j;
    }
    chosen = __reduce_min_sync(full, chosen);
    if (chosen == 0xFFFFFFFFu) {
      if (lane == 0) atomicOr(c.error_flags, 2u);
      continue;
    }
    // ---- loci of the chosen class + per-node hulls (PbCovRecorder, allele_base.cpp:221-296) ----
    for (uint32_t q = lane; q < kMultiHull; q += 32) {
      s_hull[3 * q] = kNoAllele;
      s_hull[3 * q + 1] = 0xFFFFFFFFu;
      s_hull[3 * q + 2] = 0;
    }
    for (uint32_t q = lane; q < kMultiLoci; q += 32) s_loci[q] = 0ull;
    __syncwarp();
    uint32_t n_hull = 0;  // this lane's insertions (the table is declared full at half its slots)
    auto hull_put = [&](uint32_t node, uint32_t s, uint32_t e) {
      if (v.nodes[node].len == 0) return;  // process_Node :282-287
      for (uint32_t h = (node * 2654435761u) >> (32 - 9), probes = 0;; h = (h + 1) & (kMultiHull - 1), ++probes) {
        uint32_t cur = s_hull[3 * h];
        if (cur == kNoAllele) cur = atomicCAS(&s_hull[3 * h], kNoAllele, node);
        if (cur == kNoAllele) ++n_hull;
        if (cur == kNoAllele || cur == node) {
          atomicMin(&s_hull[3 * h + 1], s);
          atomicMax(&s_hull[3 * h + 2], e);
          return;
        }
        if (probes > kMultiHull / 2) {
          bad = true;
          return;
        }
      }
    };
    for (uint32_t j = lane; j < ns; j += 32) {
      if (rep[j] != chosen) continue;
      const StateRec st = parse_rec(recs + rec_off[j]);
      ll.n_loci = ll.n_base = ll.n_used = 0;
      ll.overflow = false;
      locus_finder(v, st, ll);
      if (ll.overflow) {
        bad = true;
        continue;
      }
      for (uint32_t q = 0; q < ll.n_loci; ++q) {  // set union over the class
        const unsigned long long key = ((unsigned long long)loci_l[2 * q] << 32) | (unsigned long long)(loci_l[2 * q + 1] + 1u);
        uint32_t h = (uint32_t)((loci_l[2 * q] * 2654435761u + loci_l[2 * q + 1] * 40503u) >> 24) & (kMultiLoci - 1);
        for (uint32_t probes = 0;; h = (h + 1) & (kMultiLoci - 1), ++probes) {
          unsigned long long cur = s_loci[h];
          if (cur == 0ull) cur = atomicCAS(&s_loci[h], 0ull, key);
          if (cur == 0ull || cur == key) break;
          if (probes > kMultiLoci / 2) {
            bad = true;
            break;
          }
        }
      }
      bool first = true;
      for (uint32_t occ = st.lo;; ++occ) {
        const uint32_t pos = __ldg(v.sa + occ), nid = __ldg(v.pos2node + pos);
        Trav t;
        t.v = &v;
        t.cur = nid;
        t.remaining = L;
        t.T = st.T;
        t.ti = st.nt;
        t.first = true;
        const Node& nd = v.nodes[nid];
        t.start_pos = nd.len > 1 ? pos - nd.start : 0;
        t.end_pos = 0;
        t.bad = false;
        if (first) {  // only the first occurrence of a state gets the full traversal (:246-270)
          first = false;
          while (t.next()) hull_put(t.cur, t.start_pos, t.end_pos);
        } else if (t.next())
          hull_put(t.cur, t.start_pos, t.end_pos);
        if (t.bad) atomicOr(c.error_flags, 2u);
        if (occ == st.hi) break;
      }
    }
    __syncwarp();
    if (__reduce_add_sync(full, n_hull) > kMultiHull / 2) bad = true;
    if (__any_sync(full, bad)) {
      give_back();
      continue;
    }
    // ---- commit: lane 0 sorts the loci and settles the groups (multi-allele groups first find their table slots:
    //      a full group table gives the strand back before any counter is touched) ----
    uint32_t ok = 1;
    if (lane == 0) {
      uint32_t n_loci = 0;
      uint32_t* sl = keys + ns * kMultiKey;  // sorted (site, allele) pairs, then the group slots
      for (uint32_t q = 0; q < kMultiLoci; ++q) {
        const unsigned long long k64 = s_loci[q];
        if (k64 == 0ull) continue;
        const uint32_t s0 = (uint32_t)(k64 >> 32), a0 = (uint32_t)k64 - 1u;
        uint32_t j = n_loci++;
        while (j > 0 && (sl[2 * (j - 1)] > s0 || (sl[2 * (j - 1)] == s0 && (int32_t)sl[2 * (j - 1) + 1] > (int32_t)a0))) {
          sl[2 * j] = sl[2 * (j - 1)];
          sl[2 * j + 1] = sl[2 * (j - 1) + 1];
          --j;
        }
        sl[2 * j] = s0;
        sl[2 * j + 1] = a0;
      }
      uint32_t* slots = sl + 2 * kMultiLoci;
      uint32_t ng = 0;
      for (uint32_t q = 0; q < n_loci && ok;) {
        uint32_t e = q;
        while (e < n_loci && sl[2 * e] == sl[2 * q]) ++e;
        if (e - q > 1) {
          const uint32_t h = grouped_find_or_insert(c, (sl[2 * q] - 5) >> 1, sl + 2 * q + 1, e - q, 2);
          if (h == kNoAllele) ok = 0;
          slots[ng++] = h;
        }
        q = e;
      }
      if (ok) {
        ng = 0;
        for (uint32_t q = 0; q < n_loci;) {
          const uint32_t site = sl[2 * q], ao = c.allele_off[(site - 5) >> 1];
          uint32_t e = q;
          while (e < n_loci && sl[2 * e] == site) {
            gq_red_add(c.allele_sum + ao + sl[2 * e + 1], 1u);  // allele_sum.cpp:31-43
            ++e;
          }
          if (e - q == 1) gq_red_add(c.grouped_single + ao + sl[2 * q + 1], 1u);
          else gq_red_add(c.gcount + slots[ng++], 1u);  // grouped_allele_counts.cpp:17-49
          q = e;
        }
      }
    }
    ok = __shfl_sync(full, ok, 0);
    if (!ok) {
      give_back();
      continue;
    }
    for (uint32_t q = lane; q < kMultiHull; q += 32) {
      const uint32_t node = s_hull[3 * q];
      if (node == kNoAllele) continue;
      const Node& nd = v.nodes[node];
      if (nd.cov_off == kNoAllele) continue;
      for (uint32_t x = s_hull[3 * q + 1]; x <= s_hull[3 * q + 2]; ++x) gq_red_add(c.per_base + nd.cov_off + x, 1u);
    }
    __syncwarp();
  }
}

void launch_coverage(const IndexView& v, const BatchView& b, const SearchOut& o, const CoverageView& c,
                     uint32_t* arena, uint32_t arena_words, uint32_t n_threads, const uint32_t* list,
                     uint32_t n_list, uint32_t* overflow_list, uint32_t* n_overflow, uint32_t* work_counter,
                     cudaStream_t st, uint32_t* multi_list, uint32_t* n_multi, uint32_t* heavy_list, uint32_t* n_heavy) {
  uint32_t work = list ? n_list : 2 * (b.read_end - b.read_begin);
  if (work == 0) return;
  uint32_t blocks = (min(work, n_threads) + 255) / 256;
  cudaMemsetAsync(work_counter, 0, 4, st);
  if (v.any_nested && multi_list && !list) {
    // nested PRG, main pass: single-state strands (warp-convergent), then the strands with many states (a warp
    // each), then the rest (a thread each, handed out one by one)
    coverage_kernel<<<blocks, 256, 0, st>>>(v, b, o, c, arena, arena_words, nullptr, 0, nullptr, overflow_list, n_overflow,
                                            work_counter, 1u, multi_list, n_multi, heavy_list, n_heavy);
    const uint32_t smem = kMultiWarps * kMultiSmemWordsPerWarp * 4;
    cudaFuncSetAttribute(coverage_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t mblocks = min(blocks * 256u / (32u * kMultiWarps), 148u * 3u);
    coverage_multi_kernel<<<max(mblocks, 1u), 32 * kMultiWarps, smem, st>>>(v, b, o, c, arena, 32u * arena_words, heavy_list,
                                                                              n_heavy, overflow_list, n_overflow, work_counter);
    cudaMemsetAsync(work_counter, 0, 4, st);
    coverage_kernel<<<blocks, 256, 0, st>>>(v, b, o, c, arena, arena_words, multi_list, work, n_multi, overflow_list,
                                            n_overflow, work_counter, 2u, nullptr, nullptr, nullptr, nullptr);
    return;
  }
  coverage_kernel<<<blocks, 256, 0, st>>>(v, b, o, c, arena, arena_words, list, n_list, nullptr, overflow_list, n_overflow,
                                          work_counter, 0u, nullptr, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
__global__ void stats_kernel(const uint8_t* status, const uint32_t* len, uint32_t n_reads, unsigned long long* stats) {
  unsigned long long loc[5] = {0, 0, 0, 0, 0};
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
    loc[0] += 2;
    for (int s = 0; s < 2; ++s) {
      uint8_t stt = status[2 * r + s];
      if (stt == ST_SKIPPED) loc[1] += 1;
      else if (stt == ST_MISSING_KMER) loc[2] += 1;
      else if (stt == ST_NO_EXTENSION) loc[3] += 1;
      else if (stt == ST_MAPPED) loc[4] += 1;
    }
  }
  for (int i = 0; i < 5; ++i) {
    unsigned long long x = loc[i];
    for (int d = 16; d; d >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, d);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(stats + i, x);
  }
}

// batch counters -> the handle's totals; with `small` given only when the batch needs no overflow re-run (the host
// then counts again after the re-runs and commits unconditionally)
__global__ void stats_commit_kernel(const unsigned long long* __restrict__ batch, unsigned long long* __restrict__ total,
                                    const uint32_t* __restrict__ small) {
  if (small && (small[1] | small[2])) return;
  if (threadIdx.x < 5) total[threadIdx.x] += batch[threadIdx.x];
}

void launch_stats_commit(const unsigned long long* batch, unsigned long long* total, const uint32_t* small, cudaStream_t st) {
  stats_commit_kernel<<<1, 32, 0, st>>>(batch, total, small);
}

void launch_stats(const uint8_t* status, const uint32_t* len, uint32_t n_reads, unsigned long long* stats,
                  cudaStream_t st) {
  if (n_reads == 0) return;
  uint32_t blocks = min((n_reads + 255) / 256, 148u * 4u);
  stats_kernel<<<blocks, 256, 0, st>>>(status, len, n_reads, stats);
}

}  // namespace gq
