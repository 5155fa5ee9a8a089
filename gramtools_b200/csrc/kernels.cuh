// kernels.cuh — launch interface of the sm_100a kernels (definitions in kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "gq_core.cuh"

namespace gq {

enum StrandStatus : uint8_t {
  ST_SKIPPED = 0, ST_MISSING_KMER = 1, ST_NO_EXTENSION = 2, ST_MAPPED = 3,
  ST_UNCLASSIFIED = 4,  // search produced no state: the classify kernel decides 1 vs 2
  ST_OVERFLOW = 255
};

struct BatchView {               // one batch of reads, resident in HBM
  const uint32_t* packed;        // 2-bit base codes, 16 per word, every read starts on a word
  const uint32_t* packed_rc;     // the reverse complements in the same layout (revcomp_kernel, once per batch slice):
                                 //  both strands of a read are walked by the same instructions
  const uint32_t* word_off;      // n_reads + 1
  const uint32_t* len;           // n_reads (0 = skipped read)
  const uint32_t* seeds;         // n_reads (selection seed, shared by both strands)
  uint32_t n_reads;              // reads in the whole batch (arrays above are indexed by read id)
  uint32_t read_begin, read_end; // the slice of the batch this launch works on (H2D/compute pipelining)
  // the packed words of a strand (even = the read as stored, odd = its reverse complement)
  GQ_HD const uint32_t* strand_words(uint32_t strand, uint32_t woff) const {
    return ((strand & 1u) ? packed_rc : packed) + woff;
  }
};

struct SearchOut {
  uint8_t* status;      // 2 * n_reads
  uint32_t* st_off;     // 2 * n_reads: word offset of the strand's records in `pool`
  uint32_t* st_words;   // 2 * n_reads: words used
  uint32_t* st_count;   // 2 * n_reads: number of final states
  uint32_t* pool;       // final-state records: [lo, hi, nt, ng, (site,allele)*nt, (site,0xFFFFFFFF)*ng]
  uint32_t pool_cap;
  uint32_t* pool_used;      // bump pointer
  uint32_t* overflow_list;  // strands that ran out of arena / pool
  uint32_t* n_overflow;
  uint32_t* mapped_list;    // strands with >= 1 final state, in completion order (coverage work list)
  uint32_t* n_mapped;
  uint32_t* work_counter;   // dynamic work distribution of the search kernel
  uint32_t* listed;         // per strand flags (SeedOut::surv_cnt) or nullptr: kSurvListed = already on mapped_list
};

// Output of the seed pass (seed_kernel): one candidate per suffix of a seed state that agrees with the read on
// its left context — a width-1 search state the text kernel walks through the PRG text — plus the strands that
// need the general kernel.
struct SeedOut {
  uint32_t* rec;        // 4 words per candidate: {strand, k-mer state index, text position, pos | kind << 28}
  uint32_t cap;         // candidate records available
  uint32_t* n_surv;     // bump pointer
  uint32_t* surv_cnt;   // per strand: finished candidates (low 16 bits) | kSurvGeneral | kSurvListed
  uint32_t* gen_list;   // strands for the general search kernel (several finished candidates, seed states
  uint32_t* n_gen;      //  that cannot be split, jumps the text kernel does not take)
};

struct CoverageView {
  uint32_t* allele_sum;      // per (slot, allele): allele_off[slot] + allele
  uint32_t* per_base;        // flat in-bubble bases, PRG order
  uint32_t* grouped_single;  // per (slot, allele): reads whose allele set at the site is {allele}
  // multi-allele sets: open-addressing table of offsets into gpool records [slot, n, alleles...]
  uint32_t* gtab;            // gtab_cap entries: 0 = empty, else record offset + 1
  uint32_t* gcount;          // gtab_cap counters
  uint32_t gtab_cap;         // power of two
  uint32_t* gpool;
  uint32_t gpool_cap;
  uint32_t* gpool_used;
  unsigned long long* stats; // all_reads, skipped, missing_kmer, no_extension, exact_mapped
  const uint32_t* allele_off;
  uint32_t* error_flags;     // bit0: grouped table/pool full, bit1: inconsistent traversal
};

// bases (uint8 1..4, concatenated, device) -> packed 2-bit words for reads [r0, r1). Read r starts at word
// (offsets[r] >> 4) + r: word-aligned and non-overlapping without a prefix sum (<= 1 word wasted per read).
void launch_pack(const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1, uint32_t* word_off,
                 uint32_t* packed, uint32_t* len, cudaStream_t st);

// packed_rc for reads [b.read_begin, b.read_end) from packed (reverse_complement_read, quasimap.cpp:273-298)
void launch_revcomp(const BatchView& b, uint32_t* packed_rc, cudaStream_t st);

// Seed pass over the strands of the slice b.read_begin..b.read_end.
void launch_seed(const IndexView& v, const BatchView& b, const SearchOut& o, const SeedOut& pre, cudaStream_t st);

// Fast path: verify pass over the candidates (survivors copied to surv_rec, same capacity as pre.rec, counted
// in *n_verified), then the text kernel over the survivors: strands with one finished candidate get their
// final state here; the rest is appended to pre.gen_list.
void launch_text(const IndexView& v, const BatchView& b, const SearchOut& o, const SeedOut& pre, uint32_t* surv_rec,
                 uint32_t* n_verified, cudaStream_t st, cudaEvent_t between = nullptr);

// General search kernel. list == nullptr: every strand of the slice; otherwise the listed strands (n_list of
// them, or *n_list_dev when that pointer is given). Strands are seeded from the k-mer index in the kernel.
void launch_search(const IndexView& v, const BatchView& b, const SearchOut& o, uint32_t* arena,
                   uint32_t arena_words, uint32_t n_threads, const uint32_t* list, uint32_t n_list,
                   bool super_in_smem, uint32_t rf_thresh, uint32_t ev_thresh, cudaStream_t st,
                   uint32_t leave_opt = 0, uint32_t wait_opt = 0, const uint32_t* n_list_dev = nullptr);

// list == nullptr: the strands in o.mapped_list[0, *o.n_mapped); otherwise the listed strands.
void launch_coverage(const IndexView& v, const BatchView& b, const SearchOut& o, const CoverageView& c,
                     uint32_t* arena, uint32_t arena_words, uint32_t n_threads, const uint32_t* list,
                     uint32_t n_list, uint32_t* overflow_list, uint32_t* n_overflow, uint32_t* work_counter,
                     cudaStream_t st, uint32_t* multi_list = nullptr, uint32_t* n_multi = nullptr,
                     uint32_t* heavy_list = nullptr, uint32_t* n_heavy = nullptr);

// k-mer filter for the strands whose search found nothing (status ST_UNCLASSIFIED -> 1 or 2)
void launch_classify(const IndexView& v, const BatchView& b, const SearchOut& o, const uint32_t* list, uint32_t n_list,
                     cudaStream_t st);

void launch_stats(const uint8_t* status, const uint32_t* len, uint32_t n_reads, unsigned long long* stats,
                  cudaStream_t st);

void launch_stats_commit(const unsigned long long* batch, unsigned long long* total, const uint32_t* small, cudaStream_t st);

// out[0, n_alleles) = allele_sum mod 65536, out[n_alleles, n_alleles + n_per_base) = min(per_base, 65535)
void launch_fetch(const uint32_t* allele_sum, uint32_t n_alleles, const uint32_t* per_base, uint32_t n_per_base,
                  uint16_t* out, cudaStream_t st);

// multi-GPU exchange of the sparse multi-allele groups (comm.cu)
void launch_groups_export(const CoverageView& c, uint32_t* words, uint32_t words_cap, uint32_t* rec_off, uint32_t rec_cap,
                          uint32_t* n_out, cudaStream_t st);
void launch_groups_import(const CoverageView& c, const uint32_t* words, uint32_t stride_words, const uint32_t* rec_off,
                          uint32_t stride_recs, const uint32_t* counts, uint32_t n_ranks, uint32_t my_rank, cudaStream_t st);

int search_kernel_smem_limit_superblocks();

}  // namespace gq
