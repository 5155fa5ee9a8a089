/* gq.h — C ABI of libgq.so, the B200 quasimap back-end.
 *
 * gramtools has no plugin/FFI interface for this path; the seam this library replaces is the C++
 * function
 *     QuasimapReadsStats gram::quasimap_reads(const GenotypeParams&, const KmerIndex&,
 *                                             const PRG_Info&, ReadStats&)
 * (libgramtools/include/genotype/quasimap/quasimap.hpp:29-32, called once from
 * libgramtools/src/genotype/genotype.cpp:45-46) together with the index loading that precedes it
 * (libgramtools/src/genotype/genotype.cpp:38-40: load_prg_info + kmer_index::load).
 * INTEGRATION.md shows the few lines a maintainer adds to genotype.cpp to call these entry points.
 *
 * Conventions: every function returns 0 on success and a negative code on failure, with text in
 * gq_last_error(); the caller owns all host buffers; the library owns device memory until
 * gq_index_destroy(); calls on one handle must be serialised by the caller; one handle per GPU.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GQ_H
#define GQ_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gq_index gq_index;

typedef struct gq_layout {
  uint64_t n_symbols;      /* |prg| */
  uint64_t sa_size;        /* |prg| + 1 */
  uint32_t kmer_size;
  uint32_t n_sites;        /* number of variant sites (coverage_graph.bubble_map.size()) */
  uint32_t n_site_slots;   /* (max site id - 5)/2 + 1; == n_sites when ids are contiguous */
  uint32_t is_nested;      /* coverage_graph.is_nested */
  uint64_t n_alleles;      /* total alleles = length of the flat allele_sum vector */
  uint64_t n_per_base;     /* bases inside sites = length of the flat per-base vector */
  uint64_t n_kmer_states;  /* SearchStates stored in the k-mer index */
  uint64_t device_bytes;   /* HBM held by the index */
} gq_layout;

/* Replaces load_prg_info() + kmer_index::load() (genotype.cpp:38-40): builds FM-index, masks,
 * coverage graph and the all-k-mers index from the linearised PRG (the `prg` file of gram_dir, raw
 * LE uint32, linearised_prg.cpp:8-45) and uploads them to GPU `device`. */
int gq_index_build(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device, gq_index** out);
int gq_index_destroy(gq_index* idx);
int gq_index_describe(const gq_index* idx, gq_layout* out);
/* allele_off[n_site_slots + 1]: offset of each site's alleles in the flat allele_sum vector. */
int gq_index_allele_offsets(const gq_index* idx, uint64_t* allele_off);
/* per-base layout: for every site slot and allele, (offset,len) of the allele's bases in the flat
 * per-base vector — only meaningful for non-nested PRGs (allele_base.cpp:10-14). 2 * n_alleles. */
int gq_index_per_base_layout(const gq_index* idx, uint64_t* off_len);

/* Replaces handle_reads_buffer() (quasimap.cpp:82-118) for one batch: `bases` are the reads encoded
 * as the reference does (1..4 = A,C,G,T; encode_dna_bases, utils.cpp:72-81), concatenated; read i is
 * bases[read_offsets[i] .. read_offsets[i+1]); a read containing a non-ACGT character is passed as
 * an EMPTY read (as the reference's encoder produces) and counts as skipped. seeds[i] is the
 * selection seed of read i, used for both strands (quasimap.cpp:143-157). Coverage accumulates in
 * the handle across calls. Host buffers in, nothing out: results stay on the device. */
int gq_map_batch(gq_index* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                 const uint32_t* seeds);

/* Split form of gq_map_batch for callers that keep a batch resident in HBM: upload once ... */
int gq_batch_upload(gq_index* idx, const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads,
                    const uint32_t* seeds);
/* ... then map the resident batch (kernels only, no host<->device copies of reads). */
int gq_map_resident(gq_index* idx);

/* Per-strand results of the LAST batch, for parity checks. status: 2*n_reads
 * bytes (0 skipped, 1 missing k-mer, 2 no extension, 3 mapped). State records per strand:
 * [lo, hi, nt, ng, (site,allele)*nt, (site,0xFFFFFFFF)*ng] concatenated, unordered within a strand. */
int gq_batch_status(gq_index* idx, uint8_t* status);
int gq_batch_states_size(gq_index* idx, uint64_t* n_words);
int gq_batch_states(gq_index* idx, uint64_t* strand_off /* 2*n_reads+1 */, uint32_t* strand_count /* 2*n_reads */,
                    uint32_t* words);

/* Replaces the Coverage members of QuasimapReadsStats (coverage/types.hpp:39-43) + the per-base
 * vectors inside the graph nodes. uint16 semantics of the reference are applied here: allele_sum and
 * grouped counts wrap mod 65536 (allele_sum.cpp:41, grouped_allele_counts.cpp:47), per-base counts
 * saturate at 65535 (allele_base.cpp:239). Any pointer may be NULL. stats = {all_reads, skipped,
 * missing_kmer, no_extension, exact_mapped} (quasimap.hpp:17-24). */
int gq_coverage_fetch(gq_index* idx, uint16_t* allele_sum, uint16_t* per_base, uint64_t stats[5]);
/* Grouped allele counts as flat records [site_slot, count(uint16 wrapped), n, allele ids...],
 * sorted by (site_slot, allele ids). Call with words == NULL to get the size. */
int gq_coverage_grouped(gq_index* idx, uint32_t* words, uint64_t* n_words);
int gq_coverage_reset(gq_index* idx);
/* ReadStats::compute_coverage_depth (read_stats.cpp:119-160) on the coverage accumulated so far:
 * per level-0 site the mean per-base coverage of the allele path with the highest grouped count (or
 * that count for a direct deletion); out = {mean, variance} and counts = {num_sites_noCov,
 * num_sites_total}. Host arithmetic in double, like the reference (not part of the GPU hot path). */
int gq_read_depth_stats(gq_index* idx, double out[2], uint64_t counts[2]);

/* Multi-GPU: raw device pointers of the additive uint32 accumulators so the host layer can run one
 * NCCL all-reduce(sum) over them (reads are sharded across GPUs, index replicated).
 * counters = [allele_sum (n_alleles) | grouped_single (n_alleles) | per_base (n_per_base)] in one
 * contiguous allocation of n_counters uint32; stats = 5 x uint64. */
int gq_coverage_device_ptrs(gq_index* idx, void** counters, uint64_t* n_counters, void** stats);
/* Sparse multi-allele groups of this GPU as records [site_slot, count, n, alleles...] (raw uint32
 * counts, not wrapped) and the inverse: add such records into this handle (merge after gather). */
int gq_coverage_groups_export(gq_index* idx, uint32_t* words, uint64_t* n_words);
int gq_coverage_groups_import(gq_index* idx, const uint32_t* words, uint64_t n_words, int replace);

/* Run the kernels on a caller-owned CUDA stream (e.g. torch's current stream) so that the caller's
 * CUDA events bracket them. NULL = the library's own stream. */
int gq_set_stream(gq_index* idx, void* cuda_stream);
/* Tunables (name, value): arena_words, n_threads, cov_threads, super_in_smem, rf_thresh, ev_thresh (general
 * kernel); seed_pass (0: every strand through the general kernel), seed_recs_per_read (candidate pool);
 * chunk_reads, tail_chunk_reads (slices of the pipelined host path), resident_slices, overlap_classify.
 * Results never depend on them. */
int gq_set_option(gq_index* idx, const char* name, int64_t value);
/* Counters of the last gq_map_* call: [0] kernel launches, [1] overflow re-run strands, [2] search-phase ms
 * and [3] classify + coverage ms (CUDA events; single-slice runs only), [4] pool words used, [5] H2D bytes,
 * [6] all kernels ms (CUDA events on the caller's stream), [7] host ms spent enqueueing the call */
int gq_last_run_info(gq_index* idx, double info[8]);

const char* gq_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GQ_H */
