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
/* The k-mer index as the reference keeps it in gram_dir: the four sdsl::int_vector files `kmers`, `kmers_stats`,
 * `sa_intervals`, `paths` (written by kmer_index::dump, build/kmer_index/dump.cpp:27-141; read by kmer_index::load,
 * load.cpp:11-173). gq_kmer_index_dump writes them from a built index (what `gram build` leaves behind);
 * gq_index_build_from_gram_dir is gq_index_build with the k-mer searches replaced by loading those files (what
 * `gram genotype` does, genotype.cpp:40) — FM-index, graph and masks are still rebuilt from `prg`. The sdsl
 * serialisation is restated in kmer_index_files.cpp. The contents of the four vectors are pinned by the literals of the
 * reference's tests/build/kmer_index/test_dump_and_load.cpp; the byte serialisation is PARITY UNPINNED (no SDSL in this
 * environment, no serialised fixture in the reference): checked by hand-written golden bytes and round trips only. */
int gq_kmer_index_dump(const gq_index* idx, const char* gram_dir);
int gq_index_build_from_gram_dir(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, int device,
                                 const char* gram_dir, gq_index** out);
/* The whole index as one file of this back-end's own format (checksummed): gq_index_save after a build (`gram build`
 * writes gram_dir/gq_index), gq_index_load instead of gq_index_build — nothing is rebuilt, the arrays are uploaded as
 * they were. A file that is truncated, corrupted, of another version or record layout is refused. gq_index_prg
 * returns the PRG the index was built from (prg_out may be NULL to ask for the length): callers compare it with the
 * gram_dir/prg at hand before trusting a stored index. */
int gq_index_save(const gq_index* idx, const char* path);
int gq_index_load(const char* path, int device, gq_index** out);
int gq_index_prg(const gq_index* idx, uint32_t* prg_out, uint64_t* n_symbols);
/* The suffix array alone — what sdsl::construct computes first (make_data_structures.cpp:9-33) — built on GPU
 * `device` by prefix doubling (one radix sort per round; 28 bytes of HBM per symbol: a 3.3e9-symbol whole-genome
 * PRG takes 92 GB, one B200). sa_out: n_symbols + 1 entries (the sentinel suffix first), 32-bit text positions;
 * n_symbols < 2^32 - 3. rounds (may be NULL) = sort rounds taken. gq_index_build uses the same builder. */
int gq_suffix_array(const uint32_t* prg, uint64_t n_symbols, int device, uint32_t* sa_out, int* rounds);
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

/* 2-bit packed form of the same batch (16 bases per uint32, base j of a word at bits [2j, 2j+2), codes 0..3 =
 * A,C,G,T; read r occupies ceil(len[r] / 16) words from word_off[r]; word_off has n_reads + 1 entries, the last
 * one = total words). A third of the host-to-device bytes of the unpacked form and no packing pass on the GPU:
 * this is what a reader thread that runs ahead of the GPU should hand over (replaces the per-read
 * std::vector<uint8_t> of sequence_read/seqread.hpp:94-180 + quasimap.cpp:126-140). A read emptied by the
 * encoder has len 0. */
int gq_map_batch_packed(gq_index* idx, const uint32_t* packed, const uint32_t* word_off, const uint32_t* len,
                        uint64_t n_reads, const uint32_t* seeds);
/* Host-side packers producing that layout (OpenMP over reads, n_threads >= 1): from encoded bases 1..4
 * (encode_dna_bases, utils.cpp:72-81) or straight from sequence text (ACGTacgt; a read with any other character
 * becomes empty, utils.cpp:13-47,83-92). `packed` must hold gq_packed_words() words, word_off n_reads + 1,
 * len n_reads. Read r is placed at word (read_offsets[r] >> 4) + r. */
int gq_packed_words(const uint64_t* read_offsets, uint64_t n_reads, uint64_t* n_words);
int gq_pack_reads(const uint8_t* bases, const uint64_t* read_offsets, uint64_t n_reads, uint32_t* packed,
                  uint32_t* word_off, uint32_t* len, int n_threads);
int gq_pack_ascii(const char* text, const uint64_t* read_offsets, uint64_t n_reads, uint32_t* packed,
                  uint32_t* word_off, uint32_t* len, int n_threads);

/* Page-locked host memory for the buffers handed to gq_map_batch / gq_map_batch_packed (H2D copies of pageable
 * memory are staged by the driver and do not overlap the kernels), and the number of CUDA devices. */
int gq_host_alloc(uint64_t bytes, void** out);
int gq_host_free(void* p);
int gq_device_count(int* n);

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

/* ---- multi-GPU: reads sharded over the GPUs, index replicated, ONE exchange at the end -------
 * The reference shares one Coverage object between OpenMP threads (quasimap.cpp:90-118); here every GPU
 * accumulates its own and the totals are formed once, when all reads are mapped:
 *   - one process per GPU: rank 0 calls gq_comm_unique_id(), hands the id to the other ranks by any means
 *     (MPI, torch.distributed, a file), every rank calls gq_comm_init(handle, id, rank, n_ranks);
 *   - one process, several GPUs: gq_index_clone() the index onto each device, gq_comm_init_all(handles, n).
 * gq_coverage_allreduce() then sums, IN PLACE and over NCCL (NVLink / NVSwitch), the dense uint32 accumulators
 * (allele_sum | grouped singles | per-base) and the five counters, and merges the sparse multi-allele groups
 * (exported, all-gathered and inserted on the device). Afterwards every handle holds the totals of the whole job
 * and gq_coverage_fetch / gq_coverage_grouped apply the uint16 semantics to them. Call it once per job (or after
 * gq_coverage_reset + mapping): reducing totals again would add the other ranks' share twice. Collective: every
 * rank must call it. NCCL is loaded at run time (libnccl.so.2); without it these entry points fail, the rest of
 * the library works. */
#define GQ_COMM_ID_BYTES 128
int gq_comm_unique_id(uint8_t id[GQ_COMM_ID_BYTES]);
int gq_comm_init(gq_index* idx, const uint8_t id[GQ_COMM_ID_BYTES], int rank, int n_ranks);
int gq_comm_init_all(gq_index** per_gpu, int n);
int gq_comm_destroy(gq_index* idx);
int gq_comm_version(int* nccl_version);
int gq_coverage_allreduce(gq_index* idx);
int gq_coverage_allreduce_all(gq_index** per_gpu, int n); /* all handles of one gq_comm_init_all, one call */
/* A second handle on another GPU of this process from an index that is already built (no rebuild on the host). */
int gq_index_clone(const gq_index* src, int device, gq_index** out);

/* Raw device pointers of the additive uint32 accumulators, for callers that run their own collective.
 * counters = [allele_sum (n_alleles) | grouped_single (n_alleles) | per_base (n_per_base)] in one
 * contiguous allocation of n_counters uint32; stats = 5 x uint64. */
int gq_coverage_device_ptrs(gq_index* idx, void** counters, uint64_t* n_counters, void** stats);
/* Sparse multi-allele groups of this GPU as records [site_slot, count, n, alleles...] (raw uint32
 * counts, not wrapped) and the inverse: add such records into this handle (merge after gather). */
int gq_coverage_groups_export(gq_index* idx, uint32_t* words, uint64_t* n_words);
int gq_coverage_groups_import(gq_index* idx, const uint32_t* words, uint64_t n_words, int replace);

/* ---- genotyping: what commands::genotype::run does once quasimap has returned (genotype.cpp:68-118) ----
 * LevelGenotyper (infer/level_genotyping/{runner,model,probabilities}.cpp, infer/allele_extracter.cpp) on the coverage
 * of one sample, then the three files of geno_dir/genotype: genotyped.json (output_specs/make_json.cpp),
 * personalised_reference.fasta (personalised_reference.cpp) and genotyped.vcf.gz (output_specs/make_vcf.cpp, BGZF).
 * Host code: no device is needed. prg: the linearised PRG; per_base / n_per_base: the flat per-base vector of
 * gq_coverage_fetch (gq_layout.n_per_base entries); grouped / n_grouped_words: the records of gq_coverage_grouped;
 * stats = {mean coverage, coverage variance (gq_read_depth_stats), mean per-base error rate (ReadStats,
 * read_stats.cpp:21-70)}; ploidy 1 or 2; prg_coords_path: gram_dir/prg_coords.tsv (NULL or missing: one segment
 * named gramtools_prg); debug_path: NULL, or the file the per-site debug lines are appended to (--debug);
 * gcp_seed: seed of the genotype-confidence simulation (GCP::Model's default is 42, lib/GCP/GCP.h:26); n_threads: host
 * threads (level-1 sites are independent and are genotyped in parallel; 0 = all; the result does not depend on it).
 * Doubles follow the reference's arithmetic operation for operation; the VCF text is what htslib would print, restated
 * (htslib is absent here: parity unpinned, like the sdsl files). */
int gq_level_genotype(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                      const uint32_t* grouped, uint64_t n_grouped_words, const double stats[3], int ploidy,
                      const char* sample_id, const char* prg_coords_path, const char* genotype_dir,
                      const char* debug_path, uint32_t gcp_seed, int n_threads);
/* gq_read_depth_stats without a handle: the same arithmetic from the PRG and fetched coverage (host code). */
int gq_read_depth_stats_host(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                             const uint32_t* grouped, uint64_t n_grouped_words, double out[2], uint64_t counts[2]);
/* The same run, returning only the text of genotyped.json (one segment): json_out may be NULL to ask for the size;
 * *json_bytes is the capacity on entry and the size (with the terminating 0) on return. */
int gq_level_genotype_json(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                           const uint32_t* grouped, uint64_t n_grouped_words, const double stats[3], int ploidy,
                           const char* sample_id, uint32_t gcp_seed, int n_threads, char* json_out,
                           uint64_t* json_bytes);

/* Run the kernels on a caller-owned CUDA stream (e.g. torch's current stream) so that the caller's
 * CUDA events bracket them. NULL = the library's own stream. */
int gq_set_stream(gq_index* idx, void* cuda_stream);
/* Tunables (name, value) — results never depend on them:
 *   general kernel: arena_words (per-lane stack, >= 32), threads (lanes, multiple of 256), super_in_smem,
 *     rf_thresh, ev_thresh, leave, wait_max; big_arena_words / big_threads (overflow re-runs);
 *   seed_pass (0: every strand through the general kernel), seed_recs_per_read (candidate pool);
 *   pool_words_per_read (final-state pool; grows on demand), gtab_cap (initial multi-allele group table; grows);
 *   chunk_reads, tail_chunk_reads (slices of the pipelined host path), resident_slices (<= 64), overlap_classify,
 *   early_classify (pipelined path: k-mer filter of the early slices beside the late ones).
 * Parity precondition of the seeded selection (coverage_common.cpp:97-107): the reference's RandomInclusiveInt
 * is std::uniform_int_distribution over std::mt19937 as libstdc++ >= 11 implements it (Lemire's multiply-shift
 * with rejection); a reference built against an older libstdc++ or libc++ picks other classes for
 * multi-mapping reads. */
int gq_set_option(gq_index* idx, const char* name, int64_t value);
/* Counters of the last gq_map_* call: [0] kernel launches, [1] overflow re-run strands, [2] search-phase ms
 * and [3] classify + coverage ms (CUDA events; single-slice runs only), [4] pool words used, [5] H2D bytes,
 * [6] all kernels ms (CUDA events on the caller's stream), [7] host ms spent enqueueing the call */
int gq_last_run_info(gq_index* idx, double info[8]);

/* Per-kernel durations (ms, CUDA events on the launching streams) of the last single-slice gq_map_resident call:
 * [0] seed_kernel [1] verify_kernel [2] text_kernel [3] search_kernel (general) [4] classify_kernel (runs on a
 * second stream beside coverage) [5] coverage_kernel [6] revcomp_kernel; zero when the call was sliced. */
int gq_last_kernel_ms(gq_index* idx, double ms[8]);

const char* gq_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GQ_H */
