// index_build.hpp — host-side construction of the flat quasimap index from a linearised PRG.
//
// Rebuilds, from the `prg` integer string alone, everything `gramtools build` leaves in gram_dir
// for the quasimap path (reference: libgramtools/src/build/build.cpp:8-71): FM-index (SA, BWT,
// C array, DNA + marker masks), coverage graph (nodes, random access, target map, parent map) and
// the all-k-mers index — but laid out as flat arrays for HBM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gq_core.cuh"

namespace gq {

struct HostIndex {
  uint32_t n = 0;  // SA size
  uint32_t k = 0;
  std::vector<uint32_t> prg;
  // FM
  std::vector<uint32_t> sa;
  std::vector<RankBlk> rank_blk;
  std::vector<uint32_t> super_cnt;
  std::vector<uint32_t> mrank_blk;
  std::vector<uint32_t> marker_hit;
  // text mode
  std::vector<TextGrp> text_grp;
  std::vector<uint32_t> text_super, tmarker_hit, isa;
  uint32_t c_base[4] = {0, 0, 0, 0};
  // sites
  uint32_t n_slots = 0;   // (max site id - 5)/2 + 1
  uint32_t n_sites = 0;   // sites actually present
  std::vector<uint32_t> site_sa, allele_iv, par, tm_odd, tm_even_off, tm_even, entry_next, site_snp;
  std::vector<uint32_t> site_start_pos;  // PRG position of the site-entry marker
  std::vector<uint32_t> n_alleles;   // per slot (0 if the slot is unused)
  std::vector<uint32_t> allele_off;  // n_slots + 1: prefix sum of n_alleles (allele_sum layout)
  bool is_nested = false;
  // graph
  std::vector<uint32_t> pos2node;
  std::vector<Node> nodes;
  std::vector<uint32_t> edges;
  std::vector<uint32_t> site_start_node;  // per slot: bubble start node id
  uint32_t n_per_base = 0;                // number of in-bubble bases (flat per-base layout)
  // coverage recording without the graph (non-nested PRGs): per slot {apos offset, per-base offset of the site's
  // first base, text position of its first allele base, text position after its end marker}; apos = text
  // position of the first symbol of every allele, n_alleles + 1 entries per site (the last = position after the
  // site-end marker), so allele a spans [apos[a], apos[a + 1] - 1)
  std::vector<uint32_t> site_rec, apos;
  // k-mer index
  std::vector<uint32_t> kmer_bits, kmer_off, kmer_paths;
  std::vector<KmerState> kmer_states;
  std::vector<uint32_t> seed_off, seed_state;  // seed-pass view (KmerSeed)
  std::vector<KmerSeed> seed_ent;

  IndexView view() const;
};

// Throws std::runtime_error on malformed PRGs (same conditions as the reference:
// linearised_prg.cpp:52-80, coverage_graph.cpp:215-221,330-338).
// `sa_builder` (optional): suffix array of the compressed text (symbols in [0, sigma), unique minimum at the end);
// libgq passes the GPU builder (sa_gpu.cu), the test emulation leaves it null and gets the host SA-IS (sais.hpp).
using SaBuilder = std::vector<uint32_t> (*)(const std::vector<int32_t>& text, int32_t sigma, void* ctx);
// `kmer_index_dir` (optional): take the k-mer index from the four sdsl files of a gram_dir (kmers, kmers_stats,
// sa_intervals, paths — kmer_index::load, load.cpp:161-173) instead of running the k-mer searches.
void build_host_index(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, HostIndex& out,
                      SaBuilder sa_builder = nullptr, void* sa_ctx = nullptr, const char* kmer_index_dir = nullptr);

// Seed-pass view (seed_off / seed_ent / seed_state) of a finished k-mer index (kmer_off / kmer_states / kmer_paths)
void build_seed_view(HostIndex& ix);

// The k-mer index as the reference's gram_dir files (dump.cpp:27-141 / load.cpp:11-173); sdsl::int_vector
// serialisation restated in kmer_index_files.cpp (parity unpinned: no SDSL here).
void kmer_index_dump(const HostIndex& ix, const std::string& dir);
void kmer_index_load(HostIndex& ix, const std::string& dir);
// The whole flat index as one file of this back-end's own format (kmer_index_files.cpp): what `gram build` leaves in
// gram_dir/gq_index so that `gram genotype` does not rebuild anything. Checksummed; load throws on any mismatch.
void host_index_save(const HostIndex& ix, const std::string& path);
void host_index_load(HostIndex& ix, const std::string& path);
void write_int_vector(const std::string& path, const std::vector<uint64_t>& values, uint32_t width, bool fixed_width);
std::vector<uint64_t> read_int_vector(const std::string& path, uint32_t fixed_width, uint32_t* width_out);

// PRG symbols -> ranks among the symbols present + 1 (sdsl's char2comp), sentinel 0 appended; returns sigma
int32_t compress_text(const uint32_t* prg, uint64_t n_symbols, std::vector<uint32_t>& present, std::vector<int32_t>& text);

// sa_gpu.cu (libgq only): prefix doubling with radix sorts on `device`; rounds_out = sort rounds it took
std::vector<uint32_t> gpu_suffix_array(const std::vector<int32_t>& text, int32_t sigma, int device, int* rounds_out = nullptr);

}  // namespace gq
