// gq_oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT PATH).
//
// A dependency-free C++17 restatement of the gramtools `quasimap` algorithm
// (reference: iqbal-lab-org/gramtools @ 89c419e, libgramtools/). It exists so the
// CUDA path in gramtools_b200/ can be checked bit-for-bit against the reference
// semantics. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, link, import or execute anything here; the
// product library (libgq.so) never does.
//
// Parity pinning: the reference itself cannot be compiled in this image (SDSL 2.1.1,
// htslib, Boost, gtest are fetched from the network by its build). This restatement is
// pinned instead by the reference's own known-answer tests, transcribed into
// oracle/test_oracle.cpp (each case cites the reference test file:line) and by the
// integration fixtures IT1..IT3 (tests/golden/). SDSL / Boost on-disk *formats* are
// "parity unpinned".
//
// Every function cites the reference file:line it follows (paths relative to
// libgramtools/). Containers mirror the reference's (std::list of SearchState with two
// std::vector paths, std::set / std::map in the coverage code, hash-map k-mer index) so
// that this file doubles as the "port" CPU baseline.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <numeric>
#include <optional>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace gqo {

// ---- include/common/data_types.hpp:11-81 -------------------------------------------
using Base = uint8_t;
using Sequence = std::vector<Base>;
using Marker = uint32_t;
using AlleleId = int32_t;
using VariantLocus = std::pair<Marker, AlleleId>;
using VariantSitePath = std::vector<VariantLocus>;
using SA_Index = uint32_t;
using CovCount = uint16_t;
constexpr AlleleId ALLELE_UNKNOWN = -1;
constexpr AlleleId FIRST_ALLELE = 0;

inline bool is_site_marker(Marker m) {  // data_types.hpp:58-63
  if (!(m > 4)) throw std::invalid_argument("The given marker is not a variant marker (>4)");
  return m % 2 == 1;
}
inline bool is_allele_marker(Marker m) { return !is_site_marker(m); }  // :65-67
inline std::size_t siteID_to_index(Marker site) {                      // :78-81
  if (!is_site_marker(site)) throw std::invalid_argument("The given marker is not a site ID");
  return (site - 5) / 2;
}

// ---- include/genotype/quasimap/search/types.hpp:31-55 -------------------------------
struct SearchState {
  SA_Index lo = 0, hi = 0;  // inclusive SA interval
  VariantSitePath traversed;
  VariantSitePath traversing;
  bool has_path() const { return !traversed.empty() || !traversing.empty(); }
  bool operator==(const SearchState& o) const {
    return lo == o.lo && hi == o.hi && traversed == o.traversed && traversing == o.traversing;
  }
};
using SearchStates = std::list<SearchState>;

// ---- PRG string: src/prg/linearised_prg.cpp ----------------------------------------
std::vector<Marker> bracketed_to_ints(const std::string& s);   // prg_string_to_ints :166-215
std::vector<Marker> numbered_to_ints(const std::string& s);    // encode_prg :241-265
std::string ints_to_bracketed(const std::vector<Marker>& v);   // ints_to_prg_string :132-164
Sequence encode_read(const std::string& s);                    // utils.cpp:72-81 (non-ACGT empties)

// ---- coverage graph: include/prg/coverage_graph.hpp:40-235 -------------------------
struct Node {
  std::string seq;
  Marker site = 0;
  AlleleId allele = ALLELE_UNKNOWN;
  std::size_t pos = 0;
  std::vector<CovCount> cov;  // per-base counters, only allocated inside bubbles
  bool boundary = false;
  std::vector<int> next;      // successor node ids
  int64_t prg_start = -1;     // PRG index of seq[0] (not in the reference; canonical addressing)
  bool has_sequence() const { return !seq.empty(); }
  bool in_bubble() const { return allele != ALLELE_UNKNOWN && site != 0; }
  bool is_bubble_end() const { return next.size() == 1 && seq.empty(); }
};
struct NodeAccess {  // coverage_graph.hpp:126-141
  int node = -1;
  std::size_t offset = 0;
  VariantLocus target{0, ALLELE_UNKNOWN};
};
struct TargetedMarker {  // coverage_graph.hpp:143-158
  Marker id = 0;
  AlleleId direct_deletion_allele = ALLELE_UNKNOWN;
  bool operator==(const TargetedMarker& o) const {
    return id == o.id && direct_deletion_allele == o.direct_deletion_allele;
  }
};
struct CovGraph {
  std::vector<Node> nodes;
  int root = -1;
  std::vector<std::pair<int, int>> bubbles;  // (start,end) in creation order
  std::unordered_map<Marker, int> bubble_starts, bubble_ends;
  std::unordered_map<Marker, VariantLocus> par_map;
  std::vector<NodeAccess> random_access;
  std::unordered_map<Marker, std::vector<TargetedMarker>> target_map;
  bool is_nested = false;
};

// ---- FM index (stands in for sdsl::csa_wt<wt_int,1,..>, data_types.hpp:33-37) ------
struct FMIndex {
  std::vector<uint32_t> sa;        // size n = |prg|+1 (sentinel 0 appended by sdsl::construct)
  std::vector<uint32_t> bwt;       // bwt[i] = text[sa[i]-1], 0 where sa[i]==0
  std::vector<uint32_t> alphabet;  // sorted distinct symbols, alphabet[0]==0
  std::vector<uint64_t> C;         // C[comp] = #symbols smaller than alphabet[comp]; C[sigma]=n
  std::unordered_map<uint32_t, uint32_t> char2comp;
  uint64_t size() const { return sa.size(); }
  uint32_t sigma() const { return (uint32_t)alphabet.size(); }
};

// Bit mask + rank directory (stands in for sdsl::bit_vector + rank_support_v<1>).
struct RankedMask {
  std::vector<uint64_t> words;
  std::vector<uint32_t> cum;  // ones before each word
  uint64_t nbits = 0;
  void build(const std::vector<uint8_t>& bits);
  uint64_t rank(uint64_t i) const {  // # ones in [0,i)
    uint64_t w = i >> 6, r = i & 63;
    uint64_t c = cum[w];
    if (r) c += __builtin_popcountll(words[w] & ((1ULL << r) - 1));
    return c;
  }
  bool get(uint64_t i) const { return (words[i >> 6] >> (i & 63)) & 1; }
};

// ---- include/prg/prg_info.hpp:22-59 --------------------------------------------------
struct PRGInfo {
  std::vector<Marker> prg;
  std::unordered_map<Marker, int> last_allele_positions;  // linearised_prg.cpp:52-80
  FMIndex fm;
  RankedMask mask[4];  // make_data_structures.cpp:78-95 (a,c,g,t)
  RankedMask markers;  // make_data_structures.cpp:158-163
  CovGraph graph;
  uint64_t num_sites = 0;
};

struct KmerHash {
  std::size_t operator()(const Sequence& s) const {
    uint64_t h = 1469598103934665603ULL;
    for (auto b : s) h = (h ^ b) * 1099511628211ULL;
    return (std::size_t)h;
  }
};
using KmerIndex = std::unordered_map<Sequence, SearchStates, KmerHash>;

// ---- coverage/types.hpp:39-43 --------------------------------------------------------
using AlleleIds = std::vector<AlleleId>;
using GroupedAlleleCounts = std::map<AlleleIds, CovCount>;  // ordered stand-in for unordered_map
struct Coverage {
  std::vector<std::vector<CovCount>> allele_sum;
  std::vector<GroupedAlleleCounts> grouped;
};
struct Stats {  // quasimap.hpp:17-24
  uint64_t all_reads = 0, skipped = 0, missing_kmer = 0, no_extension = 0, exact_mapped = 0;
};

// Event counters that define "algorithmic bytes" (SURVEY.md §8d).
struct Events {
  uint64_t q_rank = 0, w_marker = 0, q_sa = 0, q_node = 0, a_cov = 0, strands = 0, bases = 0;
};

// ---- builders -------------------------------------------------------------------------
void build_end_positions(PRGInfo& info);                 // linearised_prg.cpp:52-80
void build_fm_index(const std::vector<Marker>& prg, FMIndex& fm);
void build_masks(PRGInfo& info);                         // make_data_structures.cpp:78-95,158-163
void build_cov_graph(PRGInfo& info);                     // coverage_graph.cpp:82-379
PRGInfo build_prg_info(const std::vector<Marker>& prg);  // submods/submod_resources.cpp:21-62

// ---- search ---------------------------------------------------------------------------
uint64_t dna_bwt_rank(const PRGInfo&, uint64_t upper, Marker base);                 // BWT_search.cpp:8-22
std::pair<SA_Index, SA_Index> marker_sa_interval(const PRGInfo&, Marker m);        // vBWT_jump.cpp:3-21
std::vector<VariantLocus> left_markers_search(const PRGInfo&, const SearchState&,
                                              Events* ev = nullptr);               // vBWT_jump.cpp:94-117
SearchStates search_state_vbwt_jumps(const PRGInfo&, const SearchState&,
                                     Events* ev = nullptr);                         // vBWT_jump.cpp:134-183
void process_markers_search_states(const PRGInfo&, SearchStates&, Events* ev = nullptr);  // :119-132
SearchStates search_base_backwards(const PRGInfo&, Base b, const SearchStates&,
                                   Events* ev = nullptr);                           // BWT_search.cpp:78-94
SearchStates process_read_char(const PRGInfo&, Base b, SearchStates&, Events* ev = nullptr);  // quasimap.cpp:258-268
SearchStates encapsulated_states(const PRGInfo&, const SearchStates&, Events* ev = nullptr);  // encapsulated_search.cpp:90-107
SearchStates search_read_backwards(const PRGInfo&, const KmerIndex&, const Sequence& read,
                                   uint32_t k, Events* ev = nullptr);               // quasimap.cpp:227-256
Sequence reverse_complement(const Sequence&);                                       // quasimap.cpp:288-298
bool all_kmers_in_index(const KmerIndex&, const Sequence& read, uint32_t k);        // quasimap.cpp:212-225

// ---- kmer index -------------------------------------------------------------------------
std::vector<Sequence> all_kmers_ordered(uint32_t k);                 // kmers.cpp:76-96
std::vector<Sequence> prefix_diffs(const std::vector<Sequence>&);    // kmers.cpp:38-74
KmerIndex index_kmers(const PRGInfo&, const std::vector<Sequence>& diffs, uint32_t k);  // build.cpp:101-131
KmerIndex build_kmer_index(const PRGInfo&, uint32_t k);              // build.cpp:138-148

// ---- coverage ---------------------------------------------------------------------------
struct LocusFinder {  // coverage_common.cpp:10-83
  std::set<Marker> base_sites, used_sites;
  std::set<VariantLocus> unique_loci;
  LocusFinder() = default;
  LocusFinder(const PRGInfo&, const SearchState&, Events* ev = nullptr);
  void assign_nested(const PRGInfo&, VariantLocus);
};
struct Selected {
  SearchStates states;
  std::set<VariantLocus> loci;
};
using UniqueSitePaths = std::map<std::set<Marker>, std::pair<SearchStates, std::set<VariantLocus>>>;
UniqueSitePaths equivalence_classes(const PRGInfo&, const SearchStates&, Events* ev = nullptr);
uint32_t count_nonvar(const SearchStates&);  // coverage_common.cpp:137-148
// mock_rand: if set, used in place of RandomInclusiveInt::generate (test mocks, mocks.hpp:8-13)
Selected select_mapping(const PRGInfo&, const SearchStates&, uint32_t seed,
                        std::optional<uint32_t> mock_rand = std::nullopt,
                        Events* ev = nullptr);  // coverage_common.cpp:85-177
uint32_t rng_generate(std::mt19937& g, uint32_t lo, uint32_t hi);  // random.cpp:16-19

struct NodeSpan {
  int node;
  uint32_t start, end;
};
// Traverser (allele_base.cpp:137-219): successive (node,[start,end]); node == -1 ends.
struct Traverser {
  const PRGInfo* info = nullptr;
  int cur = -1;
  std::size_t bases_remaining = 0;
  VariantSitePath traversed;
  uint32_t traversed_index = 0;
  bool first_node = true;
  uint32_t start_pos = 0, end_pos = 0;
  Traverser() = default;
  Traverser(const PRGInfo&, const NodeAccess& start, const VariantSitePath& traversed, std::size_t read_size);
  std::optional<int> next_node();
  void process_first_node();
  void go_to_next_site();
  void update_coordinates();
  void assign_end_position();
  void choose_allele();
};
using CovMapping = std::map<int, std::pair<std::pair<uint32_t, uint32_t>, bool>>;  // node -> ((start,end), full)
CovMapping pb_cov_mapping(const PRGInfo&, const SearchStates&, std::size_t read_size,
                          Events* ev = nullptr);  // PbCovRecorder :221-296 minus the write
void record_allele_base(PRGInfo&, const SearchStates&, std::size_t read_size, bool atomic, Events* ev = nullptr);
void record_allele_sum(Coverage&, const std::set<VariantLocus>&, bool atomic);  // allele_sum.cpp:31-43
void record_grouped(Coverage&, const std::set<VariantLocus>&);                  // grouped_allele_counts.cpp:17-49
Coverage empty_coverage(const PRGInfo&);                                        // coverage_common.cpp:206-212
std::vector<std::vector<std::vector<CovCount>>> allele_base_non_nested(const PRGInfo&);  // allele_base.cpp:10-38

enum StrandStatus : int { SKIPPED = 0, MISSING_KMER = 1, NO_EXTENSION = 2, MAPPED = 3 };

struct Mapper {
  PRGInfo info;
  KmerIndex kmers;
  uint32_t k = 0;
  Coverage cov;
  Stats stats;
  Events events;
  bool count_events = false;
  Mapper() = default;
  Mapper(const std::vector<Marker>& prg, uint32_t k);
  // quasimap_read (quasimap.cpp:159-194); final states optionally returned
  StrandStatus quasimap_read(const Sequence& read, uint32_t seed, bool atomic, SearchStates* out_states = nullptr);
  // handle_reads_buffer body for one read (quasimap.cpp:103-115): both strands, same seed
  void quasimap_forward_reverse(const Sequence& read, uint32_t seed, bool atomic,
                                StrandStatus* st = nullptr, SearchStates* fwd = nullptr, SearchStates* rev = nullptr);
  // canonical flat per-base coverage: every PRG position holding a base of an in-bubble node, PRG order
  std::vector<CovCount> per_base_flat() const;
};

// canonical serialisation of a state list: states sorted, each
// [lo, hi, n_traversed, n_traversing, (site, allele)...]
std::vector<uint32_t> canonical_states(const SearchStates&);

}  // namespace gqo
