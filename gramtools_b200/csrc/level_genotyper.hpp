// level_genotyper.hpp — the genotyping step that follows quasimap in `gram genotype`: host C++, no GPU.
//
// What the reference does after quasimap_reads returns (libgramtools/src/genotype/genotype.cpp:68-118): a
// LevelGenotyper walks the sites of the PRG from the most nested outwards, extracts the candidate alleles of each
// site with their per-base coverage (pasting the calls of nested sites), scores haploid / diploid genotypes from the
// grouped allele counts and calls the likeliest; then writes genotyped.json, the personalised reference and the VCF.
// Reference files this follows (semantics, not structure):
//   infer/allele_extracter.cpp:10-124            candidate alleles of a site
//   infer/level_genotyping/model.cpp:17-480      coverages, likelihoods, the call
//   infer/level_genotyping/probabilities.cpp     Poisson / negative binomial log pmfs
//   infer/level_genotyping/runner.cpp:27-337     site order, invalidation, AMBIG propagation, likelihood parameters,
//                                                genotype-confidence percentiles (lib/GCP/GCP.h)
//   infer/output_specs/{make_json,json_*_spec}.cpp, fields.hpp     genotyped.json
//   infer/personalised_reference.cpp:8-151, output_specs/segment_tracker.hpp, output_specs/make_vcf.cpp
//
// The reference walks a pointer graph (coverage_Graph) whose nodes own their coverage vectors; here the sites are
// read straight off the linearised PRG (entry / allele / end positions, reference coordinates, parent locus) and the
// per-base coverage is the flat vector gq_coverage_fetch returns: the count of base p of the PRG lives at
// (number of in-site bases before p) — the order index_build.cpp lays the per-base counters out in.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gq {
namespace lg {

using Cov = uint16_t;  // CovCount (common/data_types.hpp:52): every coverage count of the reference is 16 bits wide

// infer/types.hpp:15-70
struct Allele {
  std::string seq;
  std::vector<Cov> pb;   // per-base coverage, one entry per base of seq
  int32_t hapg = 0;      // which outgoing edge of the site this allele starts with
  bool callable = true;  // false: not compatible with the calls made in nested sites
  Allele() = default;
  Allele(std::string s, std::vector<Cov> c, int32_t h = 0, bool ok = true)
      : seq(std::move(s)), pb(std::move(c)), hapg(h), callable(ok) {}
  Allele joined(const Allele& right) const;  // operator+: the left haplogroup is kept, callable is and-ed
  bool operator==(const Allele& o) const { return seq == o.seq && pb == o.pb && hapg == o.hapg; }
  bool operator<(const Allele& o) const { return seq < o.seq; }
  double mean_cov() const;
};
using Alleles = std::vector<Allele>;
using Genotype = std::vector<int32_t>;  // indices into an allele vector; {-1} = null call

// one site's grouped allele counts: (sorted allele ids, count). The reference keeps a hash map per site
// (coverage/types.hpp:22); only sums over it are taken, so the order is irrelevant.
using GroupCounts = std::vector<std::pair<std::vector<int32_t>, Cov>>;

enum class Ploidy { Haploid = 1, Diploid = 2 };

// probabilities.hpp: log pmf of the coverage on a true allele, memoised
class LogPmf {
 public:
  virtual ~LogPmf() = default;
  double operator()(double cov);
  size_t n_memoised() const { return memo_.size(); }
  virtual bool is_poisson() const = 0;

 protected:
  virtual double compute(double cov) const = 0;
  std::map<double, double> memo_;
};
class PoissonLogPmf : public LogPmf {
 public:
  explicit PoissonLogPmf(double lambda) : lambda_(lambda) { (*this)(0); }
  bool is_poisson() const override { return true; }

 private:
  double compute(double cov) const override;
  double lambda_;
};
class NegBinomLogPmf : public LogPmf {  // cov failures before k successes of probability p
 public:
  NegBinomLogPmf(double k, double p) : k_(k), p_(p) { (*this)(0); }
  bool is_poisson() const override { return false; }

 private:
  double compute(double cov) const override;
  double k_, p_;
};

// likelihood_related_stats (probabilities.hpp:56-75), made by make_l_stats (runner.cpp:196-233)
struct LStats {
  double mean_cov = -1, mean_pb_error = -1, num_successes = -1, success_prob = -1;  // DataParams
  double log_mean_pb_error = 0, log_zero = 0, log_zero_half_depth = 0, log_no_zero = 0, log_no_zero_half_depth = 0;
  Cov credible_cov_t = 0;
  std::shared_ptr<LogPmf> pmf_full_depth, pmf_half_depth;
};
LStats make_l_stats(double mean_cov, double var_cov, double mean_pb_error);
Cov find_minimum_non_error_cov(double mean_pb_error, LogPmf& pmf);

// GenotypedSite + LevelGenotypedSite (interfaces.hpp:45-127, level_genotyping/site.hpp)
struct Site {
  Alleles alleles;
  Genotype genotype;
  std::vector<double> allele_covs;
  uint64_t total_coverage = 0;
  std::vector<int32_t> haplogroups;
  std::vector<std::string> filters;
  uint64_t pos = 0;       // reference coordinate of the site (0-based, first-allele path)
  uint32_t end_text = 0;  // PRG position of the site-end marker (stands for site_end_node)
  uint64_t end_pos = 0;   // reference coordinate just after the site
  size_t num_haplogroups = 0;
  std::optional<Alleles> extra_alleles;
  std::string debug_info;
  double gt_conf = 0, gt_conf_percentile = 0;

  bool is_null() const { return !genotype.empty() && genotype[0] == -1; }
  void make_null();
  bool has_filter(const std::string& name) const;
  void set_filter(const std::string& name) { filters.push_back(name); }
  Alleles unique_genotyped_alleles(const Alleles& all, const Genotype& gt) const;
  Alleles unique_genotyped_alleles() const { return unique_genotyped_alleles(alleles, genotype); }
  std::vector<int32_t> non_genotyped_haplogroups() const;
};

class IncorrectGenotyping : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

using Likelihoods = std::multimap<double, Genotype, std::greater<double>>;  // best first; ties in insertion order

// LevelGenotyperModel (model.hpp / model.cpp): one site's genotype from its candidate alleles and group counts
class SiteModel {
 public:
  SiteModel() = default;
  SiteModel(const Alleles& input_alleles, const GroupCounts& counts, Ploidy ploidy, const LStats* l_stats,
            bool debug = false);
  // the reference's constructor for tests (model.cpp:467-480): coverages and likelihoods given, nothing computed
  SiteModel(const LStats& l_stats, const std::vector<Cov>& covs, const Likelihoods& likelihoods);

  static uint64_t count_total_coverage(const GroupCounts& counts);
  static std::vector<bool> haplogroup_multiplicities(const Alleles& alleles);
  void set_haploid_coverages(const GroupCounts& counts, size_t num_haplogroups);
  void assign_coverage_to_empty_alleles(Alleles& alleles) const;
  double fraction_noncredible_positions(const Allele& a) const;
  std::pair<double, double> diploid_coverage(const GroupCounts& counts, std::vector<int32_t> hapgs,
                                             const std::vector<bool>& mults);
  static std::vector<Genotype> combinations(const Genotype& indices, size_t subset_size);  // get_permutations
  static Genotype rescale_genotypes(const Genotype& gt);
  static Likelihoods::const_iterator choose_max_likelihood(const Likelihoods& l, const Alleles& alleles);
  void call_genotype(const Alleles& input_alleles, const std::vector<bool>& mults, Ploidy ploidy);

  const std::vector<Cov>& haploid_covs() const { return haploid_; }
  const std::vector<Cov>& singleton_covs() const { return singleton_; }
  const Likelihoods& likelihoods() const { return likelihoods_; }
  Site& site() { return site_; }

 private:
  bool ignore_ref() const { return !alleles_.at(0).callable; }
  void add_likelihood(const Allele* const* chosen, double incompatible_coverage, const Genotype& indices);
  void haploid_likelihoods(const Alleles& used);
  void homozygous_likelihoods(const Alleles& used, const std::vector<bool>& mults);
  void heterozygous_likelihoods(const Alleles& used, const std::vector<bool>& mults);

  Alleles alleles_;
  GroupCounts counts_;
  Ploidy ploidy_ = Ploidy::Haploid;
  const LStats* l_stats_ = nullptr;
  bool debug_ = false;
  std::vector<Cov> haploid_, singleton_;  // coverage compatible with / unique to each haplogroup
  std::map<std::vector<int32_t>, std::pair<double, double>> diploid_memo_;
  uint64_t total_coverage_ = 0;
  Likelihoods likelihoods_;
  Site site_;
};

// The sites of a linearised PRG: text positions, reference coordinates, nesting (what LevelGenotyper reads from
// coverage_Graph: bubble_map, par_map, node positions — coverage_graph.cpp:82-266)
struct SiteText {
  uint32_t entry = 0, end = 0;           // PRG positions of the site-entry (odd) marker and of the site-end marker
  uint64_t pos = 0, end_pos = 0;         // reference coordinates of the site and of what follows it
  uint32_t pb_entry = 0, pb_exit = 0;    // in-site bases of the PRG before `entry` / up to and including `end`
  int32_t parent = -1, parent_hapg = -1; // index of the enclosing site and the allele of it this site sits on
  uint32_t n_alleles = 0;
};
struct PrgSites {
  std::vector<uint32_t> prg;
  std::vector<SiteText> sites;  // by site index (marker - 5) / 2
  bool is_nested = false;
  uint64_t ref_length = 0;      // reference coordinate after the last symbol
  // child_map (make_data_structures.cpp:53-69): parent site index -> haplogroup -> child site indices. The
  // reference fills it while iterating an unordered_map; only membership is ever asked of it.
  std::map<uint32_t, std::map<int32_t, std::vector<uint32_t>>> children;
};
PrgSites parse_prg_sites(const uint32_t* prg, uint64_t n_symbols);

// allele_extracter.cpp: candidate alleles of site s given the sites genotyped so far (all sites nested in s)
Alleles extract_alleles(const PrgSites& ps, uint32_t s, const Cov* per_base, const std::vector<Site>& genotyped);
Allele extract_ref_allele(const PrgSites& ps, uint32_t from, uint32_t own_site, const Cov* per_base);
// allele_combine (:26-60): every allele so far joined with every allele a nested site was called with
Alleles combine_with_site(const Alleles& existing, const Site& referent);

// flat records [site_index, count, n, allele ids…] (gq_coverage_grouped) -> one GroupCounts per site
std::vector<GroupCounts> unpack_grouped(const uint32_t* grouped, uint64_t n_words, size_t n_sites);

// ReadStats::compute_coverage_depth (read_stats.cpp:119-160)
struct DepthStats {
  double mean = 0, variance = 0;
  uint64_t num_sites_no_cov = 0, num_sites_total = 0;
};
DepthStats read_depth_stats(const PrgSites& ps, const Cov* per_base, const std::vector<GroupCounts>& counts);
std::pair<int32_t, Cov> max_cov_haplogroup(const GroupCounts& counts);  // read_stats.cpp:72-93
std::pair<Allele, Cov> extract_max_coverage_allele(const PrgSites& ps, uint32_t site, const Cov* per_base,
                                                   const std::vector<GroupCounts>& counts);  // read_stats.cpp:95-117

// segment_tracker.hpp
class SegmentTracker {
 public:
  struct Segment {
    std::string id;
    uint64_t size;
  };
  SegmentTracker() : SegmentTracker(std::string()) {}
  explicit SegmentTracker(const std::string& coords_text);  // contents of prg_coords.tsv: "ID<ws>size" per line
  const std::string& get_id(uint64_t pos);
  uint64_t relative_pos(uint64_t pos) const;
  uint64_t edge() const { return max_; }
  uint64_t global_edge() const { return global_max_ - 1; }
  void reset();
  const std::vector<Segment>& segments() const { return segments_; }

 private:
  std::vector<Segment> segments_;
  uint64_t min_ = 0, max_ = 0, global_max_ = 0;
  size_t cur_ = 0;
};

struct Fasta {
  std::string id, desc, seq;
  std::string to_string() const;  // personalised_reference.cpp:118-137: 60 columns, no trailing newline
};

struct RunOptions {
  Ploidy ploidy = Ploidy::Haploid;
  bool with_percentiles = true;  // get_gcp
  bool debug = false;
  uint32_t gcp_seed = 42;        // GCP::Model's default seed (GCP.h:26)
  int n_threads = 1;             // level-1 sites are independent: genotyped in parallel (0 = all the threads OpenMP gives)
};

// LevelGenotyper (runner.cpp:27-107)
class LevelGenotyper {
 public:
  // grouped: flat records [site_index, count, n, allele ids…] as gq_coverage_grouped returns them
  LevelGenotyper(PrgSites ps, const Cov* per_base, const uint32_t* grouped, uint64_t n_grouped_words, double mean_cov,
                 double var_cov, double mean_pb_error, const RunOptions& opt);
  // for tests of the invalidation / propagation logic on prepared sites (runner.hpp:38-39)
  LevelGenotyper(PrgSites ps, std::vector<Site> sites) : ps_(std::move(ps)), sites_(std::move(sites)) {}

  const std::vector<Site>& sites() const { return sites_; }
  std::vector<Site>& sites() { return sites_; }
  const PrgSites& prg_sites() const { return ps_; }
  const LStats& l_stats() const { return l_stats_; }
  const std::string& debug_text() const { return debug_text_; }

  std::vector<int32_t> haplogroups_with_sites(uint32_t site, const std::vector<int32_t>& candidates) const;
  void invalidate_if_needed(uint32_t parent, const std::vector<int32_t>& haplogroups);
  void run_invalidation(uint32_t site);
  void uppropagate_filter(const std::string& name, uint32_t parent);
  void downpropagate_filter(const std::string& name, uint32_t parent);
  static std::vector<double> gtconf_distribution(const std::vector<Site>& sites, const LStats& l_stats, Ploidy ploidy,
                                                 uint32_t seed);

  std::string json(const std::string& sample_id, SegmentTracker& tracker) const;       // genotyped.json (one line)
  std::vector<Fasta> personalised_reference(SegmentTracker& tracker) const;            // before deduplication
  std::string vcf(const std::string& sample_id, SegmentTracker& tracker) const;        // text of genotyped.vcf

 private:
  PrgSites ps_;
  std::vector<Site> sites_;
  LStats l_stats_;
  Ploidy ploidy_ = Ploidy::Haploid;
  std::string debug_text_;
};

// GCP::Percentiler (lib/GCP/GCP.h:104-183)
class Percentiler {
 public:
  explicit Percentiler(const std::vector<double>& sorted_confidences);
  double percentile(double conf) const;

 private:
  std::map<double, double> entries_;
};

std::string json_number(double v);                // a double the way nlohmann::json 3.7 dumps it
std::string bgzf_compress(const std::string& s);  // BGZF blocks + EOF block (what bcf_open(…, "wz") writes)
// write_deduped_p_refs (genotype.cpp:15-21): distinct sequences, in sequence order, one std::endl after each record
std::string deduped_fasta_text(std::vector<Fasta> refs, const std::string& desc);

}  // namespace lg
}  // namespace gq
