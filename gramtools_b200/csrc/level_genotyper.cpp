// level_genotyper.cpp — see level_genotyper.hpp. Host C++ only; linked into libgq.so and into tests/genotyper.
#include "level_genotyper.hpp"

#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <unordered_map>

namespace gq {
namespace lg {

// ------------------------------------------------------------------------------------------------ alleles

Allele Allele::joined(const Allele& right) const {  // infer/types.hpp:34-43
  Allele r;
  r.seq = seq + right.seq;
  r.pb.reserve(pb.size() + right.pb.size());
  r.pb.insert(r.pb.end(), pb.begin(), pb.end());
  r.pb.insert(r.pb.end(), right.pb.begin(), right.pb.end());
  r.hapg = hapg;
  r.callable = callable && right.callable;
  return r;
}

double Allele::mean_cov() const {  // infer/types.hpp:54-58
  double total = 0.0;
  for (Cov c : pb) total += c;
  return total / (double)pb.size();
}

// ------------------------------------------------------------------------------------------------ probabilities

double LogPmf::operator()(double cov) {  // probabilities.cpp:8-16
  auto it = memo_.find(cov);
  if (it != memo_.end()) return it->second;
  const double v = compute(cov);
  memo_.emplace(cov, v);
  return v;
}

double PoissonLogPmf::compute(double cov) const {  // probabilities.cpp:18-22
  return -1 * lambda_ + cov * std::log(lambda_) - std::lgamma(cov + 1);
}

double NegBinomLogPmf::compute(double cov) const {  // probabilities.cpp:34-39
  return std::lgamma(k_ + cov) - std::lgamma(cov + 1) - std::lgamma(k_) + k_ * std::log(p_) +
         cov * std::log(1 - p_);
}

Cov find_minimum_non_error_cov(double mean_pb_error, LogPmf& pmf) {  // runner.cpp:235-246
  double min_count = 1;
  if (std::isinf(pmf(min_count))) return (Cov)min_count;
  while (pmf(min_count) <= min_count * std::log(mean_pb_error)) ++min_count;
  return (Cov)min_count;
}

LStats make_l_stats(double mean_cov, double var_cov, double mean_pb_error) {  // runner.cpp:196-233
  LStats s;
  s.mean_cov = mean_cov;
  s.mean_pb_error = mean_pb_error;
  if (var_cov > mean_cov) {
    double k = std::pow(mean_cov, 2) / (var_cov - mean_cov);
    double p = k / (mean_cov + k);
    s.pmf_full_depth = std::make_shared<NegBinomLogPmf>(k, p);
    s.log_no_zero = std::log(1 - std::pow(p, k));
    s.num_successes = k;
    s.success_prob = p;
    // (the half-depth parameters square the variance, not the mean: runner.cpp:210 — kept as it is)
    k = std::pow(var_cov, 2) / (var_cov - mean_cov / 2);
    p = k / (mean_cov / 2 + k);
    s.pmf_half_depth = std::make_shared<NegBinomLogPmf>(k, p);
    s.log_no_zero_half_depth = std::log(1 - std::pow(p, k));
  } else {
    s.pmf_full_depth = std::make_shared<PoissonLogPmf>(mean_cov);
    s.log_no_zero = std::log(1 - std::exp(mean_cov * -1));
    s.pmf_half_depth = std::make_shared<PoissonLogPmf>(mean_cov / 2);
    s.log_no_zero_half_depth = std::log(1 - std::exp(mean_cov * -0.5));
  }
  s.log_mean_pb_error = std::log(mean_pb_error);
  s.log_zero = (*s.pmf_full_depth)(0);
  s.log_zero_half_depth = (*s.pmf_half_depth)(0);
  s.credible_cov_t = find_minimum_non_error_cov(mean_pb_error, *s.pmf_full_depth);
  return s;
}

// ------------------------------------------------------------------------------------------------ sites

void Site::make_null() {  // interfaces.hpp:83-87 + site.cpp:45-48: alleles, coverages and haplogroups stay
  genotype = Genotype{-1};
  total_coverage = 0;
  gt_conf = 0.;
  gt_conf_percentile = 0.;
}

bool Site::has_filter(const std::string& name) const {
  return std::find(filters.begin(), filters.end(), name) != filters.end();
}

Alleles Site::unique_genotyped_alleles(const Alleles& all, const Genotype& gt) const {  // interfaces.cpp:13-31
  std::set<int32_t> distinct;
  if (!is_null()) distinct.insert(gt.begin(), gt.end());  // sorted: the REF allele comes out first
  Alleles r;
  r.reserve(distinct.size());
  for (int32_t g : distinct) r.push_back(all.at((size_t)g));
  return r;
}

std::vector<int32_t> Site::non_genotyped_haplogroups() const {  // site.cpp:6-22
  std::set<int32_t> called;
  if (!is_null())
    for (int32_t g : genotype) called.insert(alleles.at((size_t)g).hapg);
  std::vector<int32_t> r;
  for (int32_t h = 0; h < (int32_t)num_haplogroups; ++h)
    if (!called.count(h)) r.push_back(h);
  return r;
}

// ------------------------------------------------------------------------------------------------ the site model

uint64_t SiteModel::count_total_coverage(const GroupCounts& counts) {  // model.cpp:160-165
  uint64_t t = 0;
  for (auto& e : counts) t += e.second;
  return t;
}

std::vector<bool> SiteModel::haplogroup_multiplicities(const Alleles& alleles) {  // model.cpp:177-191
  std::map<int32_t, size_t> n;
  for (auto& a : alleles) n[a.hapg] += 1;
  std::vector<bool> m(n.size(), false);
  for (auto& e : n)
    if (e.second > 1) m.at((size_t)e.first) = true;
  return m;
}

void SiteModel::set_haploid_coverages(const GroupCounts& counts, size_t num_haplogroups) {  // model.cpp:64-79
  haploid_.assign(num_haplogroups, 0);
  singleton_.assign(num_haplogroups, 0);
  for (auto& e : counts) {
    for (int32_t id : e.first) haploid_.at((size_t)id) = (Cov)(haploid_.at((size_t)id) + e.second);  // 16-bit sums
    if (e.first.size() == 1) singleton_.at((size_t)e.first[0]) = e.second;
  }
}

void SiteModel::assign_coverage_to_empty_alleles(Alleles& alleles) const {  // model.cpp:81-89: direct deletions
  for (auto& a : alleles)
    if (a.seq.empty()) a.pb = std::vector<Cov>{haploid_.at((size_t)a.hapg)};
}

double SiteModel::fraction_noncredible_positions(const Allele& a) const {  // model.cpp:148-158
  double n = 0.;
  for (Cov c : a.pb)
    if (c < l_stats_->credible_cov_t) ++n;
  return n / (double)a.pb.size();
}

std::pair<double, double> SiteModel::diploid_coverage(const GroupCounts& counts, std::vector<int32_t> hapgs,
                                                      const std::vector<bool>& mults) {  // model.cpp:91-146
  std::sort(hapgs.begin(), hapgs.end());
  auto known = diploid_memo_.find(hapgs);
  if (known != diploid_memo_.end()) return known->second;
  std::pair<double, double> r;
  if (hapgs.at(0) == hapgs.at(1)) {
    const double c = (double)haploid_.at((size_t)hapgs[0]) / 2;
    r = {c, c};
  } else {
    double c1 = (double)haploid_.at((size_t)hapgs[0]), c2 = (double)haploid_.at((size_t)hapgs[1]);
    Cov shared = 0;  // CovCount in the reference: the sum is 16 bits wide
    for (auto& e : counts) {
      const bool has1 = std::find(e.first.begin(), e.first.end(), hapgs[0]) != e.first.end();
      const bool has2 = std::find(e.first.begin(), e.first.end(), hapgs[1]) != e.first.end();
      if (has1 && has2) shared = (Cov)(shared + e.second);
    }
    const double own1 = c1 - shared, own2 = c2 - shared;
    const double belonging = (own1 == 0 && own2 == 0) ? 0.5 : own1 / (own1 + own2);
    c1 -= (1 - belonging) * shared;
    c2 -= belonging * shared;
    if (mults.at((size_t)hapgs[0])) c1 /= 2;
    if (mults.at((size_t)hapgs[1])) c2 /= 2;
    r = {c1, c2};
  }
  diploid_memo_.emplace(hapgs, r);
  return r;
}

std::vector<Genotype> SiteModel::combinations(const Genotype& indices, size_t subset_size) {  // model.cpp:214-236
  const size_t n = indices.size();
  if (subset_size > n) return {};
  std::vector<bool> pick(n, false);
  std::fill(pick.begin(), pick.begin() + (long)subset_size, true);
  std::vector<Genotype> out;
  do {
    Genotype combo;
    for (size_t i = 0; i < n; ++i)
      if (pick[i]) combo.push_back(indices[i]);
    std::sort(combo.begin(), combo.end());
    out.push_back(combo);
  } while (std::prev_permutation(pick.begin(), pick.end()));
  return out;
}

Genotype SiteModel::rescale_genotypes(const Genotype& gt) {  // model.cpp:193-212: first-seen order, 0 stays 0
  std::unordered_map<int32_t, int32_t> to{{0, 0}};
  Genotype r;
  int32_t next = 1;
  for (int32_t g : gt) {
    if (!to.count(g)) to.emplace(g, next++);
    r.push_back(to.at(g));
  }
  return r;
}

void SiteModel::add_likelihood(const Allele* const* chosen, double incompatible_coverage, const Genotype& indices) {
  // model.cpp:238-268: errors pay log(mean error) each; each chosen allele is scored on its MEAN per-base coverage
  // under the full-depth pmf (also in diploid calls), plus the zero-coverage log probability per non-credible base
  double ll = incompatible_coverage * l_stats_->log_mean_pb_error;
  const size_t n = ploidy_ == Ploidy::Haploid ? 1 : 2;
  for (size_t i = 0; i < n; ++i) {
    const Allele& a = *chosen[i];
    ll += (*l_stats_->pmf_full_depth)(a.mean_cov());
    ll += fraction_noncredible_positions(a) * l_stats_->log_zero;
  }
  likelihoods_.insert({ll, indices});
}

void SiteModel::haploid_likelihoods(const Alleles& used) {  // model.cpp:270-282
  for (int32_t i = 0; i < (int32_t)used.size(); ++i) {
    if (i == 0 && ignore_ref()) continue;
    const Allele* a = &used[(size_t)i];
    const Cov hap = haploid_.at((size_t)a->hapg);
    const uint64_t incompatible = total_coverage_ - hap;  // size_t arithmetic in the reference
    add_likelihood(&a, (double)incompatible, Genotype{i});
  }
}

void SiteModel::homozygous_likelihoods(const Alleles& used, const std::vector<bool>& mults) {  // model.cpp:284-301
  for (int32_t i = 0; i < (int32_t)used.size(); ++i) {
    if (i == 0 && ignore_ref()) continue;
    const Allele* a = &used[(size_t)i];
    auto c = diploid_coverage(counts_, {a->hapg, a->hapg}, mults);
    const double incompatible = (double)total_coverage_ - c.first - c.second;
    const Allele* both[2] = {a, a};
    add_likelihood(both, incompatible, Genotype{i, i});
  }
}

void SiteModel::heterozygous_likelihoods(const Alleles& used, const std::vector<bool>& mults) {  // model.cpp:303-335
  Genotype selected;  // only alleles whose haplogroup has coverage of its own
  for (int32_t i = 0; i < (int32_t)used.size(); ++i) {
    if (i == 0 && ignore_ref()) continue;
    if (singleton_.at((size_t)used[(size_t)i].hapg) != 0) selected.push_back(i);
  }
  if (selected.size() < 2) return;
  for (auto& combo : combinations(selected, 2)) {
    const Allele* pair[2] = {&used.at((size_t)combo[0]), &used.at((size_t)combo[1])};
    auto c = diploid_coverage(counts_, {pair[0]->hapg, pair[1]->hapg}, mults);
    const double incompatible = (double)total_coverage_ - c.first - c.second;
    add_likelihood(pair, incompatible, combo);
  }
}

Likelihoods::const_iterator SiteModel::choose_max_likelihood(const Likelihoods& l, const Alleles& alleles) {
  // model.cpp:374-399: the best genotype all of whose alleles are callable, with a runner-up left after it
  if (l.size() < 2) throw IncorrectGenotyping("Less than 2 alleles have a likelihood.\nAllele extraction bug?");
  auto it = l.begin();
  for (; it != l.end(); ++it) {
    bool callable = true;
    for (int32_t g : it->second)
      if (!alleles.at((size_t)g).callable) {
        callable = false;
        break;
      }
    if (callable) break;
  }
  if (std::distance(it, l.end()) < 2)
    throw IncorrectGenotyping("Fewer than 2 alleles are callable.\nAllele extraction bug?");
  return it;
}

void SiteModel::call_genotype(const Alleles& input, const std::vector<bool>& mults, Ploidy ploidy) {  // model.cpp:401-465
  const Allele& ref = input.at(0);
  auto it = choose_max_likelihood(likelihoods_, input);
  const double best = it->first;
  const Genotype chosen = it->second;
  ++it;
  const double conf = best - it->first;
  const Genotype runner_up = it->second;

  if (conf == 0.) {  // a tie: null call, and every allele of both genotypes is kept for the parent site (:354-361)
    site_.alleles = Alleles{ref};
    site_.make_null();
    std::set<int32_t> all(runner_up.begin(), runner_up.end());
    all.insert(chosen.begin(), chosen.end());
    Alleles extra;
    for (int32_t g : all) extra.push_back(input.at((size_t)g));
    site_.extra_alleles = extra;
    return;
  }
  {  // add_next_best_alleles (:337-363): a weakly supported call hands its runner-up to the parent, as non-callable
    const Allele& chosen_a = input.at((size_t)chosen.at(0));
    const Allele& next_a = input.at((size_t)runner_up.at(0));
    const bool low_total = (double)total_coverage_ < l_stats_->mean_cov / 4;
    const bool low_relative =
        (int)haploid_.at((size_t)chosen_a.hapg) < (int)haploid_.at((size_t)next_a.hapg) * 2;
    if (low_total || low_relative) {
      std::set<int32_t> next(runner_up.begin(), runner_up.end());
      for (int32_t g : chosen) next.erase(g);
      Alleles extra;
      for (int32_t g : next) {
        Allele a = input.at((size_t)g);
        a.callable = false;
        extra.push_back(a);
      }
      site_.extra_alleles = extra;
    }
  }

  Alleles chosen_alleles = site_.unique_genotyped_alleles(input, chosen);
  std::vector<int32_t> chosen_hapgs;
  for (int32_t g : chosen) chosen_hapgs.push_back(input.at((size_t)g).hapg);
  std::sort(chosen_hapgs.begin(), chosen_hapgs.end());
  std::vector<double> covs;
  if (ploidy == Ploidy::Haploid)
    covs = {(double)haploid_.at((size_t)chosen_hapgs.at(0))};
  else {
    const auto& c = diploid_memo_.at(chosen_hapgs);
    if (chosen.at(0) == chosen.at(1)) covs = {c.first + c.second};  // homozygous: one allele, all the coverage
    else covs = {c.first, c.second};
  }
  Genotype rescaled = rescale_genotypes(chosen);
  if (rescaled.at(0) != 0) {  // REF not called: it is reported all the same, with the coverage unique to it
    chosen_alleles.insert(chosen_alleles.begin(), ref);
    double ref_cov = (double)singleton_.at(0);
    if (mults.at(0)) ref_cov /= 2;
    covs.insert(covs.begin(), ref_cov);
  }
  site_.alleles = chosen_alleles;
  site_.genotype = rescaled;
  site_.allele_covs = covs;
  site_.total_coverage = total_coverage_;
  site_.haplogroups.clear();
  for (int32_t g : rescaled) site_.haplogroups.push_back(chosen_alleles.at((size_t)g).hapg);
  site_.gt_conf = conf;

  if (debug_) {
    std::string d = "\tnext_best_seq: ";
    for (int32_t g : runner_up) d += input.at((size_t)g).seq + ",";
    d += "\tnext_best_cov: ";
    std::vector<int32_t> hs;
    for (int32_t g : runner_up) hs.push_back(input.at((size_t)g).hapg);
    std::sort(hs.begin(), hs.end());
    for (int32_t h : hs) d += std::to_string(haploid_.at((size_t)h)) + ",";
    site_.debug_info = d;
  }
}

SiteModel::SiteModel(const Alleles& input_alleles, const GroupCounts& counts, Ploidy ploidy, const LStats* l_stats,
                     bool debug)
    : alleles_(input_alleles), counts_(counts), ploidy_(ploidy), l_stats_(l_stats), debug_(debug) {
  // model.cpp:19-62
  if (alleles_.size() < 2) throw std::logic_error("a site needs at least two candidate alleles");
  const auto mults = haplogroup_multiplicities(alleles_);
  site_.num_haplogroups = mults.size();
  {  // two candidate alleles with the same sequence: the site is ambiguous
    std::set<std::string> seen;
    for (auto& a : alleles_)
      if (!seen.insert(a.seq).second) {
        site_.set_filter("AMBIG");
        break;
      }
  }
  total_coverage_ = count_total_coverage(counts_);
  if (total_coverage_ == 0 || l_stats_->mean_cov == 0) {
    site_.alleles = Alleles{alleles_.at(0)};
    site_.make_null();
    return;
  }
  set_haploid_coverages(counts_, mults.size());
  Alleles used(alleles_);
  assign_coverage_to_empty_alleles(used);
  if (ploidy_ == Ploidy::Haploid)
    haploid_likelihoods(used);
  else {
    homozygous_likelihoods(used, mults);
    heterozygous_likelihoods(used, mults);
  }
  call_genotype(alleles_, mults, ploidy_);
}

SiteModel::SiteModel(const LStats& l_stats, const std::vector<Cov>& covs, const Likelihoods& likelihoods)
    : l_stats_(&l_stats), haploid_(covs), singleton_(covs), likelihoods_(likelihoods) {
  for (Cov c : covs) total_coverage_ += c;
}

// ------------------------------------------------------------------------------------------------ the PRG's sites

PrgSites parse_prg_sites(const uint32_t* prg, uint64_t n_symbols) {
  PrgSites ps;
  ps.prg.assign(prg, prg + n_symbols);
  uint32_t max_marker = 4;
  for (uint32_t m : ps.prg) {
    if (m < 1) throw std::runtime_error("PRG symbols must be >= 1");
    max_marker = std::max(max_marker, m);
  }
  const uint32_t S = max_marker > 4 ? ((max_marker % 2 ? max_marker : max_marker - 1) - 5) / 2 + 1 : 0;
  ps.sites.assign(S, SiteText{});
  std::vector<uint32_t> last_pos(S, 0xFFFFFFFFu);
  std::vector<uint8_t> seen(S, 0);
  for (uint64_t p = 0; p < n_symbols; ++p) {
    const uint32_t m = ps.prg[p];
    if (m <= 4) continue;
    if (m & 1u) {
      if (seen[(m - 5) / 2]) throw std::runtime_error("PRG consistency error: site marker used for two sites");
      seen[(m - 5) / 2] = 1;
    } else
      last_pos[(m - 6) / 2] = (uint32_t)p;
  }
  for (uint32_t s = 0; s < S; ++s)
    if (!seen[s] || last_pos[s] == 0xFFFFFFFFu)
      throw std::runtime_error("site ids of the PRG are not contiguous, or a site is never closed");

  struct Open {
    uint32_t site;
    int32_t allele;
  };
  std::vector<Open> open;
  uint64_t ref = 0;  // reference coordinate: bases along the path of first alleles (coverage_graph.cpp:157-246)
  uint32_t pb = 0;
  for (uint64_t p = 0; p < n_symbols; ++p) {
    const uint32_t m = ps.prg[p];
    if (m <= 4) {
      ++ref;
      if (!open.empty()) ++pb;
    } else if (m & 1u) {
      const uint32_t s = (m - 5) / 2;
      SiteText& st = ps.sites[s];
      st.entry = (uint32_t)p;
      st.pos = ref;
      st.pb_entry = pb;
      if (!open.empty()) {
        st.parent = (int32_t)open.back().site;
        st.parent_hapg = open.back().allele;
        ps.children[open.back().site][open.back().allele].push_back(s);
        ps.is_nested = true;
      }
      open.push_back({s, 0});
    } else {
      const uint32_t s = (m - 6) / 2;
      if (open.empty() || open.back().site != s) throw std::runtime_error("PRG consistency error: allele marker outside its site");
      SiteText& st = ps.sites[s];
      if (open.back().allele == 0) st.end_pos = ref;  // the site ends where its FIRST allele ends
      if (p < last_pos[s]) {
        ref = st.pos;
        open.back().allele++;
      } else {
        if (open.back().allele == 0) throw std::runtime_error("Site numbered " + std::to_string(m) + " has only one allele");
        st.n_alleles = (uint32_t)open.back().allele + 1;
        st.end = (uint32_t)p;
        st.pb_exit = pb;
        ref = st.end_pos;
        open.pop_back();
      }
    }
  }
  if (!open.empty()) throw std::runtime_error("PRG consistency error: unterminated site");
  ps.ref_length = ref;
  return ps;
}

static inline char base_char(uint32_t m) { return "?ACGT"[m]; }

Allele extract_ref_allele(const PrgSites& ps, uint32_t from, uint32_t own_site, const Cov* per_base) {
  // allele_extracter.cpp:78-92: from the first node of a haplogroup, always the first edge — i.e. the first allele
  // of every nested site — until the site's end node
  Allele r;
  const uint32_t own_even = 6 + 2 * own_site;
  // in-site bases before `from`: the caller passes the haplogroup start of the site's FIRST allele, entry + 1
  uint32_t pb = ps.sites[own_site].pb_entry;
  for (uint32_t p = from;; ++p) {
    const uint32_t m = ps.prg[p];
    if (m <= 4) {
      r.seq.push_back(base_char(m));
      r.pb.push_back(per_base[pb++]);
    } else if (m == own_even)
      break;
    else if (!(m & 1u)) {  // end of a nested site's first allele: skip its other alleles
      const SiteText& t = ps.sites[(m - 6) / 2];
      if (p < t.end) {
        p = t.end;
        pb = t.pb_exit;
      }
    }
  }
  return r;
}

Alleles combine_with_site(const Alleles& existing, const Site& referent) {  // allele_extracter.cpp:26-60
  Alleles relevant = referent.unique_genotyped_alleles();
  if (referent.extra_alleles) relevant.insert(relevant.end(), referent.extra_alleles->begin(), referent.extra_alleles->end());
  if (relevant.empty()) relevant.push_back(referent.alleles.at(0));
  constexpr size_t kMaxCombinations = 10000;  // extra alleles are dropped first, then genotyped ones
  while (existing.size() * relevant.size() > kMaxCombinations) relevant.resize(relevant.size() - 1);
  Alleles out;
  out.reserve(existing.size() * relevant.size());
  for (auto& a : existing)
    for (auto& b : relevant) out.push_back(a.joined(b));
  return out;
}

Alleles extract_alleles(const PrgSites& ps, uint32_t s, const Cov* per_base, const std::vector<Site>& genotyped) {
  // allele_extracter.cpp:10-24,94-124: one pass over the site's text; a nested site contributes the alleles it was
  // called with (times what is already there) and is stepped over
  const SiteText& st = ps.sites[s];
  const uint32_t own_even = 6 + 2 * s;
  Alleles all;
  uint32_t pb = st.pb_entry;
  int32_t hapg = 0;
  Alleles cur{Allele{"", {}, hapg}};
  std::string run_seq;
  std::vector<Cov> run_cov;
  auto flush_run = [&]() {  // allele_paste: sequence common to the haplogroup
    if (run_seq.empty()) return;
    Allele piece{run_seq, run_cov};
    for (auto& a : cur) a = a.joined(piece);
    run_seq.clear();
    run_cov.clear();
  };
  for (uint32_t p = st.entry + 1; p <= st.end; ++p) {
    const uint32_t m = ps.prg[p];
    if (m <= 4) {
      run_seq.push_back(base_char(m));
      run_cov.push_back(per_base[pb++]);
    } else if (m & 1u) {
      flush_run();
      const uint32_t t = (m - 5) / 2;
      cur = combine_with_site(cur, genotyped.at(t));
      p = ps.sites[t].end;
      pb = ps.sites[t].pb_exit;
    } else if (m == own_even) {
      flush_run();
      if (hapg == 0) {  // place_ref_as_first_allele (:69-76)
        Allele ref = extract_ref_allele(ps, st.entry + 1, s, per_base);
        auto found = std::find(cur.begin(), cur.end(), ref);
        if (found == cur.end()) {
          ref.callable = false;
          cur.insert(cur.begin(), ref);
        } else if (found != cur.begin())
          std::swap(*found, cur.front());
      }
      all.insert(all.end(), cur.begin(), cur.end());
      ++hapg;
      cur = Alleles{Allele{"", {}, hapg}};
    } else
      throw std::runtime_error("PRG consistency error: stray allele marker inside a site");
  }
  return all;
}

// ------------------------------------------------------------------------------------------------ the runner

std::vector<int32_t> LevelGenotyper::haplogroups_with_sites(uint32_t site, const std::vector<int32_t>& candidates) const {
  std::vector<int32_t> r;  // runner.cpp:146-157
  auto it = ps_.children.find(site);
  if (it == ps_.children.end()) return r;
  for (int32_t c : candidates)
    if (it->second.count(c)) r.push_back(c);
  return r;
}

void LevelGenotyper::invalidate_if_needed(uint32_t parent, const std::vector<int32_t>& haplogroups) {
  // runner.cpp:159-187: the sites on haplogroups that were not called are nulled, and so is all they contain
  if (haplogroups.empty()) return;
  std::vector<std::pair<uint32_t, int32_t>> todo;
  for (int32_t h : haplogroups) todo.emplace_back(parent, h);
  while (!todo.empty()) {
    auto locus = todo.back();
    todo.pop_back();
    const auto children = ps_.children.at(locus.first).at(locus.second);
    for (uint32_t child : children) {
      Site& c = sites_.at(child);
      if (c.is_null()) continue;
      c.make_null();
      std::vector<int32_t> all_h;
      for (int32_t h = 0; h < (int32_t)c.num_haplogroups; ++h) all_h.push_back(h);
      for (int32_t h : haplogroups_with_sites(child, all_h)) todo.emplace_back(child, h);
    }
  }
}

void LevelGenotyper::run_invalidation(uint32_t site) {  // runner.cpp:135-144
  if (!ps_.children.count(site)) return;
  invalidate_if_needed(site, haplogroups_with_sites(site, sites_.at(site).non_genotyped_haplogroups()));
}

void LevelGenotyper::uppropagate_filter(const std::string& name, uint32_t parent) {  // runner.cpp:99-112
  auto it = ps_.children.find(parent);
  if (it == ps_.children.end()) return;
  for (auto& by_hapg : it->second)
    for (uint32_t child : by_hapg.second)
      if (sites_.at(child).has_filter(name)) {
        sites_.at(parent).set_filter(name);
        return;
      }
}

void LevelGenotyper::downpropagate_filter(const std::string& name, uint32_t parent) {  // runner.cpp:114-133
  std::vector<uint32_t> todo{parent};
  while (!todo.empty()) {
    const uint32_t cur = todo.back();
    todo.pop_back();
    auto it = ps_.children.find(cur);
    if (it == ps_.children.end()) continue;
    for (auto& by_hapg : it->second)
      for (uint32_t child : by_hapg.second)
        if (!sites_.at(child).has_filter(name)) {
          sites_.at(child).set_filter(name);
          todo.push_back(child);
        }
  }
}

Percentiler::Percentiler(const std::vector<double>& v) {  // GCP.h:111-134
  if (v.size() < 2) throw std::runtime_error("Please provide at least two simulated genotype confidences.");
  auto pct = [&](std::vector<double>::const_iterator it) { return 100. * (double)(std::distance(v.begin(), it) + 1) / (double)v.size(); };
  auto cur = v.begin();
  while (cur != v.end()) {
    auto hi = std::upper_bound(v.begin(), v.end(), *cur);
    const double p = pct(cur);
    if (cur == hi - 1) entries_[*cur] = p;
    else entries_[*cur] = p + (pct(hi - 1) - p) / 2;  // a run of equal confidences: the middle of its range
    cur = hi;
  }
}

double Percentiler::percentile(double q) const {  // GCP.h:140-152
  auto above = entries_.upper_bound(q);
  if (above == entries_.end()) return 100.0;
  if (above->first == q) return above->second;
  if (above == entries_.begin()) return 0.0;
  auto below = std::prev(above);
  const double slope = (above->second - below->second) / (above->first - below->first);
  return below->second + slope * (q - below->first);
}

std::vector<double> LevelGenotyper::gtconf_distribution(const std::vector<Site>& sites, const LStats& ls, Ploidy ploidy,
                                                        uint32_t seed) {
  // runner.cpp:248-337: 10 000 confidences — those of the sites, topped up with calls on simulated two-allele sites
  // (true allele coverage from the fitted depth distribution, the other one from sequencing errors), or a random
  // sample of the sites' own when there are more than 10 000 (the reference seeds that draw from random_device)
  constexpr size_t kSize = 10000;
  std::vector<double> conf(kSize);
  size_t at = 0;
  if (sites.size() > kSize) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<> pick(0, (int)sites.size() - 1);
    while (at < kSize) conf[at++] = sites.at((size_t)pick(gen)).gt_conf;
  } else {
    for (auto& s : sites) conf[at++] = s.gt_conf;
    std::default_random_engine gen(seed);  // GCP::Model (GCP.h:26,46)
    std::vector<double> simulated;
    simulated.reserve(kSize - at);
    for (size_t i = at; i < kSize; ++i) {
      // the reference draws 16-bit counts (std::…_distribution<CovCount>); libstdc++'s Poisson rejection loop never
      // ends once its mean passes the type's maximum, which a gamma-mixed draw reaches when the depth is in the
      // thousands. Beyond 1000x (no sequencing run; reachable through the C ABI) the draw is made 32 bits wide and
      // clamped, so the call always returns; below, the reference's own distributions are used as they are.
      const bool wide = !(ls.mean_cov <= 1000.0);
      Cov correct;
      if (ls.pmf_full_depth->is_poisson()) {
        if (wide) {
          std::poisson_distribution<uint32_t> d(std::min(ls.mean_cov, 1e9));
          correct = (Cov)std::min<uint32_t>(d(gen), 65535u);
        } else {
          std::poisson_distribution<Cov> d(ls.mean_cov);
          correct = d(gen);
        }
      } else if (wide) {
        std::negative_binomial_distribution<uint32_t> d((uint32_t)(Cov)ls.num_successes, ls.success_prob);
        correct = (Cov)std::min<uint32_t>(d(gen), 65535u);
      } else {
        std::negative_binomial_distribution<Cov> d(ls.num_successes, ls.success_prob);
        correct = d(gen);
      }
      std::binomial_distribution<Cov> b(ls.mean_cov, ls.mean_pb_error);
      const Cov incorrect = b(gen);
      Alleles alleles{Allele{"C", {correct}, 0}, Allele{"A", {incorrect}, 1}};
      GroupCounts counts{{{0}, correct}, {{1}, incorrect}};
      SiteModel m(alleles, counts, ploidy, &ls);
      simulated.push_back(m.site().gt_conf);
    }
    std::sort(simulated.begin(), simulated.end());
    std::copy(simulated.begin(), simulated.end(), conf.begin() + (long)at);
  }
  std::sort(conf.begin(), conf.end());
  return conf;
}

LevelGenotyper::LevelGenotyper(PrgSites ps, const Cov* per_base, const uint32_t* grouped, uint64_t n_grouped_words,
                               double mean_cov, double var_cov, double mean_pb_error, const RunOptions& opt)
    : ps_(std::move(ps)), ploidy_(opt.ploidy) {
  const uint32_t S = (uint32_t)ps_.sites.size();
  sites_.assign(S, Site{});
  const std::vector<GroupCounts> counts = unpack_grouped(grouped, n_grouped_words, S);
  l_stats_ = make_l_stats(mean_cov, var_cov, mean_pb_error);
  if (opt.debug) {  // operator<< of likelihood_related_stats (probabilities.cpp:41-56)
    char buf[1024];
    std::snprintf(buf, sizeof buf,
                  "Model params: \nmean cov: %f\nmean per-base error: %f\nnum successes: %f\nprob of success: %f \n"
                  "log_prob_zero_cov: %f \nlog_prob_nonzero_cov: %f\n",
                  l_stats_.mean_cov, l_stats_.mean_pb_error, l_stats_.num_successes, l_stats_.success_prob,
                  l_stats_.log_zero, l_stats_.log_no_zero);
    debug_text_ += buf;
  }
  // most nested first: the reference iterates bubble_map, ordered so that children come before parents
  // (coverage_graph.hpp:181-187); a site's entry marker always follows its parent's, so descending entry position
  // gives such an order (sites that do not contain one another are independent)
  std::vector<uint32_t> order(S);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ps_.sites[a].entry > ps_.sites[b].entry; });
  static const Cov kNoCov = 0;
  const Cov* pb = per_base ? per_base : &kNoCov;
  // a level-1 site and everything nested in it form one unit of work: in `order` its descendants come right before
  // it, so the units are consecutive runs ending with a site that has no parent. Units are independent (allele
  // extraction, invalidation and filter propagation stay inside one), so they are genotyped in parallel; each thread
  // memoises the pmf in its own copy of the likelihood parameters. Debug output is ordered: one thread.
  std::vector<uint32_t> unit_end;
  for (uint32_t i = 0; i < S; ++i)
    if (ps_.sites[order[i]].parent < 0) unit_end.push_back(i + 1);
  int n_threads = opt.debug ? 1 : opt.n_threads;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
  n_threads = std::max(1, std::min<int>(n_threads, (int)unit_end.size() / 64 + 1));
  std::string first_error;
#pragma omp parallel num_threads(n_threads)
  {
    LStats local = n_threads > 1 ? make_l_stats(mean_cov, var_cov, mean_pb_error) : l_stats_;
#pragma omp for schedule(dynamic, 64)
    for (int64_t u = 0; u < (int64_t)unit_end.size(); ++u) {
      uint32_t at = 0;
      try {
        for (uint32_t i = u ? unit_end[(size_t)u - 1] : 0; i < unit_end[(size_t)u]; ++i) {
          const uint32_t s = at = order[i];
          Alleles alleles = extract_alleles(ps_, s, pb, sites_);
          SiteModel model(alleles, counts[s], ploidy_, &local, opt.debug);
          Site site = model.site();
          site.pos = ps_.sites[s].pos;
          site.end_text = ps_.sites[s].end;
          site.end_pos = ps_.sites[s].end_pos;
          if (opt.debug) {
            debug_text_ += "site index: \t" + std::to_string(s);
            debug_text_ += site.is_null() ? std::string("\tnull gt \n") : site.debug_info + "\n";
          }
          sites_[s] = std::move(site);
          run_invalidation(s);
          if (sites_[s].has_filter("AMBIG")) downpropagate_filter("AMBIG", s);
          else uppropagate_filter("AMBIG", s);
        }
      } catch (const std::exception& e) {
#pragma omp critical(gq_lg_error)
        if (first_error.empty()) first_error = "site " + std::to_string(at) + ": " + e.what();
      }
    }
  }
  if (!first_error.empty()) throw std::runtime_error(first_error);
  if (opt.with_percentiles && S > 0) {
    Percentiler pc(gtconf_distribution(sites_, l_stats_, ploidy_, opt.gcp_seed));
    for (auto& s : sites_) s.gt_conf_percentile = pc.percentile(s.gt_conf);
  }
}

// ------------------------------------------------------------------------------------------------ read depth

std::pair<int32_t, Cov> max_cov_haplogroup(const GroupCounts& counts) {
  // get_max_cov_haplogroup (read_stats.cpp:72-93): 16-bit sums per allele id, the first maximum in id order
  std::map<int32_t, Cov> per;
  for (auto& e : counts)
    for (int32_t id : e.first) per[id] = (Cov)(per[id] + e.second);
  std::pair<int32_t, Cov> best{0, 0};
  bool first = true;
  for (auto& kv : per)
    if (first || kv.second > best.second) best = kv, first = false;
  return best;
}

std::pair<Allele, Cov> extract_max_coverage_allele(const PrgSites& ps, uint32_t s, const Cov* per_base,
                                                   const std::vector<GroupCounts>& counts) {
  // extract_max_coverage_allele (read_stats.cpp:95-117): through site s along its most covered haplogroup, and
  // along the most covered haplogroup of every site met on the way; the count returned is that of site s
  const auto top = max_cov_haplogroup(counts.at(s));
  Allele path;
  struct Open {  // a site open on the chosen path: the allele wanted of it, the allele the scan is in
    uint32_t site;
    int32_t want, cur;
  };
  std::vector<Open> open{{s, top.first, 0}};
  if (top.first >= (int32_t)ps.sites[s].n_alleles) throw std::runtime_error("inconsistent grouped allele counts");
  uint32_t pb = ps.sites[s].pb_entry;
  for (uint32_t p = ps.sites[s].entry + 1; !open.empty(); ++p) {
    const uint32_t m = ps.prg[p];
    Open& o = open.back();
    if (m <= 4) {
      if (o.cur == o.want) {
        path.seq.push_back(base_char(m));
        path.pb.push_back(per_base[pb]);
      }
      ++pb;
    } else if (m & 1u) {
      const uint32_t t = (m - 5) / 2;
      if (o.cur == o.want) {
        const int32_t want = max_cov_haplogroup(counts.at(t)).first;
        if (want >= (int32_t)ps.sites[t].n_alleles) throw std::runtime_error("inconsistent grouped allele counts");
        open.push_back({t, want, 0});
      } else {  // a site on an allele that is not walked
        p = ps.sites[t].end;
        pb = ps.sites[t].pb_exit;
      }
    } else if (p == ps.sites[o.site].end)
      open.pop_back();
    else
      ++o.cur;
  }
  return {path, top.second};
}

DepthStats read_depth_stats(const PrgSites& ps, const Cov* per_base, const std::vector<GroupCounts>& counts) {
  // ReadStats::compute_coverage_depth (read_stats.cpp:119-160): per level-1 site the mean per-base coverage of that
  // path, or the haplogroup's count when the path holds no base (direct deletion); then mean and (population)
  // variance over the sites, taken in ascending order as gq_read_depth_stats does
  DepthStats out;
  std::vector<double> covs;
  double total = 0;
  for (uint32_t s = 0; s < ps.sites.size(); ++s) {
    if (ps.sites[s].parent >= 0) continue;
    const auto extraction = extract_max_coverage_allele(ps, s, per_base, counts);
    const double site_cov = extraction.first.pb.empty() ? (double)extraction.second : extraction.first.mean_cov();
    total += site_cov;
    covs.push_back(site_cov);
    if (extraction.second == 0) ++out.num_sites_no_cov;
  }
  out.num_sites_total = covs.size();
  out.mean = total / (double)covs.size();
  double var = 0;
  for (double c : covs) var += std::pow(c - out.mean, 2);
  out.variance = var / (double)covs.size();
  return out;
}

std::vector<GroupCounts> unpack_grouped(const uint32_t* grouped, uint64_t n_words, size_t n_sites) {
  std::vector<GroupCounts> counts(n_sites);
  for (uint64_t t = 0; t < n_words;) {
    if (t + 3 > n_words) throw std::runtime_error("grouped allele counts: truncated record");
    const uint32_t site = grouped[t], n = grouped[t + 2];
    if (site >= n_sites || t + 3 + n > n_words) throw std::runtime_error("grouped allele counts: bad record");
    std::vector<int32_t> ids(grouped + t + 3, grouped + t + 3 + n);
    counts[site].emplace_back(std::move(ids), (Cov)grouped[t + 1]);
    t += 3 + n;
  }
  return counts;
}

// ------------------------------------------------------------------------------------------------ segments

SegmentTracker::SegmentTracker(const std::string& coords_text) {  // segment_tracker.hpp:22-36
  std::istringstream in(coords_text);
  Segment next{"gramtools_prg", UINT64_MAX};
  while (in >> next.id >> next.size) {
    global_max_ += next.size;
    segments_.push_back(next);
  }
  if (segments_.empty()) {
    segments_.push_back(next);
    global_max_ = UINT64_MAX;
  }
  max_ = segments_[0].size - 1;
}

const std::string& SegmentTracker::get_id(uint64_t pos) {
  if (pos < min_ || pos >= global_max_) throw std::out_of_range("position outside the segments, or before the current one");
  while (pos > max_) {
    ++cur_;
    min_ = max_ + 1;
    max_ += segments_.at(cur_).size;
  }
  return segments_.at(cur_).id;
}

uint64_t SegmentTracker::relative_pos(uint64_t pos) const {
  if (pos < min_ || pos >= global_max_) throw std::out_of_range("position outside the segments, or before the current one");
  return pos - min_;
}

void SegmentTracker::reset() {
  min_ = 0;
  cur_ = 0;
  max_ = segments_[0].size - 1;
}

// ------------------------------------------------------------------------------------------------ genotyped.json

std::string json_number(double v) {
  // nlohmann::json 3.7 (detail::to_chars): shortest digits that read back as v, laid out as dddd.0 / d.ddd / 0.00ddd
  // for decimal exponents in (-4, 15], as d.ddde±XX outside; non-finite values are dumped as null
  if (!std::isfinite(v)) return "null";
  if (v == 0) return std::signbit(v) ? "-0.0" : "0.0";
  std::string out;
  if (v < 0) {
    out = "-";
    v = -v;
  }
  char buf[40];
  {  // shortest digits that round-trip, in scientific form: d[.ddd]e±XX
    auto r = std::to_chars(buf, buf + sizeof buf - 1, v, std::chars_format::scientific);
    *r.ptr = 0;
  }
  std::string digits;
  const char* e = std::strchr(buf, 'e');
  for (const char* c = buf; c < e; ++c)
    if (*c != '.') digits.push_back(*c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int k = (int)digits.size();
  const int n = std::atoi(e + 1) + 1;  // the decimal point sits after n digits
  if (k <= n && n <= 15) {
    out += digits + std::string((size_t)(n - k), '0') + ".0";
  } else if (0 < n && n <= 15) {
    out += digits.substr(0, (size_t)n) + "." + digits.substr((size_t)n);
  } else if (-4 < n && n <= 0) {
    out += "0." + std::string((size_t)(-n), '0') + digits;
  } else {
    out += digits.substr(0, 1);
    if (k > 1) out += "." + digits.substr(1);
    const int ex = n - 1;
    char eb[16];
    std::snprintf(eb, sizeof eb, "e%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    out += eb;
  }
  return out;
}

static std::string json_string(const std::string& s) {
  std::string o = "\"";
  for (unsigned char c : s) {
    if (c == '"') o += "\\\"";
    else if (c == '\\') o += "\\\\";
    else if (c == '\n') o += "\\n";
    else if (c == '\t') o += "\\t";
    else if (c == '\r') o += "\\r";
    else if (c < 0x20) {
      char b[8];
      std::snprintf(b, sizeof b, "\\u%04x", c);
      o += b;
    } else
      o.push_back((char)c);
  }
  return o + "\"";
}

template <typename T, typename F>
static std::string json_array(const std::vector<T>& v, F&& item) {
  std::string o = "[";
  for (size_t i = 0; i < v.size(); ++i) o += (i ? "," : "") + item(v[i]);
  return o + "]";
}

static const char* kDescGT = "Genotype";
static const char* kDescDP = "Total read depth on variant site";
static const char* kDescCOV = "Read coverage on each allele";
static const char* kDescFT = "Filters failed in a sample";
static const char* kDescAMBIG = "Ambiguous site. Different variant paths can produce the same sequence.";
static const char* kDescGTCONF = "Genotype confidence as likelihood ratio of called and next most likely genotype.";
static const char* kDescGCP = "Percent of calls expected to have lower GT_CONF";

std::string LevelGenotyper::json(const std::string& sample_id, SegmentTracker& tracker) const {
  // make_json.cpp:7-82, json_prg_spec.cpp, json_site_spec.hpp:25-31, fields.hpp:129-158. nlohmann::json keeps object
  // keys sorted and streams without white space; both are reproduced so that the file compares byte for byte
  // (up to the order of the children of a haplogroup in Child_Map, which the reference takes from an unordered_map:
  // ascending here)
  auto desc = [](const char* d) { return std::string("{\"Desc\":") + json_string(d) + "}"; };
  std::string o = "{\"Child_Map\":{";
  if (ps_.is_nested) {
    std::map<std::string, std::string> by_site;
    for (auto& e : ps_.children) {
      std::map<std::string, std::string> by_hapg;
      for (auto& h : e.second) {
        std::vector<uint32_t> kids(h.second);
        std::sort(kids.begin(), kids.end());
        by_hapg[std::to_string(h.first)] = json_array(kids, [](uint32_t k) { return std::to_string(k); });
      }
      std::string inner = "{";
      bool first = true;
      for (auto& h : by_hapg) {
        inner += (first ? "" : ",") + json_string(h.first) + ":" + h.second;
        first = false;
      }
      by_site[std::to_string(e.first)] = inner + "}";
    }
    bool first = true;
    for (auto& e : by_site) {
      o += (first ? "" : ",") + json_string(e.first) + ":" + e.second;
      first = false;
    }
  }
  o += "},\"Filters\":{\"AMBIG\":" + desc(kDescAMBIG) + "},\"Lvl1_Sites\":[";
  if (!ps_.is_nested)
    o += "\"all\"";
  else {
    bool first = true;
    for (size_t i = 0; i < ps_.sites.size(); ++i)
      if (ps_.sites[i].parent < 0) {
        o += (first ? "" : ",") + std::to_string(i);
        first = false;
      }
  }
  o += "],\"Model\":\"LevelGenotyping\",\"Samples\":[{\"Desc\":\"made by gramtools genotype\",\"Name\":" +
       json_string(sample_id) + "}],\"Site_Fields\":{";
  o += "\"ALS\":" + desc("Alleles at this site") + ",\"COV\":" + desc(kDescCOV) + ",\"DP\":" + desc(kDescDP) +
       ",\"FT\":" + desc(kDescFT) + ",\"GT\":" + desc(kDescGT) + ",\"GT_CONF\":" + desc(kDescGTCONF) +
       ",\"GT_CONF_PERCENTILE\":" + desc(kDescGCP) + ",\"HAPG\":" + desc("Sample haplogroups of genotyped alleles") +
       ",\"POS\":" + desc("Position on reference or pseudo-reference") + ",\"SEG\":" + desc("Segment ID") + "},\"Sites\":[";
  o.reserve(o.size() + 192 * sites_.size());
  auto ints = [&o](const std::vector<int32_t>& v) {
    o += '[';
    for (size_t i = 0; i < v.size(); ++i) {
      if (i) o += ',';
      o += std::to_string(v[i]);
    }
    o += ']';
  };
  for (size_t i = 0; i < sites_.size(); ++i) {
    const Site& s = sites_[i];
    const std::string& seg = tracker.get_id(s.pos);
    const uint64_t pos = tracker.relative_pos(s.pos) + 1;
    o += i ? ",{\"ALS\":[" : "{\"ALS\":[";
    for (size_t a = 0; a < s.alleles.size(); ++a) {
      if (a) o += ',';
      o += json_string(s.alleles[a].seq);
    }
    o += "],\"COV\":[[";
    for (size_t c = 0; c < s.allele_covs.size(); ++c) {
      if (c) o += ',';
      o += json_number(s.allele_covs[c]);
    }
    o += "]],\"DP\":[";
    o += std::to_string(s.total_coverage);
    o += "],\"FT\":[[";
    for (size_t f = 0; f < s.filters.size(); ++f) {
      if (f) o += ',';
      o += json_string(s.filters[f]);
    }
    o += "]],\"GT\":[";
    if (s.is_null()) o += "[null]";
    else ints(s.genotype);
    o += "],\"GT_CONF\":[";
    o += json_number(s.gt_conf);
    o += "],\"GT_CONF_PERCENTILE\":[";
    o += json_number(s.gt_conf_percentile);
    o += "],\"HAPG\":[";
    ints(s.haplogroups);
    o += "],\"POS\":";
    o += std::to_string(pos);
    o += ",\"SEG\":";
    o += json_string(seg);
    o += '}';
  }
  return o + "]}";
}

// ------------------------------------------------------------------------------------------------ personalised reference

std::string Fasta::to_string() const {  // personalised_reference.cpp:118-137
  std::string o = ">" + id + " " + desc;
  if (desc.empty() || desc.back() != '\n') o += "\n";
  size_t at = 0, left = seq.size();
  while (left > 60) {
    o.append(seq, at, 60);
    o += "\n";
    at += 60;
    left -= 60;
  }
  o.append(seq, at, left);
  return o;
}

std::vector<Fasta> LevelGenotyper::personalised_reference(SegmentTracker& tracker) const {
  // personalised_reference.cpp:8-116: the level-1 path of the PRG with every level-1 site replaced by its called
  // allele(s) (REF for a null call); one record per segment and haplotype, segments cut at REFERENCE coordinates
  size_t ploidy = 1;
  for (auto& s : sites_)
    if (!s.is_null()) {
      ploidy = s.genotype.size();
      break;
    }
  std::vector<Fasta> refs(tracker.segments().size() * ploidy);
  size_t offset = 0;
  auto name = [&](const std::string& id) {
    if (ploidy == 1) refs.at(offset).id = id;
    else
      for (size_t i = 0; i < ploidy; ++i) refs.at(offset + i).id = id + "_" + std::to_string(i + 1);
  };
  auto next_segment = [&]() -> uint64_t {
    if (tracker.edge() != tracker.global_edge()) {
      const std::string id = tracker.get_id(tracker.edge() + 1);
      offset += ploidy;
      name(id);
    }
    return tracker.edge();
  };
  uint64_t edge = tracker.edge();
  name(tracker.get_id(edge));
  const uint64_t n = ps_.prg.size();
  uint64_t ref = 0;
  for (uint64_t p = 0; p < n;) {
    const uint32_t m = ps_.prg[p];
    if (m > 4) {  // level-1 site entry (nested ones are never reached: their parent is stepped over)
      const uint32_t s = (m - 5) / 2;
      const Site& site = sites_.at(s);
      Genotype gt = site.is_null() ? Genotype(ploidy, 0) : site.genotype;
      if (gt.size() != ploidy) throw std::runtime_error("The sites do not all have the same GT cardinality (ploidy)");
      for (size_t i = 0; i < ploidy; ++i) refs.at(offset + i).seq += site.alleles.at((size_t)gt[i]).seq;
      p = (uint64_t)site.end_text + 1;
      ref = site.end_pos;
      if (edge == ref - 1) edge = next_segment();
      continue;
    }
    uint64_t q = p;
    std::string run;
    while (q < n && ps_.prg[q] <= 4) run.push_back(base_char(ps_.prg[q++]));
    uint64_t cur = ref;
    const uint64_t last = ref + run.size() - 1;
    while (cur <= last) {
      if (edge <= last) {
        const std::string piece = run.substr(cur - ref, edge - cur + 1);
        for (size_t i = 0; i < ploidy; ++i) refs.at(offset + i).seq += piece;
        cur = edge + 1;
        edge = next_segment();
      } else {
        const std::string piece = run.substr(cur - ref);
        for (size_t i = 0; i < ploidy; ++i) refs.at(offset + i).seq += piece;
        cur = last + 1;
      }
    }
    ref += run.size();
    p = q;
  }
  return refs;
}

std::string deduped_fasta_text(std::vector<Fasta> refs, const std::string& desc) {
  for (auto& r : refs) r.desc = desc;
  std::map<std::string, const Fasta*> by_seq;  // std::set<Fasta> ordered (and made unique) by sequence alone
  for (auto& r : refs) by_seq.emplace(r.seq, &r);
  std::string o;
  for (auto& e : by_seq) o += e.second->to_string() + "\n";
  return o;
}

// ------------------------------------------------------------------------------------------------ VCF

static std::string meta_line(const std::string& type, const std::string& id, const std::string& desc,
                             const std::string& num, const std::string& vtype, uint64_t length) {
  std::string o = "##" + type + "=<ID=" + id;  // vcf_meta_info_line::to_string (fields.hpp:38-55)
  if (!num.empty()) o += ",Number=" + num;
  if (!vtype.empty()) o += ",Type=" + vtype;
  if (!desc.empty()) o += ",Description=\"" + desc + "\"";
  if (length != 0) o += ",length=" + std::to_string(length);
  return o + ",Source=\"gramtools\">\n";
}

static std::string vcf_float(double v) {  // values go through float and htslib prints them with %g precision
  char b[32];
  std::snprintf(b, sizeof b, "%g", (double)(float)v);
  return b;
}

std::string LevelGenotyper::vcf(const std::string& sample_id, SegmentTracker& tracker) const {
  // make_vcf.cpp:7-149 through htslib's VCF text writer: the header bcf_hdr_init("w") starts (fileformat + PASS
  // filter), the lines appended by populate_vcf_hdr in their order, then one record per LEVEL-1 site with FORMAT keys
  // in the order they are set (GT, DP, COV — absent when the site never had coverages —, FT, GT_CONF,
  // GT_CONF_PERCENTILE). htslib is absent here: the record text is parity-unpinned.
  std::string o = "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n";
  for (auto& seg : tracker.segments()) o += meta_line("contig", seg.id, "", "", "", seg.size);
  o += "##source=gramtools\n##Model=LevelGenotyping\n";
  o += meta_line("FORMAT", "GT_CONF", kDescGTCONF, "1", "Float", 0);
  o += meta_line("FORMAT", "GT_CONF_PERCENTILE", kDescGCP, "1", "Float", 0);
  o += meta_line("FORMAT", "GT", kDescGT, "1", "String", 0);
  o += meta_line("FORMAT", "DP", kDescDP, "1", "Integer", 0);
  o += meta_line("FORMAT", "COV", kDescCOV, "R", "Float", 0);
  o += meta_line("FORMAT", "FT", kDescFT, "1", "String", 0);
  o += meta_line("FILTER", "AMBIG", kDescAMBIG, "", "", 0);
  o += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + sample_id + "\n";
  o.reserve(o.size() + 96 * sites_.size());
  for (size_t i = 0; i < sites_.size(); ++i) {
    if (ps_.sites[i].parent >= 0) continue;  // next_valid_idx (:52-64)
    const Site& s = sites_[i];
    o += tracker.get_id(s.pos);
    o += '\t';
    o += std::to_string(tracker.relative_pos(s.pos) + 1);
    o += "\t.\t";
    if (s.alleles.empty()) o += '.';
    else o += s.alleles[0].seq;
    o += '\t';
    if (s.alleles.size() < 2) o += '.';
    for (size_t a = 1; a < s.alleles.size(); ++a) {
      if (a > 1) o += ',';
      o += s.alleles[a].seq;
    }
    // FORMAT keys in the order they are set; COV is absent when the site never had coverages
    o += s.allele_covs.empty() ? "\t.\t.\t.\tGT:DP:FT:GT_CONF:GT_CONF_PERCENTILE\t"
                               : "\t.\t.\t.\tGT:DP:COV:FT:GT_CONF:GT_CONF_PERCENTILE\t";
    if (s.is_null()) o += '.';
    else
      for (size_t g = 0; g < s.genotype.size(); ++g) {
        if (g) o += '/';
        o += std::to_string(s.genotype[g]);
      }
    o += ':';
    o += std::to_string(s.total_coverage);
    if (!s.allele_covs.empty()) {
      o += ':';
      for (size_t c = 0; c < s.allele_covs.size(); ++c) {
        if (c) o += ',';
        o += vcf_float(s.allele_covs[c]);
      }
    }
    // only the first string reaches bcf_update_format_string (n = 1); with several filters it carries its comma
    o += ':';
    if (s.filters.empty()) o += "PASS";
    else {
      o += s.filters[0];
      if (s.filters.size() > 1) o += ',';
    }
    o += ':';
    o += vcf_float(s.gt_conf);
    o += ':';
    o += vcf_float(s.gt_conf_percentile);
    o += '\n';
  }
  return o;
}

std::string bgzf_compress(const std::string& s) {
  // BGZF (SAM spec §4.1): gzip members of at most 64 KiB each carrying a 'BC' extra field with the member's size,
  // then the empty end-of-file member; any gzip reader takes the concatenation
  std::string out;
  auto put16 = [&](std::string& b, uint32_t v) {
    b.push_back((char)(v & 0xFF));
    b.push_back((char)((v >> 8) & 0xFF));
  };
  auto put32 = [&](std::string& b, uint32_t v) {
    put16(b, v & 0xFFFF);
    put16(b, v >> 16);
  };
  auto block = [&](const char* data, size_t len) {
    z_stream z{};
    if (deflateInit2(&z, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
      throw std::runtime_error("deflateInit2 failed");
    std::vector<unsigned char> comp(deflateBound(&z, (uLong)len) + 16);
    z.next_in = (Bytef*)data;
    z.avail_in = (uInt)len;
    z.next_out = comp.data();
    z.avail_out = (uInt)comp.size();
    const int rc = deflate(&z, Z_FINISH);
    const size_t clen = z.total_out;
    deflateEnd(&z);
    if (rc != Z_STREAM_END) throw std::runtime_error("deflate failed");
    std::string b;
    const unsigned char head[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
    b.append((const char*)head, 12);
    b += "BC";
    put16(b, 2);
    put16(b, (uint32_t)(clen + 25));  // total block size - 1
    b.append((const char*)comp.data(), clen);
    put32(b, (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)data, (uInt)len));
    put32(b, (uint32_t)len);
    out += b;
  };
  constexpr size_t kBlock = 0xFF00;
  for (size_t at = 0; at < s.size(); at += kBlock) block(s.data() + at, std::min(kBlock, s.size() - at));
  block("", 0);
  return out;
}

}  // namespace lg
}  // namespace gq
