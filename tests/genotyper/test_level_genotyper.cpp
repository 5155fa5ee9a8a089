// TEST-ONLY. Known-answer checks of gramtools_b200/csrc/level_genotyper.cpp: the expectations of the reference's own
// gtest cases for the genotyping step, transcribed as data (each block cites the test it comes from, all under
// /root/reference/libgramtools/tests/genotype/infer/). Built and run by tests/test_level_genotyper.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <set>
#include <string>
#include <vector>

#include "../../gramtools_b200/csrc/level_genotyper.hpp"

using namespace gq::lg;

static int g_checks = 0, g_failed = 0;
static const char* g_case = "";
#define CHECK(cond)                                                            \
  do {                                                                         \
    ++g_checks;                                                                \
    if (!(cond)) {                                                             \
      ++g_failed;                                                              \
      std::printf("FAILED [%s] %s:%d: %s\n", g_case, __FILE__, __LINE__, #cond); \
    }                                                                          \
  } while (0)
#define CASE(name) g_case = name
static bool near(double a, double b, double rel = 1e-6) { return std::fabs(a - b) <= rel * std::max(1.0, std::fabs(b)); }
static bool ulp_eq(double a, double b) { return std::fabs(a - b) <= 4 * 2.220446049250313e-16 * std::fabs(b); }
template <typename F>
static bool throws(F&& f) {
  try {
    f();
  } catch (const std::exception&) {
    return true;
  }
  return false;
}

// "AT[GC[C,A]T,TTA]T" -> integers: sites numbered 5, 7, … in opening order, "," and "]" the site's even marker
static std::vector<uint32_t> bracketed(const std::string& s) {
  std::vector<uint32_t> out, open;
  uint32_t next = 5;
  for (char c : s) {
    if (c == '[') {
      open.push_back(next);
      out.push_back(next);
      next += 2;
    } else if (c == ',')
      out.push_back(open.back() + 1);
    else if (c == ']') {
      out.push_back(open.back() + 1);
      open.pop_back();
    } else
      out.push_back(c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : 4);
  }
  return out;
}
static PrgSites parsed(const std::string& s) {
  auto v = bracketed(s);
  return parse_prg_sites(v.data(), v.size());
}
static Site mock_site(Alleles alleles, Genotype gt) {
  Site s;
  s.alleles = std::move(alleles);
  s.genotype = std::move(gt);
  return s;
}
static bool same_alleles(const Alleles& a, const Alleles& b) { return a == b; }

static void test_probabilities() {
  CASE("LikelihoodStats.DynamicChoiceOfProbDistribution / DynamicDataParams (test_probabilities.cpp:20-58)");
  LStats l = make_l_stats(10, 5, 0.01);
  CHECK(l.pmf_full_depth->is_poisson());
  CHECK(l.mean_cov == 10. && l.mean_pb_error == 0.01 && l.num_successes == -1 && l.success_prob == -1);
  l = make_l_stats(10, 15, 0.01);
  CHECK(!l.pmf_full_depth->is_poisson());
  l = make_l_stats(10, 20, 0.01);
  CHECK(l.num_successes == 10. && l.success_prob == 0.5);
  CHECK((int)(l.num_successes * (1 - l.success_prob) / l.success_prob) == 10);
  CHECK((int)(l.num_successes * (1 - l.success_prob) / std::pow(l.success_prob, 2)) == 20);

  CASE("LogPmfs (test_probabilities.cpp:60-96)");
  PoissonLogPmf p2(2);
  CHECK(p2.n_memoised() == 1);
  CHECK(p2(0) == -2);
  NegBinomLogPmf nb(2, 0.5);
  CHECK(nb.n_memoised() == 1);
  CHECK(near(p2(2), -1.3068528194400546, 1e-7));
  PoissonLogPmf p25(2.5);
  CHECK(ulp_eq(p25(2), -1.3605657168116352));
  CHECK(ulp_eq(nb(2), -1.6739764335716716));
  NegBinomLogPmf nb25(2.5, 0.5);
  CHECK(ulp_eq(nb25(4), -2.3056313146033682));
  CHECK(p2.n_memoised() == 2);  // ProbabilityMemoisation (:8-18)

  CASE("MinCovMoreLikelyThanError (test_probabilities.cpp:98-121)");
  const double depths[3] = {10, 10, 100}, errs[3] = {0.0001, 0.001, 0.001};
  const Cov expected[3] = {1, 2, 10};
  for (int i = 0; i < 3; ++i) {
    PoissonLogPmf pmf(depths[i]);
    CHECK(find_minimum_non_error_cov(errs[i], pmf) == expected[i]);
  }
  PoissonLogPmf zero(0);
  (void)find_minimum_non_error_cov(0.01, zero);  // terminates
  CHECK(true);
}

static void test_model_coverages() {
  CASE("HaploidCoverages (test_model.cpp:14-37)");
  {
    GroupCounts g{{{0}, 5}, {{1}, 10}, {{3}, 1}};
    SiteModel m;
    m.set_haploid_coverages(g, 4);
    CHECK((m.haploid_covs() == std::vector<Cov>{5, 10, 0, 1}));
    CHECK((m.singleton_covs() == std::vector<Cov>{5, 10, 0, 1}));
  }
  {
    GroupCounts g{{{0}, 5}, {{0, 1}, 4}, {{1}, 10}, {{2, 3}, 1}};
    SiteModel m;
    m.set_haploid_coverages(g, 4);
    CHECK((m.haploid_covs() == std::vector<Cov>{9, 14, 1, 1}));
    CHECK((m.singleton_covs() == std::vector<Cov>{5, 10, 0, 0}));
  }
  CASE("DiploidCoverages (test_model.cpp:39-75)");
  {
    GroupCounts g{{{0}, 7}, {{0, 1}, 4}, {{1}, 20}, {{0, 3}, 3}, {{2, 3}, 1}};
    SiteModel m;
    m.set_haploid_coverages(g, 4);
    auto c = m.diploid_coverage(g, {0, 1}, std::vector<bool>(4, false));
    CHECK(near(c.first, 10 + 4 / 3.) && near(c.second, 20 + 8 / 3.));
  }
  {
    GroupCounts g{{{0, 1}, 3}, {{2, 3}, 1}};
    SiteModel m;
    m.set_haploid_coverages(g, 4);
    auto c = m.diploid_coverage(g, {0, 1}, std::vector<bool>(4, false));
    CHECK(near(c.first, 1.5) && near(c.second, 1.5));
  }
  CASE("LevelGenotyperModelDirectDeletion (test_model.cpp:77-97)");
  {
    Alleles a{{"C", {8}, 0}, {"G", {8}, 0}, {"", {}, 1}};
    GroupCounts g{{{0}, 8}, {{1}, 8}, {{0, 1}, 1}};
    SiteModel m;
    m.set_haploid_coverages(g, 2);
    m.assign_coverage_to_empty_alleles(a);
    CHECK((a[2].pb == std::vector<Cov>{9}) && (a[0].pb == std::vector<Cov>{8}));
  }
  CASE("DiploidCoveragesOneDominatingClass (test_model.cpp:99-136)");
  {
    GroupCounts g{{{0}, 8}, {{0, 1}, 4}};
    SiteModel m;
    m.set_haploid_coverages(g, 2);
    auto c = m.diploid_coverage(g, {0, 1}, std::vector<bool>(2, false));
    CHECK(near(c.first, 12) && near(c.second, 0));
    SiteModel m2;
    m2.set_haploid_coverages(g, 2);
    auto d = m2.diploid_coverage(g, {0, 0}, std::vector<bool>{true});
    CHECK(near(d.first, 6) && near(d.second, 6));
  }
  CASE("CountCrediblePositions / CountTotalCov / CountNumHaplogroups (test_model.cpp:138-181)");
  {
    LStats l;
    l.credible_cov_t = 3;
    SiteModel m(l, {}, {});
    CHECK(m.fraction_noncredible_positions(Allele{"ATCGCCG", {0, 0, 2, 3, 3, 5, 4, 4}, 0}) == 0.375);
    CHECK(SiteModel::count_total_coverage({}) == 0);
    CHECK(SiteModel::count_total_coverage({{{0}, 5}, {{0, 1}, 4}, {{1}, 10}, {{2, 3}, 1}}) == 20);
    CHECK((SiteModel::haplogroup_multiplicities({{"", {}}, {"", {}}}) == std::vector<bool>{true}));
    CHECK((SiteModel::haplogroup_multiplicities({{"", {}, 0}, {"", {}, 1}, {"", {}, 1}}) == std::vector<bool>{false, true}));
  }
  CASE("MakePermutations / RescaleGenotypes (test_model.cpp:183-221)");
  {
    CHECK((SiteModel::combinations({1, 4, 5}, 2) == std::vector<Genotype>{{1, 4}, {1, 5}, {4, 5}}));
    auto u = SiteModel::combinations({4, 3, 2}, 2);
    std::sort(u.begin(), u.end());
    CHECK((u == std::vector<Genotype>{{2, 3}, {2, 4}, {3, 4}}));
    CHECK(SiteModel::combinations({1}, 2).empty());
    CHECK((SiteModel::rescale_genotypes({1, 3}) == Genotype{1, 2}));
    CHECK((SiteModel::rescale_genotypes({0, 4, 4}) == Genotype{0, 1, 1}));
    CHECK((SiteModel::rescale_genotypes({4, 2}) == Genotype{1, 2}));
  }
}

static void test_model_calls() {
  CASE("TestLevelGenotyperModel_Failure (test_model.cpp:227-236)");
  {
    LStats l;
    CHECK(throws([&] { SiteModel m(Alleles{{"ACGT", {1, 1, 1, 1}, 0}}, {}, Ploidy::Haploid, &l); }));
  }
  CASE("TestLevelGenotyperModel_NullGTs (test_model.cpp:238-288)");
  {
    Alleles alleles{{"A", {0}, 0}, {"G", {0}, 1}};
    LStats l = make_l_stats(15, 0, 0.01);
    {
      Alleles dup(alleles);
      dup.emplace_back("A", std::vector<Cov>{1}, 1);
      SiteModel m(dup, {}, Ploidy::Haploid, &l);
      CHECK(m.site().is_null() && m.site().has_filter("AMBIG"));
    }
    {
      LStats l0 = l;
      l0.mean_cov = 0;
      SiteModel m(alleles, {}, Ploidy::Haploid, &l0);
      CHECK(m.site().is_null());
    }
    {
      SiteModel m(alleles, {}, Ploidy::Haploid, &l);
      CHECK(m.site().is_null());
      CHECK(m.site().alleles.size() == 1 && m.site().alleles[0].seq == "A");
    }
    {
      SiteModel m(alleles, {{{0}, 5}, {{1}, 5}}, Ploidy::Haploid, &l);
      CHECK(m.site().is_null());
      CHECK(m.site().extra_alleles.has_value() && same_alleles(*m.site().extra_alleles, alleles));
    }
  }
  CASE("TestLevelGenotyperModel_GTCalls (test_model.cpp:290-328)");
  {
    Alleles alleles{{"ATC", {0, 0, 1}, 0}, {"GGGCC", {10, 12, 12, 14, 14}, 1}};
    GroupCounts g{{{0}, 1}, {{1}, 13}};
    LStats l = make_l_stats(15, 0, 0.01);
    SiteModel dip(alleles, g, Ploidy::Diploid, &l);
    CHECK((dip.site().genotype == Genotype{1, 1}));
    SiteModel hap(alleles, g, Ploidy::Haploid, &l);
    CHECK((hap.site().genotype == Genotype{1}));
    CHECK(same_alleles(hap.site().alleles, alleles));  // REF is reported although it was not called
    CHECK((hap.site().haplogroups == std::vector<int32_t>{1}));
    CHECK(hap.site().total_coverage == 14);
    CHECK((hap.site().allele_covs == std::vector<double>{1., 13.}));
    LStats lnb = make_l_stats(15, 16, 0.01);
    SiteModel nb(alleles, g, Ploidy::Haploid, &lnb);
    CHECK((nb.site().genotype == Genotype{1}));
  }
  CASE("TestLevelGenotyperModel_ExtraAlleles (test_model.cpp:330-378)");
  {
    Alleles alleles{{"A", {0}, 0}, {"G", {0}, 1}};
    Likelihoods different{{-4, {0}}, {-2, {1}}};
    LStats l = make_l_stats(40, 0, 0.01);
    std::vector<bool> mults{false, false};
    SiteModel m1(l, {1, 39, 1}, different);
    m1.call_genotype(alleles, mults, Ploidy::Haploid);
    CHECK(!m1.site().extra_alleles.has_value());
    SiteModel m2(l, {1, 39}, Likelihoods{{-2, {0}}, {-2, {1}}});
    m2.call_genotype(alleles, mults, Ploidy::Haploid);
    CHECK(m2.site().extra_alleles.has_value() && same_alleles(*m2.site().extra_alleles, alleles));
    CHECK((*m2.site().extra_alleles)[0].callable && (*m2.site().extra_alleles)[1].callable);
    SiteModel m3(l, {1, 5}, different);  // low total coverage against a mean of 40
    m3.call_genotype(alleles, mults, Ploidy::Haploid);
    CHECK(m3.site().extra_alleles.has_value() && same_alleles(*m3.site().extra_alleles, Alleles{alleles[0]}));
    CHECK(!(*m3.site().extra_alleles)[0].callable);
    SiteModel m4(l, {20, 21}, different);  // low relative coverage
    m4.call_genotype(alleles, mults, Ploidy::Haploid);
    CHECK(m4.site().extra_alleles.has_value() && same_alleles(*m4.site().extra_alleles, Alleles{alleles[0]}));
  }
  CASE("TestLevelGenotyperModel_IgnoredREF (test_model.cpp:380-431)");
  {
    LStats l = make_l_stats(10, 0, 0.01);
    Alleles alleles{{"A", {10}, 0, false}, {"C", {9}, 1}, {"G", {10}, 2}};
    GroupCounts g{{{0}, 20}, {{1}, 9}, {{2}, 10}};
    SiteModel hap(alleles, g, Ploidy::Haploid, &l, true);
    CHECK(hap.likelihoods().size() == 2);
    SiteModel dip(alleles, g, Ploidy::Diploid, &l, true);
    CHECK(dip.likelihoods().size() == 3);
    CHECK(same_alleles(hap.site().alleles, Alleles{alleles[0], alleles[2]}));
    CHECK((hap.site().genotype == Genotype{1}));
    CHECK(same_alleles(dip.site().alleles, alleles));
    CHECK((dip.site().genotype == Genotype{1, 2}));
  }
  CASE("TestLevelGenotyperModel homozygous / nested scenario / four alleles (test_model.cpp:433-505)");
  {
    LStats l = make_l_stats(20, 0, 0.01);
    SiteModel m(Alleles{{"AA", {0, 1}, 0}, {"TT", {20, 19}, 1}}, {{{0}, 2}, {{0, 1}, 1}, {{1}, 20}}, Ploidy::Diploid, &l);
    CHECK((m.site().genotype == Genotype{1, 1}));
    LStats wide = make_l_stats(20, 200, 0.01);
    SiteModel gap(Alleles{{"AAAACAG", {0, 20, 20, 20, 20, 20, 0}, 0}, {"TAAACAT", {20, 20, 20, 20, 20, 20, 20}, 0}},
                  {{{0}, 20}}, Ploidy::Haploid, &wide);
    CHECK((gap.site().genotype == Genotype{1}));
    LStats l30 = make_l_stats(30, 0, 0.01);
    Alleles four{{"AATAA", {8, 8, 8, 8, 8}, 0}, {"AAGAA", {7, 7, 7, 7, 7}, 0}, {"GGTGG", {15, 15, 15, 16, 16}, 1},
                 {"GGCGG", {14, 14, 14, 15, 15}, 1}};
    GroupCounts g{{{0}, 15}, {{1}, 30}};
    SiteModel hap(four, g, Ploidy::Haploid, &l30);
    CHECK(hap.likelihoods().size() == 4);
    SiteModel dip(four, g, Ploidy::Diploid, &l30);
    CHECK(dip.likelihoods().size() == 10);
  }
  CASE("TestMaxLikelihoodCall (test_model.cpp:507-581)");
  {
    Likelihoods lk{{-1, {0}}, {-2, {1}}, {-3, {2}}, {-4, {3}}};
    Alleles alleles{{"A", {}}, {"B", {}}, {"C", {}}, {"D", {}}};
    CHECK(throws([&] { SiteModel::choose_max_likelihood(Likelihoods{*lk.begin()}, Alleles{}); }));
    CHECK(SiteModel::choose_max_likelihood(lk, alleles) == lk.begin());
    Alleles a0(alleles);
    a0[0].callable = false;
    CHECK(SiteModel::choose_max_likelihood(lk, a0) == std::next(lk.begin()));
    Alleles a1(alleles);
    a1[1].callable = false;
    CHECK(SiteModel::choose_max_likelihood(lk, a1) == lk.begin());
    Alleles a3(alleles);
    a3[0].callable = a3[1].callable = a3[2].callable = false;
    CHECK(throws([&] { SiteModel::choose_max_likelihood(lk, a3); }));
    LStats l = make_l_stats(20, 5, 0.01);
    SiteModel m(l, {20, 15, 12, 8}, lk);
    m.call_genotype(a0, std::vector<bool>{false}, Ploidy::Haploid);
    CHECK(same_alleles(m.site().alleles, Alleles{a0[0], a0[1]}));
    CHECK((m.site().genotype == Genotype{1}));
  }
}

static void test_allele_extraction() {
  CASE("ExtractRefAllele (test_allele_extracter.cpp:13-22)");
  {
    PrgSites ps = parsed("AT[[C,A,G]T[G[,C]C,T],TTA]T");
    std::vector<Cov> cov(64, 0);
    Allele ref = extract_ref_allele(ps, ps.sites[0].entry + 1, 0, cov.data());
    CHECK(ref.hapg == 0 && ref.seq == "CTGC");
  }
  CASE("AlleleCombineTest (test_allele_extracter.cpp:24-110)");
  {
    Alleles existing{{"ATTG", {0, 1, 2, 3}, 0}, {"ATCG", {0, 0, 1, 1}, 0}};
    auto r = combine_with_site(Alleles{existing[0]}, mock_site({{"CCC", {1, 1, 1}, 2}}, {0}));
    CHECK(same_alleles(r, Alleles{{"ATTGCCC", {0, 1, 2, 3, 1, 1, 1}, 0}}));
    Site extra = mock_site({{"CCC", {1, 1, 1}}, {"GGG", {2, 2, 2}}}, {1});
    extra.extra_alleles = Alleles{{"AAA", {2, 1, 0}, 2, false}};
    r = combine_with_site(Alleles{existing[0]}, extra);
    CHECK(same_alleles(r, Alleles{{"ATTGGGG", {0, 1, 2, 3, 2, 2, 2}, 0}, {"ATTGAAA", {0, 1, 2, 3, 2, 1, 0}, 0}}));
    CHECK(r.size() == 2 && r[0].callable && !r[1].callable);
    r = combine_with_site(Alleles{existing[0]}, mock_site({{"TTT", {1, 1, 1}}, {"CCC", {0, 1, 1}}}, {-1}));
    CHECK(same_alleles(r, Alleles{{"ATTGTTT", {0, 1, 2, 3, 1, 1, 1}, 0}}) && r[0].callable);
    r = combine_with_site(existing, mock_site({{"CCC", {1, 1, 1}, 0}, {"TTT", {5, 5, 5}, 1}}, {0, 1}));
    CHECK(same_alleles(r, Alleles{{"ATTGCCC", {0, 1, 2, 3, 1, 1, 1}, 0},
                                  {"ATTGTTT", {0, 1, 2, 3, 5, 5, 5}, 0},
                                  {"ATCGCCC", {0, 0, 1, 1, 1, 1, 1}, 0},
                                  {"ATCGTTT", {0, 0, 1, 1, 5, 5, 5}, 0}}));
  }
  CASE("AlleleExtracter_NestedPRG (test_allele_extracter.cpp:131-225)");
  {
    PrgSites ps = parsed("AT[GCC[C,A,G]T,TTA]T");
    std::vector<Cov> cov(64, 0);
    std::vector<Site> sites(2);
    auto r = extract_alleles(ps, 1, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"C", {0}, 0}, {"A", {0}, 1}, {"G", {0}, 2}}) && r[0].callable);
    sites[1] = mock_site({{"C", {0}, 0}}, {0});
    r = extract_alleles(ps, 0, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"GCCCT", {0, 0, 0, 0, 0}, 0}, {"TTA", {0, 0, 0}, 1}}));
    sites[1] = mock_site({{"C", {0}, 0}, {"A", {0}, 1}, {"G", {0}, 2}}, {0, 1, 2});
    r = extract_alleles(ps, 0, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"GCCCT", {0, 0, 0, 0, 0}, 0}, {"GCCAT", {0, 0, 0, 0, 0}, 0},
                                  {"GCCGT", {0, 0, 0, 0, 0}, 0}, {"TTA", {0, 0, 0}, 1}}) && r[0].callable);
    sites[1] = mock_site({{"C", {0}, 0}, {"G", {0}, 2}}, {1});
    r = extract_alleles(ps, 0, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"GCCCT", {0, 0, 0, 0, 0}, 0}, {"GCCGT", {0, 0, 0, 0, 0}, 0}, {"TTA", {0, 0, 0}, 1}}));
    CHECK(!r[0].callable);
    sites[1].extra_alleles = Alleles{{"A", {0}, 1}};
    r = extract_alleles(ps, 0, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"GCCCT", {0, 0, 0, 0, 0}, 0}, {"GCCGT", {0, 0, 0, 0, 0}, 0},
                                  {"GCCAT", {0, 0, 0, 0, 0}, 0}, {"TTA", {0, 0, 0}, 1}}));
  }
  CASE("AlleleExtracter_DirectDeletionPRG (test_allele_extracter.cpp:227-245)");
  {
    PrgSites ps = parsed("AT[GCC,TTA,]T");
    std::vector<Cov> cov(64, 0);
    auto r = extract_alleles(ps, 0, cov.data(), {});
    CHECK(same_alleles(r, Alleles{{"GCC", {0, 0, 0}, 0}, {"TTA", {0, 0, 0}, 1}, {"", {}, 2}}));
  }
  CASE("per-base coverage follows the flat layout (in-site bases in PRG order)");
  {
    PrgSites ps = parsed("AT[GCC[C,A,G]T,TTA]T");
    // in-site bases in PRG order: G C C | C | A | G | T | T T A  -> offsets 0..9
    std::vector<Cov> cov{10, 11, 12, 20, 21, 22, 30, 40, 41, 42};
    std::vector<Site> sites(2);
    sites[1] = mock_site({{"C", {20}, 0}, {"G", {22}, 2}}, {1});
    auto r = extract_alleles(ps, 0, cov.data(), sites);
    CHECK(same_alleles(r, Alleles{{"GCCCT", {10, 11, 12, 20, 30}, 0}, {"GCCGT", {10, 11, 12, 22, 30}, 0},
                                  {"TTA", {40, 41, 42}, 1}}));
    CHECK(ps.sites[0].pos == 2 && ps.sites[0].end_pos == 7 && ps.sites[1].pos == 5 && ps.sites[1].end_pos == 6);
    CHECK(ps.sites[1].parent == 0 && ps.sites[1].parent_hapg == 0 && ps.is_nested && ps.ref_length == 8);
  }
}

static void test_site_interface() {
  CASE("GetUniqueGenotypedAlleles / NonGenotypedHaplogroups (test_interfaces.cpp:9-86), Alleles (test_types.cpp:6-19)");
  Alleles three{{"CCC", {1, 1, 1}}, {"GGG", {1, 1, 1}}, {"TTT", {1, 1, 1}}};
  CHECK(same_alleles(mock_site(three, {0, 0, 1}).unique_genotyped_alleles(), Alleles{three[0], three[1]}));
  CHECK(same_alleles(mock_site(three, {2, 0}).unique_genotyped_alleles(), Alleles{three[0], three[2]}));
  CHECK(mock_site(three, {-1}).unique_genotyped_alleles().empty());
  Site s = mock_site({{"ACGT", {1, 1, 1, 1}, 0}, {"TTTA", {1, 8, 1, 1}, 1}, {"TATA", {1, 8, 2, 1}, 1}}, {1, 2});
  s.num_haplogroups = 5;
  CHECK((s.non_genotyped_haplogroups() == std::vector<int32_t>{0, 2, 3, 4}));
  Allele joined = Allele{"ATA", {0, 1, 0}, 0}.joined(Allele{"TT", {2, 0}, 1});
  CHECK(joined == (Allele{"ATATT", {0, 1, 0, 2, 0}, 0}));
  CHECK((Allele{"ATAT", {2, 5, 0, 3}, 0}.mean_cov() == 2.5));
  Site called = mock_site(three, {1});
  called.total_coverage = 7;
  called.gt_conf = 3.5;
  called.allele_covs = {1., 6.};
  called.haplogroups = {1};
  called.make_null();  // interfaces.hpp:83-87, site.cpp:45-48: only the genotype, the depth and the confidences go
  CHECK(called.is_null() && called.total_coverage == 0 && called.gt_conf == 0. && called.alleles.size() == 3);
  CHECK(called.allele_covs.size() == 2 && called.haplogroups.size() == 1);
}

static void test_read_depth() {
  CASE("MaxHaplogroupCoverage (test_read_stats.cpp:50-62)");
  CHECK((max_cov_haplogroup({}) == std::pair<int32_t, Cov>{0, 0}));
  CHECK((max_cov_haplogroup({{{0, 1}, 2}, {{0}, 3}, {{1}, 4}}) == std::pair<int32_t, Cov>{1, 6}));
  CHECK((max_cov_haplogroup({{{0}, 4}, {{1}, 4}}) == std::pair<int32_t, Cov>{0, 4}));  // a tie: the lower id
  CASE("TestReadMappingStats.ExtractMaxCovAllele* (test_read_stats.cpp:64-115)");
  {
    PrgSites ps = parsed("[AC[T,G]AC,GT[A,T]T]A[AA,C]T");
    std::vector<GroupCounts> counts{{{{1}, 60}}, {{{1}, 2}, {{0}, 1}}, {{{0}, 19}, {{0, 1}, 1}}, {}};
    std::vector<Cov> cov(32, 0);
    auto e = extract_max_coverage_allele(ps, 1, cov.data(), counts);
    CHECK(e.first.seq == "G" && e.second == 2);
    e = extract_max_coverage_allele(ps, 2, cov.data(), counts);
    CHECK(e.first.seq == "A" && e.second == 20);
    e = extract_max_coverage_allele(ps, 3, cov.data(), counts);
    CHECK(e.first.seq == "AA" && e.second == 0);
    e = extract_max_coverage_allele(ps, 0, cov.data(), counts);
    CHECK(e.first.seq == "GTAT" && e.second == 60);
  }
  CASE("TestMeanAndVarCovComputation (test_read_stats.cpp:117-134): sites with coverage 15 and a deletion with 5 reads");
  {
    PrgSites ps = parsed("A[AT,T]C[T,]");
    std::vector<Cov> cov{10, 20, 0, 0};
    std::vector<GroupCounts> counts{{{{0}, 20}}, {{{1}, 5}}};
    DepthStats d = read_depth_stats(ps, cov.data(), counts);
    CHECK(d.mean == 10 && d.variance == 25 && d.num_sites_total == 2 && d.num_sites_no_cov == 0);
  }
}

static void test_runner_logic() {
  CASE("LevelGenotyperInvalidation (test_runner.cpp:173-192)");
  {
    PrgSites ps;
    ps.children[0][0] = {1};
    ps.children[0][1] = {2, 3};
    LevelGenotyper g(ps, {});
    CHECK((g.haplogroups_with_sites(0, {0, 1, 2, 3}) == std::vector<int32_t>{0, 1}));
    CHECK(g.haplogroups_with_sites(1, {0, 1, 2, 3}).empty());
  }
  CASE("LevelGenotyperPropagation (test_runner.cpp:194-244)");
  {
    PrgSites ps;  // site 1 on haplogroup 0 of site 0, site 2 on haplogroup 1 of site 1
    ps.children[0][0] = {1};
    ps.children[1][1] = {2};
    std::vector<Site> sites(3);
    sites[1].num_haplogroups = 5;
    sites[2].num_haplogroups = 5;
    LevelGenotyper g(ps, sites);
    CHECK(!g.sites()[2].is_null());
    g.invalidate_if_needed(1, {1});
    CHECK(g.sites()[2].is_null());
    CHECK(!g.sites()[1].is_null());
    g.invalidate_if_needed(0, {0});
    CHECK(g.sites()[1].is_null());
    LevelGenotyper down(ps, std::vector<Site>(3));
    down.downpropagate_filter("AMBIG", 0);
    CHECK(down.sites()[1].has_filter("AMBIG") && down.sites()[2].has_filter("AMBIG"));
    LevelGenotyper up(ps, std::vector<Site>(3));
    up.sites()[1].set_filter("AMBIG");
    up.uppropagate_filter("AMBIG", 0);
    CHECK(up.sites()[0].has_filter("AMBIG"));
  }
  CASE("GCPSimulation (test_runner.cpp:153-171)");
  {
    LStats l = make_l_stats(20, 10, 0.1);
    std::vector<Site> sites(10000);
    for (auto& s : sites) s.gt_conf = 10;
    auto conf = LevelGenotyper::gtconf_distribution(sites, l, Ploidy::Haploid, 42);
    CHECK(conf.size() == 10000 && std::set<double>(conf.begin(), conf.end()).size() == 1);
    sites.resize(10);
    conf = LevelGenotyper::gtconf_distribution(sites, l, Ploidy::Haploid, 42);
    CHECK(conf.size() == 10000 && std::is_sorted(conf.begin(), conf.end()));
    LStats lnb = make_l_stats(20, 60, 0.01);
    conf = LevelGenotyper::gtconf_distribution(sites, lnb, Ploidy::Diploid, 42);
    CHECK(conf.size() == 10000 && conf.back() > 0);
  }
  CASE("confidence simulation at a depth where 16-bit Poisson draws of libstdc++ would never be accepted");
  {
    LStats deep = make_l_stats(29733.4, 112328443.1, 0.2);
    auto conf = LevelGenotyper::gtconf_distribution(std::vector<Site>(27), deep, Ploidy::Haploid, 42);
    CHECK(conf.size() == 10000);
    LStats deep_pois = make_l_stats(50000, 10, 0.01);
    conf = LevelGenotyper::gtconf_distribution(std::vector<Site>(3), deep_pois, Ploidy::Diploid, 42);
    CHECK(conf.size() == 10000);
  }
  CASE("Percentiler (lib/GCP/GCP.h:104-183)");
  {
    Percentiler p({1, 2, 2, 2, 3, 4, 5, 6, 7, 10});
    CHECK(p.percentile(0.5) == 0.0 && p.percentile(11) == 100.0 && p.percentile(10) == 100.0);
    CHECK(near(p.percentile(1), 10.0));
    CHECK(near(p.percentile(2), 30.0));    // ranks 2..4: 20 + (40 - 20) / 2
    CHECK(near(p.percentile(2.5), 40.0));  // between (2, 30) and (3, 50)
    CHECK(throws([] { Percentiler one({1.0}); }));
  }
}

static void test_segments_and_outputs() {
  CASE("SegmentTrackerTest (test_segment_tracker.cpp:21-63)");
  {
    SegmentTracker none;
    CHECK(none.get_id(1000) == "gramtools_prg" && none.get_id(40000) == "gramtools_prg");
    SegmentTracker t("chr1\t2200\nchr2\t400\n");
    CHECK(throws([&] { t.get_id(40000); }));
    CHECK(t.get_id(2200) == "chr2");
    CHECK(throws([&] { t.get_id(200); }));
    t.reset();
    CHECK(t.global_edge() == 2599 && t.edge() == 2199);
    CHECK(t.get_id(400) == "chr1" && t.get_id(2500) == "chr2" && t.edge() == 2599);
    CHECK(t.relative_pos(2500) == 300);
    t.reset();
    CHECK(t.get_id(100) == "chr1");
  }
  CASE("Personalised_Ref (test_personalised_reference.cpp:47-217)");
  {
    PrgSites ps = parsed("AT[CG[C,G]T,C]TT[AT,TT][C,G]");
    auto make = [&](std::vector<Genotype> gts) {
      std::vector<Site> sites(4);
      sites[0].alleles = {{"CGCT", {}, 0}, {"CGGT", {}, 0}, {"C", {}, 1}};
      sites[1].alleles = {{"C", {}}, {"G", {}}};
      sites[2].alleles = {{"AT", {}}, {"TT", {}}};
      sites[3].alleles = {{"C", {}}, {"G", {}}};
      for (size_t i = 0; i < 4; ++i) {
        sites[i].genotype = gts[i];
        sites[i].end_text = ps.sites[i].end;
        sites[i].end_pos = ps.sites[i].end_pos;
        sites[i].pos = ps.sites[i].pos;
      }
      return LevelGenotyper(ps, sites);
    };
    auto seqs = [](const std::vector<Fasta>& f) {
      std::vector<std::string> s;
      for (auto& r : f) s.push_back(r.seq);
      return s;
    };
    const Genotype null{-1};
    SegmentTracker one;
    CHECK((seqs(make({null, null, null, null}).personalised_reference(one)) == std::vector<std::string>{"ATCGCTTTATC"}));
    one.reset();
    CHECK((seqs(make({{2}, null, {1}, {1}}).personalised_reference(one)) == std::vector<std::string>{"ATCTTTTG"}));
    one.reset();
    CHECK((seqs(make({{1, 2}, null, {0, 1}, {0, 1}}).personalised_reference(one)) ==
           std::vector<std::string>{"ATCGGTTTATC", "ATCTTTTG"}));
    one.reset();
    auto same = make({{0, 0}, null, {1, 1}, {1, 1}}).personalised_reference(one);
    CHECK(same.size() == 2 && same[0].seq == "ATCGCTTTTTG" && same[1].seq == same[0].seq);
    const std::string text = deduped_fasta_text(same, "s personalised reference made by gramtools genotype");
    CHECK(text == ">gramtools_prg_1 s personalised reference made by gramtools genotype\nATCGCTTTTTG\n");
    SegmentTracker to_edge("chr1\t2\nchr2\t9\n"), from_edge("chr1\t6\nchr2\t5\n"), adjacent("chr1\t10\nchr2\t1\n"),
        in_seq("chr1\t7\nchr2\t4\n");
    auto all_null = make({null, null, null, null});
    CHECK((seqs(all_null.personalised_reference(to_edge)) == std::vector<std::string>{"AT", "CGCTTTATC"}));
    CHECK((seqs(all_null.personalised_reference(from_edge)) == std::vector<std::string>{"ATCGCT", "TTATC"}));
    CHECK((seqs(all_null.personalised_reference(adjacent)) == std::vector<std::string>{"ATCGCTTTAT", "C"}));
    in_seq.reset();
    auto refs = all_null.personalised_reference(in_seq);
    CHECK((seqs(refs) == std::vector<std::string>{"ATCGCTT", "TATC"}));
    CHECK(refs[0].id == "chr1" && refs[1].id == "chr2");
    CHECK(throws([&] {
      SegmentTracker t;
      make({{1, 2}, null, {0}, {0, 1}}).personalised_reference(t);  // Alleles_To_Paste.GivenInconsistentPloidy_Throws
    }));
  }
  CASE("Fasta record layout (personalised_reference.cpp:118-137)");
  {
    Fasta f{"id", "d", std::string(120, 'A')};
    CHECK(f.to_string() == ">id d\n" + std::string(60, 'A') + "\n" + std::string(60, 'A'));
    Fasta g{"id", "d", std::string(61, 'C')};
    CHECK(g.to_string() == ">id d\n" + std::string(60, 'C') + "\nC");
  }
  CASE("json_number (nlohmann::json 3.7 float layout)");
  {
    CHECK(json_number(5.0) == "5.0" && json_number(0.0) == "0.0" && json_number(0.5) == "0.5");
    CHECK(json_number(12.5) == "12.5" && json_number(-3.25) == "-3.25" && json_number(100.0) == "100.0");
    CHECK(json_number(0.001) == "0.001" && json_number(0.0001) == "0.0001" && json_number(0.00001) == "1e-05");
    CHECK(json_number(1e15) == "1e+15" || json_number(1e15) == "1000000000000000.0");
    CHECK(json_number(1e16) == "1e+16" && json_number(1.5e-7) == "1.5e-07");
    CHECK(json_number(10 + 4 / 3.) == "11.333333333333334");
    CHECK(json_number(std::nan("")) == "null");
  }
}

int main() {
  test_probabilities();
  test_model_coverages();
  test_model_calls();
  test_allele_extraction();
  test_site_interface();
  test_read_depth();
  test_runner_logic();
  test_segments_and_outputs();
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
