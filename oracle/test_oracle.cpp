// test_oracle.cpp — pins the CPU oracle (TEST INFRASTRUCTURE) against the reference's own
// known-answer tests. Each case cites the reference gtest it transcribes
// (paths relative to /root/reference/libgramtools/tests/). Expected values are copied as DATA;
// no reference source is compiled or linked.
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>

#include "gq_oracle.hpp"

using namespace gqo;

static int g_fail = 0, g_checks = 0, g_cases = 0;
static const char* g_case = "";

template <class T>
static std::string show(const T& v);
template <class T>
static std::string show(const std::vector<T>& v);
template <class A, class B>
static std::string show(const std::pair<A, B>& p);
template <class K, class V>
static std::string show(const std::map<K, V>& m);
template <class K>
static std::string show(const std::set<K>& m);
static std::string show(const uint8_t& v);
static std::string show(const TargetedMarker& t);
static std::string show(const VariantLocus& l) {
  return "(" + std::to_string(l.first) + "," + std::to_string(l.second) + ")";
}
static std::string show(const SearchState& s) {
  std::ostringstream o;
  o << "{[" << s.lo << "," << s.hi << "] T:";
  for (auto& l : s.traversed) o << show(l);
  o << " G:";
  for (auto& l : s.traversing) o << show(l);
  o << "}";
  return o.str();
}
static std::string show(const SearchStates& ss) {
  std::string o;
  for (auto& s : ss) o += show(s) + " ";
  return o;
}
template <class T>
static std::string show(const std::vector<T>& v) {
  std::ostringstream o;
  o << "[";
  for (auto& e : v) o << show(e) << ",";
  o << "]";
  return o.str();
}
template <class A, class B>
static std::string show(const std::pair<A, B>& p) {
  return "(" + show(p.first) + "," + show(p.second) + ")";
}
template <class K, class V>
static std::string show(const std::map<K, V>& m) {
  std::ostringstream o;
  o << "{";
  for (auto& e : m) o << show(e.first) << ":" << show(e.second) << ",";
  o << "}";
  return o.str();
}
template <class K>
static std::string show(const std::set<K>& m) {
  std::ostringstream o;
  o << "{";
  for (auto& e : m) o << show(e) << ",";
  o << "}";
  return o.str();
}
template <class T>
static std::string show(const T& v) {
  std::ostringstream o;
  o << v;
  return o.str();
}
static std::string show(const uint8_t& v) { return std::to_string((int)v); }
static std::string show(const TargetedMarker& t) {
  return "<" + std::to_string(t.id) + "," + std::to_string(t.direct_deletion_allele) + ">";
}

#define CHECK_EQ(a, b)                                                                      \
  do {                                                                                      \
    ++g_checks;                                                                             \
    auto _a = (a);                                                                          \
    decltype(_a) _b = (b);                                                                  \
    if (!(_a == _b)) {                                                                      \
      ++g_fail;                                                                             \
      std::printf("FAIL %s (%s:%d): %s\n   got      %s\n   expected %s\n", g_case, __FILE__, \
                  __LINE__, #a " == " #b, show(_a).c_str(), show(_b).c_str());              \
    }                                                                                       \
  } while (0)
#define CHECK(c)                                                               \
  do {                                                                         \
    ++g_checks;                                                                \
    if (!(c)) {                                                                \
      ++g_fail;                                                                \
      std::printf("FAIL %s (%s:%d): %s\n", g_case, __FILE__, __LINE__, #c);     \
    }                                                                          \
  } while (0)
#define CHECK_THROWS(expr)                                                      \
  do {                                                                          \
    ++g_checks;                                                                 \
    bool _t = false;                                                            \
    try {                                                                       \
      (void)(expr);                                                             \
    } catch (...) {                                                             \
      _t = true;                                                                \
    }                                                                           \
    if (!_t) {                                                                  \
      ++g_fail;                                                                 \
      std::printf("FAIL %s (%s:%d): no throw: %s\n", g_case, __FILE__, __LINE__, #expr); \
    }                                                                           \
  } while (0)

struct Case {
  const char* name;
  std::function<void()> fn;
};
static std::vector<Case>& cases() {
  static std::vector<Case> c;
  return c;
}
struct Reg {
  Reg(const char* n, std::function<void()> f) { cases().push_back({n, f}); }
};
#define TESTCASE(name) \
  static void name();  \
  static Reg reg_##name(#name, name); \
  static void name()

// ---- stand-ins for tests/test_resources/test_resources.hpp:26-65 (prg_setup) -------------
struct Setup {
  Mapper m;
  static Setup numbered(const std::string& prg, uint32_t k = 2) { return Setup(numbered_to_ints(prg), k); }
  static Setup bracketed(const std::string& prg, uint32_t k = 2) { return Setup(bracketed_to_ints(prg), k); }
  Setup(const std::vector<Marker>& prg, uint32_t k) : m(prg, k) {}
  StrandStatus qm(const std::string& read, uint32_t seed = 42) {  // quasimap.hpp:68 default seed 42
    return m.quasimap_read(encode_read(read), seed, false);
  }
  SearchStates search(const std::string& read) { return search_read_backwards(m.info, m.kmers, encode_read(read), m.k); }
  // collect_coverage (test_resources.cpp:9-21)
  std::vector<std::vector<CovCount>> collect(const std::vector<int>& positions) {
    std::vector<std::vector<CovCount>> out;
    for (int p : positions) out.push_back(m.info.graph.nodes[m.info.graph.random_access[p].node].cov);
    return out;
  }
};
using AS = std::vector<std::vector<CovCount>>;
using PB = std::vector<std::vector<std::vector<CovCount>>>;
using GAC = std::vector<GroupedAlleleCounts>;
static PRGInfo info_numbered(const std::string& s) { return build_prg_info(numbered_to_ints(s)); }
static PRGInfo info_bracketed(const std::string& s) { return build_prg_info(bracketed_to_ints(s)); }
static SearchState SS(SA_Index lo, SA_Index hi, VariantSitePath a = {}, VariantSitePath b = {}) {
  return SearchState{lo, hi, a, b};
}

// =========================== genotype/quasimap/test_quasimap.cpp ===========================
TESTCASE(quasimap_revcomp_26) {
  CHECK_EQ(reverse_complement(Sequence{1, 2, 1, 3, 4}), (Sequence{1, 2, 4, 3, 4}));
}
TESTCASE(quasimap_kmers_all_in_read_55) {
  KmerIndex idx{{encode_read("accg"), {}}, {encode_read("ccgt"), {}}};
  CHECK(all_kmers_in_index(idx, encode_read("accgt"), 4));
  CHECK(!all_kmers_in_index(idx, encode_read("tccgt"), 4));
}
TESTCASE(quasimap_cov_67_118) {
  { auto s = Setup::numbered("gct5c6g6t6aG7t8C8CTA"); s.qm("agccta"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 0}, {0, 1}})); }
  { auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta"); s.qm("agtcta"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 0}, {1, 0}})); }
  { auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta"); s.qm("ctgagtcta"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 1, 0}, {1, 0}})); }
  { auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta"); s.qm("tagtcta"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 1}, {1, 0}})); }
}
TESTCASE(quasimap_cov_120_172) {
  { auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta"); s.qm("tgtcta"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 0}, {0, 0}})); }
  { auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta"); s.qm("gctc"); CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 0}, {0, 0}})); }
  { auto s = Setup::numbered("gct5c6g6T6AG7T8c8cta"); s.qm("tagt"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 1}, {1, 0}})); }
  { auto s = Setup::numbered("gct5c6g6t6ag7t8ta8"); s.qm("tagc"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 0}, {0, 0}})); }
}
TESTCASE(quasimap_three_positions_seeds_174) {
  auto s = Setup::numbered("TAG5Tc6g6T6AG7T8c8cta");
  s.qm("tagt", 42);
  CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 1}, {0, 0}}));
  s.qm("tagt", 150);
  CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 2}, {1, 0}}));
}
TESTCASE(quasimap_within_allele_200_238) {
  { auto s = Setup::numbered("gct5cccc6g6t6ag"); s.qm("cccc"); CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 0}})); }
  { auto s = Setup::numbered("ac5t6cagtagtc6ta"); s.qm("gtagt"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 1}})); }
  { auto s = Setup::numbered("ac5t6cagtagttttgtagtc6ta"); s.qm("gtagt"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 1}})); }
}
TESTCASE(quasimap_within_allele_and_outside_240) {
  auto s = Setup::numbered("gtagtac5gtagtact6t6ta");
  s.qm("gtagt", 29);
  CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0}}));
  CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{1, 1, 1, 1, 1, 0, 0, 0}, {0}}}));
}
TESTCASE(quasimap_two_alleles_260_312) {
  {
    auto s = Setup::numbered("tac5gta6gtt6ta");
    s.qm("tacgt");
    CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 1}}));
    CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{1, 1, 0}, {1, 1, 0}}}));
  }
  { auto s = Setup::numbered("c5ccc6agt6ccgt6taa"); s.qm("gttaa"); CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 1, 1}})); }
  {
    auto s = Setup::numbered("ac5gtagtact6t6gggtagt6ta");
    s.qm("gtagt");
    CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 1}}));
    CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{1, 1, 1, 1, 1, 0, 0, 0}, {0}, {0, 0, 1, 1, 1, 1, 1}}}));
  }
}
TESTCASE(quasimap_multiple_reads_314_404) {
  {
    auto s = Setup::numbered("gct5c6g6T6AG7T8c8cta");
    s.qm("tagt"); s.qm("tagt");
    CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 2}, {2, 0}}));
    CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{0}, {0}, {2}}, {{2}, {0}}}));
  }
  {
    auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta");
    s.qm("gagt"); s.qm("tagt"); s.qm("cagt");
    CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 1, 1}, {3, 0}}));
    CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{1}, {1}, {1}}, {{3}, {0}}}));
  }
  {
    auto s = Setup::numbered("gct5c6g6t6ag7t8c8cta");
    s.qm("gagt"); s.qm("tagt"); s.qm("cagc");
    CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 1, 1}, {2, 1}}));
  }
  {
    auto s = Setup::numbered("gcac5t6g6c6ta7t8c8cta");
    s.qm("accta", 200); s.qm("gcact", 200);
    CHECK_EQ(s.m.cov.allele_sum, (AS{{1, 0, 0}, {0, 1}}));
  }
}
TESTCASE(quasimap_kmer_absent_406) {
  auto info = info_numbered("gcgct5c6g6t6agtcct");
  auto idx = index_kmers(info, {encode_read("tagt"), encode_read("agta"), encode_read("gtaa")}, 4);
  CHECK_EQ(search_read_backwards(info, idx, encode_read("tagtaa"), 4).size(), (size_t)0);
}
TESTCASE(quasimap_initially_in_site_420) {
  auto info = info_numbered("gcgct5c6G6t6agtcct");
  SearchStates init{SS(10, 10)};
  auto fin = process_read_char(info, 4, init);
  CHECK_EQ(fin.size(), (size_t)1);
  CHECK_EQ(fin.front().traversed, (VariantSitePath{{5, 1}}));
}
TESTCASE(quasimap_end_in_site_443_480) {
  auto s = Setup::numbered("gcgct5c6g6T6AGTCCt");
  auto st = s.search("tagtcc");
  CHECK_EQ(st, (SearchStates{SS(14, 14, {}, {{5, -1}})}));
  s.qm("tagtcc");
  CHECK_EQ(s.m.cov.allele_sum, (AS{{0, 0, 1}}));
  CHECK_EQ(allele_base_non_nested(s.m.info), (PB{{{0}, {0}, {1}}}));
}
TESTCASE(quasimap_searchstates_482_549) {
  { auto s = Setup::numbered("gcGCT5C6g6t6agtcct"); auto st = s.search("gcgctc"); CHECK_EQ(st.size(), (size_t)1); CHECK_EQ(st.front().traversed, (VariantSitePath{{5, 0}})); }
  { auto s = Setup::numbered("gcgcT5c6G6t6AGtcct"); auto st = s.search("gctgag"); CHECK_EQ(st.size(), (size_t)1); CHECK_EQ(st.front().traversed, (VariantSitePath{{5, 1}})); }
  {
    auto s = Setup::numbered("gct5c6g6t6ag7T8c8CT");
    auto st = s.search("cagtct");
    CHECK_EQ(st.size(), (size_t)1);
    CHECK_EQ(st.front().traversed, (VariantSitePath{{7, 0}}));
    CHECK_EQ(st.front().traversing, (VariantSitePath{{5, -1}}));
  }
  {
    auto s = Setup::numbered("gct5c6g6t6ag7GAG8c8ct");
    auto st = s.search("caggag");
    CHECK_EQ(st.size(), (size_t)1);
    CHECK_EQ(st.front().traversed, (VariantSitePath{{7, 0}}));
    CHECK_EQ(st.front().traversing, (VariantSitePath{{5, -1}}));
  }
}
TESTCASE(quasimap_multistep_556) {
  auto s = Setup::numbered("gct5gC6aC6C6t6Cg", 1);
  auto st = s.m.kmers.at(encode_read("c"));
  CHECK_EQ(st.size(), (size_t)1);
  CHECK_EQ(st.front().hi - st.front().lo + 1, 5u);
  st = process_read_char(s.m.info, 2, st);
  CHECK_EQ(st.size(), (size_t)1);
  CHECK_EQ(st.front().traversing.back().second, -1);
  CHECK_EQ(st.front().hi - st.front().lo + 1, 3u);
}
TESTCASE(quasimap_encapsulated_582_630) {
  { auto s = Setup::numbered("t5c6gCTTAGT6aa"); auto st = s.search("cttagt"); CHECK_EQ(st.size(), (size_t)1); CHECK_EQ(st.front().traversed.front(), (VariantLocus{5, 1})); }
  { auto s = Setup::numbered("t5c6gcttagtacgcttagt6aa"); CHECK_EQ(s.search("cttagt"), (SearchStates{SS(7, 8, {{5, 1}})})); }
  { auto s = Setup::bracketed("a[c,g[ct,t]a]c"); CHECK_EQ(s.search("agtac"), (SearchStates{SS(1, 1, {{7, 1}, {5, 1}})})); }
}
TESTCASE(quasimap_nested_deletion_exit_entry_653) {
  auto s = Setup::bracketed("t[a[c,g][c,g],]t", 1);
  CHECK_EQ(s.search("tt"), (SearchStates{SS(7, 7, {{5, 1}})}));
  CHECK_EQ(s.search("tacct"), (SearchStates{SS(7, 7, {{9, 0}, {7, 0}, {5, 0}})}));
}
TESTCASE(quasimap_nested_double_nesting_683_741) {
  std::vector<int> pos{0, 3, 5, 9, 12, 15, 17};
  {
    auto s = Setup::bracketed("A[[A[CCC,c],t],g]TA");
    s.qm("AACCCTA");
    CHECK_EQ(s.m.cov.grouped, (GAC{{{{0}, 1}}, {{{0}, 1}}, {{{0}, 1}}}));
    CHECK_EQ(s.collect(pos), (AS{{}, {1}, {1, 1, 1}, {0}, {0}, {0}, {}}));
  }
  {
    auto s = Setup::bracketed("A[[A[CCC,c],t],g]TA");
    s.qm("CTA");
    CHECK_EQ(s.m.cov.grouped, (GAC{{{{0}, 1}}, {{{0}, 1}}, {{{0, 1}, 1}}}));
    CHECK_EQ(s.collect(pos), (AS{{}, {0}, {0, 0, 1}, {1}, {0}, {0}, {}}));
  }
}
TESTCASE(quasimap_nested_single_plus_snp_743_832) {
  std::vector<int> pos{0, 2, 4, 7, 9, 11, 13, 17, 19, 21, 23};
  const char* prg = "a[t[tt,t]t,a[at,]a]g[c,g]";
  {
    auto s = Setup::bracketed(prg);
    s.qm("ATTTTGC");
    CHECK_EQ(s.m.cov.grouped, (GAC{{{{0}, 1}}, {{{0}, 1}}, {}, {{{0}, 1}}}));
    CHECK_EQ(s.collect(pos), (AS{{}, {1}, {1, 1}, {0}, {1}, {0}, {0, 0}, {0}, {}, {1}, {0}}));
  }
  {
    auto s = Setup::bracketed(prg);
    s.qm("TT");
    CHECK_EQ(s.m.cov.grouped, (GAC{{{{0}, 1}}, {{{0, 1}, 1}}, {}, {}}));
    CHECK_EQ(s.collect(pos), (AS{{}, {1}, {1, 1}, {1}, {1}, {0}, {0, 0}, {0}, {}, {0}, {0}}));
  }
  {
    auto s = Setup::bracketed(prg);
    s.qm("AAAGG");
    CHECK_EQ(s.m.cov.grouped, (GAC{{{{1}, 1}}, {}, {{{1}, 1}}, {{{1}, 1}}}));
    CHECK_EQ(s.collect(pos), (AS{{}, {0}, {0, 0}, {0}, {0}, {1}, {0, 0}, {1}, {}, {0}, {1}}));
  }
}

// =========================== genotype/quasimap/search/test_vBWT_jump.cpp ===================
TESTCASE(vbwt_marker_search_55_107) {
  auto info = info_numbered("gcgct5c6g6a6agtcct");
  CHECK_EQ(left_markers_search(info, SS(1, 2)), (std::vector<VariantLocus>{{6, -1}, {5, 2}}));
  CHECK_EQ(search_state_vbwt_jumps(info, SS(1, 2)).size(), (size_t)2);
  auto r = left_markers_search(info, SS(1, 1));
  CHECK(is_allele_marker(r[0].first));
  r = left_markers_search(info, SS(7, 7));
  CHECK(is_site_marker(r[0].first));
  CHECK_EQ(left_markers_search(info, SS(8, 11)), (std::vector<VariantLocus>{{5, 1}}));
}
TESTCASE(vbwt_jumps_simple_109_148) {
  auto info = info_numbered("gcgct5c6g6a6agtcct");
  auto ms = search_state_vbwt_jumps(info, SS(8, 11));
  CHECK_EQ(ms.size(), (size_t)1);
  CHECK_EQ(ms.front().lo, 15u);
  CHECK_EQ(ms.front().hi, 15u);
  ms = search_state_vbwt_jumps(info, SS(3, 7));
  CHECK_EQ(ms.size(), (size_t)1);
  CHECK_EQ(ms.front().lo, 15u);
  CHECK_EQ(ms.front().hi, 15u);
}
TESTCASE(vbwt_marker_sa_intervals_150_193) {
  CHECK_EQ(marker_sa_interval(info_numbered("gcgct5c6g6a6agtcct"), 6), (std::pair<SA_Index, SA_Index>{16, 18}));
  CHECK_EQ(marker_sa_interval(info_numbered("aca5g6t6catt"), 6), (std::pair<SA_Index, SA_Index>{11, 12}));
  CHECK_EQ(marker_sa_interval(info_numbered("7g8c8g9t10a10"), 8), (std::pair<SA_Index, SA_Index>{7, 8}));
}
TESTCASE(vbwt_jump_entry_exit_220_270) {
  auto info = info_numbered("gcgct5c6g6t6agtcct");
  CHECK_EQ(search_state_vbwt_jumps(info, SS(1, 1)), (SearchStates{SS(16, 18, {}, {{5, -1}})}));
  CHECK_EQ(search_state_vbwt_jumps(info, SS(7, 10)), (SearchStates{SS(15, 15, {{5, 1}})}));
  CHECK_EQ(search_state_vbwt_jumps(info, SS(2, 6)), (SearchStates{SS(15, 15, {{5, 0}})}));
}
TESTCASE(vbwt_jump_nested_292_405) {
  {
    auto info = info_bracketed("[AC,[C,G]]T");
    CHECK_EQ(search_state_vbwt_jumps(info, SS(3, 3)), (SearchStates{SS(6, 6, {{7, 0}, {5, 1}})}));
    CHECK_EQ(search_state_vbwt_jumps(info, SS(5, 5)),
             (SearchStates{SS(7, 8, {}, {{5, -1}}), SS(10, 11, {}, {{5, -1}, {7, -1}})}));
  }
  {
    auto info = info_bracketed("[C,G][C,G]");
    CHECK_EQ(search_state_vbwt_jumps(info, SS(2, 2)), (SearchStates{SS(6, 7, {{7, 0}}, {{5, -1}})}));
  }
  {
    auto info = info_bracketed("A[C,,G]T");
    CHECK_EQ(search_state_vbwt_jumps(info, SS(4, 4)),
             (SearchStates{SS(6, 8, {}, {{5, -1}}), SS(5, 5, {{5, 1}})}));
  }
}

#include "test_oracle_more.inc"

int main(int argc, char** argv) {
  std::string filter = argc > 1 ? argv[1] : "";
  for (auto& c : cases()) {
    if (!filter.empty() && std::string(c.name).find(filter) == std::string::npos) continue;
    g_case = c.name;
    ++g_cases;
    int before = g_fail;
    try {
      c.fn();
    } catch (const std::exception& e) {
      ++g_fail;
      std::printf("FAIL %s: exception %s\n", c.name, e.what());
    }
    if (g_fail == before && argc > 2) std::printf("ok   %s\n", c.name);
  }
  std::printf("%d cases, %d checks, %d failures\n", g_cases, g_checks, g_fail);
  return g_fail ? 1 : 0;
}
