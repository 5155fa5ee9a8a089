// gq_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY). See gq_oracle.hpp for scope.
// Restates libgramtools' quasimap path; each block cites the reference file:line.
#include "gq_oracle.hpp"

#include <cassert>
#include <stack>

namespace gqo {

// =====================================================================================
// Encoding (src/common/utils.cpp:13-81, src/prg/linearised_prg.cpp:132-265)
// =====================================================================================
static int encode_char_dna(char c) {  // utils.cpp:13-47
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    default: return 0;
  }
}

Sequence encode_read(const std::string& s) {  // utils.cpp:72-81
  Sequence out;
  out.reserve(s.size());
  for (char c : s) {
    int b = encode_char_dna(c);
    if (b == 0) return Sequence{};
    out.push_back((Base)b);
  }
  return out;
}

std::vector<Marker> bracketed_to_ints(const std::string& s) {  // linearised_prg.cpp:166-215
  std::stack<int> marker_stack;
  int max_var_marker = 3;
  std::vector<Marker> out;
  out.reserve(s.size());
  for (char c : s) {
    switch (c) {
      case '[':
        max_var_marker += 2;
        marker_stack.push(max_var_marker);
        out.push_back(max_var_marker);
        break;
      case ']':
        if (marker_stack.empty()) throw std::runtime_error("unbalanced ]");
        out.push_back(marker_stack.top() + 1);
        marker_stack.pop();
        break;
      case ',':
        if (marker_stack.empty()) throw std::runtime_error("stray ,");
        out.push_back(marker_stack.top() + 1);
        break;
      default: {
        int b = encode_char_dna(c);
        if (b == 0) throw std::runtime_error("not a nucleotide char");
        out.push_back(b);
      }
    }
  }
  return out;
}

std::vector<Marker> numbered_to_ints(const std::string& s) {  // linearised_prg.cpp:241-265
  std::vector<Marker> out;
  uint64_t marker = 0;
  bool in_marker = false;
  for (char c : s) {
    int b = encode_char_dna(c);
    if (b) {
      if (in_marker) out.push_back((Marker)marker);
      in_marker = false;
      marker = 0;
      out.push_back(b);
    } else {
      marker = marker * 10 + (uint32_t)(c - '0');
      in_marker = true;
    }
  }
  if (in_marker) out.push_back((Marker)marker);
  return out;
}

std::string ints_to_bracketed(const std::vector<Marker>& v) {  // linearised_prg.cpp:132-164
  std::string out(v.size(), '0');
  std::unordered_map<Marker, int> last;
  for (std::size_t pos = 0; pos < v.size(); ++pos) {
    Marker s = v[pos];
    if (s > 4) {
      if (s % 2 == 1) out[pos] = '[';
      else {
        out[pos] = ',';
        last[s] = (int)pos;
      }
      continue;
    }
    out[pos] = "?ACGT"[s];
  }
  for (auto& e : last) out[e.second] = ']';
  return out;
}

// =====================================================================================
// PRG_String::map_ends_and_check_for_duplicates (linearised_prg.cpp:52-80)
// =====================================================================================
void build_end_positions(PRGInfo& info) {
  std::set<Marker> seen;
  for (std::size_t pos = 0; pos < info.prg.size(); ++pos) {
    Marker m = info.prg[pos];
    if (m <= 4) continue;
    if (is_site_marker(m)) {
      if (seen.count(m))
        throw std::runtime_error("PRG consistency error: site marker " + std::to_string(m) +
                                 " used for two different sites");
      seen.insert(m);
    } else
      info.last_allele_positions[m] = (int)pos;
  }
}

// =====================================================================================
// FM index: stands in for sdsl::construct(fm_index, prg, cfg, 4)
// (make_data_structures.cpp:9-33). Text = prg ‖ 0; integer alphabet; full SA.
// Plain comparison sort (the SA of a sentinel-terminated string is unique).
// =====================================================================================
void build_fm_index(const std::vector<Marker>& prg, FMIndex& fm) {
  const std::size_t n = prg.size() + 1;
  std::vector<uint32_t> text(prg.begin(), prg.end());
  text.push_back(0);
  fm.sa.resize(n);
  std::iota(fm.sa.begin(), fm.sa.end(), 0u);
  const uint32_t* t = text.data();
  std::sort(fm.sa.begin(), fm.sa.end(), [t, n](uint32_t a, uint32_t b) {
    if (a == b) return false;
    while (true) {  // sentinel 0 is unique & smallest: terminates before running off
      uint32_t ca = t[a], cb = t[b];
      if (ca != cb) return ca < cb;
      ++a;
      ++b;
    }
  });
  fm.bwt.resize(n);
  for (std::size_t i = 0; i < n; ++i) fm.bwt[i] = fm.sa[i] ? text[fm.sa[i] - 1] : text[n - 1];
  std::map<uint32_t, uint64_t> hist;
  for (auto c : text) hist[c]++;
  fm.alphabet.clear();
  fm.C.clear();
  fm.char2comp.clear();
  uint64_t acc = 0;
  for (auto& e : hist) {
    fm.char2comp[e.first] = (uint32_t)fm.alphabet.size();
    fm.alphabet.push_back(e.first);
    fm.C.push_back(acc);
    acc += e.second;
  }
  fm.C.push_back(acc);
}

void RankedMask::build(const std::vector<uint8_t>& bits) {
  nbits = bits.size();
  std::size_t nw = (bits.size() + 64) / 64;
  words.assign(nw, 0);
  cum.assign(nw + 1, 0);
  for (std::size_t i = 0; i < bits.size(); ++i)
    if (bits[i]) words[i >> 6] |= 1ULL << (i & 63);
  for (std::size_t w = 0; w < nw; ++w) cum[w + 1] = cum[w] + __builtin_popcountll(words[w]);
}

void build_masks(PRGInfo& info) {  // make_data_structures.cpp:78-95, :158-163
  const auto& bwt = info.fm.bwt;
  std::vector<uint8_t> bits(bwt.size());
  for (uint32_t b = 1; b <= 4; ++b) {
    for (std::size_t i = 0; i < bwt.size(); ++i) bits[i] = bwt[i] == b;
    info.mask[b - 1].build(bits);
  }
  for (std::size_t i = 0; i < bwt.size(); ++i) bits[i] = bwt[i] > 4;
  info.markers.build(bits);
}

// =====================================================================================
// cov_Graph_Builder (coverage_graph.cpp:82-379)
// =====================================================================================
namespace {
enum class MT { sequence, site_entry, allele_end, site_end };

struct GraphBuilder {
  PRGInfo& info;
  CovGraph& g;
  const std::vector<Marker>& prg;
  int backWire = -1, cur_Node = -1;
  std::size_t cur_pos = 0;
  bool first_allele = false;
  VariantLocus cur_Locus{0, ALLELE_UNKNOWN};

  explicit GraphBuilder(PRGInfo& i) : info(i), g(i.graph), prg(i.prg) {}

  int new_node(const std::string& seq, std::size_t pos, Marker site, AlleleId allele) {
    Node n;
    n.seq = seq;
    n.pos = pos;
    n.site = site;
    n.allele = allele;
    if (n.in_bubble()) n.cov.assign(seq.size(), 0);
    g.nodes.push_back(std::move(n));
    return (int)g.nodes.size() - 1;
  }

  MT marker_type(std::size_t pos) const {  // coverage_graph.cpp:146-164
    Marker m = prg[pos];
    if (m <= 4) return MT::sequence;
    if (m % 2 == 1) return MT::site_entry;
    auto end_pos = (std::size_t)info.last_allele_positions.at(m);
    assert(pos <= end_pos);
    if (pos < end_pos) return MT::allele_end;
    return MT::site_end;
  }

  void wire(int target) {  // :260-266
    if (g.nodes[cur_Node].has_sequence()) {
      g.nodes[backWire].next.push_back(cur_Node);
      g.nodes[cur_Node].next.push_back(target);
    } else
      g.nodes[backWire].next.push_back(target);
  }

  void make_root() {  // :98-104
    cur_pos = (std::size_t)-1;
    g.root = new_node("", cur_pos, 0, ALLELE_UNKNOWN);
    backWire = g.root;
    cur_pos++;
    cur_Node = new_node("", cur_pos, 0, ALLELE_UNKNOWN);
  }
  void make_sink() {  // :106-111
    int sink = new_node("", cur_pos + 1, 0, ALLELE_UNKNOWN);
    wire(sink);
    cur_Node = backWire = -1;
  }
  void add_sequence(std::size_t prg_pos, Marker m) {  // :166-172 + coverage_Node::add_sequence :33-38
    Node& n = g.nodes[cur_Node];
    if (n.seq.empty()) n.prg_start = (int64_t)prg_pos;
    n.seq.push_back("?ACGT"[m]);
    if (n.in_bubble()) n.cov.push_back(0);
    cur_pos++;
  }
  void enter_site(Marker m) {  // :174-197
    int site_entry = new_node("", cur_pos, m, ALLELE_UNKNOWN);
    g.nodes[site_entry].boundary = true;
    wire(site_entry);
    cur_Node = new_node("", cur_pos, m, FIRST_ALLELE);
    first_allele = true;
    backWire = site_entry;
    int site_exit = new_node("", cur_pos, m, ALLELE_UNKNOWN);
    g.nodes[site_exit].boundary = true;
    g.bubbles.emplace_back(site_entry, site_exit);
    g.bubble_starts[m] = site_entry;
    g.bubble_ends[m] = site_exit;
    if (cur_Locus.first != 0) g.par_map.insert({m, cur_Locus});
    cur_Locus = {m, FIRST_ALLELE};
  }
  int reach_allele_end(Marker m) {  // :240-258
    Marker site_ID = m - 1;
    if (cur_Locus.first != site_ID) throw std::runtime_error("PRG consistency error: allele marker outside its site");
    int site_exit = g.bubble_ends.at(site_ID);
    wire(site_exit);
    if (first_allele) {
      g.nodes[site_exit].pos = cur_pos;
      first_allele = false;
    }
    return site_exit;
  }
  void end_allele(Marker m) {  // :199-213
    Marker site_ID = m - 1;
    reach_allele_end(m);
    AlleleId& allele_ID = cur_Locus.second;
    int site_entry = g.bubble_starts.at(site_ID);
    backWire = site_entry;
    cur_pos = g.nodes[site_entry].pos;
    allele_ID++;
    cur_Node = new_node("", cur_pos, site_ID, allele_ID);
  }
  void exit_site(Marker m) {  // :215-238
    Marker site_ID = m - 1;
    int site_exit = reach_allele_end(m);
    if (cur_Locus.second == FIRST_ALLELE)
      throw std::runtime_error("Site numbered " + std::to_string(m) + " has only one allele");
    auto it = g.par_map.find(site_ID);
    if (it != g.par_map.end()) {
      cur_Locus = it->second;
      if (cur_Locus.second == FIRST_ALLELE) first_allele = true;
    } else
      cur_Locus = {0, ALLELE_UNKNOWN};
    backWire = site_exit;
    cur_pos = g.nodes[site_exit].pos;
    cur_Node = new_node("", cur_pos, cur_Locus.first, cur_Locus.second);
  }

  void setup_random_access(std::size_t pos) {  // :131-144
    MT t = marker_type(pos);
    int target = t == MT::sequence ? cur_Node : backWire;
    std::size_t seq_size = g.nodes[target].seq.size();
    NodeAccess a;
    a.node = target;
    a.offset = seq_size <= 1 ? 0 : seq_size - 1;
    g.random_access[pos] = a;
  }

  void add_exit_target(Marker cur_m, TargetedMarker t) { g.target_map[cur_m].push_back(t); }  // :372-379

  void map_targets() {  // :268-311
    MT prev_t = MT::sequence;
    Marker prev_m = 0;
    // NB: the reference declares `Marker cur_allele_ID` (unsigned) and stores -1 in it; the value
    // round-trips through AlleleId (int32) unchanged.
    AlleleId cur_allele_ID = ALLELE_UNKNOWN;
    for (std::size_t pos = 0; pos < prg.size(); ++pos) {
      Marker cur_m = prg[pos];
      MT cur_t = marker_type(pos);
      switch (cur_t) {
        case MT::sequence:
          if (prev_t != MT::sequence) g.random_access[pos].target = {prev_m, cur_allele_ID};
          break;
        case MT::site_entry:
          cur_allele_ID = FIRST_ALLELE;
          if (prev_t != MT::sequence) {  // make_site_entry_target :313-328
            Marker target = prev_m;
            if (prev_t == MT::allele_end) target -= 1;
            g.target_map.insert({cur_m, {TargetedMarker{target, ALLELE_UNKNOWN}}});
          }
          break;
        case MT::site_end:
          if (prev_t != MT::sequence) {  // make_site_exit_target :330-350
            Marker target = prev_m;
            AlleleId dda = ALLELE_UNKNOWN;
            if (prev_t == MT::site_entry)
              throw std::runtime_error("PRG consistency error: site number " + std::to_string(cur_m) + " is empty");
            if (prev_t == MT::allele_end) {
              target -= 1;
              dda = cur_allele_ID;
            }
            add_exit_target(cur_m, {target, dda});
          }
          {
            auto it = g.par_map.find(cur_m - 1);
            cur_allele_ID = it != g.par_map.end() ? it->second.second : ALLELE_UNKNOWN;
          }
          break;
        case MT::allele_end:
          if (prev_t != MT::sequence) {  // make_allele_end_target :352-370
            Marker target = prev_m;
            AlleleId dda = cur_allele_ID;
            if (prev_t == MT::site_end) dda = ALLELE_UNKNOWN;
            else if (prev_t == MT::allele_end) target -= 1;
            add_exit_target(cur_m, {target, dda});
          }
          cur_allele_ID++;
          break;
      }
      prev_m = cur_m;
      prev_t = cur_t;
    }
  }

  void run() {  // :82-96
    g.random_access.assign(prg.size(), NodeAccess{});
    make_root();
    cur_Locus = {0, ALLELE_UNKNOWN};
    for (std::size_t i = 0; i < prg.size(); ++i) {
      Marker m = prg[i];
      switch (marker_type(i)) {  // process_marker :113-129
        case MT::sequence: add_sequence(i, m); break;
        case MT::site_entry: enter_site(m); break;
        case MT::allele_end: end_allele(m); break;
        case MT::site_end: exit_site(m); break;
      }
      setup_random_access(i);
    }
    make_sink();
    map_targets();
    g.is_nested = !g.par_map.empty();
  }
};
}  // namespace

void build_cov_graph(PRGInfo& info) {
  GraphBuilder b(info);
  b.run();
}

PRGInfo build_prg_info(const std::vector<Marker>& prg) {  // submod_resources.cpp:21-62
  PRGInfo info;
  info.prg = prg;
  for (auto m : prg)
    if (m < 1) throw std::runtime_error("PRG symbols must be >= 1");
  build_end_positions(info);
  build_fm_index(info.prg, info.fm);
  build_cov_graph(info);
  build_masks(info);
  info.num_sites = info.graph.bubbles.size();
  return info;
}

// =====================================================================================
// Search: BWT_search.cpp, vBWT_jump.cpp, encapsulated_search.cpp
// =====================================================================================
uint64_t dna_bwt_rank(const PRGInfo& info, uint64_t upper, Marker base) {  // BWT_search.cpp:8-22
  if (base >= 1 && base <= 4) return info.mask[base - 1].rank(upper);
  return 0;
}

std::pair<SA_Index, SA_Index> marker_sa_interval(const PRGInfo& info, Marker m) {  // vBWT_jump.cpp:3-21
  auto rank = info.fm.char2comp.at(m);
  SA_Index start = (SA_Index)info.fm.C[rank];
  SA_Index end;
  if (rank < info.fm.sigma() - 1) end = (SA_Index)(info.fm.C[rank + 1] - 1);
  else end = (SA_Index)(info.fm.size() - 1);
  return {start, end};
}

static SearchState entering_site_search_state(const PRGInfo& info, Marker allele_marker,
                                              const SearchState& cur) {  // vBWT_jump.cpp:29-44
  auto iv = marker_sa_interval(info, allele_marker);
  SearchState s = cur;
  s.lo = iv.first;
  s.hi = iv.second;
  s.traversing.push_back({allele_marker - 1, ALLELE_UNKNOWN});
  return s;
}

static void update_variant_site_path(SearchState& s, AlleleId allele_id, Marker site_ID) {  // :51-69
  if (s.traversing.empty()) {
    s.traversed.push_back({site_ID, allele_id});
  } else {
    auto existing = s.traversing.back();
    if (existing.first != site_ID || existing.second != ALLELE_UNKNOWN)
      throw std::logic_error("leaving a site that is not the innermost entered one");
    existing.second = allele_id;
    s.traversed.push_back(existing);
    s.traversing.pop_back();
  }
}

static SearchState exiting_site_search_state(const PRGInfo& info, const VariantLocus& locus,
                                             const SearchState& cur) {  // :76-92
  SearchState s = cur;
  update_variant_site_path(s, locus.second, locus.first);
  auto rank = info.fm.char2comp.at(locus.first);
  SA_Index idx = (SA_Index)info.fm.C[rank];
  s.lo = s.hi = idx;
  return s;
}

std::vector<VariantLocus> left_markers_search(const PRGInfo& info, const SearchState& s, Events* ev) {  // :94-117
  std::vector<VariantLocus> out;
  if (ev) ev->w_marker += (s.hi >> 8) - (s.lo >> 8) + 1;
  for (int64_t index = s.lo; index <= (int64_t)s.hi; ++index) {
    if (!info.markers.get(index)) continue;
    auto prg_index = info.fm.sa[index];
    if (ev) ev->q_sa++, ev->q_node++;
    VariantLocus target = info.graph.random_access[prg_index].target;
    if (is_allele_marker(target.first)) {
      if (info.last_allele_positions.at(target.first) != (int)prg_index - 1) target.first--;
    }
    out.push_back(target);
  }
  return out;
}

namespace {
struct LocusAndState {
  VariantLocus locus;
  SearchState state;
  bool commit_me = false;
};
}  // namespace

static LocusAndState extend_targets_site_exit(const PRGInfo& info, const VariantLocus& target_locus,
                                              const SearchState& state) {  // :185-228
  Marker site_marker = target_locus.first;
  bool commit_me = true;
  const auto& target_map = info.graph.target_map;
  SearchState ns = exiting_site_search_state(info, target_locus, state);
  VariantLocus next_target{0, 0};
  while (target_map.find(site_marker) != target_map.end()) {
    const auto& tms = target_map.at(site_marker);
    assert(tms.size() == 1);
    Marker next_site_marker = tms.back().id;
    if (is_allele_marker(next_site_marker)) {  // exit followed by an entry
      next_target = {next_site_marker, 0};
      commit_me = false;
      break;
    } else {  // double exit
      auto parent = info.graph.par_map.at(site_marker);
      assert(parent.first == next_site_marker);
      ns = exiting_site_search_state(info, {next_site_marker, parent.second}, ns);
      site_marker = next_site_marker;
    }
  }
  return {next_target, ns, commit_me};
}

static std::vector<LocusAndState> extend_targets_site_entry(const PRGInfo& info, const VariantLocus& target_locus,
                                                            const SearchState& state) {  // :230-265
  std::vector<LocusAndState> ext;
  Marker variant_marker = target_locus.first;
  SearchState ns = entering_site_search_state(info, variant_marker, state);
  ext.push_back({{0, 0}, ns, true});
  auto it = info.graph.target_map.find(variant_marker);
  if (it == info.graph.target_map.end()) return ext;
  for (const auto& mt : it->second) {
    if (is_site_marker(mt.id)) {  // direct deletion
      assert(mt.direct_deletion_allele != ALLELE_UNKNOWN);
      ext.push_back({{mt.id, mt.direct_deletion_allele}, ns, false});
    } else {  // double entry
      ext.push_back({{mt.id, ALLELE_UNKNOWN}, ns, false});
    }
  }
  return ext;
}

SearchStates search_state_vbwt_jumps(const PRGInfo& info, const SearchState& cur, Events* ev) {  // :134-183
  auto marker_targets = left_markers_search(info, cur, ev);
  if (marker_targets.empty()) return {};
  SearchStates out;
  std::vector<LocusAndState> to_process;
  for (auto& t : marker_targets) to_process.push_back({t, cur, false});
  while (!to_process.empty()) {
    LocusAndState item = to_process.back();
    to_process.pop_back();
    std::vector<LocusAndState> ext;
    if (is_site_marker(item.locus.first)) ext = {extend_targets_site_exit(info, item.locus, item.state)};
    else ext = extend_targets_site_entry(info, item.locus, item.state);
    for (auto& nt : ext) {
      if (nt.commit_me) out.push_back(nt.state);
      if (nt.locus.first != 0) to_process.push_back(nt);
    }
  }
  return out;
}

void process_markers_search_states(const PRGInfo& info, SearchStates& states, Events* ev) {  // :119-132
  SearchStates all;
  for (const auto& s : states) {
    auto ms = search_state_vbwt_jumps(info, s, ev);
    if (!ms.empty()) all.splice(all.end(), ms);
  }
  states.splice(states.end(), all);
}

SearchStates search_base_backwards(const PRGInfo& info, Base b, const SearchStates& states, Events* ev) {  // BWT_search.cpp:78-94
  // sdsl's int alphabet maps a symbol absent from the text to comp 0 (C[0] = 0); its rank is always 0,
  // so every state then fails the validity test below.
  auto cit = info.fm.char2comp.find(b);
  SA_Index first = (SA_Index)info.fm.C[cit == info.fm.char2comp.end() ? 0 : cit->second];
  SearchStates out;
  for (const auto& s : states) {
    // base_next_sa_interval :45-76
    SA_Index start_off = s.lo <= 0 ? 0 : (SA_Index)dna_bwt_rank(info, s.lo, b);
    SA_Index end_off = (SA_Index)dna_bwt_rank(info, (uint64_t)s.hi + 1, b);
    if (ev) ev->q_rank += 2;
    SA_Index nlo = first + start_off;
    SA_Index nhi = first + end_off - 1;
    if ((SA_Index)(nlo - 1) == nhi) continue;  // :32-37
    SearchState ns = s;
    ns.lo = nlo;
    ns.hi = nhi;
    out.push_back(std::move(ns));
  }
  return out;
}

SearchStates process_read_char(const PRGInfo& info, Base b, SearchStates& states, Events* ev) {  // quasimap.cpp:258-268
  process_markers_search_states(info, states, ev);
  return search_base_backwards(info, b, states, ev);
}

static SearchStates encapsulated_state(const PRGInfo& info, const SearchState& s, Events* ev) {  // encapsulated_search.cpp:30-88
  SearchStates out;
  SearchState cache;
  bool cache_empty = true;
  auto flush = [&]() {
    if (cache_empty) return;
    out.push_back(cache);
    cache_empty = true;
  };
  for (uint64_t i = s.lo; i <= s.hi; ++i) {
    auto prg_index = info.fm.sa[i];
    if (ev) ev->q_sa++, ev->q_node++;
    const Node& node = info.graph.nodes[info.graph.random_access[prg_index].node];
    Marker site = node.site;
    AlleleId allele = node.allele;
    if (site == 0) {
      flush();
      out.push_back(SearchState{(SA_Index)i, (SA_Index)i, {}, {}});
      continue;
    }
    VariantSitePath path{{site, allele}};
    if (cache_empty) {
      cache = SearchState{(SA_Index)i, (SA_Index)i, path, {}};
      cache_empty = false;
      continue;
    }
    if (path == cache.traversed) {
      cache.hi = (SA_Index)i;
      continue;
    }
    flush();
    cache = SearchState{(SA_Index)i, (SA_Index)i, path, {}};
    cache_empty = false;
  }
  flush();
  return out;
}

SearchStates encapsulated_states(const PRGInfo& info, const SearchStates& states, Events* ev) {  // :90-107
  SearchStates out;
  for (const auto& s : states) {
    if (s.has_path()) {
      out.push_back(s);
      continue;
    }
    auto split = encapsulated_state(info, s, ev);
    for (auto& x : split) out.push_back(x);
  }
  return out;
}

bool all_kmers_in_index(const KmerIndex& idx, const Sequence& read, uint32_t k) {  // quasimap.cpp:212-225
  for (std::size_t off = 0; off + k <= read.size(); ++off) {
    Sequence kmer(read.begin() + off, read.begin() + off + k);
    if (idx.find(kmer) == idx.end()) return false;
  }
  return true;
}

SearchStates search_read_backwards(const PRGInfo& info, const KmerIndex& idx, const Sequence& read, uint32_t k,
                                   Events* ev) {  // quasimap.cpp:227-256
  Sequence kmer(read.end() - k, read.end());
  auto it = idx.find(kmer);
  if (it == idx.end()) return {};
  SearchStates states = it->second;
  auto rit = read.rbegin();
  std::advance(rit, k);
  for (; rit != read.rend(); ++rit) {
    if (ev) ev->bases++;
    states = process_read_char(info, *rit, states, ev);
    if (states.empty()) break;
  }
  return encapsulated_states(info, states, ev);
}

Sequence reverse_complement(const Sequence& read) {  // quasimap.cpp:273-298
  Sequence out;
  out.reserve(read.size());
  for (auto it = read.rbegin(); it != read.rend(); ++it) {
    Base b = *it;
    out.push_back(b >= 1 && b <= 4 ? (Base)(5 - b) : (Base)0);
  }
  return out;
}

// =====================================================================================
// k-mer index (src/build/kmer_index/kmers.cpp, build.cpp)
// =====================================================================================
std::vector<Sequence> all_kmers_ordered(uint32_t k) {  // kmers.cpp:23-36,76-96
  // generate_all_kmers enumerates patterns in counting order (last position fastest) into an
  // insertion-ordered set, then each is reversed: consecutive kmers share the longest suffix.
  std::vector<Sequence> out;
  Sequence cur(k, 1);
  while (true) {
    Sequence rev(cur.rbegin(), cur.rend());
    out.push_back(rev);
    int64_t i = (int64_t)k - 1;
    while (i >= 0 && cur[i] == 4) --i;
    if (i < 0) break;
    cur[i]++;
    for (uint64_t j = i + 1; j < k; ++j) cur[j] = 1;
  }
  return out;
}

std::vector<Sequence> prefix_diffs(const std::vector<Sequence>& kmers) {  // kmers.cpp:38-74
  std::vector<Sequence> out;
  Sequence last;
  for (const auto& kmer : kmers) {
    if (last.empty()) {
      last = kmer;
      out.push_back(kmer);
      continue;
    }
    bool found = false;
    std::list<Base> diff;
    for (int64_t i = (int64_t)last.size() - 1; i >= 0; --i) {
      if (kmer[i] != last[i]) found = true;
      if (found) diff.push_front(kmer[i]);
    }
    last = kmer;
    out.emplace_back(diff.begin(), diff.end());
  }
  return out;
}

namespace {
struct CacheElement {
  SearchStates states;
  Base base = 0;
};
}  // namespace

KmerIndex index_kmers(const PRGInfo& info, const std::vector<Sequence>& diffs, uint32_t k) {  // build.cpp:18-131
  KmerIndex index;
  std::vector<CacheElement> cache;
  Sequence full_kmer;
  for (const auto& diff : diffs) {
    // update_full_kmer :90-99
    if (diff.size() == k) full_kmer = diff;
    else
      for (std::size_t i = 0; i < diff.size(); ++i) full_kmer[i] = diff[i];
    // build_kmer_cache :55-88
    auto it = diff.rbegin();
    if (diff.size() == k) {
      cache.clear();
      SearchStates init{SearchState{0, (SA_Index)(info.fm.size() - 1), {}, {}}};  // :37-47
      cache.push_back({search_base_backwards(info, *it, init), *it});
      ++it;
    } else
      cache.resize(k - diff.size());
    for (; it != diff.rend(); ++it) {
      SearchStates ns = cache.back().states;  // get_next_cache_element :18-29
      process_markers_search_states(info, ns);
      ns = search_base_backwards(info, *it, ns);
      cache.push_back({std::move(ns), *it});
    }
    if (!cache.back().states.empty()) index[full_kmer] = cache.back().states;
  }
  return index;
}

KmerIndex build_kmer_index(const PRGInfo& info, uint32_t k) {  // build.cpp:138-148
  return index_kmers(info, prefix_diffs(all_kmers_ordered(k)), k);
}

// =====================================================================================
// Coverage: coverage_common.cpp, allele_sum.cpp, grouped_allele_counts.cpp, allele_base.cpp
// =====================================================================================
uint32_t rng_generate(std::mt19937& g, uint32_t lo, uint32_t hi) {  // random.cpp:16-19
  std::uniform_int_distribution<uint32_t> range(lo, hi);
  return range(g);
}

LocusFinder::LocusFinder(const PRGInfo& info, const SearchState& s, Events* ev) {  // coverage_common.cpp:10-83
  {  // check_site_uniqueness :17-32
    std::set<Marker> sites;
    auto chk = [&](const VariantSitePath& p) {
      for (auto& e : p) {
        if (sites.count(e.first))
          throw std::logic_error("ERROR: A site cannot have been traversed more than once by a read");
        sites.insert(e.first);
      }
    };
    chk(s.traversed);
    chk(s.traversing);
  }
  if (!s.traversing.empty()) {  // assign_traversing_loci :52-74
    Marker parent_seed = s.traversing.back().first;
    VariantLocus new_locus;
    for (int64_t i = s.lo; i <= (int64_t)s.hi; ++i) {
      auto prg_pos = info.fm.sa[i];
      if (ev) ev->q_sa++, ev->q_node++;
      AlleleId allele = info.graph.nodes[info.graph.random_access[prg_pos].node].allele;
      new_locus = {parent_seed, allele};
      unique_loci.insert(new_locus);
    }
    assign_nested(info, new_locus);
  }
  for (const auto& l : s.traversed) assign_nested(info, l);  // :76-83
}

void LocusFinder::assign_nested(const PRGInfo& info, VariantLocus cur) {  // :34-50
  const auto& par_map = info.graph.par_map;
  while (true) {
    if (used_sites.count(cur.first)) break;
    used_sites.insert(cur.first);
    unique_loci.insert(cur);
    auto it = par_map.find(cur.first);
    if (it == par_map.end()) {
      base_sites.insert(cur.first);
      break;
    }
    cur = it->second;
  }
}

UniqueSitePaths equivalence_classes(const PRGInfo& info, const SearchStates& states, Events* ev) {  // :116-133
  UniqueSitePaths usps;
  for (const auto& s : states) {
    if (!s.has_path()) continue;
    LocusFinder l(info, s, ev);
    auto& ci = usps[l.base_sites];
    for (auto& locus : l.unique_loci) ci.second.insert(locus);
    ci.first.push_back(s);
  }
  return usps;
}

uint32_t count_nonvar(const SearchStates& states) {  // :137-148
  uint32_t c = 0;
  for (auto& s : states)
    if (!s.has_path()) c += s.hi - s.lo + 1;
  return c;
}

Selected select_mapping(const PRGInfo& info, const SearchStates& states, uint32_t seed,
                        std::optional<uint32_t> mock_rand, Events* ev) {  // :85-114, :166-177
  Selected sel;
  std::mt19937 gen;
  gen.seed(seed);  // RandomInclusiveInt ctor random.cpp:4-14
  auto usps = equivalence_classes(info, states, ev);
  if (usps.empty()) return sel;
  uint32_t nonvar = count_nonvar(states);
  uint32_t total = nonvar + (uint32_t)usps.size();
  uint32_t pick = mock_rand ? *mock_rand : rng_generate(gen, 1, total);
  if (pick <= nonvar) return sel;
  auto it = usps.begin();
  std::advance(it, pick - nonvar - 1);
  sel.states = it->second.first;
  sel.loci = it->second.second;
  return sel;
}

Traverser::Traverser(const PRGInfo& i, const NodeAccess& start, const VariantSitePath& t, std::size_t read_size)
    : info(&i), cur(start.node), bases_remaining(read_size), traversed(t), first_node(true), end_pos(0) {  // :137-148
  traversed_index = (uint32_t)traversed.size();
  start_pos = (uint32_t)start.offset;
}

std::optional<int> Traverser::next_node() {  // :150-162
  if (first_node) {
    process_first_node();
    first_node = false;
    return cur;  // may be -1 (the reference would return a null pointer here)
  } else if (bases_remaining == 0) {
    return {};
  } else {
    go_to_next_site();
    if (cur == -1) return {};
    return cur;
  }
}
void Traverser::process_first_node() {  // :164-167
  update_coordinates();
  if (!info->graph.nodes[cur].in_bubble()) go_to_next_site();
}
void Traverser::go_to_next_site() {  // :169-189
  start_pos = 0;
  while (info->graph.nodes[cur].next.size() == 1) {
    if (bases_remaining <= 0) {
      cur = -1;
      return;
    }
    cur = info->graph.nodes[cur].next[0];  // move_past_single_edge_node :196-199
    update_coordinates();
    if (info->graph.nodes[cur].in_bubble()) return;
  }
  --traversed_index;
  choose_allele();
  update_coordinates();
}
void Traverser::update_coordinates() {  // :191-194
  assign_end_position();
  if (info->graph.nodes[cur].has_sequence()) bases_remaining -= (end_pos - start_pos + 1);
}
void Traverser::assign_end_position() {  // :201-206
  end_pos = 0;
  std::size_t seq_size = info->graph.nodes[cur].seq.size();
  if (seq_size > 0) end_pos = (uint32_t)std::min(seq_size - 1, (std::size_t)start_pos + bases_remaining - 1);
}
void Traverser::choose_allele() {  // :208-219
  if (traversed_index >= traversed.size()) throw std::logic_error("Traverser ran out of traversed loci");
  auto locus = traversed[traversed_index];
  const Node& n = info->graph.nodes[cur];
  if (locus.second < 0 || (std::size_t)locus.second >= n.next.size())
    throw std::logic_error("Traverser: allele id out of range");
  int next = n.next[locus.second];
  const Node& nn = info->graph.nodes[next];
  if (nn.has_sequence()) assert(nn.site == locus.first && nn.allele == locus.second);
  cur = next;
}

static void process_node(const PRGInfo& info, CovMapping& m, int node, uint32_t s, uint32_t e) {  // :282-296
  const Node& n = info.graph.nodes[node];
  if (!n.has_sequence()) return;
  std::size_t size = n.seq.size();
  auto it = m.find(node);
  if (it == m.end()) {  // DummyCovNode ctor :109-123
    if (s > e) throw std::logic_error("start_pos must not be greater than end_pos");
    if (s >= size || e >= size) throw std::logic_error("node_size must be greater than start_pos and end_pos");
    m[node] = {{s, e}, (e - s == size - 1)};
  } else {  // extend_coordinates :125-135
    auto& d = it->second;
    if (e >= size) throw std::logic_error("end coordinate must be less than node_size");
    if (d.second) return;
    if (s < d.first.first) d.first.first = s;
    if (e > d.first.second) d.first.second = e;
    if (d.first.second - d.first.first == size - 1) d.second = true;
  }
}

CovMapping pb_cov_mapping(const PRGInfo& info, const SearchStates& states, std::size_t read_size, Events* ev) {  // :221-280
  CovMapping m;
  for (const auto& ss : states) {  // process_SearchState :244-270
    bool first = true;
    for (uint64_t occ = ss.lo; occ <= ss.hi; ++occ) {
      auto coordinate = info.fm.sa[occ];
      if (ev) ev->q_sa++, ev->q_node++;
      Traverser t(info, info.graph.random_access[coordinate], ss.traversed, read_size);
      if (first) {
        first = false;
        auto cur = t.next_node();  // record_full_traversal :272-280
        while (cur && *cur != -1) {
          process_node(info, m, *cur, t.start_pos, t.end_pos);
          cur = t.next_node();
        }
      } else {
        auto cur = t.next_node();
        if (cur && *cur != -1) process_node(info, m, *cur, t.start_pos, t.end_pos);
      }
    }
  }
  return m;
}

void record_allele_base(PRGInfo& info, const SearchStates& states, std::size_t read_size, bool atomic, Events* ev) {
  auto m = pb_cov_mapping(info, states, read_size, ev);
  for (auto& e : m) {  // write_coverage_from_dummy_nodes :230-242
    auto& cov = info.graph.nodes[e.first].cov;
    if (cov.empty()) continue;  // nodes outside bubbles hold no counters (coverage_graph.cpp:27-30)
    for (uint32_t i = e.second.first.first; i <= e.second.first.second; ++i) {
      if (ev) ev->a_cov++;
      if (atomic) {
        CovCount v;
#pragma omp atomic read
        v = cov[i];
        if (v == UINT16_MAX) continue;
#pragma omp atomic
        cov[i]++;
      } else {
        if (cov[i] == UINT16_MAX) continue;
        cov[i]++;
      }
    }
  }
}

void record_allele_sum(Coverage& c, const std::set<VariantLocus>& loci, bool atomic) {  // allele_sum.cpp:31-43
  for (auto& l : loci) {
    auto idx = siteID_to_index(l.first);
    if (atomic) {
#pragma omp atomic
      c.allele_sum[idx][l.second] += 1;
    } else
      c.allele_sum[idx][l.second] += 1;
  }
}

void record_grouped(Coverage& c, const std::set<VariantLocus>& loci) {  // grouped_allele_counts.cpp:17-49
  std::map<Marker, std::set<AlleleId>> groups;
  for (auto& l : loci) groups[l.first].insert(l.second);
  for (auto& e : groups) {
    AlleleIds ids(e.second.begin(), e.second.end());
    auto idx = siteID_to_index(e.first);
#pragma omp critical(gqo_grouped)
    c.grouped[idx][ids] += 1;
  }
}

Coverage empty_coverage(const PRGInfo& info) {  // coverage_common.cpp:206-212, allele_sum.cpp:10-29
  Coverage c;
  c.grouped.assign(info.num_sites, {});
  c.allele_sum.assign(info.num_sites, {});
  for (auto& b : info.graph.bubbles) {
    const Node& start = info.graph.nodes[b.first];
    auto idx = siteID_to_index(start.site);
    if (idx >= c.allele_sum.size()) throw std::runtime_error("site ids are not contiguous from 5");
    c.allele_sum[idx].assign(start.next.size(), 0);
  }
  return c;
}

std::vector<std::vector<std::vector<CovCount>>> allele_base_non_nested(const PRGInfo& info) {  // allele_base.cpp:10-38
  std::vector<std::vector<std::vector<CovCount>>> out;
  if (info.graph.is_nested) return out;
  out.resize(info.num_sites);
  for (auto& b : info.graph.bubbles) {
    const Node& start = info.graph.nodes[b.first];
    auto& site = out.at(siteID_to_index(start.site));
    for (int a : start.next) {
      const Node& an = info.graph.nodes[a];
      if (an.is_bubble_end()) site.emplace_back();
      else site.push_back(an.cov);
    }
  }
  return out;
}

// =====================================================================================
// quasimap_read / forward+reverse (quasimap.cpp:82-194)
// =====================================================================================
Mapper::Mapper(const std::vector<Marker>& prg, uint32_t k_) : k(k_) {
  info = build_prg_info(prg);
  kmers = build_kmer_index(info, k);
  cov = empty_coverage(info);
}

StrandStatus Mapper::quasimap_read(const Sequence& read, uint32_t seed, bool atomic, SearchStates* out_states) {
  Events* ev = count_events ? &events : nullptr;
  if (ev) ev->strands++;
  auto bump = [&](uint64_t& c) {
    if (atomic) {
#pragma omp atomic
      c += 1;
    } else
      c += 1;
  };
  // reads shorter than k are UB in the reference (quasimap.cpp:206-210); defined here as missing_kmer
  if (read.size() < k || !all_kmers_in_index(kmers, read, k)) {
    bump(stats.missing_kmer);
    return MISSING_KMER;
  }
  auto states = search_read_backwards(info, kmers, read, k, ev);
  if (states.empty()) {
    bump(stats.no_extension);
    return NO_EXTENSION;
  }
  // coverage::record::search_states (coverage_common.cpp:179-197)
  auto sel = select_mapping(info, states, seed, std::nullopt, ev);
  if (!sel.states.empty()) {
    record_allele_base(info, sel.states, read.size(), atomic, ev);
    record_allele_sum(cov, sel.loci, atomic);
    record_grouped(cov, sel.loci);
  }
  bump(stats.exact_mapped);
  if (out_states) *out_states = std::move(states);
  return MAPPED;
}

void Mapper::quasimap_forward_reverse(const Sequence& read, uint32_t seed, bool atomic, StrandStatus* st,
                                      SearchStates* fwd, SearchStates* rev) {
  if (atomic) {
#pragma omp atomic
    stats.all_reads += 2;
  } else
    stats.all_reads += 2;
  if (read.empty()) {  // quasimap.cpp:108-113
    if (atomic) {
#pragma omp atomic
      stats.skipped += 2;
    } else
      stats.skipped += 2;
    if (st) st[0] = st[1] = SKIPPED;
    return;
  }
  auto s0 = quasimap_read(read, seed, atomic, fwd);
  auto s1 = quasimap_read(reverse_complement(read), seed, atomic, rev);
  if (st) st[0] = s0, st[1] = s1;
}

std::vector<CovCount> Mapper::per_base_flat() const {
  std::vector<CovCount> out;
  const auto& ra = info.graph.random_access;
  for (std::size_t p = 0; p < info.prg.size(); ++p) {
    if (info.prg[p] > 4) continue;
    const Node& n = info.graph.nodes[ra[p].node];
    if (!n.in_bubble()) continue;
    out.push_back(n.cov[(std::size_t)((int64_t)p - n.prg_start)]);
  }
  return out;
}

std::vector<uint32_t> canonical_states(const SearchStates& states) {
  std::vector<std::vector<uint32_t>> recs;
  for (auto& s : states) {
    std::vector<uint32_t> r{s.lo, s.hi, (uint32_t)s.traversed.size(), (uint32_t)s.traversing.size()};
    for (auto& l : s.traversed) r.push_back(l.first), r.push_back((uint32_t)l.second);
    for (auto& l : s.traversing) r.push_back(l.first), r.push_back((uint32_t)l.second);
    recs.push_back(std::move(r));
  }
  std::sort(recs.begin(), recs.end());
  std::vector<uint32_t> out;
  for (auto& r : recs) out.insert(out.end(), r.begin(), r.end());
  return out;
}

}  // namespace gqo
