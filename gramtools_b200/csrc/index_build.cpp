// index_build.cpp — see index_build.hpp. Host-only; compiled with the rest of libgq.so.
#include "index_build.hpp"

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "sais.hpp"

namespace gq {

IndexView HostIndex::view() const {
  IndexView v{};
  v.n = n;
  v.rank_blk = rank_blk.data();
  v.super_cnt = super_cnt.data();
  v.mrank_blk = mrank_blk.data();
  v.marker_hit = marker_hit.data();
  v.text_grp = text_grp.data();
  v.text_super = text_super.data();
  v.tmarker_hit = tmarker_hit.data();
  v.isa = isa.data();
  for (int i = 0; i < 4; ++i) v.c_base[i] = c_base[i];
  v.n_slots = n_slots;
  v.any_nested = is_nested ? 1u : 0u;
  v.site_sa = site_sa.data();
  v.allele_iv = allele_iv.data();
  v.par = par.data();
  v.tm_odd = tm_odd.data();
  v.tm_even_off = tm_even_off.data();
  v.tm_even = tm_even.data();
  v.entry_next = entry_next.data();
  v.site_snp = site_snp.data();
  v.sa = sa.data();
  v.pos2node = pos2node.data();
  v.nodes = nodes.data();
  v.edges = edges.data();
  v.site_rec = site_rec.data();
  v.apos = apos.data();
  v.k = k;
  v.kmer_bits = kmer_bits.data();
  v.kmer_off = kmer_off.data();
  v.kmer_states = kmer_states.data();
  v.seed_off = seed_off.data();
  v.seed_ent = seed_ent.data();
  v.seed_state = seed_state.data();
  v.kmer_paths = kmer_paths.data();
  return v;
}

int32_t compress_text(const uint32_t* prg, uint64_t n_symbols, std::vector<uint32_t>& present, std::vector<int32_t>& text) {
  uint32_t maxs = 0;
  const int64_t n_i = (int64_t)n_symbols;
#pragma omp parallel for reduction(max : maxs) schedule(static)
  for (int64_t i = 0; i < n_i; ++i) maxs = std::max(maxs, prg[i]);
  present.clear();
  if ((uint64_t)maxs <= 4 * n_symbols + 1024) {
    // symbols are dense (bases and consecutive markers): a presence table instead of sorting a copy of the PRG
    std::vector<uint8_t> seen((size_t)maxs + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_i; ++i)  // relaxed atomic byte accesses: several threads may mark the same symbol
      if (!__atomic_load_n(&seen[prg[i]], __ATOMIC_RELAXED)) __atomic_store_n(&seen[prg[i]], (uint8_t)1, __ATOMIC_RELAXED);
    for (uint64_t sym = 0; sym <= maxs; ++sym)
      if (seen[sym]) present.push_back((uint32_t)sym);
  } else {
    present.assign(prg, prg + n_symbols);
    std::sort(present.begin(), present.end());
    present.erase(std::unique(present.begin(), present.end()), present.end());
  }
  const int32_t sigma = (int32_t)present.size() + 1;
  text.assign(n_symbols + 1, 0);
  if ((uint64_t)maxs <= 4 * n_symbols + 1024) {
    std::vector<int32_t> tab((size_t)maxs + 1, 0);
    for (size_t i = 0; i < present.size(); ++i) tab[present[i]] = (int32_t)i + 1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_i; ++i) text[i] = tab[prg[i]];
  } else {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_i; ++i)
      text[i] = (int32_t)(std::lower_bound(present.begin(), present.end(), prg[i]) - present.begin()) + 1;
  }
  text[n_symbols] = 0;  // the sentinel sdsl appends
  return sigma;
}

namespace {

struct OpenSite {
  uint32_t site;
  uint32_t allele;
};

// One left-to-right pass over the PRG builds the flat graph, the parent map, both target maps and
// the per-position marker targets. Reference semantics: cov_Graph_Builder
// (libgramtools/src/prg/coverage_graph.cpp:82-379); see SURVEY.md Appendix A for the derivations.
void build_graph(HostIndex& ix, std::vector<uint32_t>& hit_marker, std::vector<uint32_t>& hit_allele) {
  const auto& prg = ix.prg;
  const uint32_t L = (uint32_t)prg.size();
  uint32_t maxm = 4;
  for (auto m : prg) {
    if (m < 1) throw std::runtime_error("PRG symbols must be >= 1");
    maxm = std::max(maxm, m);
  }
  ix.n_slots = maxm > 4 ? ((maxm % 2 ? maxm : maxm - 1) - 5) / 2 + 1 : 0;
  const uint32_t S = ix.n_slots;
  // final (site end) position of every even marker; duplicate site-entry markers are an error
  std::vector<uint32_t> last_pos(S, 0xFFFFFFFFu);
  std::vector<uint8_t> seen(S, 0);
  for (uint32_t p = 0; p < L; ++p) {
    uint32_t m = prg[p];
    if (m <= 4) continue;
    if (m & 1u) {
      uint32_t s = (m - 5) / 2;
      if (seen[s])
        throw std::runtime_error("PRG consistency error: site marker " + std::to_string(m) +
                                 " used for two different sites");
      seen[s] = 1;
    } else {
      if (m < 6) throw std::runtime_error("invalid marker");
      last_pos[(m - 6) / 2] = p;
    }
  }
  ix.n_sites = 0;
  for (uint32_t s = 0; s < S; ++s) {
    if (seen[s] && last_pos[s] == 0xFFFFFFFFu)
      throw std::runtime_error("PRG consistency error: site " + std::to_string(5 + 2 * s) + " is never closed");
    if (!seen[s] && last_pos[s] != 0xFFFFFFFFu)
      throw std::runtime_error("PRG consistency error: allele marker without its site marker");
    ix.n_sites += seen[s];
  }

  ix.par.assign(2 * (size_t)S, 0);
  ix.tm_odd.assign(S, 0);
  ix.n_alleles.assign(S, 0);
  ix.site_start_node.assign(S, 0);
  ix.site_start_pos.assign(S, 0);
  std::vector<std::vector<uint32_t>> tm_even(S);
  ix.pos2node.assign(L, 0);
  hit_marker.assign(L, 0);
  hit_allele.assign(L, 0);
  std::vector<uint32_t> site_end_node(S, 0);

  // adjacency lists are built as (from,to) pairs in wiring order, then bucketed into CSR; the
  // bubble-start node of a site receives exactly one edge per allele, in allele order.
  std::vector<std::pair<uint32_t, uint32_t>> wires;
  wires.reserve(L / 2 + 16);
  auto new_node = [&](uint32_t site, int32_t allele, uint32_t start) {
    Node nd{};
    nd.site = site;
    nd.allele = allele;
    nd.start = start;
    nd.len = 0;
    nd.cov_off = kNoAllele;
    ix.nodes.push_back(nd);
    return (uint32_t)ix.nodes.size() - 1;
  };
  ix.nodes.clear();
  uint32_t back = new_node(0, -1, 0);  // root
  int64_t run = -1;                    // open sequence node
  std::vector<OpenSite> open;
  auto wire = [&](uint32_t target) {  // coverage_graph.cpp:260-266
    if (run >= 0) {
      wires.emplace_back(back, (uint32_t)run);
      wires.emplace_back((uint32_t)run, target);
    } else
      wires.emplace_back(back, target);
    run = -1;
  };
  enum { T_SEQ, T_ENTRY, T_ALLELE_END, T_SITE_END };
  int prev_t = T_SEQ;
  uint32_t prev_m = 0;
  uint32_t n_per_base = 0;
  for (uint32_t p = 0; p < L; ++p) {
    uint32_t m = prg[p];
    int t;
    if (m <= 4) t = T_SEQ;
    else if (m & 1u) t = T_ENTRY;
    else t = (p < last_pos[(m - 6) / 2]) ? T_ALLELE_END : T_SITE_END;
    const uint32_t cur_allele = open.empty() ? kNoAllele : open.back().allele;  // before this symbol
    switch (t) {
      case T_SEQ: {
        if (run < 0) {
          uint32_t site = open.empty() ? 0 : open.back().site;
          int32_t allele = open.empty() ? -1 : (int32_t)open.back().allele;
          run = new_node(site, allele, p);
          if (site != 0) ix.nodes[run].cov_off = n_per_base;
        }
        ix.nodes[run].len++;
        if (ix.nodes[run].cov_off != kNoAllele) n_per_base++;
        ix.pos2node[p] = (uint32_t)run;
        if (prev_t != T_SEQ) {  // map_targets, sequence case (:281-286) + left_markers_search conversion
          if (prev_t == T_ALLELE_END) {
            hit_marker[p] = prev_m - 1;  // an allele separator: leaving the site through this allele
            hit_allele[p] = cur_allele;
          } else {
            hit_marker[p] = prev_m;  // odd: leave through allele 0; even (site end): enter the site
            hit_allele[p] = cur_allele;
          }
        }
        break;
      }
      case T_ENTRY: {
        uint32_t s = (m - 5) / 2;
        uint32_t sn = new_node(m, -1, p), en = new_node(m, -1, p);
        ix.site_start_node[s] = sn;
        ix.site_start_pos[s] = p;
        site_end_node[s] = en;
        wire(sn);
        back = sn;
        if (!open.empty()) {
          ix.par[2 * s] = open.back().site;
          ix.par[2 * s + 1] = open.back().allele;
          ix.is_nested = true;
        }
        if (prev_t != T_SEQ)  // make_site_entry_target :313-328
          ix.tm_odd[s] = (prev_t == T_ALLELE_END) ? prev_m - 1 : prev_m;
        open.push_back({m, 0});
        ix.pos2node[p] = sn;
        break;
      }
      case T_ALLELE_END:
      case T_SITE_END: {
        uint32_t s = (m - 6) / 2;
        if (open.empty() || open.back().site != m - 1)
          throw std::runtime_error("PRG consistency error: allele marker " + std::to_string(m) +
                                   " outside its site");
        if (prev_t != T_SEQ) {
          if (t == T_SITE_END) {  // make_site_exit_target :330-350
            if (prev_t == T_ENTRY)
              throw std::runtime_error("PRG consistency error: site number " + std::to_string(m) + " is empty");
            if (prev_t == T_SITE_END) {
              tm_even[s].push_back(prev_m);
              tm_even[s].push_back(kNoAllele);
            } else {
              tm_even[s].push_back(prev_m - 1);
              tm_even[s].push_back(cur_allele);
            }
          } else {  // make_allele_end_target :352-370
            if (prev_t == T_ENTRY) {
              tm_even[s].push_back(prev_m);
              tm_even[s].push_back(cur_allele);
            } else if (prev_t == T_SITE_END) {
              tm_even[s].push_back(prev_m);
              tm_even[s].push_back(kNoAllele);
            } else {
              tm_even[s].push_back(prev_m - 1);
              tm_even[s].push_back(cur_allele);
            }
          }
        }
        wire(site_end_node[s]);
        if (t == T_ALLELE_END) {
          back = ix.site_start_node[s];
          open.back().allele++;
          ix.pos2node[p] = ix.site_start_node[s];
        } else {
          if (open.back().allele == 0)
            throw std::runtime_error("Site numbered " + std::to_string(m) + " has only one allele");
          ix.n_alleles[s] = open.back().allele + 1;
          open.pop_back();
          back = site_end_node[s];
          ix.pos2node[p] = site_end_node[s];
        }
        break;
      }
    }
    prev_t = t;
    prev_m = m;
  }
  if (!open.empty()) throw std::runtime_error("PRG consistency error: unterminated site");
  uint32_t sink = new_node(0, -1, L);
  wire(sink);
  ix.n_per_base = n_per_base;

  // CSR edges (stable in wiring order)
  std::vector<uint32_t> deg(ix.nodes.size() + 1, 0);
  for (auto& w : wires) deg[w.first + 1]++;
  for (size_t i = 0; i < ix.nodes.size(); ++i) deg[i + 1] += deg[i];
  ix.edges.assign(wires.size(), 0);
  std::vector<uint32_t> fill(deg.begin(), deg.end() - 1);
  for (auto& w : wires) ix.edges[fill[w.first]++] = w.second;
  for (size_t i = 0; i < ix.nodes.size(); ++i) {
    ix.nodes[i].edge_off = deg[i];
    ix.nodes[i].n_edges = deg[i + 1] - deg[i];
    ix.nodes[i].next0 = deg[i + 1] > deg[i] ? ix.edges[deg[i]] : 0;
  }
  // target_map[even] CSR, allele_sum layout
  ix.tm_even_off.assign(S + 1, 0);
  for (uint32_t s = 0; s < S; ++s) ix.tm_even_off[s + 1] = ix.tm_even_off[s] + (uint32_t)tm_even[s].size() / 2;
  ix.tm_even.clear();
  for (uint32_t s = 0; s < S; ++s) ix.tm_even.insert(ix.tm_even.end(), tm_even[s].begin(), tm_even[s].end());
  ix.allele_off.assign(S + 1, 0);
  for (uint32_t s = 0; s < S; ++s) ix.allele_off[s + 1] = ix.allele_off[s] + ix.n_alleles[s];
  if (ix.tm_even.empty()) ix.tm_even.assign(2, 0);
}

// Per-site text-position tables (record_strand's table route): where every allele starts in the PRG and where its
// bases sit in the flat per-base vector — which is filled in PRG order by build_graph, so for a site without
// nested sites the bases of allele a follow those of allele a - 1 directly.
void build_site_tables(HostIndex& ix) {
  const auto& prg = ix.prg;
  const uint32_t S = ix.n_slots;
  ix.site_rec.assign(4 * (size_t)S + 4, 0);
  ix.apos.assign((size_t)ix.allele_off[S] + S + 1, 0);
  std::vector<uint32_t> open, cur_allele(S, 0);
  uint32_t in_site_bases = 0;
  for (uint32_t p = 0; p < (uint32_t)prg.size(); ++p) {
    const uint32_t m = prg[p];
    if (m <= 4) {
      if (!open.empty()) ++in_site_bases;
      continue;
    }
    if (m & 1u) {
      const uint32_t s = (m - 5) / 2, a2 = ix.allele_off[s] + s;
      open.push_back(s);
      cur_allele[s] = 0;
      ix.site_rec[4 * (size_t)s] = a2;
      ix.site_rec[4 * (size_t)s + 1] = in_site_bases;
      ix.site_rec[4 * (size_t)s + 2] = p + 1;
      ix.apos[a2] = p + 1;
    } else {
      const uint32_t s = (m - 6) / 2, a2 = ix.allele_off[s] + s;
      ++cur_allele[s];
      ix.apos[a2 + cur_allele[s]] = p + 1;
      if (cur_allele[s] == ix.n_alleles[s]) {
        ix.site_rec[4 * (size_t)s + 3] = p + 1;
        open.pop_back();
      }
    }
  }
}

void build_fm(HostIndex& ix, const std::vector<uint32_t>& hit_marker, const std::vector<uint32_t>& hit_allele,
              SaBuilder sa_builder, void* sa_ctx) {
  const auto& prg = ix.prg;
  const uint32_t n = (uint32_t)prg.size() + 1;
  ix.n = n;
  // compressed alphabet: rank among the symbols present (sdsl's char2comp)
  std::vector<uint32_t> present;
  std::vector<int32_t> text;
  const int32_t sigma = compress_text(prg.data(), prg.size(), present, text);
  auto comp = [&](uint32_t sym) {
    return (int32_t)(std::lower_bound(present.begin(), present.end(), sym) - present.begin()) + 1;
  };
  const auto t_sa0 = std::chrono::steady_clock::now();
  if (sa_builder) ix.sa = sa_builder(text, sigma, sa_ctx);
  else ix.sa = suffix_array_u32(text, sigma, std::getenv("GQ_SAIS64") != nullptr);
  if (ix.sa.size() != n) throw std::runtime_error("suffix array builder returned a wrong size");
  if (std::getenv("GQ_BUILD_TIMING"))
    fprintf(stderr, "index build:   %-26s %.2f s\n", sa_builder ? "suffix array (GPU)" : "suffix array (SA-IS)",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sa0).count());
  // C array
  std::vector<uint32_t> C(sigma + 1, 0);
  for (uint32_t i = 0; i < n; ++i) C[text[i] + 1]++;
  for (int32_t c = 0; c < sigma; ++c) C[c + 1] += C[c];
  for (uint32_t b = 1; b <= 4; ++b) {
    bool has = std::binary_search(present.begin(), present.end(), b);
    int32_t cb = has ? comp(b) : (int32_t)(std::lower_bound(present.begin(), present.end(), b) - present.begin()) + 1;
    ix.c_base[b - 1] = C[cb];
  }
  const uint32_t S = ix.n_slots;
  ix.site_sa.assign(S, 0);
  ix.allele_iv.assign(2 * (size_t)S, 0);
  for (uint32_t s = 0; s < S; ++s) {
    if (ix.n_alleles[s] == 0) continue;
    int32_t co = comp(5 + 2 * s), ce = comp(6 + 2 * s);
    ix.site_sa[s] = C[co];
    ix.allele_iv[2 * s] = C[ce];          // get_allele_marker_sa_interval, vBWT_jump.cpp:3-21
    ix.allele_iv[2 * s + 1] = C[ce + 1] - 1;
  }
  // rank blocks, marker ranks and marker targets in BWT order — two passes over the blocks of 64 BWT positions, both
  // parallel: (1) the bit planes of every block and its per-base / marker counts, (2) after a prefix sum over the
  // block counts, the counters in front of every block and the jump records of its markers at their BWT rank
  const uint32_t nblk = (n >> kBlkShift) + 1;
  const uint32_t nsuper = (n >> kSuperShift) + 1;
  ix.rank_blk.assign(nblk, RankBlk{});
  ix.super_cnt.assign(4 * (size_t)nsuper, 0);
  ix.mrank_blk.assign(nblk, 0);
  std::vector<uint32_t> blk_cnt(5 * (size_t)nblk + 5, 0);  // per block: A, C, G, T, markers (then prefix sums)
  const int64_t nblk_i = (int64_t)nblk;
#pragma omp parallel for schedule(static)
  for (int64_t bi = 0; bi < nblk_i; ++bi) {
    RankBlk& b = ix.rank_blk[bi];
    uint32_t cnt[5] = {0, 0, 0, 0, 0};
    const uint64_t i0 = (uint64_t)bi << kBlkShift, i1 = std::min<uint64_t>(i0 + 64, n);
    for (uint64_t i = i0; i < i1; ++i) {
      const uint32_t p = ix.sa[i];
      const uint32_t sym = p ? prg[p - 1] : 0;
      const uint64_t bit = 1ull << (i & 63u);
      if (sym >= 1 && sym <= 4) {
        const uint32_t c = sym - 1;
        if (c & 1) b.p0 |= bit;
        if (c & 2) b.p1 |= bit;
        cnt[c]++;
      } else {
        b.p2 |= bit;
        if (sym > 4) {
          b.p0 |= bit;
          cnt[4]++;
        }
      }
    }
    for (int c = 0; c < 5; ++c) blk_cnt[5 * (size_t)bi + c] = cnt[c];
  }
  {  // exclusive prefix sums over the blocks (n / 64 entries: serial)
    uint32_t run[5] = {0, 0, 0, 0, 0};
    for (size_t bi = 0; bi < nblk; ++bi)
      for (int c = 0; c < 5; ++c) {
        const uint32_t x = blk_cnt[5 * bi + c];
        blk_cnt[5 * bi + c] = run[c];
        run[c] += x;
      }
    for (int c = 0; c < 5; ++c) blk_cnt[5 * (size_t)nblk + c] = run[c];
  }
  const uint32_t nmark_total = blk_cnt[5 * (size_t)nblk + 4];
  ix.marker_hit.assign(8 * (size_t)nmark_total, 0);
  std::vector<uint32_t> bwt_marker_pos(nmark_total);  // text position of the marker behind each BWT marker occurrence
  constexpr uint32_t kBlkPerSuper = 1u << (kSuperShift - kBlkShift);
#pragma omp parallel for schedule(static)
  for (int64_t bi = 0; bi < nblk_i; ++bi) {
    const uint32_t* tot = &blk_cnt[5 * (size_t)bi];
    const uint32_t* sup = &blk_cnt[5 * (size_t)(bi & ~(int64_t)(kBlkPerSuper - 1))];  // counts at the superblock's start
    if ((bi & (kBlkPerSuper - 1)) == 0)
      for (int c = 0; c < 4; ++c) ix.super_cnt[4 * (size_t)(bi >> (kSuperShift - kBlkShift)) + c] = ix.c_base[c] + tot[c];
    RankBlk& b = ix.rank_blk[bi];
    b.cnt = 0;
    for (int c = 0; c < 4; ++c) b.cnt |= (uint64_t)((tot[c] - sup[c]) & 0xFFFFu) << (16 * c);
    uint32_t nmark = tot[4];
    ix.mrank_blk[bi] = nmark;
    uint64_t mbits = b.p2 & b.p0;  // the block's markers, in BWT order
    while (mbits) {
      const uint64_t i = ((uint64_t)bi << kBlkShift) + (uint64_t)__builtin_ctzll(mbits);
      mbits &= mbits - 1;
      const uint32_t p = ix.sa[i];
      bwt_marker_pos[nmark] = p - 1;
      const bool at_base = p < prg.size() && prg[p] <= 4;
      const uint32_t hm = at_base ? hit_marker[p] : 0, ha = at_base ? hit_allele[p] : 0;
      uint32_t jlo = kNoAllele, jhi = kNoAllele;
      if (hm > 4) {
        if (hm & 1u) {  // exit: simple when nothing is adjacent to the left of the site (tm_odd empty)
          const uint32_t slot = (hm - 5) / 2;
          if (ix.tm_odd[slot] == 0) jlo = jhi = ix.site_sa[slot];
        } else {  // entry: simple when the site has no empty allele / nested site at an allele end
          const uint32_t slot = (hm - 6) / 2;
          if (ix.tm_even_off[slot + 1] == ix.tm_even_off[slot]) {
            jlo = ix.allele_iv[2 * slot];
            jhi = ix.allele_iv[2 * slot + 1];
          }
        }
      }
      // words 4,5 (SNP crossing table, site_sa of the entered site) are filled once the per-site tables exist; 6,7
      // (text positions behind the post-jump SA indices) below: the record is one 32 B sector
      uint32_t* rec = &ix.marker_hit[8 * (size_t)nmark];
      rec[0] = hm;
      rec[1] = ha;
      rec[2] = jlo;
      rec[3] = jhi;
      rec[4] = kNotSnp;
      ++nmark;
    }
  }
  if (ix.marker_hit.empty()) ix.marker_hit.assign(8, 0);

  // ---- pre-resolved site crossings (used by lane_event_scan) ----
  // entry_next[s][c]: interval after entering site s from its right end and consuming base c
  // (entering_site_search_state + base_next_sa_interval); site_snp[s]: the whole crossing of a site
  // made of distinct single-base alleles with no marker adjacent on either side.
  auto rank_c = [&](uint32_t c, uint32_t i) {
    const RankBlk& b = ix.rank_blk[i >> kBlkShift];
    return rank_in_blk(b, ix.super_cnt.data() + 4 * (size_t)((i >> kBlkShift) >> (kSuperShift - kBlkShift)), c, i);
  };
  ix.entry_next.assign(8 * (size_t)S + 8, 0);
  ix.site_snp.assign((size_t)S + 1, kNotSnp);
  for (uint32_t s = 0; s < S; ++s) {
    if (ix.n_alleles[s] == 0) continue;
    for (uint32_t c = 0; c < 4; ++c) {
      uint32_t r0 = rank_c(c, ix.allele_iv[2 * s]), r1 = rank_c(c, ix.allele_iv[2 * s + 1] + 1);
      ix.entry_next[8 * (size_t)s + 2 * c] = r0;
      ix.entry_next[8 * (size_t)s + 2 * c + 1] = r1 - 1;  // r1 == r0 -> hi = lo - 1: empty
    }
    if (ix.tm_odd[s] != 0 || ix.tm_even_off[s + 1] != ix.tm_even_off[s]) continue;
    const uint32_t p0 = ix.site_start_pos[s], na = ix.n_alleles[s];
    if ((size_t)p0 + 2 * na >= prg.size() + 1) continue;
    uint32_t tab = 0xFFFFFFFFu;
    bool ok = true;
    for (uint32_t a = 0; a < na && ok; ++a) {
      uint32_t base = prg[p0 + 1 + 2 * a], sep = prg[p0 + 2 + 2 * a];
      ok = base >= 1 && base <= 4 && sep == 6 + 2 * s && a < 0xFF;
      if (!ok) break;
      uint32_t c = base - 1;
      if (((tab >> (8 * c)) & 0xFFu) != 0xFFu) ok = false;  // two alleles with the same base
      else tab = (tab & ~(0xFFu << (8 * c))) | (a << (8 * c));
    }
    if (ok) ix.site_snp[s] = tab;
  }
  const int64_t n_hit_words = (int64_t)ix.marker_hit.size();
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < n_hit_words - 7; m += 8) {
    uint32_t hm = ix.marker_hit[m];
    if (hm > 4 && (hm & 1u) == 0 && ix.marker_hit[m + 2] != kNoAllele) {  // simple entry
      uint32_t slot = (hm - 6) / 2;
      ix.marker_hit[m + 4] = ix.site_snp[slot];
      ix.marker_hit[m + 5] = ix.site_sa[slot];
      ix.marker_hit[m + 7] = ix.sa[ix.site_sa[slot]];
    }
    const uint32_t jlo = ix.marker_hit[m + 2], jhi = ix.marker_hit[m + 3];
    ix.marker_hit[m + 6] = (jlo != kNoAllele && jlo == jhi) ? ix.sa[jlo] : kNoAllele;
  }

  // ---- text mode: the PRG as 2-bit codes + marker flags, jump records in text order, inverse SA ----
  const uint32_t L = (uint32_t)prg.size();
  ix.isa.assign(n, 0);
  const int64_t n_i = (int64_t)n;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_i; ++i) ix.isa[ix.sa[i]] = (uint32_t)i;
  const uint32_t ngrp = (L >> 4) + 1;
  ix.text_grp.assign(ngrp, TextGrp{0, 0});
  ix.text_super.assign((L >> kTextSuperShift) + 1, 0);
  std::vector<uint32_t> text_rank(L, 0);  // rank of a marker position among the markers of the text
  uint32_t tm = 0, tsup = 0;
  for (uint32_t q = 0; q < L; ++q) {
    if ((q & ((1u << kTextSuperShift) - 1)) == 0) ix.text_super[q >> kTextSuperShift] = tsup = tm;
    if ((q & 15u) == 0) ix.text_grp[q >> 4].info = (tm - tsup) << 16;
    const uint32_t sym = prg[q];
    if (sym > 4) {
      ix.text_grp[q >> 4].info |= 1u << (q & 15u);
      text_rank[q] = tm++;
    } else
      ix.text_grp[q >> 4].codes |= (sym - 1) << (2 * (q & 15u));
  }
  ix.tmarker_hit.assign(std::max<size_t>(8, 8 * (size_t)tm), 0);
  for (size_t mr = 0; mr < bwt_marker_pos.size(); ++mr)
    std::copy(ix.marker_hit.begin() + 8 * mr, ix.marker_hit.begin() + 8 * mr + 8,
              ix.tmarker_hit.begin() + 8 * (size_t)text_rank[bwt_marker_pos[mr]]);
}

// ---- all-k-mers index (reference: src/build/kmer_index/build.cpp:18-148) ---------------------
// The search states of one k-mer prefix: fixed-size records + one pool of path words (nt pairs, then ng sites) — no
// allocation per state (a std::vector per state made malloc the hot spot of the build).
struct HRec {
  uint32_t lo, hi, counts, path_off;
};
struct StateList {
  std::vector<HRec> recs;
  std::vector<uint32_t> pool;
  void clear() {
    recs.clear();
    pool.clear();
  }
  bool empty() const { return recs.empty(); }
  void swap(StateList& o) {
    recs.swap(o.recs);
    pool.swap(o.pool);
  }
};

// All successors of `in` after consuming each of the four bases (marker processing first unless `first`): the marker
// processing of a state does not depend on the base consumed next, so it runs once (run_stack_ready) and every
// resulting state is extended by A, C, G and T with one pair of rank-block loads. out[c] holds, per base, the same
// states as the reference's pass (build.cpp:55-131), in its order: the extended input states, then the marker-derived
// ones as its worklist commits them (tests/test_host_parity.py::test_kmer_index_holds_the_oracles_states).
struct ReadyCollect {
  const IndexView* v;
  StateList* out;      // [4]: extensions of the input state itself
  StateList* derived;  // [4]: extensions of the states its markers give
  uint32_t in_lo, in_hi, in_counts;
  void operator()(const uint32_t* t, uint32_t lo, uint32_t hi) {
    const uint32_t b0 = lo >> kBlkShift, b1 = (hi + 1) >> kBlkShift;
    const RankBlk B0 = load_blk(v->rank_blk + b0);
    const RankBlk B1 = (b1 == b0) ? B0 : load_blk(v->rank_blk + b1);
    const uint32_t* s0 = v->super_cnt + 4 * (b0 >> (kSuperShift - kBlkShift));
    const uint32_t* s1 = v->super_cnt + 4 * (b1 >> (kSuperShift - kBlkShift));
    const uint32_t w = entry_words(t[3]);
    // every jump lengthens the path, so the input state is the one record that still has its interval and counts
    const bool is_input = lo == in_lo && hi == in_hi && t[3] == in_counts;
    StateList* dst = is_input ? out : derived;
    for (uint32_t c = 0; c < 4; ++c) {
      const uint32_t r0 = rank_in_blk(B0, s0, c, lo), r1 = rank_in_blk(B1, s1, c, hi + 1);
      if (r1 <= r0) continue;
      StateList& o = dst[c];
      const HRec rec{r0, r1 - 1, t[3], (uint32_t)o.pool.size()};  // C[c] is folded into the superblock counters
      // a state is emitted after the states chained to it (they sit above it on the stack); the reference commits
      // it before them: `before` = where its chain began in this list
      if (!is_input && before && before[c] < o.recs.size()) o.recs.insert(o.recs.begin() + before[c], rec);
      else o.recs.push_back(rec);
      o.pool.insert(o.pool.end(), t + kHdr, t + w);
    }
  }
  const size_t* before = nullptr;  // set by run_stack_in_reference_order around a call
};

// run_stack_ready (gq_core.cuh) with the reference's commit order: search_state_vBWT_jumps commits an entered / exited
// state when it is made, BEFORE the loci chained to it are processed (vBWT_jump.cpp:155-183); the stack machine emits
// it when it comes back to the top, after them. The list positions at which an entry's processing began are kept per
// stack offset, and the entry's record is inserted there.
void run_stack_in_reference_order(Stack& s, const IndexView& v, ReadyCollect& col) {
  struct Open {
    uint32_t offset;
    size_t at[4];
  };
  std::vector<Open> open;
  while (!stack_empty(s) && !s.overflow) {
    uint32_t* t = s.mem + s.top;
    const uint32_t w0 = t[0], kind = w0 >> 28, pos = w0 & 0x0FFFFFFFu;
    if (kind == K_JUMP) {
      bool known = false;
      for (auto& o : open) known |= o.offset == s.top;
      if (!known) {
        Open o{s.top, {0, 0, 0, 0}};
        for (uint32_t c = 0; c < 4; ++c) o.at[c] = col.derived[c].recs.size();
        open.push_back(o);
      }
      process_jump(s, v);
      continue;
    }
    const uint32_t lo = t[1], hi = t[2];
    if (kind == K_SCAN && interval_has_marker(v, lo, hi)) {
      t[0] = pos | (K_READY << 28);
      scan_markers(s, v, pos, lo, hi);
      continue;
    }
    col.before = nullptr;
    for (size_t i = 0; i < open.size(); ++i)
      if (open[i].offset == s.top) {
        col.before = open[i].at;
        col(t, lo, hi);
        open.erase(open.begin() + (long)i);
        col.before = nullptr;
        t = nullptr;
        break;
      }
    if (t) col(t, lo, hi);
    pop(s);
  }
}

void step_states4(const IndexView& v, const StateList& in, bool first, std::vector<uint32_t>& arena,
                  StateList* out /* [4] */) {
  // the reference appends the marker-derived states of ALL input states after the input states themselves
  // (process_markers_search_states splices them at the end, vBWT_jump.cpp:119-132) and then extends the list in
  // order: per base, out[c] = extensions of the inputs, then extensions of the derived states
  StateList derived[4];
  for (uint32_t c = 0; c < 4; ++c) out[c].clear();
  for (const HRec& st : in.recs) {
    const uint32_t pw = entry_words(st.counts) - kHdr;
    while (true) {
      Stack s;
      s.mem = arena.data();
      s.limit = (uint32_t)arena.size();
      s.overflow = false;
      s.top = 0;
      uint32_t* t = s.mem;
      t[0] = 1u | ((first ? K_READY : K_SCAN) << 28);
      t[1] = st.lo;
      t[2] = st.hi;
      t[3] = st.counts;
      t[4] = kNoAllele;
      std::copy(in.pool.begin() + st.path_off, in.pool.begin() + st.path_off + pw, t + kHdr);
      size_t mark[8], pmark[8];
      for (uint32_t c = 0; c < 4; ++c) {
        mark[c] = out[c].recs.size(), pmark[c] = out[c].pool.size();
        mark[4 + c] = derived[c].recs.size(), pmark[4 + c] = derived[c].pool.size();
      }
      ReadyCollect col{&v, out, derived, st.lo, st.hi, st.counts};
      run_stack_in_reference_order(s, v, col);
      if (!s.overflow) break;
      for (uint32_t c = 0; c < 4; ++c) {
        out[c].recs.resize(mark[c]), out[c].pool.resize(pmark[c]);
        derived[c].recs.resize(mark[4 + c]), derived[c].pool.resize(pmark[4 + c]);
      }
      arena.resize(arena.size() * 2);
    }
  }
  for (uint32_t c = 0; c < 4; ++c) {
    const uint32_t shift = (uint32_t)out[c].pool.size();
    for (HRec h : derived[c].recs) {
      h.path_off += shift;
      out[c].recs.push_back(h);
    }
    out[c].pool.insert(out[c].pool.end(), derived[c].pool.begin(), derived[c].pool.end());
  }
}

struct KmerOut {
  std::vector<uint32_t> codes;     // k-mer code of each run
  std::vector<uint32_t> n_states;  // states in that run
  std::vector<KmerState> states;   // path_off relative to this thread's `paths`
  std::vector<uint32_t> paths;
  void add(uint32_t code, const StateList& l) {
    codes.push_back(code);
    n_states.push_back((uint32_t)l.recs.size());
    const uint32_t base = (uint32_t)paths.size();
    paths.insert(paths.end(), l.pool.begin(), l.pool.end());
    for (const HRec& h : l.recs) states.push_back(KmerState{h.lo, h.hi, base + h.path_off, h.counts});
  }
};

void kmer_recurse(const IndexView& v, uint32_t k, uint32_t depth, uint32_t code, const StateList& cur,
                  std::vector<std::array<StateList, 4>>& levels, std::vector<uint32_t>& arena, KmerOut& out) {
  std::array<StateList, 4>& nxt = levels[depth];
  step_states4(v, cur, depth == 0, arena, nxt.data());
  for (uint32_t c = 0; c < 4; ++c) {
    if (nxt[c].empty()) continue;
    uint32_t ncode = code | (c << (2 * (k - 1 - depth)));  // base j of the k-mer sits at bits [2j, 2j+2)
    if (depth + 1 == k) out.add(ncode, nxt[c]);
    else kmer_recurse(v, k, depth + 1, ncode, nxt[c], levels, arena, out);  // levels[depth] stays live for the siblings
  }
}

void build_kmers(HostIndex& ix) {
  const bool timing = std::getenv("GQ_BUILD_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "index build:   %-26s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  const uint32_t k = ix.k;
  if (k < 1 || k > 14) throw std::runtime_error("kmer_size must be in [1,14]");  // command_setup.py:97-99
  const uint64_t nk = 1ull << (2 * k);
  ix.kmer_bits.assign((((nk + 31) / 32) + 3) & ~3ull, 0);  // whole 16-byte groups: copied to shared memory as uint4
  ix.kmer_off.assign(nk + 1, 0);
  IndexView v = ix.view();
  // The first `pre` bases (rightmost of the k-mer) are expanded breadth first, every prefix node once — these levels
  // hold the wide intervals whose marker scans are the expensive ones — then the 4^pre subtrees are independent tasks,
  // handed out dynamically (256 of them for k >= 4: hosts with more than 16 cores stay busy, heavy subtrees do not
  // serialise the tail). Task results are merged in the order of a depth-first walk with the second base as the
  // outermost loop, which is the order this builder has always produced (the path pool's layout depends on it).
  const uint32_t pre = std::min<uint32_t>(k, 4);
  const uint32_t ntask = 1u << (2 * pre);
  std::string err;
  std::vector<StateList> level(1);
  level[0].recs.push_back(HRec{0, ix.n - 1, 0, 0});
  for (uint32_t d = 0; d < pre && err.empty(); ++d) {  // node index: first consumed base = most significant digit
    std::vector<StateList> next(level.size() * 4);
    const int n_nodes = (int)level.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int node = 0; node < n_nodes; ++node) {
      if (level[node].empty()) continue;
      try {
        std::vector<uint32_t> arena(1u << 16);
        std::array<StateList, 4> out4;
        step_states4(v, level[node], d == 0, arena, out4.data());
        for (uint32_t c = 0; c < 4; ++c) next[4 * (size_t)node + c].swap(out4[c]);
      } catch (const std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    level.swap(next);
  }
  if (!err.empty()) throw std::runtime_error(err);
  // merge order: tasks sorted by (second base, first base, third, fourth, ...)
  std::vector<uint32_t> order(ntask);
  for (uint32_t t = 0; t < ntask; ++t) order[t] = t;
  auto digit = [&](uint32_t node, uint32_t d) { return (node >> (2 * (pre - 1 - d))) & 3u; };  // d-th consumed base
  auto merge_key = [&](uint32_t node) {
    uint32_t key = 0;
    for (uint32_t d = 0; d < pre; ++d) {
      const uint32_t src = d == 0 ? (pre > 1 ? 1u : 0u) : (d == 1 ? 0u : d);
      key = key * 4 + digit(node, src);
    }
    return key;
  };
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b2) { return merge_key(a) < merge_key(b2); });
  std::vector<KmerOut> outs(ntask);  // outs[i] = result of task order[i]
#pragma omp parallel for schedule(dynamic, 1)
  for (int ti = 0; ti < (int)ntask; ++ti) {
    const uint32_t node = order[ti];
    if (level[node].empty()) continue;
    try {
      std::vector<uint32_t> arena(1u << 16);
      std::vector<std::array<StateList, 4>> levels(k);
      uint32_t code = 0;
      for (uint32_t d = 0; d < pre; ++d) code |= digit(node, d) << (2 * (k - 1 - d));
      KmerOut& out = outs[ti];
      if (pre == k) out.add(code, level[node]);
      else kmer_recurse(v, k, pre, code, level[node], levels, arena, out);
      level[node].clear();
    } catch (const std::exception& e) {
#pragma omp critical
      err = e.what();
    }
  }
  if (!err.empty()) throw std::runtime_error(err);
  lap("k-mer searches (DFS)");
  // merge into CSR ordered by k-mer code: every k-mer belongs to exactly one task, so the tasks write disjoint
  // ranges — counts, then (after the prefix sums) states and paths, in parallel over the tasks
  const int n_outs = (int)outs.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n_outs; ++t) {
    const KmerOut& o = outs[t];
    for (size_t i = 0; i < o.codes.size(); ++i) ix.kmer_off[o.codes[i] + 1] = o.n_states[i];
  }
  uint64_t n_states_total = 0;
  for (uint64_t c = 0; c < nk; ++c) {
    if (ix.kmer_off[c + 1]) ix.kmer_bits[c >> 5] |= 1u << (c & 31);
    n_states_total += ix.kmer_off[c + 1];
    ix.kmer_off[c + 1] += ix.kmer_off[c];
  }
  if (n_states_total >= 0xFFFFFFFFull) throw std::runtime_error("k-mer index exceeds 2^32 states; use a larger kmer_size");
  ix.kmer_states.assign(ix.kmer_off[nk], KmerState{});
  std::vector<uint64_t> path_base(outs.size() + 1, 0);  // the path pool keeps the tasks' merge order
  for (size_t t = 0; t < outs.size(); ++t) path_base[t + 1] = path_base[t] + outs[t].paths.size();
  const uint64_t total_paths = path_base.back();
  if (total_paths >= 0xFFFFFFFFull) throw std::runtime_error("k-mer index paths exceed 2^32 words");
  ix.kmer_paths.assign(total_paths, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < n_outs; ++t) {
    KmerOut& o = outs[t];
    const uint32_t base = (uint32_t)path_base[t];
    std::copy(o.paths.begin(), o.paths.end(), ix.kmer_paths.begin() + base);
    size_t si = 0;
    for (size_t i = 0; i < o.codes.size(); ++i) {
      const uint32_t dst = ix.kmer_off[o.codes[i]];
      for (uint32_t j = 0; j < o.n_states[i]; ++j, ++si) {
        KmerState ks = o.states[si];
        ks.path_off += base;
        ix.kmer_states[dst + j] = ks;
      }
    }
    o = KmerOut{};  // free the task's buffers as soon as they are merged
  }
  if (ix.kmer_paths.empty()) ix.kmer_paths.push_back(0);
  if (ix.kmer_states.empty()) ix.kmer_states.push_back(KmerState{});
  lap("k-mer CSR merge");
  build_seed_view(ix);
  lap("seed view");
}

}  // namespace

// Seed-pass view of the k-mer index (kmer_off / kmer_states must be final)
void build_seed_view(HostIndex& ix) {
  const uint32_t k = ix.k;
  const uint64_t nk = 1ull << (2 * k);
  // seed-pass view: per k-mer one entry per suffix of its narrow states (text position + left context), one
  // entry per wide state; bucketed by the first d context bases (gq_core.cuh, KmerSeed)
  const uint32_t d = seed_bucket_bases(k), B = seed_buckets(k);
  auto entry_of = [&](uint32_t i, uint32_t& bucket) {  // suffix entry of SA index i
    const uint32_t p = ix.sa[i];
    uint32_t ctx = 0, nctx = 0;
    while (nctx < kSeedCtxBases && nctx < p) {
      const uint32_t sym = ix.prg[p - 1 - nctx];
      if (sym > 4) break;
      ctx |= (sym - 1) << (22 - 2 * nctx);
      ++nctx;
    }
    bucket = nctx >= d ? 1u + (d ? ctx >> (24 - 2 * d) : 0u) : 0u;
    return KmerSeed{p, 0x80000000u | (nctx << 24) | ctx};
  };
  ix.seed_off.assign(nk * B + 1, 0);
  const int64_t n_codes = (int64_t)nk;
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t c = 0; c < n_codes; ++c) {  // bucket sizes (stored one slot ahead, prefix-summed below)
    for (uint32_t j = ix.kmer_off[c]; j < ix.kmer_off[c + 1]; ++j) {
      const KmerState& ks = ix.kmer_states[j];
      if (ks.hi - ks.lo + 1 > kSplitWidth) {
        ix.seed_off[c * B + 1]++;
        continue;
      }
      for (uint32_t i = ks.lo; i <= ks.hi; ++i) {
        uint32_t q;
        entry_of(i, q);
        ix.seed_off[c * B + q + 1]++;
      }
    }
  }
  uint64_t n_seed_total = 0;
  for (uint64_t t = 0; t < nk * B; ++t) {
    n_seed_total += ix.seed_off[t + 1];
    ix.seed_off[t + 1] += ix.seed_off[t];
  }
  if (n_seed_total >= 0xFFFFFFFFull) throw std::runtime_error("seed view exceeds 2^32 entries; use a larger kmer_size");
  ix.seed_ent.assign(std::max<size_t>(ix.seed_off[nk * B], 1), KmerSeed{0, 0});
  ix.seed_state.assign(std::max<size_t>(ix.seed_off[nk * B], 1), 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t c = 0; c < n_codes; ++c) {
    uint32_t cur[17];
    for (uint32_t q = 0; q < B; ++q) cur[q] = ix.seed_off[c * B + q];
    for (uint32_t j = ix.kmer_off[c]; j < ix.kmer_off[c + 1]; ++j) {
      const KmerState& ks = ix.kmer_states[j];
      if (ks.hi - ks.lo + 1 > kSplitWidth) {
        ix.seed_ent[cur[0]] = KmerSeed{ks.lo, ks.hi};
        ix.seed_state[cur[0]++] = j;
        continue;
      }
      for (uint32_t i = ks.lo; i <= ks.hi; ++i) {
        uint32_t q;
        const KmerSeed sd = entry_of(i, q);
        ix.seed_ent[cur[q]] = sd;
        ix.seed_state[cur[q]++] = j;
      }
    }
  }
}

void build_host_index(const uint32_t* prg, uint64_t n_symbols, uint32_t kmer_size, HostIndex& ix, SaBuilder sa_builder,
                      void* sa_ctx, const char* kmer_index_dir) {
  if (n_symbols == 0) throw std::runtime_error("empty PRG");
  // positions, SA indices and marker ranks are unsigned 32-bit words on the device, 0xFFFFFFFF is "none"
  if (n_symbols >= (1ull << 32) - 3) throw std::runtime_error("PRG too long: text positions are 32-bit words (max 2^32 - 4 symbols)");
  ix = HostIndex{};
  ix.k = kmer_size;
  ix.prg.assign(prg, prg + n_symbols);
  std::vector<uint32_t> hit_marker, hit_allele;
  const bool timing = std::getenv("GQ_BUILD_TIMING") != nullptr;  // developer aid: seconds per phase on stderr
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "index build: %-28s %.2f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  build_graph(ix, hit_marker, hit_allele);
  build_site_tables(ix);
  lap("coverage graph + site tables");
  build_fm(ix, hit_marker, hit_allele, sa_builder, sa_ctx);
  lap("SA + rank blocks + text mode");
  if (kmer_index_dir) {
    kmer_index_load(ix, kmer_index_dir);  // kmers / kmers_stats / sa_intervals / paths of a gram_dir (kmer_index::load)
    lap("k-mer index from files + seed view");
  } else {
    build_kmers(ix);
    lap("k-mer index + seed view");
  }
}

}  // namespace gq
