// Developer probe: phase timings of the genotyping step (profiles/r02_genotype_step.txt).
// g++ -O2 -fopenmp -std=c++17 -o /tmp/probe tools/genotype_step_probe.cpp gramtools_b200/csrc/level_genotyper.cpp -lz
#include <chrono>
#include <cstdio>
#include <random>
#include "../gramtools_b200/csrc/level_genotyper.hpp"
using namespace gq::lg;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  std::mt19937 g(1);
  std::vector<uint32_t> prg;
  const uint32_t S = 100000;
  for (uint32_t s = 0; s < S; ++s) {
    for (int i = 0; i < 40; ++i) prg.push_back(1 + g() % 4);
    prg.push_back(5 + 2 * s); prg.push_back(1); prg.push_back(6 + 2 * s); prg.push_back(2); prg.push_back(6 + 2 * s);
  }
  std::vector<Cov> pb(2 * S);
  std::vector<uint32_t> grouped;
  for (uint32_t s = 0; s < S; ++s) {
    uint32_t a = 5 + g() % 12, b = g() % 3;
    pb[2 * s] = a; pb[2 * s + 1] = b;
    grouped.insert(grouped.end(), {s, a, 1, 0});
    if (b) grouped.insert(grouped.end(), {s, b, 1, 1});
  }
  double t = now();
  PrgSites ps = parse_prg_sites(prg.data(), prg.size());
  std::printf("parse %.3f\n", now() - t); t = now();
  RunOptions opt; opt.with_percentiles = false;
  LevelGenotyper lg(ps, pb.data(), grouped.data(), grouped.size(), 10.2, 10.3, 1e-3, opt);
  std::printf("genotype (no gcp) %.3f\n", now() - t); t = now();
  opt.with_percentiles = true;
  LevelGenotyper lg2(ps, pb.data(), grouped.data(), grouped.size(), 10.2, 10.3, 1e-3, opt);
  std::printf("genotype (gcp) %.3f\n", now() - t); t = now();
  for (int nt : {2, 4, 8}) {
    opt.n_threads = nt;
    t = now();
    LevelGenotyper lgp(ps, pb.data(), grouped.data(), grouped.size(), 10.2, 10.3, 1e-3, opt);
    double dt = now() - t;
    SegmentTracker t1, t2;
    std::printf("genotype (gcp) %d threads %.3f same=%d\n", nt, dt, (int)(lgp.json("s", t1) == lg2.json("s", t2)));
  }
  t = now();
  SegmentTracker tr;
  std::string j = lg2.json("s", tr);
  std::printf("json %.3f (%zu bytes)\n", now() - t, j.size()); t = now();
  tr.reset();
  auto refs = lg2.personalised_reference(tr);
  std::string f = deduped_fasta_text(refs, "d");
  std::printf("fasta %.3f\n", now() - t); t = now();
  tr.reset();
  std::string v = lg2.vcf("s", tr);
  std::printf("vcf %.3f\n", now() - t); t = now();
  std::string z = bgzf_compress(v);
  std::printf("bgzf %.3f\n", now() - t);
}
