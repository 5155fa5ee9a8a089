// genotype_capi.cpp — C entry points of the genotyping step (level_genotyper.cpp): host code, no device needed.
#include <cstdint>
#include <fstream>
#include <future>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../../include/gq.h"
#include "level_genotyper.hpp"

void set_last_error(const std::string& what);  // capi.cu

namespace {

std::string slurp(const char* path) {
  if (!path || !*path) return std::string();
  std::ifstream f(path, std::ios::binary);
  if (!f) return std::string();  // a missing prg_coords.tsv means one segment (segment_tracker.hpp:32-35)
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

void spit(const std::string& path, const std::string& text) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  f.write(text.data(), (std::streamsize)text.size());
  if (!f) throw std::runtime_error("cannot write " + path);
}

gq::lg::LevelGenotyper run(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                           const uint32_t* grouped, uint64_t n_grouped_words, const double stats[3], int ploidy,
                           uint32_t gcp_seed, bool debug, int n_threads) {
  if (!prg || !stats || (n_grouped_words && !grouped)) throw std::runtime_error("null argument");
  if (ploidy != 1 && ploidy != 2) throw std::runtime_error("ploidy must be 1 (haploid) or 2 (diploid)");
  gq::lg::PrgSites ps = gq::lg::parse_prg_sites(prg, n_symbols);
  uint64_t in_site = 0;
  for (auto& s : ps.sites)
    if (s.parent < 0) in_site += s.pb_exit - s.pb_entry;
  if (in_site != n_per_base)
    throw std::runtime_error("per-base vector has " + std::to_string(n_per_base) + " entries, the PRG has " +
                             std::to_string(in_site) + " bases inside sites");
  if (n_per_base && !per_base) throw std::runtime_error("null argument");
  if (!ps.sites.empty() && (!(stats[0] >= 0) || !(stats[1] >= 0) || !(stats[2] >= 0 && stats[2] <= 1)))
    throw std::runtime_error("stats must be {mean coverage >= 0, coverage variance >= 0, error rate in [0, 1]}");
  gq::lg::RunOptions opt;
  opt.ploidy = ploidy == 1 ? gq::lg::Ploidy::Haploid : gq::lg::Ploidy::Diploid;
  opt.gcp_seed = gcp_seed;
  opt.debug = debug;
  opt.n_threads = n_threads;
  return gq::lg::LevelGenotyper(std::move(ps), per_base, grouped, n_grouped_words, stats[0], stats[1], stats[2], opt);
}

}  // namespace

extern "C" {

int gq_level_genotype(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                      const uint32_t* grouped, uint64_t n_grouped_words, const double stats[3], int ploidy,
                      const char* sample_id, const char* prg_coords_path, const char* genotype_dir,
                      const char* debug_path, uint32_t gcp_seed, int n_threads) {
  try {
    if (!sample_id || !genotype_dir) throw std::runtime_error("null argument");
    const bool debug = debug_path && *debug_path;
    gq::lg::LevelGenotyper g =
        run(prg, n_symbols, per_base, n_per_base, grouped, n_grouped_words, stats, ploidy, gcp_seed, debug, n_threads);
    const std::string dir = std::string(genotype_dir) + "/", coords = slurp(prg_coords_path), sample = sample_id;
    // the three files are independent of one another (each walks the sites with its own segment tracker)
    const auto policy = n_threads == 1 ? std::launch::deferred : std::launch::async;
    auto json = std::async(policy, [&] {  // genotype.cpp:92-98
      gq::lg::SegmentTracker tracker(coords);
      spit(dir + "genotyped.json", g.json(sample, tracker) + "\n");
    });
    auto fasta = std::async(policy, [&] {  // genotype.cpp:100-108
      gq::lg::SegmentTracker tracker(coords);
      spit(dir + "personalised_reference.fasta",
           gq::lg::deduped_fasta_text(g.personalised_reference(tracker),
                                      sample + " personalised reference made by gramtools genotype"));
    });
    auto vcf = std::async(policy, [&] {  // genotype.cpp:110-112
      gq::lg::SegmentTracker tracker(coords);
      spit(dir + "genotyped.vcf.gz", gq::lg::bgzf_compress(g.vcf(sample, tracker)));
    });
    std::string failure;
    for (auto* f : {&json, &fasta, &vcf}) {
      try {
        f->get();
      } catch (const std::exception& e) {
        if (failure.empty()) failure = e.what();
      }
    }
    if (!failure.empty()) throw std::runtime_error(failure);
    if (debug) {
      std::ofstream f(debug_path, std::ios::app);  // runner.cpp:45-50: appended
      f << g.debug_text();
    }
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return -1;
  }
  return 0;
}

int gq_read_depth_stats_host(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                             const uint32_t* grouped, uint64_t n_grouped_words, double out[2], uint64_t counts[2]) {
  try {
    if (!prg || !out || !counts || (n_grouped_words && !grouped) || (n_per_base && !per_base))
      throw std::runtime_error("null argument");
    gq::lg::PrgSites ps = gq::lg::parse_prg_sites(prg, n_symbols);
    static const uint16_t none = 0;
    const auto d = gq::lg::read_depth_stats(ps, per_base ? per_base : &none,
                                            gq::lg::unpack_grouped(grouped, n_grouped_words, ps.sites.size()));
    out[0] = d.mean;
    out[1] = d.variance;
    counts[0] = d.num_sites_no_cov;
    counts[1] = d.num_sites_total;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return -1;
  }
  return 0;
}

int gq_level_genotype_json(const uint32_t* prg, uint64_t n_symbols, const uint16_t* per_base, uint64_t n_per_base,
                           const uint32_t* grouped, uint64_t n_grouped_words, const double stats[3], int ploidy,
                           const char* sample_id, uint32_t gcp_seed, int n_threads, char* json_out,
                           uint64_t* json_bytes) {
  try {
    if (!sample_id || !json_bytes) throw std::runtime_error("null argument");
    gq::lg::LevelGenotyper g =
        run(prg, n_symbols, per_base, n_per_base, grouped, n_grouped_words, stats, ploidy, gcp_seed, false, n_threads);
    gq::lg::SegmentTracker tracker;
    const std::string text = g.json(sample_id, tracker);
    if (json_out) {
      if (*json_bytes < text.size() + 1) throw std::runtime_error("json_out is too small");
      std::copy(text.begin(), text.end(), json_out);
      json_out[text.size()] = 0;
    }
    *json_bytes = text.size() + 1;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return -1;
  }
  return 0;
}

}  // extern "C"
