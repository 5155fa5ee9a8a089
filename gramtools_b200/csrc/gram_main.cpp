// gram_main.cpp — `gram genotype …`: the process seam of the reference, served by libgq.so.
//
// The Python front-end runs (gramtools/commands/genotype/genotype.py:71-93)
//   gram genotype --gram_dir D --reads F… --sample_id S --ploidy {haploid|diploid} --kmer_size K
//                 --genotype_dir G --max_threads T [--seed N] [--debug]
// parsed in the reference by libgramtools/src/main.cpp:51-100 + src/genotype/parameters.cpp:45-116.
// This executable keeps that argv contract and the geno_dir layout for the QUASIMAP part of
// commands::genotype::run (genotype.cpp:24-66): it reads gram_dir/prg (raw LE uint32), maps every reads
// file on the GPU and writes
//   G/coverage/allele_sum_coverage                 (allele_sum.cpp:45-57)
//   G/coverage/allele_base_coverage.json           (allele_base.cpp:91-107)
//   G/coverage/grouped_allele_counts_coverage.json (grouped_allele_counts.cpp:93-111)
//   G/read_stats.json                              (read_stats.cpp:162-209)
// and prints the five counters as genotype.cpp:55-66 does; then the genotyping step of the reference (LevelGenotyper,
// genotype.cpp:68-118; level_genotyper.cpp, host code) writes G/genotype/genotyped.json, personalised_reference.fasta
// and genotyped.vcf.gz (and, with --debug, G/site_gtyping_debug_info.txt).
// gram_dir/prg is what is consumed: FM-index, masks and coverage graph are rebuilt from it (their SDSL / Boost
// archives are third-party formats); the k-mer index is searched again or, with
// --kmer_index_from_gram_dir, loaded from gram_dir's kmers / kmers_stats / sa_intervals / paths; with --gq_index the
// whole index comes from gram_dir/gq_index. `gram build` writes those files.
#include <sys/stat.h>
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/gq.h"
#include "read_file.hpp"

namespace {

struct Params {
  std::string gram_dir, sample_id, ploidy, genotype_dir;
  std::vector<std::string> reads;
  uint32_t kmer_size = 0, max_threads = 1;
  bool has_seed = false, debug = false;
  uint32_t seed = 0;
  int devices = 1;  // --devices N (or GQ_DEVICES): GPUs of this node the reads are sharded over; 0 = all
  bool kmers_from_gram_dir = false;  // --kmer_index_from_gram_dir: load kmers / kmers_stats / sa_intervals / paths
  bool use_gq_index = false;         // --gq_index: load gram_dir/gq_index (the whole index, written by `gram build`)
};

[[noreturn]] void usage_fail(const std::string& msg) {
  std::cout << msg << std::endl;
  std::cout << "genotype options:\n  --gram_dir arg\n  --reads arg [arg…]\n  --sample_id arg\n"
               "  --ploidy arg {haploid, diploid}\n  --kmer_size arg\n  --genotype_dir arg\n"
               "  --max_threads arg (=1)\n  --seed arg\n  --devices arg (=1; B200 back-end: GPUs to shard the reads over, 0 = all)\n"
               "  --kmer_index_from_gram_dir (B200 back-end: load the k-mer index files of gram_dir instead of rebuilding it)\n"
               "  --gq_index (B200 back-end: load gram_dir/gq_index, the whole index as `gram build` wrote it)\n";
  std::exit(1);
}

Params parse_genotype(int argc, const char* const* argv, int first) {
  Params p;
  bool got_k = false;
  for (int i = first; i < argc; ++i) {
    std::string a = argv[i];
    auto value = [&](const std::string& name) -> std::string {
      if (i + 1 >= argc) usage_fail("the required argument for option '--" + name + "' is missing");
      return argv[++i];
    };
    if (a == "--gram_dir") p.gram_dir = value("gram_dir");
    else if (a == "--reads") {
      while (i + 1 < argc && std::strncmp(argv[i + 1], "--", 2) != 0) p.reads.push_back(argv[++i]);
    } else if (a == "--sample_id") p.sample_id = value("sample_id");
    else if (a == "--ploidy") p.ploidy = value("ploidy");
    else if (a == "--kmer_size") {
      p.kmer_size = (uint32_t)std::stoul(value("kmer_size"));
      got_k = true;
    } else if (a == "--genotype_dir") p.genotype_dir = value("genotype_dir");
    else if (a == "--max_threads") p.max_threads = (uint32_t)std::stoul(value("max_threads"));
    else if (a == "--seed") {
      p.seed = (uint32_t)std::stoul(value("seed"));
      p.has_seed = true;
    } else if (a == "--devices") p.devices = std::stoi(value("devices"));
    else if (a == "--kmer_index_from_gram_dir") p.kmers_from_gram_dir = true;
    else if (a == "--gq_index") p.use_gq_index = true;
    else if (a == "--debug") p.debug = true;
    else usage_fail("unrecognised option '" + a + "'");
  }
  if (p.gram_dir.empty()) usage_fail("the option '--gram_dir' is required but missing");
  if (p.reads.empty()) usage_fail("the option '--reads' is required but missing");
  if (p.sample_id.empty()) usage_fail("the option '--sample_id' is required but missing");
  if (p.ploidy != "haploid" && p.ploidy != "diploid") usage_fail("the option '--ploidy' must be haploid or diploid");
  if (!got_k) usage_fail("the option '--kmer_size' is required but missing");
  if (p.genotype_dir.empty()) usage_fail("the option '--genotype_dir' is required but missing");
  return p;
}

void make_dir(const std::string& d) {
  if (mkdir(d.c_str(), 0775) != 0 && errno != EEXIST) {
    std::cout << "Cannot create directory " << d << std::endl;
    std::exit(1);
  }
}

// PRG_String(file) (linearised_prg.cpp:8-45): little-endian uint32 stream
std::vector<uint32_t> read_prg(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) {
    std::cout << "PRG String file not found: " << path << std::endl;
    std::exit(1);
  }
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint32_t> prg((size_t)(bytes > 0 ? bytes : 0) / 4);  // a trailing partial word is ignored
  const size_t got = prg.empty() ? 0 : std::fread(prg.data(), 4, prg.size(), f);  // one read, not one per symbol
  std::fclose(f);
  prg.resize(got);
  const uint32_t probe = 1;
  if (*(const unsigned char*)&probe != 1)  // big-endian host: the file is little-endian
    for (auto& w : prg) w = (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
  return prg;
}

// sequence reader (read_file.hpp): SeqRead / seq_file.h behaviour, lines as views into one inflate buffer
std::unique_ptr<gq::ReadFile> open_reads(const std::string& path) {
  try {
    return std::make_unique<gq::ReadFile>(path);
  } catch (const std::exception& e) {
    std::cout << e.what() << std::endl;
    std::exit(1);
  }
}

inline uint8_t encode_base(char c) {  // encode_char, utils.cpp:13-47
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    default: return 0;
  }
}

void check(int rc) {
  if (rc != 0) {
    std::cout << "libgq error: " << gq_last_error() << std::endl;
    std::exit(1);
  }
}

std::string join_path(const std::string& a, const std::string& b) { return a + (a.empty() || a.back() == '/' ? "" : "/") + b; }

// ---- ingestion pipeline (replaces SeqRead + the 5000-read buffer, seqread.hpp:94-180 / quasimap.cpp:126-140) ----
// reader thread: parses the read files, concatenates sequence text, draws the per-read seeds in file order;
// one worker thread per GPU: packs a batch to 2 bits per base straight from the text (gq_pack_ascii, OpenMP) into
// pinned buffers and maps it (gq_map_batch_packed). Batches are handed round-robin to whichever worker is free, so
// parsing, packing, the H2D copy and the kernels of different batches overlap; results do not depend on which GPU
// maps which batch (seeds travel with the reads, coverage is summed at the end).
struct Batch {
  std::string text;               // sequence characters of the batch's reads, concatenated
  std::vector<uint64_t> offs{0};  // n + 1
  std::vector<uint32_t> seeds;    // n
  void clear() {
    text.clear();
    offs.assign(1, 0);
    seeds.clear();
  }
};

class BatchQueue {
 public:
  explicit BatchQueue(size_t depth) : depth_(depth) {}
  void push(std::unique_ptr<Batch> b) {
    std::unique_lock<std::mutex> lk(m_);
    not_full_.wait(lk, [&] { return q_.size() < depth_; });
    q_.push_back(std::move(b));
    not_empty_.notify_one();
  }
  std::unique_ptr<Batch> pop() {  // nullptr: the reader is done and the queue is drained
    std::unique_lock<std::mutex> lk(m_);
    not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return nullptr;
    auto b = std::move(q_.front());
    q_.pop_front();
    not_full_.notify_one();
    return b;
  }
  void close() {
    std::lock_guard<std::mutex> lk(m_);
    closed_ = true;
    not_empty_.notify_all();
  }

 private:
  std::mutex m_;
  std::condition_variable not_full_, not_empty_;
  std::deque<std::unique_ptr<Batch>> q_;
  size_t depth_;
  bool closed_ = false;
};

// pinned packed buffers of one worker, grown on demand
struct PinnedBatch {
  uint32_t *packed = nullptr, *word_off = nullptr, *len = nullptr, *seeds = nullptr;
  uint64_t cap_words = 0, cap_reads = 0;
  void reserve(uint64_t words, uint64_t reads) {
    if (words > cap_words) {
      gq_host_free(packed);
      check(gq_host_alloc(words * 4, (void**)&packed));
      cap_words = words;
    }
    if (reads > cap_reads) {
      gq_host_free(word_off);
      gq_host_free(len);
      gq_host_free(seeds);
      check(gq_host_alloc((reads + 1) * 4, (void**)&word_off));
      check(gq_host_alloc(reads * 4 + 4, (void**)&len));
      check(gq_host_alloc(reads * 4 + 4, (void**)&seeds));
      cap_reads = reads;
    }
  }
  ~PinnedBatch() {
    gq_host_free(packed);
    gq_host_free(word_off);
    gq_host_free(len);
    gq_host_free(seeds);
  }
};

}  // namespace

// `gram build --gram_dir D --kmer_size K`: the k-mer index part of commands::build::run (build.cpp:8-71,
// kmer_index::build + dump): reads D/prg (as written by the Python front-end / an earlier reference build), builds the
// index on GPU 0 and leaves D/kmers, D/kmers_stats, D/sa_intervals, D/paths in the reference's sdsl format, plus
// D/gq_index: the whole flat index in this back-end's own format (`gram genotype --gq_index` then rebuilds nothing).
// The other files of a reference gram_dir (fm_index, masks, cov_graph: SDSL / Boost archives) are not written. Options the reference's build takes and this one does not need
// (--ref, --prg, --max_threads, --all_kmers, ...) are accepted and ignored.
int run_build(int argc, const char* const* argv, int first) {
  std::string gram_dir;
  uint32_t kmer_size = 0;
  bool got_k = false;
  for (int i = first; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--gram_dir" && i + 1 < argc) gram_dir = argv[++i];
    else if (a == "--kmer_size" && i + 1 < argc) {
      kmer_size = (uint32_t)std::stoul(argv[++i]);
      got_k = true;
    } else if (a.rfind("--", 0) == 0 && i + 1 < argc && std::strncmp(argv[i + 1], "--", 2) != 0) ++i;  // ignored option + value
  }
  if (gram_dir.empty() || !got_k) {
    std::cout << "build options:\n  --gram_dir arg\n  --kmer_size arg\n";
    return 1;
  }
  std::cout << "Executing build command" << std::endl;
  std::vector<uint32_t> prg = read_prg(join_path(gram_dir, "prg"));
  std::cout << "Loaded PRG: " << prg.size() << " symbols" << std::endl;
  gq_index* idx = nullptr;
  check(gq_index_build(prg.data(), prg.size(), kmer_size, 0, &idx));
  check(gq_kmer_index_dump(idx, gram_dir.c_str()));
  check(gq_index_save(idx, join_path(gram_dir, "gq_index").c_str()));  // the whole index, for `gram genotype --gq_index`
  gq_layout lay;
  check(gq_index_describe(idx, &lay));
  std::cout << "Indexed kmers search states: " << lay.n_kmer_states << std::endl;
  check(gq_index_destroy(idx));
  return 0;
}

int main(int argc, const char* const* argv) {
  // main.cpp:51-100: `gram` alone prints the global help and exits 0 (gramtools_main.py:80-90 relies on it)
  std::string cmd;
  int first = 1;
  bool debug_flag = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--help") break;
    if (a == "--debug") {
      debug_flag = true;
      continue;
    }
    cmd = a;
    first = i + 1;
    break;
  }
  if (cmd.empty()) {
    std::cout << "Gramtools! Global options:\n  --command arg   command to execute: {build, genotype, simulate}\n"
                 "  --help          Produce this help message\n  --debug         Turn on debug output\n"
                 "(B200 quasimap back-end: `genotype` and `build` (k-mer index files) are served)\n";
    return 0;
  }
  if (cmd == "build") return run_build(argc, argv, first);
  if (cmd != "genotype") {
    std::cout << (cmd == "simulate" ? "Command not served by the B200 quasimap back-end: " : "Unrecognised command: ")
              << cmd << std::endl;
    return 1;
  }
  Params p = parse_genotype(argc, argv, first);
  p.debug = p.debug || debug_flag;

  std::cout << "Executing genotype command" << std::endl;
  const std::string cov_dir = join_path(p.genotype_dir, "coverage");
  make_dir(p.genotype_dir);
  make_dir(cov_dir);
  make_dir(join_path(p.genotype_dir, "genotype"));

  // ReadStats::compute_base_error_rate on the first reads file (genotype.cpp:32-34, read_stats.cpp:21-70)
  uint64_t max_read_length = 0, num_bases = 0, no_qual_reads = 0, informative = 0;
  float running_qual = 0.0f;  // a float in the reference as well (read_stats.cpp:33): the rounding of its sum is kept
  {
    auto rf = open_reads(p.reads[0]);
    std::string seq, qual;
    while (informative < 10000 && rf->next(seq, qual)) {
      if (seq.size() > max_read_length) max_read_length = seq.size();
      if (qual.empty()) {
        ++no_qual_reads;
        continue;
      }
      for (char q : qual) running_qual += (q - 33);
      num_bases += qual.size();
      ++informative;
    }
  }
  double mean_pb_error = 0.0;
  if (num_bases) {  // read_stats.cpp:62-65: float / int64 is a float division
    const double mean_qual = running_qual / (int64_t)num_bases;
    mean_pb_error = std::pow(10, -mean_qual / 10);
  }

  std::cout << "Loading PRG data" << std::endl;
  std::vector<uint32_t> prg = read_prg(join_path(p.gram_dir, "prg"));
  int n_dev = p.devices;
  if (const char* e = std::getenv("GQ_DEVICES")) n_dev = std::atoi(e);
  int have = 0;
  check(gq_device_count(&have));
  if (n_dev <= 0 || n_dev > have) n_dev = have > 0 ? (n_dev <= 0 ? have : std::min(n_dev, have)) : 1;
  std::vector<gq_index*> handles(n_dev, nullptr);
  // host index built once ... (k-mer index searched, or — as kmer_index::load does, genotype.cpp:40 — taken from the
  // sdsl files `gram build` left in gram_dir)
  if (p.use_gq_index) {
    // nothing is rebuilt; the stored index must be the one of this gram_dir/prg and this kmer_size
    check(gq_index_load(join_path(p.gram_dir, "gq_index").c_str(), 0, &handles[0]));
    gq_layout l0;
    check(gq_index_describe(handles[0], &l0));
    uint64_t n_stored = 0;
    check(gq_index_prg(handles[0], nullptr, &n_stored));
    std::vector<uint32_t> stored(n_stored ? n_stored : 1);
    check(gq_index_prg(handles[0], stored.data(), &n_stored));
    stored.resize(n_stored);
    if (l0.kmer_size != p.kmer_size || stored != prg) {
      std::cout << "gram_dir/gq_index was built from another PRG or kmer_size: run `gram build` again" << std::endl;
      return 1;
    }
  } else if (p.kmers_from_gram_dir || std::getenv("GQ_KMER_INDEX_FROM_GRAM_DIR"))
    check(gq_index_build_from_gram_dir(prg.data(), prg.size(), p.kmer_size, 0, p.gram_dir.c_str(), &handles[0]));
  else
    check(gq_index_build(prg.data(), prg.size(), p.kmer_size, 0, &handles[0]));
  for (int d = 1; d < n_dev; ++d) check(gq_index_clone(handles[0], d, &handles[d]));  // ... uploaded per GPU
  if (n_dev > 1) check(gq_comm_init_all(handles.data(), n_dev));
  gq_index* idx = handles[0];
  gq_layout lay;
  check(gq_index_describe(idx, &lay));

  std::cout << "Running quasimap" << std::endl;
  uint32_t master_seed = p.has_seed ? p.seed : std::random_device{}();
  std::cout << "Master random seed for read selection: " << master_seed << std::endl;
  std::cout << "Maximum thread count: " << p.max_threads << " (host threads packing reads; reads are mapped on " << n_dev
            << " GPU" << (n_dev > 1 ? "s" : "") << ")" << std::endl;
  std::cout << "Processing reads:" << std::endl;
  const uint64_t kRefBatch = 5000;       // quasimap.cpp:126-128: seeds are drawn 5000 at a time
  uint64_t kGpuBatch = 1u << 20;         // reads per gq_map_batch_packed call (GQ_BATCH_READS overrides: tests)
  if (const char* e = std::getenv("GQ_BATCH_READS")) kGpuBatch = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
  BatchQueue queue(2 * (size_t)n_dev);
  std::atomic<uint64_t> total_reads{0};
  std::mutex print_m;
  const int pack_threads = std::max<int>(1, (int)p.max_threads / n_dev);
  std::vector<std::thread> workers;
  for (int d = 0; d < n_dev; ++d)
    workers.emplace_back([&, d] {
      PinnedBatch pin;
      while (auto b = queue.pop()) {
        const uint64_t n = b->seeds.size();
        uint64_t words = 0;
        check(gq_packed_words(b->offs.data(), n, &words));
        pin.reserve(words, n);
        check(gq_pack_ascii(b->text.data(), b->offs.data(), n, pin.packed, pin.word_off, pin.len, pack_threads));
        std::memcpy(pin.seeds, b->seeds.data(), n * 4);
        check(gq_map_batch_packed(handles[d], pin.packed, pin.word_off, pin.len, n, pin.seeds));
        const uint64_t t = total_reads.fetch_add(n) + n;
        std::lock_guard<std::mutex> lk(print_m);
        std::cout << 2 * t << std::endl;
      }
    });
  {  // reader (this thread)
    std::mt19937 master;
    master.seed(master_seed);
    for (const auto& path : p.reads) {
      auto rf = open_reads(path);
      auto cur = std::make_unique<Batch>();
      size_t seq_len = 0;
      uint64_t in_ref_batch = 0;
      // the record's sequence goes straight from the inflate buffer into the batch text (qualities are only counted)
      while (rf->next_seq(cur->text, seq_len)) {
        // read j of a 5000-read buffer gets the j-th of the 5000 draws made for that buffer
        // (quasimap.cpp:132-139): unused draws of the last, partly filled buffer are discarded
        if (in_ref_batch == kRefBatch) in_ref_batch = 0;
        cur->seeds.push_back((uint32_t)master());
        ++in_ref_batch;
        cur->offs.push_back(cur->text.size());  // non-ACGT characters empty the read in the packer (utils.cpp:72-92)
        if (cur->seeds.size() == kGpuBatch) {
          queue.push(std::move(cur));
          cur = std::make_unique<Batch>();
        }
      }
      if (!cur->seeds.empty()) queue.push(std::move(cur));
      if (in_ref_batch) master.discard(kRefBatch - in_ref_batch);
    }
    queue.close();
  }
  for (auto& w : workers) w.join();
  if (n_dev > 1) check(gq_coverage_allreduce_all(handles.data(), n_dev));  // the one exchange: every handle holds the totals

  // ---- outputs -----------------------------------------------------------------------------
  std::vector<uint16_t> allele_sum(lay.n_alleles ? lay.n_alleles : 1), per_base(lay.n_per_base ? lay.n_per_base : 1);
  uint64_t st[5];
  check(gq_coverage_fetch(idx, allele_sum.data(), per_base.data(), st));
  std::vector<uint64_t> allele_off(lay.n_site_slots + 1);
  check(gq_index_allele_offsets(idx, allele_off.data()));
  {  // allele_sum.cpp:45-57
    std::ofstream f(join_path(cov_dir, "allele_sum_coverage"));
    for (uint32_t s = 0; s < lay.n_site_slots; ++s) {
      for (uint64_t a = allele_off[s]; a < allele_off[s + 1]; ++a) f << allele_sum[a] << (a + 1 < allele_off[s + 1] ? " " : "");
      f << "\n";  // (no flush per site)
    }
  }
  {  // allele_base.cpp:91-107; empty by convention for nested PRGs (:10-14)
    std::ofstream f(join_path(cov_dir, "allele_base_coverage.json"));
    f << "{\"allele_base_counts\":[";
    if (!lay.is_nested) {
      std::vector<uint64_t> ol(2 * (lay.n_alleles ? lay.n_alleles : 1));
      check(gq_index_per_base_layout(idx, ol.data()));
      for (uint32_t s = 0; s < lay.n_site_slots; ++s) {
        f << "[";
        for (uint64_t a = allele_off[s]; a < allele_off[s + 1]; ++a) {
          f << "[";
          for (uint64_t i = 0; i < ol[2 * a + 1]; ++i) f << (int)per_base[ol[2 * a] + i] << (i + 1 < ol[2 * a + 1] ? "," : "");
          f << "]" << (a + 1 < allele_off[s + 1] ? "," : "");
        }
        f << "]" << (s + 1 < lay.n_site_slots ? "," : "");
      }
    }
    f << "]}" << std::endl;
  }
  std::vector<uint32_t> grouped_records;  // kept for the genotyper
  {  // grouped_allele_counts.cpp:51-111 (group ids numbered by first appearance, sites in order)
    uint64_t nw = 0;
    check(gq_coverage_grouped(idx, nullptr, &nw));
    std::vector<uint32_t> w(nw ? nw : 1);
    check(gq_coverage_grouped(idx, w.data(), &nw));
    grouped_records.assign(w.begin(), w.begin() + nw);
    std::map<std::vector<uint32_t>, uint64_t> group_id;
    std::vector<std::map<std::string, uint32_t>> site_counts(lay.n_site_slots);
    for (uint64_t t = 0; t < nw;) {
      uint32_t n = w[t + 2];
      std::vector<uint32_t> ids(w.begin() + t + 3, w.begin() + t + 3 + n);
      auto it = group_id.find(ids);
      if (it == group_id.end()) it = group_id.insert({ids, group_id.size()}).first;
      site_counts[w[t]][std::to_string(it->second)] = w[t + 1];
      t += 3 + n;
    }
    std::map<std::string, std::vector<uint32_t>> groups_by_name;
    for (auto& e : group_id) groups_by_name[std::to_string(e.second)] = e.first;
    std::ofstream f(join_path(cov_dir, "grouped_allele_counts_coverage.json"));
    f << "{\"grouped_allele_counts\":{\"allele_groups\":{";
    bool first_g = true;
    for (auto& e : groups_by_name) {
      f << (first_g ? "" : ",") << "\"" << e.first << "\":[";
      for (size_t i = 0; i < e.second.size(); ++i) f << e.second[i] << (i + 1 < e.second.size() ? "," : "");
      f << "]";
      first_g = false;
    }
    f << "},\"site_counts\":[";
    for (uint32_t s = 0; s < lay.n_site_slots; ++s) {
      f << "{";
      bool first_c = true;
      for (auto& e : site_counts[s]) {
        f << (first_c ? "" : ",") << "\"" << e.first << "\":" << e.second;
        first_c = false;
      }
      f << "}" << (s + 1 < lay.n_site_slots ? "," : "");
    }
    f << "]}}" << std::endl;
  }
  double depth[2];
  {  // read_stats.cpp:162-209
    uint64_t cnt[2];
    check(gq_read_depth_stats(idx, depth, cnt));
    const std::string path = join_path(p.genotype_dir, "read_stats.json");
    std::cout << "Writing read stats to " << path << std::endl;
    std::ofstream f(path);
    f << "\n{\n\"Read_depth\":\n    {\"Mean\": " << depth[0] << ",\n    \"Variance\": " << depth[1]
      << ",\n    \"num_sites_noCov\": " << cnt[0] << ",\n    \"num_sites_total\": " << cnt[1] << "\n    },\n"
      << "\"Max_read_length\": " << max_read_length << ",\n\"Quality\":\n    {\"Error_rate_mean\": " << mean_pb_error
      << ",\n    \"Num_bases\": " << num_bases << ",\n    \"No_qual_reads\": " << no_qual_reads << "\n    }}\n";
  }
  std::cout << std::endl << "The following counts include generated reverse complement reads." << std::endl;
  std::cout << "Count all reads: " << st[0] << std::endl;
  std::cout << "Count skipped reads with no sequence: " << st[1] << std::endl;
  std::cout << "Count reads with >0 kmers not in kmer index: " << st[2] << std::endl;
  std::cout << "Count reads with no exact mapping: " << st[3] << std::endl;
  std::cout << "Count exact mapped reads: " << st[4] << std::endl;
  for (gq_index* h : handles) gq_index_destroy(h);  // the device is not needed any more
  handles.clear();

  // ---- genotyping (genotype.cpp:68-118): LevelGenotyper on the fetched coverage, host code of libgq.so ----------
  std::cout << "====================" << std::endl << "Running genotyping" << std::endl;
  const std::string debug_file = p.debug ? join_path(p.genotype_dir, "site_gtyping_debug_info.txt") : std::string();
  if (p.debug) std::cout << "Logging debug genotyping stats to " << debug_file << std::endl;
  std::cout << "Running genotyping model" << std::endl;
  {
    const uint64_t nw = grouped_records.size();
    const double stats3[3] = {depth[0], depth[1], mean_pb_error};
    std::cout << "Producing json vcf, personalised reference, vcf" << std::endl;
    check(gq_level_genotype(prg.data(), prg.size(), lay.n_per_base ? per_base.data() : nullptr, lay.n_per_base,
                            nw ? grouped_records.data() : nullptr, nw, stats3, p.ploidy == "haploid" ? 1 : 2,
                            p.sample_id.c_str(), join_path(p.gram_dir, "prg_coords.tsv").c_str(),
                            join_path(p.genotype_dir, "genotype").c_str(), p.debug ? debug_file.c_str() : nullptr, 42,
                            (int)p.max_threads));
  }
  return 0;
}
