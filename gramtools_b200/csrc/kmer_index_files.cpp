// kmer_index_files.cpp — the k-mer index as the reference's gram_dir files (see index_build.hpp). Host-only.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "index_build.hpp"

namespace gq {

// ---- the k-mer index as gram_dir files (reference: src/build/kmer_index/dump.cpp:27-141, load.cpp:11-173) -----------
// Four sdsl::int_vector files. Serialised form of sdsl-lite 2.1.1 (int_vector::serialize): the length in BITS as a
// little-endian uint64; for int_vector<0> (run-time width) one byte with the width; then the elements, `width` bits
// each, packed LSB first into 64-bit words, the last word zero-padded.
//   kmers         int_vector<3>  k base codes (1..4) per indexed k-mer, k-mers in any order (the reference writes its
//                                unordered_map's order and reads the file sequentially)
//   kmers_stats   int_vector<>   per k-mer: number of SearchStates, then the path length (traversed + traversing
//                                loci) of each; bit-compressed (width = bits of the largest element, at least 1)
//   sa_intervals  int_vector<>   per SearchState: first, last SA index
//   paths         int_vector<>   per SearchState: (site marker, allele id + 1) per traversed locus, then (site, 0) per
//                                traversing one (ALLELE_UNKNOWN = -1 is shifted to 0, dump.cpp:104-107)
// PARITY UNPINNED: SDSL is not available here and the reference ships no serialised fixture, so these files are
// checked against the format as documented above (golden bytes written by hand, tests/test_sdsl_io.py) and by
// round trips, not against files written by sdsl itself.
void write_int_vector(const std::string& path, const std::vector<uint64_t>& values, uint32_t width, bool fixed_width) {
  if (width == 0 || width > 64) throw std::runtime_error("int_vector width must be in [1,64]");
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  const uint64_t bits = (uint64_t)values.size() * width;
  std::vector<uint64_t> words((bits + 63) / 64, 0);
  uint64_t pos = 0;
  for (uint64_t x : values) {
    if (width < 64 && (x >> width)) {
      fclose(f);
      throw std::runtime_error("int_vector element does not fit its width");
    }
    words[pos >> 6] |= x << (pos & 63);
    if ((pos & 63) + width > 64) words[(pos >> 6) + 1] |= x >> (64 - (pos & 63));
    pos += width;
  }
  bool ok = fwrite(&bits, 8, 1, f) == 1;
  if (!fixed_width) {
    const uint8_t w8 = (uint8_t)width;
    ok = ok && fwrite(&w8, 1, 1, f) == 1;
  }
  if (!words.empty()) ok = ok && fwrite(words.data(), 8, words.size(), f) == words.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw std::runtime_error("write error on " + path);
}

std::vector<uint64_t> read_int_vector(const std::string& path, uint32_t fixed_width, uint32_t* width_out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot read " + path);
  uint64_t bits = 0;
  uint32_t width = fixed_width;
  bool ok = fread(&bits, 8, 1, f) == 1;
  if (ok && fixed_width == 0) {
    uint8_t w8 = 0;
    ok = fread(&w8, 1, 1, f) == 1;
    width = w8;
  }
  if (!ok || width == 0 || width > 64 || bits % width != 0) {
    fclose(f);
    throw std::runtime_error("not an sdsl int_vector: " + path);
  }
  std::vector<uint64_t> words((bits + 63) / 64, 0);
  if (!words.empty() && fread(words.data(), 8, words.size(), f) != words.size()) {
    fclose(f);
    throw std::runtime_error("truncated int_vector: " + path);
  }
  fclose(f);
  std::vector<uint64_t> values(bits / width);
  const uint64_t mask = width == 64 ? ~0ull : ((1ull << width) - 1);
  uint64_t pos = 0;
  for (auto& x : values) {
    uint64_t v = words[pos >> 6] >> (pos & 63);
    if ((pos & 63) + width > 64) v |= words[(pos >> 6) + 1] << (64 - (pos & 63));
    x = v & mask;
    pos += width;
  }
  if (width_out) *width_out = width;
  return values;
}

static uint32_t compressed_width(const std::vector<uint64_t>& v) {  // sdsl::util::bit_compress: hi(max) + 1
  uint64_t mx = 0;
  for (uint64_t x : v) mx = std::max(mx, x);
  uint32_t w = 1;
  while (w < 64 && (mx >> w)) ++w;
  return w;
}

void kmer_index_dump(const HostIndex& ix, const std::string& dir) {
  const uint32_t k = ix.k;
  const uint64_t nk = 1ull << (2 * k);
  std::vector<uint64_t> kmers, stats, sa_iv, paths;
  for (uint64_t code = 0; code < nk; ++code) {
    const uint32_t b = ix.kmer_off[code], e = ix.kmer_off[code + 1];
    if (b == e) continue;  // the index holds the k-mers with at least one SearchState
    for (uint32_t j = 0; j < k; ++j) kmers.push_back(((code >> (2 * j)) & 3u) + 1);  // base j of the k-mer, 1..4
    stats.push_back(e - b);
    for (uint32_t st = b; st < e; ++st) {
      const KmerState& ks = ix.kmer_states[st];
      const uint32_t nt = ks.counts & 0xFFFFu, ng = ks.counts >> 16;
      stats.push_back(nt + ng);
      sa_iv.push_back(ks.lo);
      sa_iv.push_back(ks.hi);
      const uint32_t* p = ix.kmer_paths.data() + ks.path_off;
      for (uint32_t t = 0; t < nt; ++t) {
        paths.push_back(p[2 * t]);
        paths.push_back((uint64_t)p[2 * t + 1] + 1);
      }
      for (uint32_t g = 0; g < ng; ++g) {
        paths.push_back(p[2 * nt + g]);
        paths.push_back(0);
      }
    }
  }
  write_int_vector(dir + "/kmers", kmers, 3, true);
  write_int_vector(dir + "/kmers_stats", stats, compressed_width(stats), false);
  write_int_vector(dir + "/sa_intervals", sa_iv, compressed_width(sa_iv), false);
  write_int_vector(dir + "/paths", paths, compressed_width(paths), false);
}

void kmer_index_load(HostIndex& ix, const std::string& dir) {
  const uint32_t k = ix.k;
  if (k < 1 || k > 14) throw std::runtime_error("kmer_size must be in [1,14]");
  const uint64_t nk = 1ull << (2 * k);
  const std::vector<uint64_t> kmers = read_int_vector(dir + "/kmers", 3, nullptr);
  const std::vector<uint64_t> stats = read_int_vector(dir + "/kmers_stats", 0, nullptr);
  const std::vector<uint64_t> sa_iv = read_int_vector(dir + "/sa_intervals", 0, nullptr);
  const std::vector<uint64_t> paths = read_int_vector(dir + "/paths", 0, nullptr);
  if (kmers.size() % k != 0) throw std::runtime_error("kmers: length is not a multiple of kmer_size");
  const uint64_t n_kmers = kmers.size() / k;
  // pass 1: the code of every k-mer of the file, its number of states and where its entries start
  struct FileKmer {
    uint64_t code, stats_at, state_at, path_at;
    uint32_t n_states;
  };
  std::vector<FileKmer> fk(n_kmers);
  uint64_t si = 0, state_i = 0, path_i = 0;
  for (uint64_t q = 0; q < n_kmers; ++q) {
    uint64_t code = 0;
    for (uint32_t j = 0; j < k; ++j) {
      const uint64_t b = kmers[q * k + j];
      if (b < 1 || b > 4) throw std::runtime_error("kmers: base code outside 1..4");
      code |= (b - 1) << (2 * j);
    }
    if (si >= stats.size()) throw std::runtime_error("kmers_stats: shorter than kmers");
    const uint64_t ns = stats[si];
    if (si + 1 + ns > stats.size()) throw std::runtime_error("kmers_stats: truncated");
    fk[q] = FileKmer{code, si, state_i, path_i, (uint32_t)ns};
    for (uint64_t t = 0; t < ns; ++t) path_i += 2 * stats[si + 1 + t];
    state_i += ns;
    si += 1 + ns;
  }
  if (2 * state_i != sa_iv.size()) throw std::runtime_error("sa_intervals: length does not match kmers_stats");
  if (path_i != paths.size()) throw std::runtime_error("paths: length does not match kmers_stats");
  if (state_i >= 0xFFFFFFFFull || path_i >= 0xFFFFFFFFull) throw std::runtime_error("k-mer index exceeds 2^32 entries");
  ix.kmer_bits.assign((((nk + 31) / 32) + 3) & ~3ull, 0);
  ix.kmer_off.assign(nk + 1, 0);
  for (const FileKmer& f : fk) {
    if (ix.kmer_off[f.code + 1] != 0 && f.n_states) throw std::runtime_error("kmers: a k-mer occurs twice");
    ix.kmer_off[f.code + 1] = f.n_states;
  }
  for (uint64_t c = 0; c < nk; ++c) {
    if (ix.kmer_off[c + 1]) ix.kmer_bits[c >> 5] |= 1u << (c & 31);
    ix.kmer_off[c + 1] += ix.kmer_off[c];
  }
  // pass 2: states in CSR order; path words re-packed as [site, allele]*nt then [site]*ng per state, k-mer by k-mer
  // in ascending code order (a file in another order gives the same index)
  std::vector<uint64_t> by_code(n_kmers);
  for (uint64_t q = 0; q < n_kmers; ++q) by_code[q] = q;
  std::sort(by_code.begin(), by_code.end(), [&](uint64_t a, uint64_t b) { return fk[a].code < fk[b].code; });
  ix.kmer_states.assign(std::max<uint64_t>(state_i, 1), KmerState{});
  ix.kmer_paths.clear();
  for (uint64_t q : by_code) {
    const FileKmer& f = fk[q];
    uint64_t pa = f.path_at;
    for (uint32_t t = 0; t < f.n_states; ++t) {
      const uint64_t plen = stats[f.stats_at + 1 + t];
      uint32_t nt = 0, ng = 0;
      const uint32_t off = (uint32_t)ix.kmer_paths.size();
      for (uint64_t e = 0; e < plen; ++e)  // traversed loci (allele known) first, as dump.cpp writes them
        if (paths[pa + 2 * e + 1] != 0) {
          if (ng) throw std::runtime_error("paths: a traversed locus follows a traversing one");
          ix.kmer_paths.push_back((uint32_t)paths[pa + 2 * e]);
          ix.kmer_paths.push_back((uint32_t)(paths[pa + 2 * e + 1] - 1));
          ++nt;
        } else
          ++ng;
      for (uint64_t e = nt; e < plen; ++e) ix.kmer_paths.push_back((uint32_t)paths[pa + 2 * e]);
      if (nt > 0xFFFFu || ng > 0xFFFFu) throw std::runtime_error("paths: path too long");
      const uint64_t lo = sa_iv[2 * (f.state_at + t)], hi = sa_iv[2 * (f.state_at + t) + 1];
      if (lo > hi || hi >= ix.n) throw std::runtime_error("sa_intervals: interval outside the suffix array");
      ix.kmer_states[ix.kmer_off[f.code] + t] = KmerState{(uint32_t)lo, (uint32_t)hi, off, nt | (ng << 16)};
      pa += 2 * plen;
    }
  }
  if (ix.kmer_paths.empty()) ix.kmer_paths.push_back(0);
  build_seed_view(ix);
}

}  // namespace gq
