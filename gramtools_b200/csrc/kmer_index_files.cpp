// kmer_index_files.cpp — the k-mer index as the reference's gram_dir files (see index_build.hpp). Host-only.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
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

// ---- the whole flat index as one file (this back-end's own format; `gram build` writes gram_dir/gq_index) ------------
// [magic "GQINDEX1"][record sizes of the structs][scalars][every array: u64 count + raw little-endian elements]
// [u64 word-wise checksum of everything before it]. Loading replaces the whole host build (suffix array, FM tables,
// graph, k-mer searches, seed view); the arrays are then uploaded exactly as after a build.
namespace {
struct IoHash {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  void mix(const void* p, size_t bytes) {
    const uint8_t* b = (const uint8_t*)p;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
      uint64_t w;
      std::memcpy(&w, b + i, 8);
      h = (h ^ w) * 0xD6E8FEB86659FD93ull;
      h ^= h >> 32;
    }
    for (; i < bytes; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
  }
};
struct Writer {
  FILE* f;
  IoHash hash;
  bool ok = true;
  void raw(const void* p, size_t bytes) {
    if (bytes && fwrite(p, 1, bytes, f) != bytes) ok = false;
    hash.mix(p, bytes);
  }
  template <class T>
  void scalar(T& x) { raw(&x, sizeof(T)); }
  template <class T>
  void vec(std::vector<T>& v) {
    uint64_t n = v.size();
    raw(&n, 8);
    raw(v.data(), n * sizeof(T));
  }
};
struct Reader {
  FILE* f;
  IoHash hash;
  uint64_t left;  // bytes of the file not yet consumed (bounds every count read from it)
  void raw(void* p, size_t bytes) {
    if (bytes > left || (bytes && fread(p, 1, bytes, f) != bytes)) throw std::runtime_error("gq_index file: truncated");
    left -= bytes;
    hash.mix(p, bytes);
  }
  template <class T>
  void scalar(T& x) { raw(&x, sizeof(T)); }
  template <class T>
  void vec(std::vector<T>& v) {
    uint64_t n = 0;
    raw(&n, 8);
    if (n > left / sizeof(T)) throw std::runtime_error("gq_index file: corrupt array length");
    v.resize(n);
    raw(v.data(), n * sizeof(T));
  }
};
// every member of HostIndex, in one place for the writer and the reader
template <class Io>
void visit_index(HostIndex& h, Io& io) {
  io.scalar(h.n); io.scalar(h.k); io.scalar(h.c_base); io.scalar(h.n_slots); io.scalar(h.n_sites);
  uint32_t nested = h.is_nested ? 1u : 0u;
  io.scalar(nested);
  h.is_nested = nested != 0;
  io.scalar(h.n_per_base);
  io.vec(h.prg); io.vec(h.sa); io.vec(h.rank_blk); io.vec(h.super_cnt); io.vec(h.mrank_blk); io.vec(h.marker_hit);
  io.vec(h.text_grp); io.vec(h.text_super); io.vec(h.tmarker_hit); io.vec(h.isa);
  io.vec(h.site_sa); io.vec(h.allele_iv); io.vec(h.par); io.vec(h.tm_odd); io.vec(h.tm_even_off); io.vec(h.tm_even);
  io.vec(h.entry_next); io.vec(h.site_snp); io.vec(h.site_start_pos); io.vec(h.n_alleles); io.vec(h.allele_off);
  io.vec(h.pos2node); io.vec(h.nodes); io.vec(h.edges); io.vec(h.site_start_node); io.vec(h.site_rec); io.vec(h.apos);
  io.vec(h.kmer_bits); io.vec(h.kmer_off); io.vec(h.kmer_paths); io.vec(h.kmer_states); io.vec(h.seed_off);
  io.vec(h.seed_state); io.vec(h.seed_ent);
}
const char kIndexMagic[8] = {'G', 'Q', 'I', 'N', 'D', 'E', 'X', '1'};
const uint32_t kRecordSizes[6] = {(uint32_t)sizeof(RankBlk), (uint32_t)sizeof(Node),     (uint32_t)sizeof(KmerState),
                                  (uint32_t)sizeof(KmerSeed), (uint32_t)sizeof(TextGrp), 0x01020304u /* byte order */};
}  // namespace

void host_index_save(const HostIndex& ix, const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  Writer w{f};
  w.raw(kIndexMagic, 8);
  w.raw(kRecordSizes, sizeof(kRecordSizes));
  visit_index(const_cast<HostIndex&>(ix), w);  // the writer only reads
  const uint64_t sum = w.hash.h;
  const bool ok = w.ok && fwrite(&sum, 8, 1, f) == 1;
  if (fclose(f) != 0 || !ok) throw std::runtime_error("write error on " + path);
}

void host_index_load(HostIndex& ix, const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot read " + path);
  try {
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (size < 8 + (long)sizeof(kRecordSizes) + 8) throw std::runtime_error("gq_index file: too short");
    Reader r{f, IoHash{}, (uint64_t)size - 8};
    char magic[8];
    uint32_t sizes[6];
    r.raw(magic, 8);
    r.raw(sizes, sizeof(sizes));
    if (std::memcmp(magic, kIndexMagic, 8) != 0) throw std::runtime_error("gq_index file: not a gq index (or another version)");
    if (std::memcmp(sizes, kRecordSizes, sizeof(sizes)) != 0) throw std::runtime_error("gq_index file: written with another record layout");
    ix = HostIndex{};
    visit_index(ix, r);
    if (r.left != 0) throw std::runtime_error("gq_index file: trailing bytes");
    uint64_t sum = 0;
    if (fread(&sum, 8, 1, f) != 1 || sum != r.hash.h) throw std::runtime_error("gq_index file: checksum mismatch");
    // the cheap consistency checks a corrupted-but-checksummed file could not fail, a foreign one could
    const uint64_t nk = 1ull << (2 * ix.k);
    if (ix.k < 1 || ix.k > 14 || ix.sa.size() != ix.n || ix.isa.size() != ix.n || ix.prg.size() + 1 != ix.n ||
        ix.kmer_off.size() != nk + 1 || ix.allele_off.size() != (size_t)ix.n_slots + 1)
      throw std::runtime_error("gq_index file: inconsistent sizes");
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
}

}  // namespace gq
