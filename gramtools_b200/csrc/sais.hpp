// sais.hpp — suffix array by induced sorting (SA-IS) over an integer alphabet.
//
// Product-side replacement for `sdsl::construct(fm_index, prg, cfg, 4)`
// (reference: libgramtools/src/prg/make_data_structures.cpp:9-33). Input text must end in a
// unique smallest symbol (the sentinel 0 that sdsl appends), symbols in [0, sigma).
// Linear time. The index type I is a template parameter: int32_t for texts shorter than 2^31 symbols, int64_t
// beyond (whole-genome PRGs, BASELINE config 5: 3.3e9 symbols) — suffix_array_u32 picks it and always returns
// 32-bit unsigned positions, which is what the device index stores (text length < 2^32 - 1).
#pragma once
#include <cstdint>
#include <vector>

namespace gq {
namespace sais_detail {

template <class T, class I>
static void bucket_bounds(const T* s, I n, I sigma, std::vector<I>& bkt, bool end) {
  std::fill(bkt.begin(), bkt.end(), 0);
  for (I i = 0; i < n; ++i) bkt[s[i]]++;
  I sum = 0;
  for (I c = 0; c < sigma; ++c) {
    sum += bkt[c];
    bkt[c] = end ? sum : sum - bkt[c];
  }
}

template <class T, class I>
static void induce(const T* s, I* sa, I n, I sigma, const std::vector<bool>& is_s, std::vector<I>& bkt) {
  bucket_bounds(s, n, sigma, bkt, false);  // L-type: left to right, bucket heads
  for (I i = 0; i < n; ++i) {
    I j = sa[i] - 1;
    if (sa[i] > 0 && !is_s[j]) sa[bkt[s[j]]++] = j;
  }
  bucket_bounds(s, n, sigma, bkt, true);  // S-type: right to left, bucket tails
  for (I i = n - 1; i >= 0; --i) {
    I j = sa[i] - 1;
    if (sa[i] > 0 && is_s[j]) sa[--bkt[s[j]]] = j;
  }
}

template <class T, class I>
static void sais(const T* s, I* sa, I n, I sigma) {
  if (n == 1) {
    sa[0] = 0;
    return;
  }
  std::vector<bool> is_s(n);
  is_s[n - 1] = true;
  for (I i = n - 2; i >= 0; --i) is_s[i] = s[i] < s[i + 1] || (s[i] == s[i + 1] && is_s[i + 1]);
  auto is_lms = [&](I i) { return i > 0 && is_s[i] && !is_s[i - 1]; };

  std::vector<I> bkt(sigma);
  // 1. place LMS suffixes at bucket tails, induce
  bucket_bounds(s, n, sigma, bkt, true);
  for (I i = 0; i < n; ++i) sa[i] = -1;
  for (I i = 1; i < n; ++i)
    if (is_lms(i)) sa[--bkt[s[i]]] = i;
  induce(s, sa, n, sigma, is_s, bkt);

  // 2. compact sorted LMS substrings, name them
  I n1 = 0;
  for (I i = 0; i < n; ++i)
    if (is_lms(sa[i])) sa[n1++] = sa[i];
  for (I i = n1; i < n; ++i) sa[i] = -1;
  I name = 0, prev = -1;
  for (I i = 0; i < n1; ++i) {
    I pos = sa[i];
    bool diff = false;
    if (prev == -1) diff = true;
    else {
      for (I d = 0;; ++d) {
        if (s[pos + d] != s[prev + d] || is_s[pos + d] != is_s[prev + d]) {
          diff = true;
          break;
        }
        if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
      }
    }
    if (diff) {
      ++name;
      prev = pos;
    }
    sa[n1 + (pos >> 1)] = name - 1;
  }
  for (I i = n - 1, j = n - 1; i >= n1; --i)
    if (sa[i] >= 0) sa[j--] = sa[i];

  // 3. recurse if names are not unique
  I* sa1 = sa;
  I* s1 = sa + n - n1;
  if (name < n1) sais<I, I>(s1, sa1, n1, name);
  else
    for (I i = 0; i < n1; ++i) sa1[s1[i]] = i;

  // 4. map back and induce the final order
  bucket_bounds(s, n, sigma, bkt, true);
  for (I i = 1, j = 0; i < n; ++i)
    if (is_lms(i)) s1[j++] = i;
  for (I i = 0; i < n1; ++i) sa1[i] = s1[sa1[i]];
  for (I i = n1; i < n; ++i) sa[i] = -1;
  for (I i = n1 - 1; i >= 0; --i) {
    I j = sa[i];
    sa[i] = -1;
    sa[--bkt[s[j]]] = j;
  }
  induce(s, sa, n, sigma, is_s, bkt);
}
}  // namespace sais_detail

// text: symbols in [0,sigma), text.back() must be the unique minimum. Returns SA (size n).
inline std::vector<int32_t> suffix_array(const std::vector<int32_t>& text, int32_t sigma) {
  std::vector<int32_t> sa(text.size());
  sais_detail::sais<int32_t, int32_t>(text.data(), sa.data(), (int32_t)text.size(), sigma);
  return sa;
}
// the same with 64-bit indices inside (any length; 8 bytes per suffix while it runs)
inline std::vector<int64_t> suffix_array64(const std::vector<int32_t>& text, int32_t sigma) {
  std::vector<int64_t> sa(text.size());
  sais_detail::sais<int32_t, int64_t>(text.data(), sa.data(), (int64_t)text.size(), (int64_t)sigma);
  return sa;
}
// SA as unsigned 32-bit positions for texts of up to 2^32 - 2 symbols: 32-bit SA-IS below 2^31, 64-bit above
inline std::vector<uint32_t> suffix_array_u32(const std::vector<int32_t>& text, int32_t sigma, bool force64 = false) {
  std::vector<uint32_t> out(text.size());
  if (!force64 && text.size() < (1ull << 31)) {
    std::vector<int32_t> sa = suffix_array(text, sigma);
    for (size_t i = 0; i < sa.size(); ++i) out[i] = (uint32_t)sa[i];
  } else {
    std::vector<int64_t> sa = suffix_array64(text, sigma);
    for (size_t i = 0; i < sa.size(); ++i) out[i] = (uint32_t)sa[i];
  }
  return out;
}
}  // namespace gq
