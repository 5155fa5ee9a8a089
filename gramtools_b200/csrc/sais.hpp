// sais.hpp — suffix array by induced sorting (SA-IS) over an integer alphabet.
//
// Product-side replacement for `sdsl::construct(fm_index, prg, cfg, 4)`
// (reference: libgramtools/src/prg/make_data_structures.cpp:9-33). Input text must end in a
// unique smallest symbol (the sentinel 0 that sdsl appends), symbols in [0, sigma).
// Linear time, 32-bit indices (text length < 2^31).
#pragma once
#include <cstdint>
#include <vector>

namespace gq {
namespace sais_detail {

template <class T>
static void bucket_bounds(const T* s, int32_t n, int32_t sigma, std::vector<int32_t>& bkt, bool end) {
  std::fill(bkt.begin(), bkt.end(), 0);
  for (int32_t i = 0; i < n; ++i) bkt[s[i]]++;
  int32_t sum = 0;
  for (int32_t c = 0; c < sigma; ++c) {
    sum += bkt[c];
    bkt[c] = end ? sum : sum - bkt[c];
  }
}

template <class T>
static void induce(const T* s, int32_t* sa, int32_t n, int32_t sigma, const std::vector<bool>& is_s,
                   std::vector<int32_t>& bkt) {
  bucket_bounds(s, n, sigma, bkt, false);  // L-type: left to right, bucket heads
  for (int32_t i = 0; i < n; ++i) {
    int32_t j = sa[i] - 1;
    if (sa[i] > 0 && !is_s[j]) sa[bkt[s[j]]++] = j;
  }
  bucket_bounds(s, n, sigma, bkt, true);  // S-type: right to left, bucket tails
  for (int32_t i = n - 1; i >= 0; --i) {
    int32_t j = sa[i] - 1;
    if (sa[i] > 0 && is_s[j]) sa[--bkt[s[j]]] = j;
  }
}

template <class T>
static void sais(const T* s, int32_t* sa, int32_t n, int32_t sigma) {
  if (n == 1) {
    sa[0] = 0;
    return;
  }
  std::vector<bool> is_s(n);
  is_s[n - 1] = true;
  for (int32_t i = n - 2; i >= 0; --i) is_s[i] = s[i] < s[i + 1] || (s[i] == s[i + 1] && is_s[i + 1]);
  auto is_lms = [&](int32_t i) { return i > 0 && is_s[i] && !is_s[i - 1]; };

  std::vector<int32_t> bkt(sigma);
  // 1. place LMS suffixes at bucket tails, induce
  bucket_bounds(s, n, sigma, bkt, true);
  for (int32_t i = 0; i < n; ++i) sa[i] = -1;
  for (int32_t i = 1; i < n; ++i)
    if (is_lms(i)) sa[--bkt[s[i]]] = i;
  induce(s, sa, n, sigma, is_s, bkt);

  // 2. compact sorted LMS substrings, name them
  int32_t n1 = 0;
  for (int32_t i = 0; i < n; ++i)
    if (is_lms(sa[i])) sa[n1++] = sa[i];
  for (int32_t i = n1; i < n; ++i) sa[i] = -1;
  int32_t name = 0, prev = -1;
  for (int32_t i = 0; i < n1; ++i) {
    int32_t pos = sa[i];
    bool diff = false;
    if (prev == -1) diff = true;
    else {
      for (int32_t d = 0;; ++d) {
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
  for (int32_t i = n - 1, j = n - 1; i >= n1; --i)
    if (sa[i] >= 0) sa[j--] = sa[i];

  // 3. recurse if names are not unique
  int32_t* sa1 = sa;
  int32_t* s1 = sa + n - n1;
  if (name < n1) sais<int32_t>(s1, sa1, n1, name);
  else
    for (int32_t i = 0; i < n1; ++i) sa1[s1[i]] = i;

  // 4. map back and induce the final order
  bucket_bounds(s, n, sigma, bkt, true);
  for (int32_t i = 1, j = 0; i < n; ++i)
    if (is_lms(i)) s1[j++] = i;
  for (int32_t i = 0; i < n1; ++i) sa1[i] = s1[sa1[i]];
  for (int32_t i = n1; i < n; ++i) sa[i] = -1;
  for (int32_t i = n1 - 1; i >= 0; --i) {
    int32_t j = sa[i];
    sa[i] = -1;
    sa[--bkt[s[j]]] = j;
  }
  induce(s, sa, n, sigma, is_s, bkt);
}
}  // namespace sais_detail

// text: symbols in [0,sigma), text.back() must be the unique minimum. Returns SA (size n).
inline std::vector<int32_t> suffix_array(const std::vector<int32_t>& text, int32_t sigma) {
  std::vector<int32_t> sa(text.size());
  sais_detail::sais<int32_t>(text.data(), sa.data(), (int32_t)text.size(), sigma);
  return sa;
}
}  // namespace gq
