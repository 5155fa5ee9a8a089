// old_read_file.hpp — TEST-ONLY: the line-by-line reader `gram genotype` used before read_file.hpp (gzgets + one
// std::string per line), kept verbatim as the behavioural reference of tests/test_read_file.py.
#pragma once
#include <zlib.h>

#include <cstring>
#include <stdexcept>
#include <string>

// Minimal sequence reader with the behaviour of SeqRead / seq_file.h that matters here
// (include/sequence_read/seqread.hpp:94-180, seq_file.h:247-335): format sniffed from the first byte
// ('@' FASTQ, '>' FASTA, else one read per line), multi-line records joined, gz transparently
// inflated, a malformed record ends the file.
class OldReadFile {
 public:
  explicit OldReadFile(const std::string& path) {
    gz_ = gzopen(path.c_str(), "rb");
    if (!gz_) throw std::runtime_error("Cannot open reads file " + path);
    gzbuffer(gz_, 1 << 20);
    have_line_ = next_line(line_);
    if (have_line_) fmt_ = line_.empty() ? 'p' : (line_[0] == '@' ? 'q' : (line_[0] == '>' ? 'a' : 'p'));
  }
  ~OldReadFile() {
    if (gz_) gzclose(gz_);
  }
  bool next(std::string& seq, std::string& qual) {
    seq.clear();
    qual.clear();
    if (!have_line_) return false;
    if (fmt_ == 'p') {
      seq = line_;
      have_line_ = next_line(line_);
      return true;
    }
    if (fmt_ == 'a') {
      if (line_.empty() || line_[0] != '>') return false;
      while ((have_line_ = next_line(line_)) && (line_.empty() || line_[0] != '>')) seq += line_;
      return true;
    }
    if (line_.empty() || line_[0] != '@') return false;
    while ((have_line_ = next_line(line_)) && (line_.empty() || line_[0] != '+')) seq += line_;
    if (!have_line_) return false;  // no '+' line: malformed
    while (qual.size() < seq.size() && (have_line_ = next_line(line_))) qual += line_;
    if (qual.size() != seq.size()) return false;
    have_line_ = next_line(line_);
    return true;
  }

 private:
  bool next_line(std::string& out) {
    out.clear();
    char buf[1 << 16];
    while (gzgets(gz_, buf, sizeof buf)) {
      size_t n = std::strlen(buf);
      bool eol = n && buf[n - 1] == '\n';
      if (eol) --n;
      if (n && buf[n - 1] == '\r') --n;
      out.append(buf, n);
      if (eol) return true;
    }
    return !out.empty();
  }
  gzFile gz_ = nullptr;
  std::string line_;
  bool have_line_ = false;
  char fmt_ = 'p';
};

