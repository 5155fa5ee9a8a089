// read_file.hpp — sequence-file reader of `gram genotype` (host side of the ingestion pipeline).
//
// Behaviour of SeqRead / seq_file.h that matters here (include/sequence_read/seqread.hpp:94-180,
// seq_file.h:247-335): format sniffed from the first byte ('@' FASTQ, '>' FASTA, else one read per line),
// multi-line records joined, gz transparently inflated, a malformed record ends the file.
//
// Lines are handed out as views into one large inflate buffer (memchr for the line ends, no per-line string): the
// sequence of a record is appended straight to the caller's batch text and its quality characters are only counted
// (next_seq), which is what the mapping loop needs; next() still returns sequence and qualities as strings for the
// read-statistics pass. Header-only so that the CPU test-suite can compare it with the line-by-line reader it
// replaced (tests/emu, tests/test_read_file.py).
#pragma once
#include <zlib.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace gq {

class ReadFile {
 public:
  explicit ReadFile(const std::string& path, size_t buffer_bytes = size_t(4) << 20) : buf_(buffer_bytes < 64 ? 64 : buffer_bytes) {
    gz_ = gzopen(path.c_str(), "rb");
    if (!gz_) throw std::runtime_error("Cannot open reads file " + path);
    gzbuffer(gz_, 1 << 20);
    have_line_ = take_line();
    if (have_line_) fmt_ = line_.empty() ? 'p' : (line_[0] == '@' ? 'q' : (line_[0] == '>' ? 'a' : 'p'));
  }
  ~ReadFile() {
    if (gz_) gzclose(gz_);
  }
  ReadFile(const ReadFile&) = delete;
  ReadFile& operator=(const ReadFile&) = delete;

  // next record: sequence and (FASTQ) qualities; false at the end of the file or at a malformed record
  bool next(std::string& seq, std::string& qual) {
    seq.clear();
    qual.clear();
    size_t n_qual = 0;
    return record(seq, &qual, n_qual);
  }
  // next record: its sequence is APPENDED to `text` (rolled back when the record is malformed), its length returned
  bool next_seq(std::string& text, size_t& seq_len) {
    const size_t before = text.size();
    size_t n_qual = 0;
    if (!record(text, nullptr, n_qual)) {
      text.resize(before);
      return false;
    }
    seq_len = text.size() - before;
    return true;
  }

 private:
  // one record in the control flow of the reader this replaces: `line_` always holds the next unconsumed line
  bool record(std::string& seq, std::string* qual, size_t& n_qual) {
    if (dead_) return false;
    if (parse(seq, qual, n_qual)) return true;
    dead_ = true;  // the end of the file, or a malformed record — which ends the file
    return false;
  }
  bool parse(std::string& seq, std::string* qual, size_t& n_qual) {
    const size_t seq0 = seq.size();
    if (!have_line_) return false;
    if (fmt_ == 'p') {
      seq.append(line_);
      have_line_ = take_line();
      return true;
    }
    const char* p;
    size_t n;
    if (fmt_ == 'a') {
      if (line_.empty() || line_[0] != '>') return false;
      // sequence lines up to the next header: appended from the buffer, only the header is kept as `line_`
      while (true) {
        if (!view_line(p, n)) {
          have_line_ = false;
          return true;
        }
        if (n && p[0] == '>') {
          line_.assign(p, n);
          have_line_ = true;
          return true;
        }
        seq.append(p, n);
      }
    }
    if (line_.empty() || line_[0] != '@') return false;
    while (true) {  // sequence lines up to the '+' line
      if (!view_line(p, n)) {
        have_line_ = false;
        return false;  // no '+' line: malformed
      }
      if (n && p[0] == '+') break;
      seq.append(p, n);
    }
    const size_t seq_len = seq.size() - seq0;
    n_qual = 0;
    bool more = true;
    while (n_qual < seq_len && (more = view_line(p, n))) {  // quality lines until as many characters as bases
      n_qual += n;
      if (qual) qual->append(p, n);
    }
    if (!more) have_line_ = false;
    if (n_qual != seq_len) return false;
    have_line_ = take_line();
    return true;
  }
  bool take_line() {  // the next line, copied into line_ (headers and the lookahead line: short)
    const char* p;
    size_t n;
    if (!view_line(p, n)) {
      line_.clear();
      return false;
    }
    line_.assign(p, n);
    return true;
  }
  // the next line as a view into the buffer (valid until the next call), without its "\n" / "\r\n"; false at the
  // end of the file (a last line without a newline counts if it is not empty)
  bool view_line(const char*& p, size_t& n) {
    while (true) {
      const char* nl = pos_ < end_ ? (const char*)std::memchr(buf_.data() + pos_, '\n', end_ - pos_) : nullptr;
      if (nl) {
        p = buf_.data() + pos_;
        n = (size_t)(nl - p);
        pos_ += n + 1;
        if (n && p[n - 1] == '\r') --n;
        return true;
      }
      if (eof_) {
        if (pos_ >= end_) return false;
        p = buf_.data() + pos_;
        n = end_ - pos_;
        pos_ = end_;
        if (n && p[n - 1] == '\r') --n;
        return n != 0;
      }
      // no complete line in the buffer: keep the partial one at the front and read on (a line longer than the
      // buffer grows it)
      if (pos_ > 0) {
        std::memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
        end_ -= pos_;
        pos_ = 0;
      }
      if (end_ == buf_.size()) buf_.resize(2 * buf_.size());
      const size_t want = buf_.size() - end_;
      const int got = gzread(gz_, buf_.data() + end_, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
      if (got <= 0) eof_ = true;
      else end_ += (size_t)got;
    }
  }

  gzFile gz_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
  std::string line_;
  bool have_line_ = false, dead_ = false;
  char fmt_ = 'p';
};

}  // namespace gq
