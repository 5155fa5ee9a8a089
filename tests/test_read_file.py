"""`gram genotype`'s sequence-file reader (gramtools_b200/csrc/read_file.hpp: lines as views into one inflate buffer)
against the line-by-line reader it replaced (tests/emu/old_read_file.hpp, kept verbatim): same records, bases and
qualities on FASTQ / FASTA / one-read-per-line files, multi-line records, CRLF, gz, missing final newline, empty
lines, lines longer than the buffer, the reference's integration reads — and a malformed record ends the file."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from common import ROOT, emu_lib


def _read(path, mode, buffer_bytes=0):
    lib = emu_lib()
    out = np.zeros(4, dtype=np.uint64)
    rc = lib.emu_read_file(os.fsencode(str(path)), mode, buffer_bytes, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0, (rc, lib.emu_last_error().decode())
    return [int(x) for x in out]


def _same_everywhere(path):
    old = _read(path, 0)
    for buf in (0, 64, 100, 4096):  # tiny buffers: every line straddles a refill
        new = _read(path, 1, buf)
        assert new == old, (path, buf, new, old)
        seq_only = _read(path, 2, buf)
        assert seq_only[:3] == old[:3], (path, buf, seq_only, old)
    return old


def _fastq(rng, n, multi=False, crlf=False, lower=False, final_newline=True):
    nl = "\r\n" if crlf else "\n"
    recs = []
    for i in range(n):
        L = int(rng.integers(0 if i % 17 == 5 else 1, 400))
        s = "".join(rng.choice(list("acgtn" if lower else "ACGTN"), L, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        q = "".join(chr(int(x)) for x in rng.integers(33, 74, L))  # may start with '@' or '+'
        if multi and L > 10:
            w = int(rng.integers(5, 80))
            s = nl.join(s[j:j + w] for j in range(0, L, w))
            q = nl.join(q[j:j + w] for j in range(0, L, w))
        recs.append(f"@read{i} some description{nl}{s}{nl}+{'read%d' % i if i % 3 == 0 else ''}{nl}{q}")
    return nl.join(recs) + (nl if final_newline else "")


def test_equivalence_on_generated_files(tmp_path):
    rng = np.random.default_rng(11)
    files = {}
    files["plain.fastq"] = _fastq(rng, 300)
    files["multi.fastq"] = _fastq(rng, 200, multi=True)
    files["crlf.fastq"] = _fastq(rng, 100, crlf=True)
    files["crlf_multi.fastq"] = _fastq(rng, 100, multi=True, crlf=True)
    files["lower_nonewline.fastq"] = _fastq(rng, 50, lower=True, final_newline=False)
    files["one.fastq"] = "@r\nACGT\n+\nIIII\n"
    files["fasta.fa"] = "".join(f">s{i} d\n" + "\n".join("".join(rng.choice(list("ACGT"), int(rng.integers(0, 70))))
                                                          for _ in range(int(rng.integers(0, 6)))) + "\n" for i in range(120))
    files["fasta_nonl.fa"] = ">a\nACGT\nAC\n>b\nGG"
    files["fasta_empty_lines.fa"] = ">a\n\nACGT\n\n\n>b\n\n>c\nTT\n"
    files["lines.txt"] = "".join("".join(rng.choice(list("ACGT"), int(rng.integers(0, 90)))) + "\n" for _ in range(200))
    files["lines_first_empty.txt"] = "\nACGT\n\nGG\n"
    files["long_lines.fastq"] = "@r1\n" + "ACGT" * 50000 + "\n+\n" + "I" * 200000 + "\n@r2\nAC\n+\nII\n"
    files["empty.fastq"] = ""
    files["only_newlines.txt"] = "\n\n\n"
    for name, text in files.items():
        p = tmp_path / name
        p.write_bytes(text.encode())
        got = _same_everywhere(p)
        gz = tmp_path / (name + ".gz")
        with gzip.open(gz, "wb") as f:
            f.write(text.encode())
        assert _same_everywhere(gz) == got, name
    assert _read(tmp_path / "plain.fastq", 0)[0] == 300 and _read(tmp_path / "fasta.fa", 0)[0] == 120
    assert _read(tmp_path / "long_lines.fastq", 2, 64)[:2] == [2, 200002]


def test_reference_integration_reads():
    """The reads of the reference's integration tests (when the reference tree is present: development container)."""
    base = "/root/reference/gramtools/tests/integration_test_data"
    if not os.path.isdir(base):
        pytest.skip("reference tree not present")
    n = 0
    for it in sorted(os.listdir(base)):
        p = os.path.join(base, it, "reads.fastq")
        if os.path.exists(p):
            assert _same_everywhere(p)[0] > 0
            n += 1
    assert n >= 3


def test_malformed_record_ends_the_file(tmp_path):
    good = "@a\nACGT\n+\nIIII\n@b\nGGCC\n+\nIIII\n"
    cases = {
        "short_quality": good + "@c\nACGTACGT\n+\nIII\n@d\nAC\n+\nII\n",   # qualities run into the next record
        "no_plus": good + "@c\nACGT\n",
        "long_quality": good + "@c\nACGT\n+\nIIIIII\n@d\nAC\n+\nII\n",
        "not_a_header": good + "ACGT\n+\nIIII\n",
        "fasta_garbage_first": "ACGT\n>a\nAC\n",                          # sniffed as one read per line: 3 'reads'
    }
    expect = {"short_quality": 3,   # length-based quality reading swallows "@d", "AC", "+": a (wrong) third record, then the end
              "no_plus": 2, "long_quality": 2, "not_a_header": 2, "fasta_garbage_first": 3}
    for name, text in cases.items():
        p = tmp_path / name
        p.write_bytes(text.encode())
        old, new, seq_only = _read(p, 0), _read(p, 1), _read(p, 2)
        assert new == old and seq_only[:3] == old[:3], name   # both stop at the same record, with the same reads
        assert new[0] == expect[name], (name, new)
    with pytest.raises(AssertionError):
        _read(tmp_path / "missing.fastq", 1)


def test_random_bytes_fuzz(tmp_path):
    """Arbitrary line soup over an alphabet rich in '@', '+', '>' and '\\r': whatever the old reader makes of it (records
    until the first malformed one), the new reader makes the same of it, at every buffer size."""
    rng = np.random.default_rng(21)
    alphabet = np.frombuffer(b"ACGTNacgt@+>I#\r ", dtype=np.uint8)
    for case in range(150):
        first = [b"@", b">", b"A", b""][case % 4]
        lines = [first + bytes(rng.choice(alphabet, int(rng.integers(0, 30)))) for _ in range(int(rng.integers(0, 40)))]
        if case % 3 == 0:  # mostly well-formed FASTQ with a few random lines thrown in
            lines = []
            for i in range(int(rng.integers(1, 12))):
                L = int(rng.integers(0, 25))
                lines += [b"@r%d" % i, bytes(rng.choice(alphabet[:9], L)), b"+", bytes(rng.choice(alphabet, L)).replace(b"\r", b"I")]
                if rng.random() < 0.15:
                    lines.insert(int(rng.integers(0, len(lines))), bytes(rng.choice(alphabet, int(rng.integers(0, 10)))))
        data = b"\n".join(lines) + (b"\n" if case % 5 else b"")
        p = tmp_path / f"fuzz{case}"
        p.write_bytes(data)
        _same_everywhere(p)
