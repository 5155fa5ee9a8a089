"""Generates tests/golden/it_fixtures.json from the reference's integration fixtures
(gramtools/tests/integration_test_data/IT{1,2,3}/{prg.bin,reads.fastq}) — run in the dev container,
where /root/reference is mounted. Expected values are the ones asserted by
gramtools/tests/genotype/test_genotype_integration_tests.py:68-157 (copied here as data)."""
import json
import os
import struct

REF = "/root/reference/gramtools/tests/integration_test_data"
EXPECTED = {
    "IT1": {  # :68-101  PRG AAA[CC,TA]AC[TTTT,GGG]
        "allele_base_counts": [[[0, 1], [1, 1]], [[1, 1, 1, 1], [1, 1, 0]]],
        "grouped": {"0": {"0": 1, "1": 1}, "1": {"0": 1, "1": 1}},
    },
    "IT2": {  # :104-130  PRG TT[AAAc,AAAg]gg[cAA,gAA]TTCAA
        "allele_base_counts": [[[1, 1, 1, 0], [1, 1, 1, 0]], [[0, 1, 1], [0, 1, 1]]],
        "grouped": {"0": {"0,1": 1}, "1": {"0,1": 1}},
    },
    "IT3": {  # :133-157  PRG T[cCCC[A,g]CT,]ATTTTt (nested: no per-base dump)
        "grouped": {"0": {"0,1": 1, "0": 1}, "1": {"0": 1}},
    },
}


def read_prg(path):
    b = open(path, "rb").read()
    return list(struct.unpack("<%dI" % (len(b) // 4), b))


def read_fastq(path):
    lines = open(path).read().split("\n")
    return [lines[i + 1].strip() for i in range(0, len(lines) - 1, 4) if lines[i].startswith("@")]


out = {}
for name, exp in EXPECTED.items():
    d = os.path.join(REF, name)
    case = {"prg": read_prg(os.path.join(d, "prg.bin")), "reads": read_fastq(os.path.join(d, "reads.fastq")),
            "kmer_size": 5, "seed": 42}
    case.update(exp)
    out[name] = case
json.dump(out, open(os.path.join(os.path.dirname(__file__), "it_fixtures.json"), "w"), indent=1)
print({k: (v["prg"], v["reads"]) for k, v in out.items()})
