"""Byte model of the implemented algorithm (profiles/r02_byte_model.json), per BASELINE config.

The TEST-ONLY host emulation runs the very device functions of the kernels (gq_device.cuh) one strand at a time;
built with -DGQ_EMU_COUNTERS every global-memory access of those functions goes through GQ_LDG / GQ_AT / GQ_TOUCH,
and the emulation records, per thread-sized unit of work (one strand in seed / general / classify / coverage, one
candidate in verify / text) the DISTINCT 32-byte sectors it touches in every index structure, and the bytes it
touches in the per-batch arrays that consecutive threads access side by side (packed reads, per-strand words,
candidate / final-state records: coalesced, so a touch costs its bytes). That is the traffic the algorithm needs
when nothing is shared between threads through a cache = its algorithmic bytes. bench.py divides them by the
kernels' CUDA-event durations. Not counted: per-lane scratch arenas of the general kernel and of the general coverage
path (L1/L2-resident working memory), the 4^k-bit presence set of the classify kernel (one copy per CTA in shared
memory: 128 KB x 148 per launch at k = 10, reported separately), the coverage kernel's own work-list reads.

  python tools/byte_model.py [config ...]      (CPU only; a few minutes for config 2)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from common import Emu  # noqa: E402

SAMPLE = {1: 10_000, 2: 20_000, 3: 5_000, 4: 5_000}
out_path = os.path.join(ROOT, "profiles", "r02_byte_model.json")
model = json.load(open(out_path)) if os.path.exists(out_path) else {}
for config in [int(x) for x in sys.argv[1:]] or [1, 2, 3]:
    c = bench.CONFIGS[config]
    prg, haps = bench.make_prg(config)
    n = min(SAMPLE[config], c["n_reads"])
    bases, offs, seeds = bench.make_reads(config, haps, 0, n, first_read=0)
    e = Emu(prg, c["k"])
    e.track(True)
    e.map(bases, offs, seeds, arena_words=512)
    by_kernel, units = e.tracked()
    e.track(False)
    # revcomp_kernel (not a per-strand device function of the emulation): streams the slice's packed words in and
    # their reverse complements out, plus len and word_off
    import numpy as np  # noqa: E402
    words = int(((np.diff(offs.astype(np.int64)) + 15) // 16).sum())
    by_kernel["revcomp_kernel"] = {"packed reads": 4.0 * words, "packed reverse strands": 4.0 * words, "read len": 4.0 * n,
                                   "read word_off": 4.0 * n}
    units["revcomp_kernel"] = n
    smem = {}
    for kn, d in by_kernel.items():  # the presence set lives in shared memory on the device
        smem[kn] = sum(v for s, v in d.items() if "smem copy" in s)
        by_kernel[kn] = {s: v for s, v in d.items() if "smem copy" not in s}
    model[f"config{config}"] = {
        "workload": c["text"], "sample_reads": n,
        "bytes_per_read": {kn: sum(d.values()) / n for kn, d in by_kernel.items()},
        "bytes_per_read_by_structure": {kn: {s: v / n for s, v in sorted(d.items(), key=lambda kv: -kv[1])}
                                        for kn, d in by_kernel.items()},
        "units_per_read": {kn: u / n for kn, u in units.items()},
        "presence_set_probe_bytes_per_read_if_it_were_in_global_memory": {kn: v / n for kn, v in smem.items() if v},
        "stats": [int(x) for x in e.stats()],
    }
    tot = sum(model[f"config{config}"]["bytes_per_read"].values())
    print(f"config {config}: {tot:.0f} B/read", {k: round(v) for k, v in model[f'config{config}']['bytes_per_read'].items()})
with open(out_path, "w") as f:
    json.dump(model, f, indent=1)
