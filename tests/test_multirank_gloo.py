"""N>1 host logic on CPU: world_size-2 gloo. Each rank maps its shard of the reads (through the
test-only host emulation of the device functions), the dense counters go through ONE all-reduce(sum),
sparse groups are gathered and merged, uint16 semantics are applied once — and the result must equal
the oracle run on the whole read set."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import Emu, Oracle
from gramtools_b200 import master_seeds, synth
from gramtools_b200.distributed import finalize_counters, merge_group_records, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, prg, k, bases, offs, seeds, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = offs.size - 1
    lo, hi = shard_bounds(n, rank, world)
    e = Emu(prg, k)
    sub_offs = offs[lo:hi + 1] - offs[lo]
    e.map(bases[int(offs[lo]):int(offs[hi])], sub_offs, seeds[lo:hi])
    cnt = torch.from_numpy(e.counters_raw().astype(np.int64))
    dist.all_reduce(cnt)                       # the one data-path collective
    st = torch.from_numpy(e.stats().astype(np.int64))
    dist.all_reduce(st)
    gathered = [None] * world
    dist.all_gather_object(gathered, e.groups_raw())
    if rank == 0:
        groups = merge_group_records(gathered)
        np.savez(out_path, cnt=cnt.numpy(), st=st.numpy(), groups=groups)
    dist.destroy_process_group()


def test_two_rank_shard_and_reduce(tmp_path):
    prg = synth.make_nested_prg(3, 250, 21)
    rng = np.random.default_rng(21)
    haps = [synth.random_haplotype(prg, rng) for _ in range(4)]
    bases, offs = synth.sample_reads(haps, 601, 30, 21, frac_garbage=0.05, frac_n=0.02)
    seeds = master_seeds(42, offs.size - 1)
    k = 4
    out = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(2, _free_port(), prg, k, bases, offs, seeds, out), nprocs=2, join=True)
    z = np.load(out)
    o = Oracle(prg, k)
    o.map(bases, offs, seeds)
    ref = o.result()
    na, npb = o.n_alleles, o.n_per_base
    allele_off = _allele_offsets(prg)
    a, p, g = finalize_counters(z["cnt"], na, npb, z["groups"], allele_off)
    assert np.array_equal(a, ref.allele_sum)
    assert np.array_equal(p, ref.per_base)
    assert np.array_equal(g, ref.grouped)
    assert [int(x) for x in z["st"]] == ref.stats


def _allele_offsets(prg):
    """alleles per site from the PRG itself: separators + 1, sites in id order."""
    counts = {}
    for m in [int(x) for x in prg]:
        if m > 4 and m % 2 == 0:
            counts[m - 1] = counts.get(m - 1, 0) + 1
    off = [0]
    for s in sorted(counts):
        off.append(off[-1] + counts[s])
    return off


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
