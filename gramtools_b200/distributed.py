"""Host-side logic of the multi-GPU path: shard reads, reduce coverage.

One process per GPU, index replicated, reads sharded by contiguous ranges. Read j of the whole job
keeps the j-th draw of the master mt19937 (quasimap.cpp:136-137), so results do not depend on the
sharding. The only exchange: one all-reduce(sum) of the dense uint32 accumulators plus a gather/merge
of the sparse multi-allele groups; uint16 semantics are applied once on the reduced totals.
"""
import os

import numpy as np


def bind_to_device_numa(device):
    """Pin this process (its threads and, by first touch, the pinned buffers it allocates afterwards) to the CPUs of
    the NUMA node GPU `device` hangs off. With one process per GPU all pulling reads from host memory at PCIe rate,
    buffers on the far socket cross the inter-socket link and the ranks' copies slow each other down. Returns the
    previous affinity (restore it with os.sched_setaffinity before CPU-bound legs) or None when nothing was done:
    the topology is read from NVML + sysfs and any missing piece leaves the process as it was."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device
        if vis:  # CUDA ordinal -> NVML index when the visible set is a list of ordinals
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if device < len(ids) and ids[device].isdigit():
                index = int(ids[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


def shard_bounds(n_reads, rank, world):
    """Contiguous, balanced read range [lo, hi) of `rank`."""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_group_records(record_arrays):
    """Merge [slot, count, n, alleles...] record streams by key, summing counts (raw, no wrap)."""
    acc = {}
    for words in record_arrays:
        w = [int(x) for x in np.asarray(words).ravel()]
        i = 0
        while i < len(w):
            n = w[i + 2]
            key = (w[i],) + tuple(w[i + 3:i + 3 + n])
            acc[key] = acc.get(key, 0) + w[i + 1]
            i += 3 + n
    out = []
    for key in sorted(acc):
        out += [key[0], acc[key], len(key) - 1, *key[1:]]
    return np.asarray(out, dtype=np.uint32)


def finalize_counters(counters_u32, n_alleles, n_per_base, group_records, allele_off):
    """Reduced uint32 totals -> the reference's uint16 views.

    allele_sum / grouped counts wrap mod 65536 (allele_sum.cpp:41, grouped_allele_counts.cpp:47),
    per-base counts saturate at 65535 (allele_base.cpp:239). Returns (allele_sum, per_base, grouped
    records sorted by (slot, alleles)) in the formats of gq_coverage_fetch / gq_coverage_grouped."""
    c = np.asarray(counters_u32, dtype=np.uint64)
    allele_sum = (c[:n_alleles] & 0xFFFF).astype(np.uint16)
    single = c[n_alleles:2 * n_alleles]
    per_base = np.minimum(c[2 * n_alleles:2 * n_alleles + n_per_base], 65535).astype(np.uint16)
    acc = {}
    allele_off = [int(x) for x in allele_off]
    for s in range(len(allele_off) - 1):
        for a in range(allele_off[s + 1] - allele_off[s]):
            v = int(single[allele_off[s] + a])
            if v:
                acc[(s, a)] = v
    w = [int(x) for x in np.asarray(group_records).ravel()]
    i = 0
    while i < len(w):
        n = w[i + 2]
        key = (w[i],) + tuple(w[i + 3:i + 3 + n])
        acc[key] = acc.get(key, 0) + w[i + 1]
        i += 3 + n
    out = []
    for key in sorted(acc):
        out += [key[0], acc[key] & 0xFFFF, len(key) - 1, *key[1:]]
    return allele_sum, per_base, np.asarray(out, dtype=np.uint32)
