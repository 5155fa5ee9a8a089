#!/usr/bin/env python
"""bench.py — quasimap reads/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (k-mer filter + seeding + vBWT backward search + coverage
recording, both strands) over one batch of synthetic reads. Workload at N=1 = BASELINE config 2:
4.4 Mb PRG, 100k biallelic SNPs, 1M x 150 bp reads, kmer_size 10. For N>1 every rank maps its own
1M-read shard against a replicated index (weak scaling) and each step ends with one NCCL
all-reduce(sum) of the coverage counters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF_LEN, N_SITES, N_READS, READ_LEN, KMER = 4_400_000, 100_000, 1_000_000, 150, 10
GEN_SEED = 0x6772616D + 2
MAP_SEED = 42
# ncu --set full, config 2, 1M reads, dram__bytes_read.sum + dram__bytes_write.sum summed over seed_kernel,
# verify_kernel, text_kernel and search_kernel (profiles/r01_v13_kernels_summary.txt)
SEARCH_PHASE_DRAM_BYTES = 187977728 + 30433280 + 90977280 + 6275584 + 335387136 + 97702912 + 66304


def env_int(name, default):
    return int(os.environ.get(name, default))


def make_workload(rank, n_reads):
    from gramtools_b200 import synth
    prg, ref, pos, alt = synth.make_snp_prg(REF_LEN, N_SITES, GEN_SEED)
    haps = synth.snp_haplotypes(ref, pos, alt, 8, GEN_SEED + 1)
    bases, offs = synth.sample_reads(haps, n_reads, READ_LEN, GEN_SEED + 100 + rank)
    # seeds: read j of the whole job gets the j-th draw of mt19937(--seed) (quasimap.cpp:136-137)
    seeds = synth.master_seeds(MAP_SEED, n_reads * (rank + 1))[n_reads * rank:]
    return prg, bases, offs, seeds


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(prg, bases, offs, seeds, sample_reads, count_events=False):
    """The oracle port (reference algorithm + containers, OpenMP over reads exactly as
    quasimap.cpp:90) on the host cores, on the first `sample_reads` reads of the workload."""
    from common import Oracle
    cores = os.cpu_count() or 1
    o = Oracle(prg, KMER)
    n = min(sample_reads, offs.size - 1)
    b, of, sd = bases[:int(offs[n])], offs[:n + 1], seeds[:n]
    ev = None
    if count_events:
        m = min(n, 5000)
        o.map(bases[:int(offs[m])], offs[:m + 1], seeds[:m], threads=1, want_states=False, count_events=True)
        ev = o.events()
        ev["reads"] = m
    return o, (b, of, sd), cores, ev


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference
    binary cannot be built here — SDSL/htslib/Boost absent), all host threads, bounded sample/step."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    sample = int(os.environ.get("GQ_REF_SAMPLE", 40_000))
    prg, bases, offs, seeds = make_workload(0, sample)
    o, (b, of, sd), cores, _ = cpu_reference_run(prg, bases, offs, seeds, sample)
    for _ in range(args.warmup):
        o.map(b[:int(of[2000])], of[:2001], sd[:2000], threads=cores, want_states=False)
    t = 0.0
    for _ in range(args.steps):
        t += o.map(b, of, sd, threads=cores, want_states=False)
    v = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "quasimap reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} reads of the workload per step, OpenMP over reads"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "config2: 4.4 Mb random reference + 100k biallelic SNPs (PRG 4.8M symbols), "
                        "1M x 150 bp error-free reads per GPU, kmer_size=10, both strands, --seed 42",
            "reads_per_gpu": N_READS, "read_len": READ_LEN, "kmer_size": KMER,
            "parallelism": f"reads sharded x{n_gpus}, index replicated, 1 NCCL all-reduce of coverage counters/step"
            if n_gpus > 1 else "1 GPU",
            "l2": "256 MB device buffer written between steps (untimed) to flush L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gq")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gramtools_b200 import QuasimapIndex

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries exactly one JSON line: anything libraries write to fd 1 (NCCL prints its version banner
    # there) is sent to stderr, and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line: the NCCL version banner goes to stdout at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prg, bases, offs, seeds = make_workload(rank, N_READS)
    t0 = time.time()
    idx = QuasimapIndex(prg, KMER, device=local)
    build_s = time.time() - t0
    stream = torch.cuda.current_stream()
    idx.set_stream(stream.cuda_stream)

    # pinned host copies for the end-to-end arm
    pb = torch.from_numpy(bases).pin_memory()
    po = torch.from_numpy(offs.view(np.int64)).pin_memory()
    ps = torch.from_numpy(seeds.view(np.int32)).pin_memory()
    hb, ho, hs = pb.numpy(), po.numpy().view(np.uint64), ps.numpy().view(np.uint32)

    counters = None
    if world > 1:
        ptr, n_cnt, _ = idx.device_counters()

        class _Shim:
            __cuda_array_interface__ = {"shape": (n_cnt,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        counters = torch.as_tensor(_Shim(), device=f"cuda:{local}")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reduced = torch.empty_like(counters) if counters is not None else None

    def reduce_counters():
        # one NCCL all-reduce(sum) of allele_sum | grouped singles | per-base over NVLink; the local
        # accumulators stay local (they keep accumulating across batches), the sum lands in `reduced`
        reduced.copy_(counters)
        dist.all_reduce(reduced)

    def step_resident():
        idx.map_resident()
        if counters is not None:
            reduce_counters()

    # ---------------- device-resident arm (`value`) ----------------
    idx.upload(hb, ho, hs)
    for _ in range(args.warmup):
        step_resident()
    idx.reset_coverage()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    search_ms, cov_ms, launches, reruns = 0.0, 0.0, 0, 0
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        barrier()
        ev[s][0].record(stream)
        step_resident()
        ev[s][1].record(stream)
        info = idx.run_info()
        search_ms += info["search_ms"]   # CUDA events inside the library, on the launching stream:
        cov_ms += info["coverage_ms"]    # search phase / classify + coverage phase of this step
        launches += info["launches"] + (1 if counters is not None else 0)
        reruns += info["rerun_strands"]
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    a_sum, p_base, stats = idx.coverage()

    # ---------------- end-to-end arm (`e2e`): host buffers in, counters out, every step -------------
    idx.reset_coverage()
    for _ in range(2):
        idx.map_batch(hb, ho, hs)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for s in range(args.steps):
        idx.map_batch(hb, ho, hs)
        if counters is not None:
            reduce_counters()
        _, _, st = idx.coverage()  # D2H of the step's result: coverage vectors + the five counters
        d2h = (a_sum.size + p_base.size) * 2 + 40 + 24  # uint16 vectors + counters + the call's status words
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (resident + end to end)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = idx.run_info()["h2d_bytes"]

    if rank == 0:
        reads_total = N_READS * world * args.steps
        value = reads_total / (total_ms / 1e3)
        # ---- roofline of the dominant kernel (search_kernel) ----
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        cpu = None
        alg_bytes_per_read = None
        if not args.no_cpu_baseline:
            sample = int(os.environ.get("GQ_REF_SAMPLE", 40_000))
            o, (b, of, sd), cores, evc = cpu_reference_run(prg, bases, offs, seeds, sample, count_events=True)
            dt = o.map(b, of, sd, threads=cores, want_states=False)
            cpu = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample} reads of the workload, OpenMP over reads (oracle port of the reference "
                             "algorithm; the reference binary needs SDSL/htslib/Boost, absent here)"}
            # algorithmic bytes of the search kernel per read: one 32 B sector per rank
            # query + the 2-bit packed read (both strands) + k-mer presence bits + one 16 B seed entry
            q_rank = evc["q_rank"] / evc["reads"]
            alg_bytes_per_read = 32.0 * q_rank + 2 * ((READ_LEN + 3) // 4) + 2 * ((READ_LEN - KMER + 1 + 7) // 8) + 16
        roof = None
        kernels = {"search_ms": search_ms / args.steps, "classify_coverage_ms": cov_ms / args.steps}
        if alg_bytes_per_read is not None:
            per_launch_s = (search_ms / args.steps) / 1e3
            achieved = alg_bytes_per_read * N_READS / per_launch_s / 1e9
            # dram__bytes_read.sum + dram__bytes_write.sum of the search-phase kernels (seed + verify + text +
            # general) for one 1M-read batch, from the committed `ncu --set full` capture (profiles/)
            traffic = SEARCH_PHASE_DRAM_BYTES if N_READS == 1_000_000 else None
            roof = {"bound": "hbm", "kernel": "search phase: seed_kernel + verify_kernel + text_kernel + search_kernel",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_read": alg_bytes_per_read,
                    "kernel_ms_per_launch": search_ms / args.steps, "coverage_kernel_ms": cov_ms / args.steps,
                    "note": "algorithmic bytes are the REFERENCE algorithm's (SURVEY 8d: 32 B per rank query, 2 per state "
                            "per base); width-1 states are walked in the packed PRG text instead (16 bases per 8 B), so "
                            "frac can exceed 1 and measured DRAM traffic is far below the algorithmic bytes"}
        line = {
            "metric": "quasimap reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": reads_total / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
            "stats": {"all_reads": stats.all_reads_count, "skipped": stats.skipped_reads_count,
                      "missing_kmer": stats.missing_kmer_reads_count, "no_extension": stats.no_extension_reads_count,
                      "exact_mapped": stats.exact_mapped_reads_count, "rerun_strands": int(reruns),
                      "index_build_s": build_s},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
