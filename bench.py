#!/usr/bin/env python
"""bench.py — quasimap reads/s on B200 (BASELINE.json metric), with roofline, CPU baseline and in-run parity.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4}] [--impl reference]

A step = one pass of the hot path (k-mer seeding + vBWT backward search + k-mer filter + coverage recording,
both strands) over one batch of synthetic reads. The default workload is BASELINE config 2 (4.4 Mb PRG, 100k
biallelic SNPs, 1M x 150 bp reads, kmer_size 10); `--config` selects the other BASELINE shapes. For N > 1 every
rank maps its own shard against a replicated index and the job ends with ONE exchange — the library's own
NCCL all-reduce of the coverage counters + merge of the sparse groups (gq_coverage_allreduce) — inside the timed
region. Before anything is timed, a prefix of the workload is mapped (sharded over the ranks and reduced when
N > 1) and compared bit for bit with the CPU oracle: a mismatch fails the run.
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

BASE_SEED = 0x6772616D  # generator seed = 0x6772616D + config index
MAP_SEED = 42
CONFIGS = {
    1: dict(kind="snp", ref_len=1_000, n_sites=50, n_reads=10_000, read_len=100, k=5, scaling="weak",
            text="config1: toy PRG (1 kb random reference + 50 biallelic SNPs), 10k x 100 bp error-free reads, kmer_size=5"),
    2: dict(kind="snp", ref_len=4_400_000, n_sites=100_000, n_reads=1_000_000, read_len=150, k=10, scaling="weak",
            text="config2: 4.4 Mb random reference + 100k biallelic SNPs (PRG 4.8M symbols), 1M x 150 bp error-free "
                 "reads per GPU, kmer_size=10, both strands, --seed 42"),
    3: dict(kind="nested", n_loci=200, locus_len=5_000, n_reads=5_000_000, read_len=150, k=10, scaling="weak",
            text="config3: nested-variant PRG (200 loci x 5 kb, bracket grammar: nesting depth <= 3, 2-4 pairwise distinct "
                 "alleles per site as make_prg builds them, empty alleles, adjacent sites; 892k symbols, 44k sites), "
                 "5M x 150 bp error-free reads per GPU, kmer_size=10"),
    4: dict(kind="indel", ref_len=250_000_000, n_sites=5_000_000, n_reads=50_000_000, read_len=150, k=11,
            scaling="strong",
            text="config4: 250 Mb random reference + 5M SNP/indel sites (80/10/10 %), 50M x 150 bp error-free reads "
                 "sharded over the GPUs, kmer_size=11"),
}
# config 2 keeps the module-level names round 1 used (tests, tools)
REF_LEN, N_SITES, N_READS, READ_LEN, KMER = (CONFIGS[2][k] for k in ("ref_len", "n_sites", "n_reads", "read_len", "k"))
GEN_SEED = BASE_SEED + 2
PARITY_READS = 20_000


def env_int(name, default):
    return int(os.environ.get(name, default))


def make_prg(config):
    """-> (prg, haplotypes)"""
    from gramtools_b200 import synth
    c, gs = CONFIGS[config], BASE_SEED + config
    if c["kind"] == "snp":
        prg, ref, pos, alt = synth.make_snp_prg(c["ref_len"], c["n_sites"], gs)
        return prg, synth.snp_haplotypes(ref, pos, alt, 8, gs + 1)
    if c["kind"] == "nested":
        prg = synth.make_nested_prg(c["n_loci"], c["locus_len"], gs, distinct=True)
        rng = np.random.default_rng(3)
        return prg, [synth.random_haplotype(prg, rng) for _ in range(8)]
    prg, ref, sites = synth.make_indel_prg_np(c["ref_len"], c["n_sites"], gs)
    return prg, synth.indel_haplotypes(ref, sites, 4, gs + 1)


def make_reads(config, haps, rank, n_reads, first_read=None):
    """Reads of `rank` (weak scaling: its own n_reads; strong: reads [first_read, first_read + n_reads) of the job)
    and their selection seeds: read j of the whole job gets the j-th draw of mt19937(--seed) (quasimap.cpp:136-137)."""
    from gramtools_b200 import synth
    c, gs = CONFIGS[config], BASE_SEED + config
    chunk = 1_000_000
    if n_reads <= chunk:
        bases, offs = synth.sample_reads(haps, n_reads, c["read_len"], gs + 100 + rank)
    else:  # bounded host memory: a million reads at a time
        parts, offs = [], [np.zeros(1, dtype=np.uint64)]
        for i, r0 in enumerate(range(0, n_reads, chunk)):
            b, o = synth.sample_reads(haps, min(chunk, n_reads - r0), c["read_len"], gs + 100 + 1000 * rank + i)
            parts.append(b)
            offs.append(o[1:] + offs[-1][-1])
        bases, offs = np.concatenate(parts), np.concatenate(offs)
    j0 = n_reads * rank if first_read is None else first_read
    seeds = synth.master_seeds(MAP_SEED, j0 + n_reads)[j0:]
    return bases, offs, seeds


def make_workload(rank, n_reads, config=2):
    """(prg, bases, offsets, seeds) of one rank — config 2 by default, as round 1 generated it."""
    prg, haps = make_prg(config)
    bases, offs, seeds = make_reads(config, haps, rank, n_reads)
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


def workload_config(config, n_gpus, reads_per_gpu):
    c = CONFIGS[config]
    return {"workload": c["text"], "config_index": config, "reads_per_gpu": reads_per_gpu, "read_len": c["read_len"],
            "kmer_size": c["k"],
            "parallelism": (f"reads sharded x{n_gpus} ({c['scaling']} scaling), index replicated, one in-library NCCL "
                            "all-reduce of the coverage counters + merge of sparse groups at the end of the job "
                            "(inside the timed region)") if n_gpus > 1 else "1 GPU",
            "l2": "between steps (untimed): a 256 MB device buffer is written (flushes L2), then another 256 MB buffer is "
                  "read, so that the step starts with a cold L2 that holds no dirty lines of the flush itself (their "
                  "write-back otherwise lands on the step's first kernel: +0.06 ms)"}


def reference_sample(config):
    """reads per step of the CPU arm: ~10-30 s of host work per run of 20 steps"""
    return int(os.environ.get("GQ_REF_SAMPLE", {1: 10_000, 2: 40_000, 3: 10_000, 4: 20_000}[config]))


def oracle_ok(config):
    # the oracle builds its index with std containers: minutes beyond ~20 Mb (config 4: ~40 min) — not run there
    return config != 4


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference binary cannot
    be built here — SDSL/htslib/Boost absent), all host threads, a bounded sample of the workload per step."""
    if env_int("RANK", 0) != 0:
        return
    config = args.config
    if not oracle_ok(config):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU port builds its index in ~40 min at config 4's size"}))
        return
    from common import Oracle
    c = CONFIGS[config]
    sample = min(reference_sample(config), c["n_reads"])
    prg, haps = make_prg(config)
    bases, offs, seeds = make_reads(config, haps, 0, sample, first_read=0)
    cores = os.cpu_count() or 1
    o = Oracle(prg, c["k"])
    m = min(2000, sample)
    for _ in range(args.warmup):
        o.map(bases[:int(offs[m])], offs[:m + 1], seeds[:m], threads=cores, want_states=False)
    t = 0.0
    for _ in range(args.steps):
        t += o.map(bases, offs, seeds, threads=cores, want_states=False)
    v = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "quasimap reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": c["scaling"], "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(config, args.gpus, c["n_reads"]),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} reads of the workload per step, OpenMP over reads"},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def load_profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def coverage_result(idx):
    """(allele_sum, per_base, grouped, stats list) of a handle"""
    a, p, st = idx.coverage()
    return a, p, idx.grouped(), [st.all_reads_count, st.skipped_reads_count, st.missing_kmer_reads_count,
                                 st.no_extension_reads_count, st.exact_mapped_reads_count]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gq")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (developer runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # torchrun sets OMP_NUM_THREADS=1 for every rank unless the caller chose a value: the index build (host, OpenMP)
    # then runs on one thread (config 4: 300 s instead of ~60). Give every rank its share of the host's cores instead.
    n_local = env_int("LOCAL_WORLD_SIZE", env_int("WORLD_SIZE", 1))
    if n_local > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // n_local))
    import torch
    import torch.distributed as dist
    from gramtools_b200 import QuasimapIndex, comm_unique_id, pack_reads
    from gramtools_b200.distributed import bind_to_device_numa, shard_bounds

    config = args.config
    cfg = CONFIGS[config]
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries exactly one JSON line: anything libraries write to fd 1 (NCCL prints its version banner
    # there) is sent to stderr, and the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    # one process per GPU: keep each rank's threads and pinned buffers on its GPU's NUMA node (the end-to-end arm is
    # bound by eight ranks pulling reads from host memory at once); the CPU legs get the full affinity back
    full_affinity = bind_to_device_numa(local) if (world > 1 and not os.environ.get("GQ_NO_NUMA_BIND")) else None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # reads of this rank: weak scaling = n_reads per GPU, strong = the job's reads split over the ranks
    if cfg["scaling"] == "strong":
        total = args.reads * world if args.reads else cfg["n_reads"]
        lo, hi = shard_bounds(total, rank, world)
        n_reads, first = hi - lo, lo
    else:
        n_reads, first = (args.reads or cfg["n_reads"]), None
    prg, haps = make_prg(config)
    bases, offs, seeds = make_reads(config, haps, rank, n_reads, first_read=first)
    t0 = time.time()
    idx = QuasimapIndex(prg, cfg["k"], device=local)
    build_s = time.time() - t0
    # One non-default stream carries everything that is timed — the L2 flush, the timing events and the library's
    # kernels — so that the events really bracket the kernels. (torch's default stream has handle 0, which
    # gq_set_stream reads as "the library's own stream": flush and kernels then ran on two unordered streams and the
    # step could start while the flush was still in flight.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    idx.set_stream(stream.cuda_stream)
    for kv in os.environ.get("GQ_OPTIONS", "").split(","):  # developer switch: library tunables (A/B runs)
        if "=" in kv:
            idx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if world > 1:
        # the library's own communicator (ncclCommInitRank); the id travels over torch.distributed
        id_t = torch.from_numpy(comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(id_t, 0)
        idx.comm_init(id_t.cpu().numpy(), rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- parity before timing: a prefix of the job against the CPU oracle ----------------
    parity = None
    oracle = None
    if oracle_ok(config) and not args.no_cpu_baseline:
        m = min(PARITY_READS, cfg["n_reads"])
        pb_, po_, ps_ = make_reads(config, haps, 0, m, first_read=0)  # the same prefix on every rank
        lo, hi = shard_bounds(m, rank, world)
        idx.map_batch(pb_[int(po_[lo]):int(po_[hi])], po_[lo:hi + 1] - po_[lo], ps_[lo:hi])
        if world > 1:
            idx.coverage_allreduce()
        got = coverage_result(idx)
        if rank == 0:
            from common import Oracle
            if full_affinity:
                os.sched_setaffinity(0, full_affinity)
            oracle = Oracle(prg, cfg["k"])
            oracle.map(pb_, po_, ps_, threads=os.cpu_count() or 1, want_states=False)
            ref = oracle.result(want_states=False)
            ok = (np.array_equal(got[0], ref.allele_sum) and np.array_equal(got[1], ref.per_base)
                  and np.array_equal(got[2], ref.grouped) and got[3] == ref.stats)
            if full_affinity:
                bind_to_device_numa(local)
            parity = {"ok": bool(ok), "reads": m, "ranks": world,
                      "checked": "allele_sum, per-base, grouped allele counts, 5 counters vs the CPU oracle"
                                 + (" after the in-library NCCL all-reduce" if world > 1 else "")}
        flag = torch.tensor([1 if (rank != 0 or parity["ok"]) else 0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            raise SystemExit("bench.py: PARITY FAILURE — the GPU coverage of the prefix differs from the oracle's")
        idx.reset_coverage()
    properties = None

    # pinned host copies for the end-to-end arms
    pbases = torch.from_numpy(bases).pin_memory()
    poffs = torch.from_numpy(offs.view(np.int64)).pin_memory()
    pseeds = torch.from_numpy(seeds.view(np.int32)).pin_memory()
    hb, ho, hs = pbases.numpy(), poffs.numpy().view(np.uint64), pseeds.numpy().view(np.uint32)
    pk, pw, pl = pack_reads(hb, ho)
    pk_t, pw_t, pl_t = (torch.from_numpy(x.view(np.int32)).pin_memory() for x in (pk, pw, pl))
    hk, hw, hl = (t.numpy().view(np.uint32) for t in (pk_t, pw_t, pl_t))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")  # 256 MB, only ever read

    def flush_l2(s):
        flush.fill_(s & 0xFF)   # write a buffer larger than L2 ...
        return flush_r.max()    # ... then read one: cold AND clean (the dirty lines are written back here, untimed)


    # ---------------- device-resident arm (`value`) ----------------
    idx.upload(hb, ho, hs)
    for _ in range(args.warmup):
        idx.map_resident()
    if world > 1:
        idx.coverage_allreduce()  # warm the communicator too: NCCL sets its channels up on the first collective
    # per-kernel durations come from a separate, untimed pass in which classify and coverage run one after the other
    # (in the timed steps they share the GPU on two streams, so their own durations overlap)
    idx.set_option("overlap_classify", 0)
    kernel_ms = {}
    KSTEPS = 5
    for s in range(KSTEPS + 1):
        flush_l2(s)
        idx.map_resident()
        if s:  # the first pass warms up
            for k_, v_ in idx.kernel_ms().items():
                kernel_ms[k_] = kernel_ms.get(k_, 0.0) + v_ / KSTEPS
    idx.set_option("overlap_classify", 1)
    idx.reset_coverage()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    launches, reruns = 0, 0
    search_ms = cov_ms = 0.0
    barrier()
    for s in range(args.steps):
        flush_l2(s)
        ev[s][0].record(stream)
        idx.map_resident()
        ev[s][1].record(stream)
        info = idx.run_info()
        search_ms += info["search_ms"]
        cov_ms += info["coverage_ms"]
        launches += info["launches"]
        reruns += info["rerun_strands"]
    # the job's one exchange: coverage of all ranks summed in place (north_star: "a single NCCL allreduce ... at the end")
    ev[args.steps][0].record(stream)
    if world > 1:
        idx.coverage_allreduce()
        launches += 2  # group export + import kernels (the NCCL kernels are the library's)
    ev[args.steps][1].record(stream)
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in ev[:args.steps])
    reduce_ms = ev[args.steps][0].elapsed_time(ev[args.steps][1])
    total_ms = step_ms + reduce_ms
    if world > 1:
        t = torch.tensor([total_ms, reduce_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, reduce_ms = float(t[0].item()), float(t[1].item())
    a_sum, p_base, stats = idx.coverage()
    if parity is None:
        # no oracle at this size: size-independent properties of the whole job instead — every error-free read maps
        # on at least one strand, the five counters add up, and every mapped strand was recorded (allele_sum holds
        # at least one count per site crossing, per-base coverage is non-empty)
        status = idx.batch_status().reshape(-1, 2)
        mapped_reads = int(((status == 3).sum(axis=1) >= 1).sum())
        tot = stats.skipped_reads_count + stats.missing_kmer_reads_count + stats.no_extension_reads_count + stats.exact_mapped_reads_count
        properties = {"every_read_maps": mapped_reads == n_reads, "counters_add_up": tot == stats.all_reads_count,
                      "coverage_recorded": bool(a_sum.astype(np.int64).sum() > 0 and p_base.astype(np.int64).sum() > 0)}
        if not all(properties.values()):
            raise SystemExit(f"bench.py: PROPERTY FAILURE {properties}")

    # ---------------- end-to-end arm (`e2e`): host buffers in, reduced coverage out, every step -------------
    # step = reset the accumulators, map the batch from pinned HOST buffers (2-bit packed reads: the form a reader
    # thread hands over), reduce over the ranks, fetch the coverage to the host
    def e2e_loop(map_fn):
        for _ in range(2):
            idx.reset_coverage()
            map_fn()
            if world > 1:
                idx.coverage_allreduce()
            idx.coverage()
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            idx.reset_coverage()
            map_fn()
            if world > 1:
                idx.coverage_allreduce()
            idx.coverage()  # D2H of the step's result: coverage vectors + the five counters
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, idx.run_info()["h2d_bytes"]

    e2e_ms, h2d = e2e_loop(lambda: idx.map_batch_packed(hk, hw, hl, hs))
    e2e_u8_ms, h2d_u8 = e2e_loop(lambda: idx.map_batch(hb, ho, hs))
    d2h = (a_sum.size + p_base.size) * 2 + 40 + 24  # uint16 vectors + counters + the call's status words
    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed regions (resident + end to end)

    if rank == 0:
        job_reads = n_reads * world if cfg["scaling"] == "weak" else (args.reads * world if args.reads else cfg["n_reads"])
        reads_total = job_reads * args.steps
        value = reads_total / (total_ms / 1e3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # ---- CPU baseline beside it: oracle port, all host cores and one thread, bounded samples ----
        cpu = cpu1 = None
        ref_alg_bytes = None
        if oracle is not None:
            if full_affinity:
                os.sched_setaffinity(0, full_affinity)
            cores = os.cpu_count() or 1
            sample = min(reference_sample(config), n_reads)
            b, of, sd = bases[:int(offs[sample])], offs[:sample + 1], seeds[:sample]
            dt = float(np.median([oracle.map(b, of, sd, threads=cores, want_states=False) for _ in range(5)]))
            cpu = {"value": sample / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample} reads of the workload, median of 5 passes, OpenMP over reads (oracle port of "
                             "the reference algorithm; the reference binary needs SDSL/htslib/Boost, absent here)"}
            s1 = max(1000, sample // 10)
            dt1 = oracle.map(bases[:int(offs[s1])], offs[:s1 + 1], seeds[:s1], threads=1, want_states=False, count_events=True)
            evc = oracle.events()
            cpu1 = {"value": s1 / dt1, "unit": "reads/s", "cores": 1, "kind": "port", "sample": f"first {s1} reads, 1 thread"}
            # SURVEY §8d's figure for the REFERENCE algorithm (one 32 B sector per rank query, 2 per state per base)
            L, k = cfg["read_len"], cfg["k"]
            ref_alg_bytes = 32.0 * evc["q_rank"] / s1 + 2 * ((L + 3) // 4) + 2 * ((L - k + 1 + 7) // 8) + 16
        # ---- roofline of the dominant kernel: byte model of THIS algorithm (profiles/r02_byte_model.json:
        # distinct 32 B sectors each thread-sized unit of work touches per structure, from the instrumented host
        # emulation of the device functions) over the kernel's CUDA-event duration in this run ----
        model = (load_profile_json("r02_byte_model.json") or {}).get(f"config{config}")
        traffic = (load_profile_json("r02_dram_traffic.json") or {}).get(f"config{config}")
        kms = dict(kernel_ms)
        phases, roof = [], None
        if model and any(kms.values()):
            for name in QuasimapIndex.KERNELS:
                ms = kms.get(name, 0.0)
                mb = model["bytes_per_read"].get(name, 0.0)
                ach = mb * n_reads / (ms / 1e3) / 1e9 if ms > 0 else 0.0
                tr = traffic["dram_bytes_per_read"].get(name) if traffic else None
                phases.append({"kernel": name, "ms": ms, "algorithmic_bytes_per_read": mb, "achieved_GBps": ach,
                               "frac": ach / peak, "dram_bytes_per_read_ncu": tr})
            dom = max(phases, key=lambda p_: p_["ms"])
            whole = sum(p_["algorithmic_bytes_per_read"] for p_ in phases)
            roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"],
                    "traffic": int(dom["dram_bytes_per_read_ncu"] * n_reads) if dom["dram_bytes_per_read_ncu"] else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_read": dom["algorithmic_bytes_per_read"],
                    "kernel_ms_per_launch": dom["ms"], "reads_per_launch": n_reads,
                    "whole_step": {"algorithmic_bytes_per_read": whole, "ms": step_ms / args.steps,
                                   "frac": whole * n_reads / (step_ms / args.steps / 1e3) / 1e9 / peak},
                    "reference_algorithm_bytes_per_read": ref_alg_bytes,
                    "reference_algorithm_bytes_per_s": (ref_alg_bytes * n_reads / (search_ms / args.steps / 1e3)
                                                        if ref_alg_bytes and search_ms else None),
                    "note": "bytes = this algorithm's own model (text-mode walk), not the reference's rank-query bytes; "
                            "the kernels are bound by the latency of dependent sector loads, not by bandwidth"}
        line = {
            "metric": "quasimap reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(config, world, n_reads), "clocks": clocks, "parity": parity["ok"] if parity else None,
            "parity_check": parity if parity else {"ok": None, "why": "the CPU oracle builds its index in ~40 min at this size; "
                                                   "size-independent properties checked instead (bit parity of this site "
                                                   "mix at the same k-mer frequency: tests/test_gpu_parity.py::test_config4_shape_prefix)",
                                                   "properties": properties},
            "e2e": {"value": reads_total / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                    "input": "2-bit packed reads + word offsets + lengths + seeds in pinned host memory (gq_map_batch_packed); "
                             "each step: reset, map, " + ("NCCL all-reduce, " if world > 1 else "") + "fetch coverage to the host"},
            "e2e_u8": {"value": reads_total / (e2e_u8_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(h2d_u8),
                       "ms_per_step": e2e_u8_ms / args.steps, "input": "1 byte per base (gq_map_batch), packed on the GPU"},
            "numa_bound": bool(full_affinity),
            "gpu_launches": int(launches),
            "roofline": roof, "phases": phases, "cpu_baseline": cpu, "cpu_baseline_1t": cpu1,
            "kernels": {"search_ms": search_ms / args.steps, "classify_coverage_ms": cov_ms / args.steps,
                        "allreduce_ms_once": reduce_ms},
            "stats": {"all_reads": stats.all_reads_count, "skipped": stats.skipped_reads_count,
                      "missing_kmer": stats.missing_kmer_reads_count, "no_extension": stats.no_extension_reads_count,
                      "exact_mapped": stats.exact_mapped_reads_count, "rerun_strands": int(reruns),
                      "index_build_s": build_s, "index_device_bytes": int(idx.layout.device_bytes)},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
