#!/usr/bin/env python3
"""Benchmark of the hot path: pileup + call throughput in aligned bases/s on the BASELINE.json
configuration "single 5 Mb genome, 1000 samples at ~10x" (configs[1], "c2").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1..c5|cov]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

One "step" = one pass of index -> pileup (with mate-overlap correction) -> call -> compaction -> gather
over every read of the shard, inputs resident in HBM (`value`; the copy of the hits to the host is timed and
reported beside it). Shards that do not fit the device (c3 per-GPU shard, c5 at full size) are processed
window by window (groups of contigs, msnv_window_*): `value` then sums the windows' kernel times.
`e2e` is measured FROM BAM FILES through the drop-in programs (`samtools` stand-in | `snpCall`): BGZF inflate
and BAM decode on the host cores, upload, kernels, text output - the same boundary the reference arm is
timed at; host decode time is broken out. `e2e_h2d` is the same pass through the C ABI from pinned host
arrays (upload + kernels + hits), which is what the host link allows.
With N > 1 every rank owns one genome shard of the same shape (the sharding createOptimumSplit produces for
N equal genomes; for c3 the N bins of its LPT assignment); there is no collective on the data path (weak
scaling).
The reference arm (`--impl reference`) and the `cpu_baseline` object time the CPU pipe
`mpileup (oracle restatement) | snpCall (unmodified reference build)` on a bounded sample of the same
workload; for workloads with several genomes the unchanged metaSNV.py drives it with --threads = host cores.
Upstream samtools is not available in this image.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, description)
    "c2": ("c2", "single 5 Mb genome, 1000 samples at ~10x, population + individual calling (BASELINE.json configs[1])"),
    "c1": ("c1", "tutorial shape: 3 genomes, 160 samples (BASELINE.json configs[0])"),
    "c4": ("c4", "deep coverage: one 3 Mb genome, 20 samples at ~2000x without the >8000x spikes (BASELINE.json configs[3])"),
    "c3": ("c3", "ProGenomes2 scale: 1000 genomes x 4 Mb in 50 contigs each, 500 samples carrying ~10% of the genomes at 5x; one shard of 8 "
                 "(createOptimumSplit over 8 GPUs) per GPU is selected with --scale 0.125 (BASELINE.json configs[2])"),
    "c5": ("c5", "50 genomes x 3 Mb, 200 samples at ~10x (BASELINE.json configs[4], pileup + call part)"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", delete=False, suffix=".csv")
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------- CPU pipe
def cpu_pipe_once(data, out_prefix):
    """One pass of `mpileup | snpCall` on the CPU over the BAMs in `data`; returns seconds."""
    from metasnv_b200 import harness as H
    ref = os.path.join(data, "ref.fa")
    caller = H.oracle_bin("snpCall_ref") if os.path.exists(H.oracle_bin("snpCall_ref")) else H.oracle_bin("snpcall_oracle")
    prod = [H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", os.path.join(data, "all_samples")]
    cons = [caller] + H.snpcall_args(ref, out_prefix + ".indiv")
    t0 = time.perf_counter()
    rc, err = H._pipe(prod, cons, out_prefix + ".called")
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError("CPU pipe failed: " + err)
    return dt, os.path.basename(caller)


def cpu_sample(workload, work, samples, scale):
    """Bounded sample of the workload as BAM files (same model, same per-sample depth, shorter genome)."""
    from metasnv_b200 import harness as H
    data = os.path.join(work, "cpu_sample")
    st = H.synth(data, WORKLOADS[workload][0], scale=scale, samples=samples)
    return data, st


def cpu_baseline(workload, work, steps=1, scale=None, samples=None):
    # about 3e8 aligned bases: 10-30 s of the CPU pipe
    samples = samples or {"c2": 1000, "c1": 160, "c4": 20, "c3": 500, "c5": 200}[workload]
    scale = scale or {"c2": 0.003, "c1": 0.12, "c4": 0.0012, "c3": 0.0004, "c5": 0.0012}[workload]
    data, st = cpu_sample(workload, work, samples, scale)
    times = []
    caller = ""
    for i in range(steps):
        dt, caller = cpu_pipe_once(data, os.path.join(work, "cpu_out"))
        times.append(dt)
    best = min(times)
    desc = "%s at scale %g (%d samples, %d aligned bases in %d reads) through `oracle mpileup | %s`: 2 processes in a pipe, as metaSNV.py runs one genome" % (
        workload, scale, samples, st["aligned_bases"], st["reads"], caller)
    return {"value": st["aligned_bases"] / best, "unit": "aligned bases/s", "cores": 2,
            "kind": "port", "sample": desc, "seconds": best, "host_cores_available": os.cpu_count()}, st, times


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    work = tempfile.mkdtemp(prefix="msnv_bench_ref_")
    try:
        base, st, times = cpu_baseline(a.workload, work, steps=a.warmup + a.steps if a.ref_full_steps else max(1, min(a.steps, 3)))
        used = times[-min(len(times), a.steps):]
        ms = 1000.0 * sum(used) / len(used)
        value = st["aligned_bases"] / (ms / 1000.0)
        base["value"] = value
        line = {"impl": "reference", "metric": "aligned_bases_per_s", "value": value, "unit": "aligned bases/s", "n_gpus": a.gpus,
                "steps": len(used), "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u16", "data": "synthetic", "config": {"workload": WORKLOADS[a.workload][1], "sample": base["sample"]},
                "cpu_baseline": base,
                "e2e": {"value": value, "unit": "aligned bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return 0


# ---------------------------------------------------------------------------------------------- GPU path
def run_ours(a):
    import numpy as np
    import torch
    from metasnv_b200 import abi
    from metasnv_b200 import harness as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback (use --impl reference for the CPU pipe)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    preset = WORKLOADS[a.workload][0]
    desc = H.describe(preset, a.scale, a.samples, seed=(a.seed + rank) if a.seed else 0)
    if not a.seed and rank:
        desc["seed"] += rank                                   # a different genome per shard
    S = desc["n_samples"]
    genome_len = int(sum(desc["contig_len"]))
    ctx = abi.Context(local)
    t0 = time.perf_counter()
    n_pos, first = ctx.shard_synth(desc)
    t_synth = time.perf_counter() - t0

    # ---- host copies of every sample in pinned memory (end-to-end input), workload statistics
    t0 = time.perf_counter()
    arena = abi.PinnedArena()
    host_samples = []
    n_reads = aligned = h2d_bytes = n_segs = 0
    e2e_on = not a.no_e2e
    e2e_skipped = None
    if e2e_on:
        # The end-to-end leg keeps a host copy of the whole shard (75 GB for the headline shape) per rank. Never drive
        # the box out of memory for it: all ranks agree (min over ranks) on whether the node has room.
        import psutil
        need = 0
        for s in range(S):
            z = ctx.sample_sizes(s)
            need += z.n_reads * 16 + 8 + z.n_segs * 6 + z.n_q4 * 5 + 8 * 256
        if dist is not None:
            dist.barrier()
        ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        room = psutil.virtual_memory().available >= need * ranks_here * 1.1 + 16e9
        if dist is not None:
            flag = torch.tensor([1 if room else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            room = bool(flag.item())
        if not room:
            e2e_on = False
            e2e_skipped = "host memory: %.0f GB per rank x %d ranks needed for the host copies of the shards" % (need / 1e9, ranks_here)
    for s in range(S):
        e = ctx.export_sample(s, arena.alloc if e2e_on else None)
        aligned += int(e["seg_len"].sum(dtype=np.uint64))
        n_reads += e["pos"].size
        n_segs += e["seg_len"].size
        h2d_bytes += sum(v.nbytes for k, v in e.items() if k != "max_span")
        if e2e_on:
            host_samples.append(e)
    ref = ctx.export_ref(n_pos)
    h2d_bytes += ref.nbytes
    t_export = time.perf_counter() - t0
    if first >= 0:
        ctx.shard_mask_position(first)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-input timing
    for _ in range(a.warmup):
        ctx.shard_run(copy=False)
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    per = []
    for _ in range(a.steps):
        ctx.shard_run(copy=False)
        per.append(ctx.timings())
    sync_all()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(t["ms_total"] for t in per) / a.steps
    keys = ("ms_index", "ms_pileup", "ms_call", "ms_compact", "ms_gather")
    kern = {k: sum(t[k] for t in per) / a.steps for k in keys}
    launches = sum(t["kernel_launches"] for t in per)
    items = per[-1]["n_items"]
    hits = ctx.shard_run(copy=True)
    n_hits = hits.n_hits
    d2h_bytes = n_hits * (4 + 20 + 1 + 1 + 10 * S)

    # ---- end to end through the C ABI with host buffers
    e2e_ms = None
    if e2e_on:
        ctx2 = ctx
        e2e_times = []
        for it in range(a.e2e_warmup + a.e2e_steps):
            sync_all()
            t0 = time.perf_counter()
            ctx2.shard_begin(S, ref)
            for s, e in enumerate(host_samples):
                if e["pos"].size:
                    ctx2.shard_add_sample(s, e)
            if first >= 0:
                ctx2.shard_mask_position(first)
            h = ctx2.shard_run(copy=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert h.n_hits == n_hits
            if it >= a.e2e_warmup:
                e2e_times.append(dt)
        e2e_ms = 1000.0 * sum(e2e_times) / len(e2e_times)
    host_samples = None
    pageable_bytes = arena.pageable_bytes
    arena.close()
    ctx.close()

    # ---- reduce over ranks: time = max, work = sum
    step_ms, wall_ms = dev_ms, 1000.0 * wall / a.steps
    tot_aligned, tot_launch, tot_sp, e2e_max = aligned, launches, S * genome_len, e2e_ms or 0.0
    from metasnv_b200.sharding import reduce_over_ranks
    (step_ms, wall_ms, e2e_max), work = reduce_over_ranks(dist, "cuda", [step_ms, wall_ms, e2e_max], [tot_aligned, tot_launch, tot_sp])
    tot_aligned, tot_launch, tot_sp = [int(x) for x in work]
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peak, peak_src = measured_peak()
    sp_active = items * abi.TILE                      # sample-positions whose count tiles are written / read
    # algorithmic bytes: pileup reads 1.25 B per aligned base (the padding of the position-aligned
    # layout is NOT counted) + 12 B per read + 6 B per segment and writes 10 B per active sample-position; the call kernel
    # reads those 10 B again plus 1 B of reference.
    pile_bytes = 1.25 * aligned + 12.0 * n_reads + 6.0 * n_segs + 10.0 * sp_active
    path_bytes = pile_bytes + 10.0 * sp_active + n_pos + n_hits * (8 + 10 * S)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "pileup_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("workload") == a.workload and abs(tj.get("scale", 1.0) - a.scale) < 1e-9 and tj.get("samples", 0) == a.samples:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    achieved = pile_bytes / (kern["ms_pileup"] / 1000.0) / 1e9
    line = {
        "metric": "aligned_bases_per_s", "value": tot_aligned / (step_ms / 1000.0), "unit": "aligned bases/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOADS[a.workload][1], "preset": preset, "scale": a.scale, "samples_per_shard": S,
                   "genome_len_per_shard": genome_len, "shards": world, "reads_per_shard": n_reads, "aligned_bases_per_shard": aligned,
                   "l2": "inputs (%.1f GB per shard) are far larger than L2; nothing is cached between steps" % (h2d_bytes / 1e9),
                   "timing": "CUDA events on the library's stream around each step (max over ranks); wall clock %.3f ms/step" % wall_ms},
        "sample_positions_per_s": tot_sp / (step_ms / 1000.0),
        "gpu_launches": tot_launch,
        "kernels_ms": kern,
        "hits_per_shard": n_hits,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "pileup_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": pile_bytes,
                     "ms_per_launch": kern["ms_pileup"], "frac_of_nominal_8TBs": achieved / 8000.0,
                     "whole_path": {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (dev_ms / 1000.0) / 1e9,
                                    "frac": path_bytes / (dev_ms / 1000.0) / 1e9 / peak}},
        "setup_s": {"device_synth": t_synth, "export_to_pinned_host": t_export},
    }
    if e2e_skipped:
        line["e2e"] = {"value": None, "unit": "aligned bases/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                       "skipped": e2e_skipped}
    if e2e_on:
        line["e2e"] = {"value": tot_aligned / (e2e_max / 1000.0), "unit": "aligned bases/s", "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_max, "steps": a.e2e_steps,
                       "what": "msnv_shard_begin + msnv_shard_add_sample x samples from pinned host arrays + msnv_shard_run (hits copied back)",
                       "pageable_host_bytes": pageable_bytes}
    if world == 1 and not a.no_cpu_baseline:
        work = tempfile.mkdtemp(prefix="msnv_bench_cpu_")
        try:
            line["cpu_baseline"], _, _ = cpu_baseline(a.workload, work)
        finally:
            shutil.rmtree(work, ignore_errors=True)
    emit(line)
    return 0


_json_out = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner with
    printf, ignoring NCCL_DEBUG_FILE): keep a private handle to the real stdout for the JSON line and point file
    descriptor 1 at stderr for everybody else."""
    global _json_out
    if _json_out is None:
        sys.stdout.flush()
        _json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _json_out or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="genome length scale (1.0 = the BASELINE.json shape)")
    ap.add_argument("--samples", type=int, default=0, help="override the sample count (0 = the BASELINE.json shape)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-warmup", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-full-steps", action="store_true", help="reference arm: really run warmup+steps passes (slow)")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    claim_stdout()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
