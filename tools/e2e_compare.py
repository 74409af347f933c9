#!/usr/bin/env python3
"""End-to-end comparison on real BAM files: `samtools(stand-in) | snpCall` (GPU, decode included) against
`oracle mpileup | reference snpCall` (CPU) and product qaCompute against the reference build, same box, same files.
Prints one JSON object; host BGZF decode time is reported separately from kernel time (MSNV_PERF_JSON record).

  python tools/e2e_compare.py --preset c1 --scale 0.5 [--samples N] [--work DIR]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metasnv_b200 import harness as H  # noqa: E402
from metasnv_b200.paths import bin_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="c1")
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--work", default="/tmp/msnv_e2e")
    ap.add_argument("--skip-cpu", action="store_true")
    a = ap.parse_args()
    d = a.work
    t0 = time.time()
    st = H.synth(d, a.preset, a.scale, a.samples)
    out = {"preset": a.preset, "scale": a.scale, "synth": st, "synth_s": time.time() - t0, "host_cores": os.cpu_count()}
    bam_bytes = sum(os.path.getsize(l.strip()) for l in open(os.path.join(d, "all_samples")))
    out["bam_bytes"] = bam_bytes
    perf = os.path.join(d, "perf.jsonl")
    if os.path.exists(perf):
        os.unlink(perf)
    # the first CUDA process on a fresh box pays for loading the driver: time the pipe twice, report both
    t0 = time.time()
    rc, err = H.run_product_snpcall(d, os.path.join(d, "gpu"), env=dict(os.environ, MSNV_PERF_JSON=perf))
    out["gpu_pipe_first_process_s"] = time.time() - t0
    assert rc == 0, err
    t0 = time.time()
    rc, err = H.run_product_snpcall(d, os.path.join(d, "gpu"), env=dict(os.environ, MSNV_PERF_JSON=perf))
    out["gpu_pipe_s"] = time.time() - t0
    assert rc == 0, err
    out["gpu_perf"] = json.loads(open(perf).read().strip().splitlines()[-1])
    out["gpu_aligned_bases_per_s"] = st["aligned_bases"] / out["gpu_pipe_s"]
    if not a.skip_cpu:
        t0 = time.time()
        rc, err = H.run_oracle_snpcall(d, os.path.join(d, "cpu"))
        out["cpu_pipe_s"] = time.time() - t0
        assert rc == 0, err
        out["cpu_aligned_bases_per_s"] = st["aligned_bases"] / out["cpu_pipe_s"]
        out["identical"] = all(not H.first_diff(os.path.join(d, "cpu" + e), os.path.join(d, "gpu" + e)) for e in (".called", ".indiv"))
        out["speedup_e2e_from_bam"] = out["cpu_pipe_s"] / out["gpu_pipe_s"]
    bam = open(os.path.join(d, "all_samples")).readline().strip()
    t0 = time.time()
    r = H.run_qacompute(bin_path("qaCompute"), bam, os.path.join(d, "g.cov"))
    out["qacompute_gpu_s"] = time.time() - t0
    t0 = time.time()
    r2 = H.run_qacompute(H.oracle_bin("qaCompute_ref"), bam, os.path.join(d, "o.cov"))
    out["qacompute_ref_s"] = time.time() - t0
    out["qacompute_identical"] = r.returncode == 0 and r2.returncode == 0 and not H.first_diff(os.path.join(d, "g.cov"), os.path.join(d, "o.cov")) \
        and not H.first_diff(os.path.join(d, "g.cov.detail"), os.path.join(d, "o.cov.detail"))
    out["qacompute_bam_bytes"] = os.path.getsize(bam)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
