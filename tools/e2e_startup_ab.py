#!/usr/bin/env python3
"""A/B of snpCall's start-up on one box: first window decoded while the CUDA context comes up (default) against waiting for the
context first (MSNV_EARLY_DECODE=0). Same BAM set, runs interleaved, wall time of the whole `samtools | snpCall` pipe."""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metasnv_b200 import harness as H

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.05714); ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
work = tempfile.mkdtemp(prefix="msnv_ab_")
data = os.path.join(work, "data")
H.synth(data, "c2", scale=a.scale)
perf = os.path.join(work, "perf.jsonl")
H.run_product_snpcall(data, os.path.join(work, "out"), env=dict(os.environ))          # page cache
for r in range(a.rounds):
    for early in ("1", "0"):
        if os.path.exists(perf):
            os.unlink(perf)
        env = dict(os.environ, MSNV_PERF_JSON=perf, MSNV_EARLY_DECODE=early, MSNV_VERBOSE="1")
        t0 = time.perf_counter()
        rc, err = H.run_product_snpcall(data, os.path.join(work, "out"), env=env)
        dt = time.perf_counter() - t0
        p = json.loads(open(perf).readline())
        tr = [l for l in err.splitlines() if l.startswith("[msnv")]
        print(json.dumps({"early_decode": early, "rc": rc, "wall_s": round(dt, 3), "total_s": p.get("total_s"), "decode_wall_s": p.get("decode_wall_s"),
                          "waiting_for_decode_s": p.get("waiting_for_decode_s"), "trace": [t[6:40] for t in tr if "window" not in t or "first" in t]}), flush=True)
