#!/usr/bin/env python3
"""Tuning aid: one device-generated shard, msnv_shard_run under several MSNV_PILEUP_VARIANT / MSNV_CHUNK_Q4 settings."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metasnv_b200 import abi, harness as H

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="c2"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--samples", type=int, default=0); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--settings", default="default,0,1,2,3,4", help="comma list of <variant|default>[:chunk_q4]")
a = ap.parse_args()
desc = H.describe(a.preset, a.scale, a.samples)
ctx = abi.Context(0)
n_pos, first = ctx.shard_synth(desc)
if first >= 0:
    ctx.shard_mask_position(first)
ref_hits = None
for setting in a.settings.split(","):
    v, _, q = setting.partition(":")
    for k in ("MSNV_PILEUP_VARIANT", "MSNV_CHUNK_Q4"):
        os.environ.pop(k, None)
    if v != "default":
        os.environ["MSNV_PILEUP_VARIANT"] = v
    if q:
        os.environ["MSNV_CHUNK_Q4"] = q
    ctx.shard_run(copy=False)
    ms = []
    for _ in range(a.steps):
        h = ctx.shard_run(copy=False)
        ms.append(ctx.timings()["ms_pileup"])
    if ref_hits is None:
        ref_hits = h.n_hits
    print(json.dumps({"setting": setting, "ms_pileup": sum(ms) / len(ms), "min": min(ms), "hits_equal": h.n_hits == ref_hits}), flush=True)
