#!/usr/bin/env python3
"""Debug aid: device counts of a device-generated shard against the numpy recount (tests/pileup_counts.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from metasnv_b200 import abi, harness as H
from pileup_counts import numpy_counts
preset, scale, samples = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
desc = H.describe(preset, scale, samples)
with abi.Context(0) as ctx:
    P, first = ctx.shard_synth(desc)
    ctx.shard_run()
    nbad = 0
    for s in range(desc["n_samples"]):
        e = ctx.export_sample(s)
        if not e["pos"].size:
            continue
        want = numpy_counts(e, P)
        got = ctx.shard_counts(s, 0, P)
        bad = np.argwhere(got != want)
        print("sample", s, "reads", e["pos"].size, "mated", int((e["mate"] >= 0).sum()), "mismatches", bad.shape[0])
        if bad.size and nbad < 2:
            nbad += 1
            p, c = bad[0]
            print("  first mismatch pos %d channel %d: device %s numpy %s" % (p, c, got[p], want[p]))
            pos = e["pos"]; span = e["max_span"]
            idx = np.where((pos <= p) & (pos + span > p))[0]
            for i in idx[:12]:
                print("   read %d pos %d mate %d segs %s" % (i, pos[i], e["mate"][i], [(int(e["seg_pos"][k]), int(e["seg_len"][k])) for k in range(e["seg_off"][i], e["seg_off"][i+1])]))
            ps = np.unique(bad[:, 0])
            print("  mismatching positions:", ps[:40], "... tiles", np.unique(ps // 1024)[:10])
