#!/usr/bin/env python3
"""Generate one synthetic data set, run oracle and product on it, report byte parity.

  python tools/parity_run.py --preset c1 --scale 0.05 --samples 12 [--split] [--cov] [--text] [--work DIR]
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
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--samples", type=int, default=12)
    ap.add_argument("--depth", type=float, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--work", default="/tmp/msnv_parity")
    ap.add_argument("--split", action="store_true", help="also run with -l bed_header (split mode)")
    ap.add_argument("--cov", action="store_true", help="also compare qaCompute outputs")
    ap.add_argument("--text", action="store_true", help="also run the product in classic text mode")
    ap.add_argument("--annotation", action="store_true")
    a = ap.parse_args()
    d = a.work
    st = H.synth(d, a.preset, a.scale, a.samples, a.seed, depth=a.depth, annotation=a.annotation)
    print(json.dumps(st))
    ann = os.path.join(d, "annotation.txt") if (a.annotation or a.preset == "c5") else None
    ok = True
    modes = [("unsplit", None)]
    if a.split:
        modes.append(("split", H.bed_header(d, os.path.join(d, "bed_header"))))
    for name, bed in modes:
        t0 = time.time()
        rc, err = H.run_oracle_snpcall(d, os.path.join(d, "oracle_" + name), bed=bed, ann=ann)
        t1 = time.time()
        env = dict(os.environ, MSNV_PERF_JSON=os.path.join(d, "perf.jsonl"))
        rc2, err2 = H.run_product_snpcall(d, os.path.join(d, "gpu_" + name), bed=bed, ann=ann, env=env)
        t2 = time.time()
        print("[%s] oracle rc=%d %.2fs | product rc=%d %.2fs" % (name, rc, t1 - t0, rc2, t2 - t1))
        if rc2 != 0:
            print(err2)
            ok = False
            continue
        for ext in (".called", ".indiv"):
            df = H.first_diff(os.path.join(d, "oracle_" + name + ext), os.path.join(d, "gpu_" + name + ext))
            n = sum(1 for _ in open(os.path.join(d, "oracle_" + name + ext), "rb"))
            print("  %-8s %7d lines  %s" % (ext, n, "IDENTICAL" if not df else "DIFFERENT " + df))
            ok = ok and not df
        if a.text:
            rc3, err3 = H.run_product_snpcall_text(d, os.path.join(d, "txt_" + name), bed=bed, ann=ann)
            for ext in (".called", ".indiv"):
                df = H.first_diff(os.path.join(d, "oracle_" + name + ext), os.path.join(d, "txt_" + name + ext))
                print("  text%-4s rc=%d %s" % (ext, rc3, "IDENTICAL" if not df else "DIFFERENT " + df))
                ok = ok and not df and rc3 == 0
    if a.cov:
        bams = [l.strip() for l in open(os.path.join(d, "all_samples"))][:4]
        for i, b in enumerate(bams):
            r1 = H.run_qacompute(H.oracle_bin("qaCompute_ref"), b, os.path.join(d, "o%d.cov" % i))
            r2 = H.run_qacompute(bin_path("qaCompute"), b, os.path.join(d, "g%d.cov" % i))
            if r2.returncode != 0:
                print("qaCompute rc", r2.returncode, r2.stderr)
                ok = False
                continue
            for ext in ("", ".detail"):
                df = H.first_diff(os.path.join(d, "o%d.cov%s" % (i, ext)), os.path.join(d, "g%d.cov%s" % (i, ext)))
                print("  cov%-8s sample %d %s" % (ext, i, "IDENTICAL" if not df else "DIFFERENT " + df))
                ok = ok and not df
    if os.path.exists(os.path.join(d, "perf.jsonl")):
        print(open(os.path.join(d, "perf.jsonl")).read())
    print("PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
