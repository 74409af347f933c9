#!/usr/bin/env python3
"""Times bin/metaSNV_Filtering against the reference's unchanged metaSNV_Filtering.py on one project directory
(produced here by the unchanged metaSNV.py with the CPU oracle binaries) and checks the outputs are identical."""
import argparse, json, os, shutil, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="c2"); ap.add_argument("--scale", type=float, default=0.004)
ap.add_argument("--samples", type=int, default=400); ap.add_argument("--work", default="/tmp/msnv_filt")
ap.add_argument("--threads", type=int, default=1)
ap.add_argument("--fabricate", type=int, nargs=3, metavar=("LINES", "SAMPLES", "TAXA"),
                help="skip Part I: write a project with LINES random called_SNPs lines per taxon (the size of a full run's output)")
a = ap.parse_args()
shutil.rmtree(a.work, ignore_errors=True)
proj = os.path.join(a.work, "proj")
t0 = time.time()
if a.fabricate:
    import random
    L, S, T = a.fabricate
    rnd = random.Random(1)
    os.makedirs(os.path.join(proj, "snpCaller"))
    names = ["s%04d.bam" % i for i in range(S)]
    open(os.path.join(proj, "all_samples"), "w").write("".join("/data/bam/%s\n" % n for n in names))
    for fn, second, lo, hi in (("proj.all_cov.tab", "Average_cov", 3.0, 14.0), ("proj.all_perc.tab", "Percentage_1x", 30.0, 100.0)):
        with open(os.path.join(proj, fn), "w") as f:
            f.write("\t" + "\t".join(names) + "\nTaxId\t" + "\t".join([second] * S) + "\n")
            for t in range(T):
                f.write("%d\t%s\n" % (100001 + t, "\t".join("%.6f" % rnd.uniform(lo, hi) for _ in range(S))))
    for t in range(T):
        with open(os.path.join(proj, "snpCaller", "called_SNPs.best_split_%d" % t), "w") as f:
            for i in range(L):
                cov = [rnd.randint(0, 25) for _ in range(S)]
                alt = [rnd.randint(0, c) if rnd.random() < 0.3 else 0 for c in cov]
                f.write("%d.synth.c0\t-\t%d\t%s\t%s\t%d|%s|.|%s\n" % (100001 + t, 10 * i + 7, "ACGT"[i & 3], "|".join(map(str, cov)),
                                                                       sum(alt), "CGTA"[i & 3], "|".join(map(str, alt))))
    st = {"samples": S}
else:
    data = os.path.join(a.work, "data")
    st = H.synth(data, a.preset, a.scale, a.samples)
    script, env = H.stage_metasnv(os.path.join(a.work, "tree"), "oracle")
    r = H.run_metasnv(script, env, proj, os.path.join(data, "all_samples"), os.path.join(data, "ref.fa"), threads=4, n_splits=2)
    assert r.returncode == 0, r.stderr
out = {"preset": "fabricated" if a.fabricate else a.preset, "scale": a.scale, "samples": st["samples"], "input_setup_s": time.time() - t0,
       "called_bytes": sum(os.path.getsize(os.path.join(proj, "snpCaller", f)) for f in os.listdir(os.path.join(proj, "snpCaller")) if f.startswith("called")),
       "called_lines": sum(sum(1 for _ in open(os.path.join(proj, "snpCaller", f))) for f in os.listdir(os.path.join(proj, "snpCaller")) if f.startswith("called")),
       "host_cores": os.cpu_count(), "threads": a.threads}
pa, pb = os.path.join(a.work, "ref", "proj"), os.path.join(a.work, "new", "proj")
shutil.copytree(proj, pa); shutil.copytree(proj, pb)
opts = ["--n_threads", str(a.threads)]
t0 = time.time(); r1 = subprocess.run([sys.executable, os.path.join(H.ORACLE_BIN, "metaSNV", "metaSNV_Filtering.py"), pa] + opts, capture_output=True, text=True); out["reference_py_s"] = time.time() - t0
assert r1.returncode == 0, r1.stderr
t0 = time.time(); r2 = subprocess.run([bin_path("metaSNV_Filtering"), pb] + opts, capture_output=True, text=True); out["cpp_s"] = time.time() - t0
assert r2.returncode == 0, r2.stderr
fa = sorted(os.listdir(os.path.join(pa, "filtered", "pop")))
out["freq_files"] = len(fa)
out["freq_bytes"] = sum(os.path.getsize(os.path.join(pa, "filtered", "pop", f)) for f in fa)
out["identical"] = fa == sorted(os.listdir(os.path.join(pb, "filtered", "pop"))) and all(
    open(os.path.join(pa, "filtered", "pop", f), "rb").read() == open(os.path.join(pb, "filtered", "pop", f), "rb").read() for f in fa)
out["speedup"] = out["reference_py_s"] / out["cpp_s"]
print(json.dumps(out))
