"""The reference's UNCHANGED orchestrator (metaSNV.py + its three helper scripts, staged by oracle/Makefile into
oracle/_ref/metaSNV) drives the product binaries exactly as it drives its own: qaCompute per BAM, `samtools view -H`,
createOptimumSplit, one `samtools mpileup | snpCall` pipe per split. Every file of the project directory must be
byte-identical to the run with the oracle's binaries."""
import json
import os
import subprocess
import sys

import pytest

from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("threads,splits,ann", [(3, 3, False), (1, 1, False), (2, 2, True)])
def test_metasnv_py_drives_the_gpu_path(threads, splits, ann, built, tmp_path):
    if not os.path.exists(os.path.join(H.ORACLE_BIN, "metaSNV", "metaSNV.py")):
        pytest.skip("oracle/_ref/metaSNV not staged (needs /root/reference at build time)")
    data = str(tmp_path / "data")
    if ann:
        st = H.synth(data, "c5", 0.002, 5, annotation=True)
    else:
        st = H.synth(data, "c1", 0.03, 9)
    lst, ref = os.path.join(data, "all_samples"), os.path.join(data, "ref.fa")
    db_ann = os.path.join(data, "annotation.txt") if ann else None
    outs = {}
    for mode in ("oracle", "gpu"):
        script, env = H.stage_metasnv(str(tmp_path / ("tree_" + mode)), mode)
        out = str(tmp_path / "proj")          # same project name in both runs: it appears in file names
        perf = str(tmp_path / ("perf_" + mode + ".jsonl"))
        env["MSNV_PERF_JSON"] = perf
        r = H.run_metasnv(script, env, out, lst, ref, threads=threads, n_splits=splits, db_ann=db_ann)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        keep = str(tmp_path / ("proj_" + mode))
        os.rename(out, keep)
        outs[mode] = H.tree_files(keep)
    # the product's coverage pass leaves one extra file per BAM: where every contig starts in it (cov/<bam>.cov.tidx) ...
    tidx = [f for f in outs["gpu"] if f.endswith(".tidx")]
    assert len(tidx) == sum(1 for _ in open(lst))
    for f in tidx:
        del outs["gpu"][f]
    # ... which lets every split's snpCall seek to its contigs instead of inflating whole files (metaSNV.py:157-165)
    runs = [json.loads(l) for l in open(str(tmp_path / "perf_gpu.jsonl"))]
    calls = [x for x in runs if x.get("tool") == "snpCall"]
    assert len(calls) == splits
    if splits > 1:
        assert all(x["bams_read_through_index"] == x["samples"] for x in calls)
        # every record is parsed by the one split that owns its contig (plus the record that ends a run), not by every split
        n_records = st["reads"] + st["junk"] + st["unmapped"]
        assert sum(x["records"] for x in calls) <= n_records + 2 * splits * calls[0]["samples"]
    assert sorted(outs["oracle"]) == sorted(outs["gpu"])
    assert any(f.startswith("snpCaller/called_SNPs") for f in outs["gpu"])
    n_lines = 0
    for rel in sorted(outs["oracle"]):
        d = H.first_diff(outs["oracle"][rel], outs["gpu"][rel])
        assert not d, "%s: %s" % (rel, d)
        if rel.startswith("snpCaller/called_SNPs"):
            n_lines += sum(1 for _ in open(outs["gpu"][rel]))
    assert n_lines > 0
    # Part II's first step on both trees: the reference's metaSNV_Filtering.py on the oracle tree, the product's
    # bin/metaSNV_Filtering on the GPU tree -> identical filtered/pop/*.freq
    ref_filter = os.path.join(H.ORACLE_BIN, "metaSNV", "metaSNV_Filtering.py")
    if os.path.exists(ref_filter):
        a, b = str(tmp_path / "proj_oracle"), str(tmp_path / "proj_gpu")
        # same project name for both (it is part of the coverage table names)
        os.rename(a, str(tmp_path / "proj"))
        r1 = subprocess.run([sys.executable, ref_filter, str(tmp_path / "proj"), "-m", "2"], capture_output=True, text=True)
        assert r1.returncode == 0, r1.stderr
        os.rename(str(tmp_path / "proj"), a)
        os.rename(b, str(tmp_path / "proj"))
        r2 = subprocess.run([bin_path("metaSNV_Filtering"), str(tmp_path / "proj"), "-m", "2"], capture_output=True, text=True)
        assert r2.returncode == 0, r2.stderr
        os.rename(str(tmp_path / "proj"), b)
        fa, fb = H.tree_files(os.path.join(a, "filtered")), H.tree_files(os.path.join(b, "filtered"))
        assert sorted(fa) == sorted(fb) and len(fa) > 0
        for rel in fa:
            assert not H.first_diff(fa[rel], fb[rel]), rel
