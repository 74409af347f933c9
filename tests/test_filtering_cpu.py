"""SURVEY.md 8(f) ranks 1 and 2 (host-only C++). Rank 2, `bin/metaSNV_covSummary`, at the end of this file.
Rank 1: `bin/metaSNV_Filtering` against the reference's UNCHANGED
metaSNV_Filtering.py (staged by oracle/Makefile into oracle/_ref/metaSNV) on project directories that the
unchanged metaSNV.py produced with the CPU oracle binaries. The bar is byte-identical `.freq` files
(north_star allows 1e-6 relative on the frequencies; the C++ emitter reproduces Python's str(float))."""
import os
import shutil
import subprocess
import sys
import time

import pytest

from conftest import ROOT
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path

REF_SCRIPT = os.path.join(H.ORACLE_BIN, "metaSNV", "metaSNV_Filtering.py")


@pytest.fixture(scope="module")
def project(built, tmp_path_factory):
    """Part I of metaSNV on a small synthetic data set (3 genomes, 10 samples), CPU oracle binaries."""
    if not os.path.exists(REF_SCRIPT):
        pytest.skip("oracle/_ref/metaSNV/metaSNV_Filtering.py is missing: run `make -C oracle` where /root/reference exists")
    base = str(tmp_path_factory.mktemp("filtering"))
    data = os.path.join(base, "data")
    H.synth(data, "c1", 0.05, 10)
    script, env = H.stage_metasnv(os.path.join(base, "tree"), "oracle")
    proj = os.path.join(base, "proj")
    r = H.run_metasnv(script, env, proj, os.path.join(data, "all_samples"), os.path.join(data, "ref.fa"), threads=2, n_splits=2)
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(os.path.join(proj, "snpCaller", "called_SNPs.best_split_0")) > 0
    return proj


def _tree(root):
    out = {}
    for d, _, fs in os.walk(root):
        for f in fs:
            out[os.path.relpath(os.path.join(d, f), root)] = open(os.path.join(d, f), "rb").read()
    return out


@pytest.mark.parametrize("opts", [[], ["-c", "8", "-p", "0.9"], ["-d", "10.2", "-b", "99.8", "-m", "1", "--ind"],
                                  ["-c", "1", "-p", "0.0", "--n_threads", "3"], ["-m", "20"]],
                         ids=["defaults", "strict-positions", "few-samples-ind", "lenient-threads", "no-taxon"])
def test_freq_files_identical_to_reference_script(project, opts, tmp_path):
    a, b = str(tmp_path / "ref" / "proj"), str(tmp_path / "new" / "proj")
    shutil.copytree(project, a)
    shutil.copytree(project, b)
    r1 = subprocess.run([sys.executable, REF_SCRIPT, a] + opts, capture_output=True, text=True)
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([bin_path("metaSNV_Filtering"), b] + opts, capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    ta, tb = _tree(os.path.join(a, "filtered")), _tree(os.path.join(b, "filtered"))
    assert sorted(ta) == sorted(tb)
    for k in ta:
        assert ta[k] == tb[k], "%s differs" % k
    if opts in ([], ["-c", "1", "-p", "0.0", "--n_threads", "3"]):
        assert any(k.endswith(".filtered.freq") and len(v) > 200 for k, v in ta.items())


def test_python_float_formatting(tmp_path):
    """The emitter's number formatting over the values a frequency can take: k/n for every n <= 300 plus edge cases."""
    proj = tmp_path / "p"
    (proj / "snpCaller").mkdir(parents=True)
    S = 2
    (proj / "all_samples").write_text("/x/a.bam\n/x/b.bam\n")
    tab = "\ta.bam\tb.bam\nTaxId\tx\tx\nT\t100.0\t100.0\n"
    (proj / "p.all_cov.tab").write_text(tab)
    (proj / "p.all_perc.tab").write_text(tab)
    lines, want = [], ["\ta.bam\tb.bam"]
    pos = 0
    for n in list(range(5, 301)) + [65535, 40000, 9999]:
        for k in sorted(set([0, 1, 2, 3, n // 3, n // 2, n - 1, n])):
            pos += 1
            lines.append("T.c\t-\t%d\tA\t%d|%d\t%d|C|.|%d|%d" % (pos, n, n, 2 * k, k, k))
            want.append("T.c:-:%d:A>C:.\t%s\t%s" % (pos, str(k / n), str(k / n)))
    (proj / "snpCaller" / "called_SNPs.best_split_0").write_text("\n".join(lines) + "\n")
    r = subprocess.run([bin_path("metaSNV_Filtering"), str(proj)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = (proj / "filtered" / "pop" / "T.filtered.freq").read_text().rstrip("\n").split("\n")
    assert got == want


def test_command_line_errors(tmp_path):
    exe = bin_path("metaSNV_Filtering")
    assert subprocess.run([exe], capture_output=True).returncode == 2
    assert subprocess.run([exe, "--bogus", "x"], capture_output=True).returncode == 2
    r = subprocess.run([exe, str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and "No such file" in r.stderr


# ---- bin/metaSNV_covSummary against the UNCHANGED computeGenomeCoverage.py + collapse_coverages.py

def _cov_inputs(project, dst):
    """A project directory holding only qaCompute's files (cov/<bam>.cov, .cov.detail) of `project`."""
    os.makedirs(os.path.join(dst, "cov"))
    n = 0
    for f in sorted(os.listdir(os.path.join(project, "cov"))):
        if f.endswith(".cov") or f.endswith(".cov.detail"):
            shutil.copy(os.path.join(project, "cov", f), os.path.join(dst, "cov", f))
            n += 1
    return n


def test_coverage_summaries_identical_to_reference_scripts(project, tmp_path):
    """The per-sample `.cov.summary` files and the two taxa x samples matrices that metaSNV.py:compute_summary makes with
    S + 1 Python processes, from one host program: byte-identical to what the unchanged scripts wrote into `project` (real
    coverage files of 10 samples x 3 genomes) and to the scripts run again on a hand-made set with several contigs per taxon,
    a taxon that is not covered at all and lengths that make the weighted averages inexact in binary."""
    base = os.path.basename(project)
    new = str(tmp_path / "new" / base)
    assert _cov_inputs(project, new) >= 20
    r = subprocess.run([bin_path("metaSNV_covSummary"), new], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want, got = _tree(project), _tree(new)
    names = [k for k in want if k.endswith(".cov.summary")] + [base + ".all_cov.tab", base + ".all_perc.tab"]
    assert len(names) >= 12
    for k in names:
        assert got[k] == want[k], "%s differs" % k

    # hand-made: run the reference scripts and the program on the same inputs
    scripts = os.path.join(H.ORACLE_BIN, "metaSNV", "src")
    a, b = str(tmp_path / "a" / "hand"), str(tmp_path / "b" / "hand")
    rows = [("101.pA.c1", 1000003, 12.34567, 900001, 800002), ("101.pA.c2", 7, 0.14286, 1, 0), ("101.pA.c3", 333331, 3.00001, 333331, 111110),
            ("202.pB.x", 4999999, 0.00000, 0, 0), ("303.c", 1234567, 107.10101, 1234567, 1234566), ("101.late", 11, 1.09091, 11, 1)]
    for d in (a, b):
        os.makedirs(os.path.join(d, "cov"))
        for s, scale in (("s2.bam", 1.0), ("s10.bam", 0.37), ("S1.bam", 2.5)):
            with open(os.path.join(d, "cov", s + ".cov"), "w") as f:
                f.write("Chromosome\tSeq_lem\tAvg_Cov\n")
                for n, L, avg, x1, x2 in rows:
                    f.write("%s\t%d\t%3.5f\n" % (n, L, avg * scale))
                f.write("\n")
            with open(os.path.join(d, "cov", s + ".cov.detail"), "w") as f:
                for n, L, avg, x1, x2 in rows:
                    f.write("%s\t%d\t%d\t%d\t0\t\n" % (n, L, int(x1 * min(1.0, scale)), int(x2 * min(1.0, scale))))
    for f in sorted(os.listdir(os.path.join(a, "cov"))):
        if f.endswith(".cov"):
            p = os.path.join(a, "cov", f)
            assert subprocess.run([sys.executable, os.path.join(scripts, "computeGenomeCoverage.py"), p, p + ".detail", p + ".summary"]).returncode == 0
    assert subprocess.run([sys.executable, os.path.join(scripts, "collapse_coverages.py"), a]).returncode == 0
    r = subprocess.run([bin_path("metaSNV_covSummary"), b], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ta, tb = _tree(a), _tree(b)
    assert sorted(ta) == sorted(tb) and "hand.all_cov.tab" in ta and len(ta) == 3 * 3 + 2
    for k in ta:
        assert ta[k] == tb[k], "%s differs:\n%s\n---\n%s" % (k, ta[k].decode(), tb[k].decode())
    assert b"101\t" in ta["hand.all_cov.tab"] and ta["hand.all_cov.tab"].split(b"\n")[0] == b"\tS1.bam\ts10.bam\ts2.bam"

    # inputs the scripts die on end the program with exit code 1
    assert subprocess.run([bin_path("metaSNV_covSummary"), str(tmp_path / "nowhere")], capture_output=True).returncode == 1
    bad = str(tmp_path / "bad")
    os.makedirs(os.path.join(bad, "cov"))
    open(os.path.join(bad, "cov", "x.bam.cov"), "w").write("Chromosome\tSeq_lem\tAvg_Cov\n1.a\t10\t1.00000\n")
    open(os.path.join(bad, "cov", "x.bam.cov.detail"), "w").write("1.a\t10\t5\t1\t\n1.b\t10\t5\t1\t\n")
    assert subprocess.run([bin_path("metaSNV_covSummary"), bad], capture_output=True).returncode == 1
