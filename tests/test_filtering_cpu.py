"""`bin/metaSNV_Filtering` (host-only C++) against the reference's UNCHANGED
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
