"""Parity of the CUDA path with the oracle (B200 only: `pytest -m gpu`). Bit-exact for every output:
called_SNPs, indiv_called, cov/*.cov, cov/*.cov.detail and the per-sample per-position counts."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path

pytestmark = pytest.mark.gpu


def _same(a, b):
    d = H.first_diff(a, b)
    assert not d, "%s vs %s: %s" % (a, b, d)


# ---------------------------------------------------------------- committed golden fixtures
@pytest.mark.parametrize("name", ["c1_tiny", "c5_tiny_annotated", "c4_tiny_deep"])
def test_golden_fixture(name, datasets, tmp_path):
    recipe = json.load(open(os.path.join(GOLDEN, name, "recipe.json")))
    data = datasets(recipe["preset"], recipe["scale"], recipe["samples"], **recipe["extra"])
    ann = os.path.join(data, "annotation.txt") if recipe["extra"].get("annotation") else None
    bed = H.bed_header(data, os.path.join(data, "bed_header"))
    for mode, b in (("unsplit", None), ("split", bed)):
        out = str(tmp_path / ("gpu_" + mode))
        rc, err = H.run_product_snpcall(data, out, bed=b, ann=ann)
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(os.path.join(GOLDEN, name, mode + ext), out + ext)
        out = str(tmp_path / ("txt_" + mode))
        rc, err = H.run_product_snpcall_text(data, out, bed=b, ann=ann)      # classic text input, GPU call kernels
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(os.path.join(GOLDEN, name, mode + ext), out + ext)
    for i, bam in enumerate([l.strip() for l in open(os.path.join(data, "all_samples"))][:2]):
        out = str(tmp_path / ("s%d.cov" % i))
        r = H.run_qacompute(bin_path("qaCompute"), bam, out)
        assert r.returncode == 0, r.stderr
        for ext in ("", ".detail"):
            _same(os.path.join(GOLDEN, name, "s%d.cov%s" % (i, ext)), out + ext)


def test_hand_written_case(built, tmp_path):
    """Edge cases in one small SAM: insertion, deletion, clips, N bases, lower-case / N reference, filtered flags,
    orphan, overlapping mates that agree and disagree, low base quality, read starting at position 1."""
    tmp = str(tmp_path)
    bams = []
    for s in ("s1", "s2"):
        out = os.path.join(tmp, s + ".bam")
        subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", s + ".sam"), "--bam", out], check=True)
        bams.append(out)
    lst = os.path.join(tmp, "all_samples")
    open(lst, "w").write("\n".join(bams) + "\n")
    os.symlink(os.path.join(GOLDEN, "hand", "ref.fa"), os.path.join(tmp, "ref.fa"))
    out = os.path.join(tmp, "gpu")
    rc, err = H.run_product_snpcall(tmp, out, c=2, t=2)
    assert rc == 0, err
    _same(os.path.join(GOLDEN, "hand", "expected.called"), out + ".called")
    _same(os.path.join(GOLDEN, "hand", "expected.indiv"), out + ".indiv")
    for s, bam in zip(("s1", "s2"), bams):
        o = os.path.join(tmp, s + ".cov")
        assert H.run_qacompute(bin_path("qaCompute"), bam, o).returncode == 0
        for ext in ("", ".detail"):
            _same(os.path.join(GOLDEN, "hand", "expected_%s.cov%s" % (s, ext)), o + ext)


def test_hand_written_pairing_rules(built, tmp_path):
    """tests/golden/hand/s5_rules.sam (supplementary read, unpaired later mate, the distant-mate exit, equal-quality
    disagreement; expectations derived by hand in tests/test_oracle_cpu.py) through the product, whole and with a BED that
    has two intervals on one contig."""
    tmp = str(tmp_path)
    bams = []
    for s in ("s5_rules", "s2"):
        out = os.path.join(tmp, s + ".bam")
        subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", s + ".sam"), "--bam", out], check=True)
        bams.append(out)
    open(os.path.join(tmp, "all_samples"), "w").write("\n".join(bams) + "\n")
    os.symlink(os.path.join(GOLDEN, "hand", "ref.fa"), os.path.join(tmp, "ref.fa"))
    for mode, bed in (("whole", None), ("bed", os.path.join(GOLDEN, "hand", "rules.bed"))):
        o, g = os.path.join(tmp, "oracle_" + mode), os.path.join(tmp, "gpu_" + mode)
        rc, err = H.run_oracle_snpcall(tmp, o, bed=bed, c=1, t=1)
        assert rc == 0, err
        rc, err = H.run_product_snpcall(tmp, g, bed=bed, c=1, t=1)
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(o + ext, g + ext)
        assert os.path.getsize(o + ".called") + os.path.getsize(o + ".indiv") > 0
    # per-position counts of the first sample against the hand-derived pileup: position 27 of ctgB counts one C (the earlier mate), no A
    dump = os.path.join(tmp, "counts.bin")
    rc, err = H.run_product_snpcall(tmp, os.path.join(tmp, "gpu_counts"), c=1, t=1, env=dict(os.environ, MSNV_DUMP_COUNTS=dump))
    assert rc == 0, err
    hdr = open(dump + ".layout").read().split("\n")
    S, P = int(hdr[0].split("\t")[0]), int(hdr[0].split("\t")[1])
    off = {l.split("\t")[0]: int(l.split("\t")[1]) for l in hdr[1:] if l}
    cnt = np.fromfile(dump, np.uint16).reshape(S, P, 5)
    b = off["ctgB"]
    assert cnt[0, b + 26].tolist() == [0, 1, 0, 0, 0]             # 27 (1-based): C from the earlier mate only
    assert cnt[0, b + 5].tolist() == [0, 0, 0, 3, 0]              # 6: three T (supplementary + both p2 mates, never paired)
    assert cnt[0, b + 12].tolist() == [0, 0, 0, 0, 0]             # 13: nothing
    assert cnt[0, b + 30].tolist() == [1, 0, 0, 0, 0]             # 31: A from the later mate, behind the overlap


# ---------------------------------------------------------------- live oracle on larger seeded inputs
LIVE = [
    ("c1", 0.2, 40, {}, dict()),
    ("c1", 0.05, 12, {"seed": 7}, dict(c=10, t=2, p=0.2)),
    ("c1", 0.05, 12, {"seed": 8}, dict(c=1, t=1, p=0.0)),
    ("c3", 0.0005, 10, {}, dict()),
    ("c4", 0.005, 3, {}, dict()),
    ("c5", 0.004, 6, {"annotation": True}, dict()),
    ("c2", 0.004, 160, {}, dict()),
]


@pytest.mark.parametrize("preset,scale,samples,kw,opts", LIVE, ids=["%s-%s-%d" % (p, s, n) for p, s, n, _, _ in LIVE])
def test_live_parity(preset, scale, samples, kw, opts, datasets, tmp_path):
    data = datasets(preset, scale, samples, **kw)
    ann = os.path.join(data, "annotation.txt") if kw.get("annotation") else None
    bed = H.bed_header(data, os.path.join(data, "bed_header"))
    for mode, b in (("unsplit", None), ("split", bed)):
        o, g = str(tmp_path / ("oracle_" + mode)), str(tmp_path / ("gpu_" + mode))
        rc, err = H.run_oracle_snpcall(data, o, bed=b, ann=ann, **opts)
        assert rc == 0, err
        rc, err = H.run_product_snpcall(data, g, bed=b, ann=ann, **opts)
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(o + ext, g + ext)
    for i, bam in enumerate([l.strip() for l in open(os.path.join(data, "all_samples"))][:3]):
        o, g = str(tmp_path / ("o%d.cov" % i)), str(tmp_path / ("g%d.cov" % i))
        assert H.run_qacompute(H.oracle_bin("qacompute_oracle"), bam, o).returncode == 0
        r = H.run_qacompute(bin_path("qaCompute"), bam, g)
        assert r.returncode == 0, r.stderr
        for ext in ("", ".detail"):
            _same(o + ext, g + ext)


def test_split_files_concatenate_to_whole(datasets, tmp_path):
    """createOptimumSplit-style sharding: per-genome splits called independently (one process each, as metaSNV.py
    does) equal the oracle on the same splits, and no call is lost or duplicated across shards."""
    data = datasets("c1", 0.05, 12)
    bed = H.bed_header(data, os.path.join(data, "bed_header"))
    lines = open(bed).read().splitlines()
    total = 0
    for k, ln in enumerate(lines):                       # one genome (contig) per split
        sp = str(tmp_path / ("best_split_%d" % k))
        open(sp, "w").write(ln + "\n")
        o, g = str(tmp_path / ("o%d" % k)), str(tmp_path / ("g%d" % k))
        assert H.run_oracle_snpcall(data, o, bed=sp)[0] == 0
        rc, err = H.run_product_snpcall(data, g, bed=sp)
        assert rc == 0, err
        _same(o + ".called", g + ".called")
        _same(o + ".indiv", g + ".indiv")
        total += sum(1 for _ in open(g + ".called"))
    whole = str(tmp_path / "whole")
    assert H.run_product_snpcall(data, whole, bed=bed)[0] == 0
    # every split drops its own first pileup line, the single-process run only one: allow that difference
    n_whole = sum(1 for _ in open(whole + ".called"))
    assert n_whole - len(lines) <= total <= n_whole


# ---------------------------------------------------------------- per-position counts (pileup kernel output)
from pileup_counts import oracle_counts as _oracle_counts  # noqa: E402


@pytest.mark.parametrize("preset,scale,samples", [("c1", 0.03, 8), ("c4", 0.002, 2)])
def test_pileup_counts_match_oracle_text(preset, scale, samples, datasets, tmp_path):
    data = datasets(preset, scale, samples)
    dump = str(tmp_path / "counts.bin")
    rc, err = H.run_product_snpcall(data, str(tmp_path / "gpu"), env=dict(os.environ, MSNV_DUMP_COUNTS=dump))
    assert rc == 0, err
    lay = [l.rstrip("\n").split("\t") for l in open(dump + ".layout")]
    S, P = int(lay[0][0]), int(lay[0][1])
    layout = [(n, int(o), int(l)) for n, o, l in lay[1:]]
    got = np.fromfile(dump, np.uint16).reshape(S, P, 5)
    pile = str(tmp_path / "pile.txt")
    with open(pile, "wb") as f:
        subprocess.run([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", os.path.join(data, "ref.fa"), "-B", "-b",
                        os.path.join(data, "all_samples")], stdout=f, check=True)
    want = _oracle_counts(pile, S, layout, P)
    assert want.shape == got.shape
    bad = np.argwhere(want != got)
    assert bad.size == 0, "first mismatch (sample,pos,channel)=%s want %s got %s" % (bad[0], want[tuple(bad[0])], got[tuple(bad[0])])
    assert int(got.sum()) > 0


def test_exotic_cigar_operations_counts_match_pinned_oracle_text(built, tmp_path):
    """tests/golden/hand/s4_ops.sam: =/X, padding, hard and soft clips around indels, zero-length operations,
    one-base segments, mates whose overlap contains an insertion and a deletion, a reference skip inside a pair.
    The device counts of every position must equal the counts of the restatement's (pinned) pileup text."""
    tmp = str(tmp_path)
    bams = []
    for s in ("s4_ops", "s1"):
        out = os.path.join(tmp, s + ".bam")
        subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", s + ".sam"), "--bam", out], check=True)
        bams.append(out)
    open(os.path.join(tmp, "all_samples"), "w").write("\n".join(bams) + "\n")
    os.symlink(os.path.join(GOLDEN, "hand", "ref.fa"), os.path.join(tmp, "ref.fa"))
    dump = os.path.join(tmp, "counts.bin")
    rc, err = H.run_product_snpcall(tmp, os.path.join(tmp, "gpu"), c=2, t=2, env=dict(os.environ, MSNV_DUMP_COUNTS=dump))
    assert rc == 0, err
    lay = [l.rstrip("\n").split("\t") for l in open(dump + ".layout")]
    S, P = int(lay[0][0]), int(lay[0][1])
    layout = [(n, int(o), int(l)) for n, o, l in lay[1:]]
    got = np.fromfile(dump, np.uint16).reshape(S, P, 5)
    want = _oracle_counts(os.path.join(GOLDEN, "hand", "expected_ops.pileup"), S, layout, P)
    bad = np.argwhere(want != got)
    assert bad.size == 0, "first mismatch (sample,pos,channel)=%s want %s got %s" % (bad[0], want[tuple(bad[0])], got[tuple(bad[0])])
    assert int(got[0].sum()) > 60


@pytest.mark.parametrize("form", ["gather", "scatter"])
def test_cigars_that_open_with_a_deletion(form, built, tmp_path):
    """tests/golden/hand/s6_leading_del.sam: reads in order of POS whose first aligned bases are NOT in order (6D10M at 5,
    12M at 6, 2D8M at 7, ...). The gather form of the pileup kernel derives its per-quad segment ranges from coordinate
    order and must notice (it then offers every record to every quad); both forms must give the pinned counts and the
    oracle's calls."""
    tmp = str(tmp_path)
    bams = []
    for s in ("s6_leading_del", "s1"):
        out = os.path.join(tmp, s + ".bam")
        subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", s + ".sam"), "--bam", out], check=True)
        bams.append(out)
    open(os.path.join(tmp, "all_samples"), "w").write("\n".join(bams) + "\n")
    os.symlink(os.path.join(GOLDEN, "hand", "ref.fa"), os.path.join(tmp, "ref.fa"))
    env = dict(os.environ, MSNV_PILEUP=form)
    o, g = os.path.join(tmp, "oracle"), os.path.join(tmp, "gpu")
    rc, err = H.run_oracle_snpcall(tmp, o, c=1, t=1)
    assert rc == 0, err
    rc, err = H.run_product_snpcall(tmp, g, c=1, t=1, env=env)
    assert rc == 0, err
    for ext in (".called", ".indiv"):
        _same(o + ext, g + ext)
    assert os.path.getsize(o + ".called") + os.path.getsize(o + ".indiv") > 0
    dump = os.path.join(tmp, "counts.bin")
    rc, err = H.run_product_snpcall(tmp, os.path.join(tmp, "gpu_counts"), c=1, t=1, env=dict(env, MSNV_DUMP_COUNTS=dump))
    assert rc == 0, err
    lay = [l.rstrip("\n").split("\t") for l in open(dump + ".layout")]
    S, P = int(lay[0][0]), int(lay[0][1])
    layout = [(n, int(o_), int(l)) for n, o_, l in lay[1:]]
    got = np.fromfile(dump, np.uint16).reshape(S, P, 5)
    want = _oracle_counts(os.path.join(GOLDEN, "hand", "expected_leading_del.pileup"), S, layout, P)
    bad = np.argwhere(want != got)
    assert bad.size == 0, "first mismatch (sample,pos,channel)=%s want %s got %s" % (bad[0], want[tuple(bad[0])], got[tuple(bad[0])])
    assert int(got[0].sum()) > 60


def test_empty_inputs(built, tmp_path):
    """BAMs with a header but no reads, alone and next to a populated one."""
    d = str(tmp_path)
    sam = os.path.join(d, "e.sam")
    open(sam, "w").write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ctgA\tLN:40\n@SQ\tSN:ctgB\tLN:40\n")
    e1, e2 = os.path.join(d, "e1.bam"), os.path.join(d, "e2.bam")
    for e in (e1, e2):
        subprocess.run([bin_path("msnv_synth"), "--sam", sam, "--bam", e], check=True)
    full = os.path.join(d, "s1.bam")
    subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", "s1.sam"), "--bam", full], check=True)
    os.symlink(os.path.join(GOLDEN, "hand", "ref.fa"), os.path.join(d, "ref.fa"))
    for name, bams in (("only_empty", [e1, e2]), ("mixed", [e1, full, e2])):
        lst = os.path.join(d, name + ".list")
        open(lst, "w").write("\n".join(bams) + "\n")
        o, g = os.path.join(d, "o_" + name), os.path.join(d, "g_" + name)
        assert H.run_oracle_snpcall(d, o, all_samples=lst, c=1, t=1)[0] == 0
        rc, err = H.run_product_snpcall(d, g, all_samples=lst, c=1, t=1)
        assert rc == 0, err
        _same(o + ".called", g + ".called")
        _same(o + ".indiv", g + ".indiv")


@pytest.mark.parametrize("env", [{"MSNV_INDEX_BITMAP": "1"}, {"MSNV_INDEX_BITMAP": "0"}, {"MSNV_MAX_READS": "16"}, {"MSNV_MAX_READS": "255"},
                                 {"MSNV_CHUNK_Q4": "1280"}, {"MSNV_CHUNK_Q4": "16384"}, {"MSNV_PILEUP_CTAS": "1"},
                                 {"MSNV_MAX_READS": "40", "MSNV_CHUNK_Q4": "1280", "MSNV_PILEUP_CTAS": "2"},
                                 {"MSNV_CONSUMERS": "256"}, {"MSNV_TILE_BUDGET_MB": "1"}],
                         ids=lambda e: "-".join("%s=%s" % kv for kv in e.items()))
def test_kernel_variants_give_identical_output(env, datasets, tmp_path):
    """Every launch-time choice of the library (occupancy bitmap for sparse shards; reads, quads per staged chunk of the
    pileup kernel and therefore how many chunks an item takes; CTAs per SM) must produce the same bytes."""
    for name in ("c1_tiny", "c4_tiny_deep"):
        recipe = json.load(open(os.path.join(GOLDEN, name, "recipe.json")))
        data = datasets(recipe["preset"], recipe["scale"], recipe["samples"], **recipe["extra"])
        out = str(tmp_path / ("gpu_" + name))
        rc, err = H.run_product_snpcall(data, out, env=dict(os.environ, **env))
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(os.path.join(GOLDEN, name, "unsplit" + ext), out + ext)


# ---------------------------------------------------------------- the reference's own vectors through the product's classic mode
_VECTORS = json.load(open(os.path.join(GOLDEN, "snpcall_vectors.json")))


@pytest.mark.parametrize("case", _VECTORS, ids=[c["name"] for c in _VECTORS])
def test_reference_vectors_through_classic_mode(case, built, tmp_path):
    """tests/golden/snpcall_vectors.json (made by the unmodified reference build, SURVEY.md Annex E: token truncation at 10000
    characters, lower-case reference skip, -p, no -i file, annotation, empty input) fed as mpileup TEXT to the product's snpCall:
    host tokeniser (csrc/host/text_pileup.cc) + the GPU call / compaction / gather kernels. Cases the reference answers by
    crashing (negative return codes: abort / segfault on malformed columns) are outside the contract."""
    if case["rc"] != 0:
        pytest.skip("the reference itself fails on this input (rc %d)" % case["rc"])
    tmp = str(tmp_path)
    for fn, txt in case.get("files", {}).items():
        open(os.path.join(tmp, fn), "w").write(txt)
    indiv = os.path.join(tmp, "indiv.txt")
    args = [a if a != "@INDIV" else indiv for a in case["args"]]
    r = subprocess.run([bin_path("snpCall")] + args, input=case["stdin"].encode(), capture_output=True, cwd=tmp)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert r.stdout.decode() == case["stdout"]
    assert (open(indiv).read() if os.path.exists(indiv) else None) == case["indiv"]


# ---------------------------------------------------------------- position windows through the product
@pytest.mark.parametrize("preset,scale,samples,windows", [("c1", 0.1, 12, 5), ("c4", 0.005, 3, 4), ("c3", 0.0005, 10, 7)])
def test_snpcall_in_windows_equals_oracle(preset, scale, samples, windows, datasets, tmp_path):
    """snpCall cut into position windows (decoders resume from window to window, reads that straddle a boundary go to both,
    uploads overlap decoding, two window slots on the device) writes the same bytes as the oracle's single pass - also when the
    count planes of a window must be produced in several ranges of tiles (tiny tile budget)."""
    data = datasets(preset, scale, samples)
    bed = H.bed_header(data, os.path.join(data, "bed_header"))
    for mode, b in (("unsplit", None), ("split", bed)):
        o, g = str(tmp_path / ("oracle_" + mode)), str(tmp_path / ("gpu_" + mode))
        rc, err = H.run_oracle_snpcall(data, o, bed=b)
        assert rc == 0, err
        perf = str(tmp_path / ("perf_" + mode))
        env = dict(os.environ, MSNV_WINDOWS=str(windows), MSNV_TILE_BUDGET_MB="2", MSNV_PERF_JSON=perf)
        rc, err = H.run_product_snpcall(data, g, bed=b, env=env)
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            _same(o + ext, g + ext)
        assert json.loads(open(perf).readline())["windows"] > 1


# ---------------------------------------------------------------- two splits on two GPUs (the shipped multi-GPU path)
def test_two_splits_on_two_gpus(datasets, tmp_path):
    """metaSNV.py's own sharding: one `samtools mpileup -l best_split_k | snpCall -i ...best_split_k` pipe per genome bin, started
    at the same time; snpCall takes GPU k mod #GPUs from the name of its -i file. Needs two GPUs."""
    from metasnv_b200 import abi
    if abi.device_count() < 2:
        pytest.skip("needs two GPUs")
    import threading
    data = datasets("c1", 0.1, 12)
    beds = []
    for k, line in enumerate(open(H.bed_header(data, os.path.join(data, "bed_header")))):
        genome = line.split(".")[0]
        p = str(tmp_path / ("best_split_%d" % (k % 2)))
        open(p, "a").write(line)
        if p not in beds:
            beds.append(p)
    res = {}

    def one(k, bed):
        perf = str(tmp_path / ("perf_%d" % k))
        out = str(tmp_path / ("gpu.best_split_%d" % k))
        ref = os.path.join(data, "ref.fa")
        prod = [bin_path("samtools"), "mpileup", "-f", ref, "-l", bed, "-B", "-b", os.path.join(data, "all_samples")]
        cons = [bin_path("snpCall"), "-f", ref, "-i", str(tmp_path / ("indiv_called.best_split_%d" % k)), "-c", "4", "-t", "4"]
        res[k] = H._pipe(prod, cons, out, env=dict(os.environ, MSNV_PERF_JSON=perf)) + (json.loads(open(perf).readline()) if os.path.exists(perf) else {},)

    th = [threading.Thread(target=one, args=(k, b)) for k, b in enumerate(beds)]
    [t.start() for t in th]
    [t.join() for t in th]
    for k, bed in enumerate(beds):
        rc, err, perf = res[k]
        assert rc == 0, err
        assert perf["device"] == k % abi.device_count()
        o = str(tmp_path / ("oracle_%d" % k))
        rc, err = H.run_oracle_snpcall(data, o, bed=bed)
        assert rc == 0, err
        _same(o + ".called", str(tmp_path / ("gpu.best_split_%d" % k)))
        _same(o + ".indiv", str(tmp_path / ("indiv_called.best_split_%d" % k)))
