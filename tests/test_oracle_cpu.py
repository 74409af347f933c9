"""The oracle is only trusted after it is pinned (no GPU needed):
  * reference snpCall (compiled unmodified) == committed golden vectors == C restatement,
  * reference qaCompute == C restatement,
  * the mpileup restatement == the committed hand-reviewed pileup of the hand-written SAM case.
"""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, has_reference_build
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path

VECTORS = json.load(open(os.path.join(GOLDEN, "snpcall_vectors.json")))


def _run_snpcall(binary, case, tmp):
    for fn, txt in case.get("files", {}).items():
        open(os.path.join(tmp, fn), "w").write(txt)
    indiv = os.path.join(tmp, "indiv.txt")
    if os.path.exists(indiv):
        os.unlink(indiv)
    args = [a if a != "@INDIV" else indiv for a in case["args"]]
    r = subprocess.run([binary] + args, input=case["stdin"].encode(), capture_output=True, cwd=tmp)
    ind = open(indiv).read() if os.path.exists(indiv) else None
    return r.returncode, r.stdout.decode(), ind


@pytest.mark.parametrize("case", VECTORS, ids=[c["name"] for c in VECTORS])
def test_snpcall_golden_vectors(case, snpcall_checkers, tmp_path):
    for label, binary in snpcall_checkers:
        rc, out, ind = _run_snpcall(binary, case, str(tmp_path))
        assert rc == case["rc"], label
        assert out == case["stdout"], label
        assert ind == case["indiv"], label


def test_annex_e_known_answers():
    """Spot checks of the committed vectors against the values recorded in SURVEY.md Annex E."""
    c = {v["name"]: v for v in VECTORS}
    assert "g1.x.c1\t-\t2\tA\t8|4|0\t4|T|.|4|0|0\n" in c["annex_e_1_9"]["stdout"]
    assert "g1.x.c1\t-\t5\tg\t5|4|2\t5|A|.|5|0|0,4|C|.|0|4|0\n" in c["annex_e_1_9"]["stdout"]
    assert "g1.x.c1\t-\t8\tT\t400|0|0\t4|A|.|4|0|0\n" in c["annex_e_1_9"]["stdout"]
    assert c["annex_e_1_9"]["indiv"] == "g1.x.c1\t-\t4\tG\t500|504|100\t4|A|.|0|4|0\ng1.x.c1\t-\t9\tT\t401|0|0\t4|A|.|4|0|0\n"
    assert "\t10000|4\t5|T|.|5|0\n" in c["truncation_10000"]["indiv"]
    assert "g1.x.c1\tgeneA\t5\tC\t6|2\t6|T|N[GCT-GTT]|4|2\n" in c["annotation"]["stdout"]
    assert "g1.x.c1\tgeneA\t6\tT\t6|2\t6|A|S[GCT-GCA]|4|2\n" in c["annotation"]["stdout"]
    assert "g1.x.c1\tgeneA\t12\tT\t4|2\t4|G|N[TTT-TTG]|4|0\n" in c["annotation"]["stdout"]
    assert "g2.y.c2\tgeneD\t3\tG\t5|4\t5|A|N[ATG-ATA]|5|0\n" in c["annotation"]["stdout"]
    assert c["empty_stdin"]["rc"] == 0 and c["empty_stdin"]["stdout"] == ""


def _bam_from_sam(name, tmp):
    out = os.path.join(tmp, name + ".bam")
    subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(GOLDEN, "hand", name + ".sam"), "--bam", out], check=True)
    return out


def test_mpileup_restatement_hand_case(built, tmp_path):
    tmp = str(tmp_path)
    lst = os.path.join(tmp, "list")
    open(lst, "w").write("%s\n%s\n" % (_bam_from_sam("s1", tmp), _bam_from_sam("s2", tmp)))
    ref = os.path.join(GOLDEN, "hand", "ref.fa")
    txt = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", lst])
    assert txt == open(os.path.join(GOLDEN, "hand", "expected.pileup"), "rb").read()
    lines = txt.decode().split("\n")
    # hand-derived facts (samtools mpileup semantics, SURVEY.md Annex A)
    assert lines[5].startswith("ctgA\t6\tC\t4\t.,+2tt.^5.\tIIII")        # insertion after the column, read start marker
    assert lines[6].startswith("ctgA\t7\tG\t4\t.,.-2TA.\t")              # deletion announced with reference bases
    assert lines[7].startswith("ctgA\t8\tT\t4\t.,*.\t")                  # deleted base
    assert lines[8].startswith("ctgA\t9\tA\t3\t.,*\t")                   # base quality 2 filtered (-Q 13)
    assert lines[14].startswith("ctgA\t15\tG\t1\t.\tN\t")                # mates agree: 40+5 -> 45, the later mate drops to 0
    assert lines[16].startswith("ctgA\t17\tN\t1\t.\t]\t")                # N read base matches N reference; disagreeing pair p2 keeps 0.8*10 < 13
    assert lines[23].startswith("ctgA\t24\tT\t0\t*\t*\t")                # covered only by a filtered base: line with depth 0
    assert lines[40].startswith("ctgB\t6\tT\t3\tAAa\tIII\t3\tAAa\tIII")  # strand case of mismatches
    # exotic CIGAR operations (=, X, P, H+S around indels, zero-length, one-base segments, indels inside a mate overlap)
    lst4 = os.path.join(tmp, "list4")
    open(lst4, "w").write("%s\n%s\n" % (_bam_from_sam("s4_ops", tmp), _bam_from_sam("s1", tmp)))
    txt4 = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", lst4])
    assert txt4 == open(os.path.join(GOLDEN, "hand", "expected_ops.pileup"), "rb").read()
    l4 = txt4.decode().split("\n")
    assert l4[5].startswith("ctgA\t6\tC\t4\tA.,.\t")                  # the X of 5=1X4= is a mismatch like any other
    assert l4[11].startswith("ctgA\t12\tt\t2\t*^]a\tA5\t")             # q1's deletion meets its mate's first base: no overlap rule on '*'
    # CIGARs that open with a deletion: the reads are in order of POS, their first aligned bases are not (hand-derived:
    # column 11 sees d1 . d2 C d3 . d4 . d5 T d6 . in file order; a read shows '*' with its start marker in its first column)
    lst6 = os.path.join(tmp, "list6")
    open(lst6, "w").write("%s\n%s\n" % (_bam_from_sam("s6_leading_del", tmp), _bam_from_sam("s1", tmp)))
    txt6 = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", lst6])
    assert txt6 == open(os.path.join(GOLDEN, "hand", "expected_leading_del.pileup"), "rb").read()
    l6 = txt6.decode().split("\n")
    assert l6[4].startswith("ctgA\t5\tA\t2\t.^]*\tII\t")               # d1 opens with a deleted base
    assert l6[10].startswith("ctgA\t11\tg\t6\t.C..T.\tIIIIII\t")
    assert l6[16].startswith("ctgA\t17\tN\t5\tAA$.$GC\tIIIII\t")         # d4's last base is N on an N reference: '.'; d5's second segment (after its inner deletion) shows G
    # N operations render '>' / '<'
    lst3 = os.path.join(tmp, "list3")
    open(lst3, "w").write("%s\n" % _bam_from_sam("s3_refskip", tmp))
    txt3 = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", lst3])
    assert txt3 == open(os.path.join(GOLDEN, "hand", "expected_refskip.pileup"), "rb").read()
    assert b"<" in txt3


def test_mpileup_restatement_pairing_rules(built, tmp_path):
    """Hand-derived case for the per-read rules of SURVEY.md Annex A.1 / A.2 that the first hand case does not reach:
    a supplementary read is kept; a later mate whose partner was filtered (duplicate) is never paired (mpos < pos: not
    stored); a pair whose first mate claims a distant mate (|isize| >= 2 l_qseq and mpos >= its end) is not paired even
    though the mates overlap; mates that disagree with EQUAL qualities: the earlier read keeps int(0.8 q), the later drops
    to 0; a BED with two intervals on one contig."""
    tmp = str(tmp_path)
    lst = os.path.join(tmp, "list")
    open(lst, "w").write("%s\n" % _bam_from_sam("s5_rules", tmp))
    ref = os.path.join(GOLDEN, "hand", "ref.fa")
    txt = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-B", "-b", lst])
    assert txt == open(os.path.join(GOLDEN, "hand", "expected_rules.pileup"), "rb").read()
    ln = {l.split("\t")[1]: l for l in txt.decode().split("\n") if l}
    assert ln["1"] == "ctgB\t1\tT\t2\t^].^].\tII"            # the supplementary read (0x800) is piled up like any other
    assert ln["6"] == "ctgB\t6\tT\t3\t..,\tIII"               # p2: mate 1 says its mate is far away -> not paired -> both mates counted at 5..8
    assert ln["11"] == "ctgB\t11\tG\t1\t,\tI"                 # p1 mate 1 is a duplicate: filtered; 13 and 14 have no line at all
    assert "13" not in ln and "14" not in ln
    assert ln["16"] == "ctgB\t16\tG\t1\t,\tI"                 # p1 mate 2 (mpos < pos, partner never stored): full quality
    assert ln["26"] == "ctgB\t26\tC\t1\t.\tI"                 # p3 agree: 20 + 20 = 40 for the earlier mate, the later one drops out
    assert ln["27"] == "ctgB\t27\tC\t1\t.\t1"                 # p3 disagree, equal qualities: the earlier mate keeps int(0.8 * 20) = 16
    assert ln["31"] == "ctgB\t31\tA\t1\t,\t5"                 # behind the overlap the later mate counts again (no '^': its first column was 25)
    bed = os.path.join(GOLDEN, "hand", "rules.bed")
    txt = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref, "-l", bed, "-B", "-b", lst])
    assert txt == open(os.path.join(GOLDEN, "hand", "expected_rules_bed.pileup"), "rb").read()
    assert [l.split("\t")[1] for l in txt.decode().split("\n") if l] == ["3", "4", "5", "6", "21", "22", "23", "24"]     # BED is 0-based half open


def test_view_header(built, tmp_path):
    bam = _bam_from_sam("s1", str(tmp_path))
    a = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "view", "-H", bam])
    b = subprocess.check_output([bin_path("samtools"), "view", "-H", bam])
    assert a == b == b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ctgA\tLN:40\n@SQ\tSN:ctgB\tLN:40\n"


@pytest.fixture(scope="module")
def tiny_data(built, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("tiny"))
    H.synth(os.path.join(d, "c1"), "c1", 0.02, 6)
    H.synth(os.path.join(d, "c5"), "c5", 0.002, 4, annotation=True)
    return d


@pytest.mark.parametrize("name", ["c1_tiny", "c5_tiny_annotated"])
def test_oracle_pipe_matches_committed_golden(name, tiny_data, tmp_path):
    """oracle mpileup | snpCall restatement reproduces the golden files written with the reference build."""
    recipe = json.load(open(os.path.join(GOLDEN, name, "recipe.json")))
    data = os.path.join(tiny_data, recipe["preset"])
    ann = os.path.join(data, "annotation.txt") if recipe["extra"].get("annotation") else None
    ref = os.path.join(data, "ref.fa")
    bed = H.bed_header(data, os.path.join(data, "bed_header"))
    for mode, b in (("unsplit", None), ("split", bed)):
        prod = [H.oracle_bin("mpileup_oracle"), "mpileup", "-f", ref] + (["-l", b] if b else []) + ["-B", "-b", os.path.join(data, "all_samples")]
        out = str(tmp_path / mode)
        rc, err = H._pipe(prod, [H.oracle_bin("snpcall_oracle")] + H.snpcall_args(ref, out + ".indiv", ann), out + ".called")
        assert rc == 0, err
        for ext in (".called", ".indiv"):
            assert not H.first_diff(os.path.join(GOLDEN, name, mode + ext), out + ext), (mode, ext)


@pytest.mark.parametrize("name", ["c1_tiny", "c5_tiny_annotated"])
def test_qacompute_restatement(name, tiny_data, tmp_path):
    recipe = json.load(open(os.path.join(GOLDEN, name, "recipe.json")))
    data = os.path.join(tiny_data, recipe["preset"])
    bams = [l.strip() for l in open(os.path.join(data, "all_samples"))][:2]
    for i, b in enumerate(bams):
        out = str(tmp_path / ("s%d.cov" % i))
        assert H.run_qacompute(H.oracle_bin("qacompute_oracle"), b, out).returncode == 0
        for ext in ("", ".detail"):
            assert not H.first_diff(os.path.join(GOLDEN, name, "s%d.cov%s" % (i, ext)), out + ext)
        if has_reference_build():
            out2 = str(tmp_path / ("r%d.cov" % i))
            assert H.run_qacompute(H.oracle_bin("qaCompute_ref"), b, out2).returncode == 0
            for ext in ("", ".detail"):
                assert not H.first_diff(out + ext, out2 + ext)


def test_hand_case_coverage(built, tmp_path):
    for s in ("s1", "s2"):
        bam = _bam_from_sam(s, str(tmp_path))
        out = str(tmp_path / (s + ".cov"))
        assert H.run_qacompute(H.oracle_bin("qacompute_oracle"), bam, out).returncode == 0
        for ext in ("", ".detail"):
            assert not H.first_diff(os.path.join(GOLDEN, "hand", "expected_%s.cov%s" % (s, ext)), out + ext)
