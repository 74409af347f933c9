#!/usr/bin/env python3
"""Regenerate the committed golden vectors under tests/golden/.

Run in the BUILD container (needs /root/reference for oracle/_ref/snpCall_ref and qaCompute_ref):
    python tests/golden/make_golden.py
Everything written here is produced by the reference's own code (compiled unmodified, see
oracle/Makefile) fed either with hand-written mpileup text (SURVEY.md Annex E) or with the text of
the oracle's mpileup restatement on seeded synthetic BAMs. The GPU box has no /root/reference, so the
tests there compare against these files.
"""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from metasnv_b200 import harness as H  # noqa: E402
from metasnv_b200.paths import bin_path  # noqa: E402

REF = H.oracle_bin("snpCall_ref")


def col(b, q="I"):
    return "0\t*\t*" if b == "*" else "%d\t%s\t%s" % (len(b), b, q * len(b))


def pile(lines):
    return "".join("%s\t%d\t%s\t" % (c, p, r) + "\t".join(col(b) for b in bs) + "\n" for c, p, r, bs in lines)


def run_ref(args, stdin, workdir):
    indiv = os.path.join(workdir, "indiv.txt")
    a = [x if x != "@INDIV" else indiv for x in args]
    r = subprocess.run([REF] + a, input=stdin.encode(), capture_output=True, cwd=workdir)
    ind = open(indiv).read() if os.path.exists(indiv) else None
    if os.path.exists(indiv):
        os.unlink(indiv)
    return {"rc": r.returncode, "stdout": r.stdout.decode(), "indiv": ind}


def annex_e(work):
    cases = []
    base = [("g1.x.c1", 1, "A", ["TTTT", "TTTT", "*"]),
            ("g1.x.c1", 2, "A", ["..,,TtTt", "....", "*"]),
            ("g1.x.c1", 3, "C", ["^I.$,+2AGa-1NgGG*", "..", ".."]),
            ("g1.x.c1", 4, "G", ["." * 500, "." * 500 + "aaaa", "," * 100]),
            ("g1.x.c1", 5, "g", ["AAAAA", "cccc", "Nn.."]),
            ("g1.x.c1", 6, "T", ["...", "*", "*"]),
            ("g1.x.c1", 7, "T", ["AC", "GA", "ca"]),
            ("g1.x.c1", 8, "T", ["AAAA" + "." * 396, "*", "*"]),
            ("g1.x.c1", 9, "T", ["AAAA" + "." * 397, "*", "*"])]
    cases.append({"name": "annex_e_1_9", "args": ["-i", "@INDIV"], "stdin": pile(base)})
    cases.append({"name": "thresholds_c6_t5", "args": ["-i", "@INDIV", "-c", "6", "-t", "5"], "stdin": pile(base)})
    cases.append({"name": "fraction_p0.5", "args": ["-i", "@INDIV", "-p", "0.5"], "stdin": pile(base)})
    cases.append({"name": "no_indiv_file", "args": [], "stdin": pile(base)})
    cases.append({"name": "truncation_10000", "args": ["-i", "@INDIV"],
                  "stdin": pile([("c", 1, "A", ["....", "...."]), ("c", 3, "A", ["TTTTT" + "." * 10500, "...."])])})
    cases.append({"name": "empty_stdin", "args": ["-i", "@INDIV"], "stdin": ""})
    cases.append({"name": "lowercase_ref_skip", "args": ["-i", "@INDIV"],
                  "stdin": pile([("c", 1, "a", ["....", "...."]), ("c", 2, "a", ["AAAAaaaa", "...."]), ("c", 3, "t", ["TTTTCCCC", "cc.."]),
                                 ("c", 4, "N", ["ACGTACGTACGT", "acgt"])])})
    # annotation vectors (Annex E 10-14)
    g1 = "ATGGCTAAATTTGGGCCCTGA" + "ACGT" * 5
    g2 = "ATGAAACCCGGGTTTTAG"
    with open(os.path.join(work, "ann_ref.fa"), "w") as f:
        f.write(">g1.x.c1\n%s\n>g2.y.c2\n%s\n" % (g1, g2))
    with open(os.path.join(work, "ann.txt"), "w") as f:
        f.write("gene_id\texternal_id\tsequence_id\ttype\tgene_info\tlength\tstart\tend\tstrand\tstart_codon\tstop_codon\tgc\n")
        rows = [("geneA", "g1.x.c1", 1, 21, "+"), ("geneB", "g1.x.c1", 10, 21, "-"), ("geneC", "g1.x.c1", 25, 36, "-"), ("geneD", "g2.y.c2", 1, 18, "+")]
        for i, (n, s, a, b, st) in enumerate(rows):
            f.write("%d\t%s\t%s\tCDS\t<annotation>\t%d\t%d\t%d\t%s\t\t\t\n" % (i + 1, n, s, b - a + 1, a, b, st))
    ann_lines = [("g1.x.c1", 1, "A", ["....", "...."]),
                 ("g1.x.c1", 5, "C", ["..TTTT", "tt"]), ("g1.x.c1", 6, "T", ["..AAAA", "aa"]), ("g1.x.c1", 12, "T", ["GGGG", ".."]),
                 ("g1.x.c1", 23, "C", ["TTTT.", ".."]), ("g1.x.c1", 30, "C", ["GGGGG", "gg"]), ("g2.y.c2", 3, "G", ["AAAAA", "...."]),
                 ("g2.y.c2", 17, "A", ["CCCCC", "...."])]
    cases.append({"name": "annotation", "args": ["-f", "ann_ref.fa", "-g", "ann.txt", "-i", "@INDIV"], "stdin": pile(ann_lines),
                  "files": {"ann_ref.fa": open(os.path.join(work, "ann_ref.fa")).read(), "ann.txt": open(os.path.join(work, "ann.txt")).read()}})
    for c in cases:
        c.update(run_ref(c["args"], c["stdin"], work))
    json.dump(cases, open(os.path.join(HERE, "snpcall_vectors.json"), "w"), indent=1)
    print("snpcall_vectors.json: %d cases" % len(cases))


HAND_REF = ">ctgA desc\nACGTACGTACgtACGTNCGTACGTACGTACGTACGTACGT\n>ctgB\nTTTTTTTTTTGGGGGGGGGGCCCCCCCCCCAAAAAAAAAA\n"
HAND_SAM = [
    "@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:ctgA\tLN:40", "@SQ\tSN:ctgB\tLN:40",
    # name flag ref pos mapq cigar rnext pnext tlen seq qual
    "r1\t0\tctgA\t1\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII",
    "r2\t16\tctgA\t3\t30\t4M2I4M\t*\t0\t0\tGTACTTGTAC\tIIIIIIIIII",
    "r3\t0\tctgA\t5\t0\t3M2D5M\t*\t0\t0\tACGCGTAC\tIIIIIIII",
    "r4\t0\tctgA\t6\t20\t3S5M2S\t*\t0\t0\tTTTCGTANGG\tIIIIII#III",
    "r5\t1024\tctgA\t6\t20\t8M\t*\t0\t0\tCGTACGTA\tIIIIIIII",
    "r6\t256\tctgA\t6\t20\t8M\t*\t0\t0\tCGTACGTA\tIIIIIIII",
    "r7\t512\tctgA\t6\t20\t8M\t*\t0\t0\tCGTACGTA\tIIIIIIII",
    "r8\t73\tctgA\t7\t20\t8M\t*\t0\t0\tGTACGTAC\tIIIIIIII",
    "p1\t99\tctgA\t10\t40\t10M\t=\t14\t14\tCGTACGTNCG\tIIII5IIIII",
    "p2\t99\tctgA\t12\t40\t10M\t=\t16\t14\tTTCGTNCGTA\t++++++++++",
    "p1\t147\tctgA\t14\t40\t10M\t=\t10\t-14\tCGTNCGTACG\tI&I5IIIIII",
    "p2\t147\tctgA\t16\t40\t10M\t=\t12\t-14\tTTCGTACGTA\t**********",
    "r9\t16\tctgA\t31\t50\t5H10M\t*\t0\t0\tCGTACGTACG\t!!!!IIIIII",
    "s1\t0\tctgB\t1\t60\t12M\t*\t0\t0\tTTTTTATTTTGG\tIIIIIIIIIIII",
    "s2\t0\tctgB\t1\t60\t12M\t*\t0\t0\tTTTTTATTTTGG\tIIIIIIIIIIII",
    "s3\t16\tctgB\t2\t60\t11M\t*\t0\t0\tTTTTATTTTGG\tIIIIIIIIIII",
    "s4\t16\tctgB\t2\t60\t5M3N3M\t*\t0\t0\tTTTTAGGG\tIIIIIIII",
    "s5\t0\tctgB\t29\t60\t12M\t*\t0\t0\tCCAAAAAAAAAA\tIIIIIIIIIIII",
    "u1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII",
]

# CIGAR operations and shapes the synthetic model never produces: =/X, padding, hard + soft clips around indels,
# zero-length operations, one-base segments, mates whose overlap contains an insertion and a deletion,
# a reference skip inside a pair. Checked as per-position counts (tests/test_gpu_parity.py) against the
# restatement's text, which is pinned here.
HAND_OPS = [
    "@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:ctgA\tLN:40", "@SQ\tSN:ctgB\tLN:40",
    "x1\t0\tctgA\t1\t60\t5=1X4=\t*\t0\t0\tACGTAAGTAC\tIIIIIIIIII",
    "x2\t0\tctgA\t2\t60\t3M1P1I3M\t*\t0\t0\tCGTTACG\tIIIIIII",
    "x3\t16\tctgA\t3\t60\t2H2S3M1I2M1D3M2S\t*\t0\t0\tTTGTATCGACGGG\tIIIIIIIIIIIII",
    "x4\t0\tctgA\t4\t60\t4M0I4M\t*\t0\t0\tTACGTACG\tIIIIIIII",
    "q1\t99\tctgA\t8\t60\t4M2D6M\t=\t12\t14\tTACGCGTNCG\tIIIIIIIIII",
    "q1\t147\tctgA\t12\t60\t3M1I6M\t=\t8\t-14\tACGTTNCGAA\t5555555555",
    "q2\t99\tctgA\t20\t60\t10M\t=\t22\t10\tACGTACGTAC\t((((((((((",
    "q2\t147\tctgA\t22\t60\t2S8M\t=\t20\t-10\tGGGTTCGTAC\tIIIIIIIIII",
    "q3\t99\tctgA\t26\t60\t3M3N3M\t=\t30\t10\tGTAGTA\tIIIIII",
    "q3\t147\tctgA\t30\t60\t6M\t=\t26\t-10\tCCTACG\tIIIIII",
    "x6\t0\tctgA\t33\t60\t2=2X2=\t*\t0\t0\tACAAAC\tIIIIII",
    "y1\t0\tctgB\t1\t60\t10M\t*\t0\t0\tTTTTTTTTTA\tIIIIIIIIII",
    "y2\t16\tctgB\t5\t60\t1M1I1M1I1M1D1M1I1M\t*\t0\t0\tTATCTGGAG\tIIIIIIIII",
    "y3\t0\tctgB\t31\t60\t10M\t*\t0\t0\tAAAAACAAAA\tIIIIIIIIII",
]


def hand_case(work):
    d = os.path.join(HERE, "hand")
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d)
    open(os.path.join(d, "ref.fa"), "w").write(HAND_REF)
    # sample 1: all reads; sample 2: only the ctgB reads without the N-skip read (keeps snpCall's alphabet)
    open(os.path.join(d, "s1.sam"), "w").write("\n".join(l for l in HAND_SAM if not l.startswith("s4\t")) + "\n")
    open(os.path.join(d, "s2.sam"), "w").write("\n".join(l for l in HAND_SAM if l[0] == "@" or l[0] in "su" and not l.startswith("s4\t")) + "\n")
    open(os.path.join(d, "s3_refskip.sam"), "w").write("\n".join(HAND_SAM) + "\n")
    open(os.path.join(d, "s4_ops.sam"), "w").write("\n".join(HAND_OPS) + "\n")
    for s in ("s1", "s2", "s3_refskip", "s4_ops"):
        subprocess.run([bin_path("msnv_synth"), "--sam", os.path.join(d, s + ".sam"), "--bam", os.path.join(work, s + ".bam")], check=True)
    lst4 = os.path.join(work, "hand_list4")
    open(lst4, "w").write("%s\n%s\n" % (os.path.join(work, "s4_ops.bam"), os.path.join(work, "s1.bam")))
    txt4 = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", os.path.join(d, "ref.fa"), "-B", "-b", lst4])
    open(os.path.join(d, "expected_ops.pileup"), "wb").write(txt4)
    lst = os.path.join(work, "hand_list")
    open(lst, "w").write("%s\n%s\n" % (os.path.join(work, "s1.bam"), os.path.join(work, "s2.bam")))
    txt = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", os.path.join(d, "ref.fa"), "-B", "-b", lst])
    open(os.path.join(d, "expected.pileup"), "wb").write(txt)
    lst3 = os.path.join(work, "hand_list3")
    open(lst3, "w").write("%s\n" % os.path.join(work, "s3_refskip.bam"))
    txt3 = subprocess.check_output([H.oracle_bin("mpileup_oracle"), "mpileup", "-f", os.path.join(d, "ref.fa"), "-B", "-b", lst3])
    open(os.path.join(d, "expected_refskip.pileup"), "wb").write(txt3)
    r = subprocess.run([REF, "-f", os.path.join(d, "ref.fa"), "-i", os.path.join(d, "expected.indiv"), "-c", "2", "-t", "2"], input=txt,
                       capture_output=True)
    open(os.path.join(d, "expected.called"), "wb").write(r.stdout)
    for i, s in enumerate(("s1", "s2")):
        subprocess.run([H.oracle_bin("qaCompute_ref"), "-c", "10", "-d", "-i", os.path.join(work, s + ".bam"), os.path.join(d, "expected_%s.cov" % s)],
                       check=True, capture_output=True)
    print("hand/: pileup %d lines, called %d lines" % (txt.count(b"\n"), r.stdout.count(b"\n")))


SYNTH_CASES = [
    # name, preset, scale, samples, extra
    ("c1_tiny", "c1", 0.02, 6, {}),
    ("c5_tiny_annotated", "c5", 0.002, 4, {"annotation": True}),
    ("c4_tiny_deep", "c4", 0.002, 2, {}),
]


def synth_cases(work):
    for name, preset, scale, samples, extra in SYNTH_CASES:
        data = os.path.join(work, name)
        H.synth(data, preset, scale, samples, **extra)
        out = os.path.join(HERE, name)
        if os.path.isdir(out):
            shutil.rmtree(out)
        os.makedirs(out)
        ann = os.path.join(data, "annotation.txt") if extra.get("annotation") else None
        bed = H.bed_header(data, os.path.join(data, "bed_header"))
        for mode, b in (("unsplit", None), ("split", bed)):
            rc, err = H.run_oracle_snpcall(data, os.path.join(out, mode), bed=b, ann=ann)
            assert rc == 0, err
        bams = [l.strip() for l in open(os.path.join(data, "all_samples"))][:2]
        for i, b in enumerate(bams):
            r = H.run_qacompute(H.oracle_bin("qaCompute_ref"), b, os.path.join(out, "s%d.cov" % i))
            assert r.returncode == 0
        json.dump({"preset": preset, "scale": scale, "samples": samples, "extra": extra}, open(os.path.join(out, "recipe.json"), "w"))
        print(name, {f: os.path.getsize(os.path.join(out, f)) for f in sorted(os.listdir(out))})


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/snpCall_ref is missing: run `make -C oracle` where /root/reference exists")
    work = "/tmp/msnv_golden_work"
    if os.path.isdir(work):
        shutil.rmtree(work)
    os.makedirs(work)
    annex_e(work)
    hand_case(work)
    synth_cases(work)


if __name__ == "__main__":
    main()
