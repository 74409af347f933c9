"""Drive the drop-in programs and the oracle on the same inputs (used by tests/ and bench.py).

The oracle side (oracle/_ref/*) is test infrastructure: `run_oracle_*` may only be called from
tests, smoke() and bench.py's CPU-baseline legs.
"""
import os
import subprocess
import shutil

from .paths import BIN_DIR, ORACLE_DIR, bin_path

ORACLE_BIN = os.path.join(ORACLE_DIR, "_ref")


def oracle_bin(name):
    return os.path.join(ORACLE_BIN, name)


def synth(out_dir, preset, scale=1.0, samples=0, seed=0, threads=0, depth=None, annotation=False):
    """Write a synthetic data set (ref.fa, bam/*.bam, all_samples[, annotation.txt]) and return its stats."""
    import json
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    cmd = [bin_path("msnv_synth"), "--preset", preset, "--scale", str(scale), "--out", out_dir]
    if samples:
        cmd += ["--samples", str(samples)]
    if seed:
        cmd += ["--seed", str(seed)]
    if threads:
        cmd += ["--threads", str(threads)]
    if depth is not None:
        cmd += ["--depth", str(depth)]
    if annotation:
        cmd += ["--annotation"]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def describe(preset, scale=1.0, samples=0, seed=0, depth=None):
    """The synthetic configuration as a dict (input of abi.Context.shard_synth)."""
    import json
    cmd = [bin_path("msnv_synth"), "--preset", preset, "--scale", str(scale), "--describe"]
    if samples:
        cmd += ["--samples", str(samples)]
    if seed:
        cmd += ["--seed", str(seed)]
    if depth is not None:
        cmd += ["--depth", str(depth)]
    return json.loads(subprocess.run(cmd, check=True, capture_output=True, text=True).stdout)


def _pipe(producer, consumer, out_path, env=None):
    with open(out_path, "wb") as out:
        p1 = subprocess.Popen(producer, stdout=subprocess.PIPE, env=env)
        p2 = subprocess.Popen(consumer, stdin=p1.stdout, stdout=out, stderr=subprocess.PIPE, env=env)
        p1.stdout.close()
        _, err = p2.communicate()
        p1.wait()
    return p2.returncode, err.decode(errors="replace")


def snpcall_args(ref, indiv, ann=None, c=4, t=4, p=None):
    a = ["-f", ref]
    if ann:
        a += ["-g", ann]
    a += ["-i", indiv, "-c", str(c), "-t", str(t)]
    if p is not None:
        a += ["-p", str(p)]
    return a


def run_oracle_snpcall(data_dir, out_prefix, bed=None, ann=None, c=4, t=4, p=None, all_samples=None):
    """oracle mpileup restatement | reference snpCall  ->  <out_prefix>.called / .indiv"""
    ref = os.path.join(data_dir, "ref.fa")
    lst = all_samples or os.path.join(data_dir, "all_samples")
    prod = [oracle_bin("mpileup_oracle"), "mpileup", "-f", ref] + (["-l", bed] if bed else []) + ["-B", "-b", lst]
    cons = [oracle_bin("snpCall_ref")] + snpcall_args(ref, out_prefix + ".indiv", ann, c, t, p)
    rc, err = _pipe(prod, cons, out_prefix + ".called")
    return rc, err


def run_product_snpcall(data_dir, out_prefix, bed=None, ann=None, c=4, t=4, p=None, all_samples=None, env=None):
    """samtools stand-in | snpCall (GPU)  ->  <out_prefix>.called / .indiv"""
    ref = os.path.join(data_dir, "ref.fa")
    lst = all_samples or os.path.join(data_dir, "all_samples")
    prod = [bin_path("samtools"), "mpileup", "-f", ref] + (["-l", bed] if bed else []) + ["-B", "-b", lst]
    cons = [bin_path("snpCall")] + snpcall_args(ref, out_prefix + ".indiv", ann, c, t, p)
    return _pipe(prod, cons, out_prefix + ".called", env=env)


def run_product_snpcall_text(data_dir, out_prefix, bed=None, ann=None, c=4, t=4, p=None):
    """oracle mpileup TEXT | product snpCall (classic mode: host parse + GPU call kernels)."""
    ref = os.path.join(data_dir, "ref.fa")
    lst = os.path.join(data_dir, "all_samples")
    prod = [oracle_bin("mpileup_oracle"), "mpileup", "-f", ref] + (["-l", bed] if bed else []) + ["-B", "-b", lst]
    cons = [bin_path("snpCall")] + snpcall_args(ref, out_prefix + ".indiv", ann, c, t, p)
    return _pipe(prod, cons, out_prefix + ".called")


def run_qacompute(binary, bam, out):
    return subprocess.run([binary, "-c", "10", "-d", "-i", bam, out], capture_output=True, text=True)


def first_diff(a_path, b_path, width=300):
    """Human-readable description of the first differing line of two text files ('' if identical)."""
    with open(a_path, "rb") as fa, open(b_path, "rb") as fb:
        n = 0
        while True:
            la, lb = fa.readline(), fb.readline()
            n += 1
            if la != lb:
                return "line %d:\n  A: %r\n  B: %r" % (n, la[:width], lb[:width])
            if not la:
                return ""


def bed_header(data_dir, out_path):
    """The bed_header file metaSNV.py:81-94 derives from `samtools view -H` of the first BAM."""
    first = open(os.path.join(data_dir, "all_samples")).readline().rstrip()
    txt = subprocess.check_output([bin_path("samtools"), "view", "-H", first]).decode()
    with open(out_path, "w") as f:
        for line in txt.split("\n")[1:]:
            cols = line.rstrip().split("\t")
            if len(cols) != 3 or cols[0] != "@SQ":
                continue
            f.write(cols[1].replace("SN:", "") + "\t1\t" + cols[2].replace("LN:", "") + "\n")
    return out_path


def stage_metasnv(tree, mode):
    """Build a directory that looks like a metaSNV checkout to the reference's UNCHANGED metaSNV.py
    (staged by oracle/Makefile into oracle/_ref/metaSNV): the scripts are symlinked, the three worker
    programs are either the product's (mode 'gpu') or the oracle's (mode 'oracle').
    Returns (path to metaSNV.py, env with PATH set so that `samtools` resolves to the chosen stand-in)."""
    src = os.path.join(ORACLE_BIN, "metaSNV")
    if not os.path.exists(os.path.join(src, "metaSNV.py")):
        raise RuntimeError("oracle/_ref/metaSNV is missing: run `make -C oracle` where /root/reference exists")
    if os.path.isdir(tree):
        shutil.rmtree(tree)
    os.makedirs(os.path.join(tree, "src", "qaTools"))
    os.makedirs(os.path.join(tree, "src", "snpCaller"))
    os.makedirs(os.path.join(tree, "shim"))
    os.symlink(os.path.join(src, "metaSNV.py"), os.path.join(tree, "metaSNV.py"))
    for f in ("computeGenomeCoverage.py", "collapse_coverages.py", "createOptimumSplit.py"):
        os.symlink(os.path.join(src, "src", f), os.path.join(tree, "src", f))
    if mode == "gpu":
        os.symlink(bin_path("qaCompute"), os.path.join(tree, "src", "qaTools", "qaCompute"))
        os.symlink(bin_path("snpCall"), os.path.join(tree, "src", "snpCaller", "snpCall"))
        os.symlink(bin_path("samtools"), os.path.join(tree, "shim", "samtools"))
    else:
        qa = oracle_bin("qaCompute_ref") if os.path.exists(oracle_bin("qaCompute_ref")) else oracle_bin("qacompute_oracle")
        sc = oracle_bin("snpCall_ref") if os.path.exists(oracle_bin("snpCall_ref")) else oracle_bin("snpcall_oracle")
        os.symlink(qa, os.path.join(tree, "src", "qaTools", "qaCompute"))
        os.symlink(sc, os.path.join(tree, "src", "snpCaller", "snpCall"))
        os.symlink(oracle_bin("mpileup_oracle"), os.path.join(tree, "shim", "samtools"))
    env = dict(os.environ)
    env["PATH"] = os.path.join(tree, "shim") + os.pathsep + env.get("PATH", "")
    return os.path.join(tree, "metaSNV.py"), env


def run_metasnv(script, env, out_dir, all_samples, ref, threads=1, n_splits=1, db_ann=None):
    import sys
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    cmd = [sys.executable, script, out_dir, all_samples, ref, "--threads", str(threads), "--n_splits", str(n_splits)]
    if db_ann:
        cmd += ["--db_ann", db_ann]
    return subprocess.run(cmd, env=env, capture_output=True, text=True)


def tree_files(root):
    out = {}
    for d, _, fs in os.walk(root):
        for f in fs:
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = p
    return out
