"""Host-side checks that need no GPU: the C ABI library loads and exports what include/msnv.h declares,
the drop-in programs keep the reference's command-line behaviour, BAM I/O round-trips, and the product
refuses to run without a CUDA device (there is no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT
from metasnv_b200 import harness as H
from metasnv_b200.paths import bin_path, lib_path


def _no_gpu():
    lib = ctypes.CDLL(lib_path())
    lib.msnv_device_count.restype = ctypes.c_int
    return lib.msnv_device_count() == 0


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "msnv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(msnv_[a-z_0-9]+)\s*\(", hdr))
    assert {"msnv_create", "msnv_shard_begin", "msnv_shard_add_sample", "msnv_shard_run", "msnv_call_counts", "msnv_cov_run"} <= names
    lib = ctypes.CDLL(lib_path())
    for n in sorted(names):
        assert hasattr(lib, n), "libmsnv_gpu.so does not export %s" % n
    lib.msnv_abi_version.restype = ctypes.c_int
    assert lib.msnv_abi_version() == 6


def test_abi_python_binding_matches_struct_sizes(built):
    from metasnv_b200 import abi
    assert ctypes.sizeof(abi.SampleReads) == 16 + 8 * 8
    assert ctypes.sizeof(abi.CallParams) == 16
    assert ctypes.sizeof(abi.Hits) == 8 + 6 * 8
    assert ctypes.sizeof(abi.CovBlocks) == 8 + 4 * 8
    assert ctypes.sizeof(abi.Timings) == 7 * 4 + 4 + 3 * 8 + 4 * 4 + 4 * 4 + 2 * 8


def test_library_contains_sm100a_code_and_tma(built):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib_path()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass            # cp.async.bulk (TMA) in the pileup kernel


def test_product_fails_loudly_without_gpu(built, tmp_path):
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    d = str(tmp_path / "d")
    H.synth(d, "c1", 0.02, 2)
    rc, err = H.run_product_snpcall(d, str(tmp_path / "out"))
    assert rc != 0 and "no usable CUDA device" in err
    r = H.run_qacompute(bin_path("qaCompute"), open(os.path.join(d, "all_samples")).readline().strip(), str(tmp_path / "x.cov"))
    assert r.returncode != 0 and "no usable CUDA device" in r.stderr
    from metasnv_b200 import abi
    with pytest.raises(abi.MsnvError):
        abi.Context(0)


def test_snpcall_cli_surface(built, tmp_path):
    """Exit codes and messages of call_vC.cpp:346-416."""
    sc = bin_path("snpCall")
    r = subprocess.run([sc, "-h"], capture_output=True)
    assert r.returncode == 255
    r = subprocess.run([sc, "-f", "/nonexistent/ref.fa"], capture_output=True, text=True)
    assert r.returncode == 255 and "Cannot open /nonexistent/ref.fa" in r.stderr
    r = subprocess.run([sc, "stray"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "Non-option argument stray\n"
    r = subprocess.run([sc, "-z"], capture_output=True, text=True)
    assert r.returncode == -6 and "Unknown option `-z'." in r.stderr          # abort()
    r = subprocess.run([sc, "-f"], capture_output=True, text=True)
    assert r.returncode == -6 and "requires a reference file" in r.stderr
    ind = str(tmp_path / "i.txt")
    r = subprocess.run([sc, "-i", ind], input=b"", capture_output=True)
    assert r.returncode == 0 and r.stdout == b"" and os.path.getsize(ind) == 0 and b"Identified 0 samples" in r.stderr


def test_qacompute_cli_surface(built, tmp_path):
    qa = bin_path("qaCompute")
    assert subprocess.run([qa], capture_output=True).returncode == 1
    assert subprocess.run([qa, "-c", "10", "/nonexistent.bam", str(tmp_path / "o")], capture_output=True).returncode == 1
    assert subprocess.run([qa, "-m", "a", "b"], capture_output=True).returncode == 255


def test_samtools_standin(built, tmp_path):
    st = bin_path("samtools")
    r = subprocess.run([st, "mpileup", "-f", "ref.fa", "-l", "split", "-B", "-b", "list"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "#MSNV1\tref.fa\tsplit\tlist\n"
    r = subprocess.run([st, "mpileup", "-f", "ref.fa", "-B", "-b", "list"], capture_output=True, text=True)
    assert r.stdout == "#MSNV1\tref.fa\t-\tlist\n"
    assert subprocess.run([st, "sort", "x"], capture_output=True).returncode == 1


def test_bam_roundtrip_product_writer_oracle_reader(built, tmp_path):
    """BAMs written by the product's BGZF/BAM writer are read identically by the oracle's independent reader
    (qaCompute restatement totals) and the header survives both readers."""
    d = str(tmp_path / "d")
    st = H.synth(d, "c1", 0.02, 3)
    total = 0
    for i, b in enumerate(l.strip() for l in open(os.path.join(d, "all_samples"))):
        out = str(tmp_path / ("s%d.cov" % i))
        assert H.run_qacompute(H.oracle_bin("qacompute_oracle"), b, out).returncode == 0
        m = re.search(r"Total number of reads: (\d+)", open(out).read())
        total += int(m.group(1))
    assert total == st["reads"] + st["junk"] + st["unmapped"]
    assert H.bed_header(d, str(tmp_path / "bed")) and open(str(tmp_path / "bed")).read().count("\n") == 3


def test_fast_inflate_agrees_with_zlib(tmp_path):
    """csrc/host/fast_inflate.cc (the BGZF reader's own DEFLATE decoder) against zlib on ~2900 streams: stored, fixed
    and dynamic blocks, multi-block streams, every size up to 64 KiB, truncated and bit-flipped input, wrong expected
    sizes; guard bytes around the output buffer. The checker is tests/fast_inflate_check.cc."""
    exe = str(tmp_path / "fast_inflate_check")
    host = os.path.join(ROOT, "metasnv_b200", "csrc", "host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", host, os.path.join(ROOT, "tests", "fast_inflate_check.cc"),
                    os.path.join(host, "fast_inflate.cc"), "-lz", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches: 0" in r.stdout, r.stdout + r.stderr[-2000:]


def test_folded_crc32_agrees_with_zlib(tmp_path):
    """csrc/host/crc32_fold.cc (carry-less-multiply CRC-32 of a BGZF member, zlib for the tail and on other CPUs)."""
    exe = str(tmp_path / "crc32_fold_check")
    host = os.path.join(ROOT, "metasnv_b200", "csrc", "host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", host, os.path.join(ROOT, "tests", "crc32_fold_check.cc"),
                    os.path.join(host, "crc32_fold.cc"), "-lz", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches: 0" in r.stdout, r.stdout + r.stderr[-2000:]


def test_overlap_rule_four_lane_form_matches_scalar_rule(tmp_path):
    """csrc/gpu/overlap_rule.h (host + device): the byte-lane form the pileup kernel uses equals the scalar
    restatement of htslib's tweak_overlap_quality (SURVEY.md Annex A.2) for every (quality, quality, base, base)
    in every lane, (q * 205) >> 8 equals (int)(0.8 * q) for every byte, and the quad masks are right.
    The checker is tests/overlap_rule_check.cc (exhaustive, ~1 s)."""
    exe = str(tmp_path / "overlap_rule_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "metasnv_b200", "csrc", "gpu"),
                    os.path.join(ROOT, "tests", "overlap_rule_check.cc"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches: 0" in r.stdout, r.stdout + r.stderr


def test_bench_contract_on_cpu(built):
    """bench.py without a GPU: the product arm refuses loudly (no CPU fallback to time), the reference arm prints
    the contract's JSON line from rank 0 only."""
    import json
    import sys
    bench = os.path.join(ROOT, "bench.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, bench, "--steps", "1"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert len(r.stdout.strip().splitlines()) == 1, "exactly one line on stdout"
    line = json.loads(r.stdout.strip())
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "aligned_bases_per_s" and line["value"] > 1e6
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
    # anything a library prints to file descriptor 1 (NCCL's version banner does) must not reach stdout
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.system('echo banner from a library'); "
            "print('python print'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == '{"a": 1}\n' and "banner from a library" in r.stderr and "python print" in r.stderr
