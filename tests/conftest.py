import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have(path):
    return os.path.exists(path)


@pytest.fixture(scope="session")
def built():
    """Product binaries + oracle must exist (built by __graft_entry__.build(); prebuilt on the GPU box)."""
    from metasnv_b200.paths import bin_path, lib_path
    from metasnv_b200 import harness as H
    need = [lib_path(), bin_path("snpCall"), bin_path("qaCompute"), bin_path("samtools"), bin_path("msnv_synth"),
            H.oracle_bin("mpileup_oracle"), H.oracle_bin("snpcall_oracle"), H.oracle_bin("qacompute_oracle")]
    if not all(_have(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    missing = [p for p in need if not _have(p)]
    assert not missing, "native build missing: %s" % missing
    return True


def has_reference_build():
    from metasnv_b200 import harness as H
    return _have(H.oracle_bin("snpCall_ref")) and _have(H.oracle_bin("qaCompute_ref"))


@pytest.fixture(scope="session")
def snpcall_checkers(built):
    """The CPU checkers for snpCall text: always the restatement, plus the reference build when present."""
    from metasnv_b200 import harness as H
    c = [("restatement", H.oracle_bin("snpcall_oracle"))]
    if _have(H.oracle_bin("snpCall_ref")):
        c.append(("reference", H.oracle_bin("snpCall_ref")))
    return c


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, **kw)


@pytest.fixture(scope="module")
def datasets(built, tmp_path_factory):
    """Seeded synthetic BAM sets (bin/msnv_synth), generated once per test module and configuration."""
    from metasnv_b200 import harness as H
    cache = {}
    base = str(tmp_path_factory.mktemp("data"))

    def get(preset, scale, samples, **kw):
        key = (preset, scale, samples, tuple(sorted(kw.items())))
        if key not in cache:
            d = os.path.join(base, "%s_%d" % (preset, len(cache)))
            H.synth(d, preset, scale, samples, **kw)
            cache[key] = d
        return cache[key]
    return get
