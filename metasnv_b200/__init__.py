"""metasnv_b200 -- B200-native hot path of metaSNV Part I (`samtools mpileup | snpCall`, `qaCompute`).

The product is native: ``lib/libmsnv_gpu.so`` (CUDA kernels behind the C ABI of ``include/msnv.h``)
and the drop-in programs in ``bin/`` (``snpCall``, ``qaCompute``, ``samtools`` stand-in, ``msnv_synth``).
This Python package only locates and drives them (tests, benchmark, build); it contains no
CPU implementation of the path.
"""
from .paths import BIN_DIR, LIB_DIR, REPO_ROOT, ORACLE_DIR, lib_path, bin_path  # noqa: F401
