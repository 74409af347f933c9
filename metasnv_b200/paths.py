import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
BIN_DIR = os.path.join(PKG_DIR, "bin")
LIB_DIR = os.path.join(PKG_DIR, "lib")
ORACLE_DIR = os.path.join(REPO_ROOT, "oracle")


def lib_path(name="libmsnv_gpu.so"):
    # MSNV_LIB: alternative build of the same ABI (kernel experiments)
    if name == "libmsnv_gpu.so" and os.environ.get("MSNV_LIB"):
        return os.environ["MSNV_LIB"]
    return os.path.join(LIB_DIR, name)


def bin_path(name):
    return os.path.join(BIN_DIR, name)
