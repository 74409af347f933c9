import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
BIN_DIR = os.path.join(PKG_DIR, "bin")
LIB_DIR = os.path.join(PKG_DIR, "lib")
ORACLE_DIR = os.path.join(REPO_ROOT, "oracle")


def lib_path(name="libmsnv_gpu.so"):
    return os.path.join(LIB_DIR, name)


def bin_path(name):
    return os.path.join(BIN_DIR, name)
