"""In-tree build of libmaplab_lc_b200.so (nvcc, sm_100a only)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmaplab_lc_b200.so")


def build(force=False, jobs=8):
    """Compile every CUDA source for sm_100a (`-gencode arch=compute_100a,code=sm_100a -lineinfo`,
    see csrc/Makefile) and link the C-ABI library. nvcc cross-compiles without a GPU."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-C", CSRC, f"-j{jobs}"])
    assert os.path.exists(LIB_PATH)
    return LIB_PATH
