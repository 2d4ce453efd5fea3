"""Builds ``libxpgnn_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(verbose=False, jobs=8, force=False):
    """``force``: recompile every translation unit (``make -B``) -- what ``__graft_entry__.build()`` does, so that a
    successful build always means "these sources compile for sm_100a", not "object files were lying around"."""
    cmd = ["make", "-C", _CSRC, "-j%d" % jobs] + (["-B"] if force else [])
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout[-4000:] + proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("nvcc build of libxpgnn_b200.so failed")
    return os.path.join(os.path.dirname(_CSRC), "libxpgnn_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))
