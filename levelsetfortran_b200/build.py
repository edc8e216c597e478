"""Builds liblsf_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False, verbose: bool = False) -> str:
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    if force:
        cmd.append("-B")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        sys.stdout.write(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building liblsf_b200.so failed")
    return os.path.join(_HERE, "liblsf_b200.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
