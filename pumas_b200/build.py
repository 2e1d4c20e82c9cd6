"""Build libpumas_b200.so in-tree with nvcc for sm_100a (cross-compiles without
a GPU). Usage: python -m pumas_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
TARGET = os.environ.get(
    "PUMAS_B200_TARGET", os.path.join(HERE, "libpumas_b200.so"))

# -fmad=false: no FMA contraction, so that binary64 arithmetic follows the
# reference's order of operations bit for bit ("Numerics").
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xcompiler", "-ffp-contract=off", "--shared", "-cudart", "shared",
    "--threads", "8", "-diag-suppress", "128",
]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if f.endswith(".cu")]


def dependencies():
    deps = sources()
    for d in (CSRC, os.path.join(ROOT, "include")):
        deps += [os.path.join(d, f) for f in os.listdir(d)
                 if f.endswith((".h", ".hpp", ".cuh"))]
    return deps


def up_to_date():
    if not os.path.exists(TARGET):
        return False
    t = os.path.getmtime(TARGET)
    return all(os.path.getmtime(d) <= t for d in dependencies())


def flags():
    """NVCC_FLAGS with the PUMAS_B200_FMAD=1 override (tuning only)."""
    out = list(NVCC_FLAGS)
    if os.environ.get("PUMAS_B200_FMAD") == "1":
        out[out.index("-fmad=false")] = "-fmad=true"
    return out


def build(force=False, verbose=False):
    if not force and up_to_date():
        return TARGET
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PUMAS_B200_NVCC_EXTRA", "").split()
    cmd = [nvcc] + flags() + extra + [
        "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", TARGET
    ] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
