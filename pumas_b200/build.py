"""Build libpumas_b200.so in-tree with nvcc for sm_100a (cross-compiles without
a GPU). Usage: python -m pumas_b200.build [--force]

User `__device__` medium / locals callbacks (include/pumas_b200_geometry.h,
INTEGRATION.md section 6) must be device-linked with the transport kernels:
    python -m pumas_b200.build --user my_media.cu [-o libpumas_b200_user.so]
builds a second library from the same sources plus the user's translation
units, as relocatable device code (-rdc=true)."""
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


# The resident kernel of kernels_callback.cu runs 12 warps per CTA: functions
# compiled separately (relocatable device code) must fit its register budget.
RDC_MAXRREGCOUNT = "168"


def build_user(user_sources, target=None, verbose=False):
    """The library with the user's .cu files device-linked to the callback
    kernel (kernels_callback.cu). Only those translation units are compiled as
    relocatable device code; the others are built as in the main library."""
    import tempfile
    user_sources = [os.path.abspath(u) for u in user_sources]
    if target is None:
        target = os.path.join(
            os.path.dirname(user_sources[0]), "libpumas_b200_user.so")
    deps = dependencies() + user_sources
    for u in user_sources:
        d = os.path.dirname(u)
        deps += [os.path.join(d, f) for f in os.listdir(d)
                 if f.endswith((".h", ".hpp", ".cuh"))]
    if os.path.exists(target) and all(
            os.path.getmtime(d) <= os.path.getmtime(target) for d in deps):
        return target
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PUMAS_B200_NVCC_EXTRA", "").split()
    compile_flags = [f for f in flags() if f not in ("--shared",)]
    k = compile_flags.index("-cudart")
    del compile_flags[k:k + 2]
    includes = ["-I", CSRC, "-I", os.path.join(ROOT, "include")]
    for u in user_sources:
        includes += ["-I", os.path.dirname(u)]
    plain = [f for f in sources() if not f.endswith("kernels_callback.cu")]
    rdc = [os.path.join(CSRC, "kernels_callback.cu")] + user_sources

    def run(cmd, cwd):
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, cwd=cwd)

    def objects(files):
        return [os.path.splitext(os.path.basename(f))[0] + ".o" for f in files]

    with tempfile.TemporaryDirectory(prefix="pumas_b200_user_") as tmp:
        run([nvcc] + compile_flags + extra + includes + ["-c"] + plain, tmp)
        run([nvcc] + compile_flags + extra + includes +
            ["-dc", "-maxrregcount", RDC_MAXRREGCOUNT] + rdc, tmp)
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-dlink",
             "-Xcompiler", "-fPIC", "-o", "device_link.o"] + objects(rdc), tmp)
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared",
             "-cudart", "shared", "-o", target] +
            objects(plain) + objects(rdc) + ["device_link.o"], tmp)
    return target


if __name__ == "__main__":
    if "--user" in sys.argv:
        args = sys.argv[1:]
        out = None
        if "-o" in args:
            out = os.path.abspath(args[args.index("-o") + 1])
            del args[args.index("-o"):args.index("-o") + 2]
        users = [a for a in args if a.endswith(".cu")]
        print(build_user(users, out, verbose="-v" in args))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
