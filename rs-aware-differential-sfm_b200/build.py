"""In-tree build of librsdsfm.so (sm_100a only) with nvcc.  No JIT cache: the .so lives next to
this file so it travels with the repository snapshot to the GPU box."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build" + os.environ.get("RSDSFM_BUILD_TAG", ""))
LIB = os.path.join(HERE, os.environ.get("RSDSFM_LIB_NAME", "librsdsfm.so"))   # experiments: RSDSFM_LIB_NAME + RSDSFM_EXTRA_NVCC build a variant next to the product

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xcompiler", "-Wall"]
# Translation units and extra flags.  --fmad=false where results must be bit-identical to the
# reference's scalar IEEE arithmetic (RANSAC inlier sets, integer splat targets, alpha factors).
# api.cu / ransac.cu also hold HOST arithmetic (solve9.h: the 9-point fit): the host compiler must not
# contract a*b+c into an FMA either, or the hypotheses -- and with them the "bit-exact inlier sets" --
# would depend on the host ISA (the oracle is built with -ffp-contract=off too).
UNITS = [
    ("api.cu", ["-Xcompiler", "-ffp-contract=off"]),
    ("pipeline.cu", []),
    ("refine.cu", []),
    ("ransac.cu", ["--fmad=false", "-Xcompiler", "-ffp-contract=off"]),
    ("rectify.cu", ["--fmad=false"]),
    ("preproc.cu", ["--fmad=false"]),
]
HEADERS = ["common.cuh", "stages.h", "lm_controller.h", "lm_layout.h", "lm_kernel.cuh", "solve9.h", os.path.join("..", "..", "include", "rsdsfm.h")]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + extra + os.environ.get("RSDSFM_EXTRA_NVCC", "").split() + (os.environ.get("RSDSFM_EXTRA_NVCC_REFINE", "").split() if src in ("refine.cu", "preproc.cu") else []) + ["-Xptxas", "-v"] * int(verbose) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                print(" ".join(cmd))
                print(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
