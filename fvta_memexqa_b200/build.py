"""Builds libfvta_b200.so (the C-ABI library, include/fvta_b200.h) with nvcc for sm_100a.

In-tree on purpose: the built .so travels with the repo snapshot to the GPU box.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# FVTA_LIB_PATH: development only -- load an instrumented build (tools/tc_trace.py, tile-size sweeps)
LIB_PATH = os.environ.get("FVTA_LIB_PATH") or os.path.join(LIB_DIR, "libfvta_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_library(force=False, verbose=False, out=None, extra=()):
    """Compile every csrc/*.cu into one shared library.  Returns its path.  out / extra: an
    instrumented variant (other path, extra nvcc flags) next to the production library."""
    if out is None and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("NVCC_EXTRA", "").split() + list(extra)
    objs = []
    procs = []
    obj_dir = LIB_DIR if out is None else out + ".objs"
    os.makedirs(obj_dir, exist_ok=True)
    lib_path = os.path.join(LIB_DIR, "libfvta_b200.so") if out is None else out
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path] + objs + ["-lcudart"]
    subprocess.run(cmd, check=True)
    return lib_path


if __name__ == "__main__":
    import sys
    out, extra = None, []
    if "--out" in sys.argv:
        out = sys.argv[sys.argv.index("--out") + 1]
        extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out, extra=extra))
