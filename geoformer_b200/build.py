"""Builds geoformer_b200/libgeoformer_b200.so (the C-ABI CUDA library) for sm_100a, in-tree.

    python -m geoformer_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects are cached under geoformer_b200/build/ and rebuilt
when a source or header is newer.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgeoformer_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _newest_header():
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False, variant=None, extra=()):
    """variant: development builds (e.g. -DGF_TRACE) go to libgeoformer_b200_<variant>.so with their own object
    directory; GF_LIB=<path> makes geoformer_b200._capi load one of them instead of the product library."""
    global OBJ, LIB
    if variant:
        OBJ = os.path.join(HERE, "build", variant)
        LIB = os.path.join(HERE, "libgeoformer_b200_%s.so" % variant)
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr = _newest_header()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without libgomp; nvcc only needs a host g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr):
            return obj, ""
        cmd = ([_nvcc()] + ccbin + NVCC_FLAGS + os.environ.get("GF_NVCC_EXTRA", "").split() + list(extra)
               + ["-c", src, "-o", obj])
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
        if verbose:
            sys.stderr.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc()] + ccbin + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    ext = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=var[0] if var else None, extra=ext))
