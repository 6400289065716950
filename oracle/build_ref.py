"""Recipe: compile the UNMODIFIED reference pointnet2 CUDA extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under geoformer_b200/ imports this.

The sources are compiled where they lie under /root/reference/lib/pointnet2/_ext_src
(nothing is copied into this repo); the only output is oracle/_ref/pointnet2/_ext*.so
(git-ignored, but it travels to the GPU box with the gpurun snapshot).  We do not run the
reference's own setup.py; the flags below restate lib/pointnet2/setup.py:19-33 (-O2 for both
compilers, include dir _ext_src/include) with the arch pinned to sm_100.

The resulting module is CUDA-only ("CPU not supported", sampling.cpp:35-37), so it can only
*execute* on the GPU box.  It is used by
  * tests/test_gpu_vs_reference_ext.py  (-m gpu): oracle == reference kernels == our kernels
  * bench.py --impl reference-gpu (extra, reported timing of the reference's own CUDA ops)

Run:  python oracle/build_ref.py           (about 3-4 minutes on 8 cores)
"""
import glob
import os
import sys

REF = "/root/reference/lib/pointnet2/_ext_src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pointnet2")


def build(verbose=True):
    if not os.path.isdir(REF):
        raise SystemExit("reference sources not present at %s (GPU box?): nothing to build" % REF)
    os.makedirs(OUT, exist_ok=True)
    init = os.path.join(OUT, "__init__.py")
    if not os.path.exists(init):
        open(init, "w").write("# package shell so that `import pointnet2._ext` resolves to the reference build\n")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    srcs = sorted(glob.glob(REF + "/src/*.cpp") + glob.glob(REF + "/src/*.cu"))
    mod = load(
        name="_ext",
        sources=srcs,
        extra_include_paths=[REF + "/include"],
        extra_cflags=["-O2"],
        extra_cuda_cflags=["-O2"],
        build_directory=OUT,
        verbose=verbose,
    )
    return mod


def load_ref_ext():
    """Import the prebuilt reference extension (no compilation).  Returns None if absent."""
    so = glob.glob(os.path.join(OUT, "_ext*.so"))
    if not so:
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location("_ext", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build()
    print("built:", m.__file__, file=sys.stderr)
