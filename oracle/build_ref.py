"""Compile the reference's OWN CUDA extension (pointnet2_ops._ext) from the sources where they lie under
/root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing is copied into the repo.

The reference's setup.py hard-codes Kepler..Turing arch flags (pointnet2_ops_lib/setup.py:19) that CUDA 12.9
rejects, so the sources are compiled directly with torch.utils.cpp_extension for sm_100a.  The result can
only run on a GPU: tests/test_gpu_index_ops.py::test_against_reference_cuda_extension loads it there to pin oracle/slide_oracle.c and the
slide_b200 kernels against the reference's real kernels (tie rules included).
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EXT = "/root/reference/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT = os.path.join(HERE, "_ref")
NAME = "slide_ref_ext"


def so_path():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build(verbose=False):
    if so_path():
        return so_path()
    if not os.path.isdir(REF_EXT):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    srcs = sorted(glob.glob(os.path.join(REF_EXT, "src", "*.cpp")) + glob.glob(os.path.join(REF_EXT, "src", "*.cu")))
    load(NAME, sources=srcs, extra_include_paths=[os.path.join(REF_EXT, "include")], extra_cflags=["-O3"],
         extra_cuda_cflags=["-O3"], with_cuda=True, build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path()


def load_module():
    """Import the prebuilt reference extension (GPU box or here); returns None if it was never built."""
    p = so_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
