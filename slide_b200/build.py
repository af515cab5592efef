"""In-tree build of libslide_b200.so (hand-written sm_100a CUDA + the C ABI of include/slide_b200.h).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it travels with the
repo snapshot.  Only the CUDA runtime is linked (statically) -- no torch, no cuBLAS.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libslide_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(
        glob.glob(os.path.join(CSRC, "*.h"))) + sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for f in files + [os.path.abspath(__file__)]:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def build_variant(name, defines):
    """A/B helper for kernel tuning: compile csrc/ with extra -D flags into libslide_b200_<name>.so (select it at
    run time with SLIDE_B200_LIB=<path>).  Not used by the product path."""
    nvcc = os.environ.get("NVCC", "nvcc")
    out = os.path.join(HERE, "libslide_b200_%s.so" % name)
    objs = []
    bdir = os.path.join(HERE, "build", name)
    os.makedirs(bdir, exist_ok=True)
    for src in _sources():
        obj = os.path.join(bdir, os.path.basename(src) + ".o")
        subprocess.check_call([nvcc] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + list(defines) +
                              ["-c", src, "-o", obj], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-o", out] + objs + ["-cudart", "static"])
    return out


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libslide_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see slide_b200/build/ptxas.log")
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static"])
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
