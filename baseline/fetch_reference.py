"""Mirror the reference's PYTHON sources of the hot path into baseline/_ref/ (git-ignored, travels to the GPU box with
the snapshot like oracle/_ref; nothing is copied into the tracked tree).

  baseline/_ref/pointnet2/...                       models/, diffusion_utils/, util*.py, data_utils/, ... (*.py only)
  baseline/_ref/pointnet2_ops_lib/pointnet2_ops/    the reference's own python package (pointnet2_modules.py, ...)

Uses: (1) tests/test_gpu_reference_unmodified.py builds the REAL PointNet2CloudCondition / PointAutoencoder from these
files over slide_b200's drop-in `pointnet2_ops` / `pytorch3d` on the GPU and matches tests/golden/golden.npz -- the
proof that the reference's model files run unchanged on this library; (2) `bench.py --impl reference` drives the
unmodified modules on the host cores (cpu_baseline.kind = "reference"); (3) tools/ref_gpu_eager.py times the reference's
eager GPU path on the same box.  Run in the build container (needs /root/reference):  python baseline/fetch_reference.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference"
TREES = ("pointnet2", os.path.join("pointnet2_ops_lib", "pointnet2_ops"))


def available():
    return os.path.isdir(os.path.join(OUT, "pointnet2", "models"))


def fetch(force=False):
    """Copy *.py of the two trees (no configs, no CUDA sources, no data).  Returns OUT, or None when the reference tree is
    absent (GPU box: the mirror made in the build container is used as is)."""
    if available() and not force:
        return OUT
    if not os.path.isdir(SRC):
        return None
    n = 0
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d not in ("_ext-src", "__pycache__", ".git")]
            for f in files:
                if not f.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(root, f), SRC)
                dst = os.path.join(OUT, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), dst)
                n += 1
    with open(os.path.join(OUT, "README"), "w") as fh:
        fh.write("Unmodified python sources of SLIDE-3D/SLIDE (mirror made by baseline/fetch_reference.py; %d files).\n" % n)
    return OUT


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
