"""Copies the UNMODIFIED reference files of the hot path into baseline/_ref/ (git-ignored, shipped to
the GPU box by gpurun) so that `bench.py --impl reference` can time the reference's own modules
there.  The reference is pure Python + torch: there is nothing to pip-install (no setup.py /
pyproject in /root/reference), so the "install" is a verbatim copy of the files the path imports
 plus the trainer script and its dataset adapters.  Run by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "learning/__init__.py",
    "learning/pointnet.py",
    "learning/graphnet.py",
    "learning/modules.py",
    "learning/metrics.py",
    "learning/ecc/__init__.py",
    "learning/ecc/GraphConvInfo.py",
    "learning/ecc/GraphConvModule.py",
    "learning/ecc/GraphPoolInfo.py",
    "learning/ecc/GraphPoolModule.py",
    "learning/ecc/cuda_kernels.py",
    "learning/ecc/utils.py",
    # the trainer and its data side, for the "learning/main.py unchanged" run (compat/run_main.py)
    "learning/main.py",
    "learning/spg.py",
    "learning/s3dis_dataset.py",
    "learning/custom_dataset.py",
    "learning/sema3d_dataset.py",
    "learning/vkitti_dataset.py",
]


def install(ref_root=os.environ.get("SPG_REFERENCE", "/root/reference")):
    """-> number of files copied (0 if the reference tree is absent: the GPU box)."""
    if not os.path.isdir(ref_root):
        return 0
    n = 0
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    return n


def available():
    return os.path.exists(os.path.join(DST, "learning", "pointnet.py"))


if __name__ == "__main__":
    print("copied", install(), "files ->", DST)
