"""Recipe for oracle/_ref/: the reference's OWN files for the hot path, copied (never committed: oracle/_ref/ is
git-ignored, but it travels to the GPU box with the working tree) from /root/reference where that exists -- the build
container.  The reference is pure Python on this path, so "compiling it from its sources" is a copy:

    pygcn/layers.py            the GraphConvolution class itself            -> bench.py --impl reference, cpu_baseline
    pygcn/models.py            GCN / GeneratorGCN / GCN_OVER_MLP / Generator -> tests/test_gpu_ref_models.py (the
    pygcn/utils.py             (models.py imports ReplayBuffer from it)        reference's real model classes on the
    gt-generator/constants.py  (utils.py imports it)                          drop-in layer)

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under pygcn_b200/ reads oracle/_ref/.

    python oracle/make_ref.py          (also run by __graft_entry__.build())
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [("pygcn/layers.py", "layers.py"), ("pygcn/models.py", "models.py"), ("pygcn/utils.py", "utils.py"),
         ("gt-generator/constants.py", "constants.py")]


def make_ref():
    """Returns the list of files now present under oracle/_ref/ (unchanged when /root/reference is absent)."""
    if os.path.isdir(REF):
        os.makedirs(OUT, exist_ok=True)
        for src, dst in FILES:
            shutil.copyfile(os.path.join(REF, src), os.path.join(OUT, dst))
    return [d for _, d in FILES if os.path.exists(os.path.join(OUT, d))]


if __name__ == "__main__":
    got = make_ref()
    print("oracle/_ref:", got if got else "absent (no /root/reference here)")
    sys.exit(0)
