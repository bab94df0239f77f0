"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference travel to the GPU box.

    python -m oracle.build_ref            # /root/reference/PocketNeRF/*.py  ->  oracle/_ref/PocketNeRF/

The reference is ~5 kLoC of pure Python with no build step, so "building" it is a file copy of its own sources,
from where they lie under /root/reference, into the git-ignored directory oracle/_ref/ (listed in .gitignore, NOT in
.gpurunignore: like the built .so files it ships with a gpurun snapshot but never enters history).  Nothing is edited;
oracle/ref_shim.py imports the copy exactly as it imports /root/reference.  Only `tests/`, `__graft_entry__` and
`bench.py`'s reference legs use it; the product package never does.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("POCKETNERF_REFERENCE_SRC", "/root/reference/PocketNeRF")
DST = os.path.join(HERE, "_ref", "PocketNeRF")


def build(verbose=True):
    """Copy the reference's Python sources (top-level modules + configs).  Returns the number of files copied, or -1
    when the reference tree is not present (the GPU box: the prebuilt copy is used as is)."""
    if not os.path.isfile(os.path.join(SRC, "run_nerf.py")):
        return -1
    os.makedirs(DST, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(SRC)):
        s = os.path.join(SRC, name)
        if os.path.isfile(s) and (name.endswith(".py") or name in ("LICENSE", "requirements.txt")):
            d = os.path.join(DST, name)
            if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
                n += 1
    cfg = os.path.join(SRC, "configs")
    if os.path.isdir(cfg):
        os.makedirs(os.path.join(DST, "configs"), exist_ok=True)
        for name in sorted(os.listdir(cfg)):
            if name.endswith(".txt"):
                d = os.path.join(DST, "configs", name)
                if not (os.path.isfile(d) and filecmp.cmp(os.path.join(cfg, name), d, shallow=False)):
                    shutil.copyfile(os.path.join(cfg, name), d)
                    n += 1
    if verbose:
        print("oracle/_ref: %d file(s) refreshed from %s" % (n, SRC))
    return n


if __name__ == "__main__":
    sys.exit(0 if build() >= 0 else 1)
