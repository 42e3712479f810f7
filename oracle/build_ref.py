"""TEST INFRASTRUCTURE: make the reference's own implementation of the sampling path available next to the oracle.

The reference is a plain Python tree without setup.py / pyproject (it cannot be `pip install`ed), and the GPU box has
no /root/reference.  This recipe -- run by __graft_entry__.build() in the container where the reference is mounted --
copies the Python sources the path needs (`ldm/**/*.py`, `vocoder/**/*.py`, the shipped `configs/*.yaml`) from where
they lie under the reference root into oracle/_ref/, which is git-ignored (never enters the history: it is a build
output like a wheel unpacked by `pip install --target`) but travels to the GPU box with the snapshot.  Nothing in the
product package imports it; tests/, bench.py's reference legs and oracle/ref_loader.py do.

    python -m oracle.build_ref            # idempotent; prints what it did
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MA3_REFERENCE_ROOT", "/root/reference")
TREES = ("ldm", "vocoder")


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "ldm")):
        if verbose:
            state = "kept" if os.path.isdir(os.path.join(DEST, "ldm")) else "absent"
            print(f"oracle/_ref: reference tree not mounted at {SRC}; existing copy {state}")
        return False
    n = 0
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "tsv_dirs", "data")]
            for f in files:
                if not f.endswith((".py", ".yaml", ".yml")):
                    continue
                src = os.path.join(root, f)
                dst = os.path.join(DEST, os.path.relpath(src, SRC))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                n += 1
    os.makedirs(os.path.join(DEST, "configs"), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(SRC, "configs"))):
        if f.endswith(".yaml"):
            shutil.copyfile(os.path.join(SRC, "configs", f), os.path.join(DEST, "configs", f))
            n += 1
    with open(os.path.join(DEST, "PROVENANCE"), "w") as fh:
        fh.write(f"copied verbatim from {SRC} by oracle/build_ref.py ({n} files); git-ignored build output\n")
    if verbose:
        print(f"oracle/_ref: {n} files copied from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() or os.path.isdir(os.path.join(DEST, "ldm")) else 1)
