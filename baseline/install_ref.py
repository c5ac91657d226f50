"""Copy the reference's Python sources into the git-ignored ``baseline/_ref`` (build container only).

    python baseline/install_ref.py

The reference has no setup.py / pyproject.toml, so the base contract's ``pip install ... --target baseline/_ref`` does
not apply; a plain copy of its ``*.py`` files is the install.  ``baseline/_ref`` is listed in .gitignore (the sources
never enter this repository's history) but NOT in .gpurunignore, so it ships to the GPU box with the snapshot, where
``bench.py`` runs the UNMODIFIED reference through PyTorch eager on the same B200 (``gpu_eager_baseline``).
Weights are not copied: ``model_zoo/ffdnet_*.pth`` of this repository are the reference's own files, FastDVDnet uses
the synthetic init (its trained file is absent upstream)."""
import os
import shutil
import sys

SRC = os.environ.get("SCI_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def install():
    if not os.path.isfile(os.path.join(SRC, "utilspy.py")):
        return False
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "model_zoo", "results", "dataset")]
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), SRC)
                out = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), out)
                n += 1
    print("baseline/_ref: %d reference source files copied from %s" % (n, SRC))
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else "reference tree not present at %s" % SRC)
