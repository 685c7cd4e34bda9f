"""Recipe that makes the REAL reference available on the GPU box (which has no /root/reference).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/ref_port.py for the rules).  Run in the build container by
``__graft_entry__.build()``: packs the reference's two hot-path source files, read where they lie under
/root/reference, into ``oracle/_ref/reference_py.tar.gz``.  ``oracle/_ref/`` is git-ignored (no reference
source enters the history) but not gpurun-ignored, so the archive travels with the repo snapshot.
``oracle/ref_loader.py`` unpacks it into a temporary directory and imports it unmodified.

    python oracle/build_ref.py [--reference /root/reference]
"""
import argparse
import io
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "reference_py.tar.gz")
FILES = ("Networks.py", "Losses.py")            # the modules train.py:97's training_step lives in


def build(reference="/root/reference"):
    if not all(os.path.exists(os.path.join(reference, f)) for f in FILES):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for f in FILES:
            tar.add(os.path.join(reference, f), arcname=f)
    data = buf.getvalue()
    if not (os.path.exists(ARCHIVE) and open(ARCHIVE, "rb").read() == data):
        with open(ARCHIVE, "wb") as fh:
            fh.write(data)
    return ARCHIVE


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    out = build(ap.parse_args().reference)
    print(out or "reference sources not found: nothing built")
    sys.exit(0)
