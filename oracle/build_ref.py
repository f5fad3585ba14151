#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference's solve path, staged for the GPU box (test / bench infrastructure only).

  python oracle/build_ref.py            (run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python (nothing to compile): its "build" packs the modules the SQP path imports -- TrajoptMPCReference.py,
TrajoptPlant.py, TrajoptCost.py, TrajoptConstraint.py, overloading.py, expressions.py (imported by TrajoptCost), GBD-PCG-Python/,
GRiD/__init__.py, GRiD/RBDReference/, GRiD/URDFParser/ -- unmodified, from where they lie under /root/reference, into ONE archive
oracle/_ref/reference_solve_path.zip (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so and stays out of
the repository's history; Python imports it directly through zipimport, tests/ref/refshim.py).  Nothing in the product
package imports it; bench.py's `--impl reference` arm and cpu_baseline leg time it on the box's host cores for the
reference-pinned variant of the workload (no box limits: the reference crashes on multi-coordinate limits), next
to the oracle port that runs the full workload.  A manifest with the sha1 of every staged file is written beside it.
"""
import hashlib
import json
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2T_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

FILES = ["TrajoptMPCReference.py", "TrajoptPlant.py", "TrajoptCost.py", "TrajoptConstraint.py", "overloading.py", "expressions.py", "LICENSE",
         "GRiD/__init__.py", "GRiD/LICENSE"]
DIRS = ["GBD-PCG-Python", "GRiD/RBDReference", "GRiD/URDFParser"]
SKIP = ("RBDReference_generalized.py", "test.py")      # unparseable / stale files, not on the solve path


ARCHIVE = os.path.join(DST, "reference_solve_path.zip")


def build(verbose=True):
    if not os.path.isfile(os.path.join(REF, "TrajoptMPCReference.py")):
        if verbose:
            print("oracle/_ref: reference tree not present at %s (the GPU box uses the staged archive)" % REF)
        return staged()
    manifest = {}
    todo = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(REF, d))):
            if name.endswith((".py", "LICENSE")) and name not in SKIP:
                todo.append(os.path.join(d, name))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in todo:
            src = os.path.join(REF, rel)
            if not os.path.isfile(src):
                continue
            with open(src, "rb") as f:
                data = f.read()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)
            manifest[rel] = hashlib.sha1(data).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "archive": os.path.basename(ARCHIVE), "files": manifest}, f, indent=1)
    if verbose:
        print("oracle/_ref: packed %d unmodified reference files into %s" % (len(manifest), os.path.basename(ARCHIVE)))
    return True


def staged():
    return os.path.isfile(ARCHIVE)


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
