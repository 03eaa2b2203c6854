#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference, importable on the GPU box.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` exists only in the build container;
``bench.py --impl reference`` and the ``cpu_baseline`` leg run on the GPU box, where it does not.  This
script copies the reference's own Python sources of the hot path

    graphsage/__init__.py  graphsage/aggregators.py  graphsage/encoders.py  graphsage/model.py

byte for byte from where they lie under ``/root/reference`` into ``oracle/_ref/graphsage/`` and writes a
manifest (sha256 per file) beside them.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never
enter the history) but NOT in ``.gpurunignore``, so it travels to the GPU box with the snapshot like our own
built ``.so``.  Nothing under ``oracle/_ref`` is edited: the one incompatibility with Python >= 3.11
(``random.sample`` on a set, aggregators.py:44) is handled by ``oracle/ref_runtime.py`` with the shim
SURVEY.md s8c describes, installed at import time, outside the copied files.

    python oracle/build_ref.py            # (re)build; no-op when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSAGE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["graphsage/__init__.py", "graphsage/aggregators.py", "graphsage/encoders.py", "graphsage/model.py"]


def build(verbose=False):
    """Copy the reference sources; returns the manifest dict, or None when the reference is not here."""
    if not os.path.isdir(os.path.join(REF, "graphsage")):
        if verbose:
            print("oracle/build_ref.py: %s not present (GPU box?) -- keeping the prebuilt oracle/_ref" % REF)
        return None
    manifest = {"source": REF, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref built from %s (%d files)" % (REF, len(FILES)))
    return manifest


def verify():
    """True when oracle/_ref holds exactly the files the manifest lists, with matching digests."""
    path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    with open(path) as f:
        manifest = json.load(f)
    for rel, digest in manifest["files"].items():
        p = os.path.join(OUT, rel)
        if not os.path.exists(p):
            return False
        with open(p, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                return False
    return set(manifest["files"]) == set(FILES)


if __name__ == "__main__":
    m = build(verbose=True)
    sys.exit(0 if (m is not None or verify()) else 1)
