"""TEST INFRASTRUCTURE — recipe that populates ``oracle/_ref/`` with the UNMODIFIED reference files.

    python oracle/make_ref.py            # needs /root/reference (the build container)

The reference is pure Python (no build step), so "building" it is a byte-for-byte copy of the few
files the hot path and its caller's network live in, from where they lie under ``/root/reference`` into ``oracle/_ref/``:

    lib/car_env.py   lib/buffer.py   lib/model.py   tracks/track.json   tracks/big_track.json

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the directory travels to the GPU box exactly like the built ``*.so`` files do.
There the ``-m gpu`` tests replay the CUDA path against the reference's own code, and
``bench.py --impl reference`` / ``cpu_baseline`` time it on the box's host cores
(``kind: "reference"``).  A MANIFEST.json with the sha256 of every copied file is written beside
them; ``oracle.ref_import.reference_root()`` refuses a copy whose hashes do not match.

``gymnasium`` and ``pygame`` are not installed; ``oracle/ref_import.py`` injects two inert stub
modules for them (no arithmetic in the stubs).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PPO_CAR_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["lib/car_env.py", "lib/buffer.py", "lib/model.py", "tracks/track.json", "tracks/big_track.json"]


def sha256(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def make(verbose: bool = True) -> str | None:
    """Copy the files; returns the destination, or None when the reference tree is not mounted."""
    if not os.path.isfile(os.path.join(SRC, "lib", "car_env.py")):
        return None
    manifest = {"source": SRC, "files": {}}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst):
            os.chmod(dst, 0o644)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        os.chmod(dst, 0o644)
        manifest["files"][rel] = sha256(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} unmodified reference files copied from {SRC}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
