"""TEST INFRASTRUCTURE: stage the UNMODIFIED reference under ``oracle/_ref/`` so that it travels to the GPU box.

The reference is six Python files with no build system; "building" it is copying them, byte for byte, from
``/root/reference`` (read-only, only present in the build container) into the archive ``oracle/_ref/reference_py.zip`` (imported in place through zipimport) -- which is git-ignored
(no reference source enters the history) but NOT gpurun-ignored, so the copy ships with the snapshot like our own
``.so``.  ``oracle/ref_import.py`` imports the modules from there (with ``gym`` / ``matplotlib`` stand-ins in
``sys.modules``: drone.py:2-7, vectorized_drone.py:2-6 import them at module level) when ``/root/reference`` is
absent.  Consumers: ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and the ``not gpu`` tests.

    python -m oracle.make_ref            # also run by __graft_entry__.build() when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DRONE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "reference_py.zip")          # importable as is (zipimport): one artefact, not loose sources
FILES = ["drone.py", "vectorized_drone.py", "train.py", "test.py", "traj_tb.py", "helper.py"]


def build(verbose: bool = False) -> bool:
    """Copy the reference files; returns False (and leaves any existing copy alone) when the source tree is absent."""
    if not os.path.isfile(os.path.join(SRC, "drone.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            s = os.path.join(SRC, f)
            if not os.path.isfile(s):
                continue
            data = open(s, "rb").read()
            z.writestr(zipfile.ZipInfo(f, date_time=(2020, 1, 1, 0, 0, 0)), data)
            manifest[f] = hashlib.sha256(data).hexdigest()
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files in {ARCHIVE}")
    return True


if __name__ == "__main__":
    if not build(verbose=True):
        raise SystemExit(f"{SRC} not present: nothing staged")
