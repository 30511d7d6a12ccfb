"""ORACLE tooling (test / baseline infrastructure, never on the product path).

Stages the UNMODIFIED reference modules of the hot path under ``oracle/_ref/`` so that the reference arm of
``bench.py`` (``--impl reference`` / the ``cpu_baseline`` leg) and the oracle cross-checks can run the reference's
own code on the GPU box, where ``/root/reference`` does not exist. ``oracle/_ref/`` is git-ignored (no reference
source enters the history) but not gpurun-ignored, so the staged tree travels with the snapshot like a built ``.so``.

The reference is a PyTorch repository without a build system (no setup.py / pyproject): "building"
it is a byte-for-byte copy of the files SURVEY.md 8(c) lists, keeping their relative paths so that
``oracle/ref_import.py`` (the 3-line ``sys.modules`` shims for the broken package imports) works on either root.
A SHA-256 manifest is written next to the copies; ``verify()`` re-checks it.

    python -m oracle.build_ref            # copy + manifest (no-op when /root/reference is absent)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("FMM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# relative paths, as SURVEY.md 8(c) imports them
FILES = [
    "Fall_2_Spatial_Temporal_SR/Model/__init__.py",
    "Fall_2_Spatial_Temporal_SR/Model/stgcan.py",
    "Fall_2_Spatial_Temporal_SR/Model/graph.py",
    "Fall_2_Spatial_Temporal_SR/Model/bilstm.py",
    "Fall_2_Spatial_Temporal_SR/Model/combination.py",
    "Fall_2_Spatial_Temporal_SR/Model/build_model.py",
    "Multimodal_Fall3/model/musa_model.py",
    "TRAGCN.py",
    "GRU.py",
    "EmbGCN.py",
    "TA.py",
    "GSTCAN_HAR_conv_10kfold.ipynb",      # the only source of the CNN1D / CNN_BiLSTM sensor branch (cell 2)
]


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def build(verbose: bool = False) -> bool:
    """Copy the listed files when the reference tree is present. Returns True when ``oracle/_ref`` is usable afterwards."""
    if not os.path.isdir(SRC):
        return verify()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
        manifest[rel] = _sha(dst)
        if verbose:
            print(f"{manifest[rel][:12]}  {rel}")
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha256": manifest}, fh, indent=1)
    return True


def verify() -> bool:
    """True when every staged file exists and matches the manifest."""
    try:
        manifest = json.load(open(os.path.join(DST, "MANIFEST.json")))["sha256"]
    except (OSError, ValueError, KeyError):
        return False
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in manifest.items()) \
        and set(manifest) == set(FILES)


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref ready" if ok else "reference tree not found and oracle/_ref not staged")
    sys.exit(0 if ok else 1)
