"""Recipe that stages the UNMODIFIED reference (bl0/moco) under ``oracle/_ref/`` -- TEST/BASELINE INFRASTRUCTURE.

The reference is 16 pure-Python files with no ``setup.py`` / ``pyproject.toml`` (nothing for pip to install), so
"building" it is a byte-for-byte copy of the files its training loop imports:

    /root/reference/train.py   -> oracle/_ref/train.py
    /root/reference/moco/**    -> oracle/_ref/moco/**

``oracle/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so the staged copy travels to the GPU box with the snapshot, where ``/root/reference`` does not
exist.  ``__graft_entry__.build()`` calls ``stage()`` whenever ``/root/reference`` is present; on the GPU box the
already staged files are used.  Consumers: ``oracle/ref_runner.py`` (bench.py's ``--impl reference`` arm and
``cpu_baseline`` leg) only.  A ``MANIFEST.json`` with the sha256 of every staged file is written next to them so
a run can prove the files are the reference's own.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ["train.py"]
TREES = ["moco"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "train.py")) and os.path.isdir(os.path.join(REF_DST, "moco", "NCE"))


def stage(force: bool = False) -> bool:
    """Copy the reference's training files to oracle/_ref/.  Returns True when a staged copy exists afterwards."""
    if not os.path.isdir(REF_SRC):
        return staged()
    manifest = {}
    os.makedirs(REF_DST, exist_ok=True)
    for rel in FILES:
        shutil.copyfile(os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel))
    for tree in TREES:
        dst = os.path.join(REF_DST, tree)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF_SRC, tree), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for root, _, names in os.walk(REF_DST):
        for n in sorted(names):
            if n.endswith(".py"):
                p = os.path.join(root, n)
                manifest[os.path.relpath(p, REF_DST)] = _sha(p)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")
