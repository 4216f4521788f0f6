"""Install recipe for the UNMODIFIED reference under ``baseline/_ref`` (git-ignored, travels to the GPU box with the snapshot).

TEST / BENCH INFRASTRUCTURE ONLY - nothing under the product package imports this.

The reference has no ``setup.py`` / ``pyproject.toml`` (``pip install /root/reference`` is not applicable; recorded in
DESIGN.md section 0), so the "install" is a byte-for-byte file copy of the modules on the hot path from the read-only
checkout at ``/root/reference`` into ``baseline/_ref`` - the same thing ``pip install --target`` would do for a packaged
project.  Nothing is copied into git history: ``baseline/_ref/`` is listed in ``.gitignore``.

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference exists

Layout written:
    baseline/_ref/seg/nets/*.py, baseline/_ref/seg/utils/*.py      <- Segmentation/deeplabv3+/{nets,utils}
    baseline/_ref/mm/four/*.py                                   <- MultiModal Prediction/Four_Modal/{my_mae_model,mae_utils,util}.py
    baseline/_ref/MANIFEST.json                                  sha256 of every installed file

Consumers: ``bench.py --impl reference`` (the reference's own ``fit_one_epoch`` on the host cores), ``bench.py``'s
``torch_gpu_baseline`` (the same call on the B200 under stock PyTorch/cuDNN, fp16 autocast as train.py:82), and the
golden-vector generators.  Importing the reference needs ``matplotlib`` only for plots that are never drawn here:
``oracle/ref_runner.py`` injects an empty stub for it (the reference files themselves stay untouched)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(os.path.dirname(HERE), "baseline", "_ref")

SETS = [
    ("Segmentation/deeplabv3+/nets", "seg/nets", None),
    ("Segmentation/deeplabv3+/utils", "seg/utils", None),
    ("MultiModal Prediction/Four_Modal", "mm/four", ("my_mae_model.py", "mae_utils.py", "util.py")),
]


def install(ref: str = REF, dst: str = DST) -> bool:
    """Copy the hot-path modules of the reference; returns False when the checkout is absent (GPU box)."""
    if not os.path.isdir(ref):
        return os.path.exists(os.path.join(dst, "MANIFEST.json"))
    manifest = {}
    for src_rel, dst_rel, only in SETS:
        src_dir, dst_dir = os.path.join(ref, src_rel), os.path.join(dst, dst_rel)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(".py") or (only is not None and name not in only):
                continue
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            with open(os.path.join(dst_dir, name), "rb") as f:
                manifest[os.path.join(dst_rel, name)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "files": manifest}, f, indent=1, sort_keys=True)
    return True


def available(dst: str = DST) -> bool:
    return os.path.exists(os.path.join(dst, "MANIFEST.json"))


def verify(dst: str = DST) -> bool:
    """Every installed file still has the hash recorded at install time (nobody edited the reference copy)."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        man = json.load(f)
    for rel, digest in man["files"].items():
        with open(os.path.join(dst, rel), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                return False
    return True


if __name__ == "__main__":
    ok = install()
    print("baseline/_ref %s" % ("installed" if ok else "NOT available (no /root/reference here)"))
    sys.exit(0)
