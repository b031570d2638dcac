"""Install the UNMODIFIED reference (LDenninger/CamC2V) into oracle/_ref/ so that it travels to the GPU box.

    python oracle/refgen/install_ref.py            # copies <reference>/CamContextI2V/**/*.py and <reference>/configs/**/*.yaml

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python (no setup.py / pyproject, nothing to compile), so
"installing" it means copying its source files, byte for byte, to a place the GPU box can see: `oracle/_ref/` is listed in
.gitignore (the copies never enter this repository's history) but not in .gpurunignore (they are shipped with the snapshot, like
the built .so files).  A MANIFEST.json with the sha256 of every copied file is written next to them; `verify()` re-checks it, so
a modified copy is detected.  Consumers: `bench.py --impl reference` / `bench.py`'s `cpu_baseline` leg (times
DDIMSampler.p_sample_ddim of these files on the host cores) and tests/test_dropin_gpu.py (drives the reference's own sampler and
LatentDiffusion with this repo's UNet swapped in).  The product package `camc2v_b200/` never imports anything from here.

/root/reference itself does not exist on the GPU box; nothing there reads it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
SRC = os.environ.get("CAMC2V_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
TREES = (("CamContextI2V", (".py",)), ("configs", (".yaml", ".yml", ".json")))


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, "MANIFEST.json"))


def verify() -> bool:
    """True when every file listed in the manifest is present in oracle/_ref with the recorded sha256."""
    if not installed():
        return False
    with open(os.path.join(DST, "MANIFEST.json")) as f:
        man = json.load(f)
    return all(os.path.isfile(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in man["files"].items())


def install(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "CamContextI2V")):
        raise RuntimeError(f"reference sources not found under {SRC}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    files = {}
    for tree, exts in TREES:
        for dirpath, _dirs, names in os.walk(os.path.join(SRC, tree)):
            for n in sorted(names):
                if not n.endswith(exts):
                    continue
                s = os.path.join(dirpath, n)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                files[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "LDenninger/CamC2V (unmodified copies; see oracle/refgen/install_ref.py)", "files": files}, f, indent=0, sort_keys=True)
    if verbose:
        print(f"installed {len(files)} unmodified reference files into {DST}")
    return DST


if __name__ == "__main__":
    install()
    sys.exit(0 if verify() else 1)
