"""Stage the UNMODIFIED reference (`/root/reference/src`, pure Python) as `oracle/_ref/src`.

TEST / BENCH INFRASTRUCTURE, not product code.  `/root/reference` exists only in the authoring container; the GPU
box gets a snapshot of this repo.  `oracle/_ref/` is git-ignored (reference sources never enter the history) but NOT
gpurun-ignored, so the staged copy travels with the snapshot, exactly like the built `.so`.  With it

  * `bench.py --impl reference` times the reference's OWN modules (`build_encoder`, `build_decoder`,
    `HungarianMatcherWoL1`, `SetCriterion`; `cpu_baseline.kind = "reference"`) on the box's host cores,
  * `bench.py` reports `reference_gpu_eager`: the same modules run by stock eager torch on the same B200
    (fp32 and `autocast(bf16)`) -- the practical bar of SURVEY 8(d),
  * `tests/test_ref_integration.py` runs the reference's `ObjDetSplitTransformer` with this repo's builders swapped in.

    python oracle/stage_ref.py            # copy (idempotent); __graft_entry__.build() calls it when /root/reference exists

Only `*.py` files of `src/` are copied, byte for byte; MANIFEST.json records their sha256 so a stale or edited copy is
detectable (`verify()`).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_DEFAULT = os.environ.get("DESTR_REF", "/root/reference")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _py_files(root: str):
    for d, _, files in os.walk(os.path.join(root, "src")):
        for fn in sorted(files):
            if fn.endswith(".py"):
                yield os.path.relpath(os.path.join(d, fn), root)


def stage(src: str = SRC_DEFAULT, dst: str = DST) -> dict:
    if not os.path.isdir(os.path.join(src, "src")):
        raise FileNotFoundError(f"{src}/src not found: the reference is only present in the authoring container")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    man = {"source": src, "files": {}}
    for rel in _py_files(src):
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        man["files"][rel] = _sha(out)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    return man


def available(dst: str = DST) -> bool:
    return os.path.exists(os.path.join(dst, "MANIFEST.json"))


def verify(dst: str = DST) -> bool:
    """True when every staged file still has the sha256 recorded at staging time."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        man = json.load(f)
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h
               for rel, h in man["files"].items())


if __name__ == "__main__":
    m = stage(*(sys.argv[1:2] or [SRC_DEFAULT]))
    print(f"staged {len(m['files'])} files from {m['source']} into {DST}")
