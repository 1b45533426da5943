"""Recipe that places the UNMODIFIED reference hot-path modules under oracle/_ref/ so that the
reference itself (not only its restatement) can run where /root/reference does not exist -- the
GPU box.  oracle/_ref/ is git-ignored (never part of the history) but travels with gpurun
snapshots, like the built .so files.  Run by __graft_entry__.build() whenever /root/reference is
present; a no-op otherwise (the GPU box uses the files placed here).

Only the files the hot path imports are taken: the models/ package and the
two core/ modules holding the gate index selection and the analytic MACs.  They are byte copies;
the shims that make them importable live in oracle/ref_shim.py and edit nothing.

Test infrastructure only: imported by tests/, __graft_entry__ and bench.py's CPU legs.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"
FILES = [
    "models/__init__.py", "models/de_vit.py", "models/deit_vit.py", "models/ensemble_models.py",
    "models/cct.py", "models/utils/__init__.py", "models/utils/config.py",
    "models/utils/embedder.py", "models/utils/helpers.py", "models/utils/stochastic_depth.py",
    "models/utils/tokenizer.py", "models/utils/transformers.py",
    "core/imp_rank.py", "core/compute_metric.py",
]


def make_ref(force: bool = False) -> Path | None:
    """Copies FILES from /root/reference into oracle/_ref/.  Returns the directory, or None when
    neither the source tree nor an earlier copy exists."""
    if not SRC.is_dir():
        return DST if (DST / "MANIFEST.json").exists() else None
    manifest = {}
    for rel in FILES:
        src, dst = SRC / rel, DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        data = src.read_bytes()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if force or not dst.exists() or dst.read_bytes() != data:
            shutil.copyfile(src, dst)
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    return DST


def verify() -> bool:
    """True when every file under oracle/_ref/ still matches the hash recorded at copy time."""
    mf = DST / "MANIFEST.json"
    if not mf.exists():
        return False
    manifest = json.loads(mf.read_text())
    return all((DST / rel).exists() and
               hashlib.sha256((DST / rel).read_bytes()).hexdigest() == h
               for rel, h in manifest.items()) and set(manifest) == set(FILES)


if __name__ == "__main__":
    print(make_ref(force=True), "verified" if verify() else "NOT verified")
