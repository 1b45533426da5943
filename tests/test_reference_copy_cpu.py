"""oracle/_ref (the byte copy of the reference's hot-path modules that travels to the GPU box,
oracle/make_ref.py) must be intact and must reproduce the oracle restatement -- this is what lets
bench.py's CPU arm time the reference ITSELF (cpu_baseline.kind = "reference") on the GPU box."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_recipe_and_copy_match_the_source_tree():
    from oracle import make_ref
    dst = make_ref.make_ref()
    if dst is None:
        pytest.skip("no /root/reference and no oracle/_ref copy on this machine")
    assert make_ref.verify()
    src = Path("/root/reference")
    if src.is_dir():
        for rel in make_ref.FILES:
            assert (dst / rel).read_bytes() == (src / rel).read_bytes(), rel


def test_reference_from_the_copy_equals_the_oracle():
    """Run in a subprocess with DEVIT_REF_ROOT=oracle/_ref so that the modules are imported from
    the copy even where /root/reference exists."""
    ref = ROOT / "oracle" / "_ref"
    if not (ref / "models" / "de_vit.py").exists():
        pytest.skip("oracle/_ref not built (run __graft_entry__.build() next to /root/reference)")
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from oracle import ref_runner, ref_shim, devit_oracle as O\n"
        "from devit_b200 import synth\n"
        "assert ref_shim.REFERENCE_ROOT == %r, ref_shim.REFERENCE_ROOT\n"
        "x = synth.images(2)\n"
        "a = ref_runner.ensemble(2, 10, True)(x)\n"
        "import models.de_vit as dv; assert dv.__file__.startswith(%r), dv.__file__\n"
        "sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(2)]\n"
        "esd = synth.ensmlp_state_dict(2, num_class=10)\n"
        "gates = [synth.shrink_gates(s) for s in range(2)]\n"
        "with torch.no_grad(): b = O.ensemble_logits(sds, esd, x, gates)[0]\n"
        "r = float((a - b).abs().max() / b.abs().max()); print(r)\n"
        "assert r < 2e-5 and torch.equal(a.argmax(-1), b.argmax(-1))\n"
    ) % (str(ROOT), str(ref), str(ref))
    env = dict(os.environ, DEVIT_REF_ROOT=str(ref))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
