"""The ctypes mirrors in devit_b200/_lib.py must have exactly the layout a C compiler gives the
structs of include/devit_b200.h: compile a probe with gcc (the header is plain C) that prints
sizeof / offsetof for every struct field and compare with ctypes.  CPU only."""
import ctypes as C
import re
import shutil
import subprocess
from pathlib import Path

import pytest

from devit_b200 import _lib as L

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / 'include' / 'devit_b200.h'
PAIRS = [('devit_gemm_seg', L.GemmSeg), ('devit_gemm_args', L.GemmArgs),
         ('devit_mlp_args', L.MlpArgs), ('devit_layer_desc', L.LayerDesc),
         ('devit_vit_desc', L.VitDesc), ('devit_vit_exports', L.VitExports),
         ('devit_cct_desc', L.CctDesc), ('devit_block_weights', L.BlockWeights)]


def _c_fields(name, text):
    """Field names of `typedef struct <name> { ... } <name>;` in declaration order."""
    body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (name, name), text, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(','):
            m = re.search(r'(\w+)\s*(\[[^\]]*\])?\s*$', part.strip())
            fields.append(m.group(1))
    return fields


@pytest.mark.skipif(shutil.which('gcc') is None, reason='needs gcc')
def test_ctypes_structs_match_the_c_header(tmp_path):
    text = HEADER.read_text()
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname, ct in PAIRS:
        fields = _c_fields(cname, text)
        assert fields == [f[0] for f in ct._fields_], (cname, fields)
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-std=c11', '-o', str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True,
                                                 text=True).stdout.splitlines())
    for cname, ct in PAIRS:
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f'{cname}.{fname}']) == getattr(ct, fname).offset, (cname, fname)


def test_abi_enums_match_the_header():
    text = HEADER.read_text()
    consts = dict(re.findall(r'\b(DEVIT_[A-Z0-9_]+)\s*=\s*(\d+)', text))
    assert int(consts['DEVIT_BF16']) == L.DEVIT_BF16 and int(consts['DEVIT_FP32']) == L.DEVIT_FP32
    assert [int(consts[k]) for k in ('DEVIT_OUT_BF16', 'DEVIT_OUT_F32', 'DEVIT_OUT_F32_SPLIT')] == \
        [L.OUT_BF16, L.OUT_F32, L.OUT_F32_SPLIT]
    assert [int(consts[k]) for k in ('DEVIT_ACT_NONE', 'DEVIT_ACT_GELU_ERF', 'DEVIT_ACT_RELU')] == \
        [L.ACT_NONE, L.ACT_GELU_ERF, L.ACT_RELU]
    assert [int(consts[k]) for k in ('DEVIT_LAYOUT_NCHW', 'DEVIT_LAYOUT_NHWC')] == \
        [L.LAYOUT_NCHW, L.LAYOUT_NHWC]
    tags = {k[len('DEVIT_TAG_'):].lower(): int(v) for k, v in consts.items()
            if k.startswith('DEVIT_TAG_')}
    for i, name in enumerate(L.TAGS):
        key = {'token_prefix': 'prefix', 'gemm_mlp_fused': 'mlp_fused'}.get(name, name)
        assert tags[key] == i, name
    assert int(re.search(r'#define DEVIT_ABI_VERSION (\d+)', text).group(1)) == L.ABI_VERSION
    assert int(re.search(r'#define DEVIT_MAX_DEPTH (\d+)', text).group(1)) == len(L.VitExports().qkv)
