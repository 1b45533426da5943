"""The `devit` entrypoint reads Google-Brain Flax .npz checkpoints (models/de_vit.py:223-224,
:372-449, :506-513).  devit_b200/npz_loader.py must fill the parameters exactly like the
reference's own loader does, from the same archive -- checked against the UNMODIFIED reference
(through oracle/ref_shim.py) when it is available, and by a round trip otherwise."""
import numpy as np
import pytest
import torch

from devit_b200 import models, npz_loader, synth  # noqa: F401
from devit_b200.registry import create_model


def _archive(tmp_path, prefix='', img=224, squeeze_vec=False):
    sd = synth.vit_state_dict(77, dim=384, depth=12, num_classes=10, img=img, distilled=False)
    path = str(tmp_path / 'vit.npz')
    npz_loader.state_dict_to_npz(sd, path, num_heads=6, prefix=prefix)
    if squeeze_vec:  # some archives store vectors as [1, 1, 1, C]
        z = dict(np.load(path))
        k = prefix + 'Transformer/encoder_norm/scale'
        z[k] = z[k].reshape(1, 1, 1, -1)
        np.savez(path, **z)
    return sd, path


@pytest.mark.parametrize('prefix,squeeze', [('', False), ('opt/target/', True)])
def test_npz_round_trip(tmp_path, prefix, squeeze):
    sd, path = _archive(tmp_path, prefix, squeeze_vec=squeeze)
    m = create_model('devit', pretrained=True, pretrained_path=path, num_classes=10)
    got = m.state_dict()
    assert list(got.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(got[k], sd[k]), k


def test_npz_other_resolution_and_head_size(tmp_path):
    """A 160x160 archive into a 224x224 model: the position table is re-gridded; a head of
    another width is left at its initial value (the reference's rule)."""
    sd, path = _archive(tmp_path, img=160)
    m = create_model('devit', num_classes=7)
    head0 = m.head.weight.clone()
    m.load_pretrained(path)
    assert m.pos_embed.shape == (1, 197, 384)
    assert torch.equal(m.pos_embed[:, 0], sd['pos_embed'][:, 0])       # cls slot copied as is
    assert torch.equal(m.head.weight, head0)                           # 10-class head not loaded
    assert torch.equal(m.blocks[3].attn.qkv.weight, sd['blocks.3.attn.qkv.weight'])


def test_npz_matches_the_reference_loader(tmp_path):
    from oracle import ref_shim
    if ref_shim.reference_root() is None:
        pytest.skip('reference not available on this machine')
    _, path = _archive(tmp_path)
    _, path160 = (None, None)
    de_vit, _, _, ref_create = ref_shim.load_reference()
    ref = ref_create('devit', pretrained=True, pretrained_path=path, num_classes=10)
    mine = create_model('devit', pretrained=True, pretrained_path=path, num_classes=10)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k
    # and with a position table that has to be resized
    sd160 = synth.vit_state_dict(78, dim=384, depth=12, num_classes=10, img=160, distilled=False)
    p160 = str(tmp_path / 'vit160.npz')
    npz_loader.state_dict_to_npz(sd160, p160, num_heads=6)
    ref2 = ref_create('devit', num_classes=10)
    ref2.load_pretrained(p160)
    mine2 = create_model('devit', num_classes=10)
    mine2.load_pretrained(p160)
    assert torch.equal(ref2.pos_embed, mine2.pos_embed)
