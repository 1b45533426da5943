"""GPU parity of the two steps either side of the forward path, through
the C ABI: uint8 images -> normalised patch matrix (bit-exact against the oracle's torchvision
restatement) and the device-side eval tail (counts exact, loss <= 2e-6 relative), plus the
engine.evaluate* drop-ins against the reference's own engine.evaluate golden."""
from pathlib import Path

import numpy as np
import pytest
import torch

from devit_b200 import _lib as L
from devit_b200 import engine, synth
from devit_b200.models import IMAGENET_DEFAULT_MEAN as MEAN, IMAGENET_DEFAULT_STD as STD
from oracle import devit_oracle as O
from test_model_gpu import make_ensemble, make_sub

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).parent / 'golden' / 'edge_golden.npz')


def _patch_rows(x, num_prefix):
    """[B,C,H,W] fp32 -> token-row patch matrix [B*(prefix+P), C*256] (K order c, py, px)."""
    B, C, H, W = x.shape
    g = H // 16
    p = x.view(B, C, g, 16, g, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * 256)
    return torch.cat([torch.zeros(B, num_prefix, C * 256), p], 1).reshape(-1, C * 256)


@pytest.mark.parametrize('nhwc', [False, True])
@pytest.mark.parametrize('batch,img,prefix', [(2, 32, 0), (3, 224, 2), (1, 48, 1)])
def test_u8_patches_bit_exact(nhwc, batch, img, prefix):
    u8 = synth.images_u8(batch, img=img, nhwc=nhwc)
    ref = _patch_rows(O.to_tensor_normalize(u8, MEAN, STD, 'nhwc' if nhwc else 'nchw'), prefix)
    lay = L.LAYOUT_NHWC if nhwc else L.LAYOUT_NCHW
    split = L.im2col_tokens_u8(u8.cuda(), MEAN, STD, prefix, L.DEVIT_FP32, lay).cpu()
    assert torch.equal(split[0] + split[1], ref)          # hi + lo == the fp32 value, exactly
    assert torch.equal(split[0], L.split_tf32(ref)[0])
    b16 = L.im2col_tokens_u8(u8.cuda(), MEAN, STD, prefix, L.DEVIT_BF16, lay).cpu()
    assert torch.equal(b16, ref.bfloat16())


def test_u8_patches_golden_and_every_byte_value():
    u8 = torch.from_numpy(G['norm_u8_nhwc']).cuda()
    got = L.im2col_tokens_u8(u8, MEAN, STD, 0, L.DEVIT_FP32, L.LAYOUT_NHWC).cpu()
    assert torch.equal(got[0] + got[1], _patch_rows(torch.from_numpy(G['norm_out']), 0))
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1).expand(1, 16, 16, 3)
    got = L.im2col_tokens_u8(ramp.contiguous().cuda(), MEAN, STD, 0, L.DEVIT_FP32,
                             L.LAYOUT_NHWC).cpu()
    assert torch.equal(got[0] + got[1],
                       _patch_rows(torch.from_numpy(G['norm_ramp_out'])[None], 0))


def test_u8_single_channel_and_custom_norm():
    u8 = synth.images_u8(2, img=32, chans=1)
    ref = _patch_rows(O.to_tensor_normalize(u8, (0.5,), (0.25,)), 1)
    got = L.im2col_tokens_u8(u8.cuda(), (0.5,), (0.25,), 1, L.DEVIT_FP32).cpu()
    assert torch.equal(got[0] + got[1], ref)
    with pytest.raises(L.DevitError):
        L.im2col_tokens_u8(u8.cuda(), (0.5,), (0.0,), 1)
    with pytest.raises(L.DevitError):
        L.im2col_tokens_u8(u8.float().cuda(), (0.5,), (0.25,), 1)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('nhwc', [False, True])
def test_ensemble_on_u8_equals_ensemble_on_normalised_fp32(precision, nhwc):
    """Same patch matrix bit for bit => same logits bit for bit; and the oracle on the CPU-
    normalised batch is matched within the north-star tolerance."""
    mv, fuse = make_ensemble(precision, shrunk=True)
    u8 = synth.images_u8(3, nhwc=nhwc)
    xn = O.to_tensor_normalize(u8, MEAN, STD, 'nhwc' if nhwc else 'nchw')
    out_u8 = fuse(mv(u8.cuda()))
    out_f32 = fuse(mv(xn.cuda()))
    assert torch.equal(out_u8, out_f32)
    sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(4)]
    gates = [synth.shrink_gates(s) for s in range(4)]
    with torch.no_grad():
        ref, _, _ = O.ensemble_logits(sds, synth.ensmlp_state_dict(4), xn, gates)
    err = ((out_u8.cpu() - ref).abs().max() / ref.abs().max()).item()
    assert err < {'fp32': 1e-4, 'bf16': 2e-2}[precision], err
    if precision == 'fp32':
        assert torch.equal(out_u8.argmax(-1).cpu(), ref.argmax(-1))


def test_single_model_u8_input_and_custom_norm():
    m = make_sub(0, 'fp32')
    m.set_input_norm((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    u8 = synth.images_u8(2)
    xn = O.to_tensor_normalize(u8, (0.5,) * 3, (0.5,) * 3)
    assert torch.equal(m(u8.cuda()), m(xn.cuda()))
    with pytest.raises(AssertionError):
        m(synth.images_u8(1, img=32).cuda())


@pytest.mark.parametrize('sizes,classes', [((8, 8, 5), 100), ((16, 3), 1000), ((4, 4), 3),
                                           ((256,), 100), ((1,), 7)])
def test_eval_tail_vs_oracle(sizes, classes):
    batches = synth.eval_batches(sizes, classes)
    acc = torch.zeros(5, device='cuda', dtype=torch.float64)
    for lg, tg in batches:
        out = L.eval_tail(lg.cuda(), tg.cuda(), acc).cpu()
        loss, c1, ck = O.eval_tail(lg, tg)
        assert abs(out[0].item() - loss) <= 2e-6 * abs(loss)
        assert (int(out[1]), int(out[2])) == (c1, ck)
    ref = O.eval_epoch(batches)
    got = engine.meters_to_dict(acc.tolist())
    assert abs(got['loss'] - ref['loss']) <= 1e-6 * abs(ref['loss'])
    assert abs(got['acc1'] - ref['acc1']) < 1e-9 and abs(got['acc5'] - ref['acc5']) < 1e-9
    assert acc[1].item() == len(sizes) and acc[4].item() == sum(sizes)


def test_eval_tail_ties_strided_logits_and_bad_targets():
    logits = torch.tensor([[1.0, 1.0, 1.0, 0.0], [0.0, 2.0, 2.0, 2.0]])
    for tg, k in (([0, 1], 2), ([2, 3], 2), ([1, 2], 2), ([3, 0], 5)):
        tg = torch.tensor(tg)
        out = L.eval_tail(logits.cuda(), tg.cuda(), None, topk=k).cpu()
        loss, c1, ck = O.eval_tail(logits, tg, topk=k)
        assert (int(out[1]), int(out[2])) == (c1, ck)
        assert abs(out[0].item() - loss) <= 2e-6 * abs(loss)
    # a column slice of a wider matrix (row stride > classes)
    wide = synth.eval_batches((9,), 128)[0][0]
    tg = torch.arange(9) % 100
    out = L.eval_tail(wide.cuda()[:, :100], tg.cuda()).cpu()
    loss, c1, ck = O.eval_tail(wide[:, :100].contiguous(), tg)
    assert (int(out[1]), int(out[2])) == (c1, ck) and abs(out[0].item() - loss) <= 2e-6 * loss
    # an out-of-range target poisons the loss and is never correct
    out = L.eval_tail(logits.cuda(), torch.tensor([0, 9]).cuda(), topk=4).cpu()
    assert torch.isnan(out[0]) and (int(out[1]), int(out[2])) == (1, 1)
    with pytest.raises(L.DevitError):
        L.eval_tail(logits.cuda().half(), torch.tensor([0, 1]).cuda())


class _Replay(torch.nn.Module):
    def __init__(self, logits):
        super().__init__()
        self.logits = logits

    def forward(self, images):
        return self.logits[int(images.flatten()[0].item())].cuda()


def test_engine_evaluate_matches_reference_engine_golden():
    for name, sizes, classes in (('c100', (8, 8, 5), 100), ('c1000', (16, 3), 1000),
                                 ('c3', (4, 4), 3)):
        batches = synth.eval_batches(sizes, classes)
        loader = [(torch.full((b[0].shape[0], 1), float(i)), b[1]) for i, b in enumerate(batches)]
        res = engine.evaluate(loader, _Replay([b[0] for b in batches]), torch.device('cuda'))
        ref = G[f'eval_{name}']
        assert abs(res['loss'] - ref[0]) <= 1e-6 * abs(ref[0]), name
        assert abs(res['acc1'] - ref[1]) < 1e-9 and abs(res['acc5'] - ref[2]) < 1e-9, name


def test_evaluate_ens_disjoint_u8_loader_vs_oracle():
    mv, fuse = make_ensemble('fp32', shrunk=True)
    sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(4)]
    gates = [synth.shrink_gates(s) for s in range(4)]
    esd = synth.ensmlp_state_dict(4)
    loader, ref_batches = [], []
    for i, b in enumerate((3, 2)):
        u8 = synth.images_u8(b, seed=10 + i)
        with torch.no_grad():
            ref, _, _ = O.ensemble_logits(sds, esd, O.to_tensor_normalize(u8, MEAN, STD), gates)
        # targets: the oracle's arg-max for the first sample, its runner-up for the others
        tg = ref.topk(2, -1).indices[:, 1].clone()
        tg[0] = ref[0].argmax()
        loader.append((u8, tg))
        ref_batches.append((ref, tg))
    res = engine.evaluate_ens_disjoint(loader, mv, fuse, torch.device('cuda'))
    want = O.eval_epoch(ref_batches)
    assert abs(res['loss'] - want['loss']) <= 1e-4 * abs(want['loss'])
    assert res['acc1'] == want['acc1'] == 40.0 and res['acc5'] == want['acc5'] == 100.0
