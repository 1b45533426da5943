"""CCT rows of the hot path on the GPU: tokenizer kernels, sequence pooling, the whole sub-model
and the decomposed ensemble against the oracle / the reference goldens (fp32 mode <= 1e-4,
bf16 <= 2e-2, argmax identical in fp32 mode)."""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from devit_b200 import _lib as L
from devit_b200 import cct, synth
from oracle import cct_oracle as CO

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).resolve().parent / 'golden' / 'cct_golden.npz')
TOL = {'fp32': 1e-4, 'bf16': 2e-2}


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize('cin,cout,hw,nchw', [(3, 256, 32, True), (64, 256, 16, False),
                                              (3, 64, 32, True)])
def test_conv_relu_pool_kernels(cin, cout, hw, nchw):
    """im2col3x3 + GEMM (ReLU epilogue) + channels-last max-pool == conv2d -> relu -> max_pool2d."""
    B = 3
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, cin, hw, hw, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    pos = torch.randn((hw // 2) ** 2, cout, generator=g)
    ref = F.max_pool2d(F.relu(F.conv2d(x, w, None, 1, 1)), 3, 2, 1)
    ref = ref.flatten(2, 3).transpose(-2, -1) + pos            # [B, tokens, cout]
    kpad = (9 * cin + 7) // 8 * 8
    wk = torch.zeros(cout, kpad)
    wk[:, :9 * cin] = w.permute(0, 2, 3, 1).reshape(cout, -1)
    xin = x.cuda() if nchw else x.permute(0, 2, 3, 1).contiguous().cuda()
    st = (cin * hw * hw, hw * hw, hw, 1) if nchw else (hw * hw * cin, 1, hw * cin, cin)
    rows = B * hw * hw
    lib = L.load()
    for prec, tol in ((L.DEVIT_FP32, 1e-5), (L.DEVIT_BF16, 1.5e-2)):
        if prec == L.DEVIT_BF16:
            a = torch.empty(rows, kpad, device='cuda', dtype=torch.bfloat16)
            kind, plane = L.OUT_BF16, 0
        else:
            a = torch.empty(2, rows, kpad, device='cuda')
            kind, plane = L.OUT_F32_SPLIT, rows * kpad
        L.check(lib.devit_im2col3x3(xin.data_ptr(), a.data_ptr(), B, cin, hw, *st, kpad, kind,
                                    plane, L.stream_ptr()))
        ck = L.OUT_BF16 if prec == L.DEVIT_BF16 else L.OUT_F32
        c = L.gemm(a, L.to_operand(wk.cuda(), prec), precision=prec, act=L.ACT_RELU, out_kind=ck)
        out = torch.empty(B, (hw // 2) ** 2, cout, device='cuda')
        posd = pos.cuda()
        L.check(lib.devit_maxpool3x3s2_cl(c.data_ptr(), ck, out.data_ptr(), posd.data_ptr(),
                                          B, hw, cout, L.stream_ptr()))
        torch.cuda.synchronize()
        assert rel(out, ref) < tol, (prec, rel(out, ref))


@pytest.mark.parametrize('tokens,dim', [(256, 256), (64, 256), (100, 384)])
def test_seqpool(tokens, dim):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(5, tokens, dim, generator=g)
    w = torch.randn(dim, generator=g) * 0.2
    b = 0.3
    p = F.softmax(x @ w + b, dim=1)
    ref = torch.einsum('bt,btd->bd', p, x)
    out = torch.empty(5, dim, device='cuda')
    xd, wd = x.cuda(), w.cuda()   # keep the device tensors alive across the raw-pointer call
    L.check(L.load().devit_seqpool(xd.data_ptr(), wd.data_ptr(), b, out.data_ptr(), 5, tokens, dim,
                                   L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('name,n_conv,tokens', [('3x1', 1, 256), ('3x2', 2, 64)])
def test_cct_single_model_vs_reference_golden(precision, name, n_conv, tokens):
    m = cct.get_decct(num_classes=100, kernel_size=3, n_conv_layers=n_conv, img_size=32)
    sd = synth.cct_state_dict(0, n_conv=n_conv, tokens=tokens, num_classes=100)
    m.load_state_dict(sd)
    m = m.cuda().eval().set_precision(precision)
    x = synth.cifar_images(4).cuda()
    logits, pool = m(x, output_pool=True)
    assert rel(pool, G[f'pool_{name}']) < TOL[precision], rel(pool, G[f'pool_{name}'])
    r = rel(logits, G[f'logits_{name}'])
    assert r < TOL[precision], r
    if precision == 'fp32':
        assert (logits.argmax(-1).cpu().numpy() == G[f'logits_{name}'].argmax(-1)).all()
    # residual stream after 0 / 1 / all blocks against the oracle
    with torch.no_grad():
        _, hidden = CO.pooled_features(sd, x.cpu(), n_conv, 7, 4, return_hidden=True)
    for nl in (0, 1, 7):
        xo = torch.empty(4, tokens, 256, device='cuda')
        m.pooled_features(x, x_out=xo, num_layers=nl)
        assert rel(xo, hidden[nl]) < TOL[precision], (nl, rel(xo, hidden[nl]))


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag,ts', [('plain', None), ('mlp', 512)])
def test_cct_ensemble_vs_reference_golden(precision, tag, ts):
    n_sub = 4
    multi = cct.MultiCCT('decct_7_3x1', num_classes_list=[25] * n_sub, num_sub_models=n_sub,
                         input_size=32)
    for s in range(n_sub):
        multi.models[s].load_state_dict(synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True))
    fuse = cct.EnsembleCCT(sub_size=256, teacher_size=ts, num_sub_models=n_sub, num_classes=100)
    fuse.load_state_dict(synth.ensemble_cct_state_dict(n_sub, 256, ts, 100))
    multi = multi.cuda().eval().set_precision(precision)
    fuse = fuse.cuda().eval().set_precision(precision)
    x = synth.cifar_images(4).cuda()
    feats = multi(x)
    assert rel(torch.stack(feats), G['ens_feats']) < TOL[precision]
    logits = fuse(feats)
    r = rel(logits, G[f'ens_logits_{tag}'])
    assert r < TOL[precision], r
    if precision == 'fp32':
        assert (logits.argmax(-1).cpu().numpy() == G[f'ens_logits_{tag}'].argmax(-1)).all()


def test_cct_ragged_and_large_batch_consistency():
    """Batch 1 / 7 / 130 give the same per-image logits (no cross-image leakage, ragged M)."""
    m = cct.get_decct(num_classes=100, kernel_size=3, n_conv_layers=1, img_size=32)
    m.load_state_dict(synth.cct_state_dict(0, n_conv=1, tokens=256, num_classes=100))
    m = m.cuda().eval().set_precision('bf16')
    x = synth.cifar_images(130, seed=9).cuda()
    full = m(x)
    assert torch.equal(m(x[:1]), full[:1])
    assert torch.equal(m(x[3:10]), full[3:10])
