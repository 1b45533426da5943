"""Host-side logic that needs no GPU: C-ABI export surface, registry / constructor / state_dict
compatibility, the gate protocol, gate-aware packing (integer index selection), and the
loud-failure rule (no CPU fallback)."""
import ctypes
import os
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from devit_b200 import _lib as L
from devit_b200 import ensemble, models, packing, shrink, synth
from devit_b200.registry import create_model, is_model
from oracle import devit_oracle as O

ROOT = Path(__file__).resolve().parents[1]


def test_abi_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'devit_b200.h').read_text()
    declared = set(re.findall(r'\b(devit_[a-z0-9_]+)\s*\(', header))
    assert declared, "no prototypes found in include/devit_b200.h"
    lib = ctypes.CDLL(str(L.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.exported_symbols()), "ctypes table and header disagree"
    assert L.load().devit_abi_version() == L.ABI_VERSION


def test_no_cpu_fallback():
    m = create_model('dedeit', num_classes=10).eval()
    with pytest.raises(L.DevitError):
        m(torch.zeros(1, 3, 224, 224))
    if not torch.cuda.is_available():
        assert L.load().devit_device_check() != 0
    fuse = ensemble.EnsMLP(model='dedeit', num_class=10, sub_size=384, num_classes_list=[5, 5],
                           teacher_size=768)
    with pytest.raises(L.DevitError):
        fuse(([torch.zeros(2, 384)] * 2, [torch.zeros(2, 384)] * 2))


def test_registry_and_none_kwarg_filtering():
    assert is_model('dedeit') and is_model('devit') and is_model('deit_base_distilled_patch16_224')
    m = create_model('dedeit', num_classes=25, drop_rate=0, drop_path_rate=0.1,
                     drop_block_rate=None)  # models/ensemble_models.py:23-27
    assert m.embed_dim == 384 and m.num_tokens == 2 and len(m.blocks) == 12
    t = create_model('deit_base_distilled_patch16_224', num_classes=100)
    assert t.embed_dim == 768 and t.blocks[0].attn.num_heads == 12 and t.tuple_api


def test_state_dict_names_and_order_match_reference_layout():
    m = create_model('dedeit', num_classes=25)
    keys = list(m.state_dict())
    assert len(keys) == 155
    assert keys == synth.vit_keys(12, True, True)
    for k, shape in synth.vit_shapes(num_classes=25).items():
        assert tuple(m.state_dict()[k].shape) == shape, k
    m.load_state_dict(synth.dedeit_state_dict(0, num_classes=25))  # strict
    v = create_model('devit', num_classes=10)
    assert list(v.state_dict()) == synth.vit_keys(12, False, True)
    mv = ensemble.MultiViT(model='dedeit', num_classes_list=[25] * 4, num_div=4)
    assert len(mv.state_dict()) == 4 * 151  # heads deleted (ensemble.py:233-237 skips 4 keys)
    e = ensemble.EnsMLP(model='dedeit', num_class=100, sub_size=384, num_classes_list=[25] * 4,
                        teacher_size=768)
    assert list(e.state_dict()) == list(synth.ensmlp_state_dict(4))


def test_gate_protocol_with_mirror_functions():
    m = create_model('dedeit', num_classes=25)
    mlps = [x for x in m.modules() if shrink._is_mlp(x)]
    attns = [x for x in m.modules() if shrink._is_attn(x)]
    assert len(mlps) == 12 and len(attns) == 12  # Block / model contain both names -> skipped
    assert all(x.hidden_features == 1536 and x.gate.shape == (1536,) for x in mlps)
    assert all(x.num_heads == 6 and x.gate.shape == (6,) for x in attns)
    ng, hg = synth.shrink_gates(0)
    v0 = packing.module_version(m)
    shrink.mlp_neuron_shrink(m, ng)
    shrink.attn_head_shrink(m, hg)
    assert packing.module_version(m) != v0  # packs are invalidated by a gate assignment
    assert shrink.check_head_sparsity(m) == [(6 - g.sum().item()) / 6 for g in hg]
    shrink.mlp_neuron_restore(m)
    shrink.attn_head_restore(m)
    assert all(x.gate.sum().item() == 1536 for x in mlps)
    v1 = packing.module_version(m)
    mlps[0].gate[3] = 0  # in-place edits are noticed too
    assert packing.module_version(m) != v1


@pytest.mark.skipif(not os.path.isdir('/root/reference/core'), reason="reference not mounted")
def test_gate_protocol_with_the_reference_functions_unchanged():
    """The reference's own core/imp_rank.py functions must find and set gates on our modules."""
    from oracle import ref_shim
    ref_shim.install()
    from core import imp_rank
    m = create_model('dedeit', num_classes=25)
    rng = np.random.RandomState(0)
    ratio = [0.3] * 12
    n_rank = [rng.permutation(1536) for _ in range(12)]
    h_rank = [rng.permutation(6) for _ in range(12)]
    nm = imp_rank.mlp_neuron_mask(m, ratio, n_rank)
    hm = imp_rank.attn_head_mask(m, ratio, h_rank)
    assert len(nm) == 12 and len(hm) == 12
    imp_rank.mlp_neuron_shrink(m, nm)
    imp_rank.attn_head_shrink(m, hm)
    assert imp_rank.check_head_sparsity(m) == [2 / 6] * 12
    ours = shrink.mlp_neuron_mask(m, ratio, n_rank)
    assert all(torch.equal(a, b) for a, b in zip(nm, ours))
    imp_rank.mlp_neuron_restore(m)
    imp_rank.attn_head_restore(m)
    assert imp_rank.check_neuron_sparsity(m) == [0.0] * 12


def test_packing_kept_indices_match_oracle():
    ng, hg = synth.shrink_gates(1)
    for i in range(12):
        assert np.array_equal(packing.kept_indices(ng[i]).numpy(), O.kept_indices(ng[i].numpy()))
        assert np.array_equal(packing.kept_indices(hg[i]).numpy(), O.kept_indices(hg[i].numpy()))
    g = torch.tensor([0., 2., 0., 0.5, 0., 1.])
    assert packing.kept_indices(g).tolist() == [1, 3, 5]


def test_split_tf32_is_exact():
    t = torch.randn(1000) * torch.logspace(-6, 6, 1000)
    s = L.split_tf32(t)
    assert torch.equal(s[0] + s[1], t)
    assert (s[0].view(torch.int32) & 0x1FFF).abs().max().item() == 0
    assert ((s[1].abs() <= t.abs() * 2 ** -11 * 1.0001) | (t == 0)).all()


# ------------------------------------------------------------------------------- LN folding
def test_layernorm_fold_algebra_matches_layernorm_then_linear():
    """packing._fold: LN(x) W^T + b == rstd (x (gamma.W)^T - mean c1) + c2 (fp64 check of the
    identity the folded GEMM epilogues implement, include/devit_b200.h devit_gemm_args)."""
    g = torch.Generator().manual_seed(11)
    d, n, rows, eps = 384, 96, 37, 1e-6
    x = (torch.randn(rows, d, generator=g) * 1.7 + 0.4).double()
    norm = torch.nn.LayerNorm(d, eps=eps).double()
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(d, generator=g))
        norm.bias.copy_(0.1 * torch.randn(d, generator=g))
    w = (torch.randn(n, d, generator=g) * 0.05).double()
    b = (torch.randn(n, generator=g) * 0.1).double()
    kept = []
    wf, c1_ptr, c2 = packing.PackedVit._fold(w.float(), b.float(), norm.float(),
                                             lambda t: kept.append(t) or len(kept) - 1)
    c1 = kept[c1_ptr].double()
    # c1 is the row sum of the bf16-ROUNDED folded weights (what the tensor core multiplies)
    assert torch.equal(kept[c1_ptr], wf.to(torch.bfloat16).float().sum(1))
    mean = x.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + eps)
    wf_q = wf.to(torch.bfloat16).double()                      # operand as stored
    folded = rstd * (x @ wf_q.t() - mean * c1[None, :]) + c2.double()[None, :]
    ref = norm.double()(x) @ w.t() + b
    # the only difference is the bf16 rounding of the folded weights
    assert ((folded - ref).abs().max() / ref.abs().max()).item() < 4e-3
    exact = rstd * (x @ wf.double().t() - mean * wf.double().sum(1)[None, :]) + c2.double()[None, :]
    assert ((exact - ref).abs().max() / ref.abs().max()).item() < 1e-6


# ------------------------------------------------------------------------------- CCT host side
def test_cct_state_dict_layout_registry_and_packing_shapes():
    from devit_b200 import cct
    for n_conv, tokens, backbone in ((1, 256, False), (2, 64, False), (1, 256, True)):
        m = cct.get_decct(num_classes=100, kernel_size=3, n_conv_layers=n_conv, img_size=32,
                          backbone=backbone)
        want = synth.cct_shapes(n_conv=n_conv, tokens=tokens, num_classes=100, backbone=backbone)
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert list(got) == list(want) and got == want
        m.load_state_dict(synth.cct_state_dict(0, n_conv=n_conv, tokens=tokens, num_classes=100,
                                               backbone=backbone))
    assert is_model('cct_7_3x1_32') and is_model('cct_7_3x2_32') and is_model('cct_6_3x1_32')
    m = create_model('cct_7_3x1_32', num_classes=10)
    assert m.classifier.sequence_length == 256 and len(m.classifier.blocks) == 7
    multi = cct.MultiCCT('decct_7_3x2', num_classes_list=[25] * 4, num_sub_models=4, input_size=32)
    assert len(multi.models) == 4 and multi.models[0].backbone
    assert multi.models[0].encoders.sequence_length == 64
    with pytest.raises(L.DevitError):   # loud failure, no CPU path
        m.eval()(torch.zeros(1, 3, 32, 32))
    with pytest.raises(L.DevitError):   # only the decct_*_3xN tokenizer family is built
        cct.CCT(img_size=224, kernel_size=7, stride=2, padding=3)


def test_conv_weight_k_order_matches_im2col_definition():
    """PackedCCT lays conv weights out as [c_out, (ky, kx, c_in)] zero-padded to a multiple of 8:
    check against an explicit im2col of the same definition (devit_im2col3x3)."""
    g = torch.Generator().manual_seed(12)
    cin, cout, hw = 3, 8, 6
    x = torch.randn(2, cin, hw, hw, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    kpad = (9 * cin + 7) // 8 * 8
    wk = torch.zeros(cout, kpad)
    wk[:, :9 * cin] = w.permute(0, 2, 3, 1).reshape(cout, -1)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    a = torch.zeros(2 * hw * hw, kpad)
    for b in range(2):
        for y in range(hw):
            for xx in range(hw):
                patch = xp[b, :, y:y + 3, xx:xx + 3]                 # [c, ky, kx]
                a[(b * hw + y) * hw + xx, :9 * cin] = patch.permute(1, 2, 0).reshape(-1)
    out = (a @ wk.t()).view(2, hw, hw, cout).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(x, w, None, 1, 1)
    assert torch.allclose(out, ref, atol=1e-5)


def test_shard_plan_eight_way():
    from devit_b200 import parallel
    for world in (1, 2, 4, 8):
        seen = set()
        for r in range(world):
            p = parallel.shard_plan(world, r, 8, 1024)
            assert p.group_batch == 1024 and p.num_groups == 1
            seen.update(p.subs)
        assert seen == set(range(8))
    p = parallel.shard_plan(16, 9, 8, 1024)
    assert p.num_groups == 2 and p.group_batch == 512 and p.subs == [1] and p.batch_lo == 512


def test_gemm_tile_schedules_cover_every_tile_exactly_once():
    """Python restatement of the two persistent tile schedules of csrc/gemm.cu (round-robin and
    B-resident): every (m-pair, n-tile) unit is visited exactly once, and in B-resident mode a
    cluster never changes its n-tile."""
    def units(cluster_id, num_clusters, num_units, num_n, bres):
        first, step = cluster_id, num_clusters
        if bres:
            n_fixed = cluster_id % num_n
            group = (num_clusters - n_fixed + num_n - 1) // num_n
            first = (cluster_id // num_n) * num_n + n_fixed
            step = group * num_n
        return list(range(first, num_units, step))

    for num_mp, num_n, clusters in [(198, 3, 74), (198, 4, 74), (198, 5, 74), (99, 3, 74),
                                    (1024, 2, 74), (37, 6, 74), (198, 1, 74), (5, 3, 15)]:
        num_units = num_mp * num_n
        for bres in (False, True):
            if bres and num_n > clusters:
                continue
            seen = []
            for c in range(min(clusters, num_units) if not bres else clusters):
                u = units(c, min(clusters, num_units) if not bres else clusters, num_units, num_n,
                          bres)
                if bres:
                    assert len({x % num_n for x in u}) <= 1
                seen += u
            assert sorted(seen) == list(range(num_units)), (num_mp, num_n, clusters, bres)


def test_token_table_wraps_and_matches_token_assembly():
    """packing.token_table: row j = pos[j] + (cls/dist - patch bias on the prefix rows), first 31
    rows repeated, so that `table[j] + (patch_row W^T + bias)` is models/de_vit.py:258-264."""
    sd = synth.dedeit_state_dict(0, with_heads=False)
    pos = sd['pos_embed'].reshape(-1, 384)
    prefix = torch.cat([sd['cls_token'].reshape(1, -1), sd['dist_token'].reshape(1, -1)], 0)
    bias = sd['patch_embed.proj.bias']
    t = packing.token_table(pos, prefix, bias)
    assert t.shape == (198 + 31, 384) and torch.equal(t[198:], t[:31])
    x = synth.images(1)
    emb = O.embed_tokens(sd, x)[0]                       # reference token assembly
    patches = O.patch_embed(sd, x)[0]                    # A W^T + bias on the patch rows
    assert torch.allclose(t[:2] + bias, emb[:2], atol=1e-6)          # zero patch row + bias
    assert torch.allclose(t[2:198] + patches, emb[2:], atol=1e-5)


def test_u8_layout_detection():
    z = lambda *s: torch.zeros(*s, dtype=torch.uint8)  # noqa: E731
    assert models.u8_layout(z(2, 3, 224, 224)) == L.LAYOUT_NCHW
    assert models.u8_layout(z(2, 1, 32, 32)) == L.LAYOUT_NCHW
    assert models.u8_layout(z(2, 224, 224, 3)) == L.LAYOUT_NHWC
    assert models.u8_layout(z(2, 32, 32, 3)) == L.LAYOUT_NHWC
    for bad in (z(3, 224, 224), z(2, 224, 224, 4), z(2, 5, 32, 32)):
        with pytest.raises(L.DevitError):
            models.u8_layout(bad)


def test_engine_has_no_cpu_fallback():
    from devit_b200 import engine
    m = create_model('dedeit', num_classes=10).eval()
    loader = [(torch.zeros(1, 3, 224, 224), torch.zeros(1, dtype=torch.int64))]
    with pytest.raises(L.DevitError):
        engine.evaluate(loader, m, torch.device('cpu'))
    with pytest.raises(L.DevitError):
        L.eval_tail(torch.zeros(2, 10), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(L.DevitError):
        L.im2col_tokens_u8(torch.zeros(1, 3, 32, 32, dtype=torch.uint8), (0.5,) * 3, (0.5,) * 3, 0)


def test_timm_registration_failures_are_logged_not_swallowed(monkeypatch, caplog):
    """registry.register_model also registers with timm when it is installed; a timm that
    refuses the entrypoint must leave a warning and a record, not silence."""
    import logging
    from devit_b200 import registry

    def refusing(fn):
        raise KeyError("no default_cfg for " + fn.__name__)

    calls = []

    def accepting(fn):
        calls.append(fn.__name__)
        return fn

    def my_model(pretrained=False, **kw):
        return ('built', pretrained, kw)

    monkeypatch.setattr(registry, '_timm_register', refusing)
    with caplog.at_level(logging.WARNING, logger='devit_b200.registry'):
        registry.register_model(my_model)
    assert 'my_model' in registry.TIMM_FAILURES and 'KeyError' in registry.TIMM_FAILURES['my_model']
    assert any('timm refused to register' in r.message for r in caplog.records)
    # the local registry still serves it, with timm's None-kwarg filtering
    assert registry.create_model('my_model', drop_block_rate=None, a=1) == ('built', False, {'a': 1})
    monkeypatch.setattr(registry, '_timm_register', accepting)
    registry.register_model(my_model)
    assert calls == ['my_model']
    registry._ENTRYPOINTS.pop('my_model')
    registry.TIMM_FAILURES.pop('my_model')


def test_chain_scheduling_policy(monkeypatch):
    """ensemble.chain_tasks / batch_chunks / chain_sm_budget: the host-side plan of the concurrent
    kernel chains (no GPU needed: pure bookkeeping)."""
    from devit_b200 import ensemble
    for k in ('DEVIT_CHAINS', 'DEVIT_MIN_CHUNK', 'DEVIT_SM_SHARE', 'DEVIT_SUB_STREAMS'):
        monkeypatch.delenv(k, raising=False)
    # default: one chain per local sub-model, no batch chunking
    assert ensemble.batch_chunks(1, 256) == 1 and ensemble.batch_chunks(4, 256) == 1
    assert ensemble.chain_tasks([0, 2], 256) == [(0, 0, None), (1, 2, None)]
    assert ensemble.sub_streams() == 4
    # asking for ~4 chains on a rank with one sub-model cuts the batch, never below the minimum
    monkeypatch.setenv('DEVIT_CHAINS', '4')
    assert ensemble.batch_chunks(1, 256) == 4 and ensemble.batch_chunks(1, 128) == 2
    assert ensemble.batch_chunks(2, 256) == 2 and ensemble.batch_chunks(4, 256) == 1
    tasks = ensemble.chain_tasks([3], 250)
    assert [t[2] for t in tasks] == [(0, 83), (83, 166), (166, 250)]  # covers every image once
    monkeypatch.setenv('DEVIT_MIN_CHUNK', '200')
    assert ensemble.batch_chunks(1, 256) == 1


def test_collapsed_fusion_head_algebra():
    """EnsMLP._collapsed pre-multiplies the two Linears of a token kind (no activation between
    them, models/ensemble_models.py:79-84): W = classifier.weight @ mlp.weight,
    b = classifier.weight @ mlp.bias + classifier.bias, summed over cls / dist and halved by the
    GEMM's alpha.  Check the algebra against the two-level evaluation in fp64 on the CPU."""
    import torch
    from devit_b200 import ensemble, synth
    n, D, C = 4, 384, 100
    fuse = ensemble.EnsMLP(model='dedeit', num_class=C, sub_size=D,
                           num_classes_list=[25] * n, teacher_size=768)
    fuse.load_state_dict(synth.ensmlp_state_dict(n, num_class=C))
    fuse.set_precision('fp32')
    wcat, bias, wk, bk = fuse._collapsed(torch.device('cpu'))
    w = (wcat[0] + wcat[1]).double()          # hi + lo planes of the split-tf32 operand
    assert w.shape == (C, 2 * n * D)
    g = torch.Generator().manual_seed(3)
    fc = torch.randn(5, n * D, generator=g, dtype=torch.float64)
    fd = torch.randn(5, n * D, generator=g, dtype=torch.float64)
    sd = {k: v.double() for k, v in fuse.state_dict().items()}
    two_level = 0.5 * (
        (fc @ sd['cls_mlp.weight'].t() + sd['cls_mlp.bias']) @ sd['cls_classifier.weight'].t()
        + sd['cls_classifier.bias']
        + (fd @ sd['dist_mlp.weight'].t() + sd['dist_mlp.bias']) @ sd['dist_classifier.weight'].t()
        + sd['dist_classifier.bias'])
    one_gemm = 0.5 * (torch.cat([fc, fd], 1) @ w.t() + bias.double())
    assert (one_gemm - two_level).abs().max() / two_level.abs().max() < 1e-6
    # the per-kind pair used when 2 n > 8 K-segments: cls first (bias 0), dist second (whole bias)
    assert float(bk[0].abs().max()) == 0.0 and torch.equal(bk[1], bias)
    wc, wd = (wk[0][0] + wk[0][1]).double(), (wk[1][0] + wk[1][1]).double()
    assert torch.equal(torch.cat([wc, wd], 1), w)
