"""End-to-end parity of the CUDA path (through the nn.Module boundary -> ctypes -> C ABI) against
(1) the golden fixtures produced by the unmodified reference and (2) the CPU oracle on the same
seeded inputs.  Tolerances are the north-star's: rel = max|a-b| / max|b| <= 1e-4 in the
fp32/3xTF32 mode, <= 2e-2 in bf16; argmax and kept-index lists bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

from devit_b200 import _lib as L
from devit_b200 import ensemble, shrink, synth
from devit_b200.registry import create_model
from oracle import devit_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).parent / 'golden' / 'devit_golden.npz')
TOL = {'fp32': 1e-4, 'bf16': 2e-2}
N_SUB, B = 4, 4


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def argmax_agrees(logits, ref, err):
    """argmax must match wherever the reference's top-1/top-2 margin exceeds twice the observed
    absolute error (closer ties are not decidable at that precision and are reported)."""
    ref = np.asarray(ref, np.float64)
    got = logits.detach().double().cpu().numpy()
    top2 = np.sort(ref, -1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    decidable = margin > 2 * err
    same = got.argmax(-1) == ref.argmax(-1)
    assert same[decidable].all(), (margin, same)
    return int(decidable.sum()), int(same.sum())


def make_sub(s, precision, gates=None, num_classes=25, with_heads=True, **kw):
    m = create_model('dedeit', num_classes=num_classes)
    m.load_state_dict(synth.dedeit_state_dict(s, num_classes=num_classes, **kw))
    m = m.cuda().eval().set_precision(precision)
    if gates is not None:
        shrink.mlp_neuron_shrink(m, gates[0])
        shrink.attn_head_shrink(m, gates[1])
    return m


def make_ensemble(precision, shrunk):
    mv = ensemble.MultiViT(model='dedeit', drop=0, drop_path=0.1, num_classes_list=[25] * N_SUB,
                           num_div=N_SUB)
    for s in range(N_SUB):
        mv.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
        if shrunk:
            ng, hg = synth.shrink_gates(s)
            shrink.mlp_neuron_shrink(mv.backbones[s], ng)
            shrink.attn_head_shrink(mv.backbones[s], hg)
    fuse = ensemble.EnsMLP(model='dedeit', num_class=100, sub_size=384,
                           num_classes_list=[25] * N_SUB, teacher_size=768)
    fuse.load_state_dict(synth.ensmlp_state_dict(N_SUB))
    return mv.cuda().eval().set_precision(precision), fuse.cuda().eval().set_precision(precision)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('shrunk', [False, True])
def test_ensemble_vs_reference_golden(precision, shrunk):
    mv, fuse = make_ensemble(precision, shrunk)
    x = synth.images(B).cuda()
    cls, dist = mv(x)
    logits = fuse((cls, dist))
    tag = 'shrunk' if shrunk else 'dense'
    tol = TOL[precision]
    assert rel(torch.stack(list(cls)), G[f'{tag}_cls']) < tol
    assert rel(torch.stack(list(dist)), G[f'{tag}_dist']) < tol
    r = rel(logits, G[f'{tag}_logits'])
    assert r < tol, r
    err = np.abs(logits.double().cpu().numpy() - G[f'{tag}_logits']).max()
    dec, same = argmax_agrees(logits, G[f'{tag}_logits'], err)
    print(f'[{precision} {tag}] rel={r:.3e} abs_err={err:.3e} decidable={dec}/{B} same={same}/{B}')
    if precision == 'fp32':
        assert same == B and dec == B  # fp32 mode: argmax identical on every sample


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_eight_way_1000_classes_vs_oracle(precision):
    """BASELINE config C3 at a small batch: 8-way decomposition, ImageNet-1K fusion head
    (8 K-segments of 384 into the 768-wide mlp, 1000-class classifiers), dense gates."""
    n_sub, n_cls, bsz = 8, 1000, 3
    x = synth.images(bsz, seed=77)
    sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(n_sub)]
    esd = synth.ensmlp_state_dict(n_sub, num_class=n_cls)
    with torch.no_grad():
        ref, _, _ = O.ensemble_logits(sds, esd, x, None)
    mv = ensemble.MultiViT(model='dedeit', drop=0, drop_path=0.1,
                           num_classes_list=[n_cls // n_sub] * n_sub, num_div=n_sub)
    fuse = ensemble.EnsMLP(model='dedeit', num_class=n_cls, sub_size=384,
                           num_classes_list=[n_cls // n_sub] * n_sub, teacher_size=768)
    for s in range(n_sub):
        mv.backbones[s].load_state_dict(sds[s])
    fuse.load_state_dict(esd)
    mv = mv.cuda().eval().set_precision(precision)
    fuse = fuse.cuda().eval().set_precision(precision)
    logits = fuse(mv(x.cuda()))
    assert logits.shape == (bsz, n_cls)
    r = rel(logits, ref)
    assert r < TOL[precision], r
    if precision == 'fp32':
        assert torch.equal(logits.argmax(-1).cpu(), ref.argmax(-1))


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_plain_list_input_to_ensmlp(precision):
    """EnsMLP must also accept ordinary lists of tensors (not views of our slab)."""
    _, fuse = make_ensemble(precision, False)
    cls = [torch.from_numpy(G['dense_cls'][s]).cuda() for s in range(N_SUB)]
    dist = [torch.from_numpy(G['dense_dist'][s]).cuda() for s in range(N_SUB)]
    logits = fuse((cls, dist))
    assert rel(logits, G['dense_logits']) < TOL[precision]


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_per_block_residual_stream(precision):
    gates = synth.shrink_gates(0)
    m = make_sub(0, precision, gates, with_heads=True)
    x = synth.images(B)
    with torch.no_grad():
        _, blocks = O.forward_features(synth.dedeit_state_dict(0, with_heads=False), x, 6,
                                       gates[1], gates[0], return_blocks=True)
    xc = x.cuda()
    for nl in (0, 1, 2, 6, 12):
        xo = torch.empty(B, 198, 384, device='cuda')
        m.features_into(xc, x_out=xo, num_layers=nl)
        r = rel(xo, blocks[nl])
        assert r < TOL[precision], (nl, r)
        assert rel(xo[:, :4, :16], G['sub0_block_slice'][nl]) < TOL[precision]


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_single_model_logits_stress_and_teacher(precision):
    x = synth.images(B).cuda()
    m = make_sub(0, precision)
    assert rel(m(x), G['single_logits']) < TOL[precision]
    st = make_sub(7, precision, qkv_gain=3.0)
    assert rel(st(x), G['stress_logits']) < TOL[precision]
    t = create_model('deit_base_distilled_patch16_224', num_classes=100)
    t.load_state_dict(synth.teacher_state_dict(100))
    t = t.cuda().eval().set_precision(precision)
    out = t(x[:2])
    assert out.shape == (2, 100)
    assert rel(out, G['teacher_logits']) < TOL[precision]


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('bsz', [1, 3, 7])
def test_ragged_batches_vs_oracle(precision, bsz):
    gates = synth.shrink_gates(2)
    m = make_sub(2, precision, gates)
    x = synth.images(bsz, seed=99)
    with torch.no_grad():
        c, d = O.forward_features(synth.dedeit_state_dict(2, with_heads=False), x, 6, gates[1],
                                  gates[0])
    out = m.forward_features(x.cuda())['output']
    assert rel(out[0], c) < TOL[precision] and rel(out[1], d) < TOL[precision]


def test_gate_edge_cases_vs_oracle():
    """All heads of a layer gated off, a single kept head, ragged neuron counts (1, 17, 1535),
    non-binary gate values -- compaction must stay exact."""
    sd = synth.dedeit_state_dict(3, with_heads=False)
    ng = [torch.ones(1536) for _ in range(12)]
    hg = [torch.ones(6) for _ in range(12)]
    hg[0] = torch.zeros(6)
    hg[1] = torch.tensor([0., 0., 1., 0., 0., 0.])
    hg[2] = torch.tensor([1., 0.5, 0., 2., 0., 1.])
    ng[0] = torch.zeros(1536); ng[0][77] = 1
    ng[1] = torch.zeros(1536); ng[1][torch.arange(0, 1536, 91)] = 1   # 17 kept
    ng[2] = torch.ones(1536); ng[2][5] = 0                             # 1535 kept
    ng[3] = torch.zeros(1536)                                          # nothing kept
    ng[4] = torch.rand(1536, generator=torch.Generator().manual_seed(5)).round() * 0.5
    x = synth.images(2, seed=5)
    with torch.no_grad():
        c, d = O.forward_features(sd, x, 6, hg, ng)
    m = make_sub(3, 'fp32', (ng, hg))
    out = m.forward_features(x.cuda())['output']
    assert rel(out[0], c) < 1e-4 and rel(out[1], d) < 1e-4
    pk = m.packed()
    for i in range(12):
        want = O.kept_indices(hg[i].numpy())
        if len(want):
            assert np.array_equal(pk.kept_heads[i].numpy(), want)
        assert np.array_equal(pk.kept_neurons[i].numpy(), O.kept_indices(ng[i].numpy()))


def test_gate_set_restore_invalidates_pack():
    m = make_sub(1, 'fp32')
    x = synth.images(2, seed=3).cuda()
    dense = m.forward_features(x)['output'][0].clone()
    ng, hg = synth.shrink_gates(1)
    shrink.mlp_neuron_shrink(m, ng)
    shrink.attn_head_shrink(m, hg)
    shrunk = m.forward_features(x)['output'][0].clone()
    assert rel(shrunk, dense) > 1e-2
    shrink.mlp_neuron_restore(m)
    shrink.attn_head_restore(m)
    assert torch.equal(m.forward_features(x)['output'][0], dense)
    with torch.no_grad():
        m.blocks[0].mlp.fc1.weight.mul_(1.5)   # in-place parameter edit must be noticed too
    assert rel(m.forward_features(x)['output'][0], dense) > 1e-3


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_layerwise_outputs_and_observers(precision):
    """output_qkv / output_att / output_encoders and the lazily materialised observers
    (neuron_output, head_output) against the oracle's masked-dense intermediates."""
    gates = synth.shrink_gates(0)
    sd = synth.dedeit_state_dict(0, with_heads=False)
    m = make_sub(0, precision, gates)
    x = synth.images(2)
    tol = TOL[precision]
    with torch.no_grad():
        xe = O.embed_tokens(sd, x)
        ln = torch.nn.functional.layer_norm(xe, (384,), sd['blocks.0.norm1.weight'],
                                            sd['blocks.0.norm1.bias'], 1e-6)
        y, head_out, (q, k, v) = O.attention(sd, 'blocks.0.attn.', ln, 6, gates[1][0], True)
        x1 = xe + y
        ln2 = torch.nn.functional.layer_norm(x1, (384,), sd['blocks.0.norm2.weight'],
                                             sd['blocks.0.norm2.bias'], 1e-6)
        _, hidden = O.mlp(sd, 'blocks.0.mlp.', ln2, gates[0][0])
        feats, blocks = O.forward_features(sd, x, 6, gates[1], gates[0], return_blocks=True)
    out = m.forward_features(x.cuda(), output_qkv=True, output_att=True, output_emb=True,
                             output_encoders=True)
    assert len(out['qkv']) == 12 and len(out['attention']) == 12 and len(out['encoder']) == 13
    gq, gk, gv = out['qkv'][0]
    assert gq.shape == (2, 6, 198, 64)
    assert rel(gq, q) < tol and rel(gk, k) < tol and rel(gv, v) < tol
    assert rel(out['attention'][0], y) < tol
    assert rel(out['encoder'][0], blocks[0]) < tol and rel(out['encoder'][12], blocks[12]) < tol
    assert rel(out['output'][0], feats[0]) < tol
    assert rel(m.blocks[0].attn.head_output, head_out) < tol
    assert rel(m.blocks[0].mlp.neuron_output, hidden) < tol
    # observers after a plain fused forward are materialised lazily on first read
    m2 = make_sub(0, precision, gates)
    m2(x.cuda())
    assert m2.blocks[0].mlp.neuron_output.shape == (2, 198, 1536)
    assert rel(m2.blocks[0].mlp.neuron_output, hidden) < tol
    assert rel(m2.blocks[0].attn.head_output, head_out) < tol
    d = m2(x.cuda(), distill_token=True)
    assert set(d) >= {'output', 'last_tokens', 'qkv', 'attention', 'encoder'}


def test_devit_not_distilled():
    sd = synth.vit_state_dict(55, dim=384, depth=12, num_classes=10, distilled=False)
    m = create_model('devit', num_classes=10)
    m.load_state_dict(sd)
    m = m.cuda().eval().set_precision('fp32')
    x = synth.images(2, seed=8)
    with torch.no_grad():
        ref = O.forward_logits(sd, x)
    assert rel(m(x.cuda()), ref) < 1e-4


@pytest.mark.parametrize('precision', ['bf16', 'fp32'])
def test_full_size_batch_is_image_independent(precision):
    """BASELINE size (bs 256): the golden 4-image batch tiled 64x must reproduce the golden
    features for every copy -- a size-independent property that exercises the full-size grids,
    the M-tile boundaries (256*198 rows) and every image slot."""
    bsz = 256 if precision == 'bf16' else 64
    gates = synth.shrink_gates(0)
    m = make_sub(0, precision, gates)
    x = synth.images(B).cuda().repeat(bsz // B, 1, 1, 1)
    out = m.forward_features(x)['output']
    c = out[0].view(bsz // B, B, 384)
    assert rel(c[0], G['shrunk_cls'][0]) < TOL[precision]
    assert torch.equal(c, c[:1].expand_as(c)), "copies of the same image must agree bit-for-bit"


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_output_qkv_on_the_fused_path(precision):
    """output_qkv=True with no head gated off runs the fused forward with per-layer q/k/v exports
    (devit_vit_forward_ex): same logits as the plain forward bit for bit, q/k/v equal to the
    layer-wise path's and to the oracle's, uint8 input included; gated heads fall back."""
    sd = synth.dedeit_state_dict(0)
    m = make_sub(0, precision)
    x = synth.images(2)
    tol = TOL[precision]
    with torch.no_grad():
        xe = O.embed_tokens(sd, x)
        ln = torch.nn.functional.layer_norm(xe, (384,), sd['blocks.0.norm1.weight'],
                                            sd['blocks.0.norm1.bias'], 1e-6)
        _, _, (q, k, v) = O.attention(sd, 'blocks.0.attn.', ln, 6, None, True)
        feats, blocks = O.forward_features(sd, x, 6, None, None, return_blocks=True)
        ln5 = torch.nn.functional.layer_norm(blocks[5], (384,), sd['blocks.5.norm1.weight'],
                                             sd['blocks.5.norm1.bias'], 1e-6)
        _, _, (q5, k5, v5) = O.attention(sd, 'blocks.5.attn.', ln5, 6, None, True)
    out = m(x.cuda(), output_qkv=True)
    assert rel(out['output'], m(x.cuda())) < 1e-6
    assert len(out['qkv']) == 12 and all(t is not None for t in out['qkv'])
    gq, gk, gv = out['qkv'][0]
    assert gq.shape == (2, 6, 198, 64)
    assert rel(gq, q) < tol and rel(gk, k) < tol and rel(gv, v) < tol
    g5 = out['qkv'][5]
    assert rel(g5[0], q5) < tol and rel(g5[1], k5) < tol and rel(g5[2], v5) < tol
    # the layer-wise path (any other flag) gives the same tensors within the mode's rounding
    lw = m(x.cuda(), output_qkv=True, output_att=True)
    assert rel(lw['qkv'][5][2], g5[2].float()) < tol
    # ... and the same dtype whichever path served the call
    assert g5[2].dtype == lw['qkv'][5][2].dtype == torch.float32
    # selected layers only
    m.export_qkv_layers = [5]
    sel = m(x.cuda(), output_qkv=True)['qkv']
    assert [t is not None for t in sel] == [i == 5 for i in range(12)]
    assert torch.equal(sel[5][1], g5[1])
    m.export_qkv_layers = None
    # a gated head -> layer-wise path (dense q/k/v like the reference)
    gates = synth.shrink_gates(0)
    mg = make_sub(0, precision, gates)
    og = mg(x.cuda(), output_qkv=True)
    assert og['qkv'][0][0].shape == (2, 6, 198, 64) and rel(og['qkv'][0][0], q) < tol


def test_teacher_output_qkv_fused_tuple_api():
    t = create_model('deit_base_distilled_patch16_224', num_classes=100)
    t.load_state_dict(synth.teacher_state_dict())
    t = t.cuda().eval().set_precision('bf16')
    x = synth.images(2).cuda()
    feats, qkvs, att, enc = t.forward_features(x, output_qkv=True)
    assert len(qkvs) == 12 and att == [] and enc == []
    assert qkvs[5][0].shape == (2, 12, 198, 64) and qkvs[5][0].dtype == torch.float32
    lw = t.forward_features(x, output_qkv=True, output_att=True)[1]
    assert rel(qkvs[5][0], lw[5][0]) < 2e-2 and rel(qkvs[11][2], lw[11][2]) < 2e-2


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_collapsed_fusion_head_matches_the_four_gemm_head(precision, monkeypatch):
    """Eval-time EnsMLP = ONE pre-multiplied K-segmented GEMM (no activation sits between the two
    Linears of a token kind, models/ensemble_models.py:79-84); it must agree with the four-GEMM
    evaluation and with the reference golden logits."""
    mv, fuse = make_ensemble(precision, True)
    x = synth.images(B).cuda()
    monkeypatch.setenv('DEVIT_COLLAPSE_HEAD', '0')
    four = fuse(mv(x))
    monkeypatch.setenv('DEVIT_COLLAPSE_HEAD', '1')
    c0 = L.load().devit_launch_count()
    feats = mv(x)
    c1 = L.load().devit_launch_count()
    one = fuse(feats)
    assert L.load().devit_launch_count() - c1 == 1  # the whole head is a single launch
    assert c1 > c0
    assert rel(one, four) < (2e-5 if precision == 'fp32' else 8e-3)
    assert rel(one, G['shrunk_logits']) < TOL[precision]
    # distill=True in training mode needs the 768-wide tokens: that path keeps the two-level form
    fuse.train()
    tok, logits = fuse(mv(x), distill=True)
    fuse.eval()
    assert tok[0].shape == (B, 768) and rel(logits, G['shrunk_logits']) < TOL[precision]


def test_ensmlp_does_not_trust_a_stale_slab():
    """the FeatureList carries a second copy of the features (the operand slab); an
    entry that was replaced or edited in place must win over the slab, like torch.stack(list)."""
    mv, fuse = make_ensemble('fp32', False)
    x = synth.images(B).cuda()
    cls, dist = mv(x)
    base = fuse((cls, dist))
    # replace one entry: the result must follow the new tensor, not the stale slab
    cls2, dist2 = mv(x)
    cls2[1] = torch.zeros_like(cls2[1])
    out = fuse((cls2, dist2))
    ref_cls = [c.clone() for c in cls]
    ref_cls[1].zero_()
    want = fuse(([t for t in ref_cls], [t.clone() for t in dist]))
    assert rel(out, want) < 1e-5 and rel(out, base) > 1e-3
    # in-place edit of an entry (bumps the version counter of the slab storage)
    cls3, dist3 = mv(x)
    cls3[2].mul_(0.0)
    out3 = fuse((cls3, dist3))
    ref3 = [c.clone() for c in cls]
    ref3[2].zero_()
    want3 = fuse((ref3, [t.clone() for t in dist]))
    assert rel(out3, want3) < 1e-5


def test_batch_pipeline_two_batches_in_flight():
    """parallel.BatchPipeline: two slots (own stream / CUDA graph / workspace / side streams each)
    replayed round-robin give, slot for slot, exactly the plain forward's logits -- also for a rank
    that runs two sub-models as concurrent chains inside every slot."""
    from devit_b200 import parallel
    mv, fuse = make_ensemble('bf16', True)
    xs = [synth.images(B, seed=11).cuda(), synth.images(B, seed=12).cuda()]
    want = [fuse(mv(x)).clone() for x in xs]
    plan = parallel.shard_plan(1, 0, N_SUB, B)
    slots = [parallel.ShardedEnsemble(mv, fuse, plan) for _ in xs]
    pipe = parallel.BatchPipeline([(lambda e=e, x=x: e(x)) for e, x in zip(slots, xs)])
    assert pipe.depth == 2
    for rep in range(3):
        pipe.fork()
        for _ in range(4):
            pipe.launch()
        pipe.join()
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(pipe.outs[k], want[k]), (rep, k)
