"""Pins the CPU oracle (oracle/devit_oracle.py) to the reference: every fixture in
tests/golden/devit_golden.npz was produced by the unmodified reference modules
(tests/golden/make_golden.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest
import torch

from devit_b200 import synth
from oracle import devit_oracle as O

G = np.load(Path(__file__).parent / 'golden' / 'devit_golden.npz')
N_SUB, B = 4, 4


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope='module')
def sds():
    return [synth.dedeit_state_dict(s, with_heads=False) for s in range(N_SUB)]


@pytest.fixture(scope='module')
def x():
    return synth.images(B)


def test_dense_ensemble_matches_reference(sds, x):
    with torch.no_grad():
        logits, cls, dist = O.ensemble_logits(sds, synth.ensmlp_state_dict(N_SUB), x)
    assert rel(torch.stack(cls).numpy(), G['dense_cls']) < 2e-5
    assert rel(torch.stack(dist).numpy(), G['dense_dist']) < 2e-5
    assert rel(logits.numpy(), G['dense_logits']) < 2e-5
    assert (logits.argmax(-1).numpy() == G['dense_logits'].argmax(-1)).all()


def test_shrunk_ensemble_matches_reference(sds, x):
    gates = [synth.shrink_gates(s) for s in range(N_SUB)]
    with torch.no_grad():
        logits, cls, dist = O.ensemble_logits(sds, synth.ensmlp_state_dict(N_SUB), x, gates)
    assert rel(torch.stack(cls).numpy(), G['shrunk_cls']) < 2e-5
    assert rel(torch.stack(dist).numpy(), G['shrunk_dist']) < 2e-5
    assert rel(logits.numpy(), G['shrunk_logits']) < 2e-5
    assert (logits.argmax(-1).numpy() == G['shrunk_logits'].argmax(-1)).all()


def test_per_block_state_matches_reference(sds, x):
    ng, hg = synth.shrink_gates(0)
    with torch.no_grad():
        _, blocks = O.forward_features(sds[0], x, 6, hg, ng, return_blocks=True)
    assert len(blocks) == 13
    sl = torch.stack([b[:, :4, :16] for b in blocks]).numpy()
    assert rel(sl, G['sub0_block_slice']) < 2e-5
    am = np.array([b.abs().mean().item() for b in blocks])
    assert np.allclose(am, G['sub0_block_absmean'], rtol=1e-5)


def test_single_model_and_stress_and_teacher(x):
    with torch.no_grad():
        a = O.forward_logits(synth.dedeit_state_dict(0, num_classes=25), x)
        b = O.forward_logits(synth.dedeit_state_dict(7, num_classes=25, qkv_gain=3.0), x)
        t = O.forward_logits(synth.teacher_state_dict(100), x[:2], num_heads=12)
    assert rel(a.numpy(), G['single_logits']) < 2e-5
    assert rel(b.numpy(), G['stress_logits']) < 2e-5
    assert rel(t.numpy(), G['teacher_logits']) < 2e-5


def test_gate_index_selection_is_bit_exact():
    """core/imp_rank.py:50-62 / :132-144 and core/compute_metric.py:67 via the reference's own
    functions (golden) vs the oracle restatement and the host mirror (devit_b200/shrink.py)."""
    from devit_b200 import shrink
    for s in range(N_SUB):
        rng = np.random.RandomState(4321 + s)
        n_ratio, h_ratio = shrink.sample_policy(rng)
        assert np.array_equal(np.array(n_ratio + h_ratio), G[f'policy{s}_ratios'])
        n_rank = [rng.permutation(1536) for _ in range(12)]
        h_rank = [rng.permutation(6) for _ in range(12)]
        nm = np.stack([O.keep_mask(1536, n_ratio[i], n_rank[i]) for i in range(12)])
        hm = np.stack([O.keep_mask(6, h_ratio[i], h_rank[i]) for i in range(12)])
        assert np.array_equal(nm.astype(np.uint8), G[f'policy{s}_neuron_masks'])
        assert np.array_equal(hm.astype(np.uint8), G[f'policy{s}_head_masks'])
        ng, hg = synth.shrink_gates(s)
        assert np.array_equal(torch.stack(ng).numpy().astype(np.uint8),
                              G[f'policy{s}_neuron_masks'])
        assert np.array_equal(torch.stack(hg).numpy().astype(np.uint8), G[f'policy{s}_head_masks'])
        assert O.shrink_macs(n_ratio, h_ratio) == float(G[f'policy{s}_macs'])
        assert shrink.cal_shrink_macs(n_ratio, h_ratio, emb=384, mlp_ratio=4, seq_length=197,
                                      head=6, layer=12) == float(G[f'policy{s}_macs'])


def test_keep_mask_edge_cases():
    rank = np.arange(6)
    assert O.keep_mask(6, 0.0, rank).sum() == 6
    assert O.keep_mask(6, 1.0, rank).sum() == 0
    assert list(O.kept_indices(O.keep_mask(6, 0.5, rank))) == [3, 4, 5]
    assert list(O.kept_indices(O.keep_mask(6, 0.49, np.array([5, 0, 3, 1, 4, 2])))) == [1, 2, 4]


def test_fp32_oracle_noise_floor_against_fp64(sds, x):
    """the same graph in fp64 is the ground truth that arbitrates fp32
    disagreements.  The fp32 oracle sits ~1e-6 from it -- two orders of magnitude under the 1e-4
    bar the fp32/3xTF32 GPU mode is held to -- and picks the same classes."""
    esd = synth.ensmlp_state_dict(N_SUB)
    gates = [synth.shrink_gates(s) for s in range(N_SUB)]
    with torch.no_grad():
        l32, c32, _ = O.ensemble_logits(sds, esd, x, gates)
        l64, c64, _ = O.ensemble_logits([O.to_dtype(sd, torch.float64) for sd in sds],
                                        O.to_dtype(esd, torch.float64), x.double(), gates)
    assert l64.dtype == torch.float64
    floor = rel(l32.numpy(), l64.numpy())
    assert floor < 5e-6, floor
    assert rel(torch.stack(c32).numpy(), torch.stack(c64).numpy()) < 5e-6
    assert torch.equal(l32.argmax(-1), l64.argmax(-1))
