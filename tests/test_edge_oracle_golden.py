"""Pins the oracle's eval-transform and eval-tail restatements to
tests/golden/edge_golden.npz, which tests/golden/make_edge_golden.py produced with torchvision's
ToTensor + Normalize and the reference's UNMODIFIED engine.evaluate.  CPU only."""
from pathlib import Path

import numpy as np
import torch

from devit_b200 import engine, synth
from devit_b200.models import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
from oracle import devit_oracle as O

G = np.load(Path(__file__).parent / 'golden' / 'edge_golden.npz')


def test_to_tensor_normalize_is_bit_identical_to_torchvision():
    u8 = torch.from_numpy(G['norm_u8_nhwc'])
    assert torch.equal(u8, synth.images_u8(2, img=32, nhwc=True))
    got = O.to_tensor_normalize(u8, IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, 'nhwc')
    assert np.array_equal(got.numpy(), G['norm_out'])
    nchw = u8.permute(0, 3, 1, 2).contiguous()
    got = O.to_tensor_normalize(nchw, IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, 'nchw')
    assert np.array_equal(got.numpy(), G['norm_out'])


def test_every_byte_value_normalises_like_torchvision():
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1).expand(1, 16, 16, 3)
    got = O.to_tensor_normalize(ramp.contiguous(), IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD,
                                'nhwc')
    assert np.array_equal(got[0].numpy(), G['norm_ramp_out'])


def test_eval_epoch_matches_reference_engine_evaluate():
    for name, sizes, classes in (('c100', (8, 8, 5), 100), ('c1000', (16, 3), 1000),
                                 ('c3', (4, 4), 3)):
        res = O.eval_epoch(synth.eval_batches(sizes, classes))
        ref = G[f'eval_{name}']
        assert abs(res['loss'] - ref[0]) <= 1e-6 * abs(ref[0]), name
        assert res['acc1'] == ref[1] and res['acc5'] == ref[2], name


def test_eval_tail_tie_rule_and_topk_clamp():
    logits = torch.tensor([[1.0, 1.0, 1.0, 0.0], [0.0, 2.0, 2.0, 2.0]])
    # ties: the smaller class index sorts first
    assert O.eval_tail(logits, torch.tensor([0, 1]), topk=2)[1:] == (2, 2)
    assert O.eval_tail(logits, torch.tensor([2, 3]), topk=2)[1:] == (0, 0)
    assert O.eval_tail(logits, torch.tensor([1, 2]), topk=2)[1:] == (0, 2)
    # timm clamps maxk to the number of classes
    assert O.eval_tail(logits, torch.tensor([3, 0]), topk=5)[2] == 2


def test_host_meters_reproduce_metric_logger_averages():
    # the device meters hold {sum of batch-mean losses, #batches, #correct@1, #correct@k, #samples}
    batches = synth.eval_batches((8, 8, 5), 100)
    acc = [0.0] * 5
    for lg, tg in batches:
        l, c1, ck = O.eval_tail(lg, tg)
        acc = [acc[0] + l, acc[1] + 1, acc[2] + c1, acc[3] + ck, acc[4] + lg.shape[0]]
    res = engine.meters_to_dict(acc)
    ref = G['eval_c100']
    assert abs(res['loss'] - ref[0]) <= 1e-6 * abs(ref[0])
    assert abs(res['acc1'] - ref[1]) < 1e-9 and abs(res['acc5'] - ref[2]) < 1e-9
    assert engine.meters_to_dict([0.0] * 5) == {}
