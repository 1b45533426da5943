"""Times MultiViT.forward_slab for the sub-model / batch shapes one rank sees at N = 1, 2, 4, 8
GPUs (4-way shrunk ensemble, global batch 256) under different chain settings, CUDA-graph
replayed, on ONE GPU:   python tools/time_chain.py
Columns: subs on the rank, images, DEVIT_CHAINS, DEVIT_MIN_CHUNK, DEVIT_SM_SHARE, ms per step."""
import itertools
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import ensemble, shrink, synth  # noqa: E402

mv = ensemble.MultiViT(model='dedeit', drop=0, drop_path=0.1, num_classes_list=[25] * 4, num_div=4)
for s in range(4):
    mv.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
    ng, hg = synth.shrink_gates(s)
    shrink.mlp_neuron_shrink(mv.backbones[s], ng)
    shrink.attn_head_shrink(mv.backbones[s], hg)
mv = mv.cuda().eval().set_precision('bf16')
xs = {b: synth.images(b).cuda() for b in (256, 128)}


def timed(subs, x, steps=20):
    for _ in range(3):
        ref = mv.forward_slab(x, subs)[0]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = mv.forward_slab(x, subs)[0]
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out.clone()


shapes = [([0], 128), ([0], 256), ([0, 1], 256), ([0, 1, 2, 3], 256)]
settings = [(1, 64, 1), (2, 64, 1), (2, 64, 2), (4, 32, 1), (4, 32, 2), (8, 16, 1), (4, 64, 1),
            (4, 64, 2), (8, 32, 1), (8, 32, 2), (16, 16, 1)]
for subs, b in shapes:
    base = None
    for chains, mc, share in settings:
        if b // mc * len(subs) < chains and (chains, mc) != (1, 64):
            continue  # this many chains do not exist at this shape
        os.environ.update(DEVIT_CHAINS=str(chains), DEVIT_MIN_CHUNK=str(mc),
                          DEVIT_SM_SHARE=str(share),
                          DEVIT_SUB_STREAMS='1' if chains == 1 else str(min(chains, 8)))
        ms, out = timed(subs, xs[b])
        if base is None:
            base = out
        same = torch.equal(out, base)
        print(f'subs={len(subs)} images={b} chains={chains} min_chunk={mc} share={share}: '
              f'{ms:.3f} ms  bit_identical_to_single_chain={same}', flush=True)
