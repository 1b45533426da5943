"""Would pipelining consecutive batches help a rank that owns ONE sub-model (N = 4 / 8)?  Emulated on
one GPU: sub-model 0 on 256 (128) images, (a) one CUDA graph replayed back to back, (b) two graphs
(separate workspaces / outputs) alternating on two streams with half-chip grids, (c) the same with
whole-chip grids.   python tools/time_pipeline.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import _lib as L, ensemble, shrink, synth  # noqa: E402

mv = ensemble.MultiViT(model='dedeit', drop=0, drop_path=0.1, num_classes_list=[25] * 4, num_div=4)
for s in range(4):
    mv.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
    ng, hg = synth.shrink_gates(s)
    shrink.mlp_neuron_shrink(mv.backbones[s], ng)
    shrink.attn_head_shrink(mv.backbones[s], hg)
mv = mv.cuda().eval().set_precision('bf16')
os.environ['DEVIT_SUB_STREAMS'] = '1'
lib = L.load()


def capture(x, subs, stream, budget):
    with torch.cuda.stream(stream):
        lib.devit_set_sm_budget(budget)
        for _ in range(2):
            mv.forward_slab(x, subs)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            out = mv.forward_slab(x, subs)[0]
        lib.devit_set_sm_budget(0)
    torch.cuda.synchronize()
    return g, out


for images in (256, 128):
    for subs in ([0], [0, 1]):
        x = synth.images(images).cuda()
        steps = 20
        res = []
        for mode, budget in (('one graph', 0), ('two graphs, half-chip grids', 74),
                             ('two graphs, whole-chip grids', 0)):
            n_g = 1 if mode == 'one graph' else 2
            streams = [torch.cuda.Stream() for _ in range(n_g)]
            graphs = [capture(x, subs, st, budget) for st in streams]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream()
            for rep in range(2):
                e0.record(cur)
                for st in streams:
                    st.wait_stream(cur)
                for i in range(steps):
                    with torch.cuda.stream(streams[i % n_g]):
                        graphs[i % n_g][0].replay()
                for st in streams:
                    cur.wait_stream(st)
                e1.record(cur)
                torch.cuda.synchronize()
            res.append((mode, e0.elapsed_time(e1) / steps))
            same = all(torch.equal(graphs[0][1], gg[1]) for gg in graphs)
            assert same
        print(f'subs={len(subs)} images={images}: ' + ' | '.join(f'{m}: {t:.3f} ms/step' for m, t in res),
              flush=True)
