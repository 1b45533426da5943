"""N-rank sharded ensemble vs the same ensemble on one GPU (run under torchrun, 1 rank per GPU):
logits of every data-parallel group must equal the single-GPU logits of that batch slice
(same kernels, fixed K-segment order in the fusion head -> bit-identical), argmax included; the
staged host batch (ShardedEnsemble.stage_batch) must equal the plain device copy.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multigpu.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from devit_b200 import ensemble, parallel, shrink, synth  # noqa: E402

N_SUB, NUM_CLASS, B, D = 4, 100, int(os.environ.get("CHECK_BATCH", "16")), 384


def build(dev):
    multi = ensemble.MultiViT(model="dedeit", drop=0, drop_path=0.1,
                              num_classes_list=[25] * N_SUB, num_div=N_SUB)
    fuse = ensemble.EnsMLP(model="dedeit", num_class=NUM_CLASS, sub_size=D,
                           num_classes_list=[25] * N_SUB, teacher_size=768)
    for s in range(N_SUB):
        multi.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
        ng, hg = synth.shrink_gates(s)
        shrink.mlp_neuron_shrink(multi.backbones[s], ng)
        shrink.attn_head_shrink(multi.backbones[s], hg)
    fuse.load_state_dict(synth.ensmlp_state_dict(N_SUB, num_class=NUM_CLASS))
    return multi.to(dev).eval(), fuse.to(dev).eval()


def check_cct(world, rank, dev):
    """BASELINE config C4 shape at a small batch: 4-way decct_7_3x1 + EnsembleCCT sharded over the
    ranks (MultiCCT.forward_slab -> all-gather -> EnsembleCCT.forward_gathered) against the same
    ensemble on this GPU alone."""
    from devit_b200 import cct
    n_sub, Bc = 4, max(8, B)
    multi = cct.MultiCCT('decct_7_3x1', num_classes_list=[25] * n_sub, num_sub_models=n_sub,
                         input_size=32)
    for s in range(n_sub):
        multi.models[s].load_state_dict(synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True))
    fuse = cct.EnsembleCCT(sub_size=256, teacher_size=None, num_sub_models=n_sub, num_classes=100)
    fuse.load_state_dict(synth.ensemble_cct_state_dict(n_sub, 256, None, 100))
    multi, fuse = multi.to(dev).eval(), fuse.to(dev).eval()
    x = synth.cifar_images(Bc)
    ok = True
    for precision in ("bf16", "fp32"):
        multi.set_precision(precision)
        fuse.set_precision(precision)
        plan = parallel.shard_plan(world, rank, n_sub, Bc)
        group = parallel.make_groups(plan)
        xs = x[plan.batch_lo:plan.batch_hi].to(dev)
        sharded = parallel.ShardedEnsemble(multi, fuse, plan, group)(xs)
        single = fuse(multi(xs))
        flag = torch.tensor([int(torch.equal(sharded, single)),
                             int(torch.equal(sharded.argmax(-1), single.argmax(-1)))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[check_multigpu] CCT world={world} {precision}: bit-identical={bool(flag[0])} "
                  f"argmax-equal={bool(flag[1])}", flush=True)
        ok = ok and bool(flag[0])
    return ok


def main():
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    multi, fuse = build(dev)
    x = synth.images(B)
    ok = True
    for precision in ("bf16", "fp32"):
        multi.set_precision(precision)
        fuse.set_precision(precision)
        plan = parallel.shard_plan(world, rank, N_SUB, B)
        group = parallel.make_groups(plan)
        xs = x[plan.batch_lo:plan.batch_hi].to(dev)
        sharded = parallel.ShardedEnsemble(multi, fuse, plan, group)(xs)
        single_plan = parallel.shard_plan(1, 0, N_SUB, B)
        single = parallel.ShardedEnsemble(multi, fuse, single_plan, None)(xs)
        same = torch.equal(sharded, single)
        # host batch staged as 1/G per rank + NVLink all-gather must rebuild xs exactly
        stage_group = parallel.make_groups(plan)
        ens = parallel.ShardedEnsemble(multi, fuse, plan, group, stage_group)
        host = x[plan.batch_lo:plan.batch_hi].contiguous().pin_memory()
        staged = ens.stage_batch(host, torch.empty_like(xs))
        torch.cuda.synchronize()
        same = same and torch.equal(staged, xs)
        amax = torch.equal(sharded.argmax(-1), single.argmax(-1))
        err = ((sharded.float() - single.float()).abs().max() / single.float().abs().max()).item()
        flag = torch.tensor([int(same), int(amax)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[check_multigpu] world={world} {precision}: bit-identical={bool(flag[0])} "
                  f"argmax-equal={bool(flag[1])} rank0 rel diff={err:.2e}", flush=True)
        ok = ok and bool(flag[1]) and err < 1e-6
    ok = check_cct(world, rank, dev) and ok
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
