"""N>1 host logic on CPU: shard plans and the gathered-slab ordering, with a real world_size-2
(and 4) gloo all-gather between processes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from devit_b200 import parallel


def test_shard_plans():
    p = parallel.shard_plan(1, 0, 4, 256)
    assert p.subs == [0, 1, 2, 3] and p.group_batch == 256 and p.gathered_order() == [0, 1, 2, 3]
    p = parallel.shard_plan(2, 1, 4, 256)
    assert p.subs == [1, 3] and p.group_size == 2 and p.gathered_order() == [0, 2, 1, 3]
    p = parallel.shard_plan(4, 2, 4, 256)
    assert p.subs == [2] and p.gathered_order() == [0, 1, 2, 3] and p.group_batch == 256
    p = parallel.shard_plan(8, 6, 4, 256)
    assert p.subs == [2] and p.num_groups == 2 and p.group_id == 1
    assert (p.batch_lo, p.batch_hi) == (128, 256) and p.group_ranks == [4, 5, 6, 7]
    p = parallel.shard_plan(8, 3, 8, 1024)
    assert p.subs == [3] and p.num_groups == 1 and p.group_batch == 1024
    with pytest.raises(ValueError):
        parallel.shard_plan(3, 0, 4, 256)
    with pytest.raises(ValueError):
        parallel.shard_plan(8, 0, 4, 255)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_sub, batch, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        plan = parallel.shard_plan(world, rank, n_sub, batch)
        group = parallel.make_groups(plan)
        D, Bg = 8, plan.group_batch
        gen = torch.Generator().manual_seed(7)
        feats = torch.randn(n_sub, 2, batch, D, generator=gen)      # every sub-model, full batch
        w = torch.randn(5, n_sub * D, generator=gen)
        local = torch.stack([feats[s][:, plan.batch_lo:plan.batch_hi] for s in plan.subs])
        g = parallel.gather_blocks(local, plan, group)
        slab = g.reshape((-1,) + tuple(local.shape[1:]))             # [G*n_local, 2, Bg, D]
        order = plan.gathered_order()
        # fusion as the K-segment sum the GEMM performs
        got = sum(slab[j, 0] @ w[:, order[j] * D:(order[j] + 1) * D].t() for j in range(n_sub))
        # reference formulation: stack(list, 1).view(B, -1) then Linear
        cls_list = [feats[s, 0, plan.batch_lo:plan.batch_hi] for s in range(n_sub)]
        want = F.linear(torch.stack(cls_list, 1).view(Bg, -1), w)
        ok = torch.allclose(got, want, atol=1e-5)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_sub', [(2, 4), (4, 4)])
def test_gloo_gather_matches_stack_order(world, n_sub):
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sub, 8, ret))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


def _meters_worker(rank, world, port, ret):
    """Each rank evaluates its own shard of the batches (the reference's DistributedSampler
    split); the all-reduced meters must equal the single-process result over all batches
    (utils/dist_utils.py:35-46: count and total are summed over ranks)."""
    from devit_b200 import engine, synth
    from oracle import devit_oracle as O
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batches = synth.eval_batches((8, 8, 5, 3), 100)
        meters = engine.EvalMeters(torch.device('cpu'))
        for lg, tg in batches[rank::world]:           # host stand-in for devit_eval_tail
            loss, c1, ck = O.eval_tail(lg, tg)
            meters.acc += torch.tensor([loss, 1, c1, ck, lg.shape[0]], dtype=torch.float64)
        meters.synchronize_between_processes()
        got = meters.result()
        want = O.eval_epoch(batches)
        ok = all(abs(got[k] - want[k]) <= 1e-9 * max(1.0, abs(want[k])) for k in want)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def test_gloo_eval_meters_reduce_like_metric_logger():
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_meters_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


def _stage_worker(rank, world, port, n_sub, batch, ret):
    """stage_batch: every rank uploads 1/G of its group's batch and the all-gather over the
    second communicator rebuilds the whole group batch on every rank (uint8 and fp32)."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        plan = parallel.shard_plan(world, rank, n_sub, batch)
        group = parallel.make_groups(plan)
        stage_group = parallel.make_groups(plan)
        ens = parallel.ShardedEnsemble(None, None, plan, group, stage_group)
        gen = torch.Generator().manual_seed(11)
        full = torch.randn(batch, 3, 4, 4, generator=gen)
        full8 = torch.randint(0, 256, (batch, 4, 4, 3), generator=gen, dtype=torch.uint8)
        ok = True
        for src in (full, full8):
            host = src[plan.batch_lo:plan.batch_hi].contiguous()
            out = torch.zeros_like(host)
            ens.stage_batch(host, out)
            ok = ok and torch.equal(out, host)
        # an uneven split falls back to the plain copy
        odd = full[:plan.group_size + 1].contiguous()
        ok = ok and parallel.stage_slice(plan, odd.shape[0]) is None
        ok = ok and torch.equal(ens.stage_batch(odd, torch.zeros_like(odd)), odd)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_sub,batch', [(2, 4, 8), (4, 2, 8)])
def test_gloo_stage_batch_rebuilds_the_group_batch(world, n_sub, batch):
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stage_worker, args=(r, world, port, n_sub, batch, ret))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1


def test_stage_slice():
    p = parallel.shard_plan(4, 2, 4, 256)
    assert parallel.stage_slice(p, 256) == (128, 192) and parallel.stage_slice(p, 255) is None
    assert parallel.stage_slice(parallel.shard_plan(1, 0, 4, 256), 256) is None
    p = parallel.shard_plan(8, 5, 4, 256)          # 2 data-parallel groups of 4 model ranks
    assert p.group_batch == 128 and parallel.stage_slice(p, 128) == (32, 64)
