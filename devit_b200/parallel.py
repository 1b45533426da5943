"""Multi-GPU execution of the ensemble: one process per GPU (torchrun), sub-models sharded
across ranks, ONE exchange step -- an all-gather of each rank's LayerNormed cls/dist block into
the slab the fusion head reads (new in this build, the reference runs all
sub-models sequentially on one device, models/ensemble_models.py:33).

Partitioning (n_sub sub-models, W ranks): model-parallel groups of G = min(W, n_sub) ranks;
rank r of a group owns sub-models {s : s % G == r}; with W > n_sub the W / G groups are
data-parallel replicas that each take an equal slice of the batch (no cross-group traffic).
After the gather every rank of a group holds the group's logits.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    world: int
    rank: int
    n_sub: int
    batch: int
    group_size: int = 1
    num_groups: int = 1
    group_id: int = 0
    model_rank: int = 0
    subs: list = field(default_factory=list)
    batch_lo: int = 0
    batch_hi: int = 0
    group_ranks: list = field(default_factory=list)

    @property
    def n_local(self) -> int:
        return len(self.subs)

    @property
    def group_batch(self) -> int:
        return self.batch_hi - self.batch_lo

    def gathered_order(self) -> list:
        """Sub-model id of entry j of the gathered slab [G * n_local, ...] (rank-major)."""
        return [i * self.group_size + r for r in range(self.group_size)
                for i in range(self.n_local)]


def shard_plan(world: int, rank: int, n_sub: int, batch: int) -> ShardPlan:
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    g = min(world, n_sub)
    if world % g or n_sub % g:
        raise ValueError(f"world {world} and n_sub {n_sub} must nest (G={g})")
    groups = world // g
    if batch % groups:
        raise ValueError(f"batch {batch} is not divisible by the {groups} data-parallel groups")
    gid, mr = rank // g, rank % g
    per = batch // groups
    return ShardPlan(world=world, rank=rank, n_sub=n_sub, batch=batch, group_size=g,
                     num_groups=groups, group_id=gid, model_rank=mr,
                     subs=[s for s in range(n_sub) if s % g == mr],
                     batch_lo=gid * per, batch_hi=(gid + 1) * per,
                     group_ranks=list(range(gid * g, (gid + 1) * g)))


def make_groups(plan: ShardPlan):
    """Creates every model-parallel process group (collective over all ranks) and returns the
    one this rank belongs to (None when a group is a single rank).  Call it twice to get a
    second, independent communicator over the same ranks (ShardedEnsemble.stage_group)."""
    mine = None
    if plan.group_size == 1:
        return None
    for gid in range(plan.num_groups):
        ranks = list(range(gid * plan.group_size, (gid + 1) * plan.group_size))
        pg = dist.new_group(ranks=ranks)
        if gid == plan.group_id:
            mine = pg
    return mine


def gather_blocks(local: torch.Tensor, plan: ShardPlan, group) -> torch.Tensor:
    """all-gather of equally shaped per-rank blocks -> [G, *local.shape] (rank-major)."""
    if plan.group_size == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    out = torch.empty((plan.group_size * local.shape[0],) + tuple(local.shape[1:]),
                      dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out.view((plan.group_size,) + tuple(local.shape))


def stage_slice(plan: ShardPlan, rows: int):
    """(lo, hi) rows of a group batch of `rows` images that model rank `plan.model_rank` uploads
    in `ShardedEnsemble.stage_batch`; None when the batch does not split evenly."""
    g = plan.group_size
    if g <= 1 or rows % g:
        return None
    per = rows // g
    return plan.model_rank * per, (plan.model_rank + 1) * per


class ShardedEnsemble:
    """MultiViT + EnsMLP across the ranks of a ShardPlan.  `multi` holds (at least) this rank's
    sub-models; `fuse` is replicated on every rank."""

    def __init__(self, multi, fuse, plan: ShardPlan, group=None, stage_group=None):
        self.multi, self.fuse, self.plan, self.group = multi, fuse, plan, group
        self.order = plan.gathered_order()
        # a second communicator over the same ranks for the input exchange, so that the upload
        # of batch i+1 never queues in front of the feature all-gather of batch i
        self.stage_group = stage_group
        self._stage_buf = None

    @torch.no_grad()
    def stage_batch(self, host_batch: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Brings the group's batch from (pinned) HOST memory into `out` on this rank's GPU.
        Every sub-model consumes the same images (models/ensemble_models.py:33), so instead of
        all G model ranks pulling the whole batch over PCIe, rank r uploads rows
        [r B/G, (r+1) B/G) and one all-gather over NVLink assembles the batch on every rank:
        1/G of the host->device bytes per rank.  Runs on the current stream; falls back to a plain
        full copy for a single rank or an uneven split.  Returns `out`."""
        sl = stage_slice(self.plan, host_batch.shape[0])
        if sl is None or self.stage_group is None:
            out.copy_(host_batch, non_blocking=True)
            return out
        lo, hi = sl
        shape = (hi - lo,) + tuple(host_batch.shape[1:])
        if self._stage_buf is None or self._stage_buf.shape != shape or \
                self._stage_buf.dtype != host_batch.dtype or self._stage_buf.device != out.device:
            self._stage_buf = torch.empty(shape, dtype=host_batch.dtype, device=out.device)
        self._stage_buf.copy_(host_batch[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(out, self._stage_buf, group=self.stage_group)
        return out

    @torch.no_grad()
    def __call__(self, x_group: torch.Tensor) -> torch.Tensor:
        """x_group: this rank's copy of its group's batch slice [Bg, 3, H, W] on the local GPU.
        Returns the group's logits [Bg, num_class]."""
        from . import _lib as L
        from .models import _PREC
        _, op = self.multi.forward_slab(x_group, subs=self.plan.subs)
        g = gather_blocks(op, self.plan, self.group)
        if _PREC[self.multi.precision] == L.DEVIT_BF16:
            slab = g.view((-1,) + tuple(op.shape[1:]))           # [G*n_local, 2, Bg, D]
        else:                                                     # [G, 2, n_local, 2, Bg, D]
            slab = g.transpose(0, 1).reshape((2, -1) + tuple(op.shape[2:])).contiguous()
        return self.fuse.forward_gathered(slab, self.order)


class Replica:
    """A single model (the deit_base teacher, BASELINE config C1) on the ranks of a ShardPlan with
    n_sub = 1: the path does not shard below one model, so N ranks are N data-parallel replicas
    that each take 1/N of the batch -- no collective at all (``stage_batch`` is a plain copy)."""

    def __init__(self, model, plan: ShardPlan):
        self.model, self.plan = model, plan

    @torch.no_grad()
    def stage_batch(self, host_batch: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        out.copy_(host_batch, non_blocking=True)
        return out

    @torch.no_grad()
    def __call__(self, x_group: torch.Tensor) -> torch.Tensor:
        return self.model(x_group)


class BatchPipeline:
    """`depth` copies of one step, each with its own CUDA stream, CUDA graph, workspace and output,
    replayed round-robin so that `depth` BATCHES are in flight on the GPU at once.

    A rank that owns a single sub-model (one sub-model per GPU on 4 or 8 GPUs), or the single
    teacher model, runs one kernel chain: every persistent kernel alternates tensor-bound and
    memory-bound phases and leaves its last round of tiles partly empty (99 pair-tiles on 74
    cluster slots at 128 images).  Consecutive batches of an evaluation loop are independent, so a
    second batch on a second stream fills those holes exactly like a second sub-model does at N = 1
    (measured on one GPU, profiles/r2_time_pipeline.txt: 1 sub-model x 256 images 2.08 -> 1.91 ms per
    batch, x 128 images 1.28 -> 1.00 ms).  Latency per batch grows, throughput is what is reported.

    `steps`: one zero-argument callable per slot returning the slot's output tensor; across ranks
    every slot must use its OWN communicator (collectives of different slots run concurrently)."""

    def __init__(self, steps, capture: bool = True, n_streams: int = 0):
        """n_streams (0 = one per step): slot k runs on stream k % n_streams.  More slots than
        streams = several input / output buffers per in-flight batch (slots that share a stream
        never overlap, so they may share a communicator): an input pipeline can then fill the
        buffers of later slots while the earlier ones compute."""
        cur = torch.cuda.current_stream()
        self.steps = list(steps)
        n_streams = len(self.steps) if n_streams <= 0 else min(n_streams, len(self.steps))
        pool = [torch.cuda.Stream() for _ in range(n_streams)]
        self.streams = [pool[k % n_streams] for k in range(len(self.steps))]
        self.graphs, self.outs = [], []
        for st, fn in zip(self.streams, self.steps):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(2):  # builds this stream's workspaces outside the capture
                    out = fn()
            torch.cuda.synchronize()
            g = None
            if capture:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    out = fn()
                with torch.cuda.stream(st):
                    g.replay()  # a capture does not execute: fill `out` once
            self.graphs.append(g)
            self.outs.append(out)
        torch.cuda.synchronize()
        self._next = 0

    @property
    def depth(self) -> int:
        return len(self.steps)

    def reset(self):
        """Next launch() goes to slot 0 again (callers that pair slots with input buffers)."""
        self._next = 0

    def fork(self):
        """Order the slots' streams after the work already queued on the current stream."""
        cur = torch.cuda.current_stream()
        for st in set(self.streams):
            st.wait_stream(cur)

    def launch(self) -> int:
        """Enqueue the next step on its slot's stream; returns the slot (its result is in
        `outs[slot]` once that stream has run it; the slot's previous result is overwritten)."""
        k = self._next % len(self.steps)
        self._next += 1
        with torch.cuda.stream(self.streams[k]):
            if self.graphs[k] is not None:
                self.graphs[k].replay()
            else:
                self.outs[k] = self.steps[k]()
        return k

    def join(self):
        """Make the current stream wait for everything the slots have been given."""
        cur = torch.cuda.current_stream()
        for st in set(self.streams):
            cur.wait_stream(st)
