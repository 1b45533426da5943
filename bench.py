#!/usr/bin/env python
"""Benchmark of the B200 hot path (BASELINE.json): images/sec of the decomposed ensembles.

  python bench.py --gpus 1 --steps K --warmup W [--config headline|c1|c2|c3|c4] [--dense]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...   the reference's own CPU implementation (host cores)

Configs (BASELINE.json `configs`; `headline` is the one `metric` is quoted on):
  headline  4-way DeDeiT ensemble, shrink_ratio-0.3 head+neuron gates, 100 classes, bs 256
            (the line also carries a `dense` sub-object: the same step with all-ones gates)
  c1        deit_base_distilled_patch16_224 teacher, bs 256 (N > 1: data-parallel replicas)
  c2        the headline ensemble at bs 512
  c3        8-way DeDeiT, dense, ImageNet-1K fusion head (1000 classes), bs 1024
  c4        4-way decct_7_3x1 CCT ensemble (EnsembleCCT, 100 classes), 32x32, bs 2048

One step = one forward of the whole ensemble (backbones + fusion head) over the global batch.
`value`: inputs resident in HBM, CUDA-graph replay of the step.  `e2e`: the public module call
with the batch copied host->device from pinned memory and the logits copied back, inside the
timed region.  Global batch is fixed for every N (strong scaling): sub-models are sharded over
the ranks (one exchange: an all-gather of the feature slabs), extra ranks form data-parallel
groups.  Inputs + activations exceed the 126 MB L2, so no explicit flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D, HEADS, HIDDEN, TOKENS, PATCHES, DEPTH = 384, 6, 1536, 198, 196, 12
HEADLINE_METRIC = "images/sec, 4-way DeDeiT ensemble @224^2 bs256"

CONFIGS = {
    "headline": dict(family="dedeit", n_sub=4, classes=100, batch=256, shrunk=True,
                     metric=HEADLINE_METRIC),
    "c1": dict(family="teacher", n_sub=1, classes=100, batch=256, shrunk=False,
               metric="images/sec, deit_base_distilled_patch16_224 teacher @224^2 bs256"),
    "c2": dict(family="dedeit", n_sub=4, classes=100, batch=512, shrunk=True,
               metric="images/sec, 4-way DeDeiT ensemble @224^2 bs512"),
    "c3": dict(family="dedeit", n_sub=8, classes=1000, batch=1024, shrunk=False,
               metric="images/sec, 8-way DeDeiT ensemble (1000 classes) @224^2 bs1024"),
    "c4": dict(family="cct", n_sub=4, classes=100, batch=2048, shrunk=False,
               metric="images/sec, 4-way decct_7_3x1 CCT ensemble @32^2 bs2048"),
}


# ------------------------------------------------------------------------------- flop model
def vit_flops(dim, heads_l, hidden_l, tokens=TOKENS, patches=PATCHES):
    """Algorithmic FLOPs (2*MAC) per image of one ViT sub-model with the KEPT head /
    neuron counts per layer -> dict(gemm, attn, tail) where `tail` = proj + fc1 + fc2 (the work of
    the fused projection + MLP kernel)."""
    gemm = 2 * patches * 768 * dim
    attn = tail = 0
    for h, f in zip(heads_l, hidden_l):
        t = 2 * tokens * 64 * h * dim + 4 * tokens * dim * f
        gemm += 2 * tokens * dim * 3 * 64 * h + t
        tail += t
        attn += 4 * h * tokens * tokens * 64
    return dict(gemm=gemm, attn=attn, tail=tail)


def cct_flops(n_conv=1, tokens=256, dim=256, hidden=512, depth=7, heads=4):
    side, cin, conv = 32, 3, 0
    for i in range(n_conv):
        cout = dim if i == n_conv - 1 else 64
        conv += 2 * side * side * 9 * cin * cout
        cin, side = cout, side // 2
    tail = depth * (2 * tokens * dim * dim + 4 * tokens * dim * hidden)
    gemm = conv + depth * 2 * tokens * dim * 3 * dim + tail
    attn = depth * 4 * heads * tokens * tokens * 64
    return dict(gemm=gemm, attn=attn, tail=tail)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"bf16_burst": j["bf16_tflops"], "bf16_sustained": j["bf16_tflops_sustained"],
                "hbm": j["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples while the timed regions run."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid, self.proc, self.lines = uuid, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm_load = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None,
                "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- workloads
class Workload:
    """Everything bench.py needs to know about one config: synthetic weights / gates / inputs
    (devit_b200/synth.py seeds), the device modules, the CPU arm and the FLOP model."""

    def __init__(self, name, dense=False):
        self.name = name
        c = dict(CONFIGS[name])
        if dense:
            c["shrunk"] = False
        self.__dict__.update(c)
        self.kept = None

    # ---- descriptions
    def describe(self):
        if self.family == "teacher":
            return "deit_base_distilled_patch16_224 teacher (D=768, 12 heads, depth 12), 100 classes"
        if self.family == "cct":
            return (f"{self.n_sub}-way decct_7_3x1 CCT ensemble (D=256, 7 layers, 256 tokens), "
                    f"EnsembleCCT fusion head {self.classes} classes")
        gates = "shrink_ratio-0.3 head+neuron gates" if self.shrunk else "dense gates"
        return (f"{self.n_sub}-way DeDeiT ensemble (dedeit D=384 depth 12), {gates}, EnsMLP fusion "
                f"head {self.classes} classes")

    def config(self, world, plan):
        side = 32 if self.family == "cct" else 224
        if self.family == "teacher":
            par = "1 GPU" if world == 1 else f"{world} data-parallel replicas (batch split, no collective)"
        elif world == 1:
            par = (f"1 GPU: {self.n_sub} sub-models as concurrent kernel chains on "
                   f"{os.environ.get('DEVIT_SUB_STREAMS', '4')} streams")
        else:
            par = (f"{plan.group_size}-way sub-model sharding x {plan.num_groups} data-parallel "
                   f"group(s), all-gather of the per-rank feature slabs")
        in_mb = self.batch * 3 * side * side * 4 / 1e6
        return {"workload": f"{self.describe()}, {side}x{side}, global batch {self.batch}",
                "name": self.name, "global_batch": self.batch, "sub_models": self.n_sub,
                "parallelism": par,
                "l2": f"inputs ({in_mb:.0f} MB) + activations exceed the 126 MB L2; no flush"}

    # ---- inputs
    def images(self, n, seed=None):
        from devit_b200 import synth
        if self.family == "cct":
            return synth.cifar_images(n) if seed is None else synth.cifar_images(n, seed)
        return synth.images(n) if seed is None else synth.images(n, seed)

    # ---- device modules -> (multi, fuse) or (model, None)
    def build_device(self, dev, precision):
        from devit_b200 import cct, ensemble, shrink, synth
        from devit_b200.registry import create_model
        if self.family == "teacher":
            m = create_model("deit_base_distilled_patch16_224", num_classes=self.classes)
            m.load_state_dict(synth.teacher_state_dict(self.classes))
            self.kept = [([12] * DEPTH, [3072] * DEPTH)]
            return m.to(dev).eval().set_precision(precision), None
        n = self.n_sub
        if self.family == "cct":
            multi = cct.MultiCCT("decct_7_3x1", num_classes_list=[self.classes // n] * n,
                                 num_sub_models=n, input_size=32)
            for s in range(n):
                multi.models[s].load_state_dict(
                    synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True))
            fuse = cct.EnsembleCCT(sub_size=256, teacher_size=None, num_sub_models=n,
                                   num_classes=self.classes)
            fuse.load_state_dict(synth.ensemble_cct_state_dict(n, 256, None, self.classes))
            return (multi.to(dev).eval().set_precision(precision),
                    fuse.to(dev).eval().set_precision(precision))
        multi = ensemble.MultiViT(model="dedeit", drop=0, drop_path=0.1,
                                  num_classes_list=[self.classes // n] * n, num_div=n)
        fuse = ensemble.EnsMLP(model="dedeit", num_class=self.classes, sub_size=D,
                               num_classes_list=[self.classes // n] * n, teacher_size=768)
        self.kept = []
        for s in range(n):
            multi.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
            if self.shrunk:
                ng, hg = synth.shrink_gates(s)
                shrink.mlp_neuron_shrink(multi.backbones[s], ng)
                shrink.attn_head_shrink(multi.backbones[s], hg)
                self.kept.append(([int(g.sum()) for g in hg], [int(g.sum()) for g in ng]))
            else:
                self.kept.append(([HEADS] * DEPTH, [HIDDEN] * DEPTH))
        fuse.load_state_dict(synth.ensmlp_state_dict(n, num_class=self.classes))
        return (multi.to(dev).eval().set_precision(precision),
                fuse.to(dev).eval().set_precision(precision))

    # ---- FLOPs per image of sub-model s, and of the fusion head
    def sub_flops(self, s):
        if self.family == "teacher":
            return vit_flops(768, *self.kept[0])
        if self.family == "cct":
            return cct_flops()
        return vit_flops(D, *self.kept[s])

    def fusion_flops(self):
        if self.family == "teacher":
            return 2 * 2 * 768 * self.classes
        if self.family == "cct":
            return 2 * self.n_sub * 256 * self.classes
        return 2 * 2 * (self.n_sub * D * 768 + 768 * self.classes)

    def total_flops_per_image(self):
        return sum(self.sub_flops(s)["gemm"] + self.sub_flops(s)["attn"]
                   for s in range(self.n_sub)) + self.fusion_flops()

    # ---- CPU arm: callable x -> logits, and what it is
    def cpu_runner(self):
        """The reference's own modules (oracle/ref_runner.py over /root/reference or the
        oracle/_ref copy) when available, else the oracle port."""
        import torch
        from devit_b200 import synth
        from oracle import ref_runner
        if self.family in ("dedeit", "teacher") and ref_runner.available():
            try:
                if self.family == "teacher":
                    return ref_runner.teacher(self.classes), "reference", \
                        "reference models/deit_vit.py modules (fp32 PyTorch CPU)"
                return ref_runner.ensemble(self.n_sub, self.classes, self.shrunk), "reference", \
                    ("reference models/ensemble_models.py MultiViT + EnsMLP over models/de_vit.py "
                     "(masked-dense fp32 PyTorch CPU)")
            except Exception as e:  # noqa: BLE001
                print(f"[bench] reference modules unavailable ({type(e).__name__}: {e}); "
                      f"timing the oracle port", file=sys.stderr)
        if self.family == "cct":
            from oracle import cct_oracle as CO
            sds = [synth.cct_state_dict(s, n_conv=1, tokens=256, backbone=True)
                   for s in range(self.n_sub)]
            esd = synth.ensemble_cct_state_dict(self.n_sub, 256, None, self.classes)

            def run(x):
                with torch.no_grad():
                    return CO.ensemble_logits(sds, esd, x, 1)
            return run, "port", "oracle port of the reference CCT ensemble (fp32 PyTorch CPU)"
        from oracle import devit_oracle as O
        if self.family == "teacher":
            sd = synth.teacher_state_dict(self.classes)

            def run(x):
                with torch.no_grad():
                    return O.forward_logits(sd, x, num_heads=12)
            return run, "port", "oracle port of the reference teacher (fp32 PyTorch CPU)"
        sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(self.n_sub)]
        esd = synth.ensmlp_state_dict(self.n_sub, num_class=self.classes)
        gates = [synth.shrink_gates(s) for s in range(self.n_sub)] if self.shrunk else None

        def run(x):
            with torch.no_grad():
                return O.ensemble_logits(sds, esd, x, gates)[0]
        return run, "port", "oracle port of the reference (masked-dense fp32 PyTorch CPU)"


def cpu_time(wl, images_per_step, steps, warmup):
    """Times the CPU arm on all host cores -> (img/s, ms/step, cores, kind, what, logits, x)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind, what = wl.cpu_runner()
    x = wl.images(wl.batch)[:images_per_step].contiguous()
    out = None
    for _ in range(warmup):
        out = run(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run(x)
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt / steps * 1e3, cores, kind, what, out, x


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the same workload on the box's
    host cores.  Each step processes a bounded sample of the global batch, sized from a probe so
    that the K + W steps end within a few minutes; the sample is stated in the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = Workload(args.config, args.dense)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind, what = wl.cpu_runner()
    probe_n = min(16, wl.batch)
    xp = wl.images(wl.batch)[:probe_n].contiguous()
    run(xp)
    t0 = time.perf_counter()
    run(xp)
    rate = probe_n / (time.perf_counter() - t0)
    warm = min(args.warmup, 2)
    budget_s = float(os.environ.get("DEVIT_REF_BUDGET_S", "150"))
    sample = int(rate * budget_s / (args.steps + warm))
    sample = max(probe_n, min(wl.batch, sample // 16 * 16))
    x = wl.images(wl.batch)[:sample].contiguous()
    for _ in range(warm):
        run(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(x)
    dt = time.perf_counter() - t0
    ips, ms = sample * args.steps / dt, dt / args.steps * 1e3
    cfg = wl.config(1, None)
    cfg["reference_step"] = (f"{sample} of the {wl.batch} images of the global batch per timed step "
                             f"(bounded CPU sample; throughput in images/sec is batch-size "
                             f"independent beyond ~16 images on the CPU path)")
    line = {
        "impl": "reference", "metric": wl.metric, "value": ips, "unit": "images/sec",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": cores, "kind": kind,
                         "sample": f"{sample} images per step x {args.steps} steps, {what}"},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="devit", choices=["devit", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--dense", action="store_true", help="all-ones gates instead of shrunk")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--eager-multi", action="store_true",
                    help="N > 1: do not capture the step (NCCL all-gather included) in a CUDA graph")
    ap.add_argument("--pipeline-depth", type=int, default=0,
                    help="batches in flight (0 = auto: 2 when the rank runs a single kernel chain)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense-arm", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    # a hung collective / kernel must not eat the whole GPU lease: after the watchdog period every
    # thread's Python stack goes to stderr and the process exits
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DEVIT_BENCH_WATCHDOG_S", "480")),
                                      exit=True)

    import torch
    import torch.distributed as dist
    from devit_b200 import _lib as L
    from devit_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (devit_b200 has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().devit_device_check())
    lib = L.load()

    wl = Workload(args.config, args.dense)
    BATCH = wl.batch
    plan = parallel.shard_plan(world, rank, wl.n_sub, BATCH)
    group = parallel.make_groups(plan) if world > 1 else None
    multi, fuse = wl.build_device(dev, args.precision)
    # second communicator over the same ranks: the e2e arms upload 1/G of the batch per rank and
    # all-gather it over NVLink (ShardedEnsemble.stage_batch)
    stage_group = parallel.make_groups(plan) if world > 1 else None
    if fuse is None:
        ens = parallel.Replica(multi, plan)
    else:
        ens = parallel.ShardedEnsemble(multi, fuse, plan, group, stage_group)

    Bg = plan.group_batch
    x_host = wl.images(BATCH)[plan.batch_lo:plan.batch_hi].contiguous().pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    # ---- warm-up (also builds packs / workspaces) and launch count of one step
    logits = ens(x_dev)
    torch.cuda.synchronize()
    c0 = lib.devit_launch_count()
    logits = ens(x_dev)
    torch.cuda.synchronize()
    launches_per_step = lib.devit_launch_count() - c0
    for _ in range(args.warmup):
        ens(x_dev)
    torch.cuda.synchronize()

    # ---- parity of the sharded step, checked BEFORE anything is timed: every rank also runs its
    #      group's whole ensemble locally (all sub-models on this GPU, no collective); the
    #      sharded logits must be bit-identical (the fusion head sums the K-segments in sub-model
    #      order for every world size).
    parity = {}
    if world > 1 and fuse is not None:
        local = parallel.ShardedEnsemble(multi, fuse, parallel.shard_plan(1, 0, wl.n_sub, Bg))
        ref_local = local(x_dev)
        torch.cuda.synchronize()
        same = torch.equal(ref_local, logits)
        amax = torch.equal(ref_local.argmax(-1), logits.argmax(-1))
        parity["sharded_vs_single_rank"] = {
            "bit_identical": all_ranks(same), "argmax_equal": all_ranks(amax),
            "images_per_rank": Bg, "ranks": world}
        if not parity["sharded_vs_single_rank"]["argmax_equal"]:
            raise SystemExit("[bench] sharded logits disagree with the single-rank ensemble")
        del local, ref_local

    # ---- CUDA graph of the whole step (N > 1: the NCCL all-gather is captured with it)
    graph, g_out = None, None
    if not args.no_graph and (world == 1 or not args.eager_multi):
        ok = True
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    ens(x_dev)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_out = ens(x_dev)
            graph.replay()
            torch.cuda.synchronize()
            if not torch.equal(g_out, logits):
                raise RuntimeError("graph replay differs from the eager result")
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: CUDA graph capture unavailable ({type(e).__name__}: {e}); "
                  f"timing eager launches", file=sys.stderr)
            ok = False
        if not all_ranks(ok):  # every rank must take the same path (collectives inside)
            graph = None
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            ens(x_dev)

    for _ in range(args.warmup):
        step()
    launch_mode = "cuda_graph" if graph is not None else "eager"

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n

    # ---- value: device-resident inputs
    ms_step = timed(step, args.steps)
    value = BATCH / (ms_step / 1e3)
    one_in_flight = {"value": value, "ms_per_step": ms_step, "launch_mode": launch_mode}
    in_flight = None

    # ---- a rank with ONE kernel chain (one sub-model per GPU at N >= 4, the single teacher
    #      model) keeps a second, independent batch in flight on a second stream: consecutive
    #      batches of an evaluation loop do not depend on each other, and the second chain fills
    #      the memory-bound phases and partly empty last tile rounds of the first exactly like a
    #      second sub-model does at N = 1 (parallel.BatchPipeline).  Each slot has its own stream,
    #      CUDA graph, workspace, output and NCCL communicator.
    n_chains = 1 if fuse is None else len(plan.subs)
    depth = args.pipeline_depth if args.pipeline_depth > 0 else (2 if n_chains == 1 else 1)
    ens_slots = [ens]
    pipe = None
    if depth > 1 and graph is not None:
        for _ in range(depth - 1):
            if fuse is None:
                ens_slots.append(parallel.Replica(multi, plan))
            else:
                ens_slots.append(parallel.ShardedEnsemble(
                    multi, fuse, plan, parallel.make_groups(plan) if world > 1 else None,
                    parallel.make_groups(plan) if world > 1 else None))
        pipe = parallel.BatchPipeline([(lambda e=e: e(x_dev)) for e in ens_slots])
        if not all_ranks(all(torch.equal(o, logits) for o in pipe.outs)):
            raise SystemExit("[bench] a pipelined slot disagrees with the plain step")

        def run_pipe(n):
            pipe.reset()
            pipe.fork()
            for _ in range(n):
                pipe.launch()
            pipe.join()

        run_pipe(args.warmup)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        run_pipe(args.steps)
        e1.record()
        barrier()
        ms_pipe = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        in_flight = {"value": BATCH / (ms_pipe / 1e3), "ms_per_step": ms_pipe, "batches": depth}
        if ms_pipe < ms_step:  # keep whichever arm is faster (the teacher's big GEMMs fill the
            ms_step = ms_pipe  # chip on their own: 2 in flight measured slower there)
            value = BATCH / (ms_step / 1e3)
            launch_mode = (f"cuda_graph, {depth} batches in flight (one stream + graph + "
                           f"communicator each)")
        else:
            pipe = None

    # ---- e2e: public API with host buffers; H2D of the step's batch + D2H of its logits inside
    #      the timed region, input copies double-buffered on a copy stream
    copy_stream = torch.cuda.Stream()
    # input buffers: two when the steps run one after the other (copy i+1 under compute i); four
    # when two batches are in flight, so that the uploads can run ahead of both computing slots
    nbuf = 4 if pipe is not None else 2
    bufs = [torch.empty_like(x_dev) for _ in range(nbuf)]
    ready = [torch.cuda.Event() for _ in range(nbuf)]
    freed = [torch.cuda.Event() for _ in range(nbuf)]
    out_host = torch.empty(Bg, wl.classes).pin_memory()
    main_stream = torch.cuda.current_stream()

    # the same public call, captured once per input buffer when CUDA graphs are in use (a user
    # serving fixed-shape batches would do the same); eager launches otherwise
    e2e_graphs = None
    e2e_pipe = None
    if pipe is not None:
        # one slot (CUDA graph) per input buffer on `depth` streams: buffer b is consumed on stream
        # b % depth with that stream's communicator, so the compute of batch i overlaps that of
        # batch i + 1 as in `value`, and the uploads fill the other buffers meanwhile
        for b in range(nbuf):
            bufs[b].copy_(x_dev)
        torch.cuda.synchronize()
        e2e_pipe = parallel.BatchPipeline([(lambda e=ens_slots[b % depth], b=b: e(bufs[b]))
                                           for b in range(nbuf)], n_streams=depth)
    elif graph is not None:
        ok = True
        try:
            e2e_graphs = []
            for b in range(2):
                bufs[b].copy_(x_dev)
                torch.cuda.synchronize()
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb):
                    ob = ens(bufs[b])
                e2e_graphs.append((gb, ob))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: e2e graph capture unavailable ({type(e).__name__}: {e})",
                  file=sys.stderr)
            ok = False
        if not all_ranks(ok):
            e2e_graphs = None
            torch.cuda.synchronize()

    use_stage = world > 1 and fuse is not None and parallel.stage_slice(plan, Bg) is not None

    out_hosts = [out_host] + [torch.empty_like(out_host).pin_memory() for _ in range(nbuf - 1)]

    def e2e_loop(n):
        if e2e_pipe is not None:
            e2e_pipe.reset()
            e2e_pipe.fork()
        for i in range(n + 1):
            if i < n:  # prefetch batch i
                b = i % nbuf
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[b])
                    if use_stage:
                        ens.stage_batch(x_host, bufs[b])
                    else:
                        bufs[b].copy_(x_host, non_blocking=True)
                    ready[b].record(copy_stream)
            if i > 0 and e2e_pipe is not None:  # compute batch i-1 on slot b's stream
                b = (i - 1) % nbuf
                st = e2e_pipe.streams[b]
                st.wait_event(ready[b])
                e2e_pipe.launch()
                freed[b].record(st)
                with torch.cuda.stream(st):
                    out_hosts[b].copy_(e2e_pipe.outs[b], non_blocking=True)
                continue
            if i > 0:  # compute batch i-1
                b = (i - 1) % nbuf
                main_stream.wait_event(ready[b])
                if e2e_graphs is not None:
                    e2e_graphs[b][0].replay()
                    out = e2e_graphs[b][1]
                else:
                    out = ens(bufs[b])
                freed[b].record(main_stream)
                out_host.copy_(out, non_blocking=True)
        if e2e_pipe is not None:
            e2e_pipe.join()
        main_stream.synchronize()

    # multi-GPU: every model rank of a group needs the same images; rank r uploads rows
    # [r Bg/G, (r+1) Bg/G) and an NVLink all-gather assembles the batch (1/G of the PCIe bytes per
    # rank).  Checked against the device-resident result before it is timed; any disagreement
    # (on any rank) falls back to every rank copying the whole batch.
    for b in range(nbuf):
        freed[b].record(main_stream)
    e2e_loop(nbuf + 1)
    if use_stage:
        if not all_ranks(torch.equal(out_host.to(dev), logits)):
            if rank == 0:
                print("[bench] staged input gave different logits; timing full per-rank copies",
                      file=sys.stderr)
            use_stage = False
            for b in range(nbuf):
                freed[b].record(main_stream)
            e2e_loop(nbuf + 1)
    e2e_ok = all_ranks(torch.equal(out_host.to(dev), logits))
    parity["e2e_equals_resident"] = e2e_ok
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = BATCH / (e2e_ms / 1e3)
    # bytes that cross PCIe per step over all ranks: the whole group batch per rank, or 1/G of it
    h2d = x_host.numel() * 4 * world // (plan.group_size if use_stage else 1)
    d2h = out_host.numel() * 4 * world

    # ---- e2e_u8 (DeDeiT configs): the same public call fed DECODED uint8 batches (SURVEY.md
    #      8f-3): ToTensor + Normalize run inside the patch extraction on the device, so a step's
    #      H2D copy is a quarter of the fp32 one, and the step's result is the device-side eval tail
    #      (loss, correct@1, correct@5: 12 bytes D2H) instead of the logits.
    e2e_u8 = None
    if wl.family == "dedeit":
        try:
            from devit_b200 import synth
            u8_host = synth.images_u8(BATCH)[plan.batch_lo:plan.batch_hi].contiguous().pin_memory()
            tgt = torch.randint(0, wl.classes, (Bg,),
                                generator=torch.Generator().manual_seed(7)).to(dev)
            bufs8 = [torch.empty(u8_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
            tail_host = torch.empty(3).pin_memory()

            def tail_step(xb):
                return L.eval_tail(ens(xb), tgt)

            for b in range(2):
                bufs8[b].copy_(u8_host)
            for _ in range(2):
                tail_step(bufs8[0])
            torch.cuda.synchronize()
            u8_graphs = None
            if graph is not None:
                u8_graphs = []
                for b in range(2):
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb):
                        ob = tail_step(bufs8[b])
                    u8_graphs.append((gb, ob))
                torch.cuda.synchronize()

            def u8_loop(n):
                for i in range(n + 1):
                    if i < n:
                        b = i & 1
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(freed[b])
                            if use_stage:
                                ens.stage_batch(u8_host, bufs8[b])
                            else:
                                bufs8[b].copy_(u8_host, non_blocking=True)
                            ready[b].record(copy_stream)
                    if i > 0:
                        b = (i - 1) & 1
                        main_stream.wait_event(ready[b])
                        if u8_graphs is not None:
                            u8_graphs[b][0].replay()
                            out = u8_graphs[b][1]
                        else:
                            out = tail_step(bufs8[b])
                        freed[b].record(main_stream)
                        tail_host.copy_(out, non_blocking=True)
                main_stream.synchronize()

            for b in range(2):
                freed[b].record(main_stream)
            u8_loop(3)
            barrier()
            e0.record()
            u8_loop(args.steps)
            e1.record()
            barrier()
            u8_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
            e2e_u8 = {"value": BATCH / (u8_ms / 1e3), "unit": "images/sec", "ms_per_step": u8_ms,
                      "h2d_bytes_per_step": u8_host.numel() * world
                      // (plan.group_size if use_stage else 1),
                      "d2h_bytes_per_step": 12 * world,
                      "input": "uint8 NCHW, normalised on the device; result = loss + top-1/top-5 "
                               "counts of the batch (devit_eval_tail)"}
        except Exception as e:  # noqa: BLE001  (deterministic host-side failures hit every rank alike)
            print(f"[bench] e2e_u8 arm failed ({type(e).__name__}: {e})", file=sys.stderr)
            torch.cuda.synchronize()

    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel-family device times: CUDA events around every launch, eager, ONE stream and
    #      whole-chip grids (each kernel timed alone, which is what its roofline fraction is about;
    #      inside the step two chains share the chip, so these do not add up to ms_per_step)
    L.profile_enable(True)
    nprof = 3
    for _ in range(nprof):
        ens(x_dev)
    torch.cuda.synchronize()
    prof = L.profile_collect()
    L.profile_enable(False)
    fam = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] // nprof}
           for k, v in prof.items()}

    # ---- roofline of the dominant kernel on THIS rank's shard: the fused projection + MLP
    #      kernel (csrc/mlp.cu: proj + residual + LN-folded fc1 + GELU + fc2 + residual).
    peaks = load_peaks()
    subs_local = plan.subs if fuse is not None else [0]
    gemm_fl = attn_fl = tail_fl = 0
    for s in subs_local:
        f = wl.sub_flops(s)
        gemm_fl += f["gemm"] * Bg
        attn_fl += f["attn"] * Bg
        tail_fl += f["tail"] * Bg
    gemm_fl += wl.fusion_flops() * Bg
    gemm_ms = sum(v["ms_per_step"] for k, v in fam.items() if k.startswith("gemm"))
    gemm_launches = sum(v["launches_per_step"] for k, v in fam.items() if k.startswith("gemm"))
    attn_ms = fam.get("attention", {}).get("ms_per_step", 0.0)
    total_fl = wl.total_flops_per_image() * BATCH
    mlp = fam.get("gemm_mlp_fused")
    proj_sep = fam.get("gemm_proj", {}).get("ms_per_step", 0.0)
    if mlp and mlp["launches_per_step"]:
        dom_name = ("devit::mlp_fused_kernel<D, PROJ> (tcgen05 cta_group::2: attention-output "
                    "projection + residual + LN-folded fc1 + GELU + fc2 + residual)")
        dom_fl, dom_ms, dom_n = tail_fl, mlp["ms_per_step"] + proj_sep, mlp["launches_per_step"]
    else:  # fp32 mode / fused kernel not instantiated for this width: the GEMM family as a whole
        dom_name = "devit::gemm_kernel<BN> (tcgen05, all GEMM launches of a step)"
        dom_fl, dom_ms, dom_n = gemm_fl, gemm_ms, gemm_launches
    achieved = dom_fl / (dom_ms / 1e3) / 1e12 if dom_ms else None
    gemm_tf = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else None
    # denominator: the burst figure when the clocks sampled during the run sit near the maximum
    # (short kernels timed alone), else the sustained one; both fractions are reported
    near_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and
                    clocks["sm_mhz"] >= 0.93 * clocks["sm_max_mhz"])
    peak_key = "bf16_burst" if (near_max or clocks is None) else "bf16_sustained"
    traffic = None
    tfile = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if world == 1 and tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            traffic = tj.get(f"{wl.name}_{'shrunk' if wl.shrunk else 'dense'}_n1_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    step_tf = total_fl / (ms_step / 1e3) / 1e12
    roofline = {
        "kernel": dom_name,
        "bound": "tensor", "achieved": achieved, "peak": peaks[peak_key],
        "unit": "TFLOP/s", "frac": (achieved / peaks[peak_key]) if achieved else None,
        "frac_of_burst": (achieved / peaks["bf16_burst"]) if achieved else None,
        "frac_of_sustained": (achieved / peaks["bf16_sustained"]) if achieved else None,
        "peak_source": f"{peaks['source']} {peak_key} (sm clock {clocks['sm_mhz'] if clocks else '?'}"
                       f" of {clocks['sm_max_mhz'] if clocks else '?'} MHz during the run)",
        "flops_per_launch": dom_fl / dom_n if dom_n else None,
        "ms_per_launch": dom_ms / dom_n if dom_n else None,
        "launches_per_step": dom_n,
        "traffic": traffic,
        "traffic_source": ("ncu dram__bytes_read+write per launch, profiles/dominant_kernel_traffic.json"
                           if traffic else "not captured for this config / world size"),
        "share_of_step": dom_ms / sum(v["ms_per_step"] for v in fam.values()) if fam else None,
        "timing": "CUDA events around every launch on its stream, eager, one stream, whole-chip grid",
        "all_gemm": {"tflops": gemm_tf,
                     "frac_of_burst": gemm_tf / peaks["bf16_burst"] if gemm_tf else None,
                     "ms_per_step": gemm_ms, "launches_per_step": gemm_launches},
        "whole_step": {"tflops": step_tf,
                       "frac_of_bf16_burst": step_tf / (peaks["bf16_burst"] * world),
                       "frac_of_bf16_sustained": step_tf / (peaks["bf16_sustained"] * world),
                       "flops_per_image": total_fl / BATCH},
        "attention": {"tflops": attn_fl / (attn_ms / 1e3) / 1e12 if attn_ms else None,
                      "ms_per_step": attn_ms},
        "families": fam,
    }

    # ---- dense sub-object (headline only): the same step with all-ones gates, so that
    #      BASELINE.md's dense headline has a number from the same run
    dense = None
    if wl.name == "headline" and wl.shrunk and not args.no_dense_arm:
        try:
            del graph, e2e_graphs
            wd = Workload("headline", dense=True)
            md, fd = wd.build_device(dev, args.precision)
            ed = parallel.ShardedEnsemble(md, fd, plan, group, stage_group)
            for _ in range(3):
                ld = ed(x_dev)
            torch.cuda.synchronize()
            gd = None
            if not args.no_graph and (world == 1 or not args.eager_multi):
                ok = True
                try:
                    gd = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gd):
                        god = ed(x_dev)
                    gd.replay()
                    torch.cuda.synchronize()
                    ok = torch.equal(god, ld)
                except Exception:  # noqa: BLE001
                    ok = False
                if not all_ranks(ok):
                    gd = None
                    torch.cuda.synchronize()
            dstep = (lambda: gd.replay()) if gd is not None else (lambda: ed(x_dev))
            for _ in range(3):
                dstep()
            dms = timed(dstep, args.steps)
            dmode = "cuda_graph" if gd is not None else "eager"
            if pipe is not None and gd is not None:  # same batches-in-flight policy as `value`
                eds = [ed] + [parallel.ShardedEnsemble(md, fd, plan, e.group, None)
                              for e in ens_slots[1:]]
                dpipe = parallel.BatchPipeline([(lambda e=e: e(x_dev)) for e in eds])

                def run_dpipe(n):
                    dpipe.reset()
                    dpipe.fork()
                    for _ in range(n):
                        dpipe.launch()
                    dpipe.join()

                run_dpipe(3)
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                t0.record()
                run_dpipe(args.steps)
                t1.record()
                barrier()
                dms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
                dmode = f"cuda_graph, {depth} batches in flight"
            dfl = wd.total_flops_per_image() * BATCH
            dtf = dfl / (dms / 1e3) / 1e12
            dense = {"value": BATCH / (dms / 1e3), "unit": "images/sec", "ms_per_step": dms,
                     "whole_step_tflops": dtf,
                     "frac_of_bf16_burst": dtf / (peaks["bf16_burst"] * world),
                     "flops_per_image": dfl / BATCH,
                     "launch_mode": dmode}
        except Exception as e:  # noqa: BLE001
            print(f"[bench] dense arm failed ({type(e).__name__}: {e})", file=sys.stderr)
            torch.cuda.synchronize()

    def finish():
        """Leave without tearing the NCCL communicators down: destroy_process_group() blocks for
        ever once collectives on them have been captured into CUDA graphs that are still alive
        (seen at N = 2).  Everything has been measured and printed by now."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        return finish()

    # ---- CPU baseline (N = 1) + parity of the GPU logits against it on the same images
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = 32
        ips, ms, cores, kind, what, ref_logits, xs = cpu_time(wl, sample, 2, 1)
        cpu = {"value": ips, "unit": "images/sec", "cores": cores, "kind": kind,
               "sample": f"{sample} images/step x 2 steps (+1 warm-up) of the same workload, {what}"}
        got = logits[:sample].float().cpu()
        ref_logits = ref_logits.float()
        err = (got - ref_logits).abs()
        parity["vs_cpu_reference"] = {
            "images": sample, "kind": kind,
            "rel_err": float(err.max() / ref_logits.abs().max()),
            "argmax_mismatch": int((got.argmax(-1) != ref_logits.argmax(-1)).sum()),
            "tolerance": 2e-2 if args.precision == "bf16" else 1e-4}

    line = {
        "metric": wl.metric, "value": value, "unit": "images/sec", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision if args.precision == "bf16" else "tf32x3",
        "data": "synthetic",
        "config": wl.config(world, plan),
        "launch_mode": launch_mode,
        "batches_in_flight": depth if pipe is not None else 1,
        "one_batch_in_flight": one_in_flight,
        "several_batches_in_flight": in_flight,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/sec", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "input_staging": ("1/G of the batch per rank over PCIe + NVLink all-gather "
                                  "(checked equal to the device-resident run)" if use_stage
                                  else "every rank copies its group's whole batch")},
        "e2e_u8": e2e_u8,
        "gpu_launches": int(launches_per_step * args.steps),
        "parity": parity,
        "roofline": roofline,
        "dense": dense,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
