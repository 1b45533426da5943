#!/usr/bin/env python
"""Headline benchmark: images/sec of the 4-way DeDeiT ensemble (shrink_ratio-0.3 head/neuron
gates, 100 classes) @224^2, global batch 256, on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                       (N=1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      the reference algorithm on the host CPU (oracle port)

One step = one forward of the whole ensemble (MultiViT + EnsMLP fusion head) over the global
batch.  `value` = device-resident inputs (CUDA-graph replay), `e2e` = through the public module
API with the batch copied host->device from pinned memory and the logits read back every step.
Global batch is fixed at 256 for every N (strong scaling); sub-models are sharded one per rank
up to 4 ranks, 8 ranks = 2 data-parallel groups of 4.  Inputs (154 MB fp32 per batch) and the
activation working set (~0.4 GB) are larger than the 126 MB L2, so no explicit flush is needed.
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

N_SUB, NUM_CLASS, BATCH = 4, 100, int(os.environ.get("DEVIT_BENCH_BATCH", "256"))
D, HEADS, HIDDEN, TOKENS, PATCHES, DEPTH = 384, 6, 1536, 198, 196, 12
METRIC = "images/sec, 4-way DeDeiT ensemble @224^2 bs256"


# ------------------------------------------------------------------------------- flop model
def submodel_flops(kept_heads, kept_neurons):
    """Algorithmic FLOPs (2*MAC) per image of one sub-model: patch GEMM, QKV,
    QK^T, PV, proj, fc1, fc2 with the KEPT head / neuron counts.  -> (gemm, attention)."""
    gemm = 2 * PATCHES * 768 * D
    attn = 0
    for h, f in zip(kept_heads, kept_neurons):
        gemm += 2 * TOKENS * D * 3 * 64 * h + 2 * TOKENS * 64 * h * D + 4 * TOKENS * D * f
        attn += 4 * h * TOKENS * TOKENS * 64
    return gemm, attn


def fusion_flops():
    return 2 * 2 * (N_SUB * D * 768 + 768 * NUM_CLASS)


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


# ------------------------------------------------------------------------------- CPU arms
def cpu_oracle_ips(images_per_step, steps, warmup, shrunk=True):
    """Times the oracle port of the reference (masked-dense, fp32, all host threads)."""
    import torch
    from devit_b200 import synth
    from oracle import devit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds = [synth.dedeit_state_dict(s, with_heads=False) for s in range(N_SUB)]
    esd = synth.ensmlp_state_dict(N_SUB, num_class=NUM_CLASS)
    gates = [synth.shrink_gates(s) for s in range(N_SUB)] if shrunk else None
    x = synth.images(images_per_step)
    with torch.no_grad():
        for _ in range(warmup):
            O.ensemble_logits(sds, esd, x, gates)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.ensemble_logits(sds, esd, x, gates)
        dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 16
    warm = min(args.warmup, 2)
    ips, ms, cores = cpu_oracle_ips(sample, args.steps, warm, shrunk=not args.dense)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step of the bs-256 workload, oracle "
                                   f"port of the reference (masked-dense fp32 PyTorch CPU)"},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": ("4-way DeDeiT ensemble (dedeit D=384 depth 12), "
                         + ("dense gates" if args.dense else "shrink_ratio-0.3 head+neuron gates")
                         + ", EnsMLP fusion head 100 classes, 224x224, global batch 256"),
            "global_batch": BATCH, "sub_models": N_SUB,
            "parallelism": ("1 GPU: 4 sub-models sequential" if world == 1 else
                            f"{min(world, N_SUB)}-way sub-model sharding x "
                            f"{max(1, world // N_SUB)} data-parallel group(s), all-gather of "
                            f"[2,B,384] features"),
            "l2": "inputs (154 MB) + activations (>0.4 GB) exceed the 126 MB L2; no flush"}


# ------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="devit", choices=["devit", "reference"])
    ap.add_argument("--dense", action="store_true", help="all-ones gates instead of shrunk")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--force-graph", action="store_true", help="capture NCCL too (N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from devit_b200 import _lib as L
    from devit_b200 import ensemble, parallel, shrink, synth

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

    plan = parallel.shard_plan(world, rank, N_SUB, BATCH)
    group = parallel.make_groups(plan) if world > 1 else None

    multi = ensemble.MultiViT(model="dedeit", drop=0, drop_path=0.1,
                              num_classes_list=[NUM_CLASS // N_SUB] * N_SUB, num_div=N_SUB)
    fuse = ensemble.EnsMLP(model="dedeit", num_class=NUM_CLASS, sub_size=D,
                           num_classes_list=[NUM_CLASS // N_SUB] * N_SUB, teacher_size=768)
    kept = []
    for s in range(N_SUB):
        multi.backbones[s].load_state_dict(synth.dedeit_state_dict(s, with_heads=False))
        if args.dense:
            kept.append(([HEADS] * DEPTH, [HIDDEN] * DEPTH))
        else:
            ng, hg = synth.shrink_gates(s)
            shrink.mlp_neuron_shrink(multi.backbones[s], ng)
            shrink.attn_head_shrink(multi.backbones[s], hg)
            kept.append(([int(g.sum()) for g in hg], [int(g.sum()) for g in ng]))
    fuse.load_state_dict(synth.ensmlp_state_dict(N_SUB, num_class=NUM_CLASS))
    multi = multi.to(dev).eval().set_precision(args.precision)
    fuse = fuse.to(dev).eval().set_precision(args.precision)
    # second communicator over the same ranks: the e2e arms upload 1/G of the batch per rank and
    # all-gather it over NVLink (ShardedEnsemble.stage_batch)
    stage_group = parallel.make_groups(plan) if world > 1 else None
    ens = parallel.ShardedEnsemble(multi, fuse, plan, group, stage_group)

    Bg = plan.group_batch
    x_host = synth.images(BATCH)[plan.batch_lo:plan.batch_hi].contiguous().pin_memory()
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

    # ---- warm-up (also builds packs / workspaces) and launch count of one step
    lib = L.load()
    logits = ens(x_dev)
    torch.cuda.synchronize()
    c0 = lib.devit_launch_count()
    logits = ens(x_dev)
    torch.cuda.synchronize()
    launches_per_step = lib.devit_launch_count() - c0
    for _ in range(args.warmup):
        ens(x_dev)
    torch.cuda.synchronize()

    # ---- optional CUDA graph of the whole step (launch-bound host loop -> one replay)
    graph, g_out = None, None
    if not args.no_graph and (world == 1 or args.force_graph):
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
            if rank == 0:
                print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); "
                      f"timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            ens(x_dev)

    for _ in range(args.warmup):
        step()

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)

    # ---- value: device-resident inputs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = BATCH / (ms_step / 1e3)

    # ---- e2e: public API with host buffers; H2D of the step's batch + D2H of its logits inside
    #      the timed region, input copies double-buffered on a copy stream
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    out_host = torch.empty(Bg, NUM_CLASS).pin_memory()
    main_stream = torch.cuda.current_stream()

    # the same public call, captured once per input buffer when CUDA graphs are in use (a user
    # serving fixed-shape batches would do the same); eager launches otherwise
    e2e_graphs = None
    if graph is not None:
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
            if rank == 0:
                print(f"[bench] e2e graph capture unavailable ({type(e).__name__}: {e})",
                      file=sys.stderr)
            e2e_graphs = None
            torch.cuda.synchronize()

    def e2e_loop(n):
        for i in range(n + 1):
            if i < n:  # prefetch batch i
                b = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[b])
                    if use_stage:
                        ens.stage_batch(x_host, bufs[b])
                    else:
                        bufs[b].copy_(x_host, non_blocking=True)
                    ready[b].record(copy_stream)
            if i > 0:  # compute batch i-1
                b = (i - 1) & 1
                main_stream.wait_event(ready[b])
                if e2e_graphs is not None:
                    e2e_graphs[b][0].replay()
                    out = e2e_graphs[b][1]
                else:
                    out = ens(bufs[b])
                freed[b].record(main_stream)
                out_host.copy_(out, non_blocking=True)
        main_stream.synchronize()

    # multi-GPU: every model rank of a group needs the same images; rank r uploads rows
    # [r Bg/G, (r+1) Bg/G) and an NVLink all-gather assembles the batch (1/G of the PCIe bytes per
    # rank).  Checked against the device-resident result before it is timed; any disagreement
    # (on any rank) falls back to every rank copying the whole batch.
    use_stage = world > 1 and parallel.stage_slice(plan, Bg) is not None
    for b in range(2):
        freed[b].record(main_stream)
    e2e_loop(3)
    if use_stage:
        ok = torch.tensor([1 if torch.equal(out_host.to(dev), logits) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if rank == 0:
                print("[bench] staged input gave different logits; timing full per-rank copies",
                      file=sys.stderr)
            use_stage = False
            for b in range(2):
                freed[b].record(main_stream)
            e2e_loop(3)
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

    # ---- e2e_u8: the same public call fed DECODED uint8 batches: ToTensor +
    #      Normalize run inside the patch extraction on the device, so a step's H2D copy is a
    #      quarter of the fp32 one, and the step's result is the device-side eval tail (loss,
    #      correct@1, correct@5: 12 bytes D2H) instead of the logits.  Reported beside `e2e`.
    e2e_u8 = None
    try:
        u8_host = synth.images_u8(BATCH)[plan.batch_lo:plan.batch_hi].contiguous().pin_memory()
        tgt = torch.randint(0, NUM_CLASS, (Bg,), generator=torch.Generator().manual_seed(7)).to(dev)
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

    # ---- per-kernel-family device times (CUDA events around every launch, eager)
    L.profile_enable(True)
    nprof = 3
    for _ in range(nprof):
        ens(x_dev)
    torch.cuda.synchronize()
    prof = L.profile_collect()
    L.profile_enable(False)
    fam = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] // nprof}
           for k, v in prof.items()}

    # ---- roofline of the dominant kernel on THIS rank's shard.  The fused MLP kernel
    #      (fc1 + GELU + fc2, csrc/mlp.cu) is ~45 % of the step and 60 % of its FLOPs; the
    #      aggregate over every tcgen05 GEMM launch and the attention kernel are given beside it.
    peaks = load_peaks()
    gemm_fl = attn_fl = mlp_fl = 0
    for s in plan.subs:
        g, a = submodel_flops(*kept[s])
        gemm_fl += g * Bg
        attn_fl += a * Bg
        mlp_fl += sum(4 * TOKENS * D * f for f in kept[s][1]) * Bg
    gemm_fl += fusion_flops() * Bg
    gemm_ms = sum(v["ms_per_step"] for k, v in fam.items() if k.startswith("gemm"))
    gemm_launches = sum(v["launches_per_step"] for k, v in fam.items() if k.startswith("gemm"))
    attn_ms = fam.get("attention", {}).get("ms_per_step", 0.0)
    total_fl = sum(sum(submodel_flops(*kept[s])) for s in range(N_SUB)) * BATCH \
        + fusion_flops() * BATCH
    traffic = None
    tfile = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(
                "dense_dram_bytes_per_launch" if args.dense else "shrunk_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    mlp = fam.get("gemm_mlp_fused")
    if mlp and mlp["launches_per_step"]:
        dom_name = "devit::mlp_fused_kernel (tcgen05 cta_group::2: LN-folded fc1 + GELU + fc2 + residual)"
        dom_fl, dom_ms, dom_n = mlp_fl, mlp["ms_per_step"], mlp["launches_per_step"]
    else:  # fp32 mode / fused kernel switched off: the GEMM family as a whole
        dom_name = "devit::gemm_kernel<BN> (tcgen05, all GEMM launches of a step)"
        dom_fl, dom_ms, dom_n = gemm_fl, gemm_ms, gemm_launches
    achieved = dom_fl / (dom_ms / 1e3) / 1e12 if dom_ms else None
    gemm_tf = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else None
    roofline = {
        "kernel": dom_name,
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"],
        "unit": "TFLOP/s", "frac": (achieved / peaks["bf16_sustained"]) if achieved else None,
        "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a step)",
        "flops_per_launch": dom_fl / dom_n if dom_n else None,
        "ms_per_launch": dom_ms / dom_n if dom_n else None,
        "launches_per_step": dom_n, "traffic": traffic,
        "share_of_step": dom_ms / sum(v["ms_per_step"] for v in fam.values()) if fam else None,
        "all_gemm": {"tflops": gemm_tf, "frac": gemm_tf / peaks["bf16_sustained"] if gemm_tf else None,
                     "ms_per_step": gemm_ms, "launches_per_step": gemm_launches},
        "whole_step": {"tflops": total_fl / (ms_step / 1e3) / 1e12,
                       "frac_of_bf16_burst": total_fl / (ms_step / 1e3) / 1e12
                       / (peaks["bf16_burst"] * world),
                       "flops_per_image": total_fl / BATCH},
        "attention": {"tflops": attn_fl / (attn_ms / 1e3) / 1e12 if attn_ms else None,
                      "ms_per_step": attn_ms},
        "families": fam,
    }

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = 32
        ips, ms, cores = cpu_oracle_ips(sample, 2, 1, shrunk=not args.dense)
        cpu = {"value": ips, "unit": "images/sec", "cores": cores, "kind": "port",
               "sample": f"{sample} images/step x 2 steps (+1 warm-up) of the same 4-way "
                         f"ensemble, oracle port of the reference (masked-dense fp32 PyTorch CPU)"}

    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision if args.precision == "bf16" else "tf32x3",
        "data": "synthetic",
        "config": workload_config(args, world),
        "launch_mode": "cuda_graph" if graph is not None else "eager",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/sec", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "input_staging": ("1/G of the batch per rank over PCIe + NVLink all-gather "
                                  "(checked equal to the device-resident run)" if use_stage
                                  else "every rank copies its group's whole batch")},
        "e2e_u8": e2e_u8,
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
