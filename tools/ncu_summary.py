"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv>        -> per-kernel count / time / share
  python tools/ncu_summary.py rep <file.ncu-rep> [...]        -> per-launch key metrics

Runs on the CPU box (ncu -i reads a report without a GPU).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % of SM-active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "hmma subpipe active cycles (avg/TPC)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe insts"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory (TMA/UMMA smem) active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem bank conflicts (lsu)"),
    ("gpc__cycles_elapsed.max", "elapsed cycles"),
    ("gpc__cycles_elapsed.avg.per_second", "gpc clock"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_pipe_xu.sum", "XU-pipe (MUFU) warp instructions"),
    ("smsp__inst_executed_pipe_fma.sum", "FMA-pipe warp instructions"),
    ("smsp__inst_executed_pipe_alu.sum", "ALU-pipe warp instructions"),
    ("smsp__inst_executed_pipe_lsu.sum", "LSU-pipe warp instructions"),
    ("smsp__inst_executed_pipe_uniform.sum", "uniform-pipe warp instructions"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % of peak (active)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe % of peak (active)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe % of peak (active)"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard (cyc/inst)"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall membar"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio throttle"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg throttle"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_sleeping.ratio", "stall sleeping"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not selected"),
    ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall dispatch"),
    ("smsp__average_warp_latency_issue_stalled_tex_throttle.ratio", "stall tex throttle"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch resolving"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall no instruction"),
    ("smsp__average_warp_latency_issue_stalled_selected.ratio", "selected"),
    ("smsp__average_warp_latency_issue_stalled_gmma.ratio", "stall gmma/tensor"),
]


def launches(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms summed gpu__time_duration "
          f"(cold-cache, serialised: compare shares)")
    print(f"{'launches':>8} {'sum us':>10} {'avg us':>8} {'share':>6}  kernel (grid, block)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / 1e3 / v[0]:8.2f} {100 * v[1] / total:5.1f}%  "
              f"{k} {v[2]} {v[3]}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        print(f"## {r[col['Kernel Name']].split('(')[0]}  (launch id {r[col['ID']]})")
        for key, label in KEYS:
            if key in col:
                print(f"  {label:<44} {r[col[key]]:>16} {units[col[key]]:<12} [{key}]")
        rd, wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            def to_b(v, u):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                return float(v.replace(",", "")) * mult
            t = to_b(r[rd], units[rd]) + to_b(r[wr], units[wr])
            dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
            du = units[col["gpu__time_duration.sum"]]
            sec = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[du if du in ("ns", "us", "ms", "s") else "us"]
            print(f"  {'traffic (dram read+write)':<44} {t / 1e6:16.3f} MB  -> {t / sec / 1e9:.1f} GB/s under ncu")
        print()


if __name__ == "__main__":
    mode = sys.argv[1]
    for p in sys.argv[2:]:
        launches(p) if mode == "launches" else rep(p)
