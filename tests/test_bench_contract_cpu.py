"""bench.py's reference arm runs on the host cores alone, so its JSON contract can be checked
without a GPU: one line, the metric / config of the GPU arm, `impl`, `cpu_baseline` and `e2e`."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    import os
    env = dict(os.environ, DEVIT_REF_BUDGET_S='8')  # a small per-step sample keeps the CPU suite short
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/sec' and d['higher_is_better'] is True
    assert d['metric'].startswith('images/sec, 4-way DeDeiT ensemble') and d['value'] > 0
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['gpu_launches'] == 0
    assert d['config']['global_batch'] == 256 and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] \
        and cb['sample']
    assert 'reference_step' in d['config']  # the bounded per-step sample is stated
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--steps', '1'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_flop_model_matches_the_survey_numbers():
    """The algorithmic FLOPs `roofline.achieved` is computed from: dense
    dedeit 9.247 GFLOP / image / sub-model, 4-way dense 36.99, teacher 35.31, 8-way dense 73.99,
    fusion head 5.03 MFLOP (n = 4, C = 100) / 12.5 MFLOP (n = 8, C = 1000); work that a gate
    removes is not counted."""
    sys.path.insert(0, str(ROOT))
    import bench
    dense = bench.vit_flops(384, [6] * 12, [1536] * 12)
    assert abs((dense['gemm'] + dense['attn']) / 1e9 - 9.247) < 5e-3
    assert abs(dense['attn'] / 1e9 - 0.722) < 2e-3                       # QK^T + AV
    assert abs(dense['tail'] / 1e9 - (0.701 + 2.803 + 2.803)) < 5e-3     # proj + fc1 + fc2
    wl = bench.Workload('headline', dense=True)
    wl.kept = [([6] * 12, [1536] * 12)] * 4
    assert abs(wl.fusion_flops() / 1e6 - 5.03) < 0.01
    assert abs(wl.total_flops_per_image() / 1e9 - 36.99) < 0.01
    c3 = bench.Workload('c3')
    c3.kept = [([6] * 12, [1536] * 12)] * 8
    assert abs(c3.fusion_flops() / 1e6 - 12.5) < 0.1
    assert abs(c3.total_flops_per_image() / 1e9 - 73.99) < 0.02
    c1 = bench.Workload('c1')
    c1.kept = [([12] * 12, [3072] * 12)]
    assert abs(c1.total_flops_per_image() / 1e9 - 35.31) < 0.02
    # a shrunk sub-model is credited with its kept heads / neurons only
    half = bench.vit_flops(384, [3] * 12, [768] * 12)
    assert half['tail'] * 2 == dense['tail'] and half['attn'] * 2 == dense['attn']
