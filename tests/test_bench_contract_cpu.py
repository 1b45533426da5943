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
