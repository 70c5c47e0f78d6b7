"""bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`, the CPU oracle port on a bounded sample)
prints one JSON line with the keys of the task statement, and under a multi-rank launch only rank 0 prints it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    env.update(env_extra or {})
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '3', '--ref-problems', '8']
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'rti_iterations_per_sec' and d['higher_is_better'] is True
    assert d['steps'] == 2 and d['warmup'] == 3 and d['n_gpus'] == 1 and d['dtype'] == 'f64' and d['vs_baseline'] is None
    assert d['value'] > 0 and abs(d['ms_per_step'] * d['steps'] * 1e-3 * d['value'] - 8 * 2) < 1e-6 * 16      # 8 problems x 2 steps, one solve each
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and '8 problems x 2' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'cfg1' in d['config']['workload'] and 'model' not in d['config']


def test_reference_arm_only_rank_zero_prints():
    res = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2', 'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': '29571'})
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.startswith('{')]
