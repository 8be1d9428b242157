"""
The bench.py contract on the CPU: the reference arm runs without a GPU and
prints one JSON line with the agreed keys; the GPU arm refuses to run without
a device instead of falling back to anything.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ)
    env.pop('RANK', None)
    env.pop('WORLD_SIZE', None)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args),
                          capture_output=True, text=True, cwd=ROOT, env=env,
                          timeout=timeout)


def test_reference_arm_json_line():
    r = run_bench('--impl', 'reference', '--steps', '4', '--warmup', '1',
                  '--cpu-grid', '16')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['unit'] == 'cell-steps/s' and d['higher_is_better'] is True
    assert d['steps'] == 4 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1
    assert cb['value'] == d['value'] and '16x16' in cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_gpu_arm_needs_a_gpu():
    from myokit_b200 import capi
    if capi.device_count() > 0:
        return      # on a GPU box the arm runs; covered by the driver
    r = run_bench('--steps', '2', '--warmup', '1', '--grid', '32', '--no-cpu',
                  timeout=300)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]
