"""CPU: the reference arm of bench.py (`--impl reference`: the oracle restatement of the reference op sequence on the host cores) runs
without a GPU and prints ONE JSON line carrying every key of the driver contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
              'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'fitting_iters_per_sec' and d['higher_is_better'] is True
    assert d['scaling'] == 'weak' and d['vs_baseline'] is None and d['data'] == 'synthetic' and d['steps'] == 1 and d['n_gpus'] == 1
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value'] and cb['unit'] == d['unit']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert d['value'] > 0 and d['ms_per_step'] > 0
