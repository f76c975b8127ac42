"""bench.py's JSON contract on the arm that runs without a GPU: `--impl reference` times the reference's CPU
implementation of the path (oracle pipeline with the compiled reference aligner, or its C restatement) on a bounded
sample of the same workload and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from .conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--cpu-reads', '2'], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.split('\n') if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'reads/s' and d['unit'] == 'reads/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config'] and 'model' not in d['config']
    assert d['e2e'] == {'value': d['value'], 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['value'] == d['value'] and cb['cores'] >= 1 and cb['kind'] in ('reference', 'port', 'reference-aligner+restatement')
    assert 'sample' in cb


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ''
