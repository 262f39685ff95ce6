"""CPU: the reference arm of bench.py (`--impl reference`, the reference's algorithm on the host cores through the oracle port) prints ONE
JSON line with the contract's keys.  The B200 arm needs a GPU and is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output = True, text = True, timeout = 600, cwd = ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'tokens/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['metric'].startswith('train tokens/sec') and d['config']['seq_len'] == 1024
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value = d['value'], unit = 'tokens/s', h2d_bytes_per_step = 0, d2h_bytes_per_step = 0)


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK = '1', WORLD_SIZE = '2', LOCAL_RANK = '1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                         capture_output = True, text = True, timeout = 120, cwd = ROOT, env = env)
    assert out.returncode == 0 and out.stdout.strip() == ''
