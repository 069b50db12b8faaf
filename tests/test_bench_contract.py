"""bench.py's output contract on the CPU: the reference arm (the oracle port on the host cores) prints exactly one JSON
line on stdout with the keys the driver reads; everything else goes to stderr."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='1')
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--sims', '4'], capture_output=True, text=True, timeout=600, env=env, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'mcts_simulations_per_second' and d['unit'] == 'sims/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['steps'] == 1 and d['warmup'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'sims/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
