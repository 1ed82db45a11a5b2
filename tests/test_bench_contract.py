"""bench.py's reference arm (the unmodified reference from baseline/_ref, or its NumPy port where that copy is absent)
runs here and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
           '--atoms', '600', '--frames', '400', '--segments', '10', '--cpu-atoms', '600']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'evals/s' and line['value'] > 0
    assert line['higher_is_better'] is True and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] == 1
    assert line['scaling'] == 'strong' and 'sample_atoms' not in line['config']
    assert line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    for key in ('metric', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config'):
        assert key in line
    assert 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
           '--warmup', '0', '--atoms', '100', '--frames', '100']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ''


def test_stdout_carries_only_the_json_line():
    """Whatever a library writes to file descriptor 1 (NCCL prints its version banner there) must not end up
    in front of the result line: bench.py keeps the real stdout for the line and points fd 1 at stderr."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 0.0\\n'); print('python-level noise'); bench.emit({'ok': 1})" % ROOT)
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout == '{"ok": 1}\n'
    assert 'NCCL version 0.0' in res.stderr and 'python-level noise' in res.stderr
