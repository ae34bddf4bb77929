"""The benchmark's JSON contract, exercised on the CPU through the reference arm (no GPU needed)."""
import json
import os
import subprocess
import sys

import pytest

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS='4')
    proc = subprocess.run([sys.executable, os.path.join(REPO_ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
                           '--batch-size', '32'], capture_output=True, text=True, env=env, timeout=280)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline', 'impl'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['steps'] == 2 and line['unit'] == 'frames/s'
    assert line['vs_baseline'] is None and line['higher_is_better'] is True and line['value'] > 0
    # the unmodified reference (oracle/_ref mirror, or /root/reference in the build container); 'port' only if neither exists
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    if os.path.isdir(os.path.join(REPO_ROOT, 'oracle', '_ref', 'morgana')) or os.path.isdir('/root/reference/morgana'):
        assert line['cpu_baseline']['kind'] == 'reference'
    assert line['gpu_launches'] == 0
    assert line['e2e']['value'] == line['value'] and line['e2e']['h2d_bytes_per_step'] == 0
    assert 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    proc = subprocess.run([sys.executable, os.path.join(REPO_ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                           '--warmup', '0'], capture_output=True, text=True, env=env, timeout=120)
    assert proc.returncode == 0 and proc.stdout.strip() == ''


def test_reference_arm_does_not_load_the_product_library():
    """`--impl reference` must not map libmorgana_b200.so (the driver records which in-tree .so files the arm loads)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch-size', '8'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read();"
            "assert 'libmorgana_b200' not in maps, 'product library mapped';"
            "assert not any(m == 'morgana_b200' or m.startswith('morgana_b200.') for m in sys.modules), 'product package imported'"
            % os.path.join(REPO_ROOT, 'bench.py'))
    proc = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=280,
                          env=dict(os.environ, OMP_NUM_THREADS='4'))
    assert proc.returncode == 0, proc.stderr[-2000:]
