"""CPU checks of bench.py's contract: the reference arm (`--impl reference`) runs without a GPU, prints exactly one JSON
line with the keys the driver reads, and refuses to run the B200 arm on a machine without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, 'bench.py')


def _run(*args, env=None):
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line():
    out = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-pairs-log2', '12')
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'spd4_pair_dist_grad_evals_per_sec' and d['unit'] == 'pairs/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_stay_silent():
    out = _run('--impl', 'reference', '--gpus', '2', '--steps', '1', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_reference_arm_secondary_workloads():
    out = _run('--impl', 'reference', '--workload', '2a,3b', '--steps', '1', '--warmup', '1')
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{')]
    assert [r['config'] for r in rows] == ['2a', '3b']
    assert all(r['impl'] == 'reference' and r['epoch_ms'] > 0 and r['pairs_per_epoch'] > 10**6 for r in rows)


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without CUDA')
def test_b200_arm_refuses_to_run_without_cuda():
    out = _run('--steps', '1', '--warmup', '1')
    assert out.returncode != 0 and 'no CPU fallback' in (out.stderr + out.stdout)


def test_shipped_graphs_and_config_table_are_consistent():
    """bench.py --workload 1..4 runs on the shipped edge lists (tests/golden/graphs/*.npz): they are present, have the
    sizes the config table names, and every GPU config has a CPU twin for the baseline beside it."""
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.GPU_CONFIGS) == set(bench.CPU_CONFIGS) == {'1', '2a', '2b', '3a', '3b', '4'}
    for tag, (gname, factors, dtype, opt, batch, E, flops) in bench.GPU_CONFIGS.items():
        n, edges, src = bench.load_graph(gname)
        assert n == bench.GRAPH_SIZES[gname] and src == f'data/{gname}.edges.gz'
        assert edges.ndim == 2 and edges.shape[1] == 2 and edges.max() == n - 1
        assert bench.CPU_CONFIGS[tag][3] == n and len(E) == len(factors)
