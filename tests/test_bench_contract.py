"""CPU: bench.py's reference arm prints ONE JSON line with the keys the driver reads (metric, value, e2e, cpu_baseline, ...)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-batch', '1', '--cpu-T', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'GCRNN sequences/sec fwd+bwd' and d['unit'] == 'sequences/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['e2e'] == dict(value=d['value'], unit='sequences/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert 'workload' in d['config']


def test_traffic_reader_finds_committed_export():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic(['r01_ncu_gemm2_mb1024.raw.csv'], 'shift_gemm2_kernel')
    assert src is not None and 50e6 < t < 600e6


def test_traffic_reader_at_the_default_launch_shape():
    """bench.py's `roofline.traffic` must come from a capture of the SAME launch shape it times (micro-batch 2048)."""
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic(['r01_ncu_gemm2_mb2048.raw.csv'], 'shift_gemm2_kernel')
    operands = 2 * (2048 * 64 * 1024) * 2 + 1024 * 1024 * 2          # bf16 A in, bf16 C out, bf16 operator
    assert src == 'r01_ncu_gemm2_mb2048.raw.csv' and 0.5 * operands < t < 1.5 * operands


def test_cfg5_measured_traffic_per_sequence():
    """cfg5's roofline object reports measured DRAM bytes per sequence next to the 17.2 GB algorithmic figure; the committed
    ncu export must contain one launch of each of the seven per-step kernels."""
    sys.path.insert(0, ROOT)
    import bench
    t = bench.cfg5_dram_bytes_per_sequence(bench.CFG5)
    assert t is not None and 0.5 * bench.BYTES_PER_SEQ_CFG5 < t < 1.5 * bench.BYTES_PER_SEQ_CFG5


def test_multi_rank_stdout_carries_only_the_json_line():
    """Under torchrun, libraries write to fd 1 behind Python's back (NCCL's version banner): bench.py keeps a private copy of the real
    stdout for its ONE JSON line and points fd 1 at stderr."""
    code = ("import os, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-batch', '1', '--cpu-T', '1'];"
            "import bench; bench.json_only_stdout(); os.system('echo BANNER_FROM_FD1'); bench.main()")
    env = dict(os.environ, WORLD_SIZE='2', RANK='0')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0])['impl'] == 'reference', out.stdout
    assert 'BANNER_FROM_FD1' in out.stderr
