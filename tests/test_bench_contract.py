"""bench.py's contract on a machine without a GPU: the reference arm runs (CPU only) and prints one
valid JSON line with the required keys; the product arm refuses loudly instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "40", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["dtype"] == "f64" and d["steps"] == 3
    assert d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    import oracle
    # the reference's own loop text where oracle/_ref was built, the restatement elsewhere
    assert d["cpu_baseline"]["kind"] == ("reference" if oracle.ref_lib() is not None else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["rows"] == 40 ** 3 and d["config"]["nnz"] == 7 * 40 ** 3 - 6 * 40 ** 2
    assert d["config"]["algorithmic_bytes"] == d["config"]["nnz"] * 12 + d["config"]["rows"] * 20


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--grid", "20",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--grid", "20", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)


def test_reference_arm_does_not_map_the_product_library():
    """The reference arm must run on oracle/ alone: no libb200aij.so / libb200petsc.so in its address
    space (the driver records which in-tree .so files each arm loads)."""
    code = ("import sys, argparse; sys.path.insert(0, %r); import bench; "
            "bench.run_reference(argparse.Namespace(grid=20, steps=1, warmup=1, gpus=1, workload='matmult')); "
            "maps = open('/proc/self/maps').read(); "
            "assert 'libb200' not in maps, [l for l in maps.splitlines() if 'libb200' in l]; "
            "assert 'petsc_openacc_b200' not in sys.modules; "
            "assert 'liboracle' in maps or 'libref_matmult' in maps") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
