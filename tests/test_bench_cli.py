"""bench.py / __graft_entry__.py contract checks that need no GPU: both files compile, and the
``--impl reference`` arm (CPU oracle port on a bounded sample) prints one well-formed JSON line."""
import json
import os
import py_compile
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_entry_points_compile():
    for f in ("bench.py", "__graft_entry__.py"):
        py_compile.compile(os.path.join(ROOT, f), doraise=True)


def _run_reference(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "n=60000", "--cpu-sample", "40000"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_line():
    # torchrun exports OMP_NUM_THREADS=1; the arm must still use all host cores
    lines = _run_reference({"OMP_NUM_THREADS": "1"})
    line = json.loads(lines[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    lines = _run_reference({"RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []
