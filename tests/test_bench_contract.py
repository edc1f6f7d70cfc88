"""bench.py's output contract on the arm that runs without a GPU (`--impl reference`): exactly one JSON line on
stdout with the keys the driver reads, whatever libraries print elsewhere."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, CIAOSR_CPU_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the unmodified reference when a copy is on this box (/root/reference or baseline/_ref), else the oracle port
    from oracle import ref_harness as rh
    assert d["cpu_baseline"]["kind"] == ("reference" if rh.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] == 4
    assert "workload" in d["config"] and d["vs_baseline"] is None
    import bench
    assert d["config"] == bench.workload_config(1)            # the same static config dict as our arm's line


def test_reference_arm_falls_back_to_the_port_without_a_reference_copy(tmp_path):
    env = dict(os.environ, CIAOSR_CPU_THREADS="4", CIAOSR_REFERENCE_ROOT=str(tmp_path), CIAOSR_NO_REFERENCE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
