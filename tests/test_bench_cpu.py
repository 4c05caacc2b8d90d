"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own files from oracle/_ref, or the oracle
port, on the host cores at the workload's full size -- here cut down with the development switch --n) prints one
JSON line with the contract's keys; both kinds produce the same labels digest."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(kind, extra=()):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--n", "1536", "--ref-kind", kind, *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_prints_contract_line():
    line = _run("port")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "s" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["steps_run"] == 1 and line["config"]["N"] == 1536
    assert len(line["config"]["labels_sha256"]) == 64 and "scal" not in line["cpu_baseline"]["sample"].replace("no scaling", "")
    # thread pinning does not depend on what the launcher exported (torchrun sets OMP_NUM_THREADS=1)
    assert line["cpu_baseline"]["threads"]["blas"] == os.cpu_count()


def test_reference_arm_runs_the_reference_itself_when_present():
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.available():
        import pytest
        pytest.skip("neither /root/reference nor oracle/_ref present")
    ref = _run("_ref")
    port = _run("port")
    assert ref["cpu_baseline"]["kind"] == "reference" and port["cpu_baseline"]["kind"] == "port"
    assert ref["config"]["labels_sha256"] == port["config"]["labels_sha256"]
    assert ref["config"]["rank_sha256"] == port["config"]["rank_sha256"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--n", "512"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
