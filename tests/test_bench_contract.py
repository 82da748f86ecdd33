"""bench.py's reference arm runs on host cores only, so its JSON line can be checked without a GPU; the GPU arm
must refuse to run without a device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env, timeout=timeout,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_reference_arm_json_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-seconds", "0.5"])
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "eddsa_poseidon_verifies_per_sec" and d["unit"] == "verifies/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_gpu_arm_refuses_to_run_without_a_device():
    p = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
