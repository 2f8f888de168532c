"""CPU: bench.py's reference arm prints the contract's JSON line (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, BENCH_REF_SAMPLE="1024")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
    assert line["config"]["config_index"] == 2 and "reference_sample" in line["config"]
    assert line["cpu_baseline"]["cpu_1t"]["as_is"] > 0          # 1 thread and all cores, as-is and amortised


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_other_configs_and_same_config_keys():
    """--config 1 / 4 (one serial fsm_exec call) and the config dict both arms print: a function of the
    command line only (bench.the_config), so the driver's same_config comparison holds."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for cfg in (1, 4):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cfg), "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr
        line = json.loads(p.stdout.strip().splitlines()[-1])
        ns = argparse.Namespace(config=cfg, dist="uniform", size=1 << 20)
        assert line["config"] == bench.the_config(ns) and line["cpu_baseline"]["cores"] == 1 and line["value"] > 0
