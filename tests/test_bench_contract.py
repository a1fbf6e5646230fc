"""bench.py's output contract, checked on CPU: the committed round-2 bench lines (profiles/r2_bench_n*.json, builder-run
on the B200 pool) carry every key the driver reads and agree with themselves; the reference arm (`--impl reference`, the
CPU PyTorch path of the step - models/registration_model.py:141-171 - through oracle/torch_port) runs here and prints a
line of the same shape; the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _baseline():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_keep_the_contract(n):
    with open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")) as f:
        d = json.load(f)
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["gpu_launches"] > 0
    assert d["vs_baseline"] is None                     # BASELINE.md publishes no number for this metric
    cfg = d["config"]
    assert "workload" in cfg and "model" not in cfg and "configs[1]" in cfg["workload"]
    assert cfg["global_batch"] == cfg["batch_per_gpu"] * n
    # whole-job throughput = pairs of all ranks / max-over-ranks step time
    assert d["value"] == pytest.approx(cfg["global_batch"] / d["ms_per_step"] * 1e3, rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]     # host copies inside the timed region: never faster
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6) and 0.0 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.8 * c["sm_max_mhz"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        assert d["parity"]["warp_index_mismatches"] == 0
        for w in d["workloads"].values():                # the 3-D workloads ride on the N = 1 line
            assert {"value", "ms_per_step", "e2e", "roofline", "cpu_baseline", "config"} <= set(w)


def test_metric_is_the_baselines():
    b = _baseline()
    with open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")) as f:
        d = json.load(f)
    assert b["metric"].startswith(d["metric"])          # "volume-pairs/sec (fwd+bwd)" of BASELINE.json
    assert d["unit"] == "pairs/s"


def test_reference_arm_prints_a_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--size", "64", "--no-3d"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"]
    assert "configs[1]" in d["config"]["workload"] and "sample" in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_product_arm_has_no_cpu_fallback():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
