"""The bench line's contract, checked without a GPU: the reference arm (`bench.py --impl reference`: the CPU oracle port on
all host threads) runs here and prints the keys the driver reads with the same `config` / `metric` / `unit` the GPU arm
uses, and the committed round-2 GPU line (`profiles/r2_bench_c2.json`) carries every key of the contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _line(out):
    rows = [l for l in out.splitlines() if l.startswith("{")]
    assert len(rows) == 1, out[-2000:]   # ONE JSON line
    return json.loads(rows[0])


def test_reference_arm_runs_on_cpu_and_mirrors_the_gpu_arms_config():
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert d["impl"] == "reference" and d["metric"] == "compress pts/s" and d["unit"] == "pts/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["gpu_launches"] == 0 and d["data"] == "synthetic" and d["dtype"] == "f64" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config object as the GPU arm builds for this workload
    cloud, cfg, desc = bench._workload("c1", 0)
    assert d["config"] == bench.config_dict(desc, cloud.shape[0])


def test_committed_gpu_line_has_every_key_of_the_contract():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_c2.json")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "strong"):
        assert k in d, k
    assert d["config"]["workload"].startswith("C2") and "model" not in d["config"] and d["scaling"] == "weak" and d["vs_baseline"] is None
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["captured_once"]["dram_bytes_per_step"] == rf["traffic"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 32 * d["config"]["per_gpu_points"] and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.95 * d["clocks"]["sm_max_mhz"]
    st = d["strong"]
    assert st["scaling"] == "strong" and st["points"] == 50_000_000 and st["compress_wall_ms"] > 0
