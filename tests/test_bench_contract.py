"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm runs here, the committed lines of the last
GPU runs carry every key the contract names, the workloads are BASELINE.json's."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workloads_are_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for key, cfg in zip(("c2", "c3", "c5"), (base["configs"][1], base["configs"][2], base["configs"][4])):
        w = bench.WORKLOADS[key]
        assert w["scene"] in cfg and f"{w['width']}×{w['height']}" in cfg, (key, cfg)
    assert os.path.exists(os.path.join(ROOT, bench.WORKLOADS["c4"]["scene"]))   # the documented stand-in for the bedroom (not obtainable offline)
    assert bench.B_PRIMARY == 216 and bench.B_SHADOW == 108 and bench.B_SPLAT == 24


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "BVH4" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("scenes/diamond_scene.json 1920x1080")
    # other ranks of a torchrun launch do nothing
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract():
    for name, n in (("r6u_bench.json", 1), ("r6s_bench_c2_n2.json", 2), ("r6q_bench_c2_n4.json", 4), ("r6t_bench_c2_n8.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                  "clocks", "e2e", "gpu_launches", "roofline", "parity"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["unit"] == "Mrays/s" and d["dtype"] == "f32" and d["vs_baseline"] is None and d["scaling"] == "strong"
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 1920 * 1080 * 12 and 0 < d["e2e"]["value"] < d["value"]
        assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        assert d["parity"]["ok"] and d["parity"]["rel_l2"] <= 1e-4 and d["parity"]["ray_counts_equal"]
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0 and "issue_roofline" in d and r["traffic"] > 0
            assert abs(sum(k["share"] for k in d["kernels"].values()) - 1) < 1e-6
