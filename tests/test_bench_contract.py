"""The JSON line bench.py prints (the driver's contract): checked on the line stored from the last B200 run of the round, and on the
reference-arm line.  (bench.py itself needs a GPU; `--impl reference` runs the CPU oracle for minutes - neither runs here.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    text = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1]
    return json.loads(text)


def test_bench_line_has_the_contract_keys():
    d = _line("r01_bench_c3_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert d["config"]["workload"] == "c3_icospheres_1m" and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] <= d["value"] * 1.02 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 1920 * 1080 * 3 * 4
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and isinstance(c["sample"], str)


def test_reference_arm_line():
    d = _line("r01_bench_c3_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["config"]["workload"] == "c3_icospheres_1m"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
