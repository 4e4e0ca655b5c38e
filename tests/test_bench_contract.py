"""The bench.py line format the driver depends on, checked on the committed lines (profiles/r01_bench_v13*.json are
verbatim stdout of `python bench.py` on 1 / 4 / 8 B200s) and on a live `--impl reference` run of one small step."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
            "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict,
            "gpu_launches": int, "clocks": dict, "roofline": dict}


@pytest.mark.parametrize("name,n", [("r02_bench_v12.json", 1), ("r02_bench_v11.json", 1), ("r02_bench_v10.json", 1), ("r02_bench_v10_n2.json", 2), ("r02_bench_v6.json", 1), ("r02_bench_v5_n2.json", 2), ("r02_bench_v5_n8.json", 8), ("r01_bench_v15.json", 1), ("r01_bench_v15_n8.json", 8), ("r01_bench_v16_n2.json", 2), ("r01_bench_v14.json", 1), ("r01_bench_v13.json", 1), ("r01_bench_v13_n2.json", 2), ("r01_bench_v13_n4.json", 4),
                                    ("r01_bench_v13_n8.json", 8)])
def test_committed_lines_follow_the_contract(name, n):
    d = json.load(open(os.path.join(REPO, "profiles", name)))
    for k, t in REQUIRED.items():
        assert isinstance(d[k], t), (k, type(d[k]))
    assert d["vs_baseline"] is None and d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True
    assert d["metric"] == json.load(open(os.path.join(REPO, "BASELINE.json")))["metric"].split(";")[0]
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert d["gpu_launches"] == d["steps"] * n            # one kernel per forward per GPU
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    if name.startswith("r02"):          # round 2 additions
        assert d["value_sustained"] > 0 and d["sustained"]["frac"] > 0.5
        assert "submit" in d["e2e"]["api"] and d["e2e"]["sustained_value"] > 0
        if n == 1:
            assert set(d["batch_sweep"]) >= {"1", "16", "128", "256"}
            c4 = d["config4"]
            assert c4["dtype"] == "bf16" and "configs[3]" in c4["workload"] and c4["roofline"]["traffic"] > 0
        else:
            x = d["nccl_exchange"]
            assert x["gather_bit_equal_on_rank0"] and x["weights_spot_check_identical_on_all_ranks"] and x["gather_records"] > 1000
    if "v10" in name or "v11" in name or "v12" in name:  # the end-of-round lines: self-play with the reference's Tier-1 gate at full depth
        t1 = d["selfplay"]["tier1_deep"]
        assert t1["sims_per_s"] > 0 and t1["sims_per_s_4x_games"] > t1["sims_per_s"] and t1["games_per_gpu_4x"] == 4 * t1["games_per_gpu"]
        assert 0.0 <= t1["game_steps_waiting_for_a_gate_job_fraction"] < 1.0
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]


def test_reference_arm_prints_one_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "64"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["unit"] == "evals/s"
