"""CPU: the JSON line of ``bench.py --impl reference`` (the arm the driver runs beside ours) carries every key of the bench
contract, and ranks other than 0 stay silent.  The CPU sample is shrunk to one block on 24 x 24 window pixels so this runs in seconds;
the arithmetic timed is still the oracle port (oracle/rrdbnet_ref.py + oracle/wow_cv2.py)."""
import argparse
import json

import pytest

import bench


@pytest.fixture
def fast_sample(monkeypatch):
    orig = bench.cpu_sample
    monkeypatch.setattr(bench, "cpu_sample", lambda wl, **kw: orig(wl, blocks=1, n_windows=1, window_px=(24, 24), post_px=(64, 64)))


@pytest.mark.parametrize("workload,gpus", [("cfg2", 1), ("scene", 8), ("post4096", 1), ("cfg3", 1)])
def test_reference_arm_json_contract(fast_sample, capsys, monkeypatch, workload, gpus):
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=gpus, steps=2, warmup=1, workload=workload)
    bench.run_reference(args)
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["n_gpus"] == gpus and line["steps"] == 2 and line["warmup"] == 1
    assert line["metric"].startswith("output Mpix/s") and line["unit"] == "Mpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["dtype"] == "f32" and line["scaling"] in ("weak", "strong", "replicas") and isinstance(line["config"]["workload"], str)
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "Mpix/s" and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent(fast_sample, capsys, monkeypatch):
    monkeypatch.setenv("RANK", "3")
    bench.run_reference(argparse.Namespace(gpus=8, steps=1, warmup=0, workload="cfg2"))
    assert capsys.readouterr().out == ""


def test_workloads_name_the_baseline_configs():
    assert "64 windows of 532x532" in bench.workload("cfg2", 1)["label"]
    assert bench.workload("cfg2", 8)["H"] == 8 * 4096                       # weak scaling: 64 windows per GPU
    assert "1849 windows of 276x276" in bench.workload("scene", 8)["label"] and bench.workload("scene", 8)["H"] == 10980
    assert bench.FLOP_PER_LR_PX == 3456 + 69 * 479232 + 73728 + 294912 + 2 * 1179648 + 55296
    assert bench.workload("scene", 1)["scaling"] == "strong" and bench.workload("cfg2", 2)["scaling"] == "weak"
    assert bench.EDSR_FLOP_PER_LR_PX == 2 * 9 * (3 * 64 + 32 * 64 * 64 + 64 * 64 + 64 * 256) + 4 * (2 * 9 * 64 * 256) + 16 * (2 * 9 * 64 * 3)
    with pytest.raises(SystemExit):
        bench.workload("nope", 1)


def test_default_workload_is_the_north_star_scene(monkeypatch):
    import sys
    seen = {}
    monkeypatch.setattr(bench, "run_ours", lambda a: seen.update(vars(a)))
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    bench.main()
    assert seen["workload"] == "scene" and seen["gpus"] == 1 and seen["warmup"] >= 3


def test_clock_sampler_parses_nvidia_smi_lines():
    c = bench.ClockSampler(0)
    c.proc = type("P", (), {"terminate": lambda s: None, "wait": lambda s, timeout=None: None, "kill": lambda s: None})()
    c.lines = ["1372, 1965, 998.21, Not Active, Not Active, Not Active, Active", "1380, 1965, 1001.5, Not Active, Not Active, Not Active, Active",
               "1365, 1965, [N/A], Not Active, Active, Not Active, Not Active", "garbage"]
    got = c.stop()
    assert got == {"sm_mhz": 1372.0, "sm_max_mhz": 1965.0, "reasons": ["hw_thermal_slowdown", "sw_power_cap"], "samples": 3, "power_w": 1001.5}
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]
