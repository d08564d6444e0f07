"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the oracle port on the host
cores and prints ONE JSON line with the keys the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "train64",
                           "--steps", "1", "--warmup", "0", *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the arm states the batch its bounded sample really ran (it printed the workload's 64 while running 4),
    # keeps `config` identical to the B200 arm's (same builder), and times 1 warm-up + the requested steps
    assert d["sample_batch"] == d["cpu_baseline"]["batch"] == 4 and d["config"]["batch_per_gpu"] == 64
    assert abs(d["ms_per_step"] - d["cpu_baseline"]["ms_per_iteration"]) < 1e-6
    assert d["cpu_baseline"]["warmup_iterations"] == 1 and d["cpu_baseline"]["iterations"] == 1
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config("train64", 1, False)
    assert d["metric"] == bench.metric_name("train64")


def test_train4_reference_arm_runs_config_0_in_full():
    """BASELINE configs[0]: the 4x4 stage, one G+D step, batch 16 on the CPU — runs at its real batch."""
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "train4",
                        "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["sample_batch"] == 16 and d["config"]["batch_per_gpu"] == 16 and d["cpu_baseline"]["iterations"] == 3


def test_roofline_accounting_of_conv_calls():
    sys.path.insert(0, ROOT)
    import bench

    # 3x3 conv, N=2, 16x16, 32 -> 64: 2*N*H*W*9*Cin*Cout flops, 2*N*HW*(Cin+Cout) bytes
    fl, ex, by = bench.conv_call_work("bg_conv_fprop", (2, 16, 16, 32, 64, 3, 1, 0.2))
    assert fl == ex == 2 * 2 * 256 * 9 * 32 * 64 and by == 2 * 2 * 256 * (32 + 64)
    # folded conv+pool: credited with the 3x3 count at the INPUT resolution, executes 16/36 of it
    fl, ex, by = bench.conv_call_work("bg_conv_pool4_fprop", (2, 32, 32, 64, 64, 1, 0.2))
    assert fl == 2 * 2 * 1024 * 9 * 64 * 64 and abs(ex - fl * 16 / 36) < 1 and by == 2 * 2 * (64 * 1024 + 64 * 256)
    fl, ex, _ = bench.conv_call_work("bg_conv_pool4_dgrad", (2, 16, 16, 64, 64, 0.2))
    assert fl == 2 * 2 * 1024 * 9 * 64 * 64 and abs(ex - fl * 16 / 36) < 1
    rec = [("bg_conv_fprop", (2, 16, 16, 32, 64, 3, 1, 0.2), 0.01), ("bg_adain_apply", (2, 256, 64, 1e-8), 0.02)]
    peaks = {"hbm": 6000.0, "tf_sustained": 1400.0, "tf_burst": 1600.0, "src": "test"}
    roof, aux = bench.roofline_from_calls(rec, "train256", peaks, 100.0, 400.0, False)
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and abs(roof["frac"] - roof["achieved"] / 1400.0) < 1e-3
    assert roof["executed_frac"] == roof["frac"] and aux["worst"]["call"] == "bg_adain_apply"
    roof, _ = bench.roofline_from_calls(rec, "train512", peaks, 100.0, 500.0, False)
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] == 6000.0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
