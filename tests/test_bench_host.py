"""CPU tests of bench.py's host logic: the clock sampler's window / throttle-reason bookkeeping and the
`--impl reference` arm's JSON contract (the arm that runs on host cores only)."""
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clock_sampler_reports_only_the_timed_region():
    import bench

    s = bench.ClockSampler(0)
    s.nvml, s.max_mhz = object(), 1965.0      # pretend NVML is up; no thread was started
    s.thread = threading.Thread(target=lambda: None)
    s.thread.start()
    s.samples = [(0.5, 300.0, 0), (1.1, 1965.0, 0), (1.2, 1950.0, 0x4), (1.3, 1965.0, 0), (2.5, 210.0, 0x8)]
    s.t0, s.t1 = 1.0, 2.0
    r = s.stop()
    assert r["samples"] == 3 and r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"]      # the hw_slowdown sample lies outside the region
    # a region shorter than the sampling interval: the bracketing samples are used
    s2 = bench.ClockSampler(0)
    s2.nvml, s2.max_mhz = object(), 1965.0
    s2.thread = threading.Thread(target=lambda: None)
    s2.thread.start()
    s2.samples = [(0.99, 1900.0, 0x20), (1.02, 1800.0, 0), (3.0, 100.0, 0x40)]
    s2.t0, s2.t1 = 1.0, 1.01
    r2 = s2.stop()
    assert r2["samples"] == 2 and r2["reasons"] == ["sw_thermal_slowdown"]
    # nothing available
    s3 = bench.ClockSampler(0)
    s3.t0 = 0.0
    assert s3.stop()["sm_mhz"] is None


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20",
                          "--warmup", "5", "--cpu-sample", "2000", "--config", "cfg1"], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                       # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["unit"] == "view-iterations/s" and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    # the arm reports the iterations it actually ran (bounded: <= 2 timed, <= 1 warm-up), not the request
    assert d["steps"] == d["cpu_baseline"]["timed_iterations"] <= 2 and d["warmup"] <= 2
    assert d["steps_requested"] == 20 and d["warmup_requested"] == 5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the other ranks of a torchrun launch exit 0 without work
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_reference_arm_never_maps_the_product_library():
    """`--impl reference` is the CPU oracle alone: libpointrix_b200.so must not be loaded by that process."""
    code = ("import sys; sys.path.insert(0, %r); import bench; "
            "r = bench.cpu_oracle_run('cfg1', 'train', 1, 0, 60.0, 500); assert r['value'] > 0; "
            "r2 = bench.cpu_oracle_run('cfg1', 'render', 1, 0, 60.0, 500); assert r2['unit'] == 'Mpix/s'; "
            "m = open('/proc/self/maps').read(); assert 'libpointrix_b200' not in m, 'product .so mapped'; "
            "assert 'pointrix_b200' not in sys.modules") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]


def test_product_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the product arm of bench.py must stop with a clear message, not run the
    oracle in its place."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("this box has a GPU")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, env=env, timeout=300)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "needs a GPU (no CPU fallback)" in out.stderr


def test_algorithmic_bytes_follow_the_survey_formulas():
    """blend forward (28 + 4C) B/isect + 4 (C+2) HW; blend backward the same gather + (4C + 8) HW + 8 (6+C)
    B/isect of ideal atomics; the factored exchange drops the 192-byte SH gradient write of the per-Gaussian backward."""
    import bench

    P, N, H, W, C, S, tiles = 1_000_000, 7_600_000, 1080, 1920, 3, 12, 8160
    a = bench.algorithmic_bytes(P, N, H, W, C, S, tiles, 1, False)
    assert a["pxb_blend_forward"] == N * 40 + 4 * 5 * H * W
    assert a["pxb_blend_backward"] == N * 40 + 20 * H * W + N * 72
    b = bench.algorithmic_bytes(P, N, H, W, C, S, tiles, 8, True)
    assert a["pxb_fused_backward"] - b["pxb_fused_backward"] == P * (192 - 12)
    assert bench.metric_name("cfg4", "train", {"W": 1920, "H": 1080}).startswith("train it/s (fwd+bwd), 1M Gaussians @1080p")
    assert bench.metric_name("cfg4", "render", {"W": 1920, "H": 1080}).startswith("render Mpix/s")
