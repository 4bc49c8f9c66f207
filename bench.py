#!/usr/bin/env python
"""Headline benchmark: train it/s (render fwd + loss + bwd) and render Mpix/s at 1 M Gaussians,
1920x1080, SH degree 3, through the MsplatRender plugin (pointrix_b200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg1|cfg2|cfg3|cfg4|cfg5] [--mode train|render]
                    [--exchange factored|allreduce|nccl] [--no-cpu-baseline] [--no-ref-gpu] [--no-loss-leg]

One "step" (mode train) = one training iteration of the render path on one view per GPU:
MsplatRender.render_iter forward -> the fused 0.8 L1 + 0.2 (1 - SSIM) loss against that view's target
image (pointrix/model/base_model.py:113-124) -> backward into every Gaussian parameter
(+ intrinsics / extrinsics / camera centre for cfg5), and when N > 1 -- views sharded by rank -- the
exchange that leaves the batch gradients / ndc gradients / radii on every rank (the SH-factored exchange
over symmetric memory, parallel.ShFactoredExchange).  No eager-PyTorch kernel runs
inside the step.  `--mode render` (the cfg4 inference sweep): one step = one forward-only render of
rgb + depth + normal + flow (C = 9); the metric is render Mpix/s.

Timed regions: (1) `value`: K device-resident steps, CUDA events, barrier + synchronize on both sides,
max over ranks; only the dominant kernel is bracketed by events inside it (`roofline`), a second pass
of K steps times every stage (`roofline_stages`).  (2) `render_mpix_s`: K forward-only views.
(3) `e2e`: K steps through the plugin with the view's camera and target image coming from pinned HOST
memory and the loss read back on the host every step.  (4) rank 0, after the process group is gone:
`ref_gpu` (the compiled reference CUDA, oracle/_ref, on this GPU, every N), `cpu_baseline` (N = 1: the CPU
oracle on one full view), `photometric_loss` (N = 1).  Every side leg is guarded: a failure is
recorded as {"unavailable": ...} and never costs the line.

`--impl reference` times the CPU-PyTorch restatement of the same math (oracle/) on the host cores on
the FULL cloud of the same config, for as many iterations as fit a few minutes (the line reports the
`steps`/`warmup` it actually ran); the reference has no CPU implementation of its own.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT_TRAIN = "view-iterations/s"
UNIT_RENDER = "Mpix/s"
LAMBDA_SSIM = 0.2


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_scene_module():
    """pointrix_b200/scene.py loaded by path: the synthetic-scene generator is pure torch, and the
    `--impl reference` arm must not map libpointrix_b200.so (importing the package would)."""
    spec = importlib.util.spec_from_file_location("pxb_scene_standalone", os.path.join(ROOT, "pointrix_b200", "scene.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def metric_name(cfg: str, mode: str, c: dict) -> str:
    scale = {"cfg1": "10K", "cfg2": "300K", "cfg3": "3M", "cfg4": "1M", "cfg5": "1M"}[cfg]
    res = "1080p" if (c["W"], c["H"]) == (1920, 1080) else f"{c['W']}x{c['H']}"
    if mode == "render":
        return f"render Mpix/s (fwd), {scale} Gaussians @{res}"
    return f"train it/s (fwd+bwd), {scale} Gaussians @{res}"


def workload_name(cfg: str, mode: str, c: dict) -> str:
    base = f"{cfg}: {c['P']} Gaussians, SH degree 3, {c['W']}x{c['H']}, white bg, {c['views']} orbit cameras; "
    if mode == "render":
        return base + ("MsplatRender.render_iter forward only, C=9 channels (rgb + depth + normals[P,3] + flow[P,2] "
                       "as extra feature kwargs), one view per GPU per step")
    cam = " + intrinsics/extrinsics/camera centre" if cfg == "cfg5" else ""
    return base + (f"MsplatRender.render_iter fwd (C=3 rgb) -> fused 0.8 L1 + 0.2 (1-SSIM) loss vs a fixed random target "
                   f"-> bwd into all Gaussian parameters{cam} (one view per GPU per step)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML every 10 ms
    (two cheap queries; an `nvidia-smi -lms 20` child process, used before, cost the step 8 %: each of its
    polls takes driver locks the launch path also needs).  Falls back to nvidia-smi at 100 ms without NVML."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int, interval_s: float = 0.010):
        self.idx = gpu_index
        self.interval = interval_s
        self.samples = []  # (t, sm_mhz, reason_bits)
        self.max_mhz = None
        self.t0 = self.t1 = None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None
        self.proc = None
        self.lines = []

    def _handle(self):
        import pynvml as N

        N.nvmlInit()
        try:
            import torch

            u = str(torch.cuda.get_device_properties(self.idx).uuid)
            h = N.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
        except Exception:
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        return N, h, reasons

    def start(self):
        try:
            N, h, reasons = self._handle()
            self.nvml = N

            mode = os.environ.get("PXB_BENCH_SAMPLER", "1")

            def loop():
                while not self._stop.is_set():
                    try:
                        t0 = time.perf_counter()
                        clk = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                        t1 = time.perf_counter()
                        rs = int(reasons(h)) if mode != "2" else 0
                        t2 = time.perf_counter()
                        self.samples.append((t0, clk, rs))
                        self.query_ms = max(getattr(self, "query_ms", (0.0, 0.0)), (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
                    except Exception:
                        pass
                    self._stop.wait(self.interval)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
                self.thread = threading.Thread(target=self._pump, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def wait_first(self, timeout):
        t0 = time.time()
        while (self.nvml or self.proc) and not (self.samples or self.lines) and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Start of the timed region."""
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.t1 is None:
            self.mark_end()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(1.0)
            rows = self.samples
        else:
            time.sleep(0.12)
            self.proc.terminate()
            rows = []
            for t, ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    bits = sum(b for b, v in zip((0x8, 0x40, 0x20, 0x4), f[5:9]) if v.lower().startswith("active"))
                    rows.append((t, float(f[1]), bits))
                    self.max_mhz = float(f[2])
                except (ValueError, IndexError):
                    continue
        inside = [r_ for r_ in rows if self.t0 <= r_[0] <= self.t1]
        if not inside and rows:  # region shorter than the interval: the samples bracketing it
            inside = sorted(rows, key=lambda r_: abs(r_[0] - 0.5 * (self.t0 + self.t1)))[:2]
        bits = 0
        for r_ in inside:
            bits |= r_[2]
        sm = [r_[1] for r_ in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.BITS.items() if bits & b), "samples": len(sm),
                "source": "nvml, every %d ms" % round(1e3 * self.interval) if self.nvml is not None else "nvidia-smi -lms 100"}


def guarded(fn, *a, **kw):
    """Side legs never cost the line: a failure becomes {"unavailable": ...}."""
    try:
        return fn(*a, **kw)
    except BaseException as e:  # noqa: BLE001
        if isinstance(e, KeyboardInterrupt):
            raise
        traceback.print_exc(file=sys.stderr)
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


# ---------------------------------------------------------------------------------------------------
# CPU oracle (the `--impl reference` arm and the N = 1 `cpu_baseline` leg)
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_run(cfg_name: str, mode: str, n_timed: int, n_warm: int, budget_s: float, sample_P: int | None = None,
                   tile_stride: int | None = None):
    """CPU-PyTorch execution of the same projection/SH/sort/blend (+ loss) math on the host cores: the full
    cloud of the config (unless `sample_P`), every tile of the view (unless `tile_stride`: only the tiles with
    (tx + 3 ty) % stride == 0 are blended, stride 0 = none -- everything per-Gaussian still runs in full).
    Stops early (never below one timed iteration) when `budget_s` of wall clock is used up; reports what ran."""
    import torch

    from oracle import loss_oracle as LO
    from oracle import msplat_oracle as O

    scene = load_scene_module()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c, sc, cams = scene.make_config(cfg_name, P=sample_P, views=max(1, min(4, n_timed + n_warm)))
    H, W = c["H"], c["W"]
    V = cams["extrinsic_matrix"].shape[0]
    target = scene.target_images(1, H, W)[0]
    cam_grads = cfg_name == "cfg5" and mode == "train"
    extras = scene.extra_features(c["P"]) if mode == "render" else {}
    tile_mask = None
    if tile_stride is not None:
        gx, gy = (W + 15) // 16, (H + 15) // 16
        t = torch.arange(gx * gy)
        tile_mask = ((t % gx + 3 * (t // gx)) % tile_stride == 0) if tile_stride > 0 else torch.zeros(gx * gy, dtype=torch.bool)
    times, t_begin, warm_done, isect = [], time.perf_counter(), 0, (0, 0)
    for it in range(n_warm + n_timed):
        v = it % V
        E = cams["extrinsic_matrix"][v].clone()
        I = cams["intrinsic_params"].clone()
        Cc = cams["camera_center"][v].clone()
        t0 = time.perf_counter()
        if mode == "render":
            with torch.no_grad():
                o = O.render_iter(H, W, E, I, Cc, **sc, sh_degree=3, render_depth=True, extra_features=extras,
                                  tile_mask=tile_mask)
            per_tile = (o["_aux"]["tile_range"][:, 1] - o["_aux"]["tile_range"][:, 0]).long()
            isect = (int(per_tile.sum()), int(per_tile.sum() if tile_mask is None else per_tile[tile_mask].sum()))
        else:
            leaves = {k: t.clone().requires_grad_() for k, t in sc.items()}
            if cam_grads:
                E.requires_grad_(), I.requires_grad_(), Cc.requires_grad_()
            # same math as loss(render_iter(...)).backward(), blend differentiated chunk by chunk (bounded memory)
            o = O.render_step(H, W, E, I, Cc, **leaves, sh_degree=3, tile_mask=tile_mask,
                              loss_fn=lambda img: LO.l1_ssim_loss(img.unsqueeze(0), target.unsqueeze(0), LAMBDA_SSIM)["loss"])
            isect = (o["isect_total"], o["isect_blended"])
        dt = time.perf_counter() - t0
        if it >= n_warm:
            times.append(dt)
        else:
            warm_done += 1
        if times and time.perf_counter() - t_begin + dt > budget_s:
            break
    t_iter = sum(times) / len(times)
    full_P = scene.CONFIGS[cfg_name]["P"]
    unit = UNIT_RENDER if mode == "render" else UNIT_TRAIN
    what = "forward only, C=9" if mode == "render" else "fwd + L1/SSIM loss + bwd" + (" incl. camera gradients" if cam_grads else "")
    n_P = min(sample_P or full_P, full_P)
    sample = (f"{'the full' if n_P == full_P else 'the first ' + str(n_P) + ' Gaussians of the'} {cfg_name} cloud ({full_P} Gaussians) "
              f"at {W}x{H}, {what}: {len(times)} timed view(s) after {warm_done} warm-up, {t_iter:.2f} s/view")
    return {"value": per_view_value(mode, W, H, t_iter), "unit": unit, "cores": cores, "kind": "port", "sample": sample,
            "s_per_view": t_iter, "sample_P": n_P, "timed_iterations": len(times), "warmup_iterations": warm_done,
            "isect_total": isect[0], "isect_blended": isect[1]}


def per_view_value(mode, W, H, s_per_view):
    return (W * H / 1e6 / s_per_view) if mode == "render" else 1.0 / s_per_view


def cpu_tile_sample_estimate(cfg_name: str, mode: str, stride: int = 8):
    """Bounded sample of one full-size view: the FULL cloud goes through every per-Gaussian stage, the sort and
    the loss, but only every `stride`-th tile of the image is blended (forward and backward).  Blending is
    the part of the CPU cost that scales with the view (it is > 90 % of it), and it scales with the tile
    intersections processed, so   t_full = t_rest + (t_sample - t_rest) * N_all / N_sampled,   with t_rest
    measured by a run that blends no tile at all.  (Sub-sampling the CLOUD instead is not used: the CPU cost
    is far from linear in the Gaussian count, measured 39.7 s estimated vs 94.8 s actual at cfg4.)"""
    r0 = cpu_oracle_run(cfg_name, mode, 1, 1, 1e9, tile_stride=0)  # its warm-up absorbs the cold start
    r1 = cpu_oracle_run(cfg_name, mode, 1, 0, 1e9, tile_stride=stride)
    t_rest, t_sub = r0["s_per_view"], r1["s_per_view"]
    k = r1["isect_total"] / max(r1["isect_blended"], 1)
    est = t_rest + max(t_sub - t_rest, 0.0) * k
    scene = load_scene_module()
    c = scene.CONFIGS[cfg_name]
    r = dict(r1)
    r.update({"value": per_view_value(mode, c["W"], c["H"], est), "s_per_view": est, "estimated": True,
              "sample": (f"one view of the full {cfg_name} cloud ({c['P']} Gaussians) at {c['W']}x{c['H']}; every per-Gaussian stage, "
                         f"the sort and the loss run in full, the blend (fwd" + ("" if mode == "render" else " + bwd") + f") over every {stride}th tile "
                         f"({r1['isect_blended']} of {r1['isect_total']} tile intersections): {t_sub:.2f} s; the same with no "
                         f"tile blended: {t_rest:.2f} s; scaled by the intersection ratio => {est:.1f} s/view"),
              "timed_iterations": 1, "warmup_iterations": 1})
    return r


def run_reference(args, out_f):
    if env_int("RANK", 0) != 0:
        return
    scene = load_scene_module()
    c = scene.CONFIGS[args.config]
    P = c["P"]
    # Bounded: a full-size view costs the host minutes.  One sampled iteration (the full cloud, every 8th tile
    # blended; it doubles as the warm-up) estimates a full iteration; if that fits the wall-clock budget ONE
    # timed iteration runs on the full view and is what the line reports; otherwise the line reports the
    # estimate (and says so).  `steps` / `warmup` are what actually ran.
    if args.cpu_sample:  # developer / test override: one fixed sample of the cloud, linear in P
        r = cpu_oracle_run(args.config, args.mode, max(1, min(args.steps, 2)), min(args.warmup, 1), args.cpu_budget,
                           sample_P=args.cpu_sample)
        k = P / r["sample_P"]
        r["value"], r["s_per_view"] = r["value"] / k, r["s_per_view"] * k
        r["sample"] += f", linearly extrapolated x{k:.1f} to the full cloud"
        warm_run = r["warmup_iterations"]
    else:
        est = cpu_tile_sample_estimate(args.config, args.mode)
        warm_run = 2
        if est["s_per_view"] <= args.cpu_budget:
            r = cpu_oracle_run(args.config, args.mode, 1, 0, args.cpu_budget)
            r["sample"] += f", no extrapolation; warm-up = the sampled iterations (estimate {est['s_per_view']:.1f} s/view)"
        else:
            r, warm_run = est, 1
            r["sample"] += f" -- above the {args.cpu_budget:.0f} s budget, so the full view was not run"
    line = {"metric": metric_name(args.config, args.mode, c), "value": r["value"], "unit": r["unit"], "n_gpus": args.gpus,
            "steps": r["timed_iterations"], "warmup": warm_run, "steps_requested": args.steps,
            "warmup_requested": args.warmup, "ms_per_step": 1e3 * r["s_per_view"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.config, args.mode, dict(c)), "parallelism": "cpu (rank 0 only)",
                       "note": "the reference has no CPU implementation: this is the oracle port on host cores, one view per step"},
            "cpu_baseline": r, "gpu_launches": 0,
            "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out_f)


def _claim_stdout():
    """The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner)
    are sent to stderr; the returned file object is the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_WATCHDOG = {"line": None, "out": None}


def _start_watchdog(seconds: float):
    """A hang (a collective whose peer died, a spinning barrier) must not cost the caller its whole time limit: after
    `seconds` the process prints the line if the headline has been measured (rank 0) and exits hard."""
    def fire():
        time.sleep(seconds)
        line, out = _WATCHDOG["line"], _WATCHDOG["out"]
        try:
            if line is not None and out is not None and env_int("RANK", 0) == 0:
                line.setdefault("errors", []).append(f"watchdog: not finished after {seconds:.0f} s; legs after the headline are missing")
                print(json.dumps(line), file=out)
                out.flush()
        finally:
            os._exit(0 if line is not None else 3)

    threading.Thread(target=fire, daemon=True).start()


def main():
    out_f = _claim_stdout()
    _WATCHDOG["out"] = out_f
    try:
        _main(out_f)
    finally:
        out_f.flush()


# ---------------------------------------------------------------------------------------------------
# algorithmic bytes (SURVEY.md 8d / DESIGN.md section 4)
# ---------------------------------------------------------------------------------------------------
def algorithmic_bytes(P, N, H, W, C, S, tiles, world, factored):
    passes = max(1, ((tiles - 1).bit_length() + 7) // 8)
    fused_bwd = P * (4 * S + 12 + 12 + 16 + 192 + 4 + 4 + 12 + 12 + 16 + 4 + 8 + (12 if factored else 192))
    return {
        # read xyz 12 + scale 12 + quat 16 + opacity 4 + SH 192; write the S-float record + depth, radius, tiles
        "pxb_fused_forward": P * (12 + 12 + 16 + 4 + 192 + 4 * S + 4 + 4 + 4),
        # depth-order sort of the visible (key,id) pairs: compaction 12, passes x 16 B, gathered scan 16 B
        "pxb_bin_prepare": P * (12 + 4 + 3 * 16 + 16),
        # emit (28 B/Gaussian read, 8 B/isect written), histogram 4, tile passes x 16 B, ranges 4 B/isect + 8 B/tile
        "pxb_sort_gaussian": P * 28 + N * (8 + 4 + 16 * passes + 4) + tiles * 8,
        # (28 + 4C) B/isect gathered + 4 (C+2) HW written
        "pxb_blend_forward": N * (28 + 4 * C) + 4 * (C + 2) * H * W,
        # same gather + (4C + 8) HW read + ideal atomics 8 (6+C) B/isect
        "pxb_blend_backward": N * (28 + 4 * C) + (4 * C + 8) * H * W + N * 8 * (6 + C),
        "pxb_fused_backward": fused_bwd,
    }


def _main(out_f):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--mode", default="train", choices=["train", "render"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="reference arm: first n Gaussians only (0 = the full cloud)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="reference arm: wall-clock budget in seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-loss-leg", action="store_true")
    ap.add_argument("--exchange", default="factored", choices=["factored", "allreduce", "nccl"])
    ap.add_argument("--max-seconds", type=float, default=900.0, help="watchdog: hard exit after this many seconds")
    ap.add_argument("--train-loop", type=int, default=0,
                    help="N = 1 side leg: run this many iterations of the full training loop (raw parameters, fused Adam + "
                         "statistics, densification) of the config and report it")
    args = ap.parse_args()
    _start_watchdog(args.max_seconds)
    if args.impl == "reference":
        return run_reference(args, out_f)

    import torch
    import torch.distributed as dist

    import pointrix_b200 as pb
    from pointrix_b200 import _lib, parallel, scene
    from pointrix_b200 import loss as PL

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_, K = max(args.warmup, 3), max(args.steps, 1)
    mode, cfg = args.mode, args.config
    train = mode == "train"

    c, sc, cams = scene.make_config(cfg)
    H, W, P, V = c["H"], c["W"], c["P"], c["views"]
    cam_grads = train and cfg == "cfg5"
    params = {k: v.to(dev).requires_grad_(train) for k, v in sc.items()}
    cams_d = {k: v.to(dev) for k, v in cams.items()}
    extras = {k: v.to(dev) for k, v in scene.extra_features(P).items()} if not train else {}
    # the plugin's explicit extra-feature arguments: `normals` (examples/supervise/renderer.py) + a dict
    extra_kw = {"normals": extras["normals"], "extra_features": {"flow": extras["flow"]}} if extras else {}
    NT = min(V, 4)  # distinct target images (25 MB each at 1080p)
    targets_h = scene.target_images(NT, H, W)
    targets = targets_h.to(dev) if train else None
    r = pb.parse_renderer({"name": "MsplatRender", "render_depth": not train}, white_bg=True, device=str(dev))
    r.sh_degree = 3
    Cch = 3 if train else 3 + 1 + sum(v.shape[1] for v in extras.values())
    S = _lib.lib.pxb_record_stride(Cch)
    # the batch loss is a mean over the views of all ranks (pointrix/model/loss.py:27-46): every view's loss
    # enters with weight 1/world -- handed to backward() as a device scalar, no scaling kernel
    loss_w = torch.full((), 1.0 / world, dtype=torch.float32, device=dev)

    exch, exch_kind = None, "none"
    if world > 1 and train:
        from pointrix_b200 import renderer as _renderer_mod

        try:
            if args.exchange == "nccl":
                raise RuntimeError("NCCL requested")
            if args.exchange == "factored":
                exch = parallel.ShFactoredExchange(P, dev)
                exch_kind = (f"SH gradient rebuilt from every rank's dL/drgb read over NVLink (pxb_sh_grad_gather), "
                             f"13 floats/Gaussian + radii all-reduced in symmetric memory ({exch.mode})")
            else:
                exch = parallel.NvlsGradExchange(P, dev)
                exch_kind = f"pxb_{exch.mode}_allreduce of 61 floats/Gaussian + radii over symmetric memory"
            _renderer_mod.set_grad_sink(exch)
        except Exception as e:  # noqa: BLE001
            exch, exch_kind = None, f"NCCL all-reduce ({type(e).__name__}: {str(e)[:80]})"

    def exchange(out, rw):
        # camera gradients (cfg5) stay local: every rank optimises the camera of its own view
        if isinstance(exch, parallel.ShFactoredExchange):
            exch.exchange(out["radii"], params["position"])
        elif exch is not None:
            exch.exchange(out["radii"])
        else:
            parallel.allreduce_step([p_.grad for p_ in params.values()], out["uv_points"].grad, out["radii"], world,
                                    average=False, radii_work=rw)

    def view_of(step):  # views sharded by rank
        return (step * world + rank) % V

    def camera(cam_src, v):
        E, I, Cc = cam_src["extrinsic_matrix"][v], cam_src["intrinsic_params"], cam_src["camera_center"][v]
        if cam_grads:
            E, I, Cc = (t.detach().clone().requires_grad_() for t in (E, I, Cc))
        return E, I, Cc

    def step_fn(step, cam_src=cams_d, tgt=None):
        v = view_of(step)
        E, I, Cc = camera(cam_src, v)
        if not train:
            with torch.no_grad():
                out = r.render_iter(H, W, E, I, Cc, **params, **extra_kw)
            return None, out
        for p_ in params.values():
            p_.grad = None
        out = r.render_iter(H, W, E, I, Cc, **params)
        # NCCL path only: radii are final after the forward, their reduce hides under the backward
        rw = parallel.begin_radii_reduce(out["radii"], world) if (exch is None and world > 1) else None
        img = out["rendered_features_split"]["rgb"]
        loss = PL.l1_ssim_loss(img.unsqueeze(0), (targets[v % NT] if tgt is None else tgt).unsqueeze(0), LAMBDA_SSIM)["loss"]
        loss.backward(loss_w)
        if world > 1:
            exchange(out, rw)
        return loss, out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed, before the warm-up): one forward per view sizes the intersection buffers, as the
    # first iterations of a training run do; without it the first visit of a heavier view inside the timed
    # region re-runs its binning with a larger capacity (and pays a cudaMalloc)
    with torch.no_grad():
        for v in range(V):
            r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v],
                          **params, **extra_kw)
    torch.cuda.synchronize()

    # ---- device-resident timed region --------------------------------------------------
    sampler = ClockSampler(local)
    DEBUG_MARKS = os.environ.get("PXB_BENCH_STEP_MARKS") == "1"   # developer knobs (timeline of the timed pass)
    if rank == 0 and os.environ.get("PXB_BENCH_SAMPLER", "1") != "0":
        sampler.start()  # before the warm-up: the sampler's own start-up must not overlap the timed region
    # Untimed settling phase before the W warm-up steps: the process has spent seconds on the host building the
    # scene, the GPU has idled meanwhile, and on a fresh box the first few hundred milliseconds of steps also
    # page code in.  So the step runs continuously for PREWARM_S seconds of wall clock first (every rank the same
    # number of steps), with exactly the allocation pattern of the timed loop.
    PREWARM_S = float(os.environ.get("PXB_BENCH_PREWARM_S", "0.6"))
    n_pre = 0
    t_pre = time.perf_counter()
    while PREWARM_S > 0:
        for _ in range(16):
            # same object lifetimes as the timed loop below (the previous step's outputs stay referenced while the
            # next step runs): the caching allocator must reach THAT steady state here.  Discarding the result
            # instead let the first timed steps call cudaMalloc (measured: one 10-250 ms stall in timed step 2)
            loss, out = step_fn(n_pre)
            n_pre += 1
        stop = torch.tensor([1.0 if time.perf_counter() - t_pre >= PREWARM_S else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(stop, op=dist.ReduceOp.MAX)  # every rank leaves after the same number of steps
        if stop.item() > 0:
            break
    for s in range(W_):
        loss, out = step_fn(s)
    sync()
    if rank == 0:
        sampler.wait_first(3.0)
        sampler.mark()
    # inside the timed region only the dominant kernel is bracketed by events (2 records per step); every
    # stage is timed in a second pass of K steps right after it (8 records per step cost ~3 % of the step)
    DOMINANT = "pxb_blend_backward" if train else "pxb_blend_forward"
    timer = _lib.KernelTimer(stages={DOMINANT})
    _lib.set_timer(timer)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    marks, host_t = [], []
    e0.record()
    for s in range(K):
        loss, out = step_fn(W_ + s)
        if DEBUG_MARKS:
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
            host_t.append(time.perf_counter())
    e1.record()
    sync()
    if DEBUG_MARKS and rank == 0:
        print("timed pass: device ms since start per step:", [round(e0.elapsed_time(m), 3) for m in marks], file=sys.stderr)
        print("timed pass: host ms between step returns:", [round(1e3 * (b - a), 3) for a, b in zip(host_t, host_t[1:])], file=sys.stderr)
        print("sampler: slowest (clock, reasons) query ms:", getattr(sampler, "query_ms", None), file=sys.stderr)
    if rank == 0:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    _lib.set_timer(None)
    clocks = sampler.stop() if rank == 0 else None
    dom_timed = timer.summary().get(DOMINANT)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    per_step_units = (W * H / 1e6) if not train else 1.0
    value = world * K * per_step_units / (ms / 1e3)
    unit = UNIT_TRAIN if train else UNIT_RENDER

    line = {
        "metric": metric_name(cfg, mode, c), "value": round(value, 3), "unit": unit, "n_gpus": world, "steps": K,
        "warmup": W_, "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, mode, c),
                   "parallelism": f"view-sharded dp{world}" + (f", gradients of 59+2 floats/Gaussian summed + max(radii) per step: {exch_kind}" if (world > 1 and train) else (", no collective (inference)" if world > 1 else "")),
                   "cache": "inputs larger than L2 (236 MB Gaussian table + records vs 126 MB L2); a different view every step"},
        "gpu_launches": launches, "clocks": clocks, "untimed_settling_steps": n_pre,
    }

    _WATCHDOG["line"] = line  # from here on a hang still yields the headline

    def finish():
        _WATCHDOG["line"] = None
        print(json.dumps(line), file=out_f)
        out_f.flush()

    try:
        _rest(args, line, locals())
    except BaseException as e:  # noqa: BLE001 -- the headline is measured: print it whatever happens next
        traceback.print_exc(file=sys.stderr)
        line.setdefault("errors", []).append(f"{type(e).__name__}: {str(e)[:300]}")
        if world > 1 and dist.is_initialized():
            try:
                dist.destroy_process_group()
            except Exception:  # noqa: BLE001
                pass
    if rank == 0:
        finish()


def _rest(args, line, L):
    """Everything after the headline timed region: stage pass, forward-only leg, e2e leg, roofline, side legs.
    Any exception here is recorded on the line, never fatal to it."""
    import torch
    import torch.distributed as dist

    from pointrix_b200 import _lib, ops, parallel
    from pointrix_b200 import loss as PL

    (rank, world, dev, K, W_, H, W, P, V, NT, S, Cch, train, cfg, mode, cam_grads, c, sc, cams, cams_d, params, extra_kw,
     targets_h, targets, r, exch, step_fn, view_of, camera, sync, ms, dom_timed, DOMINANT, clocks, loss_w, exchange) = (
        L[k] for k in ("rank", "world", "dev", "K", "W_", "H", "W", "P", "V", "NT", "S", "Cch", "train", "cfg", "mode",
                       "cam_grads", "c", "sc", "cams", "cams_d", "params", "extra_kw", "targets_h", "targets", "r", "exch",
                       "step_fn", "view_of", "camera", "sync", "ms", "dom_timed", "DOMINANT", "clocks", "loss_w",
                       "exchange"))

    # ---- second pass: every stage bracketed by CUDA events ---------------------------------
    stage_timer = _lib.KernelTimer()
    _lib.set_timer(stage_timer)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    s0.record()
    step_marks[0].record()
    for s in range(K):
        step_fn(W_ + K + s)
        step_marks[s + 1].record()
    s1.record()
    sync()
    _lib.set_timer(None)
    ms_stage_pass = s0.elapsed_time(s1)
    per = sorted(step_marks[i].elapsed_time(step_marks[i + 1]) for i in range(K))
    pick = lambda q: round(per[min(K - 1, max(0, int(round(q * (K - 1)))))], 4)  # noqa: E731
    line["step_ms_distribution"] = {"p10": pick(0.10), "median": pick(0.50), "p90": pick(0.90), "min": round(per[0], 4),
                                    "max": round(per[-1], 4), "pass": "stage-timing pass (this rank)"}
    kern = stage_timer.summary()

    # ---- forward-only render Mpix/s (second half of the BASELINE metric) -----------------
    with torch.no_grad():
        def fwd(s):
            v = view_of(s)
            r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v],
                          **params, **extra_kw)
        for s in range(3):
            fwd(s)
        sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for s in range(K):
            fwd(s)
        f1.record()
        sync()
        tf = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        line["render_mpix_s"] = round(world * K * W * H / 1e6 / (tf.item() / 1e3), 2)
        line["render_channels"] = Cch

    # ---- end to end through the plugin with HOST buffers --------------------------------
    # per step: pinned host -> device copy of that view's camera (extrinsic 4x4, intrinsics 4, centre 3) and,
    # when training, of its target image [3,H,W]; device -> host read of the step's result (the loss; for the
    # inference sweep the rendered [C,H,W] maps themselves).
    host_cam = torch.cat([cams["extrinsic_matrix"].reshape(V, 16), cams["intrinsic_params"].reshape(1, 4).expand(V, 4),
                          cams["camera_center"].reshape(V, 3)], dim=1).contiguous().pin_memory()
    host_tgt = targets_h.pin_memory() if train else None
    h2d = (16 + 4 + 3) * 4 + (3 * H * W * 4 if train else 0)
    copy_stream = torch.cuda.Stream(device=dev)
    if train:
        # the loss of step k lands in pinned slot k % 2 and is consumed by the host while step k+1 is queued
        # (a trainer logs the previous iteration's loss): every step's result is read inside the timed region,
        # but the host never drains the GPU between steps
        res_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        d2h = 4
    else:
        res_host = torch.empty(2, Cch, H, W, dtype=torch.float32).pin_memory()
        d2h = Cch * H * W * 4
    res_done = [torch.cuda.Event(), torch.cuda.Event()]
    results = []

    def e2e_collect(step):
        """Host read of the result of `step` (blocks until its D2H copy has landed)."""
        res_done[step % 2].synchronize()
        results.append(float(res_host[step % 2].reshape(-1)[0]))

    def e2e_step(step):
        v = view_of(step)
        main = torch.cuda.current_stream(dev)
        # the camera is needed first (small, on the compute stream); the 25 MB target upload runs on a copy
        # stream underneath the forward render and is joined just before the loss needs it
        cam = host_cam[v].to(dev, non_blocking=True)
        cam_src = {"extrinsic_matrix": cam[0:16].view(1, 4, 4), "intrinsic_params": cam[16:20], "camera_center": cam[20:23].view(1, 3)}
        E, I, Cc = camera(cam_src, 0)
        if not train:
            with torch.no_grad():
                out = r.render_iter(H, W, E, I, Cc, **params, **extra_kw)
            feats = out["rendered_features_split"]
            base = feats["rgb"]._base if feats["rgb"]._base is not None else feats["rgb"]
            res_host[step % 2].copy_(base if base.shape[0] == Cch else torch.cat(list(feats.values())), non_blocking=True)
        else:
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                G = host_tgt[v % NT].to(dev, non_blocking=True)
            for p_ in params.values():
                p_.grad = None
            out = r.render_iter(H, W, E, I, Cc, **params)
            rw = parallel.begin_radii_reduce(out["radii"], world) if (exch is None and world > 1) else None
            img = out["rendered_features_split"]["rgb"]
            main.wait_stream(copy_stream)
            G.record_stream(main)
            loss = PL.l1_ssim_loss(img.unsqueeze(0), G.unsqueeze(0), LAMBDA_SSIM)["loss"]
            loss.backward(loss_w)
            if world > 1:
                exchange(out, rw)
            res_host[step % 2].copy_(loss.detach(), non_blocking=True)
        res_done[step % 2].record(main)
        if step > 0:
            e2e_collect(step - 1)

    for s in range(3):
        e2e_step(s)
    e2e_collect(2)
    sync()
    results.clear()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for s in range(K):
        e2e_step(3 + s)
    e2e_collect(3 + K - 1)  # the last step's result, still inside the timed region
    g1.record()
    sync()
    assert len(results) == K + 1 and all(math.isfinite(x) for x in results[1:]), "e2e: a step's result was not read"
    te = torch.tensor([g0.elapsed_time(g1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    per_step_units = (W * H / 1e6) if not train else 1.0
    line["e2e"] = {"value": round(world * K * per_step_units / (te.item() / 1e3), 3), "unit": line["unit"],
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "last_result": results[-1]}

    # ---- the collective part is over: every rank leaves the group together; rank 0 goes on alone ----
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline bookkeeping (algorithmic bytes per SURVEY.md 8d / DESIGN.md) -----------
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    def isect_counts():
        # N of the reference's lists (3-sigma squares) over the timed views; the fused path's own count
        # (the tiles the alpha >= 1/255 ellipse reaches) drives its traffic
        Ns, Nt = [], []
        with torch.no_grad():
            for s in range(min(K, V)):
                v = view_of(W_ + s)
                ex = cams_d["extrinsic_matrix"][v][:3].contiguous()
                uv, depth = ops.project_point(params["position"], cams_d["intrinsic_params"], ex, W, H, nearest=0.2)
                vis = (depth != 0).reshape(-1)
                cov = ops.compute_cov3d(params["scaling"], params["rotation"], vis)
                _, _, tl = ops.ewa_project(params["position"], cov, cams_d["intrinsic_params"], ex, uv, W, H, vis)
                Ns.append(int(tl.sum()))
                r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v],
                              **params, **extra_kw)
                Nt.append(int(ops.LAST_N.get((dev.index, True), Ns[-1])))
        return sum(Ns) / len(Ns), sum(Nt) / len(Nt)

    N_ref, N_mean = isect_counts()
    line["config"]["intersections_reference_lists"] = N_ref
    line["config"]["intersections_binned"] = N_mean
    alg = algorithmic_bytes(P, N_mean, H, W, Cch, S, tiles, world, isinstance(exch, parallel.ShFactoredExchange))
    stages = {}
    for name, st in kern.items():
        b = alg.get(name)
        stages[name] = {"ms_avg": round(st["ms_avg"], 4), "calls": st["calls"],
                        "share_of_step": round(st["ms_total"] / ms_stage_pass, 4)}
        if b:
            gbs = b / (st["ms_avg"] * 1e-3) / 1e9
            stages[name].update({"alg_bytes": int(b), "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)})
    line["roofline_stages"] = stages
    line["roofline_stages_note"] = (f"second pass of {K} steps with every stage bracketed by CUDA events "
                                    f"({round(ms_stage_pass / K, 4)} ms/step)")
    render_kernels = {k: v for k, v in kern.items() if k in alg}
    dom = max(render_kernels.items(), key=lambda kv: kv[1]["ms_total"])[0]
    dom_ms = dom_timed["ms_avg"] if (dom == DOMINANT and dom_timed) else kern[dom]["ms_avg"]
    dom_gbs = alg[dom] / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if cfg == "cfg4" and train and os.path.exists(tj):
        ent = json.load(open(tj)).get(dom, {})
        traffic = ent.get("dram_bytes_per_launch")
        traffic_src = "profiles/ncu_traffic.json: " + ent.get("source", "?") + ("; " + ent["note"] if "note" in ent else "")
    line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": round(dom_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(dom_gbs / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "alg_bytes_per_launch": int(alg[dom]),
                        "ms_avg_in_timed_region": round(dom_ms, 4), "share_of_step": round(dom_ms * K / ms, 4),
                        "limiter": "issue",
                        "note": "SURVEY.md 8d algorithmic bytes over the HBM copy peak, as the contract asks; the blend kernels "
                                "run out of L2 and are limited by FP32 issue slots + the shared-memory pipe (ncu: profiles/), "
                                "see blend_issue_roofline for the pair-test rate from device counters"}

    # blending as (pixel,Gaussian) pair tests against the FP32 issue bound: the pair tests are
    # COUNTED on the device (one 64-bit add per warp per launch, pxb_blend_counters) in a separate untimed pass
    def blend_issue():
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        pair_bound = 148 * 128 * sm_clk * 1e6 / 16.0
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib.pxb_blend_counters(ops._p(cnt)), "pxb_blend_counters")
        try:
            # LOCAL steps only: at N > 1 this runs on rank 0 after the process group is gone, so the step here is
            # render -> loss -> backward WITHOUT the exchange (the peers have left: an exchange would wait for them
            # forever -- it did, once) and with the gradient sink detached.  The blend kernels do not depend on it.
            from pointrix_b200 import renderer as _rmod

            _rmod.set_grad_sink(None)
            n = min(K, V)
            for s in range(n):
                v = view_of(W_ + s)
                E, I, Cc = camera(cams_d, v)
                if train:
                    for p_ in params.values():
                        p_.grad = None
                    o_ = r.render_iter(H, W, E, I, Cc, **params)
                    PL.l1_ssim_loss(o_["rendered_features_split"]["rgb"].unsqueeze(0), targets[v % NT].unsqueeze(0),
                                    LAMBDA_SSIM)["loss"].backward(loss_w)
                else:
                    with torch.no_grad():
                        r.render_iter(H, W, E, I, Cc, **params, **extra_kw)
            torch.cuda.synchronize()
        finally:
            _lib.check(_lib.lib.pxb_blend_counters(None), "pxb_blend_counters")
        cf, cb = (int(x) / n for x in cnt[:2].tolist())
        o = {}
        for name, cand in (("pxb_blend_forward", cf), ("pxb_blend_backward", cb)):
            if name in kern and cand > 0:
                pps = 32.0 * cand / (kern[name]["ms_avg"] * 1e-3)
                o[name] = {"warp_candidates_per_launch": cand, "pair_tests_per_launch": 32.0 * cand,
                           "pair_tests_per_s": pps, "bound_pairs_per_s": pair_bound, "frac": round(pps / pair_bound, 4)}
        o["note"] = ("pair tests = 32 lanes x (8x4 block, Gaussian) candidates the warps executed, counted on the device; "
                     "bound = 148 SMs x 128 lanes x SM clock / 16 FP32 instructions per pair (SURVEY.md 8d)")
        return o

    line["blend_issue_roofline"] = guarded(blend_issue)

    # ---- side legs on rank 0 (guarded) ---------------------------------------------------
    if not args.no_ref_gpu:
        line["ref_gpu"] = guarded(ref_gpu_run, cfg, mode, min(K, 10), dev)
    if world == 1 and train and not args.no_loss_leg:
        line["photometric_loss"] = guarded(photometric_loss_leg, H, W, K, dev)
    if world == 1 and train and args.train_loop > 0:
        line["training_loop"] = guarded(training_loop_leg, cfg, args.train_loop, dev)
    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only
        # bounded sample (tens of seconds of CPU work): the full cloud, every 8th tile blended, scaled by intersections
        line["cpu_baseline"] = guarded(cpu_tile_sample_estimate, cfg, mode, 8)


def training_loop_leg(cfg_name, iters, dev):
    """BASELINE config 2's call shape, every piece from this repo: the 3DGS training loop on the point cloud's RAW
    parameters -- camera_extrinsics -> render_iter_raw -> fused L1/SSIM loss -> backward -> GaussianAdam.step with
    the densification statistics in the same launch -> DensificationController.f_step (clone / split / prune /
    opacity reset on nerf.yaml's schedule) -> SH degree warm-up.  `--train-loop N` runs N iterations (the full
    schedule is 30 000; a few thousand are enough to cross the first densifications at 600, 700, ...) and reports
    iterations/s including everything, plus the table size it ends with."""
    import torch

    import pointrix_b200 as pb
    from pointrix_b200 import densify, optim, scene
    from pointrix_b200 import loss as PL

    c, sc, cams = scene.make_config(cfg_name)
    H, W, V = c["H"], c["W"], c["views"]
    op = sc["opacity"].clamp(1e-6, 1 - 1e-6)
    table = {"position": sc["position"], "features": sc["shs"][:, :1].contiguous(), "features_rest": sc["shs"][:, 1:].contiguous(),
             "scaling": torch.log(sc["scaling"]), "rotation": sc["rotation"], "opacity": torch.log(op / (1 - op))}
    params = {k: v.to(dev).contiguous().requires_grad_() for k, v in table.items()}
    cams = {k: v.to(dev) for k, v in cams.items()}
    targets = scene.target_images(min(V, 8), H, W).to(dev)
    r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
    opt = optim.GaussianAdam(params)
    stats = optim.DensificationStats(c["P"], dev, W, H)
    ctl = densify.DensificationController(opt, stats, cameras_extent=4.03)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_changed = 0
    for it in range(iters):
        r.update_sh_degree(it)
        v = it % V
        p = opt.params
        out = r.render_iter_raw(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v],
                                p["position"], p["opacity"], p["scaling"], p["rotation"], p["features"], p["features_rest"])
        PL.l1_ssim_loss(out["rendered_features_split"]["rgb"].unsqueeze(0), targets[v % targets.shape[0]].unsqueeze(0),
                        LAMBDA_SSIM)["loss"].backward()
        if ctl.wants_statistics():
            opt.update_model(stats=stats, uv_points=out["uv_points"], radii=out["radii"])
        else:
            opt.update_model()
        n_changed += int(ctl.f_step())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"iterations": iters, "it_per_s": round(iters / (ms / 1e3), 2), "ms_per_iteration": round(ms / iters, 4),
            "points_start": c["P"], "points_end": len(ctl), "densification_events": n_changed, "sh_degree_end": r.sh_degree,
            "note": "render_iter_raw + fused loss + backward + fused Adam/statistics + densification controller, one view per iteration"}


def photometric_loss_leg(H, W, K, dev):
    """SURVEY.md 8f row f1 beside the path: the fused (1-l)*L1 + l*(1-SSIM) loss (csrc/loss.cu) timed alone
    (fwd+bwd, CUDA events, 8 rotating image pairs = 400 MB > L2) and the same loss written with the reference's
    torch ops on the same GPU (oracle/loss_oracle.py = pointrix/model/loss.py's formulation: 5 cuDNN depthwise
    convolutions + elementwise kernels + autograd)."""
    import torch

    from oracle import loss_oracle as LO
    from pointrix_b200 import loss as PL

    g = torch.Generator(device=dev).manual_seed(2)
    gts = [torch.rand(1, 3, H, W, device=dev, generator=g) for _ in range(8)]
    preds = [(t + 0.1 * torch.randn(t.shape, device=dev, generator=g)).clamp(0, 1) for t in gts]

    def timed(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def ours(i):
        p = preds[i % 8].detach().requires_grad_()
        PL.l1_ssim_loss(p, gts[i % 8], LAMBDA_SSIM)["loss"].backward()

    def torch_ops(i):
        p = preds[i % 8].detach().requires_grad_()
        LO.l1_ssim_loss(p, gts[i % 8], LAMBDA_SSIM)["loss"].backward()

    n = max(10, min(K, 50))
    ms_ours, ms_ref = timed(ours, n), timed(torch_ops, n)
    px = 3 * H * W
    alg = px * (8 + 12) + px * (20 + 4)  # forward 8 read + 12 written, backward 20 read + 4 written per pixel-channel
    return {"fused_fwd_bwd_ms": round(ms_ours, 4), "torch_ops_fwd_bwd_ms": round(ms_ref, 4),
            "alg_bytes": alg, "achieved_gbs": round(alg / (ms_ours * 1e-3) / 1e9, 1),
            "note": "loss = 0.8*L1 + 0.2*(1-SSIM) at [1,3,H,W]; torch_ops = the reference's formulation (loss.py:27-117) on this GPU"}


def ref_gpu_run(cfg_name, mode, steps, dev):
    """The compiled unmodified reference CUDA (oracle/_ref) on the same GPU, same inputs, same step
    (render -> the reference's torch-op L1/SSIM loss -> its backward entry points): "the real bar" of
    BASELINE.md section 2.  Reported beside our number, never part of it."""
    import torch

    from oracle import loss_oracle as LO
    from oracle import ref_driver
    from pointrix_b200 import scene

    if not ref_driver.available():
        return {"unavailable": "oracle/_ref not built"}
    train = mode == "train"
    c, sc, cams = scene.make_config(cfg_name)
    H, W, V = c["H"], c["W"], c["views"]
    sc = {k: v.to(dev) for k, v in sc.items()}
    cams = {k: v.to(dev) for k, v in cams.items()}
    NT = min(V, 4)
    targets = scene.target_images(NT, H, W).to(dev)
    cam_grads = train and cfg_name == "cfg5"
    extra = torch.cat(list(scene.extra_features(c["P"]).values()), -1).to(dev) if not train else None

    def one(v, kind):
        f = ref_driver.render_forward(H, W, cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v],
                                      **sc, render_depth=not train, extra=extra)
        if kind == "fwd":
            return
        if kind == "fwd_loss_bwd":
            img = f["img"].detach().requires_grad_()
            LO.l1_ssim_loss(img.unsqueeze(0), targets[v % NT].unsqueeze(0), LAMBDA_SSIM)["loss"].backward()
            dimg = img.grad
        else:
            dimg = targets[v % NT]  # any fixed upstream gradient: render fwd + bwd alone
        ref_driver.render_backward(f, dimg, sc["position"], sc["opacity"], sc["scaling"], sc["rotation"], sc["shs"],
                                   cams["camera_center"][v], camera_grads=cam_grads)

    res = {}
    kinds = ("fwd_loss_bwd", "fwd_bwd", "fwd") if train else ("fwd",)
    for kind in kinds:
        for s in range(2):
            one(s % V, kind)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            one(s % V, kind)
        e1.record()
        torch.cuda.synchronize()
        res["ms_" + kind] = e0.elapsed_time(e1) / steps
    out = {"impl": "msplat reference CUDA (oracle/_ref: sm_100 build, -O3 --use_fast_math) driven in its own op order, "
                   "one view per step on one GPU", "steps": steps,
           "render_mpix_s": W * H / 1e6 / (res["ms_fwd"] / 1e3), **{k: round(v, 4) for k, v in res.items()}}
    if train:
        out["it_per_s"] = 1e3 / res["ms_fwd_loss_bwd"]           # the same step as `value`
        out["it_per_s_render_only"] = 1e3 / res["ms_fwd_bwd"]    # without the loss
    return out


if __name__ == "__main__":
    main()
