#!/usr/bin/env python
"""Headline benchmark: train it/s (render fwd+bwd) and render Mpix/s at 1 M Gaussians,
1920x1080, SH degree 3, through the MsplatRender plugin (pointrix_b200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg4]
                    [--exchange factored|allreduce|nccl] [--no-cpu-baseline] [--no-ref-gpu] [--no-loss-leg]

One "step" = one training iteration of the render path on one view per GPU:
MsplatRender.render_iter forward, then backward of <dL/dimg, img> with a fixed
random dL/dimg into every Gaussian parameter (+ when N > 1, views sharded by rank, the
exchange that leaves the batch gradients / ndc gradients / radii on every rank: by default the
SH-factored exchange over symmetric memory, parallel.ShFactoredExchange).

Timed regions: (1) `value`: K device-resident steps, CUDA events, barrier + synchronize on both
sides, max over ranks; only the dominant kernel is bracketed by events inside it (`roofline`),
a second pass of K steps times every stage (`roofline_stages`).  (2) `render_mpix_s`: K forward-only
views.  (3) `e2e`: K steps with the view's camera and dL/dimg coming from pinned HOST memory and
the loss read back on the host every step.  (4) `photometric_loss` (N = 1): the fused L1+SSIM loss
beside the reference's torch-op formulation, and render -> loss -> backward.  (5) `cpu_baseline`,
`ref_gpu` (N = 1): the CPU oracle on a bounded sample and the compiled reference CUDA on this GPU.

`--impl reference` times the CPU-PyTorch restatement of the same math (the oracle) on
the host cores on a bounded sample of the same workload (the reference has no CPU
implementation of its own).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train it/s (fwd+bwd), 1M Gaussians @1080p"
UNIT = "view-iterations/s"


def env_int(k, d):
    return int(os.environ.get(k, d))


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

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append((time.perf_counter(), float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                                             int(reasons(h))))
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


def cpu_oracle_run(cfg_name: str, sample_P: int, steps: int, warmup: int):
    """CPU-PyTorch execution of the same projection/SH/sort/blend math (the oracle) on a bounded sample."""
    import torch

    from oracle import msplat_oracle as O
    from pointrix_b200 import scene

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c, sc, cams = scene.make_config(cfg_name, P=sample_P, views=max(1, min(4, steps + warmup)))
    dimg = scene.upstream_gradient(3, c["H"], c["W"])
    times = []
    for it in range(warmup + steps):
        v = it % cams["extrinsic_matrix"].shape[0]
        leaves = {k: t.clone().requires_grad_() for k, t in sc.items()}
        t0 = time.perf_counter()
        o = O.render_iter(c["H"], c["W"], cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v],
                          **leaves, sh_degree=3)
        (o["rendered_features_split"]["rgb"] * dimg).sum().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    full_P = scene.CONFIGS[cfg_name]["P"]
    t_sample = sum(times) / len(times)
    t_full = t_sample * full_P / sample_P  # linear extrapolation in #Gaussians (stated in `sample`)
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample_P} of {full_P} Gaussians of {cfg_name} at {c['W']}x{c['H']}, fwd+bwd, "
                      f"{t_sample:.2f} s/view measured, linearly extrapolated x{full_P // sample_P} to the full cloud",
            "s_per_view_sample": t_sample}


def run_reference(args, out_f):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    # bounded: at most 3 timed iterations after at most 1 warm-up, whatever K and W are (2.5 s per iteration on
    # the sample), so the arm ends within a minute; the line says so
    n_timed, n_warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    r = cpu_oracle_run(args.config, args.cpu_sample, n_timed, n_warm)
    r["timed_iterations"], r["warmup_iterations"] = n_timed, n_warm
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / r["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.config), "parallelism": "cpu"},
            "cpu_baseline": r, "gpu_launches": 0,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out_f)


def workload_name(cfg):
    from pointrix_b200 import scene

    c = scene.CONFIGS[cfg]
    return (f"{cfg}: {c['P']} Gaussians, SH degree 3, {c['W']}x{c['H']}, C=3 rgb, white bg, orbit cameras; "
            f"MsplatRender.render_iter fwd + bwd of <G,img> (one view per GPU per step)")


def _claim_stdout():
    """The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner)
    are sent to stderr; the returned file object is the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    out_f = _claim_stdout()
    try:
        _main(out_f)
    finally:
        out_f.flush()


def _main(out_f):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-loss-leg", action="store_true")
    ap.add_argument("--exchange", default="factored", choices=["factored", "allreduce", "nccl"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out_f)

    import torch
    import torch.distributed as dist

    import pointrix_b200 as pb
    from pointrix_b200 import _lib, parallel, scene

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_, K = max(args.warmup, 3), args.steps

    c, sc, cams = scene.make_config(args.config)
    H, W, P, V = c["H"], c["W"], c["P"], c["views"]
    params = {k: v.to(dev).requires_grad_() for k, v in sc.items()}
    cams_d = {k: v.to(dev) for k, v in cams.items()}
    # the batch loss is a mean over the views of all ranks (pointrix/model/loss.py:27-46): the
    # upstream gradient of every view carries 1/world
    dimg = scene.upstream_gradient(3, H, W).to(dev) / world
    r = pb.parse_renderer({"name": "MsplatRender"}, white_bg=True, device=str(dev))
    r.sh_degree = 3

    # data-parallel exchange: the fused backward writes its gradients into symmetric memory and one
    # hand-written kernel all-reduces them over NVLink (in-switch from 4 GPUs on); NCCL only if the
    # group has no peer access
    exch, exch_kind = None, "none"
    if world > 1:
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
            exch, exch_kind = None, f"NCCL all-reduce ({type(e).__name__})"

    def exchange(out, rw):
        if isinstance(exch, parallel.ShFactoredExchange):
            exch.exchange(out["radii"], params["position"])
        elif exch is not None:
            exch.exchange(out["radii"])
        else:
            parallel.allreduce_step([p_.grad for p_ in params.values()], out["uv_points"].grad, out["radii"], world,
                                    average=False, radii_work=rw)

    def view_of(step):  # views sharded by rank
        return (step * world + rank) % V

    def step_fn(step, cam_src=cams_d, g_img=dimg):
        v = view_of(step)
        for p_ in params.values():
            p_.grad = None
        out = r.render_iter(H, W, cam_src["extrinsic_matrix"][v], cam_src["intrinsic_params"], cam_src["camera_center"][v], **params)
        # NCCL path only: radii are final after the forward, their reduce hides under the backward
        rw = parallel.begin_radii_reduce(out["radii"], world) if exch is None else None
        img = out["rendered_features_split"]["rgb"]
        loss = (img * g_img).sum()
        loss.backward()
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
            r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v], **params)
    torch.cuda.synchronize()

    # ---- device-resident timed region --------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: nvidia-smi's own start-up must not overlap the timed region
    for s in range(W_):
        step_fn(s)
    sync()
    if rank == 0:
        sampler.wait_first(3.0)
        sampler.mark()
    # inside the timed region only the dominant kernel is bracketed by events (2 records per step); every
    # stage is timed in a second pass of K steps right after it (8 records per step cost ~3 % of the step)
    DOMINANT = "pxb_blend_backward"
    timer = _lib.KernelTimer(stages={DOMINANT})
    _lib.set_timer(timer)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for s in range(K):
        loss, out = step_fn(W_ + s)
    e1.record()
    sync()
    if rank == 0:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    _lib.set_timer(None)
    clocks = sampler.stop() if rank == 0 else None
    dom_timed = timer.summary().get(DOMINANT)
    stage_timer = _lib.KernelTimer()
    _lib.set_timer(stage_timer)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]  # per-step boundaries (this pass only)
    s0.record()
    step_marks[0].record()
    for s in range(K):
        step_fn(W_ + K + s)
        step_marks[s + 1].record()
    s1.record()
    sync()
    _lib.set_timer(None)
    ms_stage_pass = s0.elapsed_time(s1)
    step_dist = None
    try:  # distribution of the per-step device time (median, p10, p90); never fatal
        per = sorted(step_marks[i].elapsed_time(step_marks[i + 1]) for i in range(K))
        pick = lambda q: round(per[min(K - 1, max(0, int(round(q * (K - 1)))))], 4)  # noqa: E731
        step_dist = {"p10": pick(0.10), "median": pick(0.50), "p90": pick(0.90), "min": round(per[0], 4),
                     "max": round(per[-1], 4), "pass": "stage-timing pass (this rank)"}
    except Exception as e:  # noqa: BLE001
        step_dist = {"unavailable": repr(e)[:120]}
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * K / (ms / 1e3)
    kern = stage_timer.summary()

    # ---- forward-only render Mpix/s (second half of the BASELINE metric) -----------------
    with torch.no_grad():
        for s in range(3):
            r.render_iter(H, W, cams_d["extrinsic_matrix"][view_of(s)], cams_d["intrinsic_params"], cams_d["camera_center"][view_of(s)], **params)
        sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for s in range(K):
            v = view_of(s)
            r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v], **params)
        f1.record()
        sync()
        tf = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        render_mpix = world * K * W * H / 1e6 / (tf.item() / 1e3)

    # ---- end to end through the plugin with HOST buffers --------------------------------
    # per step: pinned host -> device copy of that view's camera (extrinsic 4x4, intrinsics 4,
    # centre 3) and of its dL/dimg [3,H,W]; device -> host read of the loss.
    # one pinned record per view: extrinsic 4x4 | intrinsics 4 | centre 3 -> a single H2D copy per step
    host_cam = torch.cat([cams["extrinsic_matrix"].reshape(V, 16), cams["intrinsic_params"].reshape(1, 4).expand(V, 4),
                          cams["camera_center"].reshape(V, 3)], dim=1).contiguous().pin_memory()
    host_dimg = (scene.upstream_gradient(3, H, W) / world).pin_memory()
    # the loss of step k lands in pinned slot k % 2 and is consumed by the host while step k+1 is queued
    # (a trainer logs the previous iteration's loss): every step's result is read inside the timed region,
    # but the host never drains the GPU between steps
    res_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    res_done = [torch.cuda.Event(), torch.cuda.Event()]
    h2d = (16 + 4 + 3) * 4 + host_dimg.numel() * 4
    d2h = 4
    losses = []

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_collect(step):
        """Host read of the result of `step` (blocks until its D2H copy has landed)."""
        res_done[step % 2].synchronize()
        losses.append(float(res_host[step % 2]))

    def e2e_step(step):
        v = view_of(step)
        main = torch.cuda.current_stream(dev)
        # the camera is needed first (small, on the compute stream); the 25 MB dL/dimg upload runs on a
        # copy stream underneath the forward render and is joined just before the backward needs it
        cam = host_cam[v].to(dev, non_blocking=True)
        E, I, Cc = cam[0:16].view(4, 4), cam[16:20], cam[20:23]
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            G = host_dimg.to(dev, non_blocking=True)
        for p_ in params.values():
            p_.grad = None
        out = r.render_iter(H, W, E, I, Cc, **params)
        rw = parallel.begin_radii_reduce(out["radii"], world) if exch is None else None
        img = out["rendered_features_split"]["rgb"]
        main.wait_stream(copy_stream)
        G.record_stream(main)
        loss = (img * G).sum()
        loss.backward()
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
    losses.clear()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for s in range(K):
        e2e_step(3 + s)
    e2e_collect(3 + K - 1)  # the last step's result, still inside the timed region
    g1.record()
    sync()
    assert len(losses) == K + 1 and all(math.isfinite(x) for x in losses[1:]), "e2e: a step's result was not read"
    te = torch.tensor([g0.elapsed_time(g1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * K / (te.item() / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    loss_info = photometric_loss_leg(r, params, cams_d, view_of, H, W, K, dev) if world == 1 and not args.no_loss_leg else None

    # ---- roofline bookkeeping (algorithmic bytes per SURVEY.md 8d / DESIGN.md) -----------
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # N for the view the stages were measured on varies per view; use the mean over timed views
    Ns = []
    from pointrix_b200 import ops

    with torch.no_grad():
        for s in range(min(K, V)):
            v = view_of(W_ + s)
            uv, depth = ops.project_point(params["position"], cams_d["intrinsic_params"], cams_d["extrinsic_matrix"][v][:3].contiguous(), W, H, nearest=0.2)
            vis = (depth != 0).reshape(-1)
            cov = ops.compute_cov3d(params["scaling"], params["rotation"], vis)
            _, _, tl = ops.ewa_project(params["position"], cov, cams_d["intrinsic_params"], cams_d["extrinsic_matrix"][v][:3].contiguous(), uv, W, H, vis)
            Ns.append(int(tl.sum()))
    N_ref = sum(Ns) / len(Ns)  # length of the reference's tile lists (3-sigma squares)
    # the fused path bins only the tiles the alpha >= 1/255 ellipse reaches: its own count drives the traffic
    N_mean = float(ops.LAST_N.get((dev.index, True), N_ref))
    Cch, S = 3, 12
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    passes = max(1, ((tiles - 1).bit_length() + 7) // 8)
    alg = {
        "pxb_fused_forward": P * (12 + 12 + 16 + 4 + 192 + 4 * S + 4 + 4 + 4),
        # depth-order sort of P (key,id) pairs: init 12 + histogram 4 + 4 passes x 16 B, gathered scan 16 B
        "pxb_bin_prepare": P * (12 + 4 + 4 * 16 + 16),
        # emit (28 B/Gaussian read, 8 B/isect written), histogram 4, tile passes x 16 B, ranges 4 B/isect + 8 B/tile
        "pxb_sort_gaussian": P * 28 + N_mean * (8 + 4 + 16 * passes + 4) + tiles * 8,
        "pxb_blend_forward": N_mean * (4 + 4 * S) + 4 * (Cch + 2) * H * W,
        "pxb_blend_backward": N_mean * (4 + 4 * S) + (4 * Cch + 8) * H * W + N_mean * 8 * (6 + Cch),
        "pxb_fused_backward": P * (4 * S + 12 + 12 + 16 + 192 + 4 + 4 + 12 + 12 + 16 + 4 + 192 + 8),
    }
    stages = {}
    for name, st in kern.items():
        b = alg.get(name)
        stages[name] = {"ms_avg": round(st["ms_avg"], 4), "calls": st["calls"],
                        "share_of_step": round(st["ms_total"] / ms_stage_pass, 4)}
        if b:
            gbs = b / (st["ms_avg"] * 1e-3) / 1e9
            stages[name].update({"alg_bytes": int(b), "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)})
    dom = max(kern.items(), key=lambda kv: kv[1]["ms_total"])[0]
    dom_ms = dom_timed["ms_avg"] if (dom == DOMINANT and dom_timed) else kern[dom]["ms_avg"]
    dom_gbs = alg[dom] / (dom_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if args.config == "cfg4" and os.path.exists(tj):
        traffic = json.load(open(tj)).get(dom, {}).get("dram_bytes_per_launch")
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(dom_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(dom_gbs / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
                "ms_avg_in_timed_region": round(dom_ms, 4), "share_of_step": round(dom_ms * K / ms, 4),
                "note": "blend kernels are FP32-issue/shared-memory bound, not HBM bound: see blend_issue_roofline"}
    # blending as (pixel,Gaussian) pair tests against the FP32 issue bound
    sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
    pair_bound = 148 * 128 * sm_clk * 1e6 / 16.0
    pairs = 256.0 * N_mean
    blend_issue = {}
    for name in ("pxb_blend_forward", "pxb_blend_backward"):
        if name in kern:
            pps = pairs / (kern[name]["ms_avg"] * 1e-3)
            blend_issue[name] = {"pairs_upper_per_launch": pairs, "pairs_per_s": pps, "bound_pairs_per_s": pair_bound,
                                 "frac": round(pps / pair_bound, 4),
                                 "note": "pairs = 256 x intersections (upper bound: early exit skips part of them)"}

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config), "intersections_reference_lists": N_ref,
                   "intersections_binned": N_mean,
                   "parallelism": f"view-sharded dp{world}" + (f", gradients of 59+2 floats/Gaussian summed + max(radii) per step: {exch_kind}" if world > 1 else ""),
                   "cache": "inputs larger than L2 (236 MB Gaussian table + 48 MB records vs 126 MB L2); a different view every step"},
        "render_mpix_s": round(render_mpix, 2),
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_stages": stages,
        "roofline_stages_note": f"second pass of {K} steps with every stage bracketed by CUDA events ({round(ms_stage_pass / K, 4)} ms/step)",
        "step_ms_distribution": step_dist,
        "blend_issue_roofline": blend_issue,
    }
    if loss_info is not None:
        line["photometric_loss"] = loss_info
    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only (the other ranks of a larger run have left)
        line["cpu_baseline"] = cpu_oracle_run(args.config, args.cpu_sample, 2, 1)
    if world == 1 and not args.no_ref_gpu:
        line["ref_gpu"] = ref_gpu_run(args.config, min(K, 10))
    print(json.dumps(line), file=out_f)
    out_f.flush()
    if world > 1:
        dist.destroy_process_group()


def photometric_loss_leg(r, params, cams_d, view_of, H, W, K, dev):
    """SURVEY.md 8f row f1 beside the path: the fused (1-l)*L1 + l*(1-SSIM) loss (csrc/loss.cu) timed alone
    (fwd+bwd, CUDA events, 8 rotating image pairs = 400 MB > L2), the same loss written with the reference's
    torch ops on the same GPU (oracle/loss_oracle.py = pointrix/model/loss.py's formulation: 5 cuDNN depthwise
    convolutions + elementwise kernels + autograd), and the training step render -> loss -> backward."""
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
        PL.l1_ssim_loss(p, gts[i % 8], 0.2)["loss"].backward()

    def torch_ops(i):
        p = preds[i % 8].detach().requires_grad_()
        LO.l1_ssim_loss(p, gts[i % 8], 0.2)["loss"].backward()

    def train_step(i):
        v = view_of(i)
        for p_ in params.values():
            p_.grad = None
        out = r.render_iter(H, W, cams_d["extrinsic_matrix"][v], cams_d["intrinsic_params"], cams_d["camera_center"][v], **params)
        PL.l1_ssim_loss(out["rendered_features_split"]["rgb"].unsqueeze(0), gts[i % 8], 0.2)["loss"].backward()

    n = max(10, min(K, 50))
    ms_ours, ms_ref, ms_step = timed(ours, n), timed(torch_ops, n), timed(train_step, n)
    px = 3 * H * W
    alg = px * (8 + 12) + px * (20 + 4)  # forward 8 read + 12 written, backward 20 read + 4 written per pixel-channel
    return {"fused_fwd_bwd_ms": round(ms_ours, 4), "torch_ops_fwd_bwd_ms": round(ms_ref, 4),
            "alg_bytes": alg, "achieved_gbs": round(alg / (ms_ours * 1e-3) / 1e9, 1),
            "train_step_render_loss_bwd_it_s": round(1e3 / ms_step, 2), "train_step_ms": round(ms_step, 4),
            "note": "loss = 0.8*L1 + 0.2*(1-SSIM) at [1,3,H,W]; torch_ops = the reference's formulation (loss.py:27-117) on this GPU"}


def ref_gpu_run(cfg_name, steps):
    """The compiled unmodified reference CUDA (oracle/_ref) on the same GPU, same inputs:
    reported beside our number, never part of it."""
    try:
        import torch

        from oracle import ref_driver
        from pointrix_b200 import scene

        if not ref_driver.available():
            return {"unavailable": "oracle/_ref not built"}
        c, sc, cams = scene.make_config(cfg_name)
        dev = torch.device("cuda", env_int("LOCAL_RANK", 0))
        sc = {k: v.to(dev) for k, v in sc.items()}
        cams = {k: v.to(dev) for k, v in cams.items()}
        dimg = scene.upstream_gradient(3, c["H"], c["W"]).to(dev)

        def one(v, bwd=True):
            f = ref_driver.render_forward(c["H"], c["W"], cams["extrinsic_matrix"][v], cams["intrinsic_params"], cams["camera_center"][v], **sc)
            if bwd:
                ref_driver.render_backward(f, dimg, sc["position"], sc["opacity"], sc["scaling"], sc["rotation"], sc["shs"], cams["camera_center"][v])

        res = {}
        for mode, bwd in (("fwd_bwd", True), ("fwd", False)):
            for s in range(2):
                one(s % c["views"], bwd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(steps):
                one(s % c["views"], bwd)
            e1.record()
            torch.cuda.synchronize()
            res[mode + "_ms"] = e0.elapsed_time(e1) / steps
        return {"impl": "msplat reference CUDA (sm_100 build, -O3 --use_fast_math) driven in its own op order",
                "it_per_s": 1e3 / res["fwd_bwd_ms"], "ms_fwd_bwd": res["fwd_bwd_ms"], "ms_fwd": res["fwd_ms"],
                "render_mpix_s": c["W"] * c["H"] / 1e6 / (res["fwd_ms"] / 1e3)}
    except Exception as e:  # never let the side measurement kill the bench line
        return {"unavailable": repr(e)[:200]}


if __name__ == "__main__":
    main()
