#!/usr/bin/env python
"""bench.py — frames/s of the VToonify per-frame synthesis hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" is one ``VToonify.forward`` (+ clamp) over one batch of 4 synthetic 576x1024 frames per GPU (configs[1]:
VToonify-D, deterministic random-init weights).  ``value`` = frames/s with inputs resident in HBM; ``e2e`` = the same
through ``FramePipeline`` with HOST buffers (pinned H2D of the fp32 inputs, D2H of the uint8 frames inside the timed
region).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IN, W_IN, BATCH = 576, 1024, 4            # BASELINE.json configs[1]
FLOP_PER_FRAME_D = 6.97e6 * H_IN * W_IN     # BASELINE.md §2 (VToonify-D, per input pixel)
BYTES_PER_FRAME_D = 28.3e3 * H_IN * W_IN


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    def __init__(self, index=0, period=0.2):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80}
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(self.period)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def usable_cores():
    """Host cores this process may really use: min(affinity, cgroup cpu.max quota); os.cpu_count() alone over-reports
    inside a quota-limited container and 128 oversubscribed threads are slower than 8."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per) + 0.999)))
    except Exception:
        pass
    return max(1, n)


# ----------------------------------------------------------------------------------------------------
def cpu_reference_fps(steps, warmup, budget_s=120.0, threads=None):
    """Time the oracle port of the reference CPU path (model/stylegan/op_cpu + F.conv2d) on the host cores.
    Each step is one VToonify-D forward on a bounded sample (B=1, a frame of the same aspect ratio sized to fit the time
    budget); the value is scaled to 576x1024-frame units by pixel count (the network is fully convolutional)."""
    import torch
    from oracle import vt_oracle as O
    from vtoonify_b200.vtoonify import VToonify  # module tree only gives key names/shapes; no kernel is called
    from vtoonify_b200.weights import det_inputs, det_state_dict
    cores = usable_cores()
    with torch.no_grad():
        sd = det_state_dict(VToonify(backbone="dualstylegan"), seed=0)
        # probe cost per pixel on a small frame; pick the thread count (<= usable cores) that is actually fastest
        x, s = det_inputs(1, 144, 256, seed=0)
        cands = [threads] if threads else sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True)
        best = None
        for th in cands:
            torch.set_num_threads(th)
            O.vtoonify_forward(sd, x[:, :, :72, :128], s, 0.5)        # warm the thread pool / primitive cache
            t0 = time.time(); O.vtoonify_forward(sd, x, s, 0.5); dt_probe = time.time() - t0
            if best is None or dt_probe < best[1]:
                best = (th, dt_probe)
            if dt_probe > 20.0:
                continue
        threads, probe = best
        torch.set_num_threads(threads)
        per_px = probe / (144 * 256)
        total = max(1, steps + warmup)
        target_px = budget_s / total / per_px
        scale = min(1.0, (target_px / (H_IN * W_IN)) ** 0.5)
        h = max(72, int(H_IN * scale) // 8 * 8)
        w = max(128, int(W_IN * scale) // 8 * 8)
        x, s = det_inputs(1, h, w, seed=0)
        for _ in range(warmup):
            O.vtoonify_forward(sd, x, s, 0.5)
        t0 = time.time()
        for _ in range(steps):
            O.vtoonify_forward(sd, x, s, 0.5)
        dt = (time.time() - t0) / max(1, steps)
    frac = (h * w) / float(H_IN * W_IN)
    fps = frac / dt
    return fps, dt, {"kind": "port", "cores": threads, "value": fps, "unit": "frames/s",
                     "sample": f"oracle port of the reference op_cpu path, VToonify-D B=1 {h}x{w} frame "
                               f"({frac:.3f} of a 576x1024 frame by pixels), {dt:.2f} s/step, torch CPU fp32 "
                               f"{torch.__version__}, {threads} threads"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    fps, dt, cb = cpu_reference_fps(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": "frames/sec at 576x1024", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "VToonify-D forward, 576x1024 frames (bounded CPU sample scaled by pixels)",
                       "backbone": "dualstylegan", "batch_per_step": 1},
            "cpu_baseline": cb,
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vtoonify_b200 import _lib, ops
    from vtoonify_b200.frame_loop import FramePipeline
    from vtoonify_b200.vtoonify import VToonify
    from vtoonify_b200.weights import det_inputs, det_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    H, W, B = args.height, args.width, args.batch
    with torch.no_grad():
        model = VToonify(backbone=args.backbone).eval()
        model.load_state_dict(det_state_dict(model, seed=0), strict=True)
        model.to(dev)
        x_host, style_host = det_inputs(B, H, W, seed=rank)
        x_host = x_host.pin_memory()
        x = x_host.to(dev)
        style = style_host.to(dev)
        ops.set_precision(args.precision)

        def step():
            y = model(x, style, d_s=0.5)
            return y.clamp_(-1, 1)                      # style_transfer.py:177

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        prof = []
        ops.set_tc_profile(prof)
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        ops.set_tc_profile(None)
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None

        # ---- e2e through the public frame-loop API with host buffers (pinned H2D in, uint8 frames D2H out)
        pipe = FramePipeline(model, style_host[:1], d_s=0.5, device=dev)
        for _ in pipe.run([x_host] * max(1, min(2, args.warmup))):
            pass
        barrier()
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        t0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        nout = 0
        for out in pipe.run([x_host] * args.steps):
            nout += out.shape[0]
        ee1.record()
        barrier()
        e2e_ms_wall = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(ee0.elapsed_time(ee1), e2e_ms_wall)

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    frames = world * B * args.steps
    fps = frames / (ms * 1e-3)
    e2e_fps = frames / (e2e_ms * 1e-3)
    peaks = load_peaks()

    # ---- roofline of the dominant kernel (conv_tc_kernel): aggregate over its launches in the timed region
    tc_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in prof)
    tc_flops = sum(f for _, _, f, _, _ in prof)
    tc_bytes = sum(nb for _, _, _, nb, _ in prof)
    per = {}
    for a, b, f, nb, label in prof:
        d = per.setdefault(label, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += f; d[2] += 1
    if args.dump_layers:
        with open(args.dump_layers, "w") as f:
            for k, v in sorted(per.items(), key=lambda kv: -kv[1][0]):
                f.write(f"{v[0] / args.steps:8.3f} ms  x{v[2] / args.steps:5.1f}  {v[1] / (v[0] * 1e-3) / 1e12:6.1f} TF/s  {k}\n")
    top = sorted(per.items(), key=lambda kv: -kv[1][0])[:6]
    # algorithmic-flop peak of the mode: TF32 = bf16 / 2; bf16x3 issues 3 bf16 products per algorithmic product = bf16 / 3
    div = 3.0 if args.precision == "bf16x3" else 2.0
    tf32_peak = peaks["bf16_tflops_sustained"] / div
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "ncu_conv_tc_latest.json")
    if os.path.exists(prof_json):
        try:
            traffic = json.load(open(prof_json)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    kname = ("conv_tc_kernel (tcgen05 kind::f16 bf16x3 split-operand implicit-GEMM conv)" if args.precision == "bf16x3"
             else "conv_tc_kernel (tcgen05 kind::tf32 implicit-GEMM conv)")
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                "peak_note": (f"algorithmic fp32-product peak = {peaks['source']} bf16 sustained cuBLAS peak ({peaks['bf16_tflops_sustained']:.0f}) / {div:.0f}"
                              + (" (3 bf16 MMA products per algorithmic product)" if div == 3.0 else " (TF32 dense)")),
                "frac_of_bf16_peak": achieved / peaks["bf16_tflops_sustained"],
                "launches": len(prof), "kernel_ms_per_step": tc_ms / args.steps, "share_of_step": tc_ms / ms,
                "algorithmic_gbs": tc_bytes / (tc_ms * 1e-3) / 1e9 if tc_ms > 0 else 0.0,
                "traffic": traffic,
                "top_layers": [{"layer": k, "ms_per_step": v[0] / args.steps, "tflops": v[1] / (v[0] * 1e-3) / 1e12,
                                "launches_per_step": v[2] / args.steps} for k, v in top],
                "whole_step": {"algorithmic_tflop_per_frame": FLOP_PER_FRAME_D / 1e12,
                               "achieved_tflops": FLOP_PER_FRAME_D * frames / world / (ms * 1e-3) / 1e12,
                               "hbm_floor_gbs_needed": BYTES_PER_FRAME_D * frames / world / (ms * 1e-3) / 1e9}}
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        _, _, cb = cpu_reference_fps(1, 0, budget_s=25.0)
    line = {"metric": "frames/sec at 576x1024", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32", "bf16x3": "bf16x3 (fp32 operands split into bf16 hi+lo, 3 tensor-core products, fp32 accumulate)", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"VToonify-{'D' if args.backbone == 'dualstylegan' else 'T'} forward+clamp, "
                                   f"{H}x{W} input frames -> {4 * H}x{4 * W}, batch {B} per GPU per step (BASELINE configs[1])",
                       "backbone": args.backbone, "batch_per_gpu": B, "frames_per_step": world * B,
                       "weights": "deterministic random-init (vtoonify_b200/weights.py)",
                       "l2": "inputs (208 MB) and every activation exceed the 126 MB L2; no flush needed",
                       "parallelism": f"frames sharded round-robin over {world} GPU(s), no collective in the forward"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes // args.steps,
                    "d2h_bytes_per_step": pipe.d2h_bytes // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "api": "vtoonify_b200.frame_loop.FramePipeline.run (pinned fp32 inputs H2D, clamp+uint8 BGR frames D2H)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cb:
        line["cpu_baseline"] = cb
    emit(json.dumps(line))


_REAL_STDOUT = None


def emit(line):
    """Write the result line to the process's original stdout (fd 1 is pointed at stderr while the benchmark runs so that
    library chatter such as NCCL's version banner cannot end up next to the JSON line)."""
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backbone", default="dualstylegan", choices=["dualstylegan", "toonify"])
    ap.add_argument("--dump-layers", default=None, help="write the per-layer conv_tc timing table to this file")
    ap.add_argument("--precision", default="bf16x3", choices=["tf32", "bf16x3", "fp32"],
                    help="bf16x3 (default): split-operand tensor-core mode that meets the 1e-3 parity bar; tf32: faster, 3e-3 error")
    ap.add_argument("--height", type=int, default=H_IN)
    ap.add_argument("--width", type=int, default=W_IN)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
