#!/usr/bin/env python
"""bench.py -- headline benchmark of the VFIDKR hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 1080p_b8|4k_stream]
                    [--no-check] [--table PATH]

Workload (BASELINE.json configs[3], the largest single-GPU configuration; named in config.workload):
one "step" is the hot path of a 1080p frame-pair inference batch, B = 8 pairs padded to 1152x1984
(demo_MiddleBury.py:294-312), in the order the networks call it:
    10 x Correlation forward      (5 PWC-Net levels x 2 directions; C = 196,128,96,64,32; PWCNet.py:230-300)
     2 x DepthFlowProjection fwd  (fillhole = 1, inference; DAIN_slowmotion.py:156-159)
     2 x FilterInterpolation fwd  ("_ori", C = 3, F = 4; DAIN.py:560-573)
and produces 8 interpolated frames.  metric = interpolated Mpixel/s = frames x 1920 x 1080 / s / 1e6.
Frame pairs are independent: with N GPUs every rank runs its own batch of 8 pairs (weak scaling), no
collective on the data path; NCCL only reduces the timing / unit counts.

`value`    inputs resident in HBM, timed with CUDA events on the launch stream, max over ranks.
`e2e`      the same step through the public Python API (vfidkr_b200.PairStream) with HOST buffers: pinned host -> device copies
           of every input and the device -> host read of the interpolated frames are inside the timed region.
`roofline` FilterInterpolation "_ori" forward kernel: algorithmic bytes (96 B/pixel) / its
           average duration measured with CUDA events inside the timed region, against the measured HBM copy peak.
`cpu_baseline`  the float64 CPU oracle (a port: the reference has no CPU implementation of this path) on the
           host cores, on a bounded sample (one pair).
`check`    before anything is timed, rank 0 runs ONE step on the reference's own kernels (oracle/_ref) on the same inputs
           and compares all 14 outputs of the step (10 cost volumes, 2 projected flows, 2 warped frames) -- a fast kernel
           whose results differ from the reference's is not done.  (oracle/_ref as the checker, never as the thing timed.)
`--workload 4k_stream`  BASELINE config 5: a stream of 4K pairs (2176 x 3904 padded) dealt round-robin to the ranks, the
           interpolated frames gathered to rank 0 with NCCL inside the timed region; strong scaling.
`--impl reference`  the reference's OWN CUDA kernels -- oracle/_ref/*.so, the unmodified reference sources built
           for sm_100a by oracle/build_ref.py -- through the same step, called as the reference's Python layers call
           them (caller zero-fills every output).  If those modules or a GPU are absent: the CPU oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FRAME_PIXELS = 1920 * 1080
PAD_H, PAD_W = 1152, 1984                 # 1080p after the /128 replication padding
PAIRS_PER_GPU = 8
PWC_LEVELS = [(196, 64), (128, 32), (96, 16), (64, 8), (32, 4)]   # (channels, downscale) levels 6..2
FI_BYTES_PER_PIXEL = 4 * (2 * 3 + 18)     # 96 B/px at C = 3
METRIC = "interpolated Mpixel/s (1080p)"
WORKLOAD = ("1080p_pair_inference_b8: 10x correlation fwd (5 PWC levels x 2 dirs) + 2x DepthFlowProjection fwd (fillhole) + "
            "2x FilterInterpolation_ori fwd (C=3,F=4)")
PAD_H_4K, PAD_W_4K, FRAME_PIXELS_4K = 2176, 3904, 3840 * 2160    # 4K after the /128 replication padding


def workload_config(world: int) -> dict:
    """`config` of the JSON line -- built ONCE, identical in both arms (the driver compares the two)."""
    return {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "frame": "1920x1080 padded to 1152x1984",
            "parallelism": f"pair-sharded x{world}, no data-path collective",
            "l2": "inputs_exceed_l2 (one step streams > 4 GB, L2 is 126 MB)",
            "flow": "synthetic scene flow: quarter-res smooth field + 6 moving rectangles + 0.25 px jitter, x4 bilinear (bench.py: scene_flow)"}
FALLBACK_HBM_GBS = 6650.0                 # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


# ----------------------------------------------------------------------------- sharding / reduction helpers
def shard_pairs(n_pairs: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of independent frame pairs to ranks."""
    return list(range(rank, n_pairs, world))


def reduce_timing(t_ms, units):
    """max over ranks of the elapsed time, sum over ranks of the processed units (tensors of shape [1])."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    return t_ms, units


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the FI kernel from the committed ncu capture, if one has been recorded."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("fi_forward_ori_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------- clock sampling
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index: int, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.backend = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.backend = "nvml"
        except Exception:
            self.nv = None

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self._BITS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no clock samples available"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- the workload
def scene_flow(torch, g, device, B, H, W):
    """Synthetic optical flow with the statistics the warping ops see in the network: PWC-Net's flow lives at quarter
    resolution and is upsampled 4x bilinearly (networks/DAIN.py:306-308), so it is smooth at the 4-pixel scale; at
    larger scales a scene moves piecewise smoothly.  Quarter-resolution field = a smooth large-scale motion
    (1/64-resolution N(0, 6^2) px, bicubic) + six independently moving rectangular "objects" per item (constant extra
    motion N(0, 8^2) px, sharp motion boundaries) + N(0, 0.25^2) px estimation jitter; clipped to +-24 px; then x4
    bilinear.  (The operator table also times the rougher `up4` field -- i.i.d. N(0, 4^2) at quarter resolution --
    and a per-pixel i.i.d. field as stress cases.)"""
    F = torch.nn.functional
    h4, w4 = H // 4, W // 4
    coarse = torch.randn((B, 2, max(H // 64, 2), max(W // 64, 2)), generator=g, device=device) * 6.0
    lo = F.interpolate(coarse, size=(h4, w4), mode="bicubic", align_corners=False)
    n_obj = 6
    motion = torch.randn((B, n_obj, 2), generator=g, device=device) * 8.0
    geom = torch.rand((B, n_obj, 4), generator=g, device=device)
    yy = torch.arange(h4, device=device).view(1, h4, 1)
    xx = torch.arange(w4, device=device).view(1, 1, w4)
    for k in range(n_obj):
        y0 = (geom[:, k, 0] * 0.8 * h4).view(B, 1, 1)
        x0 = (geom[:, k, 1] * 0.8 * w4).view(B, 1, 1)
        hh = ((0.08 + 0.25 * geom[:, k, 2]) * h4).view(B, 1, 1)
        ww = ((0.08 + 0.25 * geom[:, k, 3]) * w4).view(B, 1, 1)
        inside = ((yy >= y0) & (yy < y0 + hh) & (xx >= x0) & (xx < x0 + ww)).unsqueeze(1)   # [B,1,h4,w4]
        lo = lo + inside * motion[:, k].view(B, 2, 1, 1)
    lo = lo + torch.randn((B, 2, h4, w4), generator=g, device=device) * 0.25
    lo = lo.clamp_(-24, 24)
    return F.interpolate(lo, scale_factor=4, mode="bilinear", align_corners=False).contiguous()


def build_inputs(torch, device, seed, pinned_host=False):
    """Synthetic inputs of one step.  Returns a dict of tensors; with pinned_host
    they live in pinned host memory, else on `device`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    B, H, W = PAIRS_PER_GPU, PAD_H, PAD_W

    def make(shape, kind):
        # generated on the device with a seeded generator; the pinned-host variant is then copied out, so the
        # end-to-end loop really starts from host memory
        if kind == "image":
            x = torch.rand(shape, generator=g, device=device)
        elif kind == "flow":
            x = scene_flow(torch, g, device, shape[0], shape[2], shape[3])
        elif kind == "filter":
            x = torch.softmax(torch.randn(shape, generator=g, device=device), dim=1)
        elif kind == "depth":
            x = torch.rand(shape, generator=g, device=device) * 0.9 + 0.1
        else:
            x = torch.randn(shape, generator=g, device=device)
        if not pinned_host:
            return x
        h = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        h.copy_(x)
        return h

    d = {}
    for k in (0, 1):   # the two temporal directions
        d[f"frame{k}"] = make((B, 3, H, W), "image")
        d[f"flow{k}"] = make((B, 2, H, W), "flow")
        d[f"filter{k}"] = make((B, 16, H, W), "filter")
        d[f"rawflow{k}"] = make((B, 2, H, W), "flow")
    d["depth"] = make((B, 1, H, W), "depth")
    for lvl, (C, s) in enumerate(PWC_LEVELS):
        for k in (0, 1):
            d[f"feat{lvl}_{k}"] = make((B, C, H // s, W // s), "normal")
    return d


def run_step(V, mods, d, fi_events=None):
    """One pass of the hot path over one batch.  Returns the two warped frame batches."""
    corr, dproj, fi = mods
    outs = []
    for lvl in range(len(PWC_LEVELS)):
        a, b = d[f"feat{lvl}_0"], d[f"feat{lvl}_1"]
        o01, o10 = corr.both_directions(a, b)     # direction 0 -> 1 and 1 -> 0 of this pyramid level: one launch
        outs.append(o01)
        outs.append(o10)
    p0 = dproj(d["rawflow0"], d["depth"])
    p1 = dproj(d["rawflow1"], d["depth"])
    warped = []
    for k in (0, 1):
        if fi_events is not None:
            fi_events[k][0].record()
        warped.append(fi(d[f"frame{k}"], d[f"flow{k}"], d[f"filter{k}"]))
        if fi_events is not None:
            fi_events[k][1].record()
    return warped, (p0, p1), outs


def reference_step(torch, mods, d, keep=False):
    """The same step on the reference's own kernels (oracle/_ref), driven the way its Python layers drive them
    (caller zero-fills every output: FilterInterpolationLayer.py:34, DepthFlowProjectionLayer.py:33-35; the correlation
    resizes its own, correlation.py:24-31).  keep: also return the cost volumes and projections (for the check)."""
    corr_m, dproj_m, fi_m = mods["correlation_cuda"], mods["depthflowprojection_cuda"], mods["filterinterpolation_cuda"]
    outs, proj, warped = [], [], []
    for lvl in range(len(PWC_LEVELS)):
        a, b = d[f"feat{lvl}_0"], d[f"feat{lvl}_1"]
        for x, y in ((a, b), (b, a)):
            rb1, rb2, o = x.new_empty(0), y.new_empty(0), x.new_empty(0)
            corr_m.forward(x, y, rb1, rb2, o, 4, 1, 4, 1, 1, 1)
            if keep:
                outs.append(o)
    for k in (0, 1):
        cnt = torch.zeros_like(d["depth"])
        po = torch.zeros_like(d[f"rawflow{k}"])
        dproj_m.DepthFlowProjectionLayer_gpu_forward(d[f"rawflow{k}"], d["depth"], cnt, po, 1)
        if keep:
            proj.append(po)
    for k in (0, 1):
        out = torch.zeros_like(d[f"frame{k}"])
        fi_m.FilterInterpolationLayer_gpu_forward_ori(d[f"frame{k}"], d[f"flow{k}"], d[f"filter{k}"], out)
        warped.append(out)
    return warped, proj, outs


def check_step(torch, V, mods, d):
    """Rank 0, before timing: every output of one step against the reference's own kernels on the same inputs.
    Both sides are fp32 and each is within the stated tolerance of the float64 oracle (tests/), so they may differ by
    twice that: 2e-5 (cost volumes, warped frames), 2e-4 (projections: atomically ordered sums, discontinuous hole filling
    at exact ties is the same arithmetic on both sides)."""
    ref_mods = _reference_modules()
    if ref_mods is None:
        return {"ran": False, "why": "oracle/_ref modules not available"}
    with torch.no_grad():
        warped, proj, outs = run_step(V, mods, d)
        rwarped, rproj, routs = reference_step(torch, ref_mods, d, keep=True)
        torch.cuda.synchronize()

        def err(a, b):
            scale = b.abs().max().clamp_min(1e-30)
            return float(((a - b).abs() / (b.abs() + scale)).max())

        res = {"correlation": max(err(a, b) for a, b in zip(outs, routs)),
               "depthflowprojection": max(err(a, b) for a, b in zip(proj, rproj)),
               "filterinterpolation": max(err(a, b) for a, b in zip(warped, rwarped))}
    tol = {"correlation": 2e-5, "depthflowprojection": 2e-4, "filterinterpolation": 2e-5}
    ok = all(res[k] <= tol[k] for k in res)
    del rwarped, rproj, routs
    torch.cuda.empty_cache()
    return {"ran": True, "ok": ok, "against": "oracle/_ref (the reference's kernels, same inputs, one step, 14 outputs)",
            "max_normalised_error": res, "tolerance": tol}


def h2d_ceiling(torch, device, world, dist, mib=1024, reps=4):
    """What the box's host->device path delivers with every rank copying from pinned memory at once: GB/s summed over ranks."""
    host = torch.empty(mib << 18, dtype=torch.float32, pin_memory=True)
    dev = torch.empty_like(host, device=device)
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    nbytes = torch.tensor([float(reps * host.numel() * 4)], dtype=torch.float64, device=device)
    reduce_timing(ms, nbytes)
    del host, dev
    return float(nbytes.item()) / (ms.item() * 1e-3) / 1e9


def bench_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import vfidkr_b200 as V
    numa_cpus = V.bind_to_gpu_numa_node(device)      # before any pinned allocation: first touch lands on the GPU's node
    mods = (V.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1),
            V.DepthFlowProjectionModule(requires_grad=False), V.FilterInterpolationModule())

    d = build_inputs(torch, device, seed=1004 + rank)
    n_px = PAIRS_PER_GPU * PAD_H * PAD_W

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    check = None
    if rank == 0 and not args.no_check:
        try:
            check = check_step(torch, V, mods, d)
        except Exception as e:      # the checker must never take the measurement down with it
            check = {"ran": False, "why": f"{type(e).__name__}: {e}"}

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            run_step(V, mods, d)
        sync_all()

        # ---- device-resident timing ----
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in (0, 1)]
              for _ in range(args.steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = V.launch_count()
        sync_all()
        t0.record()
        for s in range(args.steps):
            run_step(V, mods, d, ev[s])
        t1.record()
        sync_all()
        launches = V.launch_count() - launches0
        clocks = sampler.stop()
        ms_total = t0.elapsed_time(t1)
        fi_ms = [e[0].elapsed_time(e[1]) for step in ev for e in step]
        fi_avg_ms = sum(fi_ms) / len(fi_ms)

        # ---- the dominant kernel on SURVEY 8d's stated flow distributions (outside the timed step; rank 0) ----
        other_flows = {}
        if rank == 0 and not args.timed_only:
            g = torch.Generator(device=device)
            g.manual_seed(77)
            B, H, W = PAIRS_PER_GPU, PAD_H, PAD_W
            flows = {"up4": torch.nn.functional.interpolate((torch.randn((B, 2, H // 4, W // 4), generator=g, device=device) * 4).clamp_(-20, 20),
                                                            scale_factor=4, mode="bilinear", align_corners=False).contiguous(),
                     "iid": (torch.randn((B, 2, H, W), generator=g, device=device) * 4).clamp_(-20, 20)}
            for name, fl in flows.items():
                for _ in range(2):
                    mods[2](d["frame0"], fl, d["filter0"])
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    mods[2](d["frame0"], fl, d["filter0"])
                b.record()
                torch.cuda.synchronize()
                other_flows[name] = a.elapsed_time(b) / 5
            del flows

        # ---- end to end: pinned host inputs -> device -> step -> frames back on the host ----
        e2e = None
        if not args.no_e2e:
            ceiling = h2d_ceiling(torch, device, world, dist)
            hd = build_inputs(torch, device, seed=2004 + rank, pinned_host=True)
            h2d = sum(t.numel() * 4 for t in hd.values())
            host_out = torch.empty((2, PAIRS_PER_GPU, 3, PAD_H, PAD_W), dtype=torch.float32, pin_memory=True)
            d2h = host_out.numel() * 4
            e_steps = min(args.steps, args.e2e_steps)

            # vfidkr_b200.PairStream: pairs are independent, so copies in, kernels and copies out of successive pairs
            # overlap on three streams (the call a user with host-resident pairs makes)
            stream = V.PairStream(device, lambda dd: run_step(V, mods, dd)[0], pairs_per_chunk=args.e2e_chunk)

            def e2e_step():
                stream.run(hd, (host_out[0], host_out[1]))
            for _ in range(3):   # warm-up: the per-stream allocator pools and the pinned pages settle
                e2e_step()
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(e_steps):
                e2e_step()
            e1.record()      # the caller's stream waits for each run's results (PairStream.run)
            sync_all()
            e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
            e2e_units = torch.tensor([float(e_steps * PAIRS_PER_GPU)], dtype=torch.float64, device=device)
            reduce_timing(e2e_ms, e2e_units)
            h2d_gbs = world * e_steps * h2d / (e2e_ms.item() * 1e-3) / 1e9
            e2e = {"value": float(e2e_units.item()) * FRAME_PIXELS / (e2e_ms.item() * 1e-3) / 1e6, "unit": "Mpixel/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                   "how": f"vfidkr_b200.PairStream: chunks of {args.e2e_chunk} pair(s), H2D / kernels / D2H on three streams",
                   # the end-to-end number is the host->device path of the box, not the kernels: say so with numbers
                   "h2d_achieved_GBps": h2d_gbs, "h2d_ceiling_GBps": ceiling, "frac_of_h2d_ceiling": h2d_gbs / ceiling,
                   "h2d_ceiling_how": f"{world} rank(s) copying 1 GiB pinned -> device concurrently, summed",
                   "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
            del hd, host_out

    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=device)
    units = torch.tensor([float(args.steps * PAIRS_PER_GPU)], dtype=torch.float64, device=device)   # frames out
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=device)
    per_rank_ms = None
    if world > 1:   # the spread behind the max: every rank's own device time per step (GPUs of one box are not identical)
        allt = [torch.zeros_like(t_ms) for _ in range(world)]
        dist.all_gather(allt, t_ms)
        per_rank_ms = [round(float(t.item()) / args.steps, 4) for t in allt]
    reduce_timing(t_ms, units)
    if world > 1:
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    value = float(units.item()) * FRAME_PIXELS / (t_ms.item() * 1e-3) / 1e6

    peak, peak_src = measured_peak()
    alg = FI_BYTES_PER_PIXEL * n_px
    achieved = alg / (fi_avg_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Mpixel/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_ms.item() / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "roofline": {"bound": "hbm", "kernel": "fi_forward_ori_strip_kernel<3>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": recorded_traffic(),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": fi_avg_ms,
                     # the same kernel on SURVEY.md 8d's stated flow distributions (quarter-resolution N(0,4^2) x4; per-pixel i.i.d.)
                     "frac_up4_flow": alg / (other_flows["up4"] * 1e-3) / 1e9 / peak if other_flows else None,
                     "frac_iid_flow": alg / (other_flows["iid"] * 1e-3) / 1e9 / peak if other_flows else None},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(nl.item()), "check": check,
    }
    if per_rank_ms is not None:
        line["ms_per_step_by_rank"] = per_rank_ms
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_arm(steps=5, warmup=1)
            line["cpu_baseline"]["torch_cpu_fp32"] = torch_cpu_arm()
        emit_result(line)
    if args.table and rank == 0 and world == 1:
        op_table(torch, V, device, args.table)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- BASELINE config 5: a 4K stream, pair-sharded
def gather_frames(dist, frame, dst_list, rank, world, async_op=False):
    """One round of the frame gather: every rank contributes the interpolated frame of its pair of this round, rank 0
    receives them into `dst_list` (one tensor per rank).  NCCL (or gloo in the CPU tests); None when world == 1."""
    if world == 1:
        dst_list[0].copy_(frame)
        return None
    return dist.gather(frame, gather_list=dst_list if rank == 0 else None, dst=0, async_op=async_op)


def bench_4k_stream(args):
    """P 4K pairs dealt round-robin with shard_pairs; per pair: 10 correlations, 2 DepthFlowProjections (fillhole) and
    the two adaptive warps blended into ONE interpolated frame (filter_interpolate_blend); the frames are gathered to
    rank 0 over NCCL inside the timed region.  Strong scaling: the stream is the same for every N."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import vfidkr_b200 as V
    corr = V.Correlation(4, 1, 4, 1, 1, 1)
    dproj = V.DepthFlowProjectionModule(False)
    P, H, W = args.pairs, PAD_H_4K, PAD_W_4K
    mine = shard_pairs(P, rank, world)
    rounds = (P + world - 1) // world
    g = torch.Generator(device=device)

    def pair_inputs(idx):
        g.manual_seed(5000 + idx)
        d = {}
        for k in (0, 1):
            d[f"frame{k}"] = torch.rand((1, 3, H, W), generator=g, device=device)
            d[f"flow{k}"] = scene_flow(torch, g, device, 1, H, W)
            d[f"filter{k}"] = torch.softmax(torch.randn((1, 16, H, W), generator=g, device=device), dim=1)
            d[f"rawflow{k}"] = scene_flow(torch, g, device, 1, H, W)
        d["depth"] = torch.rand((1, 1, H, W), generator=g, device=device) * 0.9 + 0.1
        for lvl, (C, s) in enumerate(PWC_LEVELS):
            for k in (0, 1):
                d[f"feat{lvl}_{k}"] = torch.randn((1, C, H // s, W // s), generator=g, device=device)
        return d

    # distinct inputs per pair while they fit comfortably (1.85 GB each); beyond that the pairs of a rank share 8 sets
    sets = [pair_inputs(i) for i in mine[:8]]
    out_rank0 = torch.empty((rounds * world, 3, H, W), device=device) if rank == 0 else None
    dummy = torch.zeros((1, 3, H, W), device=device)

    def process(j):
        d = sets[j % len(sets)]
        for lvl in range(len(PWC_LEVELS)):
            corr.both_directions(d[f"feat{lvl}_0"], d[f"feat{lvl}_1"])
        dproj(d["rawflow0"], d["depth"])
        dproj(d["rawflow1"], d["depth"])
        return V.filter_interpolate_blend(d["frame0"], d["frame1"], d["flow0"], d["flow1"], d["filter0"], d["filter1"], 0.5, 0.5)

    def stream_once():
        works = []
        for r in range(rounds):
            frame = process(r) if r < len(mine) else dummy       # ranks without a pair in the last round send a blank
            dst = [out_rank0[r * world + k].unsqueeze(0) for k in range(world)] if rank == 0 else None
            w = gather_frames(dist, frame, dst if rank == 0 else [None], rank, world, async_op=True)
            if w is not None:
                works.append((w, frame))        # keep the frame alive until the gather has consumed it
        for w, _ in works:
            w.wait()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            stream_once()
        sync_all()
        sampler = ClockSampler(local_rank)
        sampler.start()
        n0 = V.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t0.record()
        for _ in range(args.steps):
            stream_once()
        t1.record()
        sync_all()
        clocks = sampler.stop()
        launches = V.launch_count() - n0
    t_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    units = torch.tensor([float(args.steps * len(mine))], dtype=torch.float64, device=device)
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=device)
    reduce_timing(t_ms, units)
    if world > 1:
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    if rank == 0:
        frames = float(units.item())
        emit_result({
            "metric": "interpolated Mpixel/s (4K)", "value": frames * FRAME_PIXELS_4K / (t_ms.item() * 1e-3) / 1e6, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_ms.item() / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"4k_stream: {P} frame pairs 3840x2160 padded to {H}x{W}, one pair per GPU step (10x correlation fwd + "
                                   "2x DepthFlowProjection fwd (fillhole) + FilterInterpolation_ori both directions blended), dealt round-robin",
                       "pairs": P, "parallelism": f"pair-sharded x{world}; NCCL gather of the interpolated frames to rank 0 inside the timed region",
                       "gathered_bytes_per_step": int((P - len(shard_pairs(P, 0, world))) * 3 * H * W * 4),
                       "l2": "inputs_exceed_l2 (1.85 GB per pair)"},
            "frames_per_step": P, "ms_per_pair_per_gpu": t_ms.item() / args.steps / max(1, len(mine)),
            "clocks": clocks, "gpu_launches": int(nl.item()), "e2e": None})
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- CPU arms (oracle port, PyTorch-CPU fp32)
def _cpu_inputs():
    import numpy as np
    r = np.random.default_rng(1004)
    H, W = PAD_H, PAD_W
    frames = [r.random((1, 3, H, W), dtype=np.float32) for _ in (0, 1)]

    def upflow():   # quarter-resolution N(0,4^2) flow, x4 (nearest: the CPU cost does not depend on the flow)
        lo = np.clip(r.standard_normal((1, 2, H // 4, W // 4)) * 4, -20, 20).astype(np.float32)
        return np.ascontiguousarray(np.repeat(np.repeat(lo, 4, axis=2), 4, axis=3))
    flows = [upflow() for _ in range(4)]
    z = [r.standard_normal((1, 16, H, W)).astype(np.float32) for _ in (0, 1)]
    filts = [np.exp(a - a.max(1, keepdims=True)) for a in z]
    filts = [(a / a.sum(1, keepdims=True)).astype(np.float32) for a in filts]
    depth = (0.1 + 0.9 * r.random((1, 1, H, W), dtype=np.float32)).astype(np.float32)
    feats = [[r.standard_normal((1, C, H // s, W // s)).astype(np.float32) for _ in (0, 1)] for C, s in PWC_LEVELS]
    return frames, flows, filts, depth, feats


def cpu_reference_arm(steps=1, warmup=0):
    """Times the CPU oracle on a bounded sample of the same workload: ONE frame pair (of the 8 per step),
    every op of the step, all host threads OpenMP gives it."""
    from oracle import oracle as O
    O.build()
    frames, flows, filts, depth, feats = _cpu_inputs()

    def step():
        for a, b in feats:
            O.correlation_forward(a, b, 4, 1, 4, 1, 1)
            O.correlation_forward(b, a, 4, 1, 4, 1, 1)
        O.flowprojection_forward(flows[2], depth, 1)
        O.flowprojection_forward(flows[3], depth, 1)
        for k in (0, 1):
            O.fi_forward("ori", frames[k], flows[k], filts[k])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": steps * 1 * FRAME_PIXELS / dt / 1e6, "unit": "Mpixel/s", "cores": O.num_threads(),
            "kind": "port", "sample": "1 frame pair of the 8 per step (all 14 op calls of the step), float64 C oracle + OpenMP",
            "seconds": dt, "steps": steps}


def torch_cpu_arm(steps=2, warmup=1):
    """The second CPU baseline of SURVEY.md 8d(ii): the step in vectorised PyTorch-CPU fp32 (oracle/torch_cpu.py:
    gather / scatter_add_ / shifted products) on the same bounded sample, all host cores."""
    import torch
    from oracle import torch_cpu as T
    torch.set_num_threads(os.cpu_count() or 1)
    frames, flows, filts, depth, feats = [[torch.from_numpy(a) for a in x] if isinstance(x, list) and not isinstance(x[0], list) else x
                                          for x in _cpu_inputs()]
    depth = torch.from_numpy(depth) if not isinstance(depth, torch.Tensor) else depth
    feats = [[torch.from_numpy(a) for a in pair] for pair in feats]

    def step():
        for a, b in feats:
            T.correlation_forward(a, b)
            T.correlation_forward(b, a)
        T.flowprojection_forward(flows[2], depth, 1)
        T.flowprojection_forward(flows[3], depth, 1)
        for k in (0, 1):
            T.fi_ori_forward(frames[k], flows[k], filts[k])

    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    return {"value": steps * FRAME_PIXELS / dt / 1e6, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 frame pair of the 8 per step, vectorised PyTorch-CPU fp32 (oracle/torch_cpu.py)", "seconds": dt, "steps": steps}


def _reference_modules():
    """The reference's own extension modules (oracle/_ref), or None when they / a GPU are not available."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        import importlib.util
        spec = importlib.util.spec_from_file_location("vfidkr_build_ref", str(ROOT / "oracle" / "build_ref.py"))
        build_ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(build_ref)
        return {n: build_ref.load(n) for n in ("correlation_cuda", "depthflowprojection_cuda", "filterinterpolation_cuda")}
    except Exception as e:   # missing .so, ABI mismatch, ...
        print(f"[bench] reference CUDA modules unavailable ({e}); using the CPU oracle port", file=sys.stderr)
        return None


def bench_reference_cuda(args, mods):
    """The same step on the reference's unmodified CUDA kernels, driven the way its Python layers drive them.
    EVERY rank runs its own batch, exactly as in our arm (weak scaling): max time over ranks, units summed."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(1, args.steps), max(args.warmup, 3)
    d = build_inputs(torch, device, seed=1004 + rank)
    with torch.no_grad():
        for _ in range(warm):
            reference_step(torch, mods, d)
        sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            reference_step(torch, mods, d)
        t1.record()
        sync_all()
        t_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
        units = torch.tensor([float(steps * PAIRS_PER_GPU)], dtype=torch.float64, device=device)
        reduce_timing(t_ms, units)
        value = float(units.item()) * FRAME_PIXELS / (t_ms.item() * 1e-3) / 1e6
        del d
        torch.cuda.empty_cache()
        # end to end from pinned host buffers, as in our arm
        hd = build_inputs(torch, device, seed=2004 + rank, pinned_host=True)
        h2d = sum(t.numel() * 4 for t in hd.values())
        host_out = torch.empty((2, PAIRS_PER_GPU, 3, PAD_H, PAD_W), dtype=torch.float32, pin_memory=True)
        e_steps = min(steps, args.e2e_steps)

        def e2e_step():
            dd = {k: t.to(device, non_blocking=True) for k, t in hd.items()}
            warped = reference_step(torch, mods, dd)[0]
            host_out[0].copy_(warped[0], non_blocking=True)
            host_out[1].copy_(warped[1], non_blocking=True)
        e2e_step()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e_steps):
            e2e_step()
        e1.record()
        sync_all()
        e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        e_units = torch.tensor([float(e_steps * PAIRS_PER_GPU)], dtype=torch.float64, device=device)
        reduce_timing(e_ms, e_units)
        e2e_value = float(e_units.item()) * FRAME_PIXELS / (e_ms.item() * 1e-3) / 1e6
    if rank == 0:
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": t_ms.item() / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "reference_kind": "reference CUDA extensions (oracle/_ref)",
            "note": "reference = its own CUDA kernels, unmodified sources compiled for sm_100a (oracle/_ref), on the same GPUs -- it "
                    "has no CPU implementation of this path; every rank runs its own batch, as in our arm",
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": host_out.numel() * 4, "steps": e_steps},
            "gpu_launches": 0,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_arm(steps=3, warmup=1)
        emit_result(line)
    if world > 1:
        dist.destroy_process_group()


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    mods = None if args.reference_cpu else _reference_modules()
    if mods is not None:
        return bench_reference_cuda(args, mods)
    if rank != 0:       # the CPU port: rank 0 alone runs and prints it
        return
    res = cpu_reference_arm(steps=max(1, min(args.steps, 3)), warmup=1 if args.warmup > 0 else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "Mpixel/s",
        "n_gpus": 1, "steps": res["steps"], "warmup": 1 if args.warmup > 0 else 0,
        "ms_per_step": res["seconds"] / res["steps"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1),
        "reference_kind": "CPU oracle port",
        "note": "the reference has no CPU implementation of this path and its CUDA modules (oracle/_ref) were not available; this arm is "
                "the float64 CPU oracle port on the host cores, rank 0 only, on a bounded sample (1 pair per step)",
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_result(line)


# ----------------------------------------------------------------------------- per-op table (not the bench line)
def op_table(torch, V, device, path):
    """Per-operator device timings at the config-4 shape with achieved GB/s against SURVEY.md 8a bytes/pixel.
    Written to `path` as JSON lines; this is the optimisation dashboard, not the headline."""
    B, H, W = PAIRS_PER_GPU, PAD_H, PAD_W
    peak, _ = measured_peak()
    rows = []

    def timeit(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    def add(name, ms, bytes_per_px, px):
        gbs = bytes_per_px * px / (ms * 1e-3) / 1e9
        rows.append({"op": name, "ms": ms, "algorithmic_GB": bytes_per_px * px / 1e9, "GBps": gbs, "frac_of_hbm": gbs / peak})

    px = B * H * W
    I = torch.rand(B, 3, H, W, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(77)
    fl = scene_flow(torch, gen, device, B, H, W)                             # the bench's flow model
    fl_up4 = torch.nn.functional.interpolate((torch.randn(B, 2, H // 4, W // 4, device=device) * 4).clamp_(-20, 20),
                                             scale_factor=4, mode="bilinear", align_corners=False).contiguous()   # rough stress field
    fl_iid = (torch.randn(B, 2, H, W, device=device) * 4).clamp_(-20, 20)   # per-pixel i.i.d. flow: worst case for the gathers
    ft = torch.softmax(torch.randn(B, 16, H, W, device=device), 1)
    off = (torch.rand(B, 32, H, W, device=device) - 0.5) * 0.9
    dep = torch.rand(B, 1, H, W, device=device) * 0.9 + 0.1
    g = torch.randn(B, 3, H, W, device=device)
    g2 = torch.randn(B, 2, H, W, device=device)
    with torch.no_grad():
        add("FI_ori_fwd_C3", timeit(lambda: V.FilterInterpolationLayer.apply(I, fl, ft)), 96, px)
        add("FI_ori_fwd_C3_up4flow", timeit(lambda: V.FilterInterpolationLayer.apply(I, fl_up4, ft)), 96, px)
        add("FI_ori_fwd_C3_iidflow", timeit(lambda: V.FilterInterpolationLayer.apply(I, fl_iid, ft)), 96, px)
        yy, xx = torch.meshgrid(torch.linspace(0, 6, H, device=device), torch.linspace(0, 6, W, device=device), indexing="ij")
        fl_smooth = torch.stack([6 * torch.sin(xx) + 3 * torch.cos(yy), 5 * torch.cos(0.7 * xx) - 3 * torch.sin(yy)], 0)[None].repeat(B, 1, 1, 1).contiguous()
        add("FI_ori_fwd_C3_smoothflow", timeit(lambda: V.FilterInterpolationLayer.apply(I, fl_smooth, ft)), 96, px)
        del fl_smooth, yy, xx
        # both directions warped and blended in two launches (2 x 84 B/px in, 12 B/px out)
        I2 = torch.rand_like(I)
        add("FI_ori_blend_two_directions_C3", timeit(lambda: V.filter_interpolate_blend(I, I2, fl, fl_up4, ft, ft)), 180, px)
        del I2
        add("FI_dkr_fwd_C3", timeit(lambda: V.FilterInterpolationLayerDKR.apply(I, fl, ft, off)), 224, px)
        add("FI_deforconv_fwd_C3", timeit(lambda: V.FilterInterpolationLayerDeforConv.apply(I, fl, ft, off)), 224, px)
        add("FI_nofilter_fwd_C3", timeit(lambda: V.FilterInterpolationLayerNoFilterWithDeforConv.apply(I, fl, off)), 160, px)
        add("Interpolation_fwd_C3", timeit(lambda: V.InterpolationLayer.apply(I, fl)), 32, px)
        add("FlowProjection_fwd_fill", timeit(lambda: V.FlowProjectionLayer.apply(fl, False)), 20, px)
        add("FlowProjection_fwd", timeit(lambda: V.FlowProjectionLayer.apply(fl, True)), 20, px)
        add("DepthFlowProjection_fwd_fill", timeit(lambda: V.DepthFlowProjectionLayer.apply(fl, dep, False)), 24, px)
        add("DepthFlowProjection_fwd_fill_up4flow", timeit(lambda: V.DepthFlowProjectionLayer.apply(fl_up4, dep, False)), 24, px)
        add("FI_dkr_fwd_C3_up4flow", timeit(lambda: V.FilterInterpolationLayerDKR.apply(I, fl_up4, ft, off)), 224, px)
    # backward timings through the C ABI directly (no autograd bookkeeping in the timed region)
    from vfidkr_b200 import _lib
    from vfidkr_b200._common import ptr, stream_ptr
    sp = stream_ptr(device)
    gi1, gi2, gi3, gi4 = torch.empty_like(I), torch.empty_like(fl), torch.empty_like(ft), torch.empty_like(off)
    add("FI_ori_bwd_C3", timeit(lambda: _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(I), ptr(fl), ptr(ft), ptr(g),
                                                   ptr(gi1), ptr(gi2), ptr(gi3), B, 3, H, W, 4, sp)), 180, px)
    add("FI_dkr_bwd_C3", timeit(lambda: _lib.call("vfidkr_filterinterpolation_backward_dkr", ptr(I), ptr(fl), ptr(ft), ptr(off),
                                                   ptr(g), ptr(gi1), ptr(gi2), ptr(gi3), ptr(gi4), B, 3, H, W, 4, sp)), 436, px)
    add("FI_deforconv_bwd_C3", timeit(lambda: _lib.call("vfidkr_filterinterpolation_backward_deforconv", ptr(I), ptr(fl), ptr(ft),
                                                         ptr(off), ptr(g), ptr(gi1), ptr(gi2), ptr(gi3), ptr(gi4), B, 3, H, W, 4, sp)), 436, px)
    add("Interpolation_bwd_C3", timeit(lambda: _lib.call("vfidkr_interpolation_backward", ptr(I), ptr(fl), ptr(g), ptr(gi1), ptr(gi2),
                                                          B, 3, H, W, 1, sp)), 52, px)
    cnt = torch.empty(B, 1, H, W, device=device)
    po = torch.empty(B, 2, H, W, device=device)
    gd = torch.empty(B, 1, H, W, device=device)
    _lib.call("vfidkr_depthflowprojection_forward", ptr(fl), ptr(dep), ptr(cnt), ptr(po), B, H, W, 0, sp)
    add("DepthFlowProjection_bwd", timeit(lambda: _lib.call("vfidkr_depthflowprojection_backward", ptr(fl), ptr(dep), ptr(cnt), ptr(po),
                                                             ptr(g2), ptr(gi2), ptr(gd), B, H, W, sp)), 44, px)
    _lib.call("vfidkr_flowprojection_forward", ptr(fl), ptr(cnt), ptr(po), B, H, W, 0, sp)
    add("FlowProjection_bwd", timeit(lambda: _lib.call("vfidkr_flowprojection_backward", ptr(fl), ptr(cnt), ptr(g2), ptr(gi2), B, H, W, sp)), 28, px)
    del I, ft, off, g, gi1, gi3, gi4
    torch.cuda.empty_cache()
    with torch.no_grad():
        corr = V.Correlation(4, 1, 4, 1, 1, 1)
        for C, s in PWC_LEVELS:
            a = torch.randn(B, C, H // s, W // s, device=device)
            b = torch.randn_like(a)
            add(f"Correlation_fwd_C{C}_{H // s}x{W // s}", timeit(lambda: corr(a, b)), 4 * (2 * C + 81), B * (H // s) * (W // s))
        # correlation backward at the two finest PWC levels (reads f1, f2, gradoutput; writes both gradients)
        for C, sdiv in ((64, 8), (32, 4)):
            a = torch.randn(B, C, H // sdiv, W // sdiv, device=device)
            b = torch.randn_like(a)
            go = torch.randn(B, 81, H // sdiv, W // sdiv, device=device)
            ga, gb = torch.empty_like(a), torch.empty_like(b)
            add(f"Correlation_bwd_C{C}_{H // sdiv}x{W // sdiv}",
                timeit(lambda: _lib.call("vfidkr_correlation_backward", ptr(a), ptr(b), ptr(go), ptr(ga), ptr(gb), B, C, H // sdiv, W // sdiv,
                                         4, 1, 4, 1, 1, 1, sp)), 4 * (4 * C + 81), B * (H // sdiv) * (W // sdiv))
            del a, b, go, ga, gb
        # PWCDCNet.warp at the PWC level-2 shape (reads C features + 2 flow planes, writes C)
        feat = torch.randn(B, 32, H // 4, W // 4, device=device)
        flo4 = torch.nn.functional.avg_pool2d(fl, 4) / 4
        add("PWC_warp_C32_288x496", timeit(lambda: V.pwc_warp(feat, flo4), iters=20), 4 * (2 * 32 + 2), B * (H // 4) * (W // 4))
        del feat, flo4
        # the 196-channel context warp of DAIN_slowmotion, once (two batch items keep it at 7 GB)
        Bc = 2
        ctx = torch.rand(Bc, 196, H, W, device=device)
        ftc = torch.softmax(torch.randn(Bc, 16, H, W, device=device), 1)
        add("FI_ori_fwd_C196_B2", timeit(lambda: V.FilterInterpolationLayer.apply(ctx, fl[:Bc].contiguous(), ftc), iters=3),
            4 * (2 * 196 + 18), Bc * H * W)
        del ctx, ftc
        torch.cuda.empty_cache()
    # BASELINE configs 3 (training step, B = 16 at 256x448) and 5 (one 4K pair, padded to 2176x3904) on the same operators
    for tag, (Bx, Hx, Wx) in (("cfg3_B16_256x448", (16, 256, 448)), ("cfg5_4K_B1", (1, 2176, 3904))):
        pxx = Bx * Hx * Wx
        gen.manual_seed(78)
        Ix = torch.rand(Bx, 3, Hx, Wx, device=device)
        flx = scene_flow(torch, gen, device, Bx, Hx, Wx)
        ftx = torch.softmax(torch.randn(Bx, 16, Hx, Wx, device=device), 1)
        offx = (torch.rand(Bx, 32, Hx, Wx, device=device) - 0.5) * 0.9
        depx = torch.rand(Bx, 1, Hx, Wx, device=device) * 0.9 + 0.1
        gx = torch.randn(Bx, 3, Hx, Wx, device=device)
        o1, o2, o3, o4 = torch.empty_like(Ix), torch.empty_like(flx), torch.empty_like(ftx), torch.empty_like(offx)
        with torch.no_grad():
            add(f"FI_ori_fwd_C3_{tag}", timeit(lambda: V.FilterInterpolationLayer.apply(Ix, flx, ftx), iters=20), 96, pxx)
            add(f"FI_dkr_fwd_C3_{tag}", timeit(lambda: V.FilterInterpolationLayerDKR.apply(Ix, flx, ftx, offx), iters=20), 224, pxx)
            add(f"DepthFlowProjection_fwd_{tag}", timeit(lambda: V.DepthFlowProjectionLayer.apply(flx, depx, True), iters=20), 24, pxx)
        add(f"FI_ori_bwd_C3_{tag}", timeit(lambda: _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(Ix), ptr(flx), ptr(ftx), ptr(gx),
                                                              ptr(o1), ptr(o2), ptr(o3), Bx, 3, Hx, Wx, 4, sp), iters=20), 180, pxx)
        add(f"FI_dkr_bwd_C3_{tag}", timeit(lambda: _lib.call("vfidkr_filterinterpolation_backward_dkr", ptr(Ix), ptr(flx), ptr(ftx), ptr(offx),
                                                              ptr(gx), ptr(o1), ptr(o2), ptr(o3), ptr(o4), Bx, 3, Hx, Wx, 4, sp), iters=20), 436, pxx)
        del Ix, flx, ftx, offx, depx, gx, o1, o2, o3, o4
        torch.cuda.empty_cache()
    # BASELINE config 1 (one 256x448 frame pair, the reference's CPU-runnable case): a latency, not a bandwidth -- the op is a
    # single wave of 18 us of work, so the row reports microseconds through the C ABI and through the autograd Function
    gen.manual_seed(79)
    I1 = torch.rand(1, 3, 256, 448, device=device)
    fl1 = scene_flow(torch, gen, device, 1, 256, 448)
    ft1 = torch.softmax(torch.randn(1, 16, 256, 448, device=device), 1)
    o1 = torch.empty_like(I1)
    add("FI_ori_fwd_C3_cfg1_B1_256x448_c_abi", timeit(lambda: _lib.call("vfidkr_filterinterpolation_forward_ori", ptr(I1), ptr(fl1), ptr(ft1),
                                                                         ptr(o1), 1, 3, 256, 448, 4, sp), iters=200), 96, 256 * 448)
    rows[-1]["latency_us"] = rows[-1]["ms"] * 1e3
    with torch.no_grad():
        add("FI_ori_fwd_C3_cfg1_B1_256x448_python", timeit(lambda: V.FilterInterpolationLayer.apply(I1, fl1, ft1), iters=200), 96, 256 * 448)
    rows[-1]["latency_us"] = rows[-1]["ms"] * 1e3
    del I1, fl1, ft1, o1
    # SeparableConv at F = 51 (the reference's own test size, test_module.py:903-907) is arithmetic-bound:
    # the row carries the fp32 rate next to the (irrelevant) byte rate.  3*C*F*F multiply-adds per output pixel forward,
    # three times that backward.
    Bs, Hs, Ws, Fs = 8, 256, 448, 51
    Hos, Wos = Hs - Fs + 1, Ws - Fs + 1
    Is = torch.rand(Bs, 3, Hs, Ws, device=device)
    vs = torch.rand(Bs, Fs, Hos, Wos, device=device) / Fs
    hs = torch.rand(Bs, Fs, Hos, Wos, device=device) / Fs
    os_, gs = torch.empty(Bs, 3, Hos, Wos, device=device), torch.randn(Bs, 3, Hos, Wos, device=device)
    g1s, g2s, g3s = torch.empty_like(Is), torch.empty_like(vs), torch.empty_like(hs)
    pxs = Bs * Hos * Wos
    add("SeparableConv_fwd_F51_B8_256x448", timeit(lambda: _lib.call("vfidkr_separableconv_forward", ptr(Is), ptr(vs), ptr(hs), ptr(os_),
                                                                      Bs, 3, Hs, Ws, Fs, sp)), 4 * (2 * 3 + 2 * Fs), pxs)
    rows[-1]["fp32_TFLOPs"] = 3 * 3 * Fs * Fs * pxs / (rows[-1]["ms"] * 1e-3) / 1e12
    add("SeparableConv_bwd_F51_B8_256x448", timeit(lambda: _lib.call("vfidkr_separableconv_backward", ptr(Is), ptr(vs), ptr(hs), ptr(gs),
                                                                      ptr(g1s), ptr(g2s), ptr(g3s), Bs, 3, Hs, Ws, Fs, sp)), 4 * (3 * 3 + 4 * Fs), pxs)
    rows[-1]["fp32_TFLOPs"] = 3 * 3 * 3 * Fs * Fs * pxs / (rows[-1]["ms"] * 1e-3) / 1e12
    fls = torch.empty(Bs, 2, Hos, Wos, device=device)
    add("SeparableConvFlow_fwd_F51_B8_206x398", timeit(lambda: _lib.call("vfidkr_separableconvflow_forward", ptr(vs), ptr(hs), ptr(fls),
                                                                          Bs, Hos, Wos, Fs, sp), iters=20), 4 * (2 * Fs + 2), pxs)
    gfl = torch.randn(Bs, 2, Hos, Wos, device=device)
    add("SeparableConvFlow_bwd_F51_B8_206x398", timeit(lambda: _lib.call("vfidkr_separableconvflow_backward", ptr(vs), ptr(hs), ptr(gfl),
                                                                          ptr(g2s), ptr(g3s), Bs, Hos, Wos, Fs, sp), iters=20), 4 * (4 * Fs + 2), pxs)
    del Is, vs, hs, os_, gs, g1s, g2s, g3s, fls, gfl
    with open(path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    for r in rows:
        print(f"[op] {r['op']:34s} {r['ms']*1e3:10.1f} us  {r['GBps']:8.1f} GB/s  {100*r['frac_of_hbm']:5.1f} % of HBM peak", file=sys.stderr)


_RESULT_FD = None


def emit_result(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5, help="cap on the (PCIe-bound) end-to-end steps")
    ap.add_argument("--e2e-chunk", type=int, default=2, help="pairs per chunk of the end-to-end PairStream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-cpu", action="store_true", help="--impl reference: force the CPU oracle port")
    ap.add_argument("--table", default=None, help="also write the per-operator timing table (JSON lines) here")
    ap.add_argument("--workload", choices=["1080p_b8", "4k_stream"], default="1080p_b8",
                    help="1080p_b8: the headline line (BASELINE config 4); 4k_stream: config 5, pair-sharded with an NCCL gather")
    ap.add_argument("--pairs", type=int, default=16, help="4k_stream: pairs in the stream (same for every N: strong scaling)")
    ap.add_argument("--no-check", action="store_true", help="skip the pre-timing comparison against the reference's kernels")
    ap.add_argument("--timed-only", action="store_true",
                    help="warm-up + timed steps only (no check, flow variants, end-to-end or CPU legs): the command the ncu launch list is taken of")
    args = ap.parse_args()
    if args.timed_only:
        args.no_check = args.no_e2e = args.no_cpu_baseline = True
    # stdout carries exactly ONE line, the JSON result: everything else a library may print there (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for instance) is sent to stderr, and the result goes to a private copy of the real stdout
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        bench_reference(args)
    elif args.workload == "4k_stream":
        bench_4k_stream(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
