"""Workloads timed by bench.py.

Each workload owns: seeded synthetic inputs generated on the CPU (so the CPU baseline and the
GPU see identical bits), the device-resident step, the end-to-end step through
the public API with pinned host buffers, the per-step result gather, and the roofline / CPU
baseline bookkeeping.
"""
from __future__ import annotations

import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json, else the profiling guide's
    stated fallback (6.65 TB/s, 1.59 PFLOP/s)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def msda_algorithmic_bytes(N, S, M, D, L, Lq, P, elem_size):
    """SURVEY.md 8d / BASELINE.md section 3: value + loc(2) + weight(1) + out, + shape tables."""
    return elem_size * (N * S * M * D + 3 * N * Lq * M * L * P + N * Lq * M * D) + 24 * L


class Workload:
    metric = "interleaved_steps_per_sec"
    unit = "steps/s"
    dtype_name = "bf16"

    def __init__(self, rank, world, local_batch):
        self.rank, self.world = rank, world
        self.local_batch = local_batch
        self._flush = None
        self._launches = 0

    # -- shared helpers -----------------------------------------------------------------
    def flush_l2(self):
        if self._flush is None:
            self._flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self._flush.zero_()

    def reset_counters(self):
        self._launches = 0
        self._kernel_events = []

    def launch_count(self):
        return self._launches

    def gather(self, dist):
        """One all_gather of the per-rank result checksum per step (the only collective)."""
        if dist is None:
            return
        chk = self.result_checksum()
        out = [torch.empty_like(chk) for _ in range(self.world)]
        dist.all_gather(out, chk)

    def kernel_stats(self):
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in self._kernel_events]
        return {"launches": len(ms), "avg_ms": sum(ms) / len(ms) if ms else None}


class MsdaCfg3(Workload):
    """BASELINE cfg 3, deformable-attention sampler only: the 10 MMFS cross-attention layers'
    MSDA launches for `local_batch` 4-image / 2048-token sequences per GPU (per layer: N =
    local_batch, L = 12, S = 5376, Lq = 2048, M = 16, D = 64, P = 8, bf16).  Sampling locations
    follow the LLM flavour (reference point (0.5, 0.5) + offsets drawn like the reference's
    U(-3,3)/16 bias init); the image-visibility mask zeroes the weights of images a token cannot
    see, exactly as the MMFS softmax produces them (exp(-1e4) == 0)."""

    name = "msda_cfg3"
    LAYERS = 10
    SHAPES = [(32, 32), (16, 16), (8, 8)] * 4
    M, D, Lq, P = 16, 64, 2048, 8

    def make_host_inputs(self, pin):
        from oracle import level_start_index
        B = self.local_batch or 2
        self.B = B
        g = torch.Generator().manual_seed(1234 + self.rank)
        L = len(self.SHAPES)
        shapes = torch.tensor(self.SHAPES, dtype=torch.long)
        self.shapes_h, self.starts_h = shapes, level_start_index(shapes)
        S = int(shapes.prod(1).sum())
        self.S, self.L = S, L
        dt = torch.bfloat16
        self.layers_h = []
        for _ in range(self.LAYERS):
            value = torch.rand((B, S, self.M, self.D), generator=g).to(dt)
            off = (torch.rand((B, self.Lq, self.M, 4, 1, self.P, 2), generator=g) * 6 - 3) / 16.0
            off = off + 0.05 * torch.randn((B, self.Lq, self.M, 4, 1, self.P, 2), generator=g)
            loc = (0.5 + off).expand(B, self.Lq, self.M, 4, 3, self.P, 2).reshape(B, self.Lq, self.M, L, self.P, 2)
            logits = torch.randn((B, self.Lq, self.M, L, self.P), generator=g)
            vis = self.visibility()                                     # (Lq, 4) 0/1
            logits = logits + (1.0 - vis)[None, :, None, :, None].repeat_interleave(3, dim=3) * -10000.0
            attn = torch.softmax(logits.flatten(-2), -1).view(B, self.Lq, self.M, L, self.P)
            lay = tuple(t.to(dt).contiguous() for t in (value, loc, attn))
            self.layers_h.append(tuple(t.pin_memory() for t in lay) if pin else lay)

    def setup(self):
        import mm_interleaved_b200 as m
        self.m = m
        self.make_host_inputs(pin=True)
        B, dt = self.B, torch.bfloat16
        self.shapes_d, self.starts_d = self.shapes_h.cuda(), self.starts_h.cuda()
        self.layers_d = [tuple(t.cuda() for t in lay) for lay in self.layers_h]
        self.out_h = torch.empty((B, self.Lq, self.M * self.D), dtype=dt).pin_memory()
        self.last = None
        self.reset_counters()

    def visibility(self):
        """cfg 3 token layout: images at token offsets 1, 512, 1024, 1536, an extra
        <bos> at 1023; image i visible to token t iff soi_i+1 > nearest_bos(t) and soi_i+1 <= t
        (mm_interleaved.py:208-221)."""
        t = torch.arange(self.Lq)
        soi = torch.tensor([1, 512, 1024, 1536]) + 1
        nearest_bos = torch.where(t >= 1023, 1023, 0)
        return ((soi[None, :] > nearest_bos[:, None]) & (soi[None, :] <= t[:, None])).float()

    def units_per_step(self):
        return self.B

    def step_device(self):
        for (value, loc, attn) in self.layers_d:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.last = self.m.ms_deform_attn_forward(value, self.shapes_d, self.starts_d, loc, attn, 64)
            e1.record()
            self._kernel_events.append((e0, e1))
            self._launches += 1

    def step_e2e(self):
        for (value, loc, attn) in self.layers_h:
            self.m.ms_deform_attn_forward_host(value, self.shapes_h, self.starts_h, loc, attn, out=self.out_h)
            self._launches += 1
        self.last = self.out_h

    def result_checksum(self):
        return self.last.float().sum().reshape(1).to("cuda")

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for lay in self.layers_h for t in lay)

    def d2h_bytes(self):
        return self.LAYERS * self.out_h.numel() * self.out_h.element_size()

    def config(self):
        return {"workload": "BASELINE cfg3 MMFS deformable-attention sampler only: 10 cross-attn layers x "
                            f"{self.B} sequences/GPU (L=12,S=5376,Lq=2048,M=16,D=64,P=8); the rest of the "
                            "interleaved forward is not part of this workload",
                "step_unit": "one 4-image/2048-token sequence (its 10 sampler launches)",
                "global_batch": self.B * self.world, "seq_len": 2048, "images_per_seq": 4,
                "parallelism": f"dp{self.world}", "l2": "192 MiB buffer written between timed steps (L2 flush)"}

    def roofline(self, kernel):
        peaks = measured_peaks()
        ab = msda_algorithmic_bytes(self.B, self.S, self.M, self.D, self.L, self.Lq, self.P, 2)
        achieved = ab / (kernel["avg_ms"] * 1e-3) / 1e9 if kernel["avg_ms"] else None
        return {"kernel": "msda_fwd_warp_kernel<bf16,64>", "bound": "hbm", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"] if achieved else None,
                "algorithmic_bytes_per_launch": ab, "avg_launch_us": kernel["avg_ms"] * 1e3 if kernel["avg_ms"] else None,
                "launches_timed": kernel["launches"], "traffic": None}

    def cpu_baseline(self):
        return cpu_baseline_msda(self, layers=1)

    def setup_cpu_only(self):
        self.make_host_inputs(pin=False)

    def reference_step(self):
        """One bounded reference-arm step: ONE layer of ONE sequence through the reference's CPU path."""
        from oracle import msda_core_pytorch
        value, loc, attn = (t[:1].float() for t in self.layers_h[0])
        return msda_core_pytorch(value, self.shapes_h, loc, attn)

    reference_step_fraction = 1.0 / LAYERS      # of one step unit (a sequence = 10 layers)
    reference_sample = "each step = 1 of the 10 sampler layers of 1 sequence, fp32, scaled x10"


def cpu_baseline_msda(wl, layers=1, threads=None):
    """The reference's CPU-runnable path for the sampler (restated `ms_deform_attn_core_pytorch`,
    oracle/msda.py) on a bounded sample: `layers` layer(s) of ONE sequence, fp32, all host cores;
    scaled to steps/s as 1 / (10 layers x time per layer)."""
    from oracle import msda_core_pytorch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    value, loc, attn = (t[:1].float() for t in wl.layers_h[0])
    msda_core_pytorch(value, wl.shapes_h, loc, attn)     # warm-up
    t0 = time.time()
    n = 0
    while n < 3 or time.time() - t0 < 8.0:
        msda_core_pytorch(value, wl.shapes_h, loc, attn)
        n += 1
    per_layer = (time.time() - t0) / n
    return {"value": 1.0 / (per_layer * wl.LAYERS), "unit": wl.unit, "cores": threads, "kind": "port",
            "sample": f"{n} runs of 1 layer x 1 sequence of the same workload ({per_layer * 1e3:.0f} ms each), fp32, "
                      f"scaled x{wl.LAYERS} layers"}


WORKLOADS = {"msda_cfg3": MsdaCfg3}
AUTO = "msda_cfg3"


def make(name, rank, world, local_batch):
    if name == "auto":
        name = AUTO
    return WORKLOADS[name](rank, world, local_batch)


def run_reference_arm(args, world):
    """`bench.py --impl reference`: the reference's own CPU implementation of the path (oracle port of
    ms_deform_attn_core_pytorch; the reference's CUDA op has no CPU implementation,
    ops/src/ms_deform_attn.h:38) timed on the host cores with all threads, bounded sample per step."""
    name = AUTO if args.workload == "auto" else args.workload
    wl = WORKLOADS[name](0, world, 1)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    wl.setup_cpu_only()
    for _ in range(max(args.warmup, 1)):
        wl.reference_step()
    t0 = time.time()
    for _ in range(args.steps):
        wl.reference_step()
    dt = time.time() - t0
    value = wl.reference_step_fraction * args.steps / dt
    cfg = wl.config()
    return {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": wl.unit, "cores": threads, "kind": "port",
                             "sample": wl.reference_sample},
            "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
