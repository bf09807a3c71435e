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


class InterleavedCfg3(Workload):
    """BASELINE cfg 3: the interleaved image-text forward on 4-image / 2048-token sequences, bf16 -- embed
    splice + image-visibility mask + MMFS feature packing (mm_interleaved.py:121-252), the 40-layer
    Llama-13B decoder with MMFS cross-attention in every 4th layer (modeling_llama_mmfs.py:623-752) and the
    text head (decoder_text.py:140-163), i.e. the body of MMInterleaved.forward up to the logits, INCLUDING the
    visual tokenizer (CLIP ViT-L/14 + ViT-Adapter + 12-layer Q-Former, visual_tokenizer.py:65-101) on the
    4 x B images.  Random-init weights of the real architecture, synthetic token layout of SURVEY.md 8d."""

    name = "interleaved_cfg3"
    T, N_IMG, TOK_PER_IMG = 2048, 4, 64
    SOI_AT = (1, 512, 1024, 1536)
    EXTRA_BOS_AT = 1023
    LAYERS_CROSS, LAYERS_TOTAL = 10, 40

    def _config(self):
        from mm_interleaved_b200.llama_mmfs import LlamaMMFSConfig
        return LlamaMMFSConfig()

    def make_host_inputs(self, pin):
        B = self.local_batch or 4
        self.B = B
        g = torch.Generator().manual_seed(4321 + self.rank)
        ids = torch.randint(3, 31999, (B, self.T), generator=g)
        ids[:, 0] = 1
        for s in self.SOI_AT:
            ids[:, s] = 32001
            ids[:, s + 1:s + 1 + self.TOK_PER_IMG] = 32000
        ids[:, self.EXTRA_BOS_AT] = 1
        dt = torch.bfloat16
        images = torch.rand((B * self.N_IMG, 3, 224, 224), generator=g)          # [0,1] like the reference's image_tensors
        host = [ids, images]
        self.host = [t.pin_memory() for t in host] if pin else host

    def setup(self):
        import mm_interleaved_b200 as m
        from mm_interleaved_b200 import ops, sampler
        from mm_interleaved_b200.mm_interleaved import InterleavedForward
        from mm_interleaved_b200.mmfs import MMFS
        self.m, self.ops = m, ops
        self.make_host_inputs(pin=True)
        cfg = self._config()
        self.cfg = cfg
        with torch.device("meta"):
            model = InterleavedForward(cfg).to(torch.bfloat16)
        model = model.to_empty(device="cuda")
        gen = torch.Generator(device="cuda").manual_seed(7)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("norm.weight") or name.endswith("layernorm.weight") or ".norm1." in name or ".norm2." in name:
                    p.fill_(1.0)
                elif name.endswith(".gate"):
                    p.fill_(0.5)
                elif name.endswith("sampling_offsets.bias"):
                    p.uniform_(-3.0, 3.0, generator=gen)                     # mmfs.py:103-110
                elif name.endswith("ignore_token") or name.endswith(".bias"):
                    p.zero_()
                elif name.endswith("sampling_offsets.weight"):
                    p.normal_(0.0, 0.004, generator=gen)
                else:
                    p.normal_(0.0, 0.02, generator=gen)
            for mod in model.modules():
                if isinstance(mod, MMFS):
                    mod.scale_ratios = torch.tensor(mod._scale_list, device="cuda")
        self.model = model.eval()
        from mm_interleaved_b200.visual_tokenizer import Injector, VisualTokenizer
        torch.manual_seed(11)
        tok = VisualTokenizer()                                  # ViT-L/14 + adapter + Q-Former(64 queries, 12 layers)
        with torch.no_grad():
            for mod in tok.modules():
                if isinstance(mod, Injector):
                    mod.gamma.fill_(0.5)                         # zero-initialised in the reference: make the branch count
            tok.proj.weight.normal_(0.0, 0.02)
        self.tok = tok.to("cuda", torch.bfloat16).eval()
        self.dev = [t.cuda() for t in self.host]
        # The tokenizer is ~2400 small launches for 16 images (launch-bound on the host): capture it once in a CUDA
        # graph over static buffers and replay it every step.
        self.tok_in = torch.zeros((self.B * self.N_IMG, 3, 224, 224), dtype=torch.bfloat16, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                self.tok(self.tok_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = ops.launch_counter[0]
        self.tok_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.tok_graph), torch.no_grad():
            self.tok_out = self.tok(self.tok_in)
        self.tok_graph_launches = ops.launch_counter[0] - before
        self.nimg = torch.full((self.B,), self.N_IMG, dtype=torch.long, device="cuda")
        self.out_h = torch.empty((self.B, self.T), dtype=torch.long).pin_memory()
        self.last = None
        # time the dominant hand-written kernels with CUDA events from inside the step
        self._sampler_events, self._attn_events = [], []
        orig_sampler, orig_attn = sampler.mmfs_sampler_forward, ops.attention

        def timed_sampler(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig_sampler(*a, **k); e1.record()
            self._sampler_events.append((e0, e1)); self._launches += 1
            return r

        def timed_attn(q, *a, **k):
            if q.shape[1] != self.T or q.shape[3] != 128 or torch.cuda.is_current_stream_capturing():
                return orig_attn(q, *a, **k)                      # only the Llama prefill attention is the roofline subject
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig_attn(q, *a, **k); e1.record()
            self._attn_events.append((e0, e1))
            return r

        sampler.mmfs_sampler_forward = timed_sampler
        import mm_interleaved_b200.mmfs as mmfs_mod
        mmfs_mod._sampler.mmfs_sampler_forward = timed_sampler
        import mm_interleaved_b200.llama_mmfs as lm
        lm.ops.attention = timed_attn
        self.reset_counters()

    def reset_counters(self):
        super().reset_counters()
        self._sampler_events, self._attn_events = [], []
        if hasattr(self, "ops"):
            self.ops.launch_counter[0] = 0

    def launch_count(self):
        return self._launches + self.ops.launch_counter[0]

    def units_per_step(self):
        return self.B

    def _forward(self, tensors):
        ids, images = tensors[0], tensors[1]
        with torch.no_grad():
            self.tok_in.copy_(images)                             # fp32 [0,1] images -> bf16 static buffer
            self.tok_graph.replay()
            self._launches += self.tok_graph_launches
            logits = self.model(ids, self.tok_out, self.nimg, self.N_IMG)
            return logits.argmax(-1)

    def step_device(self):
        self.last = self._forward(self.dev)

    def step_e2e(self):
        dev = [t.to("cuda", non_blocking=True) for t in self.host]
        pred = self._forward(dev)
        self.out_h.copy_(pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.last = pred

    def result_checksum(self):
        return self.last.sum().reshape(1).float()

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.host)

    def d2h_bytes(self):
        return self.out_h.numel() * self.out_h.element_size()

    def kernel_stats(self):
        torch.cuda.synchronize()
        s = [a.elapsed_time(b) for a, b in self._sampler_events]
        t = [a.elapsed_time(b) for a, b in self._attn_events]
        return {"launches": len(s), "avg_ms": sum(s) / len(s) if s else None,
                "attn_launches": len(t), "attn_avg_ms": sum(t) / len(t) if t else None}

    def config(self):
        return {"workload": f"BASELINE cfg3 interleaved forward, {self.B} sequences/GPU x (4 images 224^2, 2048 tokens): visual tokenizer "
                            "(CLIP ViT-L/14 + ViT-Adapter + Q-Former) + embed splice + visibility mask + MMFS feature packing + "
                            "Llama-13B decoder (40 layers, MMFS cross-attn every 4th) + text head + argmax",
                "step_unit": "one 4-image/2048-token sequence forward",
                "global_batch": self.B * self.world, "seq_len": self.T, "images_per_seq": self.N_IMG,
                "parallelism": f"dp{self.world}", "params": "13B Llama + 10 MMFS layers + 0.45B visual tokenizer, random init, bf16",
                "l2": "192 MiB buffer written between timed steps (L2 flush); weights (27 GB) exceed L2 anyway",
                "cuda_graph": "visual tokenizer captured once and replayed per step; decoder eager"}

    def roofline(self, kernel):
        peaks = measured_peaks()
        S, M, D, L, Lq, P = 5376, 16, 64, 12, self.T, 8
        ab = msda_algorithmic_bytes(self.B, S, M, D, L, Lq, P, 2)            # SURVEY.md 8d figure (34.1 MB / layer / sequence)
        C = M * P * 2 + M * 3 * (P + 1)
        fused = 2 * (self.B * S * M * D + self.B * Lq * C + 50 * C + self.B * Lq * M * D) + self.B * 4 * Lq
        t = kernel["avg_ms"] * 1e-3 if kernel["avg_ms"] else None
        achieved = ab / t / 1e9 if t else None
        flops = 4.0 * self.B * self.T * self.T * 5120 / 2                     # causal QK^T + PV per layer
        ta = kernel["attn_avg_ms"] * 1e-3 if kernel.get("attn_avg_ms") else None
        return {"kernel": "mmfs_sampler_kernel<bf16,64> (fused MMFS deform-attn sampler)", "bound": "hbm",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"] if achieved else None,
                "algorithmic_bytes_per_launch": ab, "fused_kernel_bytes_per_launch": fused,
                "achieved_fused_bytes_GBs": fused / t / 1e9 if t else None,
                "avg_launch_us": t * 1e6 if t else None, "launches_timed": kernel["launches"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch (B = 4) from the committed capture
                # profiles/r01_sampler_cfg3_fused_ncu_details.txt (ncu --set full): 20.93 MB + 10.75 MB
                "traffic": 31.68e6 * self.B / 4.0,
                "attention": {"kernel": "attn_fwd_kernel<bf16,128> (tcgen05)", "bound": "tensor",
                              "achieved": flops / ta / 1e12 if ta else None, "peak": peaks["bf16_tflops_sustained"],
                              "unit": "TFLOP/s", "frac": flops / ta / 1e12 / peaks["bf16_tflops_sustained"] if ta else None,
                              "flops_per_launch": flops, "avg_launch_us": ta * 1e6 if ta else None,
                              "launches_timed": kernel.get("attn_launches")}}

    # ---- CPU baseline / reference arm: oracle restatement of the reference forward, bounded sample ----
    def setup_cpu_only(self):
        self.local_batch = 1
        self.make_host_inputs(pin=False)

    def _cpu_layer_weights(self, cross, seed):
        import mm_interleaved_b200  # noqa: F401
        from mm_interleaved_b200.llama_mmfs import LlamaDecoderLayer
        cfg = self._config()
        layer = LlamaDecoderLayer(cfg, cross, 0)
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for k, v in layer.state_dict().items():
            if k.endswith("norm.weight") or k.endswith("layernorm.weight") or ".norm1." in k or ".norm2." in k:
                sd["layers.0." + k] = torch.ones_like(v)
            elif k.endswith(".gate"):
                sd["layers.0." + k] = torch.full_like(v, 0.5)
            elif k.endswith("sampling_offsets.bias"):
                sd["layers.0." + k] = torch.rand(v.shape, generator=g) * 6 - 3
            elif k.endswith(".bias") or k.endswith("ignore_token"):
                sd["layers.0." + k] = torch.zeros_like(v)
            else:
                sd["layers.0." + k] = torch.randn(v.shape, generator=g) * 0.02
        return sd

    def reference_step(self):
        """Bounded CPU sample: ONE plain decoder layer + ONE MMFS cross-attention layer of ONE sequence (T = 2048,
        4 images, fp32) through the oracle restatement of the reference (oracle/llama.py, oracle/mmfs.py)."""
        from oracle.llama import additive_mask_ref, llama_layer_ref
        from mm_interleaved_b200.mm_interleaved import cross_attention_mask_from_ids, pack_mmfs_features
        if not hasattr(self, "_cpu"):
            ids = self.host[0][:1]
            x = torch.randn((1, self.T, 5120), generator=torch.Generator().manual_seed(5)) * 0.5
            gf = torch.Generator().manual_seed(9)
            feats = pack_mmfs_features([torch.randn((4, 1024, sz, sz), generator=gf) for sz in (32, 16, 8)], [32, 16, 8],
                                       torch.tensor([4]), 4)
            cross = cross_attention_mask_from_ids(ids, 4, 1, 32001, torch.tensor([4]))
            cfg = dict(eps=1e-6, n_heads=40, n_layers=1, spatial_shapes=[(32, 32), (16, 16), (8, 8)])
            add_mask = additive_mask_ref(torch.ones((1, self.T)), self.T, 0, torch.float32)
            pos = torch.arange(self.T)[None]
            self._cpu = (x, feats, cross, cfg, add_mask, pos, self._cpu_layer_weights(False, 1), self._cpu_layer_weights(True, 2))
        x, feats, cross, cfg, add_mask, pos, w_plain, w_cross = self._cpu
        with torch.no_grad():
            t0 = time.time()
            llama_layer_ref(w_plain, 0, x, None, None, add_mask, pos, cfg)
            t1 = time.time()
            llama_layer_ref(w_cross, 0, x, feats, cross, add_mask, pos, cfg)
            t2 = time.time()
        self._cpu_times = (t1 - t0, t2 - t1)
        return self._cpu_times

    # one reference step covers 2 of the 40 layers; scaled as 30 plain + 10 cross layers (head + glue excluded,
    # < 2 % of the work), see reference_sample
    reference_step_fraction = None
    reference_sample = ("each step = 1 plain + 1 MMFS cross-attention decoder layer of 1 sequence (T=2048, 4 images) in fp32 "
                        "through the oracle restatement; value = 1 / (30*t_plain + 10*t_cross); the visual tokenizer, glue and "
                        "text head (< 3 % of the FLOPs) are not in the CPU sample")

    def cpu_baseline(self):
        threads = min(os.cpu_count() or 1, 64)
        torch.set_num_threads(threads)
        if not hasattr(self, "host_cpu_ready"):
            self.host_cpu_ready = True
        tp, tc = self.reference_step()
        tp, tc = self.reference_step()
        return {"value": 1.0 / (30 * tp + 10 * tc), "unit": self.unit, "cores": threads, "kind": "port",
                "sample": f"1 plain layer ({tp:.2f} s) + 1 MMFS cross layer ({tc:.2f} s) of 1 sequence, fp32, oracle "
                          "restatement of the reference forward; scaled to 30 plain + 10 cross layers"}


class SdCfg4(Workload):
    """BASELINE cfg 4: SD-2.1 UNet 50-step classifier-free-guidance denoise at 512^2 (latents 64^2) with MMFS
    conditioning on one context image, batch 8 per GPU (16 with CFG).  Non-default workload (`--workload sd_cfg4`);
    a step = one full 50-step denoise of the local batch, unit = images/s."""

    name = "sd_cfg4"
    metric = "sd_unet_denoise_images_per_sec"
    unit = "images/s"
    STEPS = 50

    def setup(self):
        import mm_interleaved_b200 as m
        from mm_interleaved_b200 import ops, unet_sd
        self.m, self.ops, self.unet_sd = m, ops, unet_sd
        self.B = self.local_batch or 8
        torch.manual_seed(42 + self.rank)                                 # sd_base_seed: 42 (mm_inference.yaml:23)
        dt = torch.bfloat16
        self.unet = unet_sd.UNet2DConditionModel().to("cuda", dt).eval().to(memory_format=torch.channels_last)
        net = m.MMFSNet(1024, (320, 640, 1280, 1280), 2)
        with torch.no_grad():
            for blk in list(net.mmfs_down_blocks) + [net.mmfs_mid_block]:
                blk.conv.weight.normal_(0, 0.02)                          # zero-initialised in the reference
        self.net = net.to("cuda", dt).eval()
        g = torch.Generator().manual_seed(42)
        self.host = [torch.randn((self.B, 4, 64, 64), generator=g), torch.randn((self.B, 77, 1024), generator=g) * 0.02] + \
                    [torch.randn((self.B, 1, 1024, s, s), generator=g) for s in (64, 32, 16, 8)]
        self.host = [t.to(dt).pin_memory() for t in self.host]
        self.dev = [t.cuda() for t in self.host]
        self.mask = torch.ones((self.B, 1), device="cuda")
        self.out_h = torch.empty((self.B, 4, 64, 64), dtype=dt).pin_memory()
        self._attn_events = []
        orig = ops.attention

        def timed(q, kk, *a, **k):
            if q.shape[1] != 4096 or kk.shape[1] != 4096 or torch.cuda.is_current_stream_capturing():
                return orig(q, kk, *a, **k)                         # only the eager T = 4096 self-attention calls are timed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = orig(q, kk, *a, **k); e1.record()
            self._attn_events.append((e0, e1))
            return r

        unet_sd.ops.attention = timed
        self.reset_counters()

    def reset_counters(self):
        super().reset_counters()
        self._attn_events = []
        if hasattr(self, "ops"):
            self.ops.launch_counter[0] = 0

    def launch_count(self):
        return self.ops.launch_counter[0]

    def units_per_step(self):
        return self.B

    def _run(self, t):
        lat, cond, feats = t[0], t[1], list(t[2:])
        return self.unet_sd.denoise_loop(self.unet, lat, cond, torch.zeros_like(cond), feats, self.mask, self.net,
                                         num_steps=self.STEPS, guidance=7.5)

    def step_device(self):
        self.last = self._run(self.dev)

    def step_e2e(self):
        dev = [t.to("cuda", non_blocking=True) for t in self.host]
        self.last = self._run(dev)
        self.out_h.copy_(self.last, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def result_checksum(self):
        return self.last.float().sum().reshape(1)

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.host)

    def d2h_bytes(self):
        return self.out_h.numel() * self.out_h.element_size()

    def kernel_stats(self):
        torch.cuda.synchronize()
        t = [a.elapsed_time(b) for a, b in self._attn_events]
        return {"launches": len(t), "avg_ms": sum(t) / len(t) if t else None}

    def config(self):
        return {"workload": f"BASELINE cfg4: SD-2.1-base UNet (866 M params, random init) {self.STEPS}-step CFG denoise, latents "
                            f"({2 * self.B},4,64,64), ctx (77,1024), MMFSNet (13 blocks) on 1 context image; DDIM update as scheduler "
                            "stand-in; 3x3/1x1 convs (tcgen05 implicit GEMM), GroupNorm+SiLU, GEGLU, attention, LayerNorm, MMFS = this repo's kernels; conv_in/conv_out + Linear GEMMs = cuDNN/cuBLAS",
                "step_unit": "one 512^2 image (50 UNet evaluations at batch 2 for CFG)", "global_batch": self.B * self.world,
                "parallelism": f"dp{self.world}", "l2": "192 MiB buffer written between timed steps"}

    def roofline(self, kernel):
        peaks = measured_peaks()
        flops = 4.0 * (2 * self.B) * 4096 * 4096 * 320                   # self-attn at T=4096, C=320 (5 x 64), plus kv=77 cross (small)
        t = kernel["avg_ms"] * 1e-3 if kernel["avg_ms"] else None
        return {"kernel": "attn_fwd_kernel<bf16,64> (tcgen05), UNet self-attention at T=4096 (5 heads x 64)", "bound": "tensor",
                "achieved": flops / t / 1e12 if t else None, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": flops / t / 1e12 / peaks["bf16_tflops_sustained"] if t else None,
                "note": "self-attention calls at T = T_kv = 4096 only",
                "avg_launch_us": t * 1e6 if t else None, "launches_timed": kernel["launches"], "traffic": None}

    def cpu_baseline(self):
        return {"value": None, "unit": self.unit, "cores": 0, "kind": "port",
                "sample": "no CPU restatement of the diffusers UNet exists in this repo (parity unpinned, DESIGN.md); "
                          "the MMFS branch alone is covered by oracle/sd_mmfs.py"}


WORKLOADS = {"msda_cfg3": MsdaCfg3, "interleaved_cfg3": InterleavedCfg3, "sd_cfg4": SdCfg4}
AUTO = "interleaved_cfg3"


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
    if name == "interleaved_cfg3":
        threads = min(threads, 64)
        torch.set_num_threads(threads)
    for _ in range(max(min(args.warmup, 2), 1)):
        wl.reference_step()
    t0 = time.time()
    acc = [0.0, 0.0]
    for _ in range(args.steps):
        r = wl.reference_step()
        if wl.reference_step_fraction is None:
            acc[0] += r[0]; acc[1] += r[1]
    dt = time.time() - t0
    if wl.reference_step_fraction is None:
        value = 1.0 / (30 * acc[0] / args.steps + 10 * acc[1] / args.steps)
    else:
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
