"""Pipeline workload of bench.py: the full fusion recipe (PhaseNet + AdaCoF + FusionNet blend) at 1080p,
batch 16 on one B200 (BASELINE.json configs[2]); frame pairs shard across ranks by batch."""
import os
import time

import numpy as np


class PipelineWorkload:
    H, W = 1080, 1920
    B = int(os.environ.get("FVFI_BENCH_BATCH", "16"))
    name = "fusion_pipeline_1080p_batch%d (BASELINE.json configs[2]): PhaseNet + 4x AdaCoFNet + FusionNet" % B
    dtype = "f32"
    cpu_kind = "port"

    def __init__(self, device, seed):
        import torch
        from fvfi.pipeline import FusionPipeline
        from oracle import fusion_pipeline as fp  # seeded synthetic frames / weights only (not timed, not the product path)
        self.torch, self.device = torch, device
        self.tf32 = os.environ.get("FVFI_TF32", "0") == "1"
        torch.backends.cudnn.allow_tf32 = self.tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        torch.backends.cudnn.benchmark = True
        self.pipe = FusionPipeline(self.H, self.W, device, phase_plane_chunk=int(os.environ.get("FVFI_PLANE_CHUNK", "24")))
        self.pipe.max_batch = int(os.environ.get("FVFI_MAX_BATCH", "8" if self.H <= 1080 else "2"))
        self.pipe.load_state(fp.seeded_state(0))
        r1, r2 = fp.seeded_frames(1, self.H, self.W, seed)
        g = torch.Generator().manual_seed(seed)
        # B distinct pairs: per-sample brightness/shift jitter of one smooth synthetic pair
        gains = 0.8 + 0.2 * torch.rand((self.B, 1, 1, 1), generator=g)
        self.h1 = (r1 * gains).clamp(0, 1).contiguous().pin_memory()
        self.h2 = (r2 * gains).clamp(0, 1).contiguous().pin_memory()
        self.out_host = torch.empty((self.B, 3, self.H, self.W)).pin_memory()
        self.d1, self.d2 = self.h1.to(device), self.h2.to(device)
        self.frames_per_step = self.B
        self.launches_per_step = None
        self.stage_ms = {}
        self.nsteps = 0
        self.config_extra = {"tf32_convs": self.tf32, "convs": "tcgen05 implicit GEMM, %s operand split, persistent (csrc/conv_tc.cu)" % os.environ.get("FVFI_CONV_PREC", "f16x3"),
                             "phase_plane_chunk": self.pipe.phase_net.plane_chunk,
                             "sub_batch": self.pipe.max_batch}

    def step(self, timed=False):
        self.pipe.timing = [] if timed else None
        self.out = self.pipe(self.d1, self.d2)
        if timed:
            self._pending = getattr(self, "_pending", [])
            self._pending.append(self.pipe.timing)
            self.pipe.timing = None

    def _collect(self):
        for tl in getattr(self, "_pending", []):
            for (n0, e0), (n1, e1) in zip(tl[:-1], tl[1:]):
                if n1 == 'start':
                    continue  # boundary between two sub-batches
                self.stage_ms[n1] = self.stage_ms.get(n1, 0.0) + e0.elapsed_time(e1)
            self.nsteps += 1
        self._pending = []

    def _peaks(self):
        import json
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
        if os.path.exists(path):
            d = json.load(open(path))
            return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1399.0))), "measured sustained bf16 (MEASURED_PEAKS.json)"
        return 1400.0, "fallback (B200_PROFILING.md, sustained)"

    def _time_conv(self):
        """Per-launch CUDA events around every tcgen05 convolution of ONE extra pipeline step: the step's dominant
        kernel (conv_split_kernel, ~60 % of the step).  achieved = algorithmic FLOPs (2*B*H*W*Cin*Cout*K*K) / time."""
        from fvfi import conv as tc
        torch = self.torch
        tc.timing = []
        self.pipe.timing = None
        self.pipe(self.d1, self.d2)
        torch.cuda.synchronize()
        rec, tc.timing = tc.timing, None
        flops = sum(r[0] for r in rec)
        ms = sum(r[1].elapsed_time(r[2]) for r in rec)
        return flops, ms, sum(r[3] for r in rec)

    def _time_hbm_kernels(self, peak):
        """The hand-written HBM-bound kernels of the step, each timed alone with CUDA events on the step's shapes."""
        torch = self.torch
        from fvfi import adacof
        B, H, W = min(self.B, self.pipe.max_batch), 1088, 1920
        g = torch.Generator(device=self.device).manual_seed(0)
        mk = lambda *s: torch.rand(s, device=self.device, generator=g)
        i1, i2 = mk(B, 3, H + 4, W + 4), mk(B, 3, H + 4, W + 4)
        w1 = torch.softmax(mk(B, 25, H, W), 1)
        a1, b1 = mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5
        occ = mk(B, 1, H, W)

        def timeit(fn, reps=5):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        out = []
        ms = timeit(lambda: adacof.adacofnet_warp_blend(i1, i2, w1, a1, b1, w1, b1, a1, occ, 1, want_t=False))
        px = B * H * W
        nbytes = 4 * (6 * 25 * px + 2 * 3 * B * (H + 4) * (W + 4) + px + 3 * px + px)
        out.append({"kernel": "adacof_fwd_tma<2,2,3> (two warps + blend + uncertainty, TMA-streamed coefficients, smooth offsets)", "bound": "hbm",
                    "achieved": round(nbytes / ms / 1e6, 1), "unit": "GB/s", "frac": round(nbytes / ms / 1e6 / peak, 4),
                    "ms_per_launch": round(ms, 4), "algorithmic_bytes_per_launch": nbytes})
        del i1, i2, w1, a1, b1, occ
        N = 12 * B
        x = mk(N, self.H, self.W)
        pyr = self.pipe.pyr
        vals = pyr.filter(x, want_high=False)
        msd = timeit(lambda: pyr.filter(x, want_high=False), 3)
        msr = timeit(lambda: pyr.inv_filter_sparse(vals, use_high=False), 3)
        pb = 72 * self.H * self.W * N - 4 * self.H * self.W * N      # 72 HW per plane-op minus the skipped high residual
        for name, t in (("pyramid decompose (k_rows_fwd/k_cols_fwd + per level k_cols_inv_decomp/k_rows_inv)", msd),
                        ("pyramid reconstruct (per level k_rows_fwd/k_cols_fwd + k_cols_inv_gather/k_rows_inv)", msr)):
            out.append({"kernel": name, "bound": "hbm", "achieved": round(pb / t / 1e6, 1), "unit": "GB/s",
                        "frac": round(pb / t / 1e6 / peak, 4), "ms_per_call": round(t, 4), "planes": N,
                        "algorithmic_bytes_per_call": pb})
        return out

    def roofline(self, peak, peak_src):
        """Dominant kernel of the step = the tcgen05 convolution (tensor-bound).  `achieved` counts ALGORITHMIC FLOPs
        (one multiply-add per tap, channel pair and pixel); every one of them is executed as three fp16 tensor-core
        products (3xFP16 split), so the tensor pipe runs 3x that rate (`mma_frac`).  The hand-written HBM-bound kernels
        of the step (fused AdaCoF synthesis, pyramid) follow in `other_kernels`."""
        self._collect()
        tpeak, tsrc = self._peaks()
        flops, ms, launches = self._time_conv()
        ach = flops / (ms * 1e-3) / 1e12
        stages = {k: round(v / max(self.nsteps, 1), 3) for k, v in self.stage_ms.items()}
        step_ms = sum(stages.values())
        return {"bound": "tensor", "kernel": "conv_split_kernel<ACT, PREC_F16X3> (all %d launches of one step)" % launches,
                "achieved": round(ach, 1), "peak": tpeak, "unit": "TFLOP/s", "frac": round(ach / tpeak, 4),
                "mma_frac": round(3 * ach / tpeak, 4), "traffic": None, "peak_source": tsrc,
                # `achieved` aggregates ~660 launches of different shapes, so there is no single per-launch DRAM figure; one
                # representative launch from the ncu --set full capture (profiles/r01_conv_v7_summary.txt):
                "traffic_example": {"launch": "64->64 3x3 ReLU @544x960, batch 4", "dram_bytes": 1020526336,
                                    "algorithmic_bytes": 1069842432, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum"},
                "ms_per_step_in_kernel": round(ms, 2), "share_of_step": round(ms / max(step_ms, 1e-9), 3),
                "algorithmic_flops_per_step": flops,
                "other_kernels": self._time_hbm_kernels(peak), "hbm_peak": peak, "hbm_peak_source": peak_src,
                "stage_ms_per_step": stages}

    def e2e(self, steps):
        torch = self.torch
        self.pipe.timing = None
        self.pipe.interpolate_host(self.h1, self.h2, self.out_host)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.pipe.interpolate_host(self.h1, self.h2, self.out_host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        nb = lambda t: t.numel() * 4
        ok = bool(torch.equal(self.out_host, self.out.cpu()))
        return dt, nb(self.h1) + nb(self.h2), nb(self.out_host), ok

    @classmethod
    def cpu_sample(cls, threads, seed=0):
        """CPU oracle of the same recipe (oracle/fusion_pipeline.py: restated reference modules, scipy filters)
        on ONE frame pair at 272x480 (1/15.9 of the 1080p area); frames/s is scaled by the pixel ratio."""
        import torch
        from oracle import fusion_pipeline as fp
        torch.set_num_threads(threads)
        H, W = 272, 480
        be = fp.oracle_backend(fp.seeded_state(0), hw=(H, W), threads=threads)
        r1, r2 = fp.seeded_frames(1, H, W, seed)
        t0 = time.perf_counter()
        fp.interp(be, r1, r2)
        dt = time.perf_counter() - t0
        scale = (H * W) / float(cls.H * cls.W)
        return scale / dt, dt, ("1 frame pair at %dx%d (%.4f of 1080p area), frames/s scaled by the pixel ratio; "
                                "oracle port of the reference recipe, torch CPU + scipy + C warp" % (H, W, scale))


class Pipeline4KWorkload(PipelineWorkload):
    """BASELINE.json configs[3]: the same recipe at 3840x2160 (pyramid height 19, AdaCoFNet pads to 2176 rows); frame pairs
    shard across ranks exactly like the 1080p workload.  Not the default bench line (the metric is quoted at 1080p)."""
    H, W = 2160, 3840
    B = int(os.environ.get("FVFI_BENCH_BATCH_4K", "4"))
    name = "fusion_pipeline_4k_batch%d (BASELINE.json configs[3])" % B

    @classmethod
    def cpu_sample(cls, threads, seed=0):
        fps, dt, sample = PipelineWorkload.cpu_sample(threads, seed)
        return fps * (PipelineWorkload.H * PipelineWorkload.W) / float(cls.H * cls.W), dt, sample.replace("1080p", "4K-scaled 1080p")
