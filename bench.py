#!/usr/bin/env python
"""Headline benchmark: 3 s @ 16 kHz clips/s, ESC encode + decode (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = ``model.encode(x, 6)`` then ``model.decode(codes, feat_shape)`` over one batch of 36 synthetic 3 s
clips per GPU (BASELINE configs[1]: ESC-Base 9 kbps, batch 36; configs[4] is the same 36 clips per GPU on 8 GPUs,
so the scaling is weak) followed, for N > 1, by the one NCCL all-gather of codes + reconstructed audio the north
star names.  Prints ONE JSON line (rank 0).

* ``value``      device-resident throughput: inputs already in HBM, CUDA-event timed, L2 flushed between steps.
* ``e2e``        the same metric through the public API with HOST tensors: pinned host -> H2D -> kernels -> D2H
                 inside the timed region (``escb_encode_host`` / ``escb_decode_host`` behind ``ESC.encode/decode``).
* ``roofline``   the dominant kernel class, timed live with CUDA events around every launch of a profiled pass
                 (``escb_profile_begin/end``), algorithmic flops / bytes as defined in DESIGN.md.
* ``rvq``        the "RVQ argmin HBM GB/s vs peak" half of the metric: argmin-only and the three PVQ kernels.
* ``cpu_baseline`` the CPU oracle port (reference algorithm, fp32 ATen) on this box's host cores, bounded sample.

``--impl reference`` times that CPU port alone (the reference is pure Python on PyTorch CPU ops and cannot travel
to the GPU box; oracle/esc_oracle.py is its pinned restatement).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)

BASE = dict(backbone="transformer", in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6,
            win_len=20, hop_len=5, sr=16000, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2,
            window_size=4, mlp_ratio=4.0, overlap=2, group_size=3, codebook_size=1024,
            codebook_dims=[32, 32, 16, 12, 8, 6], l2norm=True)
LARGE = dict(BASE, swin_depth=4, codebook_dims=[8] * 6)
CLIP_SAMPLES = 48000
GFLOP_PER_CLIP = {"base": 56.73, "large": 99.23}          # 1x multiply-add count, S=6
METRIC = "clips_per_sec_3s_16khz_encode_decode"
UNIT = "clips/s"

# The contract is ONE JSON line on stdout.  Libraries chat on fd 1 (NCCL prints its version banner there), so the real
# stdout is kept aside for the result line and fd 1 is pointed at stderr for everything else.
_RESULT = None


def claim_stdout():
    global _RESULT
    if _RESULT is None:
        sys.stdout.flush()
        _RESULT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    tensor_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def build_cpu_reference(cfg):
    """(encode+decode callable, kind): the REAL reference staged under baseline/_ref (oracle/ref_loader.py) when it
    travelled to this machine, else the pinned oracle port."""
    import torch
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    from oracle import ref_loader
    sd = synth_state_dict(CodecSpec.from_kwargs(**cfg), 0)
    model = None
    try:
        model = ref_loader.make_reference_model(cfg, sd)
    except Exception as e:                                   # a broken staging must not kill the bench line
        print(f"bench.py: reference import failed ({e!r}); timing the oracle port instead", file=sys.stderr)
    if model is not None:
        def run(x):
            with torch.no_grad():
                codes, fs = model.encode(x, 6)
                return codes, model.decode(codes, fs)
        return run, "reference", "yzGuu830/efficient-speech-codec esc.models.make_model(...).encode/decode (baseline/_ref), fp32 ATen"
    from oracle.esc_oracle import EscOracle
    o = EscOracle(cfg, sd)

    def run(x):
        codes, fs = o.encode(x, 6)
        return codes, o.decode(codes, fs)
    return run, "port", "oracle/esc_oracle.py (pinned restatement of the reference), fp32 ATen"


def cpu_reference(cfg_name, budget_s, steps=None, warmup=1):
    """Time the reference's CPU path (encode + decode, S=6) on the host cores; returns (clips/s, ms/step, info)."""
    import torch
    from escb200.synthetic import synth_audio
    cfg = BASE if cfg_name == "base" else LARGE
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind, what = build_cpu_reference(cfg)

    t0 = time.perf_counter()
    run(synth_audio(1, CLIP_SAMPLES, seed=0))               # warm-up + per-clip cost estimate
    t1 = time.perf_counter() - t0
    if steps is None:
        sample, steps = 4, max(1, min(6, int(budget_s / max(4 * t1, 1e-3))))
    else:
        sample = max(1, min(8, int(budget_s / max((steps + warmup) * t1, 1e-3))))
    x = synth_audio(sample, CLIP_SAMPLES, seed=0)
    for _ in range(max(0, warmup - 1)):
        run(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(x)
    dt = time.perf_counter() - t0
    info = {"cores": cores, "kind": kind, "threads": torch.get_num_threads(),
            "sample": f"{steps} x (encode+decode of {sample} synthetic 3 s clips, S=6; clips/s does not depend on the "
                      f"sample's batch), {what}"}
    return sample * steps / dt, dt / steps * 1e3, info


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, info = cpu_reference(args.config, 150.0, steps=args.steps, warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ESC-{args.config} 9kbps encode+decode, 3 s synthetic clips, num_streams=6 (CPU sample)"},
            "cpu_baseline": dict(info, value=value, unit=UNIT),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ B200 arm helpers
def probe_tf32_peak(dev, seconds=0.6):
    """Dense TF32 matmul rate of this GPU (cuBLAS through torch.matmul, allow_tf32=True, 8192^3, back to back)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters, best, t_end = 0, 0.0, time.perf_counter() + seconds
        tot_ms = 0.0
        while time.perf_counter() < t_end:
            e0.record()
            for _ in range(5):
                a @ b
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            tot_ms += ms
            iters += 5
            best = max(best, 5 * 2.0 * n ** 3 / (ms * 1e-3) / 1e12)
        return {"burst": best, "sustained": iters * 2.0 * n ** 3 / (tot_ms * 1e-3) / 1e12,
                "how": "torch.matmul fp32 8192^3, torch.backends.cuda.matmul.allow_tf32=True, CUDA events"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def timed_ms(fn, iters, flush, dev):
    """Mean device ms of fn() over `iters` calls, CUDA events on the current stream, L2 flushed between calls."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return sum(a.elapsed_time(b) for a, b in ev) / iters


def rvq_microbench(model, dev, flush, peaks, iters=20):
    """BASELINE configs[3]: 1024 VQ frames (B=1, W=2048) through the six stream steps of ESC-Base, via the C ABI's
    unit entry points.  (A) argmin only on pre-projected vectors; (B) the stream step enc, dec -> codes, dec_refine."""
    import torch
    from escb200 import native
    lib, h, spec = native.lib(), model._handle(dev), model.spec
    W, B = 2048, 1
    T = W // spec.overlap
    ws = model._ws(dev, h.workspace_bytes(B, W))
    st = model._stream(dev)
    g = torch.Generator().manual_seed(100)
    streams = []
    for q, qs in enumerate(spec.quantizers()):
        C, Hq, d = qs.in_dim, qs.in_freq, qs.codebook_dim
        enc = torch.randn(B, Hq * W, C, generator=g).to(dev)
        dec = None if q == 0 else torch.randn(B, Hq * W, C, generator=g).to(dev)
        z = [torch.randn(T, d, generator=g).to(dev) for _ in range(3)]
        streams.append(dict(q=q, C=C, Hq=Hq, d=d, enc=enc, dec=dec, z=z, out=torch.empty_like(enc),
                            codes=torch.empty((B, 3, T), dtype=torch.int64, device=dev),
                            idx=torch.empty((T,), dtype=torch.int64, device=dev)))

    def argmin_only():
        for s in streams:
            for grp in range(3):
                native.check(lib.escb_codebook_argmin(h.ptr, s["q"], grp, native.ptr(s["z"][grp]), T, native.ptr(s["idx"]), st))

    def stream_steps():
        for s in streams:
            native.check(lib.escb_pvq_stream(h.ptr, s["q"], native.ptr(s["enc"]), native.ptr(s["dec"]), B, W,
                                             native.ptr(s["codes"]), native.ptr(s["out"]), native.ptr(ws), ws.numel(), st))

    for _ in range(3):
        argmin_only()
        stream_steps()
    l0 = h.launch_count()
    ms_a = timed_ms(argmin_only, iters, flush, dev)
    l1 = h.launch_count()
    ms_b = timed_ms(stream_steps, iters, flush, dev)
    l2 = h.launch_count()
    sum_d = sum(s["d"] for s in streams)
    bytes_a = T * (3 * 4 * sum_d + 6 * 3 * 8)                                   # 1416 B / frame
    flops_a = T * 2.0 * 3 * spec.codebook_size * sum_d                          # 0.651 MFLOP / frame
    frame = [2 * s["C"] * s["Hq"] for s in streams]
    bytes_b = T * (4 * (sum(frame) * 2 - frame[0]) + 4 * sum(frame) + 6 * 3 * 8)   # 169 104 B / frame
    flops_b = flops_a + T * sum(2.0 * 2 * f * s["d"] for f, s in zip(frame, streams))
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    return {"workload": "ESC-Base RVQ only: 6 streams x 3 groups, 1024 VQ frames (B=1, W=2048), L2 flushed between iterations",
            "argmin_only": {"ms": ms_a, "launches": (l1 - l0) // iters, "algorithmic_bytes": bytes_a,
                            "achieved": bytes_a / (ms_a * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": bytes_a / (ms_a * 1e-3) / 1e9 / peaks["hbm"],
                            "tflops": flops_a / (ms_a * 1e-3) / 1e12, "fp32_fma_peak_tflops_nominal": fp32_peak,
                            "note": "460 flop/B: FP32-FMA / launch bound by construction, the HBM fraction is reported because the metric names it"},
            "stream_step": {"ms": ms_b, "launches": (l2 - l1) // iters, "algorithmic_bytes": bytes_b,
                            "achieved": bytes_b / (ms_b * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": bytes_b / (ms_b * 1e-3) / 1e9 / peaks["hbm"], "tflops": flops_b / (ms_b * 1e-3) / 1e12,
                            "bound": "hbm", "api": "escb_pvq_stream: one fused launch per stream step (enc, dec -> codes, dec_refine)"}}


def small_batch_latency(model, dev, batch=1, iters=50):
    """Latency of encode + decode of `batch` 3 s clips: the ~190 launches issued eagerly, and
    the same launches captured once in a CUDA graph and replayed (the stream chain is sequential, so at B=1 the step is
    launch-latency bound)."""
    import torch
    from escb200.synthetic import synth_audio
    x = synth_audio(batch, CLIP_SAMPLES, seed=77).to(dev)

    def step():
        codes, fs = model.encode(x, 6)
        return codes, model.decode(codes, fs)
    for _ in range(3):
        ref_codes, ref_audio = step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    eager_ms = e0.elapsed_time(e1) / iters
    out = {"batch": batch, "eager_ms": eager_ms, "eager_clips_per_s": batch / (eager_ms * 1e-3)}
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            step()
            with torch.cuda.graph(g, stream=s):
                g_codes, g_audio = step()
        torch.cuda.current_stream(dev).wait_stream(s)
        g.replay()
        torch.cuda.synchronize(dev)
        same = bool(torch.equal(g_codes, ref_codes) and torch.equal(g_audio, ref_audio))
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        graph_ms = e0.elapsed_time(e1) / iters
        out.update({"graph_ms": graph_ms, "graph_clips_per_s": batch / (graph_ms * 1e-3), "graph_equals_eager": same,
                    "how": "torch.cuda.graph around ESC.encode + ESC.decode (the C-ABI calls launch on the capturing stream)"})
    except Exception as e:                               # capture is an optimisation, not part of the contract
        out["graph_error"] = repr(e)[:200]
    return out


def flip_census(cfg, dev, ref_model, my_model, batches=4, per=36):
    """Code decisions that differ from the reference's (CPU, fp32) over `batches` x `per` fresh clips, for three
    implementations of the same fp32 arithmetic: the reference itself run on this GPU, this library's tcgen05 engine
    (the product path) and this library's fp32 SIMT engine.  A clip counts once: after its first differing decision the
    residual chain diverges, so later codes of the same clip are not independent events."""
    import torch
    from escb200.codec import ESC
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_audio, synth_state_dict
    old_env = os.environ.get("ESCB_GEMM")
    os.environ["ESCB_GEMM"] = "simt"                      # read when the native handle is created
    try:
        simt = ESC(**cfg)
        simt.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**cfg), 0))
        simt = simt.eval().to(dev)
        simt.encode(synth_audio(1, CLIP_SAMPLES, seed=1).to(dev), 6)
    finally:
        if old_env is None:
            del os.environ["ESCB_GEMM"]
        else:
            os.environ["ESCB_GEMM"] = old_env
    cpu_model = ref_model.to("cpu")
    ref_cpu = []
    xs = [synth_audio(per, CLIP_SAMPLES, seed=1000 + b) for b in range(batches)]     # batch 0 = the bench input of rank 0
    with torch.no_grad():
        for x in xs:
            ref_cpu.append(cpu_model.encode(x, 6)[0])
        gpu_model = ref_model.to(dev)
        out = {"clips": batches * per, "codes_per_clip": int(ref_cpu[0][0].numel()),
               "what": "clips with at least one code index different from the reference on the CPU (fp32 ATen)"}
        for name, enc in (("reference_on_this_gpu", lambda x: gpu_model.encode(x, 6)[0]),
                          ("escb200_tcgen05", lambda x: my_model.encode(x, 6)[0]),
                          ("escb200_fp32_simt", lambda x: simt.encode(x, 6)[0])):
            clips, codes = 0, 0
            for x, rc in zip(xs, ref_cpu):
                bad = enc(x.to(dev)).cpu() != rc
                clips += int((bad.flatten(1).sum(1) > 0).sum())
                codes += int(bad.sum())
            out[name] = {"clips_with_a_flip": clips, "codes_differing": codes}
    return out


def incumbent_and_noise_floor(cfg, dev, x_dev, my_codes, my_audio, steps, my_model=None):
    """The reference itself in eager PyTorch ON THIS GPU (fp32, TF32 off for matmul and cuDNN): the practical incumbent,
    since the reference has no native kernels - plus the parity noise floor reference-CPU vs reference-GPU."""
    import torch
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    from oracle import ref_loader
    try:
        model = ref_loader.make_reference_model(cfg, synth_state_dict(CodecSpec.from_kwargs(**cfg), 0))
    except Exception as e:
        return {"unavailable": f"reference import failed: {e!r}"}
    if model is None:
        return {"unavailable": "baseline/_ref is not staged on this machine (oracle/ref_loader.py stage)"}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        B = x_dev.shape[0]
        with torch.no_grad():
            n_cpu = B                                    # ~10 clips/s on the host: a few seconds, every clip attributed
            c_cpu, fs = model.encode(x_dev[:n_cpu].cpu(), 6)
            a_cpu = model.decode(c_cpu, fs)
            gm = model.to(dev)

            def step():
                c, f = gm.encode(x_dev, 6)
                return c, gm.decode(c, f)
            for _ in range(2):
                c_gpu, a_gpu = step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(2, min(steps, 5))
            e0.record()
            for _ in range(n):
                c_gpu, a_gpu = step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / n
        census = None
        if my_model is not None:
            try:
                census = flip_census(cfg, dev, model, my_model)
            except Exception as e:
                census = {"error": repr(e)[:200]}
        return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B, "flip_census": census,
                "what": "reference esc.ESC.encode+decode, eager PyTorch on this B200, fp32, allow_tf32=False (matmul and cuDNN)",
                "noise_floor": {
                    "clips": n_cpu, "codes_per_clip": int(c_cpu[0].numel()),
                    "ref_cpu_vs_ref_gpu_code_mismatches": int((c_cpu != c_gpu[:n_cpu].cpu()).sum()),
                    "ref_cpu_vs_ref_gpu_audio_max_abs": float((a_cpu - a_gpu[:n_cpu].cpu()).abs().max()),
                    "escb200_vs_ref_cpu_code_mismatches": int((c_cpu != my_codes[:n_cpu].cpu()).sum()),
                    "escb200_vs_ref_cpu_audio_max_abs": float((a_cpu - my_audio[:n_cpu].cpu()).abs().max()),
                    "escb200_vs_ref_gpu_code_mismatches_all_clips": int((c_gpu != my_codes).sum())}}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


# ------------------------------------------------------------------------------------------------ B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    from escb200.codec import ESC
    from escb200.parallel import gather_results, shard_bounds
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_audio, synth_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; esc-b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = world
    cfg = BASE if args.config == "base" else LARGE
    B, S = args.batch, 6
    spec = CodecSpec.from_kwargs(**cfg)
    model = ESC(**cfg)
    model.load_state_dict(synth_state_dict(spec, 0))
    model = model.eval().to(dev)

    x_host = synth_audio(B, CLIP_SAMPLES, seed=1000 + rank).pin_memory()
    x_dev = x_host.to(dev)
    W = model.time_patches(CLIP_SAMPLES)
    n_out = spec.decoded_samples(W)
    g_codes = g_audio = None
    if N > 1:
        g_codes = torch.empty((N * B, S, 3, W // 2), dtype=torch.int64, device=dev)
        g_audio = torch.empty((N * B, n_out), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # 2x the 126 MB L2

    def step_device(xd=x_dev, s=S, gather=True):
        codes, fs = model.encode(xd, s)
        audio = model.decode(codes, fs)
        if N > 1 and gather:
            gather_results(codes, audio, g_codes, g_audio)       # the one NCCL all-gather of the north star
        return codes, audio

    def step_host():
        codes, fs = model.encode(x_host, S)          # CPU tensors: pinned H2D + kernels + D2H inside
        audio = model.decode(codes, fs)
        if N > 1:                                    # same workload as `value`: the gather of the device-side results
            gather_results(codes.to(dev, non_blocking=True), audio.to(dev, non_blocking=True), g_codes, g_audio)
        return codes, audio

    def barrier():
        torch.cuda.synchronize(dev)
        if N > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if N == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn, k):
        """K steps, each bracketed by CUDA events, L2 flushed (outside the event pairs) between steps; max over ranks."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))

    h = model._handle(dev)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed region
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = h.launch_count()
    total_ms = timed_steps(step_device, args.steps)
    launches = h.launch_count() - launches0
    clocks = sampler.finish()
    value = N * B * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API with host tensors
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    codes_bytes = B * S * 3 * (W // 2) * 8
    e2e = {"value": N * B * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": B * CLIP_SAMPLES * 4 + codes_bytes, "d2h_bytes_per_step": codes_bytes + B * n_out * 4,
           "api": "ESC.encode(x_cpu, 6); ESC.decode(codes_cpu, feat_shape) -> escb_encode_host / escb_decode_host"
                  + ("; + escb200.parallel.gather_results (NCCL)" if N > 1 else "")}

    # ---- multi-rank result check: rank 0 recomputes the last rank's shard and compares it with its slice of the gather
    gather_check = None
    if N > 1:
        codes_l, audio_l = step_device()
        barrier()
        if rank == 0:
            r = N - 1
            lo, hi = shard_bounds(N * B, r, N)
            xr = synth_audio(B, CLIP_SAMPLES, seed=1000 + r).to(dev)
            cr, ar = step_device(xr, S, gather=False)
            ok = bool(torch.equal(g_codes[lo:hi], cr) and torch.equal(g_audio[lo:hi], ar)
                      and torch.equal(g_codes[:B], codes_l) and torch.equal(g_audio[:B], audio_l))
            gather_check = {"ok": ok, "what": f"rank 0 recomputed rank {r}'s shard: its codes and audio equal rows [{lo}, {hi}) "
                                              "of the all-gathered tensors bit for bit; rank 0's own rows too"}
            if not ok:
                raise SystemExit("bench.py: all-gathered results differ from a local recomputation")
        barrier()

    # ---- strong scaling (BASELINE configs[4]): the same 288 clips whatever N is
    strong = None
    if args.config == "base" and args.strong_batch > 0 and args.strong_batch % N == 0:
        Bs = args.strong_batch // N
        lo, hi = shard_bounds(args.strong_batch, rank, N)
        xs = torch.cat([synth_audio(1, CLIP_SAMPLES, seed=5000 + i) for i in range(lo, hi)]).to(dev)
        gc = torch.empty((args.strong_batch, S, 3, W // 2), dtype=torch.int64, device=dev) if N > 1 else None
        ga = torch.empty((args.strong_batch, n_out), dtype=torch.float32, device=dev) if N > 1 else None

        def step_strong():
            c, fs = model.encode(xs, S)
            a = model.decode(c, fs)
            if N > 1:
                gather_results(c, a, gc, ga)
        for _ in range(2):
            step_strong()
        k = max(2, min(args.steps, 5))
        ms = timed_steps(step_strong, k)
        strong = {"global_batch": args.strong_batch, "per_gpu_batch": Bs, "value": args.strong_batch * k / (ms * 1e-3),
                  "unit": UNIT, "ms_per_step": ms / k, "steps": k, "scaling": "strong",
                  "what": "BASELINE configs[4]: 288 clips sharded over the N ranks + one all-gather of codes and audio"}
        del xs, gc, ga

    if N > 1:
        launches_t = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(launches_t)
        launches = int(launches_t.item())
    if rank != 0:
        if N > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ================================================================================ rank 0 only from here
    peaks = load_peaks()
    tf32 = probe_tf32_peak(dev)
    peaks["tf32"] = tf32["sustained"]

    # ---- per-kernel-class timing (profiled pass: CUDA events around every launch on the launch stream)
    h.profile_begin()
    for _ in range(args.steps):
        flush.zero_()
        model.decode(*model.encode(x_dev, S))
    prof = h.profile_end()
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k]["ms"])
    tv = prof[top]
    tflops = tv["flops"] / (tv["ms"] * 1e-3) / 1e12
    gbs = tv["bytes"] / (tv["ms"] * 1e-3) / 1e9
    ai = tv["flops"] / max(tv["bytes"], 1.0)
    # Every tcgen05 class computes fp32-grade products as 3 TF32 MMAs, so its tensor ceiling is a third of the TF32
    # peak and its ridge is (TF32 peak / 3) / HBM; the SIMT classes are bounded by the fp32 FMA rate.
    simt = top in ("stft_gemm", "istft_gemm", "pvq_down_gemm", "codebook_argmin", "patch_embed", "deembed_conv3x3",
                   "window_attention", "vq_loss", "layout")
    ceil_t = (148 * 128 * 2 * 1.965e9 / 1e12) if simt else peaks["tf32"] / 3.0
    ridge = ceil_t * 1e3 / peaks["hbm"]
    if ai >= ridge:
        roof = {"bound": "tensor", "achieved": tflops, "peak": peaks["tf32"], "unit": "TFLOP/s", "frac": tflops / peaks["tf32"],
                "frac_of_3xtf32_ceiling": tflops / (peaks["tf32"] / 3.0),
                "peak_source": "TF32 dense matmul rate probed in this run (sustained); the honest ceiling of a 3xTF32 kernel is a third of it"}
    else:
        roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                "peak_source": peaks["src"] + " MEASURED_PEAKS.json hbm copy"}
    traffic, traffic_src = None, None
    for tf in ("r2d_traffic.json", "r2b_traffic.json", "r2_traffic.json", "r1_traffic.json"):
        try:                                 # measured DRAM bytes per launch of that class (tools/ncu_traffic.py over one step)
            tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if args.config == "base" and B == 36 and top in tj["classes"]:
                traffic = tj["classes"][top]["dram_bytes_per_launch"]
                traffic_src = f"profiles/{tf}: ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over one step's launches of this class (a committed capture of this build's step, not measured in this run)"
                break
        except (OSError, ValueError, KeyError):
            pass
    whole_tflops = value / N * GFLOP_PER_CLIP[args.config] / 1e3
    roof.update({"traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": tv["bytes"] / max(tv["launches"], 1),
                 "algorithmic_flops_per_launch": tv["flops"] / max(tv["launches"], 1),
                 "kernel": top, "launches": tv["launches"], "avg_launch_ms": tv["ms"] / max(tv["launches"], 1),
                 "share_of_step": tv["ms"] / tot_ms, "flop_per_byte": ai, "ridge_flop_per_byte": ridge,
                 "achieved_tflops": tflops, "achieved_gbs": gbs,
                 "tf32_peak_tflops": tf32, "bf16_peak_tflops_measured": peaks["tensor"], "hbm_peak_gbs": peaks["hbm"],
                 "whole_step_tflops": whole_tflops,
                 "whole_step_frac_of_tf32_peak": whole_tflops / peaks["tf32"],
                 "whole_step_frac_of_3xtf32_ceiling": whole_tflops / (peaks["tf32"] / 3.0),
                 "how": "escb_profile_begin/end: CUDA events around every launch on the launching stream, separate pass of the same K steps"})
    breakdown = {k: {"share": round(v["ms"] / tot_ms, 4), "ms_per_step": round(v["ms"] / args.steps, 4),
                     "launches_per_step": v["launches"] // args.steps,
                     "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 3),
                     "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if v["launches"]}

    def rvq_entry(names):
        ms = sum(prof[n]["ms"] for n in names if n in prof)
        by = sum(prof[n]["bytes"] for n in names if n in prof)
        return {"achieved": by / max(ms, 1e-9) / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": by / max(ms, 1e-9) / 1e6 / peaks["hbm"], "ms_per_step": ms / args.steps}
    rvq = {"in_step_b36": {"argmin_only": rvq_entry(["codebook_argmin"]) if prof.get("codebook_argmin", {}).get("launches") else None,
                           "stream_step": rvq_entry(["pvq_down_gemm", "codebook_argmin", "pvq_up_gemm", "pvq_stream_fused"])},
           "note": "argmin-only is FMA-issue bound by construction (460 flop/B, SURVEY 8d); the stream step is the HBM-bound one"}
    extra = {}
    if not args.quick:
        if args.config == "base":
            rvq["config4_1024_frames"] = rvq_microbench(model, dev, flush, peaks)
        # ---- BASELINE configs[1] "all 6 bitrates swept": clips/s per num_streams
        sweep = {}
        for s in range(1, 7):
            for _ in range(2):
                step_device(x_dev, s, gather=False)
            k = max(3, min(args.steps, 10))
            ms = timed_steps(lambda s=s: step_device(x_dev, s, gather=False), k) if N == 1 else None
            if ms is not None:
                sweep[str(s)] = {"kbps": 1.5 * s, "value": B * k / (ms * 1e-3), "ms_per_step": ms / k}
        extra["num_streams_sweep"] = {"unit": UNIT, "batch": B, "per_num_streams": sweep}
        # ---- small-batch latency: one clip, eager launches vs CUDA graph replay
        if N == 1:
            extra["latency_b1"] = small_batch_latency(model, dev, 1)
        # ---- the reference in eager PyTorch on this GPU + parity noise floor
        my_codes, my_audio = step_device(x_dev, S, gather=False)
        extra["incumbent"] = incumbent_and_noise_floor(cfg, dev, x_dev, my_codes, my_audio, args.steps,
                                                       my_model=model if (args.config == "base" and N == 1) else None)
        # ---- BASELINE configs[2]: ESC-Large, batch 64
        if args.config == "base" and N == 1:
            del my_codes, my_audio
            lspec = CodecSpec.from_kwargs(**LARGE)
            lm = ESC(**LARGE)
            lm.load_state_dict(synth_state_dict(lspec, 0))
            lm = lm.eval().to(dev)
            xl = synth_audio(64, CLIP_SAMPLES, seed=2000).to(dev)

            def step_large():
                c, fs = lm.encode(xl, 6)
                return lm.decode(c, fs)
            for _ in range(3):
                step_large()
            k = max(3, min(args.steps, 5))
            ms = timed_steps(step_large, k)
            lt = 64 * k / (ms * 1e-3) * GFLOP_PER_CLIP["large"] / 1e3
            extra["large_b64"] = {"workload": "BASELINE configs[2]: ESC-Large 9kbps, batch 64 x 3 s clips, num_streams=6, encode+decode",
                                  "value": 64 * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k, "steps": k,
                                  "whole_step_tflops": lt, "whole_step_frac_of_tf32_peak": lt / peaks["tf32"],
                                  "whole_step_frac_of_3xtf32_ceiling": lt / (peaks["tf32"] / 3.0)}
            del lm, xl

    cpu_v, _, cpu_info = cpu_reference(args.config, args.cpu_budget)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ESC-{args.config} 9kbps, batch {B} x 3 s synthetic clips per GPU, num_streams=6, encode+decode"
                                   + (", + NCCL all-gather of codes and audio" if N > 1 else ""),
                       "per_gpu_batch": B, "global_batch": N * B, "num_streams": S,
                       "l2": "256 MiB buffer written between timed steps (outside the event pairs)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "rvq": rvq,
            "cpu_baseline": dict(cpu_info, value=cpu_v, unit=UNIT), "kernels": breakdown}
    if gather_check is not None:
        line["gather_check"] = gather_check
    if strong is not None:
        line["strong_scaling"] = strong
    line.update(extra)
    emit(line)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="base", choices=["base", "large"])
    ap.add_argument("--batch", type=int, default=36, help="clips per GPU")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--strong-batch", type=int, default=288, help="global batch of the strong-scaling line (0: skip)")
    ap.add_argument("--quick", action="store_true", help="headline + roofline only (skip the sweep / microbench / incumbent / Large sections)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU (the driver uses torchrun itself)
        import subprocess
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)


if __name__ == "__main__":
    main()
