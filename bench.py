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
def cpu_reference(cfg_name, budget_s, steps=None, warmup=1):
    """Time the oracle port (encode + decode, S=6) on the host cores; returns (clips/s, info dict)."""
    import torch
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_audio, synth_state_dict
    from oracle.esc_oracle import EscOracle
    cfg = BASE if cfg_name == "base" else LARGE
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = EscOracle(cfg, synth_state_dict(CodecSpec.from_kwargs(**cfg), 0))

    def run(x):
        codes, fs = o.encode(x, 6)
        return o.decode(codes, fs)

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
    info = {"cores": cores, "kind": "port", "threads": torch.get_num_threads(),
            "sample": f"{steps} x (encode+decode of {sample} synthetic 3 s clips, S=6), oracle/esc_oracle.py fp32 ATen"}
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


# ------------------------------------------------------------------------------------------------ B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    from escb200.codec import ESC
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
    if N > 1:
        g_codes = torch.empty((N * B, S, 3, W // 2), dtype=torch.int64, device=dev)
        g_audio = torch.empty((N * B, n_out), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # 2x the 126 MB L2

    def step_device():
        codes, fs = model.encode(x_dev, S)
        audio = model.decode(codes, fs)
        if N > 1:
            dist.all_gather_into_tensor(g_codes, codes)
            dist.all_gather_into_tensor(g_audio, audio)
        return codes, audio

    def step_host():
        codes, fs = model.encode(x_host, S)          # CPU tensors: pinned H2D + kernels + D2H inside
        audio = model.decode(codes, fs)
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

    h = model._handle(dev)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed region: K steps, each bracketed by events, L2 flushed (outside the events) between steps
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = h.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step_device()
        b.record()
    barrier()
    launches = h.launch_count() - launches0
    clocks = sampler.finish()
    total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
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
           "api": "ESC.encode(x_cpu, 6); ESC.decode(codes_cpu, feat_shape) -> escb_encode_host / escb_decode_host"}

    if N > 1:
        launches_t = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(launches_t)
        launches = int(launches_t.item())
    if rank != 0:
        if N > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel-class timing (rank 0, profiled pass: CUDA events around every launch on the launch stream)
    peaks = load_peaks()
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
    ridge = peaks["tensor"] * 1e3 / peaks["hbm"]
    if ai >= ridge:
        roof = {"bound": "tensor", "achieved": tflops, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": tflops / peaks["tensor"]}
    else:
        roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"]}
    traffic, traffic_src = None, None
    try:                                     # measured DRAM bytes per launch of that class (tools/ncu_traffic.py over one step)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if args.config == "base" and B == 36 and top in tj["classes"]:
            traffic = tj["classes"][top]["dram_bytes_per_launch"]
            traffic_src = "profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's launches of this class)"
    except (OSError, ValueError, KeyError):
        pass
    roof.update({"traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": tv["bytes"] / max(tv["launches"], 1),
                 "kernel": top, "launches": tv["launches"], "avg_launch_ms": tv["ms"] / max(tv["launches"], 1),
                 "share_of_step": tv["ms"] / tot_ms, "flop_per_byte": ai, "achieved_tflops": tflops, "achieved_gbs": gbs,
                 "peak_source": peaks["src"] + " (MEASURED_PEAKS.json bf16 sustained / hbm copy)",
                 "whole_step_tflops": value / N * GFLOP_PER_CLIP[args.config] / 1e3,
                 # the step as a whole is tensor work (97 % of its flops are GEMM / conv): algorithmic TFLOP/s per GPU against
                 # the TF32 ceiling (nominally half the measured bf16 rate) and against a third of it, the honest ceiling
                 # of fp32-grade 3xTF32 products
                 "whole_step_frac_of_tf32_peak": value / N * GFLOP_PER_CLIP[args.config] / 1e3 / (peaks["tensor"] / 2.0),
                 "whole_step_frac_of_3xtf32_ceiling": value / N * GFLOP_PER_CLIP[args.config] / 1e3 / (peaks["tensor"] / 6.0),
                 "how": "escb_profile_begin/end: CUDA events around every launch on the launching stream, separate pass of the same K steps"})
    breakdown = {k: {"share": round(v["ms"] / tot_ms, 4), "ms_per_step": round(v["ms"] / args.steps, 4),
                     "launches_per_step": v["launches"] // args.steps,
                     "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 3),
                     "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if v["launches"]}

    def rvq_entry(names):
        ms = sum(prof[n]["ms"] for n in names)
        by = sum(prof[n]["bytes"] for n in names)
        return {"achieved": by / max(ms, 1e-9) / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": by / max(ms, 1e-9) / 1e6 / peaks["hbm"], "ms_per_step": ms / args.steps}
    rvq = {"argmin_only": rvq_entry(["codebook_argmin"]),
           "fused_stream_step": rvq_entry(["pvq_down_gemm", "codebook_argmin", "pvq_up_gemm"]),
           "note": "argmin-only is FMA-issue bound by construction (460 flop/B, SURVEY 8d); the stream step is the HBM-bound one"}

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
