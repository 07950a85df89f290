#!/usr/bin/env python
"""Benchmark of the hot path: audio -> HCQT -> CNN -> note events, audio-seconds per second.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (CPU baseline arm: the oracle port on host cores)

Workload (BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is quoted on): 10 s synthetic
22 050 Hz clips, full pipeline to note events, sharded by file with no data-path collective.  Weak scaling:
every rank transcribes `--clips` (default 1250 = 10 000 / 8) clips per step.  One step = one pass of the
path over that batch.  `value` is measured with the audio already resident in HBM (bp_transcribe_device);
`e2e` goes through the host entry point (bp_transcribe_host) from pinned host buffers, H2D of the audio and
D2H of the note events inside the timed region.  Inputs per step (1.1 GB) exceed the 126 MB L2, so no
explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_SECONDS = 10.0
SR = 22050
WORKLOAD = "BASELINE configs[3] per-GPU shard: 10 s synthetic 22 050 Hz clips, full pipeline (HCQT+CNN+note decode) to note events, sharded by file"
FLOP_PER_WINDOW = 1_048_159_296
CONTOUR1_FLOP_PER_WINDOW = 680_030_208  # the dominant kernel (3x39 conv, 8->8 channels) ...
CONTOUR2_FLOP_PER_WINDOW = 18_163_200  # ... whose epilogue also does the MACs of the 5x5 8->1 conv (2 x 200 x 172 x 264)


def make_clips(n_clips: int, seed0: int):
    from basic_pitch_b200 import synth

    distinct = min(n_clips, 125)
    base = [synth.random_notes_clip(CLIP_SECONDS, seed=seed0 + i) for i in range(distinct)]
    return [base[i % distinct] for i in range(n_clips)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """Cores this process may really use (affinity mask and cgroup CPU quota), capped at 32: torch's CPU kernels
    collapse when oversubscribed on a shared host."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()[:2]
            if quota != "max":
                n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 32))


def cpu_baseline_bounded(budget_s: float, threads: int, seed0: int = 1000):
    """Time one clip, then as many as fit in ~budget_s seconds."""
    v1, _ = cpu_baseline(1, threads, seed0)
    per_clip = CLIP_SECONDS / v1
    n = int(max(2, min(48, budget_s / max(per_clip, 1e-3))))
    return cpu_baseline(n, threads, seed0 + 1)


def cpu_baseline(n_clips: int, threads: int, seed0: int = 1000):
    """The oracle port (torch-CPU fp32 restatement of the deployed graph + NumPy restatement of the reference
    decode) on the host cores.  Returns (audio-s/s, description)."""
    import torch

    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, weights
    from oracle import decode_ref, host_ref, model_ref

    torch.set_num_threads(threads)
    w = weights.load(ICASSP_2022_MODEL_PATH)
    clips = make_clips(n_clips, seed0)
    t0 = time.perf_counter()
    n_notes = 0
    for c in clips:
        out = model_ref.forward_batched(host_ref.window_audio(c), w)
        post = {k: host_ref.unwrap(out[k], len(c)) for k in out}
        with np.errstate(all="ignore"):
            _wb, ev = decode_ref.model_output_to_note_events(post, 0.5, 0.3)
        n_notes += len(ev)
    dt = time.perf_counter() - t0
    return n_clips * CLIP_SECONDS / dt, f"{n_clips} x {CLIP_SECONDS:.0f} s clips, model (torch-CPU fp32, {threads} threads) + decode (NumPy restatement, 1 thread), {n_notes} notes, {dt:.1f} s"


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = host_threads()
    v1, _ = cpu_baseline(2, threads)  # warm-up, also sizes the step
    per_step = int(max(2, min(24, 8.0 / (CLIP_SECONDS / v1))))
    vals = []
    t0 = time.perf_counter()
    for s in range(args.steps):
        v, desc = cpu_baseline(per_step, threads, seed0=2000 + 10 * s)
        vals.append(v)
    dt = time.perf_counter() - t0
    value = args.steps * per_step * CLIP_SECONDS / dt
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clip_seconds": CLIP_SECONDS, "sample_clips_per_step": per_step,
                   "sample": f"bounded sample of the workload: {per_step} of its 10 s clips per step",
                   "note": "restated CPU baseline (onnxruntime / TensorFlow are not installable offline; the reference's own decode is pure Python like this port)"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=1250, help="10 s clips per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="time budget of the cpu_baseline sample")
    ap.add_argument("--profile-kernel", type=int, default=0, help="kernel family timed for the roofline line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    # NCCL prints its version banner on stdout at VERSION/INFO level; the contract is ONE JSON line on stdout
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"

    import torch
    import torch.distributed as dist

    from basic_pitch_b200 import ICASSP_2022_MODEL_PATH, engine
    from basic_pitch_b200.inference import Model

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    model = Model(ICASSP_2022_MODEL_PATH, device=local)
    engine.broadcast_weights(model)  # the one collective of the path

    clips = make_clips(args.clips, seed0=3 + 100000 * rank)
    audio_s = args.clips * CLIP_SECONDS
    packed = engine.PackedAudio(clips, pinned=True)
    lib = model._lib
    n_windows = sum(int(lib.bp_num_windows(len(c))) for c in clips)
    n_frames = sum(int(lib.bp_num_frames(len(c))) for c in clips)
    out = engine.NoteBuffers(args.clips, max(4096, 2 * n_frames), max(65536, 24 * n_frames))
    d_audio = packed.to_device(local)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    dev_step = lambda: engine.transcribe_packed_device(model, d_audio, packed.offsets, out)  # noqa: E731
    host_step = lambda: engine.transcribe_packed_host(model, packed, out)  # noqa: E731

    for _ in range(args.warmup):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.bp_model_profile(model.handle, args.profile_kernel)
    l0 = model.launch_count
    ms = timed(dev_step, args.steps)
    launches = model.launch_count - l0
    import ctypes as C

    tot, nint, nwin = C.c_double(), C.c_int64(), C.c_int64()
    lib.bp_model_profile_read(model.handle, C.byref(tot), C.byref(nint), C.byref(nwin))
    lib.bp_model_profile(model.handle, -1)
    n_notes = out.n_notes()
    for _ in range(2):
        host_step()
    ms_e2e = timed(host_step, args.steps)
    d2h = out.result_bytes()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        k_ms = tot.value / max(nint.value, 1)
        k_windows = nwin.value / max(nint.value, 1)
        fam_flop = {0: CONTOUR1_FLOP_PER_WINDOW + CONTOUR2_FLOP_PER_WINDOW, 1: 193_740_800, 2: 57_065_472, 3: 22_359_552}.get(args.profile_kernel, 0)
        achieved = fam_flop * k_windows / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        threads = host_threads()
        cpu_v, cpu_desc = cpu_baseline_bounded(args.cpu_seconds, threads)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                traffic = json.load(fh)["dram_bytes_per_window"] * k_windows if args.profile_kernel == 0 else None
        except (OSError, KeyError, ValueError):
            pass
        value = world * args.steps * audio_s / (ms * 1e-3)
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "clips_per_gpu_per_step": args.clips, "clip_seconds": CLIP_SECONDS, "windows_per_gpu_per_step": n_windows,
                "frames_per_gpu_per_step": n_frames, "notes_per_gpu_per_step": n_notes, "parallelism": f"files x{world}, no data-path collective",
                "l2": f"inputs {packed.nbytes / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                "windows_per_second_per_gpu": n_windows * args.steps / (ms * 1e-3),
                "model_flops_fraction_of_peak": n_windows * args.steps * FLOP_PER_WINDOW / (ms * 1e-3) / 1e12 / peak_tf,
            },
            "e2e": {"value": world * args.steps * audio_s / (ms_e2e * 1e-3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": packed.nbytes, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "bp_transcribe_host (pinned host audio in, note events out)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": {0: "contour conv 8->8 3x39 + fused conv2 8->1 5x5 (conv_tc_kernel<3>, tcgen05 split-bf16)", 1: "onset conv 8->32 5x5/3 (conv_tc_kernel<1>)", 2: "cqt + lognorm", 3: "decimation chain"}.get(args.profile_kernel),
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu, profiles/roofline_traffic.json)",
                         "peak_source": peak_src, "avg_launch_ms": k_ms, "windows_per_launch": k_windows,
                         "flop_per_window": fam_flop},
            "cpu_baseline": {"value": cpu_v, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": cpu_desc},
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
