#!/usr/bin/env python
"""Benchmark of the laughter-detection hot path on B200 (BASELINE.json metric: audio-hours/sec of
features + ResNetBigger inference (+ segmenter)).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference's CPU path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # one rank per GPU, weak scaling

One step = one pass of the hot path over one meeting-shaped batch per GPU (default 6 channels x 60 min of
synthetic 16 kHz int16 audio): K1 log-mel -> K2/K3 ResNetBigger on the window starting at every frame -> K4 run
extraction for the reference's 29-threshold grid (x 3 min lengths on the host).  `value` times K steps with the PCM
resident in HBM (CUDA events, max over ranks); `e2e` times the public pipeline call with pinned HOST PCM (H2D, all
kernels, D2H of the run lists, float64 min-length filter).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-hours/sec (features+ResNetBigger inference+segmenter)"
DENSE_FLOP_PER_WINDOW = 2 * 708330784  # one full ResNetBigger forward per 10 ms frame


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons in the background during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(sd, pcm_i16, thresholds, min_lengths):
    """The reference's CPU path on one bounded sample: Lhotse-style Fbank (restated), ResNetBigger on every frame's
    window with batch 32 (torch CPU, all host threads), get_laughter_instances over the grid.  Oracle port."""
    from oracle import fbank_oracle, resnet_oracle, segmenter_oracle
    x = pcm_i16.astype(np.float32) / 32768.0
    feats = fbank_oracle.fbank(x).numpy()
    probs = resnet_oracle.window_probs(sd, feats, batch_size=32)
    fps = len(probs) / (len(pcm_i16) / 16000.0)
    inst = segmenter_oracle.get_laughter_instances(probs, thresholds, min_lengths, fps)
    return feats, probs, inst


def time_cpu_reference(sample_seconds, steps, warmup):
    from laughter_detection_icsi_b200 import synth
    sd = synth.synthetic_state_dict()
    thresholds, min_lengths = synth.eval_grid()
    pcm = synth.synth_channel(int(sample_seconds * 16000)).numpy()
    for _ in range(warmup):
        cpu_reference_step(sd, pcm[: 16000 * 2], thresholds, min_lengths)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, pcm, thresholds, min_lengths)
    dt = (time.perf_counter() - t0) / steps
    return (sample_seconds / 3600.0) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core
    sample_s = args.cpu_sample_seconds
    value, dt = time_cpu_reference(sample_s, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{sample_s:g} s of one synthetic channel per step ({int(sample_s * 100)} windows), batch 32, fp32, no_grad, 87 settings"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CPU path (oracle port: restated Lhotse Fbank + torch-CPU ResNetBigger on every frame window + "
                               "get_laughter_instances) on a bounded sample of the bench workload", "sample": sample,
                   "host_cores": os.cpu_count(), "torch_threads": cores},
        "cpu_baseline": {"value": value, "unit": "audio-hours/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- training (secondary)
def time_training(local_rank, world, batch, steps, warmup):
    """BASELINE configs 4/5: ResNetBigger forward-backward + clip + Adam on synthetic LAD windows, `batch` per GPU, bf16
    operands; data parallel = one flat-bucket NCCL all-reduce of the gradients per step.  Every step copies its batch from
    pinned host memory.  Returns (ms per step on this rank, last loss)."""
    from laughter_detection_icsi_b200 import models, synth, train as ld_train
    dev = torch.device("cuda", local_rank)
    model = models.ResNetBigger(dropout_rate=0.5, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
    model.load_state_dict(synth.synthetic_state_dict(head_gain=1.0, head_bias_shift=0.0))
    model.set_device(dev)
    opt = ld_train.B200Adam(model)   # K8: clip_grad_norm_(1.0) + Adam fused on the flat parameter vector
    rank = int(os.environ.get("RANK", "0"))
    batches = []
    for s in range(4):
        b = ld_train.synthetic_lad_batch(batch, seed=1000 * rank + s)
        batches.append({k: v.pin_memory() for k, v in b.items()})
    loss = 0.0
    for i in range(warmup):
        loss = ld_train.train_batch_fused(model, opt, batches[i % 4], dev, world_size=world)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = ld_train.train_batch_fused(model, opt, batches[i % 4], dev, world_size=world)[0]
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, loss


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch.distributed as dist
    from laughter_detection_icsi_b200 import _native, synth
    from laughter_detection_icsi_b200.pipeline import LaughterPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    n_gpus = world

    thresholds, min_lengths = synth.eval_grid()
    pipe = LaughterPipeline(synth.synthetic_state_dict(), device=local_rank, thresholds=thresholds, min_lengths=min_lengths)
    eng = pipe.engine
    n_samples = int(args.minutes * 60 * 16000)
    pcm_dev, chan_len = synth.synth_meeting(args.channels, n_samples, meeting=rank, device=f"cuda:{local_rank}")
    pcm_host = torch.empty(pcm_dev.shape, dtype=torch.int16, pin_memory=True)
    pcm_host.copy_(pcm_dev)
    torch.cuda.synchronize()
    hours_per_step = args.channels * args.minutes / 60.0

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        runs, frames = pipe.step_device(pcm_dev, chan_len)
    barrier()
    eng.timing_read(reset=True)
    eng.timing_enable(True)
    launches0 = eng.kernel_launches
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        runs, frames = pipe.step_device(pcm_dev, chan_len)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    eng.timing_enable(False)
    conv_ms = eng.timing_read_convs(reset=True)
    timing = eng.timing_read(reset=True)
    launches = eng.kernel_launches - launches0
    windows_per_step = sum(frames)

    # ---- end to end through the public pipeline call with host buffers ------------------------------------------
    pipe(pcm_host, chan_len)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        inst, _ = pipe(pcm_host, chan_len)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    d2h = pipe.d2h_bytes()
    n_segments = sum(len(v) for d in inst for v in d.values())

    train_ms, train_loss = (0.0, 0.0)
    if args.train_steps > 0:
        barrier()
        train_ms, train_loss = time_training(local_rank, world, args.train_batch, args.train_steps, 3)

    t = torch.tensor([ms, e2e_s * 1e3, train_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, train_ms_max = t.tolist()

    if rank == 0:
        peaks = load_peaks()
        gemm_ms, gemm_launches = timing["conv_gemm"]
        fbank_ms, fbank_launches = timing["fbank"]
        dense_tf = DENSE_FLOP_PER_WINDOW * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        exec_tf = 2 * eng.gemm_macs_per_row * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        hbm_gbs = eng.gemm_plane_bytes_per_row * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e9 if gemm_ms else None
        prof = {}
        prof_path = os.path.join(ROOT, "profiles", "latest.json")
        if os.path.exists(prof_path):
            with open(prof_path) as f:
                prof = json.load(f)
        line = {
            "metric": METRIC, "value": n_gpus * hours_per_step * args.steps / (ms_max * 1e-3), "unit": "audio-hours/sec",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "dtype_note": "f16 operands / f32 accumulation in the conv stack; f32 front-end, stem and head",
            "data": "synthetic",
            "config": {
                "workload": f"config-3 shaped inference: per GPU per step one synthetic meeting = {args.channels} channels x "
                            f"{args.minutes:g} min of 16 kHz int16 audio; K1 log-mel + ResNetBigger(resnet_base) on the window of EVERY "
                            f"frame + run extraction for {len(thresholds)} thresholds x {len(min_lengths)} min lengths",
                "channels_per_gpu": args.channels, "minutes_per_channel": args.minutes, "windows_per_step_per_gpu": windows_per_step,
                "checkpoint": "random-init (seeded), BatchNorm statistics randomised, head calibrated to logit std 2",
                "l2": f"inputs larger than L2 ({2 * sum(chan_len) / 1e6:.0f} MB PCM and {windows_per_step * 176 / 1e6:.0f} MB features per step)",
                "segments_found_last_step": n_segments,
            },
            "e2e": {"value": n_gpus * hours_per_step * args.steps / (e2e_ms_max * 1e-3), "unit": "audio-hours/sec",
                    "h2d_bytes_per_step": pipe.h2d_bytes(chan_len), "d2h_bytes_per_step": d2h, "timed": "wall clock around the "
                    "public LaughterPipeline call (pinned host PCM in, per-channel segment lists out), max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "gemm_taps_kernel (tcgen05 shifted-plane implicit-GEMM conv, all 19 conv launches per chunk)",
                # After cross-window reuse (11.3x fewer FLOPs than the dense evaluation) the conv stack is closer to the HBM
                # roof than to the tensor roof: `achieved` = activation-plane bytes the launches must move (every input,
                # residual and output plane of a launch once) / their CUDA-event time, against the measured
                # copy bandwidth.  The tensor-pipe view (executed and SURVEY.md section 8(d)'s dense-equivalent FLOPs against the
                # measured bf16 peak) is reported beside it.
                "bound": "hbm", "unit": "GB/s",
                "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "frac": hbm_gbs / peaks["hbm_gbs"] if hbm_gbs else None,
                "algorithmic_bytes_per_frame": eng.gemm_plane_bytes_per_row,
                "achieved_note": "fp16 activation planes read + written by the conv GEMM launches (algorithmic: each plane once per "
                                 "launch that touches it; weights are negligible) / CUDA-event time of those launches",
                "tensor": {"unit": "TFLOP/s", "peak": peaks["tflops"],
                           "executed_tflops": exec_tf, "executed_frac": exec_tf / peaks["tflops"] if exec_tf else None,
                           "algorithmic_tflops": dense_tf, "algorithmic_frac": dense_tf / peaks["tflops"] if dense_tf else None,
                           "note": "executed = 2 x 62.4 MMAC per frame after cross-window reuse; algorithmic = the dense 1.41666 GFLOP "
                                   "forward per frame the reference computes (bit-identical results)"},
                "kernel_ms_per_step": gemm_ms / args.steps, "kernel_launches_per_step": gemm_launches / args.steps,
                "kernel_share_of_step": gemm_ms / ms if ms else None,
                "per_conv_ms_per_step": {name: round(v / args.steps, 3) for name, v in conv_ms},
                "class_ms_per_step": {name: round(v[0] / args.steps, 3) for name, v in timing.items()},
                "peak_source": peaks["source"], "traffic": prof.get("gemm_dram_bytes_per_launch"),
                "traffic_note": "DRAM read+write bytes of the block1.1.conv2 launch (largest conv launch, 32768 window starts) from ncu --set full, profiles/",
                "traffic_algorithmic": _native.plan_plane_bytes_per_row(eng.cfg, conv="block1.1.conv2") * (32768 + 100),
                "traffic_algorithmic_note": "plane bytes of that same launch (32768 window starts + 100 halo rows)",
                "fbank": {"bound": "hbm", "unit": "GB/s", "achieved": 496.0 * windows_per_step * args.steps / (fbank_ms * 1e-3) / 1e9
                          if fbank_ms else None, "peak": peaks["hbm_gbs"], "ms_per_step": fbank_ms / args.steps,
                          "note": "K1 is fp32-ALU bound (exact 512-point FFT), see DESIGN.md"},
            },
        }
        if args.train_steps > 0:
            line["train"] = {
                "metric": "training samples/sec (ResNetBigger forward+backward+clip+Adam, bf16 operands)", "value": n_gpus * args.train_batch /
                (train_ms_max * 1e-3), "unit": "samples/sec", "ms_per_step": train_ms_max, "batch_per_gpu": args.train_batch,
                "steps": args.train_steps, "last_loss": train_loss, "scaling": "weak",
                "includes": "H2D of the batch from pinned host memory, forward, backward, flat-bucket gradient all-reduce (N>1), "
                            "fused clip_grad_norm_ 1.0 + Adam (ld_clip_adam_step), per-step loss/accuracy read-back like train.py:297", "data": "synthetic LAD windows (100 x 44 log-mel-like), labels recoverable"}
        if n_gpus == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            v, dt = time_cpu_reference(args.cpu_sample_seconds, 1, 1)
            line["cpu_baseline"] = {
                "value": v, "unit": "audio-hours/sec", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{args.cpu_sample_seconds:g} s of one synthetic channel ({int(args.cpu_sample_seconds * 100)} windows; "
                          "restated Lhotse Fbank + torch-CPU ResNetBigger batch 32 fp32 no_grad + 87-setting segmenter), "
                          f"{dt:.1f} s of CPU time, linear in frames"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=6, help="channels per GPU per step (one meeting)")
    ap.add_argument("--minutes", type=float, default=60.0, help="minutes of audio per channel")
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0, help="audio seconds per CPU-reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps for the secondary `train` object (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=256, help="training windows per GPU per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
