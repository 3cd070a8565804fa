#!/usr/bin/env python
"""Benchmark of the laughter-detection hot path on B200 (BASELINE.json metric: audio-hours/sec of
features + ResNetBigger inference (+ segmenter)).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (default workload)
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference's CPU path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # one rank per GPU, weak scaling
    python bench.py --config features                               # BASELINE config 2: K1 alone, 64 channel-hours per launch
    python bench.py --config corpus [--gpus N]                      # BASELINE config 3: 75 meetings x 6 channels x 1 h, strong scaling
    python bench.py --config cli                                    # BASELINE config 1: one 10-minute channel through the CLI path

Default: one step = one pass of the hot path over one meeting-shaped batch per GPU (6 channels x 60 min of synthetic
16 kHz int16 audio): K1 log-mel -> K2/K3 ResNetBigger on the window starting at every frame -> K4 run extraction for
the reference's 29-threshold grid (x 3 min lengths on the host).  `value` times K steps with the PCM resident in HBM
(CUDA events, max over ranks); `e2e` times the public pipeline call with pinned HOST PCM (H2D, all kernels, D2H of the
run lists, float64 min-length filter).  Prints ONE JSON line on rank 0.
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
FBANK_BYTES_PER_FRAME = 496.0           # 320 B of int16 PCM in + 176 B of fp32 log-mel out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))), "hbm_gbs": float(p["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS / copy bandwidth)"}
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
def cpu_reference_step(sd, pcm_i16, thresholds, min_lengths, no_grad=True):
    """The reference's CPU path on one bounded sample: Lhotse-style Fbank (restated), ResNetBigger on every frame's
    window with batch 32 (torch CPU, all host threads), get_laughter_instances over the grid.  Oracle port (pinned to the
    reference's own models.py / laugh_segmenter.py outputs at 2e-7, tests/test_oracle_golden.py).  no_grad=False
    reproduces the reference literally: segment_laughter.py:95 runs the model WITHOUT torch.no_grad()."""
    from oracle import fbank_oracle, resnet_oracle, segmenter_oracle
    x = pcm_i16.astype(np.float32) / 32768.0
    feats = fbank_oracle.fbank(x).numpy()
    if no_grad:
        probs = resnet_oracle.window_probs(sd, feats, batch_size=32)
    else:
        probs = resnet_oracle.window_probs_autograd(sd, feats, batch_size=32)
    fps = len(probs) / (len(pcm_i16) / 16000.0)
    inst = segmenter_oracle.get_laughter_instances(probs, thresholds, min_lengths, fps)
    return feats, probs, inst


def time_cpu_reference(sample_seconds, steps, warmup, no_grad=True):
    from laughter_detection_icsi_b200 import synth
    sd = synth.synthetic_state_dict()
    thresholds, min_lengths = synth.eval_grid()
    pcm = synth.synth_channel(int(sample_seconds * 16000)).numpy()
    for _ in range(warmup):
        cpu_reference_step(sd, pcm[: 16000 * 2], thresholds, min_lengths, no_grad)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, pcm, thresholds, min_lengths, no_grad)
    dt = (time.perf_counter() - t0) / steps
    return (sample_seconds / 3600.0) / dt, dt


def cpu_baseline_object(sample_seconds, steps=1, warmup=1):
    """`cpu_baseline` of the bench line: the oracle port on the box's host cores, with torch.no_grad() (value) and the way
    the reference literally runs it, building an autograd graph per batch (value_autograd_graph)."""
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core
    v, dt = time_cpu_reference(sample_seconds, steps, warmup, no_grad=True)
    v_g, dt_g = time_cpu_reference(sample_seconds, 1, 0, no_grad=False)
    sample = (f"{sample_seconds:g} s of one synthetic channel per step ({int(sample_seconds * 100)} windows; restated Lhotse Fbank + "
              "torch-CPU ResNetBigger batch 32 fp32 + 87-setting segmenter), linear in frames")
    return {"value": v, "unit": "audio-hours/sec", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
            "cpu_seconds_per_step": dt, "value_autograd_graph": v_g,
            "note": "value: with torch.no_grad(); value_autograd_graph: as segment_laughter.py:95 runs it (no no_grad). The port "
                    "is pinned to the reference's own models.py/laugh_segmenter.py outputs (tests/golden)"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box
    sample_s = args.cpu_sample_seconds
    base, dt = cpu_baseline_object(sample_s, args.steps, args.warmup)
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference CPU path (oracle port: restated Lhotse Fbank + torch-CPU ResNetBigger on every frame window + "
                               "get_laughter_instances) on a bounded sample of the bench workload", "sample": base["sample"],
                   "host_cores": os.cpu_count(), "torch_threads": base["cores"]},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- training (secondary)
def time_training(local_rank, world, batch, steps, warmup):
    """BASELINE configs 4/5: ResNetBigger forward-backward + clip + Adam on synthetic LAD windows, `batch` per GPU, bf16
    operands; data parallel = one flat-bucket NCCL all-reduce of the gradients per step.  Every step copies its batch from
    pinned host memory.  Returns (ms per step on this rank, last loss, kernel launches per step)."""
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
    stepper = ld_train.make_stepper(model, opt, dev, world_size=world, graph=os.environ.get("LD_TRAIN_GRAPH", "1") != "0")
    for i in range(max(warmup, 3)):
        stepper(batches[i % 4])
    stepper.flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m = None
    for i in range(steps):
        m = stepper(batches[i % 4]) or m
    m = stepper.flush() or m   # (inside the clock: every step's loss / accuracy reaches the host)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(m[0]), stepper.launches_per_step, stepper.mode


# ----------------------------------------------------------------------------------------------------- parity (outside the timed region)
def parity_report(pipe, sd, pcm_dev_10min, thresholds, min_lengths, n_blocks=8, block=300):
    """Agreement of the BENCHMARKED configuration (default chunk_rows, calibrated checkpoint) with the fp64 oracle on one
    10-minute channel of the bench workload: probabilities on n_blocks x block sampled windows (contiguous blocks incl. the
    one straddling sequence row 32768 and the zero-padded tail) and, per block, the segment lists of all 87 settings from
    the GPU's K4 + float64 filter against get_laughter_instances on the ORACLE's probabilities (reference
    datasets.py:85-93, laugh_segmenter.py:87-108)."""
    from oracle import fbank_oracle, resnet_oracle, segmenter_oracle
    eng = pipe.engine
    n = pcm_dev_10min.numel()
    feats, frames = eng.fbank(pcm_dev_10min, [n])
    probs = eng.infer_windows(feats, frames)
    T = frames[0]
    ref_feats = fbank_oracle.fbank(pcm_dev_10min.cpu().numpy().astype(np.float32) / 32768.0).numpy()   # the reference's fp32 front-end
    feat_err = float(np.max(np.abs(feats.cpu().numpy() - ref_feats) / np.maximum(1.0, np.abs(ref_feats))))
    starts = sorted(set([int(x) for x in np.linspace(0, T - block, n_blocks - 2)] + [max(0, min(T - block, 32768 - block // 2)),
                                                                                      T - block]))
    fps = T / (n / 16000.0)
    p_gpu = probs.cpu().numpy()
    max_dp, sq, cnt = 0.0, 0.0, 0
    settings_equal = {(t, m): True for t in thresholds for m in min_lengths}
    shift_hist = {"0": 0, "1": 0, "2": 0, ">2": 0}
    seg_total, seg_unmatched, flips, flips_outside_band = 0, 0, 0, 0
    for a in starts:
        ref = resnet_oracle.window_probs(sd, ref_feats, dtype=torch.float64, start=a, stop=a + block)
        got = p_gpu[a:a + block]
        d = np.abs(got.astype(np.float64) - ref)
        max_dp = max(max_dp, float(d.max())); sq += float((d * d).sum()); cnt += block
        inst_ref = segmenter_oracle.get_laughter_instances(ref, thresholds, min_lengths, fps)
        blk = probs[a:a + block].contiguous()
        runs = pipe.runs(blk, [block])
        inst_gpu = pipe.instances(runs, [block], [block / fps])[0]
        for th in thresholds:
            f = (got > np.float32(th)) != (ref > th)
            flips += int(f.sum())
            flips_outside_band += int(np.sum(np.abs(ref[f] - th) > d.max() + 1e-7))
        for key, exp in inst_ref.items():
            g = [tuple(r) for r in inst_gpu[key].tolist()]
            seg_total += len(exp)
            if g != exp:
                settings_equal[key] = False
                if len(g) == len(exp):
                    for (gs, ge), (es, ee) in zip(g, exp):
                        for dv in (abs(gs - es), abs(ge - ee)):
                            k = int(round(dv * fps))
                            shift_hist[str(k) if k <= 2 else ">2"] += 1
                else:
                    seg_unmatched += abs(len(g) - len(exp))
            else:
                shift_hist["0"] += 2 * len(exp)
    n_set = len(settings_equal)
    return {
        "on": f"first 10 minutes of channel 0 of the bench meeting as one channel ({T} windows), default chunk_rows, bench checkpoint",
        "windows_compared": cnt, "blocks": [[a, a + block] for a in starts],
        "feat_max_rel_err_log_domain": feat_err, "feat_tolerance": 1e-4,
        "prob_max_abs_err_vs_fp64_oracle": max_dp, "prob_rms_err": (sq / max(cnt, 1)) ** 0.5,
        "settings": n_set, "settings_with_identical_segment_lists": sum(settings_equal.values()),
        "settings_identical_frac": sum(settings_equal.values()) / n_set,
        "oracle_segments": seg_total, "segment_count_mismatch": seg_unmatched,
        "boundary_shift_frames_hist": shift_hist,
        "threshold_flips": flips, "threshold_flips_outside_tie_band": flips_outside_band,
        "note": "segments are bit-exact given the same probabilities (tests/test_gpu_parity.py); differences here are threshold "
                "ties: frames where |p_oracle - thr| is within the probability error",
    }


# ----------------------------------------------------------------------------------------------------- GPU arm (default workload)
def dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    return world, rank, local_rank, barrier


def max_over_ranks(values, local_rank, world):
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def run_b200(args):
    import torch.distributed as dist
    from laughter_detection_icsi_b200 import _native, synth
    from laughter_detection_icsi_b200.pipeline import LaughterPipeline

    world, rank, local_rank, barrier = dist_setup()
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    n_gpus = world

    thresholds, min_lengths = synth.eval_grid()
    sd = synth.synthetic_state_dict()
    pipe = LaughterPipeline(sd, device=local_rank, thresholds=thresholds, min_lengths=min_lengths, precision=args.precision,
                            chunk_rows=args.chunk_rows)
    eng = pipe.engine
    n_samples = int(args.minutes * 60 * 16000)
    pcm_dev, chan_len = synth.synth_meeting(args.channels, n_samples, meeting=rank, device=f"cuda:{local_rank}")
    pcm_host = torch.empty(pcm_dev.shape, dtype=torch.int16, pin_memory=True)
    pcm_host.copy_(pcm_dev)
    torch.cuda.synchronize()
    hours_per_step = args.channels * args.minutes / 60.0

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        runs, frames = pipe.step_device(pcm_dev, chan_len)
    barrier()
    eng.timing_enable(False)   # no per-launch events inside the headline region (they would sit between the conv launches)
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
    launches = eng.kernel_launches - launches0
    windows_per_step = sum(frames)
    # ---- the same K steps once more with a CUDA event pair around every launch: the per-class / per-conv breakdown behind
    #      `roofline` (its kernel times are measured live, on the launch stream, but outside the headline clock)
    eng.timing_read(reset=True)
    eng.timing_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        pipe.step_device(pcm_dev, chan_len)
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    eng.timing_enable(False)
    conv_ms = eng.timing_read_convs(reset=True)
    timing = eng.timing_read(reset=True)

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

    train_ms, train_loss, train_launches, train_mode = (0.0, 0.0, 0, "")
    if args.train_steps > 0:
        barrier()
        train_ms, train_loss, train_launches, train_mode = time_training(local_rank, world, args.train_batch, args.train_steps, 3)

    ms_max, e2e_ms_max, train_ms_max = max_over_ranks([ms, e2e_s * 1e3, train_ms], local_rank, world)

    if rank == 0:
        peaks = load_peaks()
        gemm_ms, gemm_launches = timing["conv_gemm"]
        fbank_ms, fbank_launches = timing["fbank"]
        bytes_per_row = eng.gemm_plane_bytes_per_row
        dense_tf = DENSE_FLOP_PER_WINDOW * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        exec_tf = 2 * eng.gemm_macs_per_row * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None
        hbm_gbs = bytes_per_row * windows_per_step * args.steps / (gemm_ms * 1e-3) / 1e9 if gemm_ms else None
        # third floor: shared-memory traffic of the SS-mode MMAs (operand reads) and of the ring (bulk-copy writes), 128 B/clk/SM at
        # the SM clock sampled during the timed region
        smem_r, smem_w = _native.plan_gemm_smem_bytes_per_row(eng.cfg)
        sm_hz = (clocks or {}).get("sm_mhz") and clocks["sm_mhz"] * 1e6
        smem_bclk = ((smem_r + smem_w) * windows_per_step * args.steps / (gemm_ms * 1e-3) / torch.cuda.get_device_properties(local_rank).multi_processor_count / sm_hz) if (gemm_ms and sm_hz) else None
        prof = {}
        prof_path = os.path.join(ROOT, "profiles", "latest.json")
        if os.path.exists(prof_path):
            with open(prof_path) as f:
                prof = json.load(f)
        line = {
            "metric": METRIC, "value": n_gpus * hours_per_step * args.steps / (ms_max * 1e-3), "unit": "audio-hours/sec",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "dtype_note": "f16 operands / f32 accumulation in the conv stack; f32 front-end, stem and head"
                          + ("; blocks 2-4 carry weights and activations as hi+lo f16 pairs (precision=split)" if args.precision == "split" else ""),
            "data": "synthetic",
            "config": {
                "workload": f"config-3 shaped inference: per GPU per step one synthetic meeting = {args.channels} channels x "
                            f"{args.minutes:g} min of 16 kHz int16 audio; K1 log-mel + ResNetBigger(resnet_base) on the window of EVERY "
                            f"frame + run extraction for {len(thresholds)} thresholds x {len(min_lengths)} min lengths",
                "channels_per_gpu": args.channels, "minutes_per_channel": args.minutes, "windows_per_step_per_gpu": windows_per_step,
                "checkpoint": "random-init (seeded), BatchNorm statistics randomised, head calibrated to logit std 2",
                "precision": args.precision, "chunk_rows": int(eng.cfg.chunk_rows) or 32768,
                "l2": f"inputs larger than L2 ({2 * sum(chan_len) / 1e6:.0f} MB PCM and {windows_per_step * 176 / 1e6:.0f} MB features per step)",
                "segments_found_last_step": n_segments,
                "timing_note": "value: K steps without per-launch events; roofline.* kernel times: the same K steps repeated with a CUDA "
                               f"event pair around every launch ({ms_profiled / args.steps:.1f} ms per step in that pass)",
            },
            "e2e": {"value": n_gpus * hours_per_step * args.steps / (e2e_ms_max * 1e-3), "unit": "audio-hours/sec",
                    "h2d_bytes_per_step": pipe.h2d_bytes(chan_len), "d2h_bytes_per_step": d2h, "timed": "wall clock around the "
                    "public LaughterPipeline call (pinned host PCM in, per-channel segment lists out), max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "gemm_taps_kernel (tcgen05 shifted-plane implicit-GEMM conv, all conv launches of a pass)",
                # After cross-window reuse (11.3x fewer FLOPs than the dense evaluation) the conv stack sits just left of the
                # ridge: `achieved` = activation-plane bytes the launches must move (every input, residual and output plane
                # of a launch once) / their CUDA-event time, against the measured copy bandwidth.  The
                # tensor-pipe view (executed and SURVEY.md section 8(d)'s dense-equivalent FLOPs against the measured bf16
                # peak) is reported beside it.
                "bound": "hbm", "unit": "GB/s",
                "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "frac": hbm_gbs / peaks["hbm_gbs"] if hbm_gbs else None,
                "algorithmic_bytes_per_frame": bytes_per_row,
                "flop_per_byte": 2 * eng.gemm_macs_per_row / bytes_per_row,
                "ridge_flop_per_byte": peaks["tflops"] * 1e3 / peaks["hbm_gbs"],
                "achieved_note": "fp16 activation planes read + written by the conv GEMM launches of THIS layer-by-layer design (each "
                                 "plane once per launch that touches it; weights are negligible; not a lower bound: fusing conv1->conv2 "
                                 "of a block would remove about a third) / CUDA-event time of those launches",
                "tensor": {"unit": "TFLOP/s", "peak": peaks["tflops"],
                           "executed_tflops": exec_tf, "executed_frac": exec_tf / peaks["tflops"] if exec_tf else None,
                           "algorithmic_tflops": dense_tf, "algorithmic_frac": dense_tf / peaks["tflops"] if dense_tf else None,
                           "note": f"executed = 2 x {eng.gemm_macs_per_row / 1e6:.1f} MMAC per frame after cross-window reuse (identity "
                                   "residual taps not counted); algorithmic = the dense 1.41666 GFLOP forward per frame the reference "
                                   "computes (bit-identical results)"},
                "smem": {"unit": "B/clk/SM", "peak": 128.0, "achieved": smem_bclk, "frac": smem_bclk / 128.0 if smem_bclk else None,
                         "operand_read_bytes_per_frame": smem_r, "bulk_copy_write_bytes_per_frame": smem_w,
                         "note": "shared-memory traffic of the conv launches: A (4 KB) + B (N x 32 B) operand slabs every SS-mode MMA reads, from "
                                 "the tap programs, plus the bulk-copy writes of the ring, against 128 B/clk/SM at the SM clock sampled under load. "
                                 "With cout <= 64 this is the HIGHEST of the three floors of the stack (DESIGN.md section 5): the one that binds"},
                "kernel_ms_per_step": gemm_ms / args.steps, "kernel_launches_per_step": gemm_launches / args.steps,
                "kernel_share_of_step": gemm_ms / ms_profiled if ms_profiled else None,
                "per_conv_ms_per_step": {name: round(v / args.steps, 3) for name, v in conv_ms},
                "class_ms_per_step": {name: round(v[0] / args.steps, 3) for name, v in timing.items()},
                "peak_source": peaks["source"], "traffic": prof.get("gemm_dram_bytes_per_launch"),
                "traffic_note": prof.get("gemm_dram_note", "DRAM read+write bytes of the block1.1.conv2 launch (largest conv launch, 32768 window "
                                                           "starts) from ncu --set full, profiles/"),
                "traffic_algorithmic": _native.plan_plane_bytes_per_row(eng.cfg, conv="block1.1.conv2") * (32768 + 100),
                "traffic_algorithmic_note": "plane bytes of that same launch (32768 window starts + 100 halo rows)",
                "fbank": {"bound": "hbm", "unit": "GB/s", "achieved": FBANK_BYTES_PER_FRAME * windows_per_step * args.steps / (fbank_ms * 1e-3) / 1e9
                          if fbank_ms else None, "peak": peaks["hbm_gbs"], "ms_per_step": fbank_ms / args.steps,
                          "note": "K1 inside the step, one launch per channel-hour; `bench.py --config features` measures it on 64 "
                                  "channel-hours per launch (BASELINE config 2)"},
            },
        }
        if args.train_steps > 0:
            train_flop = 3 * DENSE_FLOP_PER_WINDOW * args.train_batch   # forward + data gradient + weight gradient
            line["train"] = {
                "metric": "training samples/sec (ResNetBigger forward+backward+clip+Adam, bf16 operands)", "value": n_gpus * args.train_batch /
                (train_ms_max * 1e-3), "unit": "samples/sec", "ms_per_step": train_ms_max, "batch_per_gpu": args.train_batch,
                "steps": args.train_steps, "last_loss": train_loss, "scaling": "weak", "kernel_launches_per_step": train_launches,
                "launch_mode": train_mode,
                "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peaks["tflops"],
                             "achieved": train_flop / (train_ms_max * 1e-3) / 1e12,
                             "frac": train_flop / (train_ms_max * 1e-3) / 1e12 / peaks["tflops"],
                             "note": "3 x 1.41666 GFLOP per sample (forward, data gradient, weight gradient) over the WHOLE step time, "
                                     "BatchNorm/element-wise passes, head, optimiser and H2D included"},
                "includes": "H2D of the batch from pinned host memory, forward, backward, flat-bucket gradient all-reduce (N>1), "
                            "fused clip_grad_norm_ 1.0 + Adam (ld_clip_adam_step_dev), per-step loss/accuracy/precision/recall read-back "
                            "like train.py:297 (fetched one step late so that the host never drains the stream)",
                "data": "synthetic LAD windows (100 x 44 log-mel-like), labels recoverable"}
        if n_gpus == 1 and not args.no_parity:
            line["parity"] = parity_report(pipe, sd, pcm_dev[: min(n_samples, 9600000)].contiguous(), thresholds, min_lengths)
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline_object(args.cpu_sample_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------- config 2: features only
def run_features(args):
    """BASELINE config 2: batched log-mel/Fbank extraction, 64 channel-hours of synthetic 16 kHz audio per ld_fbank_i16 call
    on one B200 (and the 1-hour case), against the restated Lhotse Fbank and torchaudio's Kaldi fbank on the host cores."""
    from laughter_detection_icsi_b200 import synth
    from laughter_detection_icsi_b200.engine import get_engine
    world, rank, local_rank, barrier = dist_setup()
    if rank != 0:
        return
    eng = get_engine(local_rank, chunk_rows=256)   # K1 needs no conv workspace
    peaks = load_peaks()
    n_hour = 3600 * 16000
    one = synth.synth_channel(n_hour, device=f"cuda:{local_rank}")
    results = {}
    sampler = None
    for label, n_chan in (("1_channel_hour", 1), (f"{args.feature_channels}_channel_hours", args.feature_channels)):
        # distinct content per channel: channel c is the base hour rotated by c * 7919 samples (no extra generation time)
        pcm = torch.cat([torch.roll(one, shifts=7919 * c) for c in range(n_chan)]) if n_chan > 1 else one
        lens = [n_hour] * n_chan
        for _ in range(max(args.warmup, 3)):
            feats, frames = eng.fbank(pcm, lens)
        torch.cuda.synchronize()
        if n_chan > 1:
            sampler = ClockSampler(local_rank)
        eng.timing_read(reset=True)
        eng.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            feats, frames = eng.fbank(pcm, lens)
        e1.record()
        torch.cuda.synchronize()
        eng.timing_enable(False)
        k_ms, k_launches = eng.timing_read(reset=True)["fbank"]
        ms = e0.elapsed_time(e1) / args.steps
        n_frames = sum(frames)
        results[label] = {"ms_per_call": ms, "kernel_ms_per_call": k_ms / args.steps, "frames": n_frames,
                          "audio_hours_per_sec": n_chan / (ms * 1e-3),
                          "algorithmic_GBps": FBANK_BYTES_PER_FRAME * n_frames / (k_ms / args.steps * 1e-3) / 1e9,
                          "launches_per_call": k_launches / args.steps}
        del feats
    clocks = sampler.stop() if sampler else None
    # end to end: pinned host PCM in, host features out (1 channel-hour)
    host = torch.empty(n_hour, dtype=torch.int16, pin_memory=True); host.copy_(one)
    out_host = torch.empty((int(eng.lib.ld_fbank_num_frames(n_hour)), 44), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d = host.to(f"cuda:{local_rank}", non_blocking=True)
        f, _ = eng.fbank(d, [n_hour])
        out_host.copy_(f, non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    # CPU arm: restated Lhotse Fbank and torchaudio's Kaldi fbank on 1 hour (config 2), 1 thread and all threads
    from oracle import fbank_oracle
    x = one.cpu().numpy().astype(np.float32) / 32768.0
    cpu = {}
    for threads in (1, os.cpu_count() or 1):
        torch.set_num_threads(threads)
        t0 = time.perf_counter(); ref = fbank_oracle.fbank(x); dt = time.perf_counter() - t0
        cpu[f"restated_lhotse_fbank_{threads}_threads_s_per_hour"] = dt
        try:
            import torchaudio
            xt = torch.from_numpy(x)[None]
            t0 = time.perf_counter()
            torchaudio.compliance.kaldi.fbank(xt, num_mel_bins=44, frame_length=25, frame_shift=10, snip_edges=False, dither=0,
                                              energy_floor=0, low_freq=20, high_freq=-400, sample_frequency=16000)
            cpu[f"torchaudio_kaldi_fbank_{threads}_threads_s_per_hour"] = time.perf_counter() - t0
        except Exception as e:  # torchaudio is optional on the box
            cpu["torchaudio"] = f"unavailable: {e}"
    f1, _ = eng.fbank(one, [n_hour])
    err = float(np.max(np.abs(f1.cpu().numpy() - ref.numpy()) / np.maximum(1.0, np.abs(ref.numpy()))))
    big = results[f"{args.feature_channels}_channel_hours"]
    best_cpu = min(v for k, v in cpu.items() if k.startswith("restated") and isinstance(v, float))
    line = {
        "metric": "audio-hours/sec (log-mel Fbank features only)", "value": big["audio_hours_per_sec"], "unit": "audio-hours/sec",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": big["ms_per_call"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE config 2: feature-only, {args.feature_channels} channel-hours of 16 kHz int16 audio per "
                               "ld_fbank_i16 call (and the 1-hour case)", "l2": "inputs larger than L2 (115 MB PCM per channel-hour)"},
        "gpu_launches": int(big["launches_per_call"] * args.steps), "clocks": clocks,
        "e2e": {"value": 1.0 / (e2e_ms * 1e-3), "unit": "audio-hours/sec", "h2d_bytes_per_step": 2 * n_hour,
                "d2h_bytes_per_step": out_host.numel() * 4, "timed": "1 channel-hour: pinned host PCM -> H2D -> K1 -> D2H of the (T, 44) features"},
        "roofline": {"kernel": "fbank_kernel (K1)", "bound": "hbm", "unit": "GB/s", "achieved": big["algorithmic_GBps"],
                     "peak": peaks["hbm_gbs"], "frac": big["algorithmic_GBps"] / peaks["hbm_gbs"], "peak_source": peaks["source"],
                     "algorithmic_bytes_per_frame": FBANK_BYTES_PER_FRAME, "traffic": None,
                     "note": "exact fp32 512-point FFT: about 12 kFLOP per 496 B frame, so the fp32 pipe, not HBM, is the physical "
                             "bound (DESIGN.md section 4); ncu pipe utilisation under profiles/"},
        "results": results,
        "parity": {"max_rel_err_log_domain_vs_fp32_oracle_1h": err, "tolerance": 1e-4},
        "cpu_baseline": {"value": 1.0 / best_cpu, "unit": "audio-hours/sec", "cores": os.cpu_count(), "kind": "port",
                         "sample": "1 hour of synthetic audio, restated Lhotse Wav2LogFilterBank (torch CPU ops)", **cpu},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- config 3: corpus, strong scaling
def run_corpus(args):
    """BASELINE config 3: 75 distinct synthetic meetings x 6 channels x 60 min (450 audio-hours) sharded over the ranks at
    (meeting, channel) granularity -- the reference's unit of scale-out, cluster_scripts/gen_eval_exp.py:99-114 -- with the
    final gather of every channel's segment lists to rank 0 INSIDE the clock.  Strong scaling: the corpus is fixed."""
    import torch.distributed as dist
    from laughter_detection_icsi_b200 import distributed as ldd, synth
    from laughter_detection_icsi_b200.pipeline import LaughterPipeline
    world, rank, local_rank, barrier = dist_setup()
    thresholds, min_lengths = synth.eval_grid()
    sd = synth.synthetic_state_dict()
    pipe = LaughterPipeline(sd, device=local_rank, thresholds=thresholds, min_lengths=min_lengths, precision=args.precision)
    n_samples = int(args.minutes * 60 * 16000)
    units = [(m, c) for m in range(args.corpus_meetings) for c in range(args.channels)]
    durations = [n_samples / 16000.0] * len(units)
    mine = ldd.shard_units(durations, world)[rank]
    # this rank's shard in pinned host memory (generated on the GPU, outside the clock)
    host = torch.empty((len(mine), n_samples), dtype=torch.int16, pin_memory=True)
    for i, u in enumerate(mine):
        host[i].copy_(synth.synth_channel(n_samples, meeting=units[u][0], channel=units[u][1], device=f"cuda:{local_rank}"))
    torch.cuda.synchronize()
    group = max(1, args.channels)
    settings = [(t, m) for t in thresholds for m in min_lengths]
    shards = ldd.shard_units(durations, world)
    stream = ldd.StreamingGather([-(-len(s) // group) for s in shards], settings, len(units), dst=0)   # (gloo side group: before the clock)
    # warm-up: one group
    if len(mine):
        pipe(host[:min(group, len(mine))].reshape(-1), [n_samples] * min(group, len(mine)))
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    for g0 in range(0, len(mine), group):
        k = min(group, len(mine) - g0)
        inst, frames = pipe(host[g0:g0 + k].reshape(-1), [n_samples] * k)
        stream.submit(mine[g0:g0 + k], inst)   # packed and sent to rank 0 on a background thread while the next group runs
    t_compute = time.perf_counter() - t0
    merged = stream.finish()
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    total_s, compute_s = max_over_ranks([t_total, t_compute], local_rank, world)
    if rank == 0:
        clocks = sampler.stop()
        hours = len(units) * args.minutes / 60.0
        n_seg = sum(len(v) for d in merged for v in d.values())
        # parity spot-check: three channels (slices of 3 s) against the oracle's probabilities + segments
        spot = corpus_spot_check(pipe, sd, thresholds, min_lengths, [units[i] for i in (0, len(units) // 2, len(units) - 1)], n_samples)
        line = {
            "metric": METRIC, "value": hours / total_s, "unit": "audio-hours/sec", "n_gpus": world, "steps": 1, "warmup": 1,
            "ms_per_step": total_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"BASELINE config 3: {args.corpus_meetings} distinct synthetic meetings x {args.channels} channels x "
                                   f"{args.minutes:g} min = {hours:g} audio-hours, sharded by (meeting, channel) over {world} rank(s), "
                                   "host PCM in, all segment lists gathered on rank 0 inside the clock (streamed per meeting over a host-side gloo group "
                                   "while the next meeting runs)",
                       "units": len(units), "units_rank0": len(mine), "precision": args.precision,
                       "l2": "inputs larger than L2 (115 MB PCM per channel)"},
            "e2e": {"value": hours / total_s, "unit": "audio-hours/sec", "h2d_bytes_per_step": 2 * n_samples * len(units),
                    "d2h_bytes_per_step": None, "timed": "wall clock from the first H2D copy to the gathered result on rank 0, max over ranks"},
            "seconds_total": total_s, "seconds_before_gather": compute_s, "segments_gathered": n_seg, "clocks": clocks,
            "gpu_launches": pipe.engine.kernel_launches, "parity": spot,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def corpus_spot_check(pipe, sd, thresholds, min_lengths, units, n_samples, seconds=3):
    from laughter_detection_icsi_b200 import synth
    from oracle import fbank_oracle, resnet_oracle, segmenter_oracle
    out = []
    n = seconds * 16000
    for m, c in units:
        pcm = synth.synth_channel(n_samples, meeting=m, channel=c, device=str(pipe.engine.device))[:n].contiguous()
        probs, frames = pipe.probabilities(pcm, [n])
        feats = fbank_oracle.fbank(pcm.cpu().numpy().astype(np.float32) / 32768.0).numpy()
        ref = resnet_oracle.window_probs(sd, feats, dtype=torch.float64)
        got = probs.cpu().numpy()
        err = float(np.abs(got - ref).max())
        fps = frames[0] / float(seconds)
        inst = pipe.instances(pipe.runs(probs, frames), frames, [float(seconds)])[0]
        same = segmenter_oracle.get_laughter_instances(got, thresholds, min_lengths, fps)   # bit-exact given the same probabilities
        ok = all([tuple(r) for r in inst[k].tolist()] == same[k] for k in same)
        out.append({"meeting": m, "channel": c, "windows": int(frames[0]), "prob_max_abs_err": err, "segments_bit_exact_on_gpu_probs": bool(ok)})
    return out


# ----------------------------------------------------------------------------------------------------- config 1: the CLI on one 10-minute channel
def run_cli(args):
    """BASELINE config 1: segment_laughter.py inference on one synthetic 10-minute mono channel (WAV on disk -> TextGrid tree),
    this repo's CLI on the GPU beside the reference CPU path (oracle port) with and without torch.no_grad()."""
    import tempfile
    import scipy.io.wavfile
    from laughter_detection_icsi_b200 import models, segment_laughter, synth
    from laughter_detection_icsi_b200.utils import torch_utils
    world, rank, local_rank, barrier = dist_setup()
    if rank != 0:
        return
    sd = synth.synthetic_state_dict()
    with tempfile.TemporaryDirectory() as tmp:
        m = models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
        m.load_state_dict(sd)
        torch_utils.save_checkpoint(torch_utils.make_state_dict(m, None, 0, 0, np.inf), True, os.path.join(tmp, "ck"))
        pcm = synth.synth_channel(600 * 16000).numpy()
        wav = os.path.join(tmp, "chan0.wav")
        scipy.io.wavfile.write(wav, 16000, pcm)
        argv = ["--config", "resnet_base", "--model_path", os.path.join(tmp, "ck"), "--input_audio_file", wav, "--thresholds",
                ",".join(str(t) for t in synth.eval_grid()[0]), "--min_lengths", "0.0,0.1,0.2", "--save_to_textgrid", "True",
                "--save_to_audio_files", "False", "--output_dir", os.path.join(tmp, "out")]
        import contextlib
        import io
        times = []
        for _ in range(max(args.warmup, 1) + args.steps):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):   # the CLI prints one line per setting; stdout carries ONE JSON line
                segment_laughter.main(argv)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        gpu_s = statistics.median(times[max(args.warmup, 1):])
    base, dt = cpu_baseline_object(args.cpu_sample_seconds)
    line = {"metric": METRIC, "value": (600 / 3600.0) / gpu_s, "unit": "audio-hours/sec", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": gpu_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "BASELINE config 1: the segment_laughter CLI (model load from best.pth.tar, WAV read, features, "
                                   "ResNetBigger on every frame, 87 settings, TextGrid tree) on one synthetic 10-minute channel",
                       "includes": "checkpoint load, WAV read from disk, TextGrid files written"},
            "e2e": {"value": (600 / 3600.0) / gpu_s, "unit": "audio-hours/sec", "h2d_bytes_per_step": 2 * 600 * 16000, "d2h_bytes_per_step": 4 * 60000},
            "cpu_baseline": base}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="inference", choices=["inference", "features", "corpus", "cli"])
    ap.add_argument("--corpus", action="store_true", help="same as --config corpus")
    ap.add_argument("--channels", type=int, default=6, help="channels per GPU per step (one meeting)")
    ap.add_argument("--minutes", type=float, default=60.0, help="minutes of audio per channel")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "split"], help="conv-stack arithmetic (DESIGN.md section 7)")
    ap.add_argument("--chunk-rows", type=int, default=0, help="window starts per pass of the conv stack (0 = the library default, 32768)")
    ap.add_argument("--corpus-meetings", type=int, default=75)
    ap.add_argument("--feature-channels", type=int, default=64, help="--config features: channel-hours per ld_fbank_i16 call")
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0, help="audio seconds per CPU-reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps for the secondary `train` object (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=256, help="training windows per GPU per step")
    args = ap.parse_args()
    if args.corpus:
        args.config = "corpus"
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "features":
        run_features(args)
    elif args.config == "corpus":
        run_corpus(args)
    elif args.config == "cli":
        run_cli(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
