"""Bring-up diagnostics on a B200: every kernel against the CPU oracle, with per-plane localisation
for the conv stack.  Usage: python tools/gpu_check.py <section> [...]; sections: gemm fbank seg net perf.
Not part of the product; complements tests/ (-m gpu) with more verbose output for debugging.
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from laughter_detection_icsi_b200 import _native  # noqa: E402
from laughter_detection_icsi_b200.engine import Engine  # noqa: E402
from oracle import fbank_oracle, resnet_oracle, segmenter_oracle  # noqa: E402


def synth_pcm(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = 0.02 * rng.normal(size=n) + 0.25 * np.sin(2 * np.pi * 233.0 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5 * t)) + 0.003
    return np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)


def section_net():
    from plan_emulator import PlanEmulator
    eng = Engine(0, chunk_rows=512)
    sd = resnet_oracle.random_state_dict(seed=11)
    eng.load_state_dict(sd)
    rng = np.random.default_rng(2)
    T = 200
    feats = rng.normal(-4.0, 3.0, (T, 44)).astype(np.float32)
    probs = eng.infer_windows(torch.from_numpy(feats).cuda()).cpu().numpy()
    ref = resnet_oracle.window_probs(sd, feats)
    print(f"[net] single channel T={T}: max|gpu-oracle| = {np.abs(probs - ref).max():.3e}  (probs {ref.min():.4f}..{ref.max():.4f})")
    # per-plane localisation against the fp16-rounding emulator
    plan = _native.plan_json(eng.cfg)
    emu = PlanEmulator(plan, sd, half=True)
    nb = T + plan["H"]
    emu_probs = emu.run(torch.from_numpy(feats), nb).numpy()[:T]
    print(f"[net] emulator(fp16) vs oracle: {np.abs(emu_probs - ref).max():.3e}; gpu vs emulator: {np.abs(emu_probs - probs).max():.3e}")
    rows = nb + plan["H"]
    worst = []
    for p in plan["planes"]:
        g = eng.read_plane(p["id"], rows, p["wp"], p["C"])
        e = emu.plane_as_rows(p["id"], rows).numpy()
        d = np.abs(g - e)
        bad = np.argwhere(d > 1e-2 + 1e-2 * np.abs(e))
        worst.append((float(d.max()), p["id"], p["tag"], len(bad), bad[:3].tolist()))
    nbad = 0
    for dmax, pid, tag, cnt, where in worst:
        if cnt:
            nbad += 1
            if nbad <= 12:
                print(f"[net]   plane {pid:3d} {tag:28s} max diff {dmax:.3e} mismatches {cnt} first {where}")
    print(f"[net] planes with mismatches: {nbad} of {len(worst)}; overall max plane diff {max(w[0] for w in worst):.3e}")
    # multi-channel + multi-chunk
    T2 = [230, 1500, 101]
    feats2 = rng.normal(-4.0, 3.0, (sum(T2), 44)).astype(np.float32)
    probs2 = eng.infer_windows(torch.from_numpy(feats2).cuda(), T2).cpu().numpy()
    off = 0
    for t in T2:
        ref2 = resnet_oracle.window_probs(sd, feats2[off:off + t])
        print(f"[net] channel T={t}: max|gpu-oracle| = {np.abs(probs2[off:off + t] - ref2).max():.3e}")
        off += t


def section_fbank():
    for mode, name, mel in ((_native.LD_PREPROC_UTTERANCE, "utterance", "lhotse"), (_native.LD_PREPROC_FRAME, "frame", "kaldi")):
        eng = Engine(0, chunk_rows=256, fbank_preproc=mode)
        for n in (400, 16037, 160000 + 37):
            pcm = synth_pcm(n, seed=n)
            feats, frames = eng.fbank(torch.from_numpy(pcm).cuda(), mel=mel)
            x = pcm.astype(np.float32) / 32768.0
            ref = fbank_oracle.fbank(x, mel=mel, preproc=name).numpy()
            ref64 = fbank_oracle.fbank(x.astype(np.float64), mel=mel, preproc=name, dtype=torch.float64).numpy()
            g = feats.cpu().numpy()
            rel = np.abs(g - ref64) / np.maximum(1.0, np.abs(ref64))
            print(f"[fbank] {name}/{mel} n={n} T={frames[0]} max|gpu-f32 oracle|={np.abs(g - ref).max():.3e} "
                  f"max|gpu-f64 oracle|={np.abs(g - ref64).max():.3e} (rel {rel.max():.3e}); f32 oracle vs f64: {np.abs(ref - ref64).max():.3e}")
        # two channels in one call
        a, b = synth_pcm(8000, 1), synth_pcm(12345, 2)
        feats, frames = eng.fbank(torch.from_numpy(np.concatenate([a, b])).cuda(), [len(a), len(b)], mel=mel)
        ra = fbank_oracle.fbank(a.astype(np.float32) / 32768.0, mel=mel, preproc=name).numpy()
        rb = fbank_oracle.fbank(b.astype(np.float32) / 32768.0, mel=mel, preproc=name).numpy()
        g = feats.cpu().numpy()
        print(f"[fbank] {name} two channels: {np.abs(g[:frames[0]] - ra).max():.3e} {np.abs(g[frames[0]:] - rb).max():.3e}")
        eng.close()


def section_seg():
    eng = Engine(0, chunk_rows=256)
    rng = np.random.default_rng(4)
    z = np.cumsum(rng.normal(0, 0.35, 50000))
    p = (1.0 / (1.0 + np.exp(-(z - z.mean())))).astype(np.float32)
    p[100] = 1.5; p[200] = -0.2; p[300] = 0.0
    thr = [0.0, 0.3, 0.5, 0.9, 1.0]
    T = [20000, 30000]
    runs = eng.segment_runs(torch.from_numpy(p).cuda(), [float(np.float32(t)) for t in thr], thr, T)
    ok = True
    for k, t in enumerate(thr):
        exp = []
        off = 0
        for ci, n in enumerate(T):
            exp += [(s, e, ci) for s, e in segmenter_oracle.runs_above(p[off:off + n], t)]
            off += n
        got = list(zip(runs[k][0].tolist(), runs[k][1].tolist(), runs[k][2].tolist()))
        same = got == exp
        ok = ok and same
        print(f"[seg] thr={t}: {len(got)} runs, match={same}")
    y = eng.lowpass(torch.from_numpy(p).cuda()).cpu().numpy()
    ref = segmenter_oracle.lowpass(p)
    print(f"[seg] filtfilt max|gpu-scipy| = {np.abs(y - ref).max():.3e}; all runs match: {ok}")


def section_perf():
    eng = Engine(0)
    sd = resnet_oracle.random_state_dict(seed=11)
    eng.load_state_dict(sd)
    n = 16000 * 600
    pcm = torch.from_numpy(synth_pcm(n, 3)).cuda()
    for name, reps in (("10min", 3),):
        for _ in range(2):
            feats, frames = eng.fbank(pcm)
            probs = eng.infer_windows(feats)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(reps):
            feats, frames = eng.fbank(pcm)
        e1.record()
        for _ in range(reps):
            probs = eng.infer_windows(feats)
        e2.record()
        torch.cuda.synchronize()
        tf, tn = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
        hours = n / 16000 / 3600
        print(f"[perf] {name}: fbank {tf:.3f} ms ({frames[0] * 496 / tf / 1e6:.1f} GB/s algorithmic), net {tn:.2f} ms "
              f"-> {hours / ((tf + tn) / 1e3):.2f} audio-h/s; executed {eng.macs_per_row * 2 * frames[0] / tn / 1e9:.1f} TFLOP/s, "
              f"dense-equivalent {1.416661568e9 * frames[0] / tn / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    for sec in sys.argv[1:]:
        t0 = time.time()
        try:
            globals()["section_" + sec]()
        except Exception:
            traceback.print_exc()
            print(f"[{sec}] FAILED")
        print(f"[{sec}] {time.time() - t0:.1f} s", flush=True)
