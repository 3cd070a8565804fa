"""Synthetic meeting-shaped inputs for benchmarks and tests: there is no network for the
ICSI corpus or a trained checkpoint, so audio and weights are generated from seeds.

Audio: 16 kHz int16 mono; white-noise floor (sigma 0.02) plus voiced "laugh-like" bursts roughly every 7 s
(0.3-1.5 s long, f0 in [180, 420] Hz, 5 harmonics, 5 Hz amplitude modulation, amplitude in [0.05, 0.4]).
Checkpoint: ResNetBigger(resnet_base) with fan-in-scaled weights, BatchNorm statistics/affine away from identity,
and a head gain so that probabilities spread over (0, 1) -- a default-initialised network emits a ~1e-3 wide band
around 0.5, which makes thresholding vacuous.
"""
import math

import numpy as np
import torch

SAMPLE_RATE = 16000
BASE_SEED = 20221018

# linear2 calibration for synthetic_state_dict(seed=BASE_SEED) on synth audio features (logit mean 0, std ~2 over the
# first 4096 windows of synth_channel(…, meeting=0, channel=0)); computed once with the fp64 oracle, see DESIGN.md.
HEAD_GAIN = 217.7408888272181
HEAD_BIAS_SHIFT = 0.3061996540460959


def synth_channel(n_samples, seed=BASE_SEED, meeting=0, channel=0, device="cpu"):
    """int16 tensor (n_samples,) on `device`; the stream is keyed by (seed, meeting, channel)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed((seed * 1000003 + meeting * 1009 + channel) % (2 ** 63 - 1))
    x = torch.randn(n_samples, generator=g, device=dev, dtype=torch.float32) * 0.02
    host = np.random.default_rng([seed, meeting, channel])
    t0 = float(host.uniform(0.5, 7.0))
    two_pi = 2.0 * math.pi
    while t0 * SAMPLE_RATE < n_samples - 1:
        dur = float(host.uniform(0.3, 1.5))
        f0 = float(host.uniform(180.0, 420.0))
        amp = float(host.uniform(0.05, 0.4))
        a = int(t0 * SAMPLE_RATE)
        b = min(n_samples, a + int(dur * SAMPLE_RATE))
        t = torch.arange(b - a, device=dev, dtype=torch.float32) / SAMPLE_RATE
        env = torch.sin(math.pi * t / max(dur, 1e-3)).clamp_min(0.0) * (0.6 + 0.4 * torch.sin(two_pi * 5.0 * t))
        tone = sum(torch.sin(two_pi * h * f0 * t) / h for h in range(1, 6))
        x[a:b] += amp * env * tone
        t0 += dur + float(host.uniform(3.0, 11.0))
    return torch.clamp(torch.round(x.clamp(-1.0, 1.0) * 32767.0), -32768, 32767).to(torch.int16)


def synth_meeting(n_channels, n_samples, seed=BASE_SEED, meeting=0, device="cpu"):
    """(n_channels * n_samples,) int16, channels laid end to end, plus the per-channel lengths."""
    chans = [synth_channel(n_samples, seed, meeting, c, device) for c in range(n_channels)]
    return torch.cat(chans), [n_samples] * n_channels


def param_shapes(filter_sizes=(64, 32, 16, 16), linear_layer_size=48):
    """(name, shape) of ResNetBigger's state_dict in registration order."""
    out = []

    def bn(prefix, c):
        for leaf, shape in (("weight", (c,)), ("bias", (c,)), ("running_mean", (c,)), ("running_var", (c,)),
                            ("num_batches_tracked", ())):
            out.append((f"{prefix}.{leaf}", shape))

    out.append(("conv1.weight", (64, 1, 3, 3)))
    bn("bn1", 64)
    cin = 64
    for b, cout in enumerate(filter_sizes, start=1):
        for r in range(2):
            ic, s = (cin, 1 if b == 1 else 2) if r == 0 else (cout, 1)
            p = f"block{b}.{r}"
            out += [(p + ".conv1.weight", (cout, ic, 3, 3)), (p + ".conv1.bias", (cout,))]
            bn(p + ".bn1", cout)
            out += [(p + ".conv2.weight", (cout, cout, 3, 3)), (p + ".conv2.bias", (cout,))]
            bn(p + ".bn2", cout)
            if s != 1 or ic != cout:
                out.append((p + ".shortcut.0.weight", (cout, ic, 1, 1)))
                bn(p + ".shortcut.1", cout)
        cin = cout
    bn("bn2", linear_layer_size)
    bn("bn3", 32)
    out += [("linear1.weight", (32, linear_layer_size)), ("linear1.bias", (32,)), ("linear2.weight", (1, 32)),
            ("linear2.bias", (1,))]
    return out


def synthetic_state_dict(seed=BASE_SEED, filter_sizes=(64, 32, 16, 16), linear_layer_size=48, head_gain=None,
                         head_bias_shift=None):
    """Seeded checkpoint in the reference's state_dict layout (same stream as oracle.resnet_oracle.random_state_dict)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in param_shapes(filter_sizes, linear_layer_size):
        leaf = name.rsplit(".", 1)[1]
        is_bn = ".bn" in name or name.startswith("bn") or ".shortcut.1" in name
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif leaf == "running_mean":
            sd[name] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif leaf == "running_var":
            sd[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif is_bn and leaf == "weight":
            sd[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif is_bn:
            sd[name] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif leaf == "bias":
            sd[name] = torch.from_numpy(rng.uniform(-0.05, 0.05, shape).astype(np.float32))
        else:
            bound = 1.0 / math.sqrt(int(np.prod(shape[1:])))
            sd[name] = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
    gain = HEAD_GAIN if head_gain is None else head_gain
    shift = HEAD_BIAS_SHIFT if head_bias_shift is None else head_bias_shift
    sd["linear2.bias"] = (sd["linear2.bias"] - shift) * gain
    sd["linear2.weight"] = sd["linear2.weight"] * gain
    return sd


def eval_grid():
    """The reference's evaluation grid (cluster_scripts/gen_eval_exp.py:30-36): 29 thresholds x 3 min lengths."""
    thr = [round(float(t), 2) for t in np.linspace(0, 0.9, 19)] + [round(float(t), 2) for t in np.linspace(0.91, 1, 10)]
    return thr, [0.0, 0.1, 0.2]
