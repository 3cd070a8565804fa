"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and torchaudio):
    python tests/golden/make_golden.py
The reference's models.py and laugh_segmenter.py are imported unmodified (laugh_segmenter.py with a stub
``librosa`` module, which it imports but never uses on this path).  Inputs are derived from seeded
numpy PCG64 streams so the fixtures only need to store the reference's OUTPUTS.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.modules.setdefault("librosa", types.ModuleType("librosa"))

import laugh_segmenter as ref_seg  # noqa: E402  (reference)
import models as ref_models  # noqa: E402  (reference)

from oracle import resnet_oracle  # noqa: E402
from tests.golden.make_golden_inputs import expand_probs, golden_inputs_resnet  # noqa: E402


def make_resnet():
    sd = resnet_oracle.random_state_dict(seed=11)
    model = ref_models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
    model.load_state_dict(sd)
    model.eval()
    x = golden_inputs_resnet()
    with torch.no_grad():
        y = model(torch.from_numpy(x)).numpy()
        y64 = model.double()(torch.from_numpy(x).double()).numpy()
    keys = [k for k in model.state_dict().keys()]
    np.savez(os.path.join(HERE, "resnet_golden.npz"), probs=y, probs_f64=y64, sd_seed=11, x_seed=5,
             n_keys=len(keys))
    with open(os.path.join(HERE, "resnet_state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in model.state_dict().items()}, f, indent=0)
    print("resnet golden:", y.reshape(-1))


def make_resnet_train():
    """Reference models.py in .train() mode (dropout_rate=0 so that no random mask is involved): outputs, BCE loss and
    the gradients of a few parameter tensors from the reference module's own autograd -- pins oracle.forward_train."""
    sd = resnet_oracle.random_state_dict(seed=12)
    model = ref_models.ResNetBigger(dropout_rate=0.0, linear_layer_size=48, filter_sizes=[64, 32, 16, 16]).double()
    model.load_state_dict(sd)
    model.train()
    rng = np.random.default_rng(13)
    x = rng.normal(-4.0, 3.0, (8, 1, 100, 44)).astype(np.float32)
    labels = (rng.uniform(size=8) < 0.5).astype(np.float32)
    out = model(torch.from_numpy(x).double()).squeeze()
    loss = torch.nn.BCELoss()(out, torch.from_numpy(labels).double())
    loss.backward()
    keep = ["conv1.weight", "bn1.weight", "block1.0.conv2.weight", "block2.0.shortcut.0.weight", "block2.0.shortcut.1.bias",
            "block3.1.bn2.weight", "block4.0.conv1.weight", "bn2.bias", "linear1.weight", "linear2.bias"]
    grads = {k.replace(".", "__"): dict(model.named_parameters())[k].grad.numpy() for k in keep}
    np.savez(os.path.join(HERE, "resnet_train_golden.npz"), sd_seed=12, x_seed=13, probs=out.detach().numpy(), loss=float(loss),
             running_mean_bn1=model.bn1.running_mean.numpy(), running_var_block2=model.block2[0].bn1.running_var.numpy(), **grads)
    print("resnet train golden: loss", float(loss))


def seg_cases():
    rng = np.random.default_rng(7)
    grid_thr = [round(float(t), 2) for t in np.linspace(0, 0.9, 19)] + [round(float(t), 2) for t in np.linspace(0.91, 1, 10)]
    cases = []
    # SURVEY.md section 8c known-answer vector
    ka = [0.1, 0.6, 0.7, 0.2, 0.9, 0.9, 0.9, 0.9, 0.5, 0.5, 0.51, 0.0, 1.0, 1.2, -0.1, 0.8]
    cases.append(dict(name="known_answer", probs=ka, dtype="float64", thresholds=[0.5, 0.0, 1.0], min_lengths=[0.0, 0.02], fps=100.0))
    cases.append(dict(name="known_answer_f32", probs=ka, dtype="float32", thresholds=[0.5, 0.0, 1.0], min_lengths=[0.0, 0.02], fps=100.0))
    # smooth random walk squashed to (0,1): realistic run structure, full eval grid, non-integer fps
    z = np.cumsum(rng.normal(0, 0.35, 3000))
    p = 1.0 / (1.0 + np.exp(-(z - z.mean())))
    cases.append(dict(name="walk_f32_grid", probs=[float(v) for v in p.astype(np.float32)], dtype="float32",
                      thresholds=grid_thr, min_lengths=[0.0, 0.1, 0.2], fps=3000 / 30.0037))
    cases.append(dict(name="walk_f64", probs=[float(v) for v in p], dtype="float64", thresholds=[0.3, 0.5, 0.8],
                      min_lengths=[0.2], fps=100.0))
    # threshold ties: p == thr exactly (float32-representable and not)
    tie = [0.5, 0.5, 0.75, 0.5, 0.3, 0.3, 0.30000001192092896, 0.9, 0.9, 0.1]
    cases.append(dict(name="ties_f32", probs=tie, dtype="float32", thresholds=[0.5, 0.3, 0.9], min_lengths=[0.0], fps=100.0))
    cases.append(dict(name="ties_f64", probs=tie, dtype="float64", thresholds=[0.5, 0.3, 0.9], min_lengths=[0.0], fps=100.0))
    # min-length float64 edge: 21 frames above threshold at every start 0..199
    for start in range(0, 200):
        cases.append(dict(name=f"minlen_edge_{start}", gen={"kind": "minlen_edge", "start": start}, dtype="float32",
                          thresholds=[0.5], min_lengths=[0.2], fps=100.0))
    # degenerate: all above, none above, single frames, run touching the end, empty
    cases.append(dict(name="all_above", probs=[0.9] * 50, dtype="float32", thresholds=[0.5], min_lengths=[0.0, 0.2], fps=100.0))
    cases.append(dict(name="none_above", probs=[0.1] * 50, dtype="float32", thresholds=[0.5], min_lengths=[0.0], fps=100.0))
    cases.append(dict(name="singles", probs=[0.9, 0.1] * 20 + [0.9], dtype="float32", thresholds=[0.5], min_lengths=[0.0], fps=100.0))
    cases.append(dict(name="empty", probs=[], dtype="float32", thresholds=[0.5], min_lengths=[0.0], fps=100.0))
    return cases


def make_segmenter():
    import contextlib
    import io
    out = []
    for c in seg_cases():
        probs = np.array(expand_probs(c), dtype=c["dtype"])
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a WARN line per clamped frame
            d = ref_seg.get_laughter_instances(probs, thresholds=c["thresholds"], min_lengths=c["min_lengths"], fps=c["fps"])
        c = dict(c)
        c["expected"] = [[thr, ml, [[float(a), float(b)] for a, b in inst]] for (thr, ml), inst in d.items()]
        out.append(c)
    kept = sum(1 for c in out if c["name"].startswith("minlen_edge_") and len(c["expected"][0][2]) == 1)
    print("segmenter golden: min-length edge kept", kept, "of 200; numpy", np.__version__)
    with open(os.path.join(HERE, "segmenter_golden.json"), "w") as f:
        json.dump({"numpy_version": np.__version__, "cases": out}, f)


def make_lowpass():
    rng = np.random.default_rng(3)
    z = np.cumsum(rng.normal(0, 0.3, 4000))
    p = (1.0 / (1.0 + np.exp(-(z - z.mean())))).astype(np.float32)
    y = ref_seg.lowpass(p)
    np.savez(os.path.join(HERE, "lowpass_golden.npz"), seed=3, n=4000, out=y)
    print("lowpass golden: min", y.min(), "max", y.max())


def make_fbank():
    import torchaudio
    rng = np.random.default_rng(9)
    for n in (400, 16037):
        t = np.arange(n) / 16000.0
        x = 0.02 * rng.normal(size=n) + 0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5 * t))
        pcm = np.clip(np.round(x * 32767), -32768, 32767).astype(np.int16)
        wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]
        ref = torchaudio.compliance.kaldi.fbank(wav, num_mel_bins=44, frame_length=25.0, frame_shift=10.0, snip_edges=False,
                                                dither=0.0, energy_floor=0.0, low_freq=20.0, high_freq=-400.0,
                                                sample_frequency=16000.0, preemphasis_coefficient=0.97, remove_dc_offset=True,
                                                window_type="povey", use_energy=False, use_log_fbank=True, use_power=True)
        np.savez(os.path.join(HERE, f"fbank_kaldi_{n}.npz"), seed=9, n=n, pcm=pcm, feats=ref.numpy())
        print("fbank golden", n, ref.shape)


if __name__ == "__main__":
    make_resnet()
    make_resnet_train()
    make_segmenter()
    make_lowpass()
    make_fbank()
