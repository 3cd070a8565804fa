"""The oracle against the fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import fbank_oracle, resnet_oracle, segmenter_oracle
from golden.make_golden_inputs import expand_probs, golden_inputs_resnet


def test_resnet_oracle_matches_reference_models_py(golden_dir):
    g = np.load(os.path.join(golden_dir, "resnet_golden.npz"))
    sd = resnet_oracle.random_state_dict(int(g["sd_seed"]))
    x = torch.from_numpy(golden_inputs_resnet(int(g["x_seed"])))
    y = resnet_oracle.forward(sd, x).numpy()
    assert np.abs(y - g["probs"]).max() < 2e-7
    y64 = resnet_oracle.forward(sd, x.double()).numpy()
    assert np.abs(y64 - g["probs_f64"]).max() < 1e-12


def test_resnet_state_dict_layout_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "resnet_state_dict_keys.json")) as f:
        ref = json.load(f)
    shapes = resnet_oracle.param_shapes()
    assert [n for n, _ in shapes] == list(ref.keys())
    assert all(list(s) == ref[n] for n, s in shapes)
    assert len(shapes) == 150
    assert sum(int(np.prod(s)) for n, s in shapes if not n.endswith(("running_mean", "running_var", "num_batches_tracked"))) == 221217


def test_segmenter_oracle_matches_reference_laugh_segmenter(golden_dir):
    with open(os.path.join(golden_dir, "segmenter_golden.json")) as f:
        g = json.load(f)
    np_major = int(np.__version__.split(".")[0])
    gold_major = int(g["numpy_version"].split(".")[0])
    kept = 0
    for c in g["cases"]:
        if c["name"].startswith("ties_f32") and np_major != gold_major:
            continue  # float32 threshold ties follow NumPy's scalar promotion rules, which changed in NumPy 2
        probs = np.array(expand_probs(c), dtype=c["dtype"])
        got = segmenter_oracle.get_laughter_instances(probs, c["thresholds"], c["min_lengths"], c["fps"])
        exp = {(t, m): [tuple(x) for x in inst] for t, m, inst in c["expected"]}
        assert list(got.keys()) == list(exp.keys()), c["name"]
        assert got == exp, c["name"]
        if c["name"].startswith("minlen_edge_"):
            kept += len(got[(0.5, 0.2)])
    assert kept == 46  # float64 rounding keeps 46 of the 200 exactly-20-frame spans


def test_known_answer_vector():
    probs = [0.1, 0.6, 0.7, 0.2, 0.9, 0.9, 0.9, 0.9, 0.5, 0.5, 0.51, 0.0, 1.0, 1.2, -0.1, 0.8]
    d = segmenter_oracle.get_laughter_instances(probs, [0.5, 0.0, 1.0], [0.0, 0.02], 100.0)
    assert d[(0.5, 0.0)] == [(0.01, 0.02), (0.04, 0.07), (0.12, 0.13)]
    assert d[(0.5, 0.02)] == [(0.04, 0.07)]
    assert d[(0.0, 0.0)] == [(0.0, 0.15)]
    assert d[(1.0, 0.0)] == []


def test_lowpass_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "lowpass_golden.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    z = np.cumsum(rng.normal(0, 0.3, int(g["n"])))
    p = (1.0 / (1.0 + np.exp(-(z - z.mean())))).astype(np.float32)
    assert np.abs(segmenter_oracle.lowpass(p) - g["out"]).max() < 1e-9


def test_fbank_oracle_frame_mode_matches_torchaudio_kaldi(golden_dir):
    for n in (400, 16037):
        g = np.load(os.path.join(golden_dir, f"fbank_kaldi_{n}.npz"))
        x = g["pcm"].astype(np.float32) / 32768.0
        out = fbank_oracle.fbank(x, mel="kaldi", preproc="frame").numpy()
        assert out.shape == g["feats"].shape == (fbank_oracle.num_frames(n), 44)
        assert np.abs(out - g["feats"]).max() < 5e-5


def test_fbank_oracle_frame_counts_and_silence():
    for n, t in ((399, 2), (400, 3), (16000, 100), (16037, 100), (160037, 1000), (9600000, 60000)):
        assert fbank_oracle.num_frames(n) == t
    out = fbank_oracle.fbank(np.zeros(16000, dtype=np.float32)).numpy()
    assert np.allclose(out, np.log(np.finfo(np.float32).eps))  # silence -> log(eps) = -15.9424
    m = fbank_oracle.mel_matrix_lhotse()
    assert m.shape == (257, 44) and np.all(m[256] == 0) and np.all(m >= 0) and m.max() <= 1.0
    assert np.all((m > 0).sum(axis=1) <= 2)  # triangular bank: a bin feeds at most two filters


def test_train_mode_oracle_matches_reference_autograd(golden_dir):
    """oracle.forward_train + autograd against the reference module in .train() mode (dropout 0)."""
    g = np.load(os.path.join(golden_dir, "resnet_train_golden.npz"))
    sd = resnet_oracle.random_state_dict(int(g["sd_seed"]))
    rng = np.random.default_rng(int(g["x_seed"]))
    x = torch.from_numpy(rng.normal(-4.0, 3.0, (8, 1, 100, 44)).astype(np.float32))
    labels = torch.from_numpy((rng.uniform(size=8) < 0.5).astype(np.float32))
    probs, loss, grads, stats = resnet_oracle.train_step_reference(sd, x, labels, torch.ones(8, 48), torch.ones(8, 32), 0.0)
    assert np.abs(probs.numpy() - g["probs"]).max() < 1e-12 and abs(loss - float(g["loss"])) < 1e-12
    for key in g.files:
        if "__" in key:
            assert np.abs(grads[key.replace("__", ".")].numpy() - g[key]).max() < 1e-10, key
    # running statistics the reference module accumulated in that one step: 0.9 * init + 0.1 * batch (unbiased variance)
    mean, var = stats["bn1"]
    assert np.abs(0.9 * sd["bn1.running_mean"].double().numpy() + 0.1 * mean.numpy() - g["running_mean_bn1"]).max() < 1e-10
    n = 8 * 50 * 22
    var2 = stats["block2.0.bn1"][1].numpy() * n / (n - 1)
    assert np.abs(0.9 * sd["block2.0.bn1.running_var"].double().numpy() + 0.1 * var2 - g["running_var_block2"]).max() < 1e-10
