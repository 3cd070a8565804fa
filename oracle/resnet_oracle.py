"""CPU restatement of ResNetBigger's eval-mode forward pass and of the InferenceDataset windowing.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows the reference's models.py:181-244
(ResNetBigger), models.py:82-115 (ResidualBlock) and datasets.py:72-93 (InferenceDataset), written as a
function of a state_dict so that no nn.Module of the product is involved.  Pinned by
tests/golden/resnet_golden.npz, produced by the reference's own models.py (tests/golden/make_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

PARAM_SHAPES_CACHE = {}


def param_shapes(filter_sizes=(64, 32, 16, 16), linear_layer_size=48):
    """Ordered (name, shape) of the 150 state_dict entries of ResNetBigger (probe in SURVEY.md section 5)."""
    out = []

    def bn(prefix, c):
        out.extend([(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
                    (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())])

    out.append(("conv1.weight", (64, 1, 3, 3)))
    bn("bn1", 64)
    cin = 64
    for b, cout in enumerate(filter_sizes, start=1):
        stride = 1 if b == 1 else 2
        for r in range(2):
            p = f"block{b}.{r}"
            ic, s = (cin, stride) if r == 0 else (cout, 1)
            out.extend([(p + ".conv1.weight", (cout, ic, 3, 3)), (p + ".conv1.bias", (cout,))])
            bn(p + ".bn1", cout)
            out.extend([(p + ".conv2.weight", (cout, cout, 3, 3)), (p + ".conv2.bias", (cout,))])
            bn(p + ".bn2", cout)
            if s != 1 or ic != cout:
                out.append((p + ".shortcut.0.weight", (cout, ic, 1, 1)))
                bn(p + ".shortcut.1", cout)
        cin = cout
    bn("bn2", linear_layer_size)
    bn("bn3", 32)
    out.extend([("linear1.weight", (32, linear_layer_size)), ("linear1.bias", (32,)),
                ("linear2.weight", (1, 32)), ("linear2.bias", (1,))])
    return out


def random_state_dict(seed, filter_sizes=(64, 32, 16, 16), linear_layer_size=48, head_gain=1.0):
    """Seeded synthetic checkpoint: fan-in-scaled weights, BatchNorm statistics/affine away from identity
    (SURVEY.md section 8d).  numpy's PCG64 stream is version-stable, so fixtures can store outputs only."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in param_shapes(filter_sizes, linear_layer_size):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif name.endswith("running_mean"):
            sd[name] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif name.endswith("running_var"):
            sd[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif ".bn" in name or name.startswith("bn") or ".shortcut.1" in name:
            if name.endswith("weight"):
                sd[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
            else:
                sd[name] = torch.from_numpy(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif name.endswith("bias"):
            sd[name] = torch.from_numpy(rng.uniform(-0.05, 0.05, shape).astype(np.float32))
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            sd[name] = torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))
    sd["linear2.weight"] = sd["linear2.weight"] * head_gain
    return sd


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


def _residual_block(sd, p, x, stride):
    h = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=stride, padding=1)))
    h = _bn(sd, p + ".bn2", F.conv2d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=1, padding=1))
    if (p + ".shortcut.0.weight") in sd:
        x = _bn(sd, p + ".shortcut.1", F.conv2d(x, sd[p + ".shortcut.0.weight"], None, stride=stride))
    return F.relu(h + x)


def forward(sd, x, return_logit=False):
    """x: (B, 1, 100, 44) -> (B, 1) sigmoid output, eval mode (BatchNorm running stats, dropout off)."""
    dt = x.dtype
    sd = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    out = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], None, stride=1, padding=1)))
    for b in range(1, 5):
        out = _residual_block(sd, f"block{b}.0", out, 1 if b == 1 else 2)
        out = _residual_block(sd, f"block{b}.1", out, 1)
    out = F.avg_pool2d(out, 4)
    out = out.reshape(out.shape[0], -1)
    out = _bn(sd, "bn2", out)
    out = F.linear(out, sd["linear1.weight"], sd["linear1.bias"])
    out = F.relu(_bn(sd, "bn3", out))
    out = F.linear(out, sd["linear2.weight"], sd["linear2.bias"])
    return out if return_logit else torch.sigmoid(out)


def window(feats, index, n_frames=100):
    """InferenceDataset.__getitem__: feats[index:index+100], right-padded with ZERO rows."""
    w = np.asarray(feats)[index:index + n_frames]
    if w.shape[0] != n_frames:
        w = np.pad(w, ((0, n_frames - w.shape[0]), (0, 0)))
    return w


def window_probs(sd, feats, batch_size=32, dtype=torch.float32, start=0, stop=None, return_logit=False):
    """The inference loop of segment_laughter.load_and_pred: one probability per frame."""
    feats = np.asarray(feats, dtype=np.float32)
    stop = len(feats) if stop is None else stop
    out = []
    with torch.no_grad():
        for i0 in range(start, stop, batch_size):
            idx = range(i0, min(stop, i0 + batch_size))
            x = torch.from_numpy(np.stack([window(feats, i) for i in idx]))[:, None].to(dtype)
            out.append(forward(sd, x, return_logit).reshape(-1))
    return torch.cat(out).numpy()


def calibrate_head(sd, feats, n_windows=512, target_std=2.0):
    """Rescale linear2 so that logits over the first windows have mean 0 / std `target_std`: a random-init
    network otherwise emits probabilities in a ~1e-3 wide band around 0.5."""
    n = min(n_windows, len(feats))
    z = window_probs(sd, feats, stop=n, dtype=torch.float64, return_logit=True)
    gain = target_std / max(float(z.std()), 1e-12)
    sd = dict(sd)
    sd["linear2.weight"] = sd["linear2.weight"] * gain
    sd["linear2.bias"] = (sd["linear2.bias"] - float(z.mean())) * gain
    return sd
