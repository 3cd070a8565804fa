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


def window_probs_autograd(sd, feats, batch_size=32):
    """window_probs the way the reference literally runs it: segment_laughter.py:95 calls ``model(x).cpu().detach()`` WITHOUT
    torch.no_grad(), so every batch builds (and drops) an autograd graph over parameters that require grad.  Same
    numbers, more CPU time -- used by bench.py's CPU baseline only."""
    feats = np.asarray(feats, dtype=np.float32)
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    out = []
    for i0 in range(0, len(feats), batch_size):
        idx = range(i0, min(len(feats), i0 + batch_size))
        x = torch.from_numpy(np.stack([window(feats, i) for i in idx]))[:, None]
        out.append(forward(sd, x).detach().reshape(-1))
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


# ------------------------------------------------------------------------------------------------ training mode
_QUANT = None   # optional rounding model of the CUDA path (bf16 storage), see forward_train(quant=...)


def _q(t):
    """Straight-through rounding: value rounded like the kernels store it, gradient of the identity."""
    return t if _QUANT is None else t + (_QUANT(t) - t).detach()


def _conv_q(x, w, b, **kw):
    return _q(F.conv2d(_q(x) if _QUANT is None else x, _q(w), None if _QUANT is not None else b, **kw)) if _QUANT is not None \
        else F.conv2d(x, w, b, **kw)


def _bn_train(sd, prefix, x, stats=None):
    """BatchNorm in .train() mode: batch statistics (biased variance) normalise, like nn.BatchNorm2d/1d."""
    if stats is not None:
        dims = [d for d in range(x.dim()) if d != 1]
        stats[prefix] = (x.mean(dim=dims).detach(), x.var(dim=dims, unbiased=False).detach())
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], training=True, eps=1e-5)


def _residual_block_train(sd, p, x, stride, stats):
    h = _q(F.relu(_bn_train(sd, p + ".bn1", _conv_q(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=stride, padding=1), stats)))
    h = _bn_train(sd, p + ".bn2", _conv_q(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=1, padding=1), stats)
    if (p + ".shortcut.0.weight") in sd:
        x = _bn_train(sd, p + ".shortcut.1", _conv_q(x, sd[p + ".shortcut.0.weight"], None, stride=stride), stats)
    return _q(F.relu(h + x))


def forward_train(sd, x, mask1, mask2, dropout_p, stats=None, quant=None):
    """ResNetBigger.forward in .train() mode (models.py:222-239 with BatchNorm batch statistics and nn.Dropout at
    :232 and :235), as a function of a state_dict whose tensors may require grad.  The two dropout sites take explicit
    0/1 keep masks (kept units scaled by 1/(1-p), like nn.Dropout) so that the CUDA path can be given the same masks."""
    global _QUANT
    _QUANT = quant   # e.g. lambda t: t.bfloat16().to(t.dtype): rounds conv weights and stored activations like the CUDA path
    try:
        return _forward_train(sd, x, mask1, mask2, dropout_p, stats)
    finally:
        _QUANT = None


def _forward_train(sd, x, mask1, mask2, dropout_p, stats):
    scale = 1.0 / (1.0 - dropout_p)
    out = _q(F.relu(_bn_train(sd, "bn1", _q(F.conv2d(x, sd["conv1.weight"], None, stride=1, padding=1)), stats)))
    for b in range(1, 5):
        out = _residual_block_train(sd, f"block{b}.0", out, 1 if b == 1 else 2, stats)
        out = _residual_block_train(sd, f"block{b}.1", out, 1, stats)
    out = F.avg_pool2d(out, 4)
    out = out.reshape(out.shape[0], -1)
    out = _bn_train(sd, "bn2", out, stats) * mask1 * scale
    out = F.linear(out, sd["linear1.weight"], sd["linear1.bias"])
    out = _bn_train(sd, "bn3", out, stats) * mask2 * scale
    out = F.relu(out)
    out = F.linear(out, sd["linear2.weight"], sd["linear2.bias"])
    return torch.sigmoid(out)


def train_step_reference(sd, x, labels, mask1, mask2, dropout_p, dtype=torch.float64, quant=None):
    """loss = BCELoss(model(x).squeeze(), labels) and its gradients (train.py:277-289), in `dtype` on the CPU.
    Returns (probs, loss, {name: grad}, {bn: (batch mean, biased var)})."""
    params = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))}
    stats = {}
    probs = forward_train(params, x.to(dtype), mask1.to(dtype), mask2.to(dtype), dropout_p, stats, quant)
    loss = F.binary_cross_entropy(probs.reshape(-1), labels.to(dtype))
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}
    return probs.detach().reshape(-1), float(loss), grads, stats
