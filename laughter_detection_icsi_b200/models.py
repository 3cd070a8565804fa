"""``ResNetBigger`` with the reference's constructor, attributes and state_dict layout (reference
models.py:82-115, 181-244) whose forward pass runs on the B200 kernels.

The module tree (conv1, bn1, block1..4 = 2 x ResidualBlock{conv1, bn1, conv2, bn2, shortcut}, bn2, bn3,
linear1, linear2) only HOLDS the parameters, so that the 150 state_dict keys/shapes match and reference
checkpoints load unchanged.  ``forward`` hands them to the C ABI (BatchNorm folded, convs as tcgen05
implicit GEMMs); there is no PyTorch/CPU compute path.
"""
import numpy as np
import torch
from torch import nn

from . import engine as _engine
from ._native import LdError


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=stride, padding=1, bias=True)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=1, padding=1, bias=True)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1), stride=stride, bias=False),
                nn.BatchNorm2d(out_channels))

    def forward(self, x):
        raise LdError("ResidualBlock is a parameter container; call ResNetBigger.forward (B200 kernels)")


class _TrainFunction(torch.autograd.Function):
    """probs = ResNetBigger_train(flat_params, x): forward/backward on the CUDA training network (ld_train_*)."""

    @staticmethod
    def forward(ctx, flat, feats, mask1, mask2, p, eng):
        probs, bn_stats = eng.train_forward(flat.detach().contiguous(), feats, mask1, mask2, p)
        ctx.eng = eng
        ctx.mark_non_differentiable(bn_stats)
        return probs, bn_stats

    @staticmethod
    def backward(ctx, dprobs, _dstats):
        return ctx.eng.train_backward(dprobs), None, None, None, None, None


class ResNetBigger(nn.Module):
    def __init__(self, num_classes=1, dropout_rate=0.5, linear_layer_size=192, filter_sizes=[64, 32, 16, 16]):
        super().__init__()
        print(f"training with dropout={dropout_rate}")
        if num_classes != 1:
            raise LdError("the B200 head kernel implements the reference's single-logit classifier (num_classes=1)")
        self.conv1 = nn.Conv2d(1, 64, kernel_size=(3, 3), stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.linear_layer_size = linear_layer_size
        self.filter_sizes = filter_sizes
        self.block1 = self._create_block(64, filter_sizes[0], stride=1)
        self.block2 = self._create_block(filter_sizes[0], filter_sizes[1], stride=2)
        self.block3 = self._create_block(filter_sizes[1], filter_sizes[2], stride=2)
        self.block4 = self._create_block(filter_sizes[2], filter_sizes[3], stride=2)
        self.bn2 = nn.BatchNorm1d(linear_layer_size)
        self.bn3 = nn.BatchNorm1d(32)
        self.linear1 = nn.Linear(linear_layer_size, 32)
        self.linear2 = nn.Linear(32, num_classes)
        self.dropout = nn.Dropout(dropout_rate)
        self.global_step = 0
        self.epoch = 0
        self.best_val_loss = np.inf
        self._ld_engine = None
        self._ld_fingerprint = None

    def _create_block(self, in_channels, out_channels, stride):
        return nn.Sequential(ResidualBlock(in_channels, out_channels, stride), ResidualBlock(out_channels, out_channels, 1))

    def set_device(self, device):
        for b in [self.block1, self.block2, self.block3, self.block4]:
            b.to(device)
        self.to(device)

    # -------------------------------------------------------------------------------------- B200 path
    def _fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def train(self, mode=True):
        self._ld_fingerprint = None   # parameters usually changed between a training phase and the next eval forward
        return super().train(mode)

    def mark_weights_dirty(self):
        """Call after writing parameters or BatchNorm buffers through raw pointers / ``.data`` (ld_clip_adam_step,
        init_weights): such writes do not bump ``tensor._version``, so the fingerprint below would not see them and the next
        eval forward would run on stale folded weights."""
        self._ld_fingerprint = None

    def b200_engine(self):
        """The process-wide ld_ctx of the device the parameters live on, with these weights loaded."""
        p = self.conv1.weight
        if not p.is_cuda:
            raise LdError("ResNetBigger parameters are not on a CUDA device: call model.set_device('cuda') "
                          "(this build has no CPU path)")
        eng = _engine.get_engine(p.device.index or 0, filter_sizes=tuple(self.filter_sizes),
                                 linear_layer_size=self.linear_layer_size)
        fp = self._fingerprint()
        if self._ld_engine is not eng or self._ld_fingerprint != fp or eng.weights_owner is not self:
            eng.load_state_dict(self.state_dict())
            eng.weights_owner = self
            self._ld_engine, self._ld_fingerprint = eng, fp
        return eng

    # -------------------------------------------------------------------------------------- training path
    def _train_engine(self, batch):
        p = self.conv1.weight
        if not p.is_cuda:
            raise LdError("ResNetBigger parameters are not on a CUDA device: call model.set_device('cuda') "
                          "(this build has no CPU path)")
        eng = _engine.get_engine(p.device.index or 0, filter_sizes=tuple(self.filter_sizes),
                                 linear_layer_size=self.linear_layer_size)
        if getattr(eng, "train_table", None) is None or eng.train_max_batch < batch:
            eng.train_create(max(int(batch), 256))
            names = [n for n, _ in self.named_parameters()]
            table = eng.train_table["params"]
            if [t[0] for t in table] != names or [t[2] for t in table] != [q.numel() for q in self.parameters()]:
                raise LdError("parameter order of the module and of the CUDA training network disagree")
        return eng

    # ---- flat parameter storage: one fp32 vector whose slices ARE the parameters (views), so the training kernels read it
    #      directly and the fused optimiser (train.B200Adam / ld_clip_adam_step) updates every parameter in one launch
    def flatten_parameters(self):
        params = list(self.parameters())
        flat = getattr(self, "_ld_flat", None)
        if flat is not None and flat.device == params[0].device and all(
                q.data_ptr() == flat.data_ptr() + 4 * off for q, off in zip(params, self._ld_flat_offsets)):
            return flat
        flat = torch.cat([q.detach().reshape(-1).float() for q in params]).contiguous()
        offsets, off = [], 0
        for q in params:
            n = q.numel()
            q.data = flat[off:off + n].view_as(q)
            offsets.append(off)
            off += n
        self._ld_flat, self._ld_flat_offsets = flat, offsets
        self._ld_flat_leaves = []
        self._ld_fingerprint = None   # storage moved: re-fold on the next eval forward
        return flat

    def flat_gradient(self):
        """The flat gradient accumulated since the last zero (same layout as flatten_parameters()), or None."""
        grads = [l.grad for l in (getattr(self, "_ld_flat_leaves", None) or []) if l.grad is not None]
        if not grads:
            return None
        return grads[0] if len(grads) == 1 else torch.stack(grads).sum(0)

    def zero_flat_gradient(self):
        self._ld_flat_leaves = []

    def _forward_train(self, x):
        """Training-mode forward on the B200 kernels (batch-statistics BatchNorm, dropout); autograd-connected to
        the module parameters through one flat parameter vector, so loss.backward() fills .grad like the reference."""
        B = x.shape[0]
        eng = self._train_engine(B)
        p = float(self.dropout.p)
        dev = self.conv1.weight.device
        if p > 0.0:
            mask1 = (torch.rand(B, self.linear_layer_size, device=dev) >= p).float()
            mask2 = (torch.rand(B, 32, device=dev) >= p).float()
        else:
            mask1 = torch.ones(B, self.linear_layer_size, device=dev)
            mask2 = torch.ones(B, 32, device=dev)
        self._ld_last_masks = (mask1, mask2)
        feats = x.detach().float().reshape(B, x.shape[2], x.shape[3]).contiguous()
        if getattr(self, "_ld_fused", False):
            # fused-optimiser mode (train.B200Adam): the flat vector is the leaf; its gradient stays flat
            leaf = self.flatten_parameters().detach().requires_grad_(True)   # shares storage; its .grad is the flat gradient
            self._ld_flat_leaves = getattr(self, "_ld_flat_leaves", None) or []
            self._ld_flat_leaves.append(leaf)
            probs, bn_stats = _TrainFunction.apply(leaf, feats, mask1, mask2, p, eng)
        else:
            flat = torch.cat([q.reshape(-1) for q in self.parameters()])
            probs, bn_stats = _TrainFunction.apply(flat, feats, mask1, mask2, p, eng)
        self._update_running_stats(eng, bn_stats, B)
        return probs.reshape(B, 1)

    @torch.no_grad()
    def _update_running_stats(self, eng, bn_stats, batch):
        """nn.BatchNorm's running-statistics update (momentum 0.1, unbiased variance), from the kernel's batch statistics.
        Multi-tensor (foreach) ops: six launches for the 22 BatchNorms instead of ~180."""
        cache = getattr(self, "_bn_update_cache", None)
        if cache is None or cache["batch"] != batch or any(bn.running_mean is not t for (bn, _, _), t in zip(cache["entries"], cache["rm"])):
            # (buffers are replaced, not updated in place, by Module.to()/set_device)
            modules = dict(self.named_modules())
            entries = [(modules[name], off, C) for name, off, C in eng.train_table["batchnorms"]]
            keep, mom, var_scale = [], [], []
            for (bn, _, _), (name, _, _) in zip(entries, eng.train_table["batchnorms"]):
                n = batch * self._bn_pixels(name) if isinstance(bn, nn.BatchNorm2d) else batch
                m = bn.momentum if bn.momentum is not None else 0.1
                keep.append(1.0 - m)
                mom.append(m)
                var_scale.append(m * (n / max(n - 1, 1)))
            cache = {"batch": batch, "entries": entries, "keep": keep, "mom": mom, "var_scale": var_scale,
                     "rm": [bn.running_mean for bn, _, _ in entries], "rv": [bn.running_var for bn, _, _ in entries],
                     "nbt": [bn.num_batches_tracked for bn, _, _ in entries]}
            self._bn_update_cache = cache
        means = [bn_stats[off:off + C] for _, off, C in cache["entries"]]
        variances = [bn_stats[off + C:off + 2 * C] for _, off, C in cache["entries"]]
        torch._foreach_mul_(cache["rm"], cache["keep"])
        torch._foreach_add_(cache["rm"], torch._foreach_mul(means, cache["mom"]))
        torch._foreach_mul_(cache["rv"], cache["keep"])
        torch._foreach_add_(cache["rv"], torch._foreach_mul(variances, cache["var_scale"]))
        torch._foreach_add_(cache["nbt"], 1)

    def _bn_pixels(self, name):
        """Spatial positions per sample seen by a BatchNorm2d of the 100 x 44 input."""
        if name == "bn1":
            return 100 * 44
        b = int(name.split(".")[0][len("block"):])
        return {1: 100 * 44, 2: 50 * 22, 3: 25 * 11, 4: 13 * 6}[b]

    def forward(self, x):
        """x: (B, 1, 100, 44) float -> (B, 1) sigmoid probabilities."""
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (B, 1, T, F), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise LdError("input is not on a CUDA device (no CPU path)")
        if self.training:
            if x.shape[2] != 100 or x.shape[3] != 44:
                raise LdError("ResNetBigger on B200 is built for 100 x 44 windows")
            return self._forward_train(x)
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (B, 1, T, F), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise LdError("input is not on a CUDA device (no CPU path)")
        eng = self.b200_engine()
        B, _, T, Fdim = x.shape
        if T != eng.cfg.num_frames or Fdim != eng.cfg.num_filters:
            raise LdError(f"ResNetBigger on B200 is built for {eng.cfg.num_frames} x {eng.cfg.num_filters} windows")
        # Windows laid back to back form one sequence: the window that starts at row 100*k is item k, its
        # zero padding handled by the window-specific planes, so no item sees its neighbours.
        feats = x.detach().float().reshape(B * T, Fdim)
        probs = eng.infer_windows(feats)
        return probs[::T].reshape(B, 1).clone()

    def infer_channel(self, feats, chan_frames=None):
        """Fast path: probability of the window starting at EVERY frame (InferenceDataset semantics)."""
        return self.b200_engine().infer_windows(feats, chan_frames)
