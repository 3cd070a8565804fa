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

    def forward(self, x):
        """x: (B, 1, 100, 44) float -> (B, 1) sigmoid probabilities (eval mode)."""
        if self.training:
            raise LdError("training-mode forward (batch-statistics BatchNorm, dropout, backward) is not part of "
                          "this round's CUDA path; call model.eval() for inference")
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
