"""Mirror of the hot part of the reference's train.py: ``_train_batch`` (:261-297), ``_calc_metrics`` (:203-224) and the
per-step optimiser recipe (clip_grad_norm_ 1.0, Adam with PyTorch defaults, :151,292-295,336), running ResNetBigger's
forward/backward on the B200 training kernels (ld_train_*).  Data-parallel training all-reduces ONE flat fp32 gradient
bucket (221 217 elements) per step over NCCL (`distributed.allreduce_gradients`); BatchNorm statistics stay per GPU
like in the reference (no SyncBN).

    python -m laughter_detection_icsi_b200.train --config resnet_base --checkpoint_dir ck --synthetic_steps 50
"""
import argparse
import os
import time

import numpy as np
import torch
from torch import nn, optim

from . import config as config_module
from . import distributed as ld_dist
from . import synth
from .utils import torch_utils


def _calc_metrics(output, trgs):
    """accuracy, precision, recall of round(output) against the targets (train.py:203-224)."""
    preds = torch.round(output)
    acc = torch.sum(preds == trgs).float() / len(trgs)
    corr_pred_laughs = torch.sum(preds * trgs).float()
    total_pred_laughs = torch.sum(preds == 1).float()
    total_trg_laughs = torch.sum(trgs == 1).float()
    prec = corr_pred_laughs / total_pred_laughs if total_pred_laughs > 0 else torch.tensor(1.0)
    recall = corr_pred_laughs / total_trg_laughs if total_trg_laughs > 0 else torch.tensor(1.0)
    return float(acc), float(prec), float(recall)


def train_batch(model, optimizer, batch, device, clip=1.0, gradient_accumulation_steps=1, step=1, world_size=1):
    """One optimisation step like train.py:_train_batch: returns (loss, accuracy, precision, recall)."""
    model.train()
    segs = batch['inputs'][:, None, :, :].to(device, non_blocking=True)
    labs = batch['is_laugh'].float().to(device, non_blocking=True)
    output = model(segs).squeeze()
    criterion = nn.BCELoss()
    loss = criterion(output, labs)
    acc, prec, recall = _calc_metrics(output.detach(), labs)
    (loss / gradient_accumulation_steps).backward()
    if step % gradient_accumulation_steps == 0:
        if world_size > 1:
            ld_dist.allreduce_gradients(model.parameters(), world_size)
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        optimizer.step()
        model.zero_grad()
    return float(loss.detach()), acc, prec, recall


class B200Adam:
    """Fused clip_grad_norm_(max_norm) + Adam (train.py:292-295,336) for a ResNetBigger with flat parameter storage: ONE kernel
    pair (global norm, update) on the flat fp32 vector (ld_clip_adam_step, K8) instead of ~190 small PyTorch launches.
    Same arithmetic as torch.optim.Adam with default hyper-parameters; data parallel: pass world_size > 1 and the flat
    gradient is all-reduced (mean) first."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        self.model, self.lr, self.betas, self.eps, self.max_norm = model, lr, betas, eps, max_norm
        flat = model.flatten_parameters()
        model._ld_fused = True
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        self.grad_norm = torch.zeros(1, device=flat.device)
        self.steps = 0

    def zero_grad(self):
        self.model.zero_flat_gradient()

    def step(self, world_size=1):
        import torch.distributed as dist
        g = self.model.flat_gradient()
        if g is None:
            raise RuntimeError("B200Adam.step() without a backward pass")
        g = g.contiguous()
        if world_size > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g.div_(world_size)
        self.steps += 1
        eng = self.model._train_engine(1)
        eng.clip_adam_step(self.model.flatten_parameters(), g, self.exp_avg, self.exp_avg_sq, self.steps, self.max_norm, self.lr,
                           self.betas, self.eps, self.grad_norm)

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "steps": self.steps}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.steps = int(sd["steps"])


def train_batch_fused(model, optimizer, batch, device, world_size=1, sync_metrics=True):
    """train_batch with the fused optimiser: forward, BCELoss, backward, (all-reduce,) clip + Adam in one kernel pair."""
    model.train()
    segs = batch['inputs'][:, None, :, :].to(device, non_blocking=True)
    labs = batch['is_laugh'].float().to(device, non_blocking=True)
    optimizer.zero_grad()
    output = model(segs).squeeze()
    loss = nn.BCELoss()(output, labs)
    loss.backward()
    optimizer.step(world_size)
    if not sync_metrics:
        return loss.detach(), None, None, None
    acc, prec, recall = _calc_metrics(output.detach(), labs)
    return float(loss.detach()), acc, prec, recall


def synthetic_lad_batch(batch_size, seed, device="cpu"):
    """LadDataset-shaped batch ({'inputs': (B,100,44) float32, 'is_laugh': (B,) int32}) of log-mel-like windows whose
    label is recoverable (laugh windows carry a 5 Hz harmonic modulation), SURVEY.md section 8d config 4."""
    rng = np.random.default_rng(seed)
    t = np.arange(100)[None, :, None] / 100.0
    f = np.arange(44)[None, None, :] / 44.0
    y = (rng.uniform(size=batch_size) < 0.5).astype(np.int32)
    level = rng.uniform(-9.0, -2.0, (batch_size, 1, 1))
    tilt = rng.uniform(-4.0, 4.0, (batch_size, 1, 1))
    x = level + tilt * (f - 0.5) + rng.normal(0.0, 1.0, (batch_size, 100, 44))
    x = x + y[:, None, None] * 3.0 * np.sin(2 * np.pi * (5.0 * t + rng.uniform(0, 1, (batch_size, 1, 1)))) * (f < 0.5)
    x = np.clip(x, -15.9424, 6.0).astype(np.float32)
    return {"inputs": torch.from_numpy(x).to(device), "is_laugh": torch.from_numpy(y).to(device)}


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--config', type=str, required=True)
    p.add_argument('--checkpoint_dir', type=str, required=True)
    p.add_argument('--data_root', type=str, default=None)
    p.add_argument('--lhotse_dir', type=str, default=None)
    p.add_argument('--data_dfs_dir', type=str, default=None)
    p.add_argument('--num_epochs', type=int, default=1)
    p.add_argument('--batch_size', type=int, default=32)
    p.add_argument('--torch_device', type=str, default='cuda')
    p.add_argument('--num_workers', type=int, default=0)
    p.add_argument('--dropout_rate', type=float, default=0.5)
    p.add_argument('--gradient_accumulation_steps', type=int, default=1)
    p.add_argument('--synthetic_steps', type=int, default=0, help='train on synthetic LAD windows for this many steps (no corpus)')
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    cfg = config_module.MODEL_MAP[args.config]
    if not torch.cuda.is_available():
        raise Exception("No CUDA device found: this build trains on B200 only (no CPU path)")
    device = torch.device(args.torch_device)
    model = cfg['model'](dropout_rate=args.dropout_rate, linear_layer_size=cfg['linear_layer_size'], filter_sizes=cfg['filter_sizes'])
    model.set_device(device)
    torch_utils.init_weights(model)   # every parameter ~ N(0, 0.01), as train.py:425
    optimizer = optim.Adam(model.parameters())
    last = os.path.join(args.checkpoint_dir, 'last.pth.tar')
    if os.path.exists(last):
        torch_utils.load_checkpoint(last, model, optimizer)
    if args.synthetic_steps <= 0:
        from . import load_data
        loader = load_data.create_training_dataloader(args.lhotse_dir or args.data_root, 'train', shuffle=True)
    else:
        loader = (synthetic_lad_batch(args.batch_size, s) for s in range(args.synthetic_steps))
    t0 = time.time()
    for batch in loader:
        loss, acc, prec, rec = train_batch(model, optimizer, batch, device, gradient_accumulation_steps=args.gradient_accumulation_steps,
                                           step=model.global_step + 1)
        model.global_step += 1
        if model.global_step % 10 == 0:
            print(f"step {model.global_step}: loss {loss:.4f} acc {acc:.3f} prec {prec:.3f} rec {rec:.3f}")
    print(f"trained {model.global_step} steps in {time.time() - t0:.1f}s")
    state = torch_utils.make_state_dict(model, optimizer, model.epoch, model.global_step, model.best_val_loss)
    torch_utils.save_checkpoint(state, False, args.checkpoint_dir)


if __name__ == '__main__':
    main()
