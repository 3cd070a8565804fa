"""Mirror of the reference's train.py: the CLI (:68-145), ``run_training_loop`` (:150-167), ``run_epoch`` with
``_train_batch`` (:261-297), ``_eval_batch`` (:226-259), ``_calc_metrics`` (:203-224), ``_eval_for_logging`` (:178-201), the
logging / checkpoint cadence (:363-412), ``train_params.csv`` (:314-322) and ``metrics.csv`` (:488-504) -- with
ResNetBigger's forward/backward on the B200 training kernels (ld_train_*) and the optimiser recipe (clip_grad_norm_ 1.0,
Adam with PyTorch defaults, re-created every epoch, :151,292-295,336) fused into one kernel pair (K8).

Data parallel (BASELINE config 5): one process per GPU under torchrun; rank r takes batches r, r + W, ... of the training
loader, the gradients travel as ONE flat fp32 bucket (221 217 elements) per step over NCCL, BatchNorm statistics stay per
GPU like in the reference (no SyncBN); rank 0 evaluates, logs and writes checkpoints.

    python -m laughter_detection_icsi_b200.train --config resnet_base --checkpoint_dir ck --data_root <root> [--num_epochs N]
    python -m laughter_detection_icsi_b200.train --config resnet_base --checkpoint_dir ck --synthetic_steps 50
    torchrun --nproc-per-node 8 -m laughter_detection_icsi_b200.train --config resnet_base --checkpoint_dir ck --data_root <root>
"""
import argparse
import csv
import os
import time
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch
from torch import nn, optim

from . import config as config_module
from . import distributed as ld_dist
from . import synth
from .utils import torch_utils


@dataclass
class MetricEntry:
    accuracy: float
    precision: float
    recall: float
    loss: float
    epoch: int

    def to_list(self):
        """[precision, recall, accuracy, loss] -- the column order of metrics.csv (train.py:37-42)."""
        return [self.precision, self.recall, self.accuracy, self.loss]


METRICS_COLUMNS = ['batch_num', 'epoch', 'train_prec', 'train_rec', 'train_acc', 'train_loss', 'val_prec', 'val_rec', 'val_acc', 'val_loss']
TRAIN_PARAMS_COLUMNS = ['train_samples', 'val_samples', 'val_samples_per_log', 'log_freq', 'batchsize']


def _calc_metrics(output, trgs):
    """accuracy, precision, recall of round(output) against the targets (train.py:203-224)."""
    preds = torch.round(output)
    acc = torch.sum(preds == trgs).float() / len(trgs)
    corr_pred_laughs = torch.sum(preds * trgs).float()
    total_pred_laughs = torch.sum(preds == 1).float()
    total_trg_laughs = torch.sum(trgs == 1).float()
    prec = corr_pred_laughs / total_pred_laughs if total_pred_laughs > 0 else torch.tensor(1.0)
    recall = corr_pred_laughs / total_trg_laughs   # 0/0 = NaN without positive targets, as in the reference (train.py:218-222)
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


def eval_batch(model, batch, device, return_raw=False):
    """train.py:_eval_batch (:226-259): eval-mode forward under no_grad, BCELoss, round(output)."""
    with torch.no_grad():
        src = torch.as_tensor(np.asarray(batch['inputs'])).float().to(device)[:, None, :, :]
        trgs = torch.as_tensor(np.asarray(batch['is_laugh'])).float().to(device)
        output = model(src).squeeze(-1).reshape(-1)
        bce_loss = nn.BCELoss()(output, trgs)
        preds = torch.round(output)
        if return_raw:
            return bce_loss.item(), trgs, preds
        acc, prec, recall = _calc_metrics(output, trgs)
        return bce_loss.item(), acc, prec, recall


class B200Adam:
    """Fused clip_grad_norm_(max_norm) + Adam (train.py:292-295,336) for a ResNetBigger with flat parameter storage: ONE kernel
    pair (global norm, update) on the flat fp32 vector (ld_clip_adam_step_dev, K8) instead of ~190 small PyTorch launches.
    Same arithmetic as torch.optim.Adam with default hyper-parameters; data parallel: pass world_size > 1 and the flat
    gradient is all-reduced (mean) first.  The step count lives in device memory, so a captured CUDA graph of the training
    step replays with the right bias corrections."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        self.model, self.lr, self.betas, self.eps, self.max_norm = model, lr, betas, eps, max_norm
        flat = model.flatten_parameters()
        model._ld_fused = True
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        self.grad_norm = torch.zeros(1, device=flat.device)
        self.step_d = torch.zeros(1, dtype=torch.int64, device=flat.device)

    @property
    def steps(self):
        return int(self.step_d.item())

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
        eng = self.model._train_engine(1)
        eng.clip_adam_step_dev(self.model.flatten_parameters(), g, self.exp_avg, self.exp_avg_sq, self.step_d, self.max_norm, self.lr,
                               self.betas, self.eps, self.grad_norm)
        self.model.mark_weights_dirty()   # the kernel wrote the parameters through raw pointers (no _version bump)

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "steps": self.steps}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_d.fill_(int(sd["steps"]))


def train_batch_fused(model, optimizer, batch, device, world_size=1, sync_metrics=True, gradient_accumulation_steps=1, step=1):
    """train_batch with the fused optimiser: forward, BCELoss, backward, (all-reduce,) clip + Adam in one kernel pair, then the
    flat gradient is dropped (model.zero_grad() of train.py:295)."""
    model.train()
    segs = batch['inputs'][:, None, :, :].to(device, non_blocking=True)
    labs = batch['is_laugh'].float().to(device, non_blocking=True)
    output = model(segs).squeeze()
    loss = nn.BCELoss()(output, labs)
    (loss / gradient_accumulation_steps).backward()
    if step % gradient_accumulation_steps == 0:
        optimizer.step(world_size)
        optimizer.zero_grad()
    if not sync_metrics:
        return loss.detach(), None, None, None
    acc, prec, recall = _calc_metrics(output.detach(), labs)
    return float(loss.detach()), acc, prec, recall


def _metrics_vector(loss, output, labs):
    """[loss, #correct, #correct laughs, #predicted laughs, #target laughs] on the device (no synchronisation)."""
    preds = torch.round(output)
    return torch.stack([loss.detach().float(), torch.sum(preds == labs).float(), torch.sum(preds * labs).float(),
                        torch.sum(preds == 1).float(), torch.sum(labs == 1).float()])


def _decode_metrics(v, n):
    loss, correct, corr_laughs, n_pred, n_trg = (float(x) for x in v)
    prec = corr_laughs / n_pred if n_pred > 0 else 1.0
    recall = corr_laughs / n_trg if n_trg > 0 else float('nan')   # 0/0 in the reference (train.py:222)
    return loss, correct / n, prec, recall


class _Stepper:
    """Callable training step for a fixed batch shape (bench.py): forward, BCELoss, backward, (all-reduce,) fused clip + Adam and
    the per-step loss / accuracy / precision / recall read-back of train.py:297.

    The read-back is pipelined by one step -- step k's numbers are fetched while step k + 1 runs -- so the host never drains the
    stream; `__call__` returns the metrics of the PREVIOUS step (None on the first call) and `flush()` the last one.
    With `graph=True` the whole device side of a step (dropout masks, kernels of ld_train_forward / ld_train_backward, loss,
    running-statistics update, NCCL all-reduce, optimiser) is captured once into a CUDA graph and replayed; the batch is copied
    into static buffers first.  Falls back to eager launches when the capture fails (`mode` says which)."""

    def __init__(self, model, optimizer, device, world_size=1, graph=True):
        self.model, self.optimizer, self.device, self.world_size = model, optimizer, device, world_size
        self.want_graph, self.graph = graph, None
        self.mode = "eager, metrics read back one step late"
        self.pending = None
        self.x = self.y = None
        self.launches_per_step = None

    def _body(self):
        model = self.model
        model.train()
        output = model(self.x[:, None, :, :]).squeeze()
        labs = self.y.float()
        loss = nn.BCELoss()(output, labs)
        loss.backward()
        self.optimizer.step(self.world_size)
        self.optimizer.zero_grad()
        self.metrics_dev.copy_(_metrics_vector(loss, output.detach(), labs))

    def _setup(self, batch):
        B = batch['inputs'].shape[0]
        self.x = torch.empty((B, 100, 44), dtype=torch.float32, device=self.device)
        self.y = torch.empty(B, dtype=batch['is_laugh'].dtype, device=self.device)
        self.metrics_dev = torch.zeros(5, device=self.device)
        self.host = [torch.zeros(5).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        if not self.want_graph:
            return
        try:
            eng = self.model._train_engine(B)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):   # eager warm-up on a side stream: first-use allocations, lazy module state
                for _ in range(3):
                    self.x.copy_(batch['inputs'], non_blocking=True); self.y.copy_(batch['is_laugh'], non_blocking=True)
                    l0 = eng.train_kernel_launches
                    self._body()
                    self.launches_per_step = eng.train_kernel_launches - l0
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
            self.mode = "CUDA graph replay of the whole step, metrics read back one step late"
        except Exception as e:   # noqa: BLE001 -- any capture problem: run eagerly, say so
            self.graph = None
            self.model.zero_flat_gradient()
            torch.cuda.synchronize(self.device)
            self.mode = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]}), metrics read back one step late"

    def __call__(self, batch):
        if self.x is None or self.x.shape[0] != batch['inputs'].shape[0]:
            self.flush()
            self._setup(batch)
        self.x.copy_(batch['inputs'], non_blocking=True)
        self.y.copy_(batch['is_laugh'], non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            eng = self.model._train_engine(self.x.shape[0])
            l0 = eng.train_kernel_launches
            self._body()
            self.launches_per_step = eng.train_kernel_launches - l0
        slot = self.k & 1
        self.host[slot].copy_(self.metrics_dev, non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        prev, self.pending = self.pending, slot
        self.k += 1
        if prev is None:
            return None
        self.events[prev].synchronize()
        return _decode_metrics(self.host[prev].tolist(), self.x.shape[0])

    def flush(self):
        if self.pending is None:
            return None
        self.events[self.pending].synchronize()
        out = _decode_metrics(self.host[self.pending].tolist(), self.x.shape[0])
        self.pending = None
        return out


def make_stepper(model, optimizer, device, world_size=1, graph=True):
    return _Stepper(model, optimizer, device, world_size, graph=graph)


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


class SyntheticLoader:
    """Re-iterable stand-in for create_training_dataloader: `n_batches` synthetic LAD batches, `num_cuts` like a sampler."""

    def __init__(self, n_batches, batch_size, seed0=0):
        self.n_batches, self.batch_size, self.seed0 = n_batches, batch_size, seed0
        self.num_cuts = n_batches * batch_size

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        return (synthetic_lad_batch(self.batch_size, self.seed0 + s) for s in range(self.n_batches))


def num_cuts(loader):
    """`iterator.sampler.num_cuts` of the reference (train.py:305-306), for lhotse loaders and for the lhotse-free lists."""
    sampler = getattr(loader, "sampler", None)
    if sampler is not None and hasattr(sampler, "num_cuts"):
        return int(sampler.num_cuts)
    if hasattr(loader, "num_cuts"):
        return int(loader.num_cuts)
    return int(sum(len(b['is_laugh']) for b in loader))


class Trainer:
    """run_training_loop / run_epoch of the reference around one model.  `fused=True` uses K8 (B200Adam) on a ResNetBigger with
    flat parameter storage; `fused=False` is the literal reference recipe (optim.Adam + clip_grad_norm_) and also what the
    CPU tests drive with a stand-in module."""

    def __init__(self, model, device, checkpoint_dir, batch_size=32, log_frequency=900, gradient_accumulation_steps=1, clip=1.0,
                 world_size=1, rank=0, fused=True, verbose=True):
        self.model, self.device, self.checkpoint_dir = model, device, checkpoint_dir
        self.batch_size, self.log_frequency, self.accum, self.clip = batch_size, log_frequency, gradient_accumulation_steps, clip
        self.world_size, self.rank, self.fused, self.verbose = world_size, rank, fused, verbose
        self.metrics = {}   # METRICS_DICT: global step -> {'train': MetricEntry, 'val': MetricEntry}
        self.metrics_file = os.path.join(checkpoint_dir, 'metrics.csv')
        self.train_params_file = os.path.join(checkpoint_dir, 'train_params.csv')
        self.optimizer = None
        if rank == 0:
            Path(checkpoint_dir).mkdir(parents=True, exist_ok=True)

    def _new_optimizer(self):
        # run_epoch re-creates Adam with PyTorch defaults at the start of every epoch (train.py:336): its state is discarded
        return B200Adam(self.model, max_norm=self.clip) if self.fused else optim.Adam(self.model.parameters())

    def _train_batch(self, batch):
        # the reference tests `model.global_step % accum == 0` BEFORE incrementing global_step (train.py:291)
        step = self.model.global_step + self.accum   # == 0 mod accum exactly when global_step is
        if self.fused:
            return train_batch_fused(self.model, self.optimizer, batch, self.device, world_size=self.world_size,
                                     gradient_accumulation_steps=self.accum, step=step)
        return train_batch(self.model, self.optimizer, batch, self.device, clip=self.clip, gradient_accumulation_steps=self.accum,
                           step=step, world_size=self.world_size)

    def eval_for_logging(self, val_itr, val_iterator, val_batches_per_log):
        """train.py:_eval_for_logging (:178-201): `val_batches_per_log` dev batches in eval mode, metrics over their union."""
        self.model.eval()
        val_losses, val_trgs, val_preds = [], [], []
        for _ in range(val_batches_per_log):
            try:
                val_batch = next(val_itr)
            except StopIteration:
                val_itr = iter(val_iterator)
                val_batch = next(val_itr)
            val_loss, trgs, preds = eval_batch(self.model, val_batch, self.device, return_raw=True)
            val_trgs.append(trgs); val_preds.append(preds); val_losses.append(val_loss)
        if val_trgs:
            trgs, preds = torch.cat(val_trgs), torch.cat(val_preds)
            acc = float(torch.sum(preds == trgs).float() / len(trgs))
            corr = torch.sum((preds == trgs) * (preds == 1)).float()
            n_pred, n_trg = torch.sum(preds == 1).float(), torch.sum(trgs == 1).float()
            prec = float(corr / n_pred) if n_pred > 0 else 1.0
            recall = float(corr / n_trg)
        else:
            acc, prec, recall = float('nan'), 1.0, float('nan')
        self.model.train()
        return val_itr, (float(np.mean(val_losses)) if val_losses else float('nan')), acc, prec, recall

    def run_epoch(self, iterator, epoch_num, val_iterator=None):
        model = self.model
        validate_online = val_iterator is not None and self.log_frequency is not None
        val_itr, val_batches_per_log = None, 0
        if validate_online:
            n_train, n_val = num_cuts(iterator), num_cuts(val_iterator)
            validations_per_epoch = n_train / (self.batch_size * self.log_frequency)
            val_batches_per_log = int(n_val / validations_per_epoch)
            if self.rank == 0:
                if self.verbose:
                    print(f'Training sampler has {n_train} cuts.')
                    print(f'Validation sampler has {n_val} cuts.')
                    print(f'Using batchsize {self.batch_size}.')
                    print(f'Logging every {self.log_frequency} batches.')
                    print(f'Evaluting {val_batches_per_log} batches per log.')
                with open(self.train_params_file, 'w', newline='') as f:
                    w = csv.writer(f)
                    w.writerow(TRAIN_PARAMS_COLUMNS)
                    w.writerow([n_train, n_val, val_batches_per_log, self.log_frequency, self.batch_size])
            val_itr = iter(val_iterator)
        model.train()
        self.optimizer = self._new_optimizer()
        epoch_loss, num_batches = 0.0, 0
        losses, accs, precs, recalls = [], [], [], []
        is_best = False
        batches = list(iterator) if self.world_size > 1 and not hasattr(iterator, "__len__") else iterator
        n_usable = (len(batches) // self.world_size) * self.world_size if self.world_size > 1 else None
        for i, batch in enumerate(batches):
            if self.world_size > 1:
                if i >= n_usable:
                    break   # every rank takes the same number of steps (the all-reduce is collective)
                if i % self.world_size != self.rank:
                    continue
            batch_loss, batch_acc, batch_prec, batch_recall = self._train_batch(batch)
            epoch_loss += batch_loss
            model.global_step += 1
            num_batches = +1   # sic (train.py:357): the reference's epoch loss is the SUM of the batch losses
            losses.append(batch_loss); accs.append(batch_acc); precs.append(batch_prec); recalls.append(batch_recall)
            if self.log_frequency is not None and (model.global_step + 1) % self.log_frequency == 0:
                if self.rank == 0 and validate_online:
                    val_itr, val_loss, val_acc, val_prec, val_recall = self.eval_for_logging(val_itr, val_iterator, val_batches_per_log)
                    is_best = val_loss < model.best_val_loss
                    if is_best:
                        model.best_val_loss = val_loss
                    val_metrics = MetricEntry(accuracy=val_acc, precision=val_prec, recall=val_recall, loss=val_loss, epoch=epoch_num)
                    with np.errstate(all='ignore'):
                        train_metrics = MetricEntry(accuracy=float(np.mean(accs)), precision=float(np.mean(precs)),
                                                    recall=float(np.nanmean(recalls)) if not np.all(np.isnan(recalls)) else float('nan'),
                                                    loss=float(np.mean(losses)), epoch=epoch_num)
                    self.metrics[model.global_step] = {'val': val_metrics, 'train': train_metrics}
                    if self.verbose:
                        print("\nLogging at step: ", model.global_step)
                        print("Train metrics: ", train_metrics)
                        print("Validation metrics: ", val_metrics)
                losses, accs, precs, recalls = [], [], [], []
                # checkpoint_frequency == log_frequency in the reference (train.py:160)
                if self.rank == 0:
                    opt_for_ckpt = self.optimizer if not self.fused else None
                    state = torch_utils.make_state_dict(model, opt_for_ckpt, model.epoch, model.global_step, model.best_val_loss)
                    if self.fused:
                        state["optim_dict"] = self.optimizer.state_dict()
                    torch_utils.save_checkpoint(state, is_best=is_best, checkpoint=self.checkpoint_dir)
        model.epoch += 1
        return epoch_loss / num_batches if num_batches else 0.0

    def run_training_loop(self, n_epochs, iterator, val_iterator=None):
        for epoch in range(n_epochs):
            start_time = time.time()
            self.run_epoch(iterator, epoch_num=epoch + 1, val_iterator=val_iterator)
            if self.verbose and self.rank == 0:
                secs = int(time.time() - start_time)
                print(f'Epoch: {epoch + 1:02} | Time: {secs // 60}m {secs % 60}s')

    def update_metrics_on_disk(self):
        """metrics.csv: one row per logged step, appended to an existing file (train.py:488-504)."""
        if self.rank != 0:
            return
        rows = []
        if os.path.isfile(self.metrics_file):
            with open(self.metrics_file, newline='') as f:
                r = list(csv.reader(f))
            rows = r[1:] if r else []
        for batch_num, entry in self.metrics.items():
            rows.append([batch_num, entry['train'].epoch] + entry['train'].to_list() + entry['val'].to_list())
        with open(self.metrics_file, 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(METRICS_COLUMNS)
            w.writerows(rows)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--config', type=str, required=True)
    p.add_argument('--checkpoint_dir', type=str, required=True)
    p.add_argument('--data_root', type=str, default=None, help='required unless --synthetic_steps is given')
    p.add_argument('--num_epochs', type=int, default=1)
    p.add_argument('--lhotse_dir', type=str, default='lhotse')
    p.add_argument('--data_dfs_dir', type=str, default='data_dfs')
    p.add_argument('--batch_size', type=str)
    p.add_argument('--torch_device', type=str, default='cuda')
    p.add_argument('--num_workers', type=str, default='8')
    p.add_argument('--dropout_rate', type=str, default='0.5')
    p.add_argument('--gradient_accumulation_steps', type=str, default='1')
    p.add_argument('--include_words', type=str, default=None)
    p.add_argument('--train_on_noisy_audioset', type=str, default=None)
    p.add_argument('--synthetic_steps', type=int, default=0, help='train on this many synthetic LAD batches per epoch (no corpus)')
    p.add_argument('--log_frequency', type=int, default=None, help='overrides config[log_frequency] (900 for resnet_base)')
    return p


def init_distributed():
    """(world_size, rank, local_rank); initialises torch.distributed when launched by torchrun (NCCL on GPUs, gloo otherwise)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 1, 0, 0
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")
    return world, rank, local_rank


def main(argv=None, model_factory=None):
    """The reference's module body (train.py:119-145, 414-433, 506-537).  `model_factory(config, dropout_rate)` lets the CPU
    tests drive the loop with a stand-in module; the product path builds config['model'] = ResNetBigger on a B200."""
    args = build_parser().parse_args(argv)
    cfg = config_module.MODEL_MAP[args.config]
    batch_size = int(args.batch_size or cfg['batch_size'])
    log_frequency = args.log_frequency if args.log_frequency is not None else cfg['log_frequency']
    dropout_rate = float(args.dropout_rate)
    accum = int(args.gradient_accumulation_steps)
    world, rank, local_rank = init_distributed()
    if model_factory is None:
        if not torch.cuda.is_available():
            raise Exception("No CUDA device found: this build trains on B200 only (no CPU path)")
        device = torch.device('cuda', local_rank) if args.torch_device == 'cuda' else torch.device(args.torch_device)
        print("Initializing model...")
        print("Using device", device)
        model = cfg['model'](dropout_rate=dropout_rate, linear_layer_size=cfg['linear_layer_size'], filter_sizes=cfg['filter_sizes'])
        model.set_device(device)
        torch_utils.count_parameters(model)
        model.apply(torch_utils.init_weights)   # every parameter ~ N(0, 0.01), as train.py:425
        fused = True
    else:
        device = torch.device('cpu')
        model = model_factory(cfg, dropout_rate)
        fused = False
    if world > 1:
        # data parallel: every rank starts from rank 0's initial parameters
        import torch.distributed as dist
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
        if hasattr(model, "mark_weights_dirty"):
            model.mark_weights_dirty()
    optimizer = optim.Adam(model.parameters())
    last = os.path.join(args.checkpoint_dir, 'last.pth.tar')
    if os.path.exists(args.checkpoint_dir) and os.path.isfile(last):
        ck = torch.load(last, map_location='cpu', weights_only=False)
        # (the optimiser state is discarded by run_epoch anyway, train.py:336; fused checkpoints store B200Adam's state)
        torch_utils.load_checkpoint(last, model, optimizer if isinstance(ck.get("optim_dict"), dict) and "param_groups" in ck["optim_dict"] else None)
    elif rank == 0:
        print("Saving checkpoints to ", args.checkpoint_dir)
        print("Beginning training...")
    if args.synthetic_steps > 0:
        train_loader = SyntheticLoader(args.synthetic_steps, batch_size, seed0=0)
        dev_loader = SyntheticLoader(max(1, args.synthetic_steps // 4), batch_size, seed0=10 ** 6)
    else:
        if args.data_root is None:
            raise SystemExit("--data_root is required (or --synthetic_steps N)")
        from . import load_data
        if rank == 0:
            print("Preparing training set...")
        cutset_dir = os.path.join(args.data_root, args.lhotse_dir, 'cutsets')
        # Shuffle dev set such that evaluated cuts aren't always the same when the script is called (train.py:509-510)
        dev_loader = load_data.create_training_dataloader(cutset_dir, 'dev', shuffle=True)
        train_loader = load_data.create_training_dataloader(cutset_dir, 'train')
    trainer = Trainer(model, device, args.checkpoint_dir, batch_size=batch_size, log_frequency=log_frequency,
                      gradient_accumulation_steps=accum, world_size=world, rank=rank, fused=fused)
    start_time = time.time()
    trainer.run_training_loop(args.num_epochs, train_loader, val_iterator=dev_loader)
    tot = time.time() - start_time
    if rank == 0:
        print(f"Ran {args.num_epochs} epochs.")
        print(f"Total training time[in three different formats s/min/h]:\n{tot:.2f}s\n{tot / 60:.2f}m\n{tot / 3600:.2f}h")
        print('---------------')
        per = tot / max(args.num_epochs, 1)
        print(f"Time per epoch time[in three different formats s/min/h]:\n{per:.2f}s\n{per / 60:.2f}m\n{per / 3600:.2f}h")
    trainer.update_metrics_on_disk()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return trainer


if __name__ == '__main__':
    main()
