"""Host logic of the train.py mirror (reference train.py:150-167 epochs loop, :178-201 dev-set evaluation, :363-412 logging and
checkpoint cadence, :314-322 train_params.csv, :488-504 metrics.csv) driven on the CPU with a stand-in module -- the product
model only runs on a B200 -- and the data-parallel path under a world-size-2 gloo group."""
import csv
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from laughter_detection_icsi_b200 import train as ld_train


class TinyNet(nn.Module):
    """Same surface as ResNetBigger as far as train.py is concerned: (B,1,100,44) -> (B,1) in (0,1), step counters."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(1, 4, 3, padding=1)
        self.bn = nn.BatchNorm2d(4)
        self.linear = nn.Linear(4, 1)
        self.global_step, self.epoch, self.best_val_loss = 0, 0, np.inf

    def forward(self, x):
        h = torch.relu(self.bn(self.conv(x))).mean((2, 3))
        return torch.sigmoid(self.linear(h))

    def set_device(self, device):
        self.to(device)


def factory(cfg, dropout_rate):
    torch.manual_seed(1234)
    return TinyNet()


def read_csv(path):
    with open(path, newline='') as f:
        return list(csv.reader(f))


def test_cli_epochs_logging_checkpoints_and_csv(tmp_path):
    ck = str(tmp_path / "ck")
    argv = ["--config", "resnet_base", "--checkpoint_dir", ck, "--synthetic_steps", "12", "--batch_size", "8", "--num_epochs", "2",
            "--log_frequency", "4"]
    tr = ld_train.main(argv, model_factory=factory)
    m = tr.model
    assert m.epoch == 2 and m.global_step == 24
    # logging when (global_step + 1) % log_frequency == 0, after the increment (train.py:355,363)
    assert sorted(tr.metrics) == [3, 7, 11, 15, 19, 23]
    rows = read_csv(os.path.join(ck, "metrics.csv"))
    assert rows[0] == ['batch_num', 'epoch', 'train_prec', 'train_rec', 'train_acc', 'train_loss', 'val_prec', 'val_rec', 'val_acc', 'val_loss']
    assert [int(r[0]) for r in rows[1:]] == [3, 7, 11, 15, 19, 23] and [int(r[1]) for r in rows[1:]] == [1, 1, 1, 2, 2, 2]
    assert all(np.isfinite(float(r[5])) and np.isfinite(float(r[9])) for r in rows[1:])
    params = read_csv(os.path.join(ck, "train_params.csv"))
    assert params[0] == ['train_samples', 'val_samples', 'val_samples_per_log', 'log_freq', 'batchsize']
    # 96 train cuts, 24 dev cuts: validations per epoch = 96 / (8 * 4) = 3 -> 8 dev cuts... int(24 / 3) = 8 batches per log
    assert [int(x) for x in params[1]] == [96, 24, 8, 4, 8]
    last = torch.load(os.path.join(ck, "last.pth.tar"), weights_only=False)
    assert set(last) == {"epoch", "global_step", "best_val_loss", "state_dict", "optim_dict"} and last["global_step"] == 23
    assert os.path.isfile(os.path.join(ck, "best.pth.tar"))
    best = torch.load(os.path.join(ck, "best.pth.tar"), weights_only=False)
    assert best["best_val_loss"] == min(e['val'].loss for e in tr.metrics.values())
    # resume: last.pth.tar is loaded (global_step + 1, torch_utils.load_checkpoint), metrics.csv is appended to
    tr2 = ld_train.main(argv[:-4] + ["--num_epochs", "1", "--log_frequency", "4"], model_factory=factory)
    assert tr2.model.global_step == 24 + 12 and tr2.model.epoch == last["epoch"] + 1
    rows2 = read_csv(os.path.join(ck, "metrics.csv"))
    assert len(rows2) == 1 + 6 + 3 and rows2[:7] == rows


def test_training_reduces_the_loss_on_separable_synthetic_batches(tmp_path):
    tr = ld_train.main(["--config", "resnet_base", "--checkpoint_dir", str(tmp_path / "ck"), "--synthetic_steps", "40", "--batch_size", "16",
                        "--num_epochs", "3", "--log_frequency", "20"], model_factory=factory)
    losses = [tr.metrics[k]['train'].loss for k in sorted(tr.metrics)]
    assert losses[-1] < losses[0]


def test_metric_entry_and_recall_nan():
    e = ld_train.MetricEntry(accuracy=0.5, precision=0.25, recall=0.75, loss=1.5, epoch=2)
    assert e.to_list() == [0.25, 0.75, 0.5, 1.5]
    acc, prec, rec = ld_train._calc_metrics(torch.tensor([0.2, 0.1]), torch.tensor([0.0, 0.0]))
    assert acc == 1.0 and prec == 1.0 and np.isnan(rec)   # no predicted laughs -> precision 1; no target laughs -> 0/0 (train.py:213-220)


def _dp_worker(rank, world, port, ck):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    tr = ld_train.main(["--config", "resnet_base", "--checkpoint_dir", ck, "--synthetic_steps", "12", "--batch_size", "8", "--num_epochs", "1",
                        "--log_frequency", "3"], model_factory=factory)
    flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"steps": tr.model.global_step, "same": bool(torch.equal(gathered[0], gathered[1])), "flat": flat}, os.path.join(ck, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_cli_over_gloo(tmp_path):
    ck = str(tmp_path / "ck")
    mp.spawn(_dp_worker, args=(2, 29231 + os.getpid() % 500, ck), nprocs=2, join=True)
    r = torch.load(os.path.join(ck, "dp.pt"))
    assert r["steps"] == 6            # 12 batches sharded over 2 ranks
    assert r["same"]                   # identical parameters on both ranks after averaged-gradient updates
    init = torch.cat([p.detach().reshape(-1) for p in factory(None, 0.5).parameters()])
    assert not torch.equal(init, r["flat"])
    rows = read_csv(os.path.join(ck, "metrics.csv"))
    assert [int(x[0]) for x in rows[1:]] == [2, 5]   # rank 0 logs at (global_step + 1) % 3 == 0
    assert os.path.isfile(os.path.join(ck, "last.pth.tar"))
