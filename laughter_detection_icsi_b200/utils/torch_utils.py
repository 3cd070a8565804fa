"""Mirror of the reference's utils/torch_utils.py: checkpoint format (:36-93) and initialisation (:17-24)."""
import os
import shutil

import torch
from torch import nn


def count_parameters(model):
    counts = sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f"The model has {counts:,} trainable parameters")


def init_weights(model):
    # every parameter, BatchNorm affine and biases included, ~ N(0, 0.01)
    for _, param in model.named_parameters():
        nn.init.normal_(param.data, mean=0, std=0.01)
    if hasattr(model, "mark_weights_dirty"):
        model.mark_weights_dirty()   # writes through .data do not bump tensor._version (models.ResNetBigger.b200_engine)


def save_checkpoint(state, is_best, checkpoint):
    """Writes <checkpoint>/last.pth.tar and, if is_best, copies it to best.pth.tar."""
    filepath = os.path.join(checkpoint, "last.pth.tar")
    if not os.path.exists(checkpoint):
        print("Checkpoint Directory does not exist! Making directory {}".format(checkpoint))
        os.mkdir(checkpoint)
    torch.save(state, filepath)
    if is_best:
        shutil.copyfile(filepath, os.path.join(checkpoint, "best.pth.tar"))


def load_checkpoint(checkpoint, model, optimizer=None):
    if not os.path.exists(checkpoint):
        raise FileNotFoundError("File doesn't exist {}".format(checkpoint))
    print("Loading checkpoint at:", checkpoint)
    checkpoint = torch.load(checkpoint, map_location=lambda storage, loc: storage, weights_only=False)
    model.load_state_dict(checkpoint["state_dict"])
    if optimizer:
        optimizer.load_state_dict(checkpoint["optim_dict"])
    if "epoch" in checkpoint:
        model.epoch = checkpoint["epoch"]
    if "global_step" in checkpoint:
        model.global_step = checkpoint["global_step"] + 1
        print("Loading checkpoint at step: ", model.global_step)
    if "best_val_loss" in checkpoint:
        model.best_val_loss = checkpoint["best_val_loss"]
    return checkpoint


def make_state_dict(model, optimizer=None, epoch=None, global_step=None, best_val_loss=None):
    return {"epoch": epoch, "global_step": global_step, "best_val_loss": best_val_loss,
            "state_dict": model.state_dict(), "optim_dict": optimizer.state_dict() if optimizer is not None else None}
