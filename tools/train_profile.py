"""Short training run for ncu launch lists (not part of the product): B=256, a few steps."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laughter_detection_icsi_b200 import models, synth, train as ld_train
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
model = models.ResNetBigger(dropout_rate=0.5, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
model.load_state_dict(synth.synthetic_state_dict(head_gain=1.0, head_bias_shift=0.0))
model.set_device(dev)
opt = ld_train.B200Adam(model)   # the bench configuration: fused clip + Adam on the flat parameter vector
b = {k: v.pin_memory() for k, v in ld_train.synthetic_lad_batch(256, seed=1).items()}
for _ in range(2):
    ld_train.train_batch_fused(model, opt, b, dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    ld_train.train_batch_fused(model, opt, b, dev)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / steps * 1e3:.2f} ms per step (wall)")
