"""Training step time at batch 256 through the CUDA-graph stepper (bench configuration): python tools/train_time.py [steps] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laughter_detection_icsi_b200 import models, synth, train as ld_train  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
model = models.ResNetBigger(dropout_rate=0.5, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
model.load_state_dict(synth.synthetic_state_dict(head_gain=1.0, head_bias_shift=0.0))
model.set_device(dev)
opt = ld_train.B200Adam(model)
stepper = ld_train.make_stepper(model, opt, dev, graph=True)
batch = {k: v.pin_memory() for k, v in ld_train.synthetic_lad_batch(256, seed=0).items()}
for _ in range(5):
    stepper(batch)
stepper.flush()
torch.cuda.synchronize()
for _ in range(reps):
    t0 = time.perf_counter()
    for _ in range(steps):
        stepper(batch)
    last = stepper.flush()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    print(f"{ms:.3f} ms per step, {256 / ms:.1f} k samples/s, mode {stepper.mode}, last {last}")
