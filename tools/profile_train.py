"""A few training steps at batch 256 (for ncu launch lists): python tools/profile_train.py [steps] [graph 0/1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laughter_detection_icsi_b200 import models, synth, train as ld_train  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
graph = (sys.argv[2] if len(sys.argv) > 2 else "0") != "0"
dev = torch.device("cuda", 0)
model = models.ResNetBigger(dropout_rate=0.5, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
model.load_state_dict(synth.synthetic_state_dict(head_gain=1.0, head_bias_shift=0.0))
model.set_device(dev)
opt = ld_train.B200Adam(model)
stepper = ld_train.make_stepper(model, opt, dev, graph=graph)
batch = {k: v.pin_memory() for k, v in ld_train.synthetic_lad_batch(256, seed=0).items()}
for _ in range(steps):
    stepper(batch)
print(stepper.flush(), stepper.mode)
torch.cuda.synchronize()
