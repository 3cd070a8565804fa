"""Small workload (not part of the product): one short inference pass through every kernel of the hot path and two training
steps at batch 4 -- the smallest run that touches every launch, for sanitizer-style checks where such tools are available."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laughter_detection_icsi_b200 import models, synth, train as ld_train  # noqa: E402
from laughter_detection_icsi_b200.engine import Engine  # noqa: E402

eng = Engine(0, chunk_rows=256)
eng.load_state_dict(synth.synthetic_state_dict())
rng = np.random.default_rng(0)
pcm = torch.from_numpy((rng.normal(size=16000 * 4) * 3000).astype(np.int16)).cuda()
feats, _ = eng.fbank(pcm)
probs = eng.infer_windows(feats)
runs = eng.segment_runs(probs, [0.3, 0.6])
torch.cuda.synchronize()
print("inference ok:", tuple(feats.shape), float(probs.mean()), [len(r[0]) for r in runs])
eng.close()

dev = torch.device("cuda", 0)
model = models.ResNetBigger(dropout_rate=0.5, linear_layer_size=48, filter_sizes=[64, 32, 16, 16])
model.load_state_dict(synth.synthetic_state_dict(head_gain=1.0, head_bias_shift=0.0))
model.set_device(dev)
opt = ld_train.B200Adam(model)
b = {k: v.pin_memory() for k, v in ld_train.synthetic_lad_batch(4, seed=1).items()}
for _ in range(2):
    out = ld_train.train_batch_fused(model, opt, b, dev)
torch.cuda.synchronize()
print("training ok:", out[0])
