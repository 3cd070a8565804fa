"""K1 alone on one channel-hour (for ncu): python tools/profile_fbank.py [n_calls]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laughter_detection_icsi_b200 import synth  # noqa: E402
from laughter_detection_icsi_b200.engine import get_engine  # noqa: E402

eng = get_engine(0, chunk_rows=256)
pcm = synth.synth_channel(3600 * 16000, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    feats, frames = eng.fbank(pcm)
torch.cuda.synchronize()
print(frames, float(feats.sum()))
