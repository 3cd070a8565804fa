"""K1 alone on 6 x 1 h of synthetic audio (for ncu); prints the CUDA-event time."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laughter_detection_icsi_b200.engine import get_engine
eng = get_engine(0)
n = 16000 * 3600
pcm = (torch.randn(6 * n, device="cuda") * 2000).to(torch.int16)
for _ in range(2):
    eng.fbank(pcm, [n] * 6)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    feats, frames = eng.fbank(pcm, [n] * 6)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"fbank 6 x 1 h: {ms:.3f} ms -> {sum(frames) * 496 / ms / 1e6:.1f} GB/s algorithmic, {sum(frames) / ms / 1e3:.1f} M frames/s")
