"""Where the end-to-end pipeline call spends its wall time (not part of the product)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laughter_detection_icsi_b200 import synth
from laughter_detection_icsi_b200.pipeline import LaughterPipeline
thr, ml = synth.eval_grid()
pipe = LaughterPipeline(synth.synthetic_state_dict(), device=0, thresholds=thr, min_lengths=ml)
n = 16000 * 3600
pcm_dev, lens = synth.synth_meeting(6, n, device="cuda:0")
host = torch.empty(pcm_dev.shape, dtype=torch.int16, pin_memory=True); host.copy_(pcm_dev); torch.cuda.synchronize()
for _ in range(2):
    pipe(host, lens)
torch.cuda.synchronize()
def t(): torch.cuda.synchronize(); return time.perf_counter()
t0 = t(); inst, frames = pipe(host, lens); t1 = t()
print(f"whole call {1e3 * (t1 - t0):.1f} ms")
t0 = t(); d = host.to("cuda:0", non_blocking=True); t1 = t(); print(f"  H2D alone {1e3 * (t1 - t0):.1f} ms")
t0 = t(); probs, frames = pipe.probabilities(d, lens); t1 = t(); print(f"  probabilities (all channels, one call) {1e3 * (t1 - t0):.1f} ms")
t0 = t(); runs = pipe.runs(probs, frames); t1 = t(); print(f"  runs (K4 + D2H) {1e3 * (t1 - t0):.1f} ms")
t0 = t(); out = pipe.instances(runs, frames, [n / 16000.0] * 6); t1 = t(); print(f"  instances (host) {1e3 * (t1 - t0):.1f} ms")
t0 = t(); x = torch.cat([p for p in probs.split(360000)]); t1 = t(); print(f"  cat {1e3 * (t1 - t0):.2f} ms")
