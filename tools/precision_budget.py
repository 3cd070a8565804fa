"""Where does the fp16 error of the calibrated-head checkpoint come from?  ("weak" item 2)

CPU experiment on the plan emulator (tests/plan_emulator.py): the streaming plan is executed in fp32 with fp16 rounding
switched on selectively -- weights only, stored activations only, and one block at a time -- and compared with the fp64
oracle on the probabilities of the bench checkpoint (synth.synthetic_state_dict(), head gain ~218).

    python tools/precision_budget.py [n_windows]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from laughter_detection_icsi_b200 import _native, synth  # noqa: E402
from oracle import fbank_oracle, resnet_oracle  # noqa: E402
import plan_emulator  # noqa: E402


class SelectiveEmulator(plan_emulator.PlanEmulator):
    """q_w / q_a: predicates on the conv name ('stem', 'block1.0.conv1', ...) deciding whether that layer's weights /
    stored output are rounded to fp16.  split_w / split_a: the value is kept as hi + lo fp16 pair (22 bits) instead."""

    def __init__(self, plan, sd, q_w, q_a, split_w=lambda n: False, split_a=lambda n: False):
        super().__init__(plan, sd, half=False)
        self.q_w, self.q_a, self.split_w, self.split_a = q_w, q_a, split_w, split_a
        self._layer = "stem"
        self._seen_stem = False

    @staticmethod
    def _round(x, split):
        hi = x.half().float()
        if not split:
            return hi
        return hi + (x - hi).half().float()

    def _q(self, x):
        # PlanEmulator calls _q on: stem outputs (3x), then per conv launch: weights once, then outputs per job
        kind, name = self._next_kind()
        if kind == "w":
            return self._round(x, self.split_w(name)) if self.q_w(name) else x
        return self._round(x, self.split_a(name)) if self.q_a(name) else x

    def run(self, feats, nb):
        order = [("a", "stem")] * len(self.plan["stem"])
        for c in self.plan["convs"]:
            order.append(("w", c["conv"]))
            n_out = len(c["jobs"])
            order += [("a", c["conv"])] * n_out
        self._order = iter(order)
        return super().run(feats, nb)

    def _next_kind(self):
        return next(self._order)


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    torch.set_num_threads(os.cpu_count() or 8)
    sd = synth.synthetic_state_dict()
    pcm = synth.synth_channel(nb * 160 + 16000 + 37, meeting=0, channel=0)
    feats = fbank_oracle.fbank(pcm.numpy().astype(np.float32) / 32768.0)
    ref = resnet_oracle.window_probs(sd, feats.numpy()[: nb + 100], dtype=torch.float64)[:nb] \
        if "dtype" in resnet_oracle.window_probs.__code__.co_varnames else None
    plan = _native.plan_json()
    feats_t = feats[: nb + 100]

    def run(q_w, q_a, **kw):
        emu = SelectiveEmulator(plan, sd, q_w, q_a, **kw)
        return emu.run(feats_t, nb).double().numpy()

    base = run(lambda n: False, lambda n: False)
    if ref is None:
        ref = base  # fp32 emulator as the reference (3e-6 from fp64)
    print(f"windows {nb}; probs {ref.min():.3f}..{ref.max():.3f}; fp32 emulator vs reference: max {np.max(np.abs(base - ref)):.2e}")
    always, never = (lambda n: True), (lambda n: False)

    def blk(prefix):
        return lambda n: n.startswith(prefix)

    rows = [("all fp16 (what the kernels do)", always, always, {}),
            ("fp16 weights only", always, never, {}),
            ("fp16 activations only", never, always, {}),
            ("fp16 stem output only", never, blk("stem"), {})]
    for b in ("block1", "block2", "block3", "block4"):
        rows.append((f"fp16 activations of {b} only", never, blk(b), {}))
        rows.append((f"fp16 weights of {b} only", blk(b), never, {}))
    rows += [("all fp16, weights split hi+lo", always, always, dict(split_w=always)),
             ("all fp16, activations split hi+lo", always, always, dict(split_a=always)),
             ("all fp16, both split", always, always, dict(split_w=always, split_a=always)),
             ("all fp16, block3+4 both split", always, always,
              dict(split_w=lambda n: n.startswith(("block3", "block4")), split_a=lambda n: n.startswith(("block3", "block4")))),
             ("all fp16, block2+3+4 both split", always, always,
              dict(split_w=lambda n: n.startswith(("block2", "block3", "block4")),
                   split_a=lambda n: n.startswith(("block2", "block3", "block4")))),
             ("all fp16, block1 both split", always, always,
              dict(split_w=lambda n: n.startswith(("block1", "stem")), split_a=lambda n: n.startswith(("block1", "stem"))))]
    b234 = ("block2", "block3", "block4")
    is_y = lambda n: n.endswith(("conv2", "shortcut.0"))
    rows += [("all fp16, weights of block2-4 split", always, always, dict(split_w=lambda n: n.startswith(b234))),
             ("all fp16, weights of block4 split", always, always, dict(split_w=lambda n: n.startswith("block4"))),
             ("all fp16, block2-4: w split + y/sc planes split", always, always,
              dict(split_w=lambda n: n.startswith(b234), split_a=lambda n: n.startswith(b234) and is_y(n))),
             ("all fp16, block2-4 split + block1.1.conv2 out split", always, always,
              dict(split_w=lambda n: n.startswith(b234), split_a=lambda n: n.startswith(b234) or n == "block1.1.conv2")),
             ("all fp16, block2-4 split + stem out split", always, always,
              dict(split_w=lambda n: n.startswith(b234), split_a=lambda n: n.startswith(b234) or n == "stem")),
             ("all fp16, block2-4 split + block1 y planes split", always, always,
              dict(split_w=lambda n: n.startswith(b234), split_a=lambda n: n.startswith(b234) or (n.startswith("block1") and is_y(n)))),
             ("all fp16, all w split + block2-4 act split", always, always,
              dict(split_w=always, split_a=lambda n: n.startswith(b234))),
             ]
    if os.environ.get("ONLY_NEW"):
        rows = rows[-7:]
    print(f"{'variant':45s} {'max |dp|':>10s} {'rms |dp|':>10s}")
    for name, qw, qa, kw in rows:
        p = run(qw, qa, **kw)
        d = np.abs(p - ref)
        print(f"{name:45s} {d.max():10.2e} {np.sqrt(np.mean(d * d)):10.2e}")


if __name__ == "__main__":
    main()
