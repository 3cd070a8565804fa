"""Bring-up diagnostics for the training path on a B200 (not part of the product): determinism and parity of
ld_train_forward / ld_train_backward against the CPU oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from laughter_detection_icsi_b200.engine import get_engine  # noqa: E402
from oracle import resnet_oracle  # noqa: E402
from test_gpu_train import flat_params, make_case  # noqa: E402

eng = get_engine(0, filter_sizes=(64, 32, 16, 16), linear_layer_size=48)
eng.train_create(64)
for seed, B, p in ((5, 16, 0.5), (6, 33, 0.0), (7, 64, 0.5)):
    sd, x, labels, mask1, mask2 = make_case(seed, B)
    if p == 0.0:
        mask1, mask2 = torch.ones_like(mask1), torch.ones_like(mask2)
    ref_probs, ref_loss, ref_grads, ref_stats = resnet_oracle.train_step_reference(sd, x, labels, mask1, mask2, p)
    flat = flat_params(eng, sd)
    xs = x.reshape(B, 100, 44).cuda().contiguous()
    outs, gs, sums = [], [], []
    for rep in range(3):
        probs, bn_stats = eng.train_forward(flat, xs, mask1.cuda(), mask2.cuda(), p)
        pr = probs.detach().clone().requires_grad_(True)
        torch.nn.functional.binary_cross_entropy(pr, labels.cuda()).backward()
        g = eng.train_backward(pr.grad)
        outs.append(probs.cpu().numpy()); gs.append(g.cpu().numpy())
        import ctypes
        buf = (ctypes.c_double * 256)()
        k = eng.lib.ld_train_debug_checksums(eng._h, buf, 256)
        sums.append([buf[i] for i in range(k)])
    print(f"seed {seed} B {B} p {p}: probs err vs oracle {[float(np.abs(o - ref_probs.numpy()).max()) for o in outs]}  "
          f"run-to-run {float(np.abs(outs[0] - outs[1]).max()):.2e} {float(np.abs(outs[0] - outs[2]).max()):.2e}")
    a, b = np.array(sums[0]), np.array(sums[1])
    nconv, nlev = 20, 17
    labels_ = [f"z{i}" for i in range(nconv)] + [f"y{i}" for i in range(nlev)] + [f"dz{i}" for i in range(nconv)] + [f"dy{i}" for i in range(nlev)]
    print("   checksum run0 vs run1 (first differing planes):", [(labels_[i], f"{abs(a[i]-b[i])/max(a[i],1e-30):.1e}") for i in range(len(a)) if a[i] != b[i]][:12])
    st = bn_stats.cpu().numpy()
    worst = []
    for name, off, C in eng.train_table["batchnorms"]:
        mean, var = ref_stats[name]
        em = np.abs(st[off:off + C] - mean.numpy()).max() / (np.abs(mean.numpy()).max() + np.sqrt(var.numpy().max()))
        ev = np.abs(st[off + C:off + 2 * C] - var.numpy()).max() / np.abs(var.numpy()).max()
        worst.append((max(em, ev), name))
    print("   worst BN stat errors:", [(f"{e:.2e}", n) for e, n in sorted(worst, reverse=True)[:4]])
    rep = []
    for name, off, numel in eng.train_table["params"]:
        g, r = gs[0][off:off + numel].astype(np.float64), ref_grads[name].reshape(-1).numpy()
        if np.linalg.norm(r) < 1e-12:
            continue
        rep.append((float(np.linalg.norm(g - r) / np.linalg.norm(r)), name))
    rep.sort(reverse=True)
    print("   grad rel errors (worst 10):", [(f"{e:.2e}", n) for e, n in rep[:10]])
    print("   grad rel errors (best 5):", [(f"{e:.2e}", n) for e, n in rep[-5:]])
    print(f"   grads run-to-run {float(np.abs(gs[0] - gs[1]).max()):.2e}")
