"""CPU emulator of the streaming plan emitted by the C++ planner (ld_plan_json).

TEST INFRASTRUCTURE ONLY -- it executes the plan's stem / shifted-plane conv jobs / head with plain torch
ops on the CPU so that the planner's bookkeeping (which rows are window-specific, tap shifts, even/odd
column planes, residual routing, head row order) can be checked against the network evaluated densely
on every window, without a GPU.  With ``half=True`` it rounds stored activations and conv weights to
fp16 exactly where the CUDA kernels do, which gives a per-plane reference for ld_debug_read_plane.
"""
import numpy as np
import torch


def fold_bn(sd, bn, bias=None, eps=1e-5):
    g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
    m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + eps)
    sh = b + s * ((bias.double() if bias is not None else 0.0) - m)
    return s.float(), sh.float()


class PlanEmulator:
    def __init__(self, plan, state_dict, half=False):
        self.plan = plan
        self.sd = {k: v.detach().cpu().float() for k, v in state_dict.items() if v.dtype.is_floating_point}
        self.half = half
        self.planes = {}

    def _q(self, x, split=False):
        """fp16 rounding where the kernels round; split=True keeps the rounding residual as a second fp16 value (hi + lo),
        what precision='split' stores for the weights and planes of blocks 2-4."""
        if not self.half:
            return x
        hi = x.half().float()
        return hi + (x - hi).half().float() if split else hi

    def run(self, feats, nb):
        """feats: (T, W) float tensor of ONE channel starting at sequence row 0; evaluates window starts
        0..nb-1 (rows past T read as zeros). Returns probs[nb]."""
        plan = self.plan
        H, W = plan["H"], plan["W"]
        rows = nb + H
        G = plan["guard_rows"]
        feats = torch.as_tensor(feats, dtype=torch.float32)
        T = feats.shape[0]

        def frow(idx):  # feature rows with zero outside [0, T)
            out = torch.zeros(len(idx), W)
            ok = (idx >= 0) & (idx < T)
            out[ok] = feats[idx[ok]]
            return out

        self.geom = {}
        for p in plan["planes"]:
            g = G * p["wp"] + 256
            self.planes[p["id"]] = torch.zeros(rows * p["wp"] + 2 * g, p["C"])
            self.geom[p["id"]] = (p["wp"], g)

        # ---- stem: conv1 (no bias) + bn1 + relu in fp32 on the raw features
        w = self.sd["conv1.weight"].reshape(64, 3, 3)
        s, sh = fold_bn(self.sd, "bn1")
        wp = plan["stem_wp"]
        r = torch.arange(rows)
        for job in plan["stem"]:
            acc = torch.zeros(rows, W, 64)
            for ky in range(3):
                if not (job["mask"] >> ky) & 1:
                    continue
                f = frow(r + job["row_shift"] + ky - 1)  # (rows, W)
                fpad = torch.nn.functional.pad(f, (1, 1))
                for kx in range(3):
                    acc += fpad[:, kx:kx + W, None] * w[:, ky, kx][None, None, :]
            y = torch.relu(acc * s + sh)
            full = torch.zeros(rows, wp, 64)
            full[:, 1:1 + W] = y
            g = self.geom[job["out"]][1]
            self.planes[job["out"]][g:g + rows * wp] = self._q(full.reshape(rows * wp, 64))

        # ---- convs
        for c in plan["convs"]:
            wt = self.sd[c["conv"] + ".weight"]  # (cout, cin, k, k)
            bias = self.sd.get(c["conv"] + ".bias")
            s, sh = fold_bn(self.sd, c["bn"], bias)
            k = c["ksize"]
            wtaps = self._q(wt.reshape(c["cout"], c["cin"], k * k) * s[:, None, None], c.get("split_w", 0))  # BN scale folded into fp16 weights
            wp = c["wp"]
            M = rows * wp
            col = torch.arange(M) % wp
            inner = (col >= 1) & (col <= c["w_real"])
            for job in c["jobs"]:
                acc = torch.zeros(M, c["cout"])
                for plane, shift, wtap in job["taps"]:
                    pw, g = self.geom[plane]
                    assert pw == wp
                    acc += self.planes[plane][g + shift:g + shift + M] @ wtaps[:, :, wtap].T
                y = acc + sh
                if job["res"] >= 0:
                    pw, g = self.geom[job["res"]]
                    assert pw == wp
                    y = y + self.planes[job["res"]][g + job["res_shift"]:g + job["res_shift"] + M]
                if c["relu"]:
                    y = torch.relu(y)
                y = self._q(y * inner[:, None], c.get("split_out", 0))
                if c["out_mode"] == 0:
                    g = self.geom[job["out0"]][1]
                    self.planes[job["out0"]][g:g + M] = y
                else:
                    wp2 = c["wp2"]
                    row = torch.arange(M) // wp
                    c0 = col - 1
                    dstpix = row * wp2 + (c0 // 2) + 1
                    for par, pid in ((0, job["out0"]), (1, job["out1"])):
                        sel = inner & ((c0 % 2) == par)
                        g = self.geom[pid][1]
                        assert self.geom[pid][0] == wp2
                        self.planes[pid][g + dstpix[sel]] = y[sel]

        # ---- head
        hd = plan["head"]
        wp, C, G4 = hd["wp"], hd["C"], hd["pool_groups"]
        b = torch.arange(nb)
        x = torch.zeros(nb, C, G4)
        for i, (plane, rs) in enumerate(hd["rows"]):
            g = self.geom[plane][1]
            for cpx in range(4):
                x[:, :, i // 4] += self.planes[plane][g + (b + rs) * wp + 1 + cpx]
        x = (x / 16.0).reshape(nb, C * G4)
        s2, h2 = fold_bn(self.sd, "bn2")
        s3, h3 = fold_bn(self.sd, "bn3", self.sd["linear1.bias"])
        x = x * s2 + h2
        h = torch.relu((x @ self.sd["linear1.weight"].T) * s3 + h3)
        z = h @ self.sd["linear2.weight"].T + self.sd["linear2.bias"]
        return torch.sigmoid(z).reshape(-1)

    def plane_as_rows(self, pid, rows):
        """(rows, wp, C) view of a plane, the layout ld_debug_read_plane returns."""
        wp, g = self.geom[pid]
        return self.planes[pid][g:g + rows * wp].reshape(rows, wp, -1)


def dense_window_probs(model, feats, n_frames=100, batch=64):
    """The reference semantics: model(feats[i:i+100] zero-padded) for every frame i (datasets.py:82-93)."""
    feats = torch.as_tensor(feats, dtype=torch.float32)
    T, W = feats.shape
    padded = torch.cat([feats, torch.zeros(n_frames, W)])
    out = []
    with torch.no_grad():
        for i0 in range(0, T, batch):
            idx = torch.arange(i0, min(T, i0 + batch))
            win = torch.stack([padded[i:i + n_frames] for i in idx])[:, None]
            out.append(model(win).reshape(-1))
    return torch.cat(out)
