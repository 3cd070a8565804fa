"""The C++ streaming planner, executed on the CPU by tests/plan_emulator.py, against the oracle's dense
evaluation of every window -- validates row classification, tap shifts, even/odd column planes, residual
routing and the head's row order without a GPU."""
import numpy as np
import pytest
import torch

from laughter_detection_icsi_b200 import _native
from oracle import resnet_oracle
from plan_emulator import PlanEmulator


def test_plan_structure():
    plan = _native.plan_json()
    assert plan["H"] == 100 and plan["W"] == 44
    names = [c["conv"] for c in plan["convs"]]
    assert names[0] == "block1.0.conv1" and names[-1] == "block4.1.conv2" and len(names) == 19
    ids = {p["id"] for p in plan["planes"]}
    written = set()
    for j in plan["stem"]:
        written.add(j["out"])
    for c in plan["convs"]:
        for job in c["jobs"]:
            assert 1 <= len(job["taps"]) <= 9
            for plane, shift, wtap in job["taps"]:
                assert plane in written, "a conv reads a plane nobody wrote yet"
                assert 0 <= wtap < c["ksize"] ** 2
                assert abs(shift) <= 99 * c["wp"] + 1
            if job["res"] >= 0:
                assert job["res"] in written
            written.add(job["out0"])
            if c["out_mode"] == 1:
                written.add(job["out1"])
    assert written == ids, "every allocated plane is produced by exactly the jobs of the plan"
    assert [r[0] in written for r in plan["head"]["rows"]] == [True] * 12
    # cross-window reuse: ~62 MMAC per frame instead of 708 for a dense window
    assert 55e6 < plan["macs_per_row"] < 70e6


def test_emulated_plan_matches_dense_windows():
    plan = _native.plan_json()
    sd = resnet_oracle.random_state_dict(seed=3)
    rng = np.random.default_rng(0)
    T = 117  # includes 99 tail windows that see zero padding
    feats = rng.normal(-4.0, 3.0, (T, 44)).astype(np.float32)
    ref = resnet_oracle.window_probs(sd, feats)
    out = PlanEmulator(plan, sd).run(torch.from_numpy(feats), T).numpy()
    assert np.abs(out - ref).max() < 5e-7


def test_emulated_plan_fp16_error_budget():
    """Rounding stored activations and conv weights to fp16 (what the CUDA kernels do) stays within the 1e-3
    probability tolerance of the north star for a random-init network."""
    plan = _native.plan_json()
    sd = resnet_oracle.random_state_dict(seed=4)
    rng = np.random.default_rng(1)
    feats = rng.normal(-4.0, 3.0, (40, 44)).astype(np.float32)
    ref = resnet_oracle.window_probs(sd, feats)
    out = PlanEmulator(plan, sd, half=True).run(torch.from_numpy(feats), 40).numpy()
    assert np.abs(out - ref).max() < 1e-3


def _expand_tap_program(conv):
    """Decode the encoded MMA taps of one conv launch (ld_types.h) back into (out plane, src plane, shift, weight tap)
    products, checking the stage bookkeeping flags on the way."""
    cin, cout, gps = conv["cin"], conv["cout"], conv["groups_per_stage"]
    kchunks, box16 = cin // 8, conv["ext_alloc"] * (cin // 8)
    products = []
    for job in conv["jobs"]:
        stage, n_first, n_last, n_pass, open_stage = -1, 0, 0, 0, False
        assert len(job["outs"]) * cout <= 512 // conv["n_issuers"]
        for x, y, z, w in job["taps"]:
            a16, b16, lbo16 = x & 0x3FFF, y & 0x3FFF, (y >> 16) & 0x3FFF
            first, last, passing = bool(x & (1 << 28)), bool(x & (1 << 29)), bool(x & (1 << 30))
            slab = lbo16 == cout
            assert lbo16 in (cout, 3 * cout) and y >> 30 == 0
            if first:
                assert not open_stage
                stage += 1; n_first += 1; open_stage = True
            assert open_stage, "tap outside a stage"
            if passing:
                assert first
                n_pass += 1
            col, n = z, (w >> 17) * 8
            assert w & ((1 << 17) - 1) == 0 and col + n <= 512 // conv["n_issuers"]
            assert n % cout == 0 and col % cout == 0 and n % 16 == 0 and 16 <= n <= 256
            g = stage * gps + a16 // box16
            off = a16 % box16
            assert g < len(job["groups"]) and off <= conv["ext_alloc"] - 128
            src, gshift = job["groups"][g]
            for i in range(n // cout):
                if slab:
                    assert n == cout and b16 % (kchunks * cout) == 0
                    wtap = b16 // (kchunks * cout)
                else:
                    assert conv["w_stack"] in (1, 2)
                    kx, rem = divmod(b16, kchunks * 3 * cout)
                    assert rem % cout == 0 and rem // cout + n // cout <= 3
                    ky = {1: (2, 1, 0), 2: (2, 0, 1)}[conv["w_stack"]][rem // cout + i]   # stacking order of the weight rows
                    wtap = ky * 3 + kx
                products.append((job["outs"][col // cout + i][0], src, gshift + off, wtap))
            if last:
                n_last += 1; open_stage = False
        assert not open_stage and n_first == n_last == job["n_stages"] == -(-len(job["groups"]) // gps) and n_pass == 1
        assert stage == job["n_stages"] - 1
    return products


@pytest.mark.parametrize("filters", [(64, 32, 16, 16), (64, 48, 32, 16), (64, 64, 32, 16), (64, 16, 16, 16)])
def test_tap_program_covers_the_plan(filters):
    """The N-stacked tap programs (chains of outputs, merged MMAs) multiply exactly the products the plan lists -- for
    resnet_base and for other channel widths the kernels are instantiated for."""
    cfg = _native.default_config(filter_sizes=filters)
    plan = _native.plan_json(cfg)
    prog = _native.gemm_program_json(cfg)
    assert [c["conv"] for c in prog["convs"]] == [c["conv"] for c in plan["convs"]]
    merged = 0
    for pc, gc in zip(plan["convs"], prog["convs"]):
        want = []
        for job in pc["jobs"]:
            want += [(job["out0"], p, s, t) for p, s, t in job["taps"]]
            if job["res"] >= 0:
                want.append((job["out0"], job["res"], job.get("res_shift", 0), pc["ksize"] ** 2))
        got = _expand_tap_program(gc)
        assert sorted(got) == sorted(want), gc["conv"]
        assert [tuple(o) for j in gc["jobs"] for o in j["outs"]] == [(j["out0"], j["out1"]) for j in pc["jobs"]]
        merged += len(want) - sum(len(j["taps"]) for j in gc["jobs"])
        assert gc["n_stages"] >= 2 and gc["n_rings"] in (1, 2) and gc["n_issuers"] in (2, 4)
    assert merged > 300, "chains of window-specific rows share their input loads and MMAs"


def test_plane_traffic_per_row():
    """The algorithmic HBM traffic of the conv stack behind bench.py's roofline: ~733 KB of fp16 planes per frame, i.e. 85 FLOP
    per byte with the 62.4 MMAC the plan executes -- left of the B200 ridge, the stack is HBM-bound."""
    b = _native.plan_plane_bytes_per_row()
    assert b == 732928.0
    plan = _native.plan_json()
    assert 80 < 2 * plan["macs_per_row"] / b < 180
