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


@pytest.mark.parametrize("pad_cols", [1, 2])
def test_emulated_plan_matches_dense_windows(pad_cols, monkeypatch):
    """Executing the plan's tap lists plane by plane reproduces the dense per-window forward -- with ONE zero column shared by
    consecutive rows (default: wp = W + 1, pixel p + 1 of a row's last column is the next row's pad) and with the round-1 layout
    of a zero column on either side (LD_PLAN_PAD_COLS=2, kept for A/B runs)."""
    monkeypatch.setenv("LD_PLAN_PAD_COLS", str(pad_cols))
    plan = _native.plan_json()
    assert all(c["wp"] == c["w_real"] + pad_cols for c in plan["convs"]) and plan["stem_wp"] == 44 + pad_cols
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
    n_stacked = 9 * conv["w_blocks"]
    products = []
    for job in conv["jobs"]:
        stage, n_first, n_last, n_pass, open_stage = -1, 0, 0, 0, False
        assert len(job["outs"]) * cout <= conv["tmem_cols"] // conv["n_issuers"]
        for x, y, z, w in job["taps"]:
            a16, b16, lbo16 = x & 0x3FFF, y & 0x3FFF, (y >> 16) & 0x3FFF
            first, last, passing, half_k = bool(x & (1 << 28)), bool(x & (1 << 29)), bool(x & (1 << 30)), int(bool(x & (1 << 27)))
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
            assert w & ((1 << 17) - 1) == 0 and col + n <= conv["tmem_cols"] // conv["n_issuers"]
            assert n % cout == 0 and col % cout == 0 and n % 16 == 0 and 16 <= n <= 256
            g = stage * gps + a16 // box16
            off = a16 % box16
            assert g < len(job["groups"]) and off <= conv["ext_alloc"] - 128
            src, gshift = job["groups"][g]
            for i in range(n // cout):
                if slab:
                    assert n == cout and b16 % (kchunks * cout) == 0
                    wtap = b16 // (kchunks * cout)
                    assert wtap >= n_stacked
                else:
                    assert conv["w_stack"] in (1, 2)
                    blk, in_blk = divmod(b16, 9 * kchunks * cout)    # split precision: block 0 = hi weights, block 1 = lo weights
                    kx, rem = divmod(in_blk, kchunks * 3 * cout)
                    assert blk < conv["w_blocks"] and rem % cout == 0 and rem // cout + n // cout <= 3
                    ky = {1: (2, 1, 0), 2: (2, 0, 1)}[conv["w_stack"]][rem // cout + i]   # stacking order of the weight rows
                    wtap = blk * 9 + ky * 3 + kx
                products.append((job["outs"][col // cout + i][0], src, gshift + off, wtap, half_k))
            if last:
                n_last += 1; open_stage = False
        assert not open_stage and n_first == n_last == job["n_stages"] == -(-len(job["groups"]) // gps) and n_pass == 1
        assert stage == job["n_stages"] - 1
    return products


@pytest.mark.parametrize("filters,precision", [((64, 32, 16, 16), 0), ((64, 48, 32, 16), 0), ((64, 64, 32, 16), 0), ((64, 16, 16, 16), 0),
                                               ((64, 32, 16, 16), 1), ((64, 32, 32, 16), 1)])
def test_tap_program_covers_the_plan(filters, precision):
    """The N-stacked tap programs (chains of outputs, merged MMAs) multiply exactly the products the plan lists -- for
    resnet_base and for other channel widths the kernels are instantiated for; with precision=split (1) every product of
    blocks 2-4 appears against the hi weights (full [hi | lo] depth) and against the lo weights (hi half only)."""
    cfg = _native.default_config(filter_sizes=filters, precision=precision)
    plan = _native.plan_json(cfg)
    prog = _native.gemm_program_json(cfg)
    assert [c["conv"] for c in prog["convs"]] == [c["conv"] for c in plan["convs"]]
    merged = 0
    for pc, gc in zip(plan["convs"], prog["convs"]):
        want = []
        kk = pc["ksize"] ** 2
        for job in pc["jobs"]:
            want += [(job["out0"], p, s, t, 0) for p, s, t in job["taps"]]
            if pc["split_w"]:
                want += [(job["out0"], p, s, kk + t, pc["split_in"]) for p, s, t in job["taps"]]
            if job["res"] >= 0:
                want.append((job["out0"], job["res"], job.get("res_shift", 0), kk * (2 if pc["split_w"] else 1), 0))
        assert gc["cin"] == pc["cin"] * (2 if pc["split_in"] else 1)
        got = _expand_tap_program(gc)
        assert sorted(got) == sorted(want), gc["conv"]
        assert [tuple(o) for j in gc["jobs"] for o in j["outs"]] == [(j["out0"], j["out1"]) for j in pc["jobs"]]
        merged += len(want) - sum(len(j["taps"]) for j in gc["jobs"])
        assert gc["n_stages"] >= 2 and gc["n_rings"] in (1, 2) and gc["n_issuers"] in (2, 4)
        # narrow layers run two CTAs per SM (256 accumulator columns, half the shared memory) when their ring still has >= 4 stages
        assert gc["tmem_cols"] == 512 if gc["cout"] > 32 else gc["tmem_cols"] in (256, 512)
    assert merged > 300, "chains of window-specific rows share their input loads and MMAs"


def test_split_precision_plan_and_error_budget():
    """precision='split' (DESIGN.md section 7): blocks 2-4 keep weights and stored activations as hi + lo fp16 pairs.  On the
    calibrated-head bench checkpoint (head gain ~218) the emulated rounding model stays within 2e-3 of the fp64 oracle where
    plain fp16 is at ~5e-3; block1 stays plain fp16 (tools/precision_budget.py: splitting it buys nothing)."""
    from laughter_detection_icsi_b200 import synth
    from oracle import fbank_oracle
    cfg = _native.default_config(precision=1)
    plan = _native.plan_json(cfg)
    split = {p["tag"].split(".")[0] for p in plan["planes"] if p["split"]}
    plain = {p["tag"].split(".")[0] for p in plan["planes"] if not p["split"]}
    assert split == {"block2", "block3", "block4"} and plain == {"stem", "block1"}
    for c in plan["convs"]:
        assert c["split_w"] == c["split_out"] == int(not c["conv"].startswith("block1"))
        assert c["split_in"] == int(c["conv"].startswith(("block3", "block4", "block2.1")) or c["conv"] == "block2.0.conv2")
    # blocks 2-4 are 11.8 of the 60.7 MMAC: three (two for block2.0's fp16 input) products per tap
    assert 60.7e6 < plan["macs_per_row"] < 60.7e6 + 2.0 * 12.0e6
    assert _native.plan_plane_bytes_per_row(cfg) > _native.plan_plane_bytes_per_row() + 200e3
    sd = synth.synthetic_state_dict()
    nb = 48
    pcm = synth.synth_channel(nb * 160 + 16000 + 37, meeting=0, channel=0)
    feats = fbank_oracle.fbank(pcm.numpy().astype(np.float32) / 32768.0)[: nb + 100]
    ref = resnet_oracle.window_probs(sd, feats.numpy(), dtype=torch.float64)[:nb]
    err_split = np.abs(PlanEmulator(plan, sd, half=True).run(feats, nb).double().numpy() - ref).max()
    err_fp16 = np.abs(PlanEmulator(_native.plan_json(), sd, half=True).run(feats, nb).double().numpy() - ref).max()
    assert err_split < 2e-3 and err_split < err_fp16


def test_shared_memory_traffic_per_row():
    """The third floor of the conv stack behind bench.py's `roofline.smem`: the SS-mode MMAs of the tap programs read ~2.6 MB of
    operand slabs per frame from shared memory and the ring takes ~0.77 MB of bulk copies -- 3.4 MB per frame at 128 B/clk/SM is
    257 ms per 6-channel-hour step at 1.5 GHz, above the HBM (231 ms) and tensor-pipe (186 ms) floors."""
    r, w = _native.plan_gemm_smem_bytes_per_row()
    assert 2.5e6 < r < 2.9e6 and 0.7e6 < w < 0.9e6
    floor_ms = (r + w) * 2.16e6 / 128 / 148 / 1.5e9 * 1e3
    assert 245 < floor_ms < 270
    assert floor_ms > _native.plan_plane_bytes_per_row() * 2.16e6 / 6553.3e9 * 1e3


def test_plane_traffic_per_row():
    """The algorithmic HBM traffic of the conv stack behind bench.py's roofline: ~702 KB of fp16 planes per frame (rows of W + 1 pixels:
    one shared zero column), i.e. 86 MAC = 173 FLOP per byte with the 60.7 MMAC the plan executes -- just left of the B200 ridge
    (~215 FLOP/B)."""
    b = _native.plan_plane_bytes_per_row()
    assert b == 702080.0
    plan = _native.plan_json()
    assert 80 < 2 * plan["macs_per_row"] / b < 180
