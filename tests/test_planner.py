"""The C++ streaming planner, executed on the CPU by tests/plan_emulator.py, against the oracle's dense
evaluation of every window -- validates row classification, tap shifts, even/odd column planes, residual
routing and the head's row order without a GPU."""
import numpy as np
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
