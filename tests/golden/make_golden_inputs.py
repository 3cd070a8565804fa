"""Seeded INPUT generators shared by make_golden.py (which runs the reference on them) and the tests."""
import numpy as np


def golden_inputs_resnet(seed=5, n=6):
    rng = np.random.default_rng(seed)
    x = rng.normal(-4.0, 3.0, (n, 1, 100, 44)).astype(np.float32)
    x[1, :, 60:] = 0.0   # a tail window: zero rows, as InferenceDataset pads
    x[2, :, 1:] = 0.0
    return x


def expand_probs(case):
    """Cases either store their probabilities or a tiny generator spec (keeps the fixture small)."""
    if "probs" in case:
        return case["probs"]
    g = case["gen"]
    if g["kind"] == "minlen_edge":
        v = [0.1] * 260
        for i in range(g["start"], g["start"] + 21):
            v[i] = 0.9
        return v
    raise ValueError(g)
