"""CPU restatement of the segmenter and of the low-pass filter.  TEST INFRASTRUCTURE ONLY.

Follows the reference's laugh_segmenter.py:57-71 (fix_over_underflow), :74-111 (get_laughter_instances),
:23-24 (frame_span_to_time_span) and :49-55 (lowpass).  Pinned by tests/golden/segmenter_golden.json,
produced by the reference's own laugh_segmenter.py (tests/golden/make_golden.py).
"""
import numpy as np
import scipy.signal


def _clamp(p):
    if p > 1:
        return 1
    if p <= 0:
        return 0.0000001
    return p


def runs_above(probs, thr):
    """Maximal runs of frames whose clamped probability is strictly above thr: [(first, last), ...].
    The comparison is made the way the reference makes it: ``np.min([p]) > thr`` on the element as stored,
    i.e. with NumPy's scalar promotion rules for the element's dtype."""
    runs, start = [], None
    for i, p in enumerate(probs):
        if np.min([_clamp(p)]) > thr:
            if start is None:
                start = i
        elif start is not None:
            runs.append((start, i - 1))
            start = None
    if start is not None:
        runs.append((start, len(probs) - 1))
    return runs


def get_laughter_instances(probs, thresholds=(0.5,), min_lengths=(0.2,), fps=100.0):
    out = {}
    probs = list(probs)
    for thr in thresholds:
        runs = runs_above(probs, thr)
        spans = [(s / fps, e / fps) for s, e in runs]
        for min_l in min_lengths:
            out[(thr, min_l)] = [sp for sp in spans if sp[1] - sp[0] > min_l]
    return out


def lowpass(sig, cutoff=0.01):
    b, a = scipy.signal.butter(2, cutoff, output="ba")
    return scipy.signal.filtfilt(b, a, sig)
