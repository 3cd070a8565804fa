"""Timing of the evaluation consumer (not part of the product): one synthetic meeting of 6 participants x 60 minutes, 87
(threshold, min-length) settings, ~150 predicted laughs per channel and setting.  Compares the NumPy endpoint sweeps of
laughter_detection_icsi_b200.analysis with the point-enumeration arithmetic of the oracle (what `portion` + P.iterate at
1 ms frames cost the reference).  Usage: python tools/eval_bench.py [settings for the oracle leg, default 3]"""
import os
import random
import sys
import time

import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from laughter_detection_icsi_b200.analysis import analyse, preprocess  # noqa: E402
from oracle import analysis_oracle as oracle  # noqa: E402

SEG_COLS = ['meeting_id', 'part_id', 'chan', 'start', 'end', 'length', 'type', 'laugh_type']
rng = random.Random(0)
meeting, length_s = "Bmr000", 3600.0
rows = {"invalid": [], "laugh": [], "speech": [], "noise": []}
info = []
for p in range(6):
    part, chan = f"me0{p:02d}", f"chan{p}"
    info.append({"meeting_id": meeting, "part_id": part, "chan": chan, "length": length_s})
    for kind, n in (("invalid", 20), ("laugh", 150), ("speech", 600), ("noise", 40)):
        for _ in range(n):
            start = round(rng.uniform(0, length_s - 10), 3)
            dur = round(rng.uniform(0.1, 4.0), 3)
            rows[kind].append({"meeting_id": meeting, "part_id": part, "chan": chan, "start": start, "end": start + dur, "length": dur,
                               "type": kind, "laugh_type": "laugh" if kind == "laugh" else None})
settings = [(f"{0.02 + 0.035 * i:.3f}", ml) for i in range(29) for ml in ("0.0", "0.1", "0.2")]
preds = {s: [(f"me0{p:02d}", t, t + rng.uniform(0.2, 3.0)) for p in range(6) for t in sorted(rng.uniform(0, length_s - 5) for _ in range(150))]
         for s in settings}

t0 = time.perf_counter()
dfs = {k: pd.DataFrame(v, columns=SEG_COLS) for k, v in rows.items()}
idx = preprocess.build_indices(dfs["invalid"], dfs["laugh"], dfs["speech"], dfs["noise"], pd.DataFrame(info))
t1 = time.perf_counter()
for (thr, ml), pr in preds.items():
    df = pd.DataFrame([[meeting, p, "chanX", s, e, e - s, float(thr), ml, "laugh"] for p, s, e in pr], columns=analyse.PRED_COLUMNS)
    analyse.eval_preds(df, meeting, thr, ml, idx)
t2 = time.perf_counter()
print(f"interval sweeps: indices {t1 - t0:.2f} s, {len(settings)} settings {t2 - t1:.2f} s = {(t2 - t1) / len(settings) * 1e3:.1f} ms per (meeting, setting)")

n_oracle = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t0 = time.perf_counter()
o_invalid = oracle.index_from_rows(rows["invalid"])
o = {"invalid": o_invalid, "laugh": oracle.laugh_index_from_rows(rows["laugh"], o_invalid), "speech": oracle.index_from_rows(rows["speech"]),
     "noise": oracle.index_from_rows(rows["noise"])}
o["silence"] = oracle.silence_index(info, o["laugh"], o["invalid"], o["noise"], o["speech"])
t1 = time.perf_counter()
for s in settings[:n_oracle]:
    oracle.eval_preds(preds[s], meeting, o)
t2 = time.perf_counter()
print(f"point enumeration (oracle): indices {t1 - t0:.2f} s, {(t2 - t1) / n_oracle * 1e3:.1f} ms per (meeting, setting)")
