"""Evaluation consumer: laughter_detection_icsi_b200/analysis against the point-set oracle
(oracle/analysis_oracle.py) on random transcripts and predictions, and end to end over a TextGrid tree written by the
package's own TextGrid writer in the directory layout segment_laughter.py produces."""
import math
import os
import random

import numpy as np
import pandas as pd
import pytest

from laughter_detection_icsi_b200 import textgrid
from laughter_detection_icsi_b200.analysis import analyse, preprocess, utils
from laughter_detection_icsi_b200.analysis.intervals import IntervalSet
from oracle import analysis_oracle as oracle

SEG_COLS = ['meeting_id', 'part_id', 'chan', 'start', 'end', 'length', 'type', 'laugh_type']


def _points(s):
    out = set()
    for a, b in s.pairs():
        out |= set(range(a + 1, b + 1))
    return out


def test_interval_algebra_matches_point_sets():
    rng = random.Random(0)
    for _ in range(300):
        def rand_set():
            pairs = [(a, a + rng.randint(-2, 9)) for a in (rng.randint(0, 60) for _ in range(rng.randint(0, 7)))]
            return IntervalSet.from_pairs(pairs), set().union(*[set(range(a + 1, b + 1)) for a, b in pairs]) if pairs else set()
        (a, pa), (b, pb) = rand_set(), rand_set()
        assert _points(a) == pa and a.length() == len(pa)
        assert all(x[1] < y[0] for x, y in zip(a.pairs(), a.pairs()[1:])), "sorted, disjoint and merged when touching"
        assert _points(a | b) == pa | pb
        assert _points(a & b) == pa & pb
        assert _points(a - b) == pa - pb
        assert a.contains(b) == (pb <= pa)
        assert a.contains_each(b.lo, b.hi).tolist() == [set(range(x + 1, y + 1)) <= pa for x, y in b.pairs()]
        assert a.overlaps(b) == bool(pa & pb)
    assert IntervalSet.openclosed(1, 3) | IntervalSet.openclosed(3, 5) == IntervalSet.openclosed(1, 5)
    assert IntervalSet.openclosed(4, 4).is_empty() and IntervalSet.openclosed(5, 2).is_empty()
    assert utils.to_frames(1.2345) == 1234 and utils.to_frames(0.0005) == 0 and utils.to_sec(1500) == 1.5   # Python round (half to even)


def _random_corpus(seed, n_meetings=3, n_parts=4, length_s=60.0):
    rng = random.Random(seed)
    rows = {"invalid": [], "laugh": [], "speech": [], "noise": []}
    info = []
    for m in range(n_meetings):
        meeting = f"Bmr{m:03d}"
        for p in range(n_parts):
            part, chan = f"me{m}{p:02d}", f"chan{p}"
            info.append({"meeting_id": meeting, "part_id": part, "chan": chan, "length": length_s})
            for kind in rows:
                if m == 2 and kind == "laugh":
                    continue                     # a meeting without any transcribed laughter: recall is NaN
                for _ in range(rng.randint(0, 6)):
                    start = round(rng.uniform(0, length_s - 3), 3)
                    dur = round(rng.choice([0.05, 0.15, 0.4, 1.0, 2.5]), 3)
                    rows[kind].append({"meeting_id": meeting, "part_id": part, "chan": chan, "start": start, "end": start + dur,
                                       "length": dur, "type": kind,
                                       "laugh_type": rng.choice(["laugh", "breath-laugh"]) if kind == "laugh" else None})
    return rows, info


def _build(rows, info):
    dfs = {k: pd.DataFrame(v, columns=SEG_COLS) for k, v in rows.items()}
    idx = preprocess.build_indices(dfs["invalid"], dfs["laugh"], dfs["speech"], dfs["noise"], pd.DataFrame(info))
    o_invalid = oracle.index_from_rows(rows["invalid"])
    o = {"invalid": o_invalid, "laugh": oracle.laugh_index_from_rows(rows["laugh"], o_invalid),
         "speech": oracle.index_from_rows(rows["speech"]), "noise": oracle.index_from_rows(rows["noise"])}
    o["silence"] = oracle.silence_index(info, o["laugh"], o["invalid"], o["noise"], o["speech"])
    for m in {r["meeting_id"] for r in info}:
        for k in ("invalid", "laugh", "speech", "noise"):
            o[k].setdefault(m, {"tot_len": 0, "tot_events": 0})
    return idx, o


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_indices_and_eval_match_the_oracle(seed):
    rows, info = _random_corpus(seed)
    idx, o = _build(rows, info)
    for kind in ("invalid", "laugh", "speech", "noise", "silence"):
        mine = getattr(idx, kind)
        for m, entry in o[kind].items():
            for key, val in entry.items():
                if isinstance(val, set):
                    assert _points(mine[m][key]) == val, (kind, m, key)
                elif key in ("tot_len", "tot_events"):
                    assert mine[m][key] == pytest.approx(val, abs=1e-9)
    rng = random.Random(100 + seed)
    for m in sorted({r["meeting_id"] for r in info}):
        parts = [r["part_id"] for r in info if r["meeting_id"] == m]
        for n_pred in (0, 1, 25):
            preds = []
            for _ in range(n_pred):
                s = round(rng.uniform(0, 57), 2)
                preds.append((rng.choice(parts), s, round(s + rng.choice([0.1, 0.3, 1.0, 3.0]), 2)))
            df = pd.DataFrame([[m, p, "chanX", s, e, e - s, 0.5, "l_0.2", "laugh"] for p, s, e in preds], columns=analyse.PRED_COLUMNS)
            row = analyse.eval_preds(df, m, "0.5", "0.2", idx)
            want = oracle.eval_preds(preds, m, o)
            assert row[:3] == [m, "0.5", "0.2"] and row[10] == len([r for r in rows["laugh"] if r["meeting_id"] == m])
            got = [row[3], row[4], row[5], row[6], row[7], row[8], row[9], row[11], row[12], row[13]]
            for g, w in zip(got, want):
                assert (math.isnan(g) and math.isnan(w)) or g == pytest.approx(w, abs=1e-9), (m, n_pred, got, want)
            if n_pred == 0:
                assert row[3] == 1                                       # no predictions: precision 1 (analyse.py:205-207)
            if m == "Bmr002":
                assert math.isnan(row[4])                                # no transcribed laughter: recall NaN (:209-212)


def test_prediction_inside_an_invalid_region_is_not_evaluated():
    rows = {"invalid": [{"meeting_id": "Bmr000", "part_id": "me001", "chan": "chan0", "start": 10.0, "end": 20.0, "length": 10.0,
                         "type": "invalid", "laugh_type": None}],
            "laugh": [{"meeting_id": "Bmr000", "part_id": "me001", "chan": "chan0", "start": 30.0, "end": 31.0, "length": 1.0,
                       "type": "laugh", "laugh_type": "laugh"},
                      {"meeting_id": "Bmr000", "part_id": "me001", "chan": "chan0", "start": 40.0, "end": 40.1, "length": 0.1,
                       "type": "laugh", "laugh_type": "laugh"}],      # shorter than 0.2 s: becomes invalid (preprocess.py:13-24)
            "speech": [], "noise": []}
    info = [{"meeting_id": "Bmr000", "part_id": "me001", "chan": "chan0", "length": 60.0}]
    idx, _ = _build(rows, info)
    assert idx.invalid["Bmr000"]["me001"].pairs() == [(10000, 20000), (40000, 40100)]
    assert idx.laugh["Bmr000"]["tot_len"] == 1.0 and idx.laugh["Bmr000"]["tot_events"] == 1
    preds = [("me001", 12.0, 13.0), ("me001", 19.5, 20.5), ("me001", 30.5, 31.5)]
    df = pd.DataFrame([["Bmr000", p, "chan0", s, e, e - s, 0.5, "l_0.2", "laugh"] for p, s, e in preds], columns=analyse.PRED_COLUMNS)
    row = dict(zip(analyse.EVAL_COLUMNS, analyse.eval_preds(df, "Bmr000", "0.5", "0.2", idx)))
    assert row["num_of_pred_laughs"] == 3 and row["valid_pred_laughs"] == 2      # the first lies wholly in the invalid region
    assert row["tot_pred_time"] == pytest.approx(0.5 + 1.0) and row["corr_pred_time"] == pytest.approx(0.5)
    assert row["precision"] == pytest.approx(1 / 3) and row["recall"] == pytest.approx(0.5)
    assert row["tot_fp_silence_time"] == pytest.approx(1.0) and row["tot_fp_speech_time"] == 0


def test_textgrid_tree_end_to_end(tmp_path):
    """<out>/<meeting>/t_<thr>/l_<min_len>/chanN.TextGrid as written by segment_laughter.py -> evaluation dataframe."""
    rows, info = _random_corpus(7, n_meetings=2, n_parts=3)
    idx, o = _build(rows, info)
    rng = random.Random(5)
    expected = {}
    for m in ("Bmr000", "Bmr001"):
        for thr in ("0.2", "0.8"):
            for ml in ("0.1", "0.2"):
                d = tmp_path / m / f"t_{thr}" / f"l_{ml}"
                d.mkdir(parents=True)
                preds = []
                for chan in ("chan0", "chan1", "chan2", "chan7"):      # chan7 has no participant: ignored (analyse.py:27-28)
                    t, inst = 0.0, []
                    for _ in range(rng.randint(0, 5)):
                        s = round(t + rng.uniform(0.5, 8.0), 2)
                        e = round(s + rng.choice([0.2, 0.5, 1.5]), 2)
                        inst.append((s, e)); t = e
                    if inst:                                            # segment_laughter writes nothing for empty settings
                        textgrid.write_laughter_textgrid(str(d / f"{chan}.TextGrid"), inst)
                    if chan != "chan7":
                        preds += [(idx.chan_to_part[m][chan], s, e) for s, e in inst]
                expected[(m, thr, ml)] = oracle.eval_preds(preds, m, o)
    out_csv = tmp_path.parent / (tmp_path.name + "_eval") / "eval.csv"
    df = analyse.create_evaluation_df(str(tmp_path), str(out_csv), idx)
    assert os.path.isfile(out_csv) and len(analyse.create_evaluation_df(str(tmp_path), str(out_csv), idx, use_cache=True)) == 8
    assert list(df.columns) == analyse.EVAL_COLUMNS and len(df) == 8
    for _, r in df.iterrows():
        want = expected[(r.meeting, r.threshold, r.min_len)]
        got = [r.precision, r.recall, r.corr_pred_time, r.tot_pred_time, r.tot_transc_laugh_time, r.num_of_pred_laughs,
               r.valid_pred_laughs, r.tot_fp_speech_time, r.tot_fp_noise_time, r.tot_fp_silence_time]
        for g, w in zip(got, want):
            assert (math.isnan(g) and math.isnan(w)) or g == pytest.approx(w, abs=1e-9)
    stats = analyse.calc_sum_stats(df)
    assert list(stats.columns) == ['threshold', 'min_len', 'precision', 'recall'] and len(stats) == 4
    # the command line does the same from CSV files
    seg_csv, info_csv, out_dir = tmp_path.parent / "seg.csv", tmp_path.parent / "info.csv", tmp_path.parent / (tmp_path.name + "_cli")
    pd.concat([pd.DataFrame(v, columns=SEG_COLS) for v in rows.values()]).to_csv(seg_csv, index=False)
    pd.DataFrame(info).to_csv(info_csv, index=False)
    os.makedirs(out_dir)
    assert analyse.main(["--textgrid_dir", str(tmp_path), "--segments_csv", str(seg_csv), "--info_csv", str(info_csv),
                         "--out_dir", str(out_dir)]) == 0
    cli_stats = pd.read_csv(out_dir / "sum_stats.csv")
    assert np.allclose(cli_stats[["precision", "recall"]].to_numpy(), stats[["precision", "recall"]].to_numpy(), equal_nan=True)
    for _, srow in stats.iterrows():
        sel = df[(df.threshold == srow.threshold) & (df.min_len == srow.min_len)]
        tot = sel.tot_pred_time.sum()
        assert srow.precision == pytest.approx(sel.corr_pred_time.sum() / tot if tot else 1)
        assert srow.recall == pytest.approx(sel.corr_pred_time.sum() / sel.tot_transc_laugh_time.sum())


def test_get_params_from_path():
    p = analyse.get_params_from_path("out/Bmr021/t_0.4/l_0.2/chan3_laughter.TextGrid")
    assert p == {"chan_id": "chan3", "min_len": "l_0.2", "threshold": 0.4, "meeting_id": "Bmr021"}
    with pytest.raises(NameError):
        analyse.get_params_from_path("out/Bmr021/t_0.4/l_0.2/mic3.TextGrid")
    with pytest.raises(NameError):
        analyse.get_params_from_path("out/meeting21/t_0.4/l_0.2/chan3.TextGrid")
