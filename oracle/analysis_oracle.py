"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's evaluation arithmetic (analysis/analyse.py:102-236,
analysis/preprocess.py:13-167, analysis/utils.py:8-37) with the `portion` intervals written out as plain Python sets of
integer frame indices: P.openclosed(a, b) on the 1 ms grid is the point set {a+1, ..., b}, `|`, `&`, `-` are set operations,
`contains` is subset, and utils.p_len (the number of points P.iterate(x, step=1) yields) is len().

Parity unpinned: `portion` and `textgrids` are not installable offline and the ICSI transcripts are not in the tree, so the
reference's own analyse.py cannot be run here; this file follows it line by line instead and the product
(laughter_detection_icsi_b200/analysis) is tested against it.  Only tests/ may import this module."""
import math

FRAME_MS = 1          # config.py:51
MIN_LENGTH = 0.2      # config.py:48


def to_frames(t):     # utils.py:8-15
    return round(t * (1000 / FRAME_MS))


def to_sec(n):        # utils.py:18-24
    return n / (1000 / FRAME_MS)


def openclosed(a, b):
    return set(range(a + 1, b + 1))


def seg_invalid(row):   # preprocess.py:13-24
    return row["length"] < MIN_LENGTH or row["laugh_type"] == "breath-laugh"


def _append(index, row):   # preprocess.py:27-46
    meeting = index.setdefault(row["meeting_id"], {"tot_len": 0, "tot_events": 0})
    seg = openclosed(to_frames(row["start"]), to_frames(row["end"]))
    meeting[row["part_id"]] = meeting.get(row["part_id"], set()) | seg
    meeting["tot_len"] += to_sec(len(seg))
    meeting["tot_events"] += 1


def index_from_rows(rows):   # preprocess.py:96-121 (rows: list of dicts, any order -- the reference sorts by start per meeting)
    index = {}
    for row in sorted(rows, key=lambda r: (r["meeting_id"], r["start"])):
        _append(index, row)
    return index


def laugh_index_from_rows(rows, invalid_index):   # preprocess.py:49-93
    laugh = {}
    for row in sorted(rows, key=lambda r: (r["meeting_id"], r["start"])):
        meeting = laugh.setdefault(row["meeting_id"], {"tot_len": 0, "tot_events": 0})
        meeting.setdefault(row["part_id"], set())
        if seg_invalid(row):
            _append(invalid_index, row)
        else:
            _append(laugh, row)
    return laugh


def silence_index(info_rows, laugh, invalid, noise, speech):   # preprocess.py:134-167
    out = {}
    for row in info_rows:
        m, p = row["meeting_id"], row["part_id"]
        seg = openclosed(0, to_frames(row["length"]))
        for index in (laugh, invalid, speech, noise):
            seg = seg - index.get(m, {}).get(p, set())
        out.setdefault(m, {})[p] = seg
    return out


def _overlap(index, seg, m, p):   # analyse.py:102-117
    if p not in index[m]:
        return 0
    return to_sec(len(index[m][p] & seg))


def eval_preds(preds, m, idx):
    """analyse.py:153-236 for one meeting; preds = list of (part_id, start_s, end_s).  Returns the numeric tail of the row:
    (precision, recall, correct, predicted, transcribed, n_pred, n_valid, fp_speech, fp_noise, fp_silence)."""
    corr = incorr = fp_speech = fp_noise = fp_silence = 0
    n_valid = 0
    for p in sorted({q for q, _, _ in preds}):
        frames = set()
        for q, s, e in preds:
            if q != p:
                continue
            seg = openclosed(to_frames(s), to_frames(e))
            if p not in idx["invalid"][m] or not seg <= idx["invalid"][m][p]:
                n_valid += 1
            frames |= seg
        if p in idx["invalid"][m]:                          # laugh_match, analyse.py:119-150
            frames = frames - idx["invalid"][m][p]
        length = to_sec(len(frames))
        c, i = 0, length
        if p in idx["laugh"][m]:
            c = _overlap(idx["laugh"], frames, m, p)
            i = length - c
        corr += c
        incorr += i
        fp_speech += _overlap(idx["speech"], frames, m, p)
        fp_silence += _overlap(idx["silence"], frames, m, p)
        fp_noise += _overlap(idx["noise"], frames, m, p)
    predicted = corr + incorr
    transcribed = idx["laugh"][m]["tot_len"]
    prec = 1 if predicted == 0 else corr / predicted
    recall = math.nan if transcribed == 0 else corr / transcribed
    return prec, recall, corr, predicted, transcribed, len(preds), n_valid, fp_speech, fp_noise, fp_silence
