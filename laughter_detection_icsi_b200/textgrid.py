"""Praat TextGrid writer producing what ``tgt.write_to_file(tg, path)`` (format='short') writes for the single
'laughter' IntervalTier built in the reference's segment_laughter.py:150-158: gaps between laughs (and before the
first one, from time 0) are filled with empty intervals, times are printed with Python's ``str(float)``.
tgt is not installable offline, so byte-level parity is unpinned; the consumer (analysis/analyse.py:39-45) reads
interval boundaries and labels only.
"""


def _fill_gaps(intervals, start_time, end_time):
    out, t = [], start_time
    for s, e, text in intervals:
        if s > t:
            out.append((t, s, ""))
        out.append((s, e, text))
        t = e
    if end_time > t:
        out.append((t, end_time, ""))
    return out


def textgrid_short(tier_name, intervals):
    """intervals: [(start, end, text)] sorted and non-overlapping.  Returns the file text."""
    start_time = 0.0 if intervals else 0.0
    end_time = max((e for _, e, _ in intervals), default=0.0)
    filled = _fill_gaps(intervals, min(start_time, intervals[0][0]) if intervals else 0.0, end_time)
    tier_start = filled[0][0] if filled else 0.0
    lines = ['File type = "ooTextFile"', 'Object class = "TextGrid"', '', str(float(tier_start)), str(float(end_time)),
             '<exists>', '1', '"IntervalTier"', '"' + tier_name.replace('"', '""') + '"', str(float(tier_start)),
             str(float(end_time)), str(len(filled))]
    for s, e, text in filled:
        lines += [str(float(s)), str(float(e)), '"' + text.replace('"', '""') + '"']
    return "\n".join(lines)


def write_laughter_textgrid(path, instances, label="laugh", tier_name="laughter"):
    with open(path, "w", encoding="utf-8") as f:
        f.write(textgrid_short(tier_name, [(float(s), float(e), label) for s, e in instances]))


def read_intervals(path):
    """Parses a short-format single-tier TextGrid back into [(start, end, text)] (used by tests)."""
    with open(path, encoding="utf-8") as f:
        lines = [l.rstrip("\n") for l in f]
    n = int(lines[11])
    out = []
    for i in range(n):
        s, e, text = lines[12 + 3 * i: 15 + 3 * i]
        out.append((float(s), float(e), text.strip('"')))
    return out
