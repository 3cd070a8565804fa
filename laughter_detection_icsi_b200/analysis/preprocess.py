"""Mirror of the reference's analysis/preprocess.py: per-meeting, per-participant indices of transcribed segments.

The reference builds five module-level dicts at import time from the parsed ICSI transcripts (preprocess.py:172-213) and the
evaluation reads them as globals.  Here they are fields of one `Indices` object built from dataframes with the columns the
reference's transcript parser produces:
    segments: ['meeting_id', 'part_id', 'chan', 'start', 'end', 'length', 'type', 'laugh_type']
    info:     ['meeting_id', 'part_id', 'chan', 'length']            (one row per recorded participant channel)
Index layout as in the reference: {meeting_id: {'tot_len': seconds, 'tot_events': n, part_id: IntervalSet, ...}}."""
from dataclasses import dataclass, field

from ..config import ANALYSIS as cfg
from . import utils
from .intervals import IntervalSet


def seg_invalid(row):
    """preprocess.py:13-24: shorter than the transcript min-length, or a breath-laugh."""
    return row["length"] < cfg["model"]["min_length"] or row["laugh_type"] == "breath-laugh"


def append_to_index(index, row, meeting_id, part_id):
    """preprocess.py:27-46: add (to_frames(start), to_frames(end)] to the participant's set; tot_len counts every segment's own
    length (overlapping transcriptions count twice, as in the reference)."""
    meeting = index.setdefault(meeting_id, {"tot_len": 0, "tot_events": 0})
    seg = IntervalSet.openclosed(utils.to_frames(row["start"]), utils.to_frames(row["end"]))
    meeting[part_id] = meeting[part_id] | seg if part_id in meeting else seg
    meeting["tot_len"] += utils.to_sec(utils.p_len(seg))
    meeting["tot_events"] += 1
    return index


def create_index_from_df(df):
    """preprocess.py:96-121."""
    index = {}
    for meeting_id, meeting_df in df.groupby("meeting_id"):
        index[meeting_id] = {"tot_len": 0, "tot_events": 0}
        for part_id, part_df in meeting_df.sort_values("start").groupby("part_id"):
            for _, row in part_df.iterrows():
                append_to_index(index, row, meeting_id, part_id)
    return index


def create_laugh_index(df, invalid_index):
    """preprocess.py:49-93: valid laughs go to the laugh index, invalid ones (seg_invalid) are added to `invalid_index`."""
    laugh_index = {}
    for meeting_id, meeting_df in df.groupby("meeting_id"):
        laugh_index[meeting_id] = {"tot_len": 0, "tot_events": 0}
        for part_id, part_df in meeting_df.sort_values("start").groupby("part_id"):
            laugh_index[meeting_id][part_id] = IntervalSet.empty()
            for _, row in part_df.iterrows():
                if seg_invalid(row):
                    append_to_index(invalid_index, row, meeting_id, part_id)
                    continue
                append_to_index(laugh_index, row, meeting_id, part_id)
    return laugh_index


def get_seg_from_index(index, meeting_id, part_id):
    """preprocess.py:124-131."""
    if meeting_id in index:
        return index[meeting_id].get(part_id, IntervalSet.empty())
    return IntervalSet.empty()


def create_silence_index(info_df, laugh_index, invalid_index, noise_index, speech_index):
    """preprocess.py:134-167: the whole channel minus everything transcribed (the reference reads parse.info_df)."""
    silence_index = {}
    for _, row in info_df.iterrows():
        meeting = silence_index.setdefault(row.meeting_id, {})
        full = IntervalSet.openclosed(0, utils.to_frames(row.length))
        seg = (full - get_seg_from_index(laugh_index, row.meeting_id, row.part_id)
               - get_seg_from_index(invalid_index, row.meeting_id, row.part_id)
               - get_seg_from_index(speech_index, row.meeting_id, row.part_id)
               - get_seg_from_index(noise_index, row.meeting_id, row.part_id))
        meeting[row.part_id] = seg
        meeting["tot_length"] = utils.to_sec(utils.p_len(seg))
    return silence_index


@dataclass
class Indices:
    """What analyse.py reads from `prep.*` and `parse.*` in the reference."""
    invalid: dict = field(default_factory=dict)
    laugh: dict = field(default_factory=dict)
    speech: dict = field(default_factory=dict)
    noise: dict = field(default_factory=dict)
    silence: dict = field(default_factory=dict)
    chan_to_part: dict = field(default_factory=dict)      # parse.chan_to_part: {meeting_id: {'chanN': part_id}}
    num_transcribed_laughs: dict = field(default_factory=dict)   # rows of parse.laugh_only_df per meeting


def build_indices(invalid_df, laugh_only_df, speech_df, noise_df, info_df):
    """The 'create indices from scratch' branch of preprocess.py:190-202, plus the two lookups analyse.py takes from parse."""
    idx = Indices()
    idx.invalid = create_index_from_df(invalid_df)
    idx.laugh = create_laugh_index(laugh_only_df, invalid_index=idx.invalid)
    idx.speech = create_index_from_df(speech_df)
    idx.noise = create_index_from_df(noise_df)
    idx.silence = create_silence_index(info_df, idx.laugh, idx.invalid, idx.noise, idx.speech)
    for _, row in info_df.iterrows():
        idx.chan_to_part.setdefault(row.meeting_id, {})[row.chan] = row.part_id
    idx.num_transcribed_laughs = laugh_only_df.groupby("meeting_id").size().to_dict() if len(laugh_only_df) else {}
    # every evaluated meeting has an entry in every index (the reference's corpus-wide dataframes guarantee it)
    for meeting_id in idx.chan_to_part:
        for index in (idx.invalid, idx.laugh, idx.speech, idx.noise):
            index.setdefault(meeting_id, {"tot_len": 0, "tot_events": 0})
    return idx
