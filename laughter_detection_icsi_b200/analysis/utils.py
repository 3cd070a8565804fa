"""Mirror of the reference's analysis/utils.py:8-53 on IntervalSet instead of portion intervals."""
from ..config import ANALYSIS as cfg
from .intervals import IntervalSet


def to_frames(time_in_sec):
    """Time in seconds -> number of frames of cfg['model']['frame_duration'] ms (Python round, as in utils.py:8-15)."""
    factor = 1000 / cfg["model"]["frame_duration"]
    return round(time_in_sec * factor)


def to_sec(num_of_frames):
    factor = 1000 / cfg["model"]["frame_duration"]
    return num_of_frames / factor


def p_len(p_interval):
    """Accumulated length in frames (the reference counts the points of P.iterate(x, step=1), utils.py:27-37)."""
    return p_interval.length()


def seg_overlaps(seg, indices, meeting_id, part_id):
    """True if `seg` overlaps any segment of this participant in any of the passed indices (utils.py:40-53)."""
    for index in indices:
        if meeting_id not in index or part_id not in index[meeting_id]:
            continue
        if seg.overlaps(index[meeting_id].get(part_id, IntervalSet.empty())):
            return True
    return False
