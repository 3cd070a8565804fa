"""Evaluation consumer: TextGrid tree -> per-meeting precision / recall over the (threshold, min-length) grid.

Mirror of the reference's analysis/analyse.py, analysis/preprocess.py and analysis/utils.py with the `portion` interval
arithmetic at 1 ms frames replaced by NumPy endpoint sweeps (intervals.IntervalSet) and the import-time global indices
replaced by an explicit `preprocess.Indices` object.  The ICSI transcript parser (analysis/transcript_parsing) is not part
of this package: the indices are built from dataframes of transcribed segments with the columns that parser produces."""
from . import analyse, intervals, preprocess, utils  # noqa: F401
