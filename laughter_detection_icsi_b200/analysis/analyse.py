"""Mirror of the evaluation half of the reference's analysis/analyse.py (:23-298): TextGrid tree -> predicted-laughter
dataframe -> per-meeting precision / recall / false-positive split for every (threshold, min-length) setting, and the
corpus-level summary.  Same function names, argument meaning, column names and edge-case values (precision 1 without
predictions, recall NaN without transcribed laughter); the transcript indices come in as a `preprocess.Indices` object
instead of the reference's import-time globals, and the interval arithmetic is NumPy endpoint sweeps (intervals.py).
The plotting half of analyse.py (:300-547) is not mirrored."""
import os

import pandas as pd

from .. import textgrid
from . import utils
from .intervals import IntervalSet

EVAL_COLUMNS = ['meeting', 'threshold', 'min_len', 'precision', 'recall', 'corr_pred_time', 'tot_pred_time',
                'tot_transc_laugh_time', 'num_of_pred_laughs', 'valid_pred_laughs', 'num_of_transc_laughs',
                'tot_fp_speech_time', 'tot_fp_noise_time', 'tot_fp_silence_time']
PRED_COLUMNS = ['meeting_id', 'part_id', 'chan', 'start', 'end', 'length', 'threshold', 'min_len', 'laugh_type']


# ------------------------------------------------------------------------------------------------ parse the TextGrid tree
def get_params_from_path(path):
    """analyse.py:59-94: .../<meeting_id>/t_<thr>/l_<min_len>/chanN[_...].TextGrid -> parameters."""
    params = {}
    path = os.path.normpath(path)
    params_list = path.replace('.TextGrid', '').split('/')
    chan_id = params_list[-1].split('_')[0]
    if not chan_id.startswith('chan'):
        raise NameError("Did you follow the naming convention for channel .TextGrid-files -> 'chanN.TextGrid'")
    params['chan_id'] = chan_id
    params['min_len'] = params_list[-2]
    params['threshold'] = float(params_list[-3].replace('t_', ''))
    meeting_id = params_list[-4]
    if not len(meeting_id) == 6:
        raise NameError("Did you follow the required directory structure? all chanN.TextGrid files "
                        "need to be in a directory with its meeting ID as name -> e.g. B**NNN")
    params['meeting_id'] = meeting_id
    return params


def textgrid_to_list(full_path, params, indices):
    """analyse.py:23-45: the 'laugh' intervals of the 'laughter' tier of one channel, [] for channels without a participant."""
    chan_to_part = indices.chan_to_part.get(params['meeting_id'], {})
    if params['chan_id'] not in chan_to_part:
        return []
    if os.stat(full_path).st_size == 0:
        print(f"WARNING: Found an empty .TextGrid file for {params['meeting_id']}: {params['chan_id']}")
        return []
    part_id = chan_to_part[params['chan_id']]
    interval_list = []
    for xmin, xmax, text in textgrid.read_intervals(full_path):
        if text == 'laugh':
            interval_list.append([params['meeting_id'], part_id, params['chan_id'], xmin, xmax, xmax - xmin,
                                  params['threshold'], params['min_len'], text])
    return interval_list


def textgrid_to_df(file_path, indices):
    """analyse.py:48-56."""
    tot_list = []
    for filename in sorted(os.listdir(file_path)):
        if filename.endswith('.TextGrid'):
            full_path = os.path.join(file_path, filename)
            tot_list += textgrid_to_list(full_path, get_params_from_path(full_path), indices)
    return pd.DataFrame(tot_list, columns=PRED_COLUMNS)


# ------------------------------------------------------------------------------------------------ analyse
def seg_index_overlap(index, segment, meeting_id, part_id):
    """analyse.py:102-117: seconds of `segment` inside the participant's entry of `index`, 0 without an entry."""
    if part_id not in index[meeting_id]:
        return 0
    return utils.to_sec(utils.p_len(index[meeting_id][part_id] & segment))


def laugh_match(pred_laugh, meeting_id, part_id, indices):
    """analyse.py:119-150: (correct, incorrect, speech, noise, silence) seconds of one participant's predicted laughter, after
    removing what falls into invalid (not evaluated) regions."""
    if part_id in indices.invalid[meeting_id]:
        pred_laugh = pred_laugh - indices.invalid[meeting_id][part_id]
    pred_length = utils.to_sec(utils.p_len(pred_laugh))
    correct, incorrect = 0, pred_length
    if part_id in indices.laugh[meeting_id]:
        correct = seg_index_overlap(indices.laugh, pred_laugh, meeting_id, part_id)
        incorrect = pred_length - correct
    speech_mismatch = seg_index_overlap(indices.speech, pred_laugh, meeting_id, part_id)
    silence_mismatch = seg_index_overlap(indices.silence, pred_laugh, meeting_id, part_id)
    noise_mismatch = seg_index_overlap(indices.noise, pred_laugh, meeting_id, part_id)
    remain_mismatch = incorrect - speech_mismatch - silence_mismatch - noise_mismatch
    assert remain_mismatch < 0.001, \
        f"Accumulated false positives don't match the total incorrect time. Difference: {remain_mismatch}"
    return correct, incorrect, speech_mismatch, noise_mismatch, silence_mismatch


def eval_preds(pred_per_meeting_df, meeting_id, threshold, min_len, indices, print_stats=False):
    """analyse.py:153-236: one row of the evaluation dataframe."""
    tot_corr_pred_time = tot_incorr_pred_time = 0
    tot_fp_speech_time = tot_fp_noise_time = tot_fp_silence_time = 0
    tot_transc_laugh_time = indices.laugh[meeting_id]['tot_len']
    num_of_tranc_laughs = indices.num_transcribed_laughs.get(meeting_id, 0)
    num_of_pred_laughs = pred_per_meeting_df.shape[0]
    num_of_VALID_pred_laughs = 0

    if pred_per_meeting_df.size != 0:
        for part_id, part_df in pred_per_meeting_df.groupby('part_id'):
            starts = [utils.to_frames(v) for v in part_df['start']]
            ends = [utils.to_frames(v) for v in part_df['end']]
            invalid = indices.invalid[meeting_id].get(part_id)
            # a prediction counts as valid unless it lies completely inside an invalid region (analyse.py:185-188)
            num_of_VALID_pred_laughs += len(starts) if invalid is None else int((~invalid.contains_each(starts, ends)).sum())
            part_pred_frames = IntervalSet(starts, ends)     # union of this participant's predictions
            corr, incorr, speech, noise, silence = laugh_match(part_pred_frames, meeting_id, part_id, indices)
            tot_corr_pred_time += corr
            tot_incorr_pred_time += incorr
            tot_fp_speech_time += speech
            tot_fp_noise_time += noise
            tot_fp_silence_time += silence

    tot_predicted_time = tot_corr_pred_time + tot_incorr_pred_time
    prec = 1 if tot_predicted_time == 0 else tot_corr_pred_time / tot_predicted_time
    recall = float('NaN') if tot_transc_laugh_time == 0 else tot_corr_pred_time / tot_transc_laugh_time

    if print_stats:
        print(f'total transcribed time: {tot_transc_laugh_time:.2f}\ntotal predicted time: {tot_predicted_time:.2f}\n'
              f'correct: {tot_corr_pred_time:.2f}\nincorrect: {tot_incorr_pred_time:.2f}\n')
        print(f'Meeting: {meeting_id}\nThreshold: {threshold}\nPrecision: {prec:.4f}\nRecall: {recall:.4f}\n')

    return [meeting_id, threshold, min_len, prec, recall, tot_corr_pred_time, tot_predicted_time, tot_transc_laugh_time,
            num_of_pred_laughs, num_of_VALID_pred_laughs, num_of_tranc_laughs, tot_fp_speech_time, tot_fp_noise_time,
            tot_fp_silence_time]


def create_evaluation_df(path, out_path, indices, use_cache=False):
    """analyse.py:238-283: walk <path>/<meeting>/t_<thr>/l_<min_len>/ and evaluate every setting of every meeting."""
    if use_cache and os.path.isfile(out_path):
        return pd.read_csv(out_path)
    all_evals = []
    for meeting in sorted(os.listdir(path)):
        meeting_path = os.path.join(path, meeting)
        if not os.path.isdir(meeting_path):
            continue
        for threshold in sorted(os.listdir(meeting_path)):
            threshold_dir = os.path.join(meeting_path, threshold)
            for min_length in sorted(os.listdir(threshold_dir)):
                pred_laughs = textgrid_to_df(os.path.join(threshold_dir, min_length), indices)
                all_evals.append(eval_preds(pred_laughs, meeting, threshold.replace('t_', ''), min_length.replace('l_', ''),
                                            indices))
    eval_df = pd.DataFrame(all_evals, columns=EVAL_COLUMNS)
    if out_path:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
        eval_df.to_csv(out_path, index=False)
    return eval_df


def calc_sum_stats(eval_df):
    """analyse.py:286-316: precision / recall once for the whole corpus per (min_len, threshold) -- weighted by time, which
    solves the problem of meetings of different length."""
    sum_vals = eval_df.groupby(['min_len', 'threshold'])[['corr_pred_time', 'tot_pred_time', 'tot_transc_laugh_time']] \
        .sum().reset_index()
    sum_vals['precision'] = sum_vals['corr_pred_time'] / sum_vals['tot_pred_time']
    sum_vals.loc[sum_vals.tot_pred_time == 0, 'precision'] = 1
    sum_vals['recall'] = sum_vals['corr_pred_time'] / sum_vals['tot_transc_laugh_time']
    return sum_vals[['threshold', 'min_len', 'precision', 'recall']]


# ------------------------------------------------------------------------------------------------ command line
def main(argv=None):
    """python -m laughter_detection_icsi_b200.analysis.analyse --textgrid_dir OUT --segments_csv SEG.csv --info_csv INFO.csv

    OUT is the tree segment_laughter.py writes (<meeting>/t_<thr>/l_<min_len>/chanN.TextGrid); SEG.csv holds the transcribed
    segments (columns meeting_id, part_id, chan, start, end, length, type in {invalid, laugh, speech, noise}, laugh_type), INFO.csv
    one row per recorded participant channel (meeting_id, part_id, chan, length).  Writes the per-meeting evaluation dataframe and
    the corpus summary (config.ANALYSIS file names) into --out_dir, like analyse.py's __main__ does before plotting."""
    import argparse

    from ..config import ANALYSIS as cfg
    from . import preprocess

    ap = argparse.ArgumentParser(description=main.__doc__.split("\n")[0])
    ap.add_argument("--textgrid_dir", required=True)
    ap.add_argument("--segments_csv", required=True)
    ap.add_argument("--info_csv", required=True)
    ap.add_argument("--out_dir", default=".")
    args = ap.parse_args(argv)
    seg = pd.read_csv(args.segments_csv)
    info = pd.read_csv(args.info_csv)
    by_type = {t: seg[seg["type"] == t] for t in ("invalid", "laugh", "speech", "noise")}
    indices = preprocess.build_indices(by_type["invalid"], by_type["laugh"], by_type["speech"], by_type["noise"], info)
    eval_df = create_evaluation_df(args.textgrid_dir, os.path.join(args.out_dir, cfg["eval_df_cache_file"]), indices)
    stats = calc_sum_stats(eval_df)
    stats.to_csv(os.path.join(args.out_dir, cfg["sum_stats_cache_file"]), index=False)
    print(stats.to_string(index=False))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
