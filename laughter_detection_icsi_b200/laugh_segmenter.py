"""Mirror of the reference's laugh_segmenter.py (live functions only: :19-24, :35-42, :49-71, :74-111,
:141-149).  Run detection (K4) and the low-pass (K5) execute on the GPU; the float64 frame->seconds
conversion and the strict min-length comparison run in C doubles on the host so that they round exactly
like the reference's Python floats.
"""
import numpy as np
import torch

from . import engine as _engine


def collapse_to_start_and_end_frame(instance_list):
    return (instance_list[0], instance_list[-1])


def frame_span_to_time_span(frame_span, fps=100.):
    return (frame_span[0] / fps, frame_span[1] / fps)


def seconds_to_frames(s, fps=100):
    return int(s * fps)


def seconds_to_samples(s, sr):
    return s * sr


def cut_laughter_segments(instance_list, y, sr):
    new_audio = []
    for start, end in instance_list:
        sample_start = int(seconds_to_samples(start, sr))
        sample_end = int(seconds_to_samples(end, sr))
        new_audio = np.concatenate([new_audio, y[sample_start:sample_end]])
    return new_audio


def fix_over_underflow(prob):
    """Scalar form of the clamp the GPU kernel applies: > 1 -> 1, <= 0 -> 1e-7 (so threshold 0 keeps it)."""
    if prob > 1:
        print('WARN: Fixed probability > 1')
        return 1
    if prob <= 0:
        print('WARN: Fixed probability <= 0')
        return 0.0000001
    return prob


def lowpass(sig, filter_order=2, cutoff=0.01):
    """Zero-phase 2nd-order Butterworth low-pass (the reference hard-codes order 2, laugh_segmenter.py:51)."""
    eng = _engine.get_engine(_device_of(sig))
    x = _to_device(sig, eng)
    out = eng.lowpass(x, cutoff=cutoff)
    return out if torch.is_tensor(sig) and sig.is_cuda else out.cpu().numpy()


def _device_of(x):
    """The GPU the work runs on: the tensor's own device, else the process's CURRENT device (a rank whose model lives on
    cuda:k must not create a second context on GPU 0)."""
    if torch.is_tensor(x) and x.is_cuda:
        return x.device.index or 0
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


def _to_device(probs, eng):
    if torch.is_tensor(probs):
        t = probs.detach()
        if t.dtype not in (torch.float32, torch.float64):
            t = t.double()
        return t.to(eng.device).reshape(-1)
    a = np.asarray(probs)
    if a.dtype != np.float32:
        a = a.astype(np.float64)  # lists of Python floats compare as doubles in the reference
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).to(eng.device)


def comparison_thresholds(thresholds, prob_is_f32):
    """The value each in-range probability is compared with.  The reference evaluates ``np.min([p]) > thr`` on
    float32 elements: under NumPy >= 2 (NEP 50) the Python-float threshold is cast to float32 first, under NumPy 1.x
    the comparison is made in float64.  Results differ only for probabilities exactly at the float32 rounding of
    the threshold ("threshold ties")."""
    if prob_is_f32 and int(np.__version__.split(".")[0]) >= 2:
        return [float(np.float32(t)) for t in thresholds]
    return [float(t) for t in thresholds]


def get_laughter_runs(probs, thresholds, chan_frames=None):
    """GPU part: for each threshold the (first_frame, last_frame, channel) arrays of all maximal runs."""
    eng = _engine.get_engine(_device_of(probs))
    x = _to_device(probs, eng)
    if x.numel() == 0:
        e = np.zeros(0, dtype=np.int32)
        return [(e, e, e) for _ in thresholds], eng
    thr = [float(t) for t in thresholds]
    return eng.segment_runs(x, comparison_thresholds(thr, x.dtype == torch.float32), thr, chan_frames), eng


def get_laughter_instances(probs, thresholds=[0.5], min_lengths=[0.2], fps=100.):
    """{(threshold, min_length): [(start_s, end_s), ...]} for every setting, thresholds-major like the reference."""
    runs, eng = get_laughter_runs(probs, thresholds)
    instance_dict = {}
    for (starts, ends, _), thr in zip(runs, thresholds):
        for min_l in min_lengths:
            s, e = eng.filter_min_length(starts, ends, fps, min_l)
            instance_dict[(thr, min_l)] = list(zip(s.tolist(), e.tolist()))
    return instance_dict


def format_outputs(instances, wav_paths=None):
    outs = []
    for i in range(len(instances)):
        if wav_paths is not None:
            outs.append({'filename': wav_paths[i], 'start': instances[i][0], 'end': instances[i][1]})
        else:
            outs.append({'start': instances[i][0], 'end': instances[i][1]})
    return outs
