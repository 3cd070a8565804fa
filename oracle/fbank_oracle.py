"""CPU restatement of the feature front-end.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party algorithm: lhotse @ git f1b66b8a8db2ea93e87dcb9db3991f6dd473b89d (reference requirements.txt:1),
lhotse/features/kaldi/layers.py ``Wav2LogFilterBank`` (= ``Wav2Win`` -> rfft -> |X|^2 -> mel -> log),
reached from the reference at utils/utils.py:25 ``Fbank(FbankConfig(num_filters=44, frame_shift=0.01))`` and
load_data.py:47-49 ``cut.compute_features(extractor)``.  PARITY UNPINNED against Lhotse itself (not
installable here); ``preproc='frame', mel='kaldi'`` is pinned against torchaudio.compliance.kaldi.fbank
(tests/golden/fbank_kaldi_*.npz).

Defaults restated: sampling_rate 16000, frame_length 25 ms (400), frame_shift 10 ms (160), FFT 512,
remove_dc_offset, preemph 0.97, povey window, dither 0, snip_edges False, low_freq 20, high_freq -400,
num_filters 44, norm_filters False, log(max(., float32 eps)).
"""
import numpy as np
import torch

FRAME_LEN, FRAME_SHIFT, FFT_LEN = 400, 160, 512
EPS = float(torch.finfo(torch.float32).eps)


def num_frames(num_samples):
    return (num_samples + FRAME_SHIFT // 2) // FRAME_SHIFT


def _mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_matrix_lhotse(num_filters=44, sampling_rate=16000, low_freq=20.0, high_freq=-400.0):
    """``create_mel_scale`` of the pinned Lhotse: bin mels from linspace(0, sr, fft_length) (its spacing quirk),
    strict left < m < right support, row 256 all zero, float64 arithmetic cast to float32."""
    if high_freq <= 0:
        high_freq = sampling_rate / 2 + high_freq
    centers = np.linspace(_mel(low_freq), _mel(high_freq), num_filters + 2)
    bins = _mel(np.linspace(0, sampling_rate, FFT_LEN))
    out = np.zeros((FFT_LEN // 2 + 1, num_filters), dtype=np.float32)
    for k in range(num_filters):
        lo, mid, hi = centers[k:k + 3]
        m = bins[: FFT_LEN // 2]
        rising = (m - lo) / (mid - lo)
        falling = (hi - m) / (hi - mid)
        w = np.where(m <= mid, rising, falling)
        out[: FFT_LEN // 2, k] = np.where((m > lo) & (m < hi), w, 0.0)
    return out


def mel_matrix_kaldi(num_filters=44, sampling_rate=16000, low_freq=20.0, high_freq=-400.0):
    """Kaldi ``MelBanks`` / torchaudio ``get_mel_banks`` without VTLN."""
    nyq = 0.5 * sampling_rate
    if high_freq <= 0:
        high_freq += nyq
    lo_m, hi_m = _mel(low_freq), _mel(high_freq)
    step = (hi_m - lo_m) / (num_filters + 1)
    m = _mel(sampling_rate / FFT_LEN * np.arange(FFT_LEN // 2))
    out = np.zeros((FFT_LEN // 2 + 1, num_filters), dtype=np.float32)
    for k in range(num_filters):
        left, center, right = lo_m + k * step, lo_m + (k + 1) * step, lo_m + (k + 2) * step
        out[: FFT_LEN // 2, k] = np.maximum(0.0, np.minimum((m - left) / (center - left), (right - m) / (right - center)))
    return out


def povey_window(dtype=torch.float32):
    return torch.hann_window(FRAME_LEN, periodic=False, dtype=dtype).pow(0.85)


def _frames(x):
    """``_get_strided_batch(snip_edges=False)``: flip-padding 120 left, the rest right; (T, 400) frames."""
    n = x.shape[-1]
    t = num_frames(n)
    left = (FRAME_LEN - FRAME_SHIFT) // 2
    right = (t - 1) * FRAME_SHIFT + FRAME_LEN - n - left
    pieces = [torch.flip(x[:left], (0,)), x]
    if right > 0:
        pieces.append(torch.flip(x[n - right:], (0,)))
    padded = torch.cat(pieces)
    return padded.unfold(0, FRAME_LEN, FRAME_SHIFT)[:t]


def fbank(samples, mel="lhotse", preproc="utterance", dtype=torch.float32):
    """samples: float in [-1, 1) (int16 / 32768), 1-D.  Returns (T, F) log-mel energies.

    preproc='utterance' -- Lhotse ``Wav2Win.forward``: DC offset removed and pre-emphasis applied over the
        whole recording (replicate-padded by one sample), THEN framed.
    preproc='frame'     -- Kaldi / torchaudio: both applied inside each 400-sample frame.
    """
    x = torch.as_tensor(np.asarray(samples), dtype=dtype).reshape(-1)
    B = torch.as_tensor(mel_matrix_lhotse() if isinstance(mel, str) and mel == "lhotse"
                        else mel_matrix_kaldi() if isinstance(mel, str) else np.asarray(mel)).to(dtype)
    win = povey_window(dtype)
    if preproc == "utterance":
        x = x - x.mean()
        prev = torch.cat([x[:1], x[:-1]])
        x = x - 0.97 * prev
        fr = _frames(x)
    elif preproc == "frame":
        fr = _frames(x)
        fr = fr - fr.mean(dim=1, keepdim=True)
        prev = torch.cat([fr[:, :1], fr[:, :-1]], dim=1)
        fr = fr - 0.97 * prev
    else:
        raise ValueError(preproc)
    fr = fr * win
    fr = torch.nn.functional.pad(fr, (0, FFT_LEN - FRAME_LEN))
    spec = torch.fft.rfft(fr, dim=1)
    power = spec.real ** 2 + spec.imag ** 2
    return torch.log(torch.clamp(power @ B, min=EPS))
