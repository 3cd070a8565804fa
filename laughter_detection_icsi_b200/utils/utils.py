"""Mirror of the reference's utils/utils.py:6-26: ``get_feat_extractor`` returns the log-mel extractor the
pipeline uses -- here the fused B200 kernel (K1) instead of lhotse's CPU ``Fbank``."""
import numpy as np
import torch

from .. import engine as _engine


class B200Fbank:
    """``extract(samples, sampling_rate) -> (T, num_filters) float32`` like lhotse's ``Fbank.extract``.

    ``samples`` are floats in [-1, 1) as lhotse loads 16-bit audio (int16 / 32768) or int16 directly; they
    are sent to the GPU as int16, which is exact for 16-bit sources."""

    def __init__(self, num_filters, frame_shift, mel="lhotse", device=0):
        if abs(frame_shift - 0.01) > 1e-12:
            raise ValueError("the B200 front-end is built for frame_shift = 1/100 s (config.FEAT['num_samples'] = 100)")
        self.num_filters = num_filters
        self.frame_shift = frame_shift
        self.mel = mel
        self.device = device

    @property
    def sampling_rate(self):
        return _engine.SAMPLE_RATE

    @staticmethod
    def to_int16(samples):
        a = np.asarray(samples)
        if a.dtype == np.int16:
            return a.reshape(-1)
        a = a.astype(np.float64).reshape(-1) * 32768.0
        return np.clip(np.rint(a), -32768, 32767).astype(np.int16)

    def extract(self, samples, sampling_rate):
        assert sampling_rate == _engine.SAMPLE_RATE, f"Fbank was instantiated for sampling_rate {_engine.SAMPLE_RATE}"
        pcm = torch.from_numpy(self.to_int16(samples))
        eng = _engine.get_engine(self.device)
        feats, _ = eng.fbank(pcm.to(eng.device), mel=self.mel)
        return feats.cpu().numpy()


def get_feat_extractor(num_samples, num_filters, use_kaldi=False):
    """frame_shift = 1/num_samples seconds.  (In the reference the Kaldifeat branch is dead code: the CPU Fbank
    always wins, utils/utils.py:25; ``use_kaldi`` is accepted and ignored the same way.)"""
    frame_shift = 1 / num_samples
    return B200Fbank(num_filters=num_filters, frame_shift=frame_shift)
