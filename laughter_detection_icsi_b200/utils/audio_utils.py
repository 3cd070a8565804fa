"""Mirror of the reference's utils/audio_utils.py:7-9 (duration via audioread) for WAV files, plus the PCM
loader the inference path needs (lhotse ``Recording.from_file`` / ``load_audio`` in load_data.py:44-45)."""
import wave

import numpy as np


def get_audio_length(path):
    with wave.open(path, "rb") as f:
        return f.getnframes() / float(f.getframerate())


def load_wav_int16(path):
    """Mono 16-bit PCM WAV -> (int16 samples, sampling_rate); multi-channel files use channel 0 (MonoCut channel=0)."""
    with wave.open(path, "rb") as f:
        if f.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported")
        n, ch, sr = f.getnframes(), f.getnchannels(), f.getframerate()
        data = np.frombuffer(f.readframes(n), dtype="<i2")
    if ch > 1:
        data = data.reshape(-1, ch)[:, 0]
    return np.ascontiguousarray(data), sr
