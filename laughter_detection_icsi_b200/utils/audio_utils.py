"""Mirror of the reference's utils/audio_utils.py:7-9 (duration via audioread), plus the PCM loader the inference path needs
(lhotse ``Recording.from_file`` / ``load_audio`` in load_data.py:44-45): 16-bit PCM WAV and NIST SPHERE files, uncompressed or
shorten-compressed (``sample_coding pcm,embedded-shorten-v2.00``, how ICSI ships its channels; decoded by ld_shorten_decode in
the C library -- the reference's helpers call ``sph2pipe`` for that, analysis/output_processing/laughs_to_wav.py:95-96).  The
decoded stream is checked against the header's sample_count and sample_checksum."""
import ctypes
import os
import wave

import numpy as np


def _decode_shorten(path, h):
    from .. import _native
    lib = _native.load_library()
    with open(path, "rb") as f:
        f.seek(h["_header_bytes"])
        payload = np.frombuffer(f.read(), dtype=np.uint8)
    ch = int(h.get("channel_count", 1))
    want = int(h["sample_count"]) * ch if "sample_count" in h else None
    n_out, n_chan = ctypes.c_int64(0), ctypes.c_int32(0)
    if want is None:
        if lib.ld_shorten_decode(payload.ctypes.data, payload.size, None, 0, ctypes.byref(n_chan), ctypes.byref(n_out)) != 0:
            raise ValueError(f"{path}: {lib.ld_shorten_last_error().decode()}")
        want = int(n_out.value)
    out = np.empty(want, dtype=np.int16)
    if lib.ld_shorten_decode(payload.ctypes.data, payload.size, out.ctypes.data, out.size, ctypes.byref(n_chan), ctypes.byref(n_out)) != 0:
        raise ValueError(f"{path}: {lib.ld_shorten_last_error().decode()}")
    if int(n_out.value) != want or int(n_chan.value) != ch:
        raise ValueError(f"{path}: shorten stream decodes to {n_out.value} samples in {n_chan.value} channel(s), the SPHERE header "
                         f"announces {want} in {ch}")
    if "sample_checksum" in h and not os.environ.get("LD_SPHERE_IGNORE_CHECKSUM"):
        got = int(out.view(np.uint16).astype(np.uint64).sum() & 0xFFFF)
        if got != int(h["sample_checksum"]) & 0xFFFF:
            raise ValueError(f"{path}: decoded samples have checksum {got}, the SPHERE header says {h['sample_checksum']} "
                             "(corrupt file or an unsupported shorten variant; LD_SPHERE_IGNORE_CHECKSUM=1 skips this check)")
    return out


def _sphere_header(path):
    with open(path, "rb") as f:
        magic = f.readline().strip()
        if magic != b"NIST_1A":
            return None
        size = int(f.readline().strip())
        f.seek(0)
        text = f.read(size).decode("latin-1")
    fields = {}
    for line in text.split("\n")[2:]:
        parts = line.strip().split(None, 2)
        if not parts or parts[0] == "end_head":
            break
        if len(parts) == 3:
            fields[parts[0]] = int(parts[2]) if parts[1] == "-i" else parts[2]
    fields["_header_bytes"] = size
    return fields


def _load_sphere_int16(path, h):
    coding = str(h.get("sample_coding", "pcm"))
    if int(h.get("sample_n_bytes", 2)) != 2:
        raise ValueError(f"{path}: only 16-bit PCM SPHERE is supported")
    ch = int(h.get("channel_count", 1))
    if coding.startswith("pcm,embedded-shorten"):
        data = _decode_shorten(path, h)
        if ch > 1:
            data = data.reshape(-1, ch)[:, 0]
        return np.ascontiguousarray(data), int(h["sample_rate"])
    if coding != "pcm":
        raise ValueError(f"{path}: SPHERE sample_coding '{coding}' is not supported (pcm and pcm,embedded-shorten-v* are); "
                         "unpack it first (sph2pipe -f wav)")
    dtype = ">i2" if str(h.get("sample_byte_format", "01")) == "10" else "<i2"
    data = np.fromfile(path, dtype=dtype, offset=h["_header_bytes"], count=int(h["sample_count"]) * ch if "sample_count" in h else -1)
    if ch > 1:
        data = data.reshape(-1, ch)[:, 0]
    return np.ascontiguousarray(data.astype(np.int16)), int(h["sample_rate"])


def get_audio_length(path):
    h = _sphere_header(path)
    if h is not None:
        return _load_sphere_int16(path, h)[0].shape[0] / float(h["sample_rate"]) if "sample_count" not in h \
            else int(h["sample_count"]) / float(h["sample_rate"])
    with wave.open(path, "rb") as f:
        return f.getnframes() / float(f.getframerate())


def load_wav_int16(path):
    """Mono 16-bit PCM WAV (or uncompressed SPHERE) -> (int16 samples, sampling_rate); multi-channel files use channel 0
    (MonoCut channel=0)."""
    h = _sphere_header(path)
    if h is not None:
        return _load_sphere_int16(path, h)
    with wave.open(path, "rb") as f:
        if f.getsampwidth() != 2:
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported")
        n, ch, sr = f.getnframes(), f.getnchannels(), f.getframerate()
        data = np.frombuffer(f.readframes(n), dtype="<i2")
    if ch > 1:
        data = data.reshape(-1, ch)[:, 0]
    return np.ascontiguousarray(data), sr
