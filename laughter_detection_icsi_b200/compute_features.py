"""GPU feature store and LAD cut construction: the part of the reference's compute_features.py that feeds training
(`compute_features_per_split` :66-111, whole-track features; `compute_features_for_cuts` :114-195, one 1-second cut per
data-frame row with an `is_laugh` supervision) on top of the fused B200 Fbank kernel (K1), without lhotse.

Storage is an own raw format: one float32 ``(T, 44)`` ``.npy`` per (meeting, channel) track plus
``feats.jsonl`` with one record per track.  Cuts follow lhotse's ``truncate(offset, duration).pad(duration=1.0)`` on a
feature matrix with a 10 ms frame shift: first frame = round_half_up(sub_start / 0.01), frames = round_half_up(
min(sub_duration, 1.0) / 0.01), right-padded to 100 frames with lhotse's LOG_EPSILON (recalled: log(1e-10); not pinned --
lhotse is not installable offline).
"""
import json
import math
import os

import numpy as np
import torch

from . import config as cfg
from . import engine as _engine
from .datasets import FeatureCut, LadDataset
from .utils.audio_utils import load_wav_int16

FRAME_SHIFT = 0.01
LOG_EPSILON = math.log(1e-10)   # lhotse.utils.LOG_EPSILON, the pad value of log-domain features
SPLITS = ['train', 'dev', 'test']


def _round_half_up(x):
    return int(math.floor(x + 0.5))


class FeatureStore:
    """(meeting_id, chan_id) -> (T, 44) float32 log-mel features of the whole track, computed by K1."""

    def __init__(self, root=None):
        self.root = root
        self.tracks = {}      # key -> numpy (T, 44)
        self.meta = {}        # key -> {"audio_path", "num_frames", "duration"}

    @staticmethod
    def key(meeting_id, chan_id):
        return f"{meeting_id}/{chan_id}"

    def add_track(self, meeting_id, chan_id, audio_path, device=0, mel="lhotse"):
        """Whole-track features of one WAV (16 kHz, 16 bit) -- compute_features_per_split's per-recording work."""
        pcm, sr = load_wav_int16(audio_path)
        if sr != _engine.SAMPLE_RATE:
            raise ValueError(f"{audio_path}: sampling rate {sr} != {_engine.SAMPLE_RATE}")
        eng = _engine.get_engine(device)
        feats, _ = eng.fbank(torch.from_numpy(pcm).to(eng.device), mel=mel)
        return self.add_features(meeting_id, chan_id, feats.cpu().numpy(), audio_path, len(pcm) / float(sr))

    def add_features(self, meeting_id, chan_id, feats, audio_path="", duration=None):
        k = self.key(meeting_id, chan_id)
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        if feats.ndim != 2 or feats.shape[1] != cfg.FEAT['num_filters']:
            raise ValueError(f"features must be (T, {cfg.FEAT['num_filters']})")
        self.tracks[k] = feats
        self.meta[k] = {"audio_path": audio_path, "num_frames": int(feats.shape[0]),
                        "duration": float(duration if duration is not None else feats.shape[0] * FRAME_SHIFT)}
        return k

    def save(self, root=None):
        root = root or self.root
        os.makedirs(os.path.join(root, 'feats'), exist_ok=True)
        with open(os.path.join(root, 'feats.jsonl'), 'w') as f:
            for k, feats in self.tracks.items():
                fname = os.path.join('feats', k.replace('/', '_') + '.npy')
                np.save(os.path.join(root, fname), feats)
                f.write(json.dumps({"id": k, "features": fname, **self.meta[k]}) + "\n")
        return root

    @classmethod
    def load(cls, root):
        store = cls(root)
        with open(os.path.join(root, 'feats.jsonl')) as f:
            for line in f:
                rec = json.loads(line)
                store.tracks[rec["id"]] = np.load(os.path.join(root, rec["features"]))
                store.meta[rec["id"]] = {k: rec[k] for k in ("audio_path", "num_frames", "duration")}
        return store

    def cut(self, meeting_id, chan_id, sub_start, sub_duration, label, min_seg_duration=1.0, cut_id=None):
        """One LAD cut: truncate(offset=sub_start, duration=sub_duration).pad(duration=min_seg_duration) + is_laugh."""
        feats = self.tracks[self.key(meeting_id, chan_id)]
        n_target = _round_half_up(min_seg_duration / FRAME_SHIFT)
        first = _round_half_up(float(sub_start) / FRAME_SHIFT)
        n = min(_round_half_up(min(float(sub_duration), min_seg_duration) / FRAME_SHIFT), n_target)
        first = max(0, min(first, feats.shape[0]))
        window = feats[first:first + n]
        if window.shape[0] < n_target:
            pad = np.full((n_target - window.shape[0], feats.shape[1]), LOG_EPSILON, dtype=np.float32)
            window = np.concatenate([window, pad], axis=0)
        return FeatureCut(window, int(label), cut_id or f"{meeting_id}_{chan_id}_{first}")


class GpuCutSampler:
    """LAD cuts as index triples over whole-track features that stay in HBM: the batch tensor
    is gathered on the GPU by ld_gather_windows, so training reads no features from the host.  Same frame arithmetic as
    FeatureStore.cut (truncate(sub_start, sub_duration).pad(1 s) on the 10 ms grid, LOG_EPSILON padding)."""

    def __init__(self, store, device=0):
        self.engine = _engine.get_engine(device)
        dev = self.engine.device
        self.keys = list(store.tracks)
        self.index = {k: i for i, k in enumerate(self.keys)}
        lens = [store.tracks[k].shape[0] for k in self.keys]
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if lens else np.zeros(0, dtype=np.int64)
        self.track_len_host = np.asarray(lens, dtype=np.int64)
        self.tracks = torch.from_numpy(np.concatenate([store.tracks[k] for k in self.keys]).astype(np.float32)).to(dev)
        self.track_off = torch.from_numpy(offs).to(dev)
        self.track_len = torch.from_numpy(self.track_len_host).to(dev)

    def triples(self, rows, min_seg_duration=1.0, shuffle_seed=None):
        """(int32 (n, 3) triples, int32 (n,) labels) for data-frame rows, shuffled like cuts_from_dataframe."""
        n_target = _round_half_up(min_seg_duration / FRAME_SHIFT)
        tri = np.zeros((len(rows), 3), dtype=np.int32)
        lab = np.zeros(len(rows), dtype=np.int32)
        for i, r in enumerate(rows):
            t = self.index[FeatureStore.key(r["meeting_id"], r["chan_id"])]
            first = _round_half_up(float(r["sub_start"]) / FRAME_SHIFT)
            n = min(_round_half_up(min(float(r["sub_duration"]), min_seg_duration) / FRAME_SHIFT), n_target)
            first = max(0, min(first, int(self.track_len_host[t])))
            tri[i] = (t, first, n)
            lab[i] = int(r["label"])
        if shuffle_seed is not None:
            order = np.random.default_rng(shuffle_seed).permutation(len(rows))
            tri, lab = tri[order], lab[order]
        return tri, lab

    def batches(self, tri, lab, max_cuts=32):
        """LadDataset-shaped batches whose 'inputs' are CUDA tensors gathered on the device."""
        dev = self.engine.device
        tri_d = torch.from_numpy(np.ascontiguousarray(tri)).to(dev)
        lab_d = torch.from_numpy(np.ascontiguousarray(lab)).to(dev)
        for i in range(0, len(tri), max_cuts):
            t = tri_d[i:i + max_cuts].contiguous()
            inputs = self.engine.gather_windows(self.tracks, self.track_off, self.track_len, t, LOG_EPSILON)
            yield {"inputs": inputs, "input_lens": torch.full((t.shape[0],), cfg.FEAT['num_samples'], dtype=torch.int32),
                   "is_laugh": lab_d[i:i + max_cuts], "cut": None}


def read_data_df(path):
    """Rows of a `{split}_df.csv` (create_data_df.py; columns start,duration,sub_start,sub_duration,audio_path,meeting_id,
    chan_id,label) as dicts."""
    import csv
    with open(path, newline='') as f:
        return [dict(r) for r in csv.DictReader(f)]


def cuts_from_dataframe(rows, store, min_seg_duration=1.0, shuffle_seed=None):
    """compute_features_for_cuts for one split: one cut per data-frame row; shuffled like the reference when a seed is given
    (the frames are sorted speech-first / laugh-last, compute_features.py:186-190)."""
    cuts = [store.cut(r["meeting_id"], r["chan_id"], float(r["sub_start"]), float(r["sub_duration"]), int(r["label"]),
                      min_seg_duration, cut_id=f"cut_{i}") for i, r in enumerate(rows)]
    if shuffle_seed is not None:
        order = np.random.default_rng(shuffle_seed).permutation(len(cuts))
        cuts = [cuts[i] for i in order]
    return cuts


def training_batches(cuts, max_cuts=32):
    """What create_training_dataloader yields: LadDataset batches of `max_cuts` cuts (SingleCutSampler(max_cuts=32))."""
    ds = LadDataset()
    for i in range(0, len(cuts), max_cuts):
        yield ds[cuts[i:i + max_cuts]]
