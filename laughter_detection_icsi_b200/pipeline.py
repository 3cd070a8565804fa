"""Batch inference over many channels: the whole hot path of segment_laughter.load_and_pred
(reference segment_laughter.py:79-122) for a list of channels at once -- PCM -> log-mel (K1) -> per-frame
ResNetBigger probabilities (K2+K3) -> threshold runs (K4) -> float64 min-length filter -- with one H2D copy
of the PCM and one D2H copy of the run lists.  This is the call `bench.py` times end to end and the unit
that is sharded across GPUs (one meeting's channels per rank, no collective in the loop).
"""
import numpy as np
import torch

from . import engine as _engine
from . import laugh_segmenter


class LaughterPipeline:
    def __init__(self, state_dict, device=0, thresholds=(0.5,), min_lengths=(0.2,), mel="lhotse", **engine_kw):
        self.engine = _engine.get_engine(device, **engine_kw)
        self._state_dict = {k: v.detach().clone() for k, v in state_dict.items() if torch.is_tensor(v)}
        self._ensure_weights()
        self.thresholds = [float(t) for t in thresholds]
        self.min_lengths = [float(m) for m in min_lengths]
        self.mel = mel
        self._cap = None

    def _ensure_weights(self):
        """The engine is shared per device: another pipeline / model may have loaded its own checkpoint since."""
        if self.engine.weights_owner is not self:
            self.engine.load_state_dict(self._state_dict)
            self.engine.weights_owner = self

    # --- device-resident stages ------------------------------------------------------------------------
    def probabilities(self, pcm_dev, chan_len):
        """int16 CUDA PCM (channels end to end) -> (float32 CUDA probs, frames per channel)."""
        self._ensure_weights()
        feats, frames = self.engine.fbank(pcm_dev, chan_len, mel=self.mel)
        return self.engine.infer_windows(feats, frames), frames

    def runs(self, probs_dev, frames):
        """Per threshold: (starts, ends, channel) int32 numpy arrays of all maximal runs above the threshold."""
        thr_cmp = laugh_segmenter.comparison_thresholds(self.thresholds, probs_dev.dtype == torch.float32)
        out = self.engine.segment_runs(probs_dev, thr_cmp, self.thresholds, frames, cap=self._cap)
        self._cap = max(self._cap or 0, max(len(s) for s, _, _ in out) + 1024)
        return out

    def step_device(self, pcm_dev, chan_len):
        """One pass of the hot path with inputs already in HBM; results stay on the device except the run lists."""
        probs, frames = self.probabilities(pcm_dev, chan_len)
        return self.runs(probs, frames), frames

    # --- host in, host out ----------------------------------------------------------------------------
    def instances(self, runs, frames, durations_s):
        """runs -> per channel {(thr, min_len): float64 array (n, 2) of (start_s, end_s)}, fps = frames / duration as in
        the reference (segment_laughter.py:103-105).  The run lists arrive in frame order, i.e. already grouped by
        channel; rows compare equal to the reference's list of tuples (``.tolist()``)."""
        out = [dict() for _ in frames]
        fps = [frames[c] / float(durations_s[c]) for c in range(len(frames))]
        min_lengths = self.min_lengths

        def one(args):
            # float64 frame / fps and the strict `end - start > min_len` (laugh_segmenter.py:23-24,108): IEEE double
            # division and subtraction, element-wise -- the same arithmetic as ld_filter_min_length, done once per
            # (threshold, channel) and masked per min_length
            starts, ends, c = args
            s = starts.astype(np.float64) / fps[c]
            e = ends.astype(np.float64) / fps[c]
            d = e - s
            se = np.stack([s, e], axis=1)
            return [se[d > ml] for ml in min_lengths]

        work, keys = [], []
        for (starts, ends, chans), thr in zip(runs, self.thresholds):
            bounds = np.searchsorted(chans, np.arange(len(frames) + 1))
            for c in range(len(frames)):
                work.append((starts[bounds[c]:bounds[c + 1]], ends[bounds[c]:bounds[c + 1]], c))
                keys.append((thr, c))
        n_runs = sum(len(w[0]) for w in work)
        if n_runs > 200000:   # large batches: NumPy releases the GIL, the (threshold, channel) groups run on host threads
            if not hasattr(self, "_pool"):
                from concurrent.futures import ThreadPoolExecutor
                import os
                self._pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
            results = list(self._pool.map(one, work))
        else:
            results = [one(w) for w in work]
        # insertion order of the reference: thresholds-major, min_lengths-minor (laugh_segmenter.py:87)
        for (thr, c), res in zip(keys, results):
            for ml, arr in zip(min_lengths, res):
                out[c][(thr, ml)] = arr
        return out

    def __call__(self, pcm_host, chan_len, durations_s=None):
        """pcm_host: int16 host tensor (pinned for full copy bandwidth). Returns (per-channel instance dicts, frames).
        The channels are copied on a side stream one by one and the network starts on channel c as soon as it has
        landed, so all but the first channel's H2D copy overlaps with compute."""
        dev = self.engine.device
        chan_len = [int(n) for n in chan_len]
        pcm_host = pcm_host.reshape(-1)
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        pcm_dev = torch.empty(pcm_host.numel(), dtype=torch.int16, device=dev)
        self._copy_stream.wait_stream(main)   # the destination was allocated on the main stream
        events, off = [], 0
        with torch.cuda.stream(self._copy_stream):
            for n in chan_len:
                pcm_dev[off:off + n].copy_(pcm_host[off:off + n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
                off += n
        probs, frames, off = [], [], 0
        for n, ev in zip(chan_len, events):
            main.wait_event(ev)
            p, f = self.probabilities(pcm_dev[off:off + n], [n])
            probs.append(p)
            frames += f
            off += n
        pcm_dev.record_stream(self._copy_stream)
        runs = self.runs(torch.cat(probs) if len(probs) > 1 else probs[0], frames)
        if durations_s is None:
            durations_s = [n / float(_engine.SAMPLE_RATE) for n in chan_len]
        return self.instances(runs, frames, durations_s), frames

    @staticmethod
    def h2d_bytes(chan_len):
        return 2 * int(sum(chan_len))

    def d2h_bytes(self):
        """Bytes the last `runs` call copied back (counts + the used part of the start/end/channel lists)."""
        return int(self.engine.last_d2h_bytes)
