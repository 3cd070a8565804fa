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
                # one process per GPU shares the host cores with its siblings (torchrun sets LOCAL_WORLD_SIZE)
                siblings = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
                self._pool = ThreadPoolExecutor(max_workers=max(2, min(16, (os.cpu_count() or 1) // siblings)))
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
        Software pipeline over the channels: they are copied on a side stream one by one, the network starts on channel c as
        soon as it has landed (all but the first H2D copy overlap with compute), and the run lists of channel c are copied
        back and min-length filtered on the host while the GPU already works on channel c + 1 (K4 in two phases,
        Engine.segment_runs_launch / _collect) -- only the last channel's D2H + filter is exposed."""
        dev = self.engine.device
        chan_len = [int(n) for n in chan_len]
        pcm_host = pcm_host.reshape(-1)
        if durations_s is None:
            durations_s = [n / float(_engine.SAMPLE_RATE) for n in chan_len]
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
        out, frames, off, d2h = [], [], 0, 0
        caps = self.__dict__.setdefault("_chan_caps", {})
        pending = None   # (handle, frames of that channel, duration)

        def finish(p):
            nonlocal d2h
            handle, f, dur, c = p
            runs = self.engine.segment_runs_collect(handle)
            caps[c] = max(caps.get(c, 0), max(len(s) for s, _, _ in runs) + 1024)
            d2h += int(self.engine.last_d2h_bytes)
            out.append(self.instances(runs, f, [dur])[0])

        for c, (n, ev) in enumerate(zip(chan_len, events)):
            main.wait_event(ev)
            p, f = self.probabilities(pcm_dev[off:off + n], [n])
            thr_cmp = laugh_segmenter.comparison_thresholds(self.thresholds, p.dtype == torch.float32)
            handle = self.engine.segment_runs_launch(p, thr_cmp, self.thresholds, f, cap=caps.get(c), slot=c & 1)
            if pending is not None:
                finish(pending)       # host work for channel c - 1 while the GPU runs channel c
            pending = (handle, f, durations_s[c], c)
            frames += f
            off += n
        if pending is not None:
            finish(pending)
        pcm_dev.record_stream(self._copy_stream)
        self._last_d2h = d2h
        return out, frames

    @staticmethod
    def h2d_bytes(chan_len):
        return 2 * int(sum(chan_len))

    def d2h_bytes(self):
        """Bytes the last `runs` call copied back (counts + the used part of the start/end/channel lists)."""
        return int(getattr(self, "_last_d2h", self.engine.last_d2h_bytes))
