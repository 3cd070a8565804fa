"""Host-side handle on one GPU's ld_ctx: torch tensors in, torch tensors out, everything computed by
the hand-written sm_100a kernels behind the C ABI (include/ld_b200.h).  PyTorch is used for device
memory and streams only.
"""
import ctypes
import json

import numpy as np
import torch

from . import _native
from ._native import LdError, LdTensor, check, f64_array, i64_array

NUM_BINS = 257  # 512-point real FFT
SAMPLE_RATE = 16000  # ICSI audio; lhotse FbankConfig default (train.py:65 in the reference)


def lhotse_mel_matrix(num_filters=44, sampling_rate=SAMPLE_RATE, fft_length=512, low_freq=20.0, high_freq=-400.0):
    """(257, num_filters) float32 filterbank of lhotse@f1b66b8 ``create_mel_scale`` (norm_filters=False):
    HTK-style mel scale 1127*ln(1+f/700), centres linspace(mel(20), mel(7600), F+2), bin frequencies
    ``linspace(0, sampling_rate, fft_length)`` (the pinned version's spacing), triangular weights.
    Pure host-side setup data for K1 -- the kernel takes the matrix at run time."""
    if high_freq <= 0:
        high_freq = sampling_rate / 2 + high_freq
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    melfc = np.linspace(mel(low_freq), mel(high_freq), num_filters + 2)
    mels = mel(np.linspace(0, sampling_rate, fft_length))
    B = np.zeros((fft_length // 2 + 1, num_filters), dtype=np.float32)
    for k in range(num_filters):
        left, center, right = melfc[k], melfc[k + 1], melfc[k + 2]
        for j in range(fft_length // 2):
            m = mels[j]
            if left < m < right:
                B[j, k] = (m - left) / (center - left) if m <= center else (right - m) / (right - center)
    return B


def kaldi_mel_matrix(num_filters=44, sampling_rate=SAMPLE_RATE, fft_length=512, low_freq=20.0, high_freq=-400.0):
    """(257, num_filters) float32 Kaldi/torchaudio ``get_mel_banks`` filterbank (no VTLN)."""
    nyquist = 0.5 * sampling_rate
    if high_freq <= 0:
        high_freq += nyquist
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    fft_bin_width = sampling_rate / fft_length
    mel_low, mel_high = mel(low_freq), mel(high_freq)
    delta = (mel_high - mel_low) / (num_filters + 1)
    B = np.zeros((fft_length // 2 + 1, num_filters), dtype=np.float32)
    melj = mel(fft_bin_width * np.arange(fft_length // 2))
    for k in range(num_filters):
        left, center, right = mel_low + k * delta, mel_low + (k + 1) * delta, mel_low + (k + 2) * delta
        up = (melj - left) / (center - left)
        down = (right - melj) / (right - center)
        B[: fft_length // 2, k] = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)
    return B


class Engine:
    """One ld_ctx on one CUDA device."""

    def __init__(self, device=0, chunk_rows=0, fbank_preproc=_native.LD_PREPROC_UTTERANCE, filter_sizes=(64, 32, 16, 16),
                 linear_layer_size=48, precision="fp16"):
        self.lib = _native.load_library()
        if not torch.cuda.is_available():
            raise LdError("no CUDA device: the B200 kernels cannot run and there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}")
        self.precision = precision
        self.cfg = _native.default_config(chunk_rows=chunk_rows, fbank_preproc=fbank_preproc, filter_sizes=filter_sizes,
                                          linear_layer_size=linear_layer_size, precision=_native.PRECISIONS[precision])
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            h = ctypes.c_void_p()
            check(self.lib.ld_create(self.device.index, ctypes.byref(self.cfg), ctypes.byref(h)))
        self._h = h
        self._mel_cache = {}
        self.weights_loaded = False
        self.weights_owner = None  # the nn.Module whose parameters are currently loaded (models.py)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ld_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict):
        """model.load_state_dict + eval(): fold BatchNorm, repack convs (segment_laughter.py:63-72)."""
        keep, tensors = [], []
        for name, t in state_dict.items():
            if not torch.is_tensor(t) or not t.dtype.is_floating_point:
                continue
            a = np.ascontiguousarray(t.detach().cpu().float().numpy())
            keep.append((name.encode(), a))
        arr = (LdTensor * len(keep))()
        for i, (n, a) in enumerate(keep):
            arr[i].name = n
            arr[i].data = a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            arr[i].numel = a.size
        check(self.lib.ld_resnet_load_weights(self._h, arr, len(keep)))
        self.weights_loaded = True

    # ------------------------------------------------------------------------------------------ K1
    def mel_device(self, mel):
        key = id(mel) if not isinstance(mel, str) else mel
        if key not in self._mel_cache:
            if isinstance(mel, str):
                m = {"lhotse": lhotse_mel_matrix, "kaldi": kaldi_mel_matrix}[mel](self.cfg.num_filters)
            else:
                m = np.asarray(mel, dtype=np.float32)
            if m.shape != (NUM_BINS, self.cfg.num_filters):
                raise ValueError(f"mel matrix must be ({NUM_BINS}, {self.cfg.num_filters})")
            self._mel_cache[key] = (torch.from_numpy(np.ascontiguousarray(m)).to(self.device), mel)
        return self._mel_cache[key][0]

    def fbank(self, pcm, chan_len=None, mel="lhotse"):
        """pcm: int16 CUDA tensor (channels laid end to end); returns ((sum T, F) float32, [T_c])."""
        if pcm.dtype != torch.int16 or not pcm.is_cuda:
            raise ValueError("pcm must be an int16 CUDA tensor")
        pcm = pcm.contiguous().reshape(-1)
        chan_len = [pcm.numel()] if chan_len is None else [int(x) for x in chan_len]
        if sum(chan_len) != pcm.numel():
            raise ValueError("chan_len does not add up to the number of samples")
        frames = [int(self.lib.ld_fbank_num_frames(n)) for n in chan_len]
        feats = torch.empty((sum(frames), self.cfg.num_filters), dtype=torch.float32, device=self.device)
        mel_d = self.mel_device(mel)
        with torch.cuda.device(self.device):
            check(self.lib.ld_fbank_i16(self._h, pcm.data_ptr(), i64_array(chan_len), len(chan_len), mel_d.data_ptr(),
                                        feats.data_ptr(), None, self._stream()))
        return feats, frames

    # ------------------------------------------------------------------------------------------ K2 + K3
    def infer_windows(self, feats, chan_frames=None):
        """Probability of every frame's 100-frame window (InferenceDataset semantics). feats: (sum T, F) CUDA fp32."""
        if feats.dtype != torch.float32 or not feats.is_cuda or feats.dim() != 2 or feats.shape[1] != self.cfg.num_filters:
            raise ValueError("feats must be a float32 CUDA tensor of shape (T, num_filters)")
        feats = feats.contiguous()
        chan_frames = [feats.shape[0]] if chan_frames is None else [int(x) for x in chan_frames]
        if sum(chan_frames) != feats.shape[0]:
            raise ValueError("chan_frames does not add up to the number of feature rows")
        probs = torch.empty(feats.shape[0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ld_resnet_infer_windows(self._h, feats.data_ptr(), i64_array(chan_frames), len(chan_frames),
                                                   probs.data_ptr(), self._stream()))
        return probs

    def infer_pcm_host(self, pcm_host, chan_len, mel="lhotse"):
        """End to end with HOST buffers (H2D of the PCM and D2H of the probabilities inside the call)."""
        pcm_host = pcm_host.reshape(-1)
        if pcm_host.dtype != torch.int16 or pcm_host.is_cuda:
            raise ValueError("pcm_host must be an int16 host tensor")
        chan_len = [int(x) for x in chan_len]
        if sum(chan_len) != pcm_host.numel():
            raise ValueError("chan_len does not add up to the number of samples")
        frames = [int(self.lib.ld_fbank_num_frames(n)) for n in chan_len]
        probs = torch.empty(sum(frames), dtype=torch.float32, pin_memory=True)
        m = self.mel_device(mel).cpu().contiguous()
        with torch.cuda.device(self.device):
            check(self.lib.ld_infer_pcm_host(self._h, pcm_host.data_ptr(), i64_array(chan_len), len(chan_len), m.data_ptr(),
                                             probs.data_ptr(), self._stream()))
        return probs, frames

    # ------------------------------------------------------------------------------------------ K4 / K5
    def segment_runs(self, probs, thr_cmp, thr_raw=None, chan_frames=None, cap=None):
        """Maximal runs of clamp(p) > thr per threshold: list (per threshold) of (starts, ends, chans) int32 numpy."""
        return self.segment_runs_collect(self.segment_runs_launch(probs, thr_cmp, thr_raw, chan_frames, cap))

    # K4 in two phases, so that a caller can queue more GPU work between the launch and the host-side collection (the run lists
    # of channel c are copied back and filtered on the host while the GPU already works on channel c + 1, pipeline.py)
    def segment_runs_launch(self, probs, thr_cmp, thr_raw=None, chan_frames=None, cap=None, slot=0):
        """Launches K4 into the buffers of `slot` on the current stream and starts the copy of the per-threshold counts to
        pinned host memory; returns a handle for segment_runs_collect.  A slot may be reused once its handle is collected."""
        if not probs.is_cuda or probs.dtype not in (torch.float32, torch.float64):
            raise ValueError("probs must be a float32/float64 CUDA tensor")
        probs = probs.contiguous().reshape(-1)
        chan_frames = [probs.numel()] if chan_frames is None else [int(x) for x in chan_frames]
        if sum(chan_frames) != probs.numel():
            raise ValueError("chan_frames does not add up to the number of probabilities")
        thr_raw = thr_cmp if thr_raw is None else thr_raw
        n_thr = len(thr_cmp)
        cap = int(cap) if cap else max(1024, probs.numel() // 64)
        slots = self.__dict__.setdefault("_seg_slots", {})
        st = slots.get(slot)
        if st is None or st["key"] != (n_thr, cap):   # device lists are reused across calls
            st = {"key": (n_thr, cap),
                  "bufs": tuple(torch.empty((n_thr, cap), dtype=torch.int32, device=self.device) for _ in range(3)),
                  "counts": torch.zeros(n_thr, dtype=torch.int32, device=self.device),
                  "counts_host": torch.zeros(n_thr, dtype=torch.int32, pin_memory=True),
                  "event": torch.cuda.Event(), "host": None}
            slots[slot] = st
        starts, ends, chans = st["bufs"]
        with torch.cuda.device(self.device):
            check(self.lib.ld_segment_runs(self._h, probs.data_ptr(), int(probs.dtype == torch.float64),
                                           i64_array(chan_frames), len(chan_frames), f64_array(thr_cmp), f64_array(thr_raw),
                                           n_thr, starts.data_ptr(), ends.data_ptr(), chans.data_ptr(),
                                           st["counts"].data_ptr(), cap, self._stream()))
        st["counts_host"].copy_(st["counts"], non_blocking=True)
        st["event"].record(torch.cuda.current_stream(self.device))
        return {"slot": slot, "probs": probs, "thr_cmp": list(thr_cmp), "thr_raw": list(thr_raw), "chan_frames": chan_frames,
                "cap": cap, "n_thr": n_thr}

    def segment_runs_collect(self, handle):
        """Waits for the K4 launch of `handle` only (not for work queued after it) and copies the USED part of every run list
        back through one pinned staging buffer, on a side stream (3 x sum(count) int32)."""
        st = self._seg_slots[handle["slot"]]
        n_thr, cap = handle["n_thr"], handle["cap"]
        st["event"].synchronize()
        cnt = st["counts_host"].numpy().copy()
        if cnt.max(initial=0) > cap:   # rare: a list overflowed its capacity -- run again with room, synchronously
            return self.segment_runs_collect(self.segment_runs_launch(handle["probs"], handle["thr_cmp"], handle["thr_raw"],
                                                                      handle["chan_frames"], cap=int(cnt.max()) + 16, slot="retry"))
        starts, ends, chans = st["bufs"]
        total = int(cnt.sum())
        if st["host"] is None or st["host"].numel() < 3 * total:
            st["host"] = torch.empty(max(3 * total, 3 * 1024), dtype=torch.int32, pin_memory=True)
        host = st["host"]
        if getattr(self, "_seg_copy_stream", None) is None:
            self._seg_copy_stream = torch.cuda.Stream(device=self.device)
        side = self._seg_copy_stream
        side.wait_event(st["event"])
        with torch.cuda.stream(side):
            off = 0
            for k in range(n_thr):
                n = int(cnt[k])
                for j, buf in enumerate((starts, ends, chans)):
                    if n:
                        host[off + j * n: off + (j + 1) * n].copy_(buf[k, :n], non_blocking=True)
                off += 3 * n
        side.synchronize()
        arr = host.numpy()
        out, off = [], 0
        for k in range(n_thr):
            n = int(cnt[k])
            out.append((arr[off:off + n].copy(), arr[off + n:off + 2 * n].copy(), arr[off + 2 * n:off + 3 * n].copy()))
            off += 3 * n
        self.last_d2h_bytes = 4 * n_thr + 12 * total
        return out

    def filter_min_length(self, starts, ends, fps, min_len):
        """float64 frame->seconds and strict `end - start > min_len` (laugh_segmenter.py:23-24,108)."""
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        ends = np.ascontiguousarray(ends, dtype=np.int32)
        n = len(starts)
        os_, oe = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        I32, F64 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
        kept = self.lib.ld_filter_min_length(starts.ctypes.data_as(I32), ends.ctypes.data_as(I32), n, float(fps), float(min_len),
                                             os_.ctypes.data_as(F64), oe.ctypes.data_as(F64))
        return os_[:kept], oe[:kept]

    def lowpass(self, probs, cutoff=0.01):
        """filtfilt(butter(2, cutoff)) -> float64 CUDA tensor (laugh_segmenter.py:49-55)."""
        probs = probs.contiguous().reshape(-1)
        b, a = f64_array([0, 0, 0]), f64_array([0, 0, 0])
        self.lib.ld_butter2_lowpass(float(cutoff), b, a)
        out = torch.empty(probs.numel(), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ld_lowpass_filtfilt(self._h, probs.data_ptr(), int(probs.dtype == torch.float64), probs.numel(), b, a,
                                               out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------------------------------ LAD window gather
    def gather_windows(self, tracks, track_off, track_len, triples, pad_value):
        """(B, num_frames, num_filters) float32 CUDA batch of LAD windows from device-resident whole-track features.
        tracks: (sum T, F) float32 CUDA; track_off / track_len: int64 CUDA [n_tracks]; triples: int32 CUDA (B, 3) =
        (track, first frame, frames taken)."""
        if not (tracks.is_cuda and tracks.dtype == torch.float32 and tracks.is_contiguous() and tracks.dim() == 2
                and tracks.shape[1] == self.cfg.num_filters):
            raise ValueError("tracks must be a contiguous float32 CUDA tensor (sum T, num_filters)")
        for t, dt in ((track_off, torch.int64), (track_len, torch.int64), (triples, torch.int32)):
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
                raise ValueError("track tables must be int64 and triples int32 contiguous CUDA tensors")
        if triples.dim() != 2 or triples.shape[1] != 3 or track_off.numel() != track_len.numel():
            raise ValueError("triples must be (B, 3); track_off and track_len must have one entry per track")
        B = triples.shape[0]
        out = torch.empty((B, self.cfg.num_frames, self.cfg.num_filters), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ld_gather_windows(self._h, tracks.data_ptr(), track_off.data_ptr(), track_len.data_ptr(), triples.data_ptr(), B,
                                             float(pad_value), out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------------------------------ training
    def train_create(self, max_batch=256):
        """Allocates the dense training network (bf16 operands) for batches of up to max_batch windows."""
        with torch.cuda.device(self.device):
            check(self.lib.ld_train_create(self._h, int(max_batch)))
        need = self.lib.ld_train_table_json(self._h, None, 0)
        buf = ctypes.create_string_buffer(need)
        self.lib.ld_train_table_json(self._h, buf, need)
        self.train_table = json.loads(buf.value.decode())
        self.train_max_batch = int(max_batch)
        return self.train_table

    def train_forward(self, flat_params, x, mask1, mask2, dropout_p):
        """flat_params: (n_params,) fp32 CUDA in ResNetBigger.parameters() order; x: (B,100,44) fp32 CUDA; masks: float 0/1
        keep masks (B,48), (B,32).  Returns (probs (B,), bn_stats (n_bn_stats,))."""
        B = x.shape[0]
        for t in (flat_params, x, mask1, mask2):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("training tensors must be contiguous float32 CUDA tensors")
        if flat_params.numel() != self.train_table["n_params"]:
            raise ValueError("flat parameter vector has the wrong length")
        probs = torch.empty(B, dtype=torch.float32, device=self.device)
        bn_stats = torch.empty(self.train_table["n_bn_stats"], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ld_train_forward(self._h, flat_params.data_ptr(), x.data_ptr(), B, mask1.data_ptr(), mask2.data_ptr(),
                                            float(dropout_p), probs.data_ptr(), bn_stats.data_ptr(), self._stream()))
        return probs, bn_stats

    def train_backward(self, dprobs):
        """dprobs: dLoss/dprobs (B,) fp32 CUDA of the last train_forward.  Returns the flat gradient (n_params,)."""
        dprobs = dprobs.contiguous().float()
        grads = torch.empty(self.train_table["n_params"], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ld_train_backward(self._h, dprobs.data_ptr(), grads.data_ptr(), self._stream()))
        return grads

    def clip_adam_step(self, params, grads, exp_avg, exp_avg_sq, step, max_norm=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                       grad_norm_out=None):
        """K8: clip_grad_norm_(max_norm) + Adam step on flat fp32 CUDA vectors, in place (ld_clip_adam_step)."""
        for t in (params, grads, exp_avg, exp_avg_sq):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
                raise ValueError("clip_adam_step needs contiguous float32 CUDA vectors of equal length")
        with torch.cuda.device(self.device):
            check(self.lib.ld_clip_adam_step(self._h, params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                             params.numel(), float(max_norm), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                             int(step), grad_norm_out.data_ptr() if grad_norm_out is not None else None, self._stream()))

    def clip_adam_step_dev(self, params, grads, exp_avg, exp_avg_sq, step_d, max_norm=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                           grad_norm_out=None):
        """K8 with the step counter in device memory (int64 CUDA tensor, incremented by the call): graph-capturable."""
        for t in (params, grads, exp_avg, exp_avg_sq):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
                raise ValueError("clip_adam_step needs contiguous float32 CUDA vectors of equal length")
        if not (step_d.is_cuda and step_d.dtype == torch.int64 and step_d.numel() == 1):
            raise ValueError("step_d must be a one-element int64 CUDA tensor")
        with torch.cuda.device(self.device):
            check(self.lib.ld_clip_adam_step_dev(self._h, params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                                 params.numel(), float(max_norm), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                                 step_d.data_ptr(), grad_norm_out.data_ptr() if grad_norm_out is not None else None,
                                                 self._stream()))

    def train_debug_read(self, kind, index):
        """Dense (B, C, H, W) float32 numpy copy of one training tensor (see ld_train_debug_read)."""
        dims = (ctypes.c_int32 * 4)()
        n = self.lib.ld_train_debug_read(self._h, int(kind), int(index), None, dims)
        if n < 0:
            raise LdError("no such training tensor")
        out = np.empty(tuple(dims), dtype=np.float32)
        self.lib.ld_train_debug_read(self._h, int(kind), int(index), out.ctypes.data_as(ctypes.c_void_p), dims)
        return out

    # ------------------------------------------------------------------------------------------ introspection
    def read_plane(self, plane_id, rows, wp, C):
        out = np.empty((rows, wp, C), dtype=np.float32)
        check(self.lib.ld_debug_read_plane(self._h, plane_id, rows, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    @property
    def macs_per_row(self):
        return float(self.lib.ld_plan_macs_per_row(self._h))

    @property
    def gemm_macs_per_row(self):
        return float(self.lib.ld_plan_gemm_macs_per_row(self._h))

    def conv_pipeline_groups(self):
        """Per conv launch of the plan: (group id or -1, CTAs of its role) -- which consecutive conv layers run as the roles of one
        layer-pipelined kernel launch (ld_conv_pipeline_groups)."""
        grp, ctas = (ctypes.c_int32 * 64)(), (ctypes.c_int32 * 64)()
        n = self.lib.ld_conv_pipeline_groups(self._h, grp, ctas, 64)
        if n < 0:
            check(n)
        return [(int(grp[i]), int(ctas[i])) for i in range(n)]

    def conv_launch_names(self):
        """Name of the kernel launch each conv of the plan belongs to: its own name, or the '+'-joined names of its pipelined group."""
        names = [c["conv"] for c in _native.plan_json(self.cfg)["convs"]]
        groups = self.conv_pipeline_groups()
        joined = {}
        for name, (g, _) in zip(names, groups):
            if g >= 0:
                joined[g] = joined.get(g, []) + [name]
        return [("+".join(joined[g]) if g >= 0 else name) for name, (g, _) in zip(names, groups)]

    @property
    def gemm_plane_bytes_per_row(self):
        """Algorithmic HBM traffic of the conv stack per sequence row (see _native.plan_plane_bytes_per_row): every plane once per
        kernel launch that touches it, a layer-pipelined group counting as one launch."""
        return _native.plan_plane_bytes_per_row(self.cfg, groups=[g for g, _ in self.conv_pipeline_groups()])

    def timing_enable(self, enable=True):
        check(self.lib.ld_timing_enable(self._h, int(bool(enable))))

    def timing_read(self, reset=True):
        """{class: (milliseconds, launches)} accumulated since the last reset (synchronises the recorded events)."""
        n = len(_native.TIMING_CLASSES)
        ms, cnt = (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        check(self.lib.ld_timing_read(self._h, ms, cnt, int(bool(reset))))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_native.TIMING_CLASSES)}

    def timing_read_convs(self, reset=True):
        """[(conv name, accumulated ms)] per conv launch of the plan."""
        ms = (ctypes.c_double * 64)()
        n = self.lib.ld_timing_read_convs(self._h, ms, 64, int(bool(reset)))
        if n < 0:
            check(n)
        names = self.conv_launch_names()
        groups = self.conv_pipeline_groups()
        # a pipelined group is timed as one launch under its first conv
        return [(names[i], ms[i]) for i in range(n) if groups[i][0] < 0 or i == 0 or groups[i - 1][0] != groups[i][0]]

    def gemm_counters(self, reset=True):
        """LD_GEMM_PROF=1 only: [(conv name, [8 cycle counters])], see include/ld_b200.h."""
        buf = (ctypes.c_uint64 * (64 * 8))()
        n = self.lib.ld_debug_gemm_counters(self._h, buf, 64, int(bool(reset)))
        if n < 0:
            check(n)
        names = [c["conv"] for c in _native.plan_json(self.cfg)["convs"]]
        return [(names[i], [int(buf[i * 8 + k]) for k in range(8)]) for i in range(n)]

    def gemm_sync_wait(self):
        """LD_GEMM_PROF=1 only: per conv launch, (cycles producer warp 0 waited for the neighbouring roles of its pipelined launch, the
        share of them spent waiting for upstream tiles -- the rest is back-pressure from the consumer)."""
        buf = (ctypes.c_uint64 * 64)()
        n = self.lib.ld_debug_gemm_sync_wait(self._h, buf, 64)
        if n < 0:
            check(n)
        return [((int(buf[i]) & 0xFFFFFFFF) << 10, (int(buf[i]) >> 32) << 10) for i in range(n)]

    def gemm_cta_spread(self):
        """LD_GEMM_PROF=1 only: per conv launch, (min, max) CTA lifetime per tile in cycles among the CTAs of its launches."""
        lo, hi = (ctypes.c_double * 64)(), (ctypes.c_double * 64)()
        n = self.lib.ld_debug_gemm_cta_spread(self._h, lo, hi, 64)
        if n < 0:
            check(n)
        return [(lo[i], hi[i]) for i in range(n)]

    @property
    def kernel_launches(self):
        return int(self.lib.ld_kernel_launches(self._h))

    @property
    def train_kernel_launches(self):
        return int(self.lib.ld_train_kernel_launches(self._h))


_engines = {}


def get_engine(device=0, **kw):
    """Process-wide engine per device (one context per GPU per process)."""
    idx = device if isinstance(device, int) else (torch.device(device).index or 0)
    full = dict(chunk_rows=0, fbank_preproc=_native.LD_PREPROC_UTTERANCE, filter_sizes=(64, 32, 16, 16), linear_layer_size=48,
                precision="fp16")
    full.update(kw)
    full["filter_sizes"] = tuple(int(f) for f in full["filter_sizes"])
    kw = full
    key = (idx, tuple(sorted(kw.items())))
    if key not in _engines:
        _engines[key] = Engine(idx, **kw)
    return _engines[key]
