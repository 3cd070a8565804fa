/* ld_b200.h -- C ABI of the B200-native laughter-detection hot path.
 *
 * The reference (LasseWolter/laughter-detection-icsi) is pure Python and has no FFI of its own; the
 * drop-in boundary is its Python surface.  Each entry point below names the
 * reference call it replaces (file:line in the reference tree).  The Python host mirror in
 * laughter_detection_icsi_b200/ binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions: plain C types only; every `*_d` / "device" pointer is a CUDA device pointer owned by the
 * caller, every "host" pointer is ordinary host memory; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream); functions return 0 on success or a negative ld_status and record a
 * message retrievable with ld_last_error() (thread local).  One ld_ctx per GPU per process; a context
 * is not thread-safe, distinct contexts are independent.  No entry point ever falls back to the CPU.
 */
#ifndef LD_B200_H_
#define LD_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define LD_API __attribute__((visibility("default")))
#else
#define LD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ld_ctx ld_ctx;

enum ld_status {
    LD_OK = 0,
    LD_ERR_INVALID = -1,   /* bad argument */
    LD_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    LD_ERR_STATE = -3,     /* e.g. weights not loaded */
    LD_ERR_CAPACITY = -4,  /* output buffer too small */
    LD_ERR_UNSUPPORTED = -5
};

/* Feature front-end variant (ambiguity ii). */
enum ld_fbank_preproc {
    LD_PREPROC_UTTERANCE = 0, /* Lhotse Wav2Win: DC removal + pre-emphasis over the whole recording */
    LD_PREPROC_FRAME = 1      /* Kaldi / torchaudio.compliance.kaldi: per 400-sample frame */
};

/* Arithmetic of the conv stack. */
enum ld_precision {
    LD_PRECISION_FP16 = 0,  /* fp16 operands, fp32 accumulation everywhere (default; 1e-5 on random-init weights) */
    LD_PRECISION_SPLIT = 1  /* blocks 2-4 carry weights AND stored activations as hi + lo fp16 pairs (three MMAs per tap):
                               <= 2e-3 on the calibrated-head checkpoint whose 218x head gain amplifies fp16 rounding */
};

typedef struct ld_config {
    int32_t struct_size;     /* sizeof(ld_config), for forward compatibility */
    int32_t num_frames;      /* window length in frames, config.FEAT['num_samples'] (config.py:29) = 100 */
    int32_t num_filters;     /* mel bins, config.FEAT['num_filters'] (config.py:30) = 44 */
    int32_t filter_sizes[4]; /* config.MODEL_MAP[..]['filter_sizes'] (config.py:16) = 64,32,16,16 */
    int32_t linear_layer_size; /* config.py:13 = 48 */
    int32_t chunk_rows;      /* window starts evaluated per pass of the conv stack (0 = default 32768) */
    int32_t fbank_preproc;   /* enum ld_fbank_preproc */
    int32_t precision;       /* enum ld_precision */
    int32_t reserved[5];
} ld_config;

/* One named tensor of a checkpoint's state_dict (host memory, fp32, C-contiguous). */
typedef struct ld_tensor {
    const char* name;    /* e.g. "block2.0.shortcut.0.weight" */
    const float* data;
    int64_t numel;
} ld_tensor;

LD_API const char* ld_last_error(void);
LD_API const char* ld_version(void);

/* Fills *cfg with the resnet_base / FEAT defaults of the reference's config.py:9-31. */
LD_API void ld_default_config(ld_config* cfg);

/* Replaces: model construction + model.set_device(device) (segment_laughter.py:59-61). */
LD_API int ld_create(int device, const ld_config* cfg, ld_ctx** out);
LD_API void ld_destroy(ld_ctx* ctx);

/* K1. Replaces lhotse Fbank(FbankConfig(num_filters=44, frame_shift=0.01)).extract as reached from
 * cut.compute_features(extractor) (load_data.py:47-49, utils/utils.py:25).
 * pcm_d: int16 mono 16 kHz samples of n_chan recordings laid end to end; chan_len[c] samples each (host
 * array).  mel_d: (257, num_filters) fp32 filterbank matrix, row-major (runtime data so that either the
 * Lhotse or the Kaldi bank can be matched).  feats_d: fp32 (sum_c T_c, num_filters), T_c =
 * (chan_len[c] + 80) / 160 (snip_edges=False); frames_out (host, optional) receives T_c. */
LD_API int ld_fbank_i16(ld_ctx* ctx, const int16_t* pcm_d, const int64_t* chan_len, int32_t n_chan,
                 const float* mel_d, float* feats_d, int64_t* frames_out, void* stream);
LD_API int64_t ld_fbank_num_frames(int64_t num_samples);
/* The filterbank behind mel_d is packed into its sparse kernel form when mel_d differs from the previous call's pointer
 * (one D2H copy + stream synchronisation, then none in the steady state).  A caller that rewrites the matrix IN PLACE must
 * call this before the next ld_fbank_i16 so that the new contents are packed. */
LD_API int ld_fbank_reset_mel(ld_ctx* ctx);

/* Replaces model.load_state_dict(checkpoint['state_dict']); model.eval() (segment_laughter.py:63-72):
 * folds every BatchNorm's running statistics into a per-channel scale/shift and repacks the conv
 * weights into the tensor-core operand layout.  Needs the 150 state_dict entries of ResNetBigger
 * (num_batches_tracked entries may be omitted). */
LD_API int ld_resnet_load_weights(ld_ctx* ctx, const ld_tensor* tensors, int32_t n);

/* K2+K3. Replaces the whole inference loop of load_and_pred (segment_laughter.py:90-100) together with
 * InferenceDataset windowing (datasets.py:82-93): for every frame i of every channel, the sigmoid
 * output of ResNetBigger on feats[i:i+100] zero-padded at the tail.
 * feats_d: fp32 (sum_c chan_frames[c], num_filters); probs_d: fp32 (sum_c chan_frames[c]). */
LD_API int ld_resnet_infer_windows(ld_ctx* ctx, const float* feats_d, const int64_t* chan_frames, int32_t n_chan,
                            float* probs_d, void* stream);

/* LAD window gather (training input).  Replaces the reference's cut construction row_track.truncate(offset=sub_start,
 * duration=sub_duration).pad(duration=1.0) (compute_features.py:167) + PrecomputedFeatures()(cuts) (datasets.py:56): builds
 * out_d fp32 (n_windows, num_frames, num_filters) from whole-track features resident on the device.
 * tracks_d: fp32 (sum_t T_t, num_filters), tracks laid end to end; track_off_d / track_len_d: int64 [n_tracks] first row and
 * rows of each track; triples_d: int32 [n_windows][3] = (track, first frame, frames taken); rows past `frames` or past the
 * end of the track are filled with pad_value (Lhotse's LOG_EPSILON for log-domain features). */
LD_API int ld_gather_windows(ld_ctx* ctx, const float* tracks_d, const int64_t* track_off_d, const int64_t* track_len_d,
                      const int32_t* triples_d, int32_t n_windows, float pad_value, float* out_d, void* stream);

/* Audio ingest (host only, no ld_ctx): decodes a Shorten stream (format versions 1-3, 16-bit signed PCM) -- the payload of
 * the ICSI corpus' NIST SPHERE files ("sample_coding pcm,embedded-shorten-v2.00"), which the reference reads through lhotse's
 * Recording.from_file / load_audio (load_data.py:44-45).  data: the bytes after the SPHERE header, starting with "ajkg".
 * out (may be NULL to query the size): interleaved int16 samples, at most cap; *n_out receives the total sample count over
 * all channels, *n_chan_out the channel count.  Errors: ld_shorten_last_error(). */
LD_API int ld_shorten_decode(const uint8_t* data, int64_t n_bytes, int16_t* out, int64_t cap, int32_t* n_chan_out, int64_t* n_out);
LD_API const char* ld_shorten_last_error(void);

/* K4. Replaces the run detection of laugh_segmenter.get_laughter_instances (laugh_segmenter.py:87-105)
 * for n_thr thresholds at once: maximal runs of fix_over_underflow(p) > thr inside each channel, as
 * (first_frame, last_frame) pairs relative to the channel start.  prob_is_f64: probs_d holds doubles.
 * thr_cmp[k] is the value in-range probabilities are compared with and thr_raw[k] the value the clamped
 * constants 1 and 1e-7 are compared with (they differ only when emulating NumPy>=2 float32 scalar
 * comparison, see laugh_segmenter.py in the host mirror).
 * Outputs (device): starts_d/ends_d int32 [n_thr][cap] per-threshold lists in frame order,
 * chan_d int32 [n_thr][cap] channel index of each run, counts_d int32 [n_thr] (may exceed cap: the
 * caller must then retry with a larger cap). */
LD_API int ld_segment_runs(ld_ctx* ctx, const void* probs_d, int32_t prob_is_f64, const int64_t* chan_frames,
                    int32_t n_chan, const double* thr_cmp, const double* thr_raw, int32_t n_thr,
                    int32_t* starts_d, int32_t* ends_d, int32_t* chan_d, int32_t* counts_d, int32_t cap,
                    void* stream);

/* Host helper: the float64 time conversion and strict min-length filter of
 * laugh_segmenter.py:23-24,104-108: keep (s/fps, e/fps) iff e/fps - s/fps > min_len. Returns #kept. */
LD_API int64_t ld_filter_min_length(const int32_t* starts, const int32_t* ends, int64_t n, double fps,
                             double min_len, double* out_start_s, double* out_end_s);

/* K5. Replaces laugh_segmenter.lowpass = scipy.signal.filtfilt(b, a, sig) for a second-order section
 * (laugh_segmenter.py:49-55; padtype='odd', padlen=9, lfilter_zi initial state).  b, a: 3 doubles each
 * (host).  probs_d fp32 or fp64 (prob_is_f64) of length n; out_d: fp64 length n. */
LD_API int ld_lowpass_filtfilt(ld_ctx* ctx, const void* probs_d, int32_t prob_is_f64, int64_t n, const double* b,
                        const double* a, double* out_d, void* stream);
/* Host helper: scipy.signal.butter(2, cutoff) coefficients (laugh_segmenter.py:52). */
LD_API void ld_butter2_lowpass(double cutoff, double* b3, double* a3);

/* End-to-end convenience used by segment_laughter / bench: host int16 PCM in, per-frame probabilities
 * out (host), H2D and D2H copies included.  Equivalent to ld_fbank_i16 + ld_resnet_infer_windows. */
LD_API int ld_infer_pcm_host(ld_ctx* ctx, const int16_t* pcm_host, const int64_t* chan_len, int32_t n_chan,
                      const float* mel_host, float* probs_host, void* stream);

/* ---- Training (reference train.py:261-297: model(src) in .train() mode, loss.backward()) ------------------------------
 * ld_train_create allocates the dense training network for batches of up to max_batch windows (bf16 operands, fp32
 * accumulation and master parameters).  Parameters and gradients travel as ONE flat fp32 device buffer in the order of
 * ResNetBigger.parameters() (ld_train_table_json lists name/offset/numel, and the BatchNorm modules with the offset of
 * their batch mean[C] / biased var[C] inside bn_stats, for the running-statistics update the caller performs).
 * ld_train_forward: x_d (batch,100,44) fp32, mask1_d (batch,48) / mask2_d (batch,32) float 0/1 keep masks of the two dropout
 * sites (models.py:232,235), dropout_p the rate (kept units are scaled by 1/(1-p)); writes probs_d (batch) = sigmoid
 * outputs.  ld_train_backward: dprobs_d = dLoss/dprobs (batch); writes grads_d (n_params).  Conv biases that feed a
 * training-mode BatchNorm have exactly zero gradient. */
LD_API int ld_train_create(ld_ctx* ctx, int32_t max_batch);
LD_API int64_t ld_train_table_json(const ld_ctx* ctx, char* buf, int64_t cap);
LD_API int ld_train_forward(ld_ctx* ctx, const float* params_d, const float* x_d, int32_t batch, const float* mask1_d,
                     const float* mask2_d, float dropout_p, float* probs_d, float* bn_stats_d, void* stream);
LD_API int ld_train_backward(ld_ctx* ctx, const float* dprobs_d, float* grads_d, void* stream);
LD_API int64_t ld_train_kernel_launches(const ld_ctx* ctx);
/* K8. Replaces torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) + optimizer.step() of train.py:292-295 with
 * optim.Adam (train.py:336; PyTorch's update arithmetic) on flat fp32 device vectors of n elements: params updated in place,
 * exp_avg / exp_avg_sq are Adam's state, `step` counts from 1.  max_norm <= 0 disables clipping.  grad_norm_d (optional,
 * device) receives the un-clipped global L2 norm.  For data-parallel training all-reduce grads_d before the call. */
LD_API int ld_clip_adam_step(ld_ctx* ctx, float* params_d, const float* grads_d, float* exp_avg_d, float* exp_avg_sq_d, int64_t n,
                      float max_norm, float lr, float beta1, float beta2, float eps, int64_t step, float* grad_norm_d, void* stream);
/* Same update with the step count kept in DEVICE memory: *step_d (int64, starts at 0) is incremented by the call and the bias
 * corrections are computed on the device, so a CUDA graph that captured the call replays correctly step after step. */
LD_API int ld_clip_adam_step_dev(ld_ctx* ctx, float* params_d, const float* grads_d, float* exp_avg_d, float* exp_avg_sq_d, int64_t n,
                          float max_norm, float lr, float beta1, float beta2, float eps, int64_t* step_d, float* grad_norm_d, void* stream);
/* Debug: sum |value| of every conv output, activation, conv-output gradient and input-gradient plane of the last step. */
LD_API int32_t ld_train_debug_checksums(ld_ctx* ctx, double* out, int32_t cap);
/* Debug: one tensor of the last training step as dense (B, C, H, W) fp32 in host memory.  kind 0: conv output z of conv
 * `index` (order of the parameter table: conv1, block1.0.conv1, block1.0.conv2, ...), 1: activation level, 2: gradient
 * of a conv output, 3: gradient wrt a level, 4: block g, 5: block dh.  dims4 receives (B, C, H, W).  Returns the element
 * count (also when out_host is NULL) or -1. */
LD_API int64_t ld_train_debug_read(ld_ctx* ctx, int32_t kind, int32_t index, float* out_host, int32_t* dims4);

/* Introspection (no GPU needed): JSON description of the streaming plan (planes, conv jobs, taps),
 * consumed by tests/ to check the planner against the reference network on the CPU.  Returns the number
 * of bytes required (including the terminating NUL); writes at most cap bytes. */
LD_API int64_t ld_plan_json(const ld_config* cfg, char* buf, int64_t cap);

/* Introspection (no GPU needed): JSON description of the tensor-core tap program of every conv launch of the plan --
 * jobs (chains of output planes), their load groups (plane id, first pixel shift) and encoded MMA taps
 * (ld_types.h: x = A offset | flags, y = B offset | LBO << 16, z = accumulator column, w = N/8 << 17) -- with plane ids in place of
 * device addresses.  tests/ expand it back into (output, input, shift, weight tap) products and compare with ld_plan_json.
 * Same return convention as ld_plan_json. */
LD_API int64_t ld_gemm_program_json(const ld_config* cfg, char* buf, int64_t cap);

/* Debug: copy one activation plane of the last processed chunk to the host as fp32 [rows][wp][C]. */
LD_API int ld_debug_read_plane(ld_ctx* ctx, int32_t plane_id, int64_t rows, float* out_host);
/* Executed multiply-accumulates per sequence row of the streaming plan, and kernel launches so far. */
LD_API double ld_plan_macs_per_row(const ld_ctx* ctx);
LD_API double ld_plan_gemm_macs_per_row(const ld_ctx* ctx);

/* Device timing per kernel class, for roofline reporting.  While enabled every launch is bracketed by CUDA events
 * recorded on the launch stream.  Classes: 0 conv GEMM (K2), 1 stem, 2 head, 3 fbank (K1), 4 segmenter/low-pass.
 * ld_timing_read synchronises the recorded events and returns the accumulated milliseconds and launch counts
 * (arrays of LD_TIMING_CLASSES); reset != 0 clears the accumulators. */
#define LD_TIMING_CLASSES 5
LD_API int ld_timing_enable(ld_ctx* ctx, int32_t enable);
LD_API int ld_timing_read(ld_ctx* ctx, double* out_ms, int64_t* out_launches, int32_t reset);
LD_API int64_t ld_kernel_launches(const ld_ctx* ctx);
/* Accumulated milliseconds of each conv launch of the plan (order of ld_plan_json's "convs"); returns their number. */
LD_API int32_t ld_timing_read_convs(ld_ctx* ctx, double* out_ms, int32_t cap, int32_t reset);
/* Debug (context created with LD_GEMM_PROF=1 in the environment): 8 cycle counters per conv launch, summed over CTAs:
 * producer wait, MMA wait-operands, MMA wait-accumulator, MMA issue, epilogue wait, epilogue work, CTA lifetime, tiles.
 * Returns the number of conv launches (0 when the counters are off). */
LD_API int32_t ld_debug_gemm_counters(ld_ctx* ctx, uint64_t* out, int32_t cap_convs, int32_t reset);
/* Debug (LD_GEMM_PROF=1): cycles producer warp 0 of every conv launch spent waiting for the neighbouring layers of its
 * layer-pipelined launch, summed over CTAs, in units of 1024 cycles: low 32 bits all waits (dataflow + back-pressure), high
 * 32 bits the dataflow share; read before ld_debug_gemm_counters resets. */
LD_API int32_t ld_debug_gemm_sync_wait(ld_ctx* ctx, uint64_t* out, int32_t cap_convs);
/* Debug (LD_GEMM_PROF=1): per conv launch, the smallest and the largest CTA lifetime per tile (cycles) seen among the CTAs of
 * its launches since the last reset -- how far the persistent CTAs of a launch drift apart; read before ld_debug_gemm_counters
 * resets. */
LD_API int32_t ld_debug_gemm_cta_spread(ld_ctx* ctx, double* min_cycles_per_tile, double* max_cycles_per_tile, int32_t cap_convs);
/* Layer-pipelined conv launches: consecutive conv launches of one shape (cin, cout) and resolution run as the roles of ONE
 * kernel launch, each on its own share of the SMs, and hand their output tiles over through L2 (replaces the per-layer
 * launches of the loop at reference models.py:222-228).  Per conv launch of the plan (order of
 * ld_plan_json's "convs"): its group (-1: launched on its own) and the CTAs its role gets.  Returns the number of conv
 * launches.  Environment at context creation: LD_GEMM_PIPE = 0 (off, default: measured no faster than separate launches) /
 * 1 (cout >= 48 only) / 2 (all). */
LD_API int32_t ld_conv_pipeline_groups(ld_ctx* ctx, int32_t* group_of_conv, int32_t* ctas_of_conv, int32_t cap_convs);

#ifdef __cplusplus
}
#endif
#endif /* LD_B200_H_ */
