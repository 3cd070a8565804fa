"""ctypes binding of lib/libld_b200.so (C ABI declared in include/ld_b200.h).

The library is the product; there is no Python/CPU fallback.  Importing this module only needs the
shared object to exist (it is built in-tree by ``python -m laughter_detection_icsi_b200.build``);
creating a context needs a B200.
"""
import ctypes
import json
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libld_b200.so")

LD_PREPROC_UTTERANCE = 0
LD_PREPROC_FRAME = 1
LD_PRECISION_FP16 = 0
LD_PRECISION_SPLIT = 1
PRECISIONS = {"fp16": LD_PRECISION_FP16, "split": LD_PRECISION_SPLIT}
TIMING_CLASSES = ("conv_gemm", "stem", "head", "fbank", "segment")


class LdError(RuntimeError):
    pass


class LdConfig(ctypes.Structure):
    _fields_ = [
        ("struct_size", c_int32),
        ("num_frames", c_int32),
        ("num_filters", c_int32),
        ("filter_sizes", c_int32 * 4),
        ("linear_layer_size", c_int32),
        ("chunk_rows", c_int32),
        ("fbank_preproc", c_int32),
        ("precision", c_int32),
        ("reserved", c_int32 * 5),
    ]


class LdTensor(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("data", POINTER(ctypes.c_float)), ("numel", c_int64)]


# name -> (restype, argtypes); mirrors include/ld_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "ld_last_error": (c_char_p, []),
    "ld_version": (c_char_p, []),
    "ld_default_config": (None, [POINTER(LdConfig)]),
    "ld_create": (c_int, [c_int, POINTER(LdConfig), POINTER(c_void_p)]),
    "ld_destroy": (None, [c_void_p]),
    "ld_fbank_i16": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "ld_fbank_num_frames": (c_int64, [c_int64]),
    "ld_fbank_reset_mel": (c_int, [c_void_p]),
    "ld_resnet_load_weights": (c_int, [c_void_p, POINTER(LdTensor), c_int32]),
    "ld_resnet_infer_windows": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p]),
    "ld_gather_windows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, ctypes.c_float, c_void_p, c_void_p]),
    "ld_shorten_decode": (c_int, [c_void_p, c_int64, c_void_p, c_int64, POINTER(c_int32), POINTER(c_int64)]),
    "ld_shorten_last_error": (c_char_p, []),
    "ld_segment_runs": (c_int, [c_void_p, c_void_p, c_int32, POINTER(c_int64), c_int32, POINTER(c_double), POINTER(c_double),
                                c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "ld_filter_min_length": (c_int64, [POINTER(c_int32), POINTER(c_int32), c_int64, c_double, c_double,
                                       POINTER(c_double), POINTER(c_double)]),
    "ld_lowpass_filtfilt": (c_int, [c_void_p, c_void_p, c_int32, c_int64, POINTER(c_double), POINTER(c_double), c_void_p, c_void_p]),
    "ld_butter2_lowpass": (None, [c_double, POINTER(c_double), POINTER(c_double)]),
    "ld_infer_pcm_host": (c_int, [c_void_p, c_void_p, POINTER(c_int64), c_int32, c_void_p, c_void_p, c_void_p]),
    "ld_plan_json": (c_int64, [POINTER(LdConfig), c_char_p, c_int64]),
    "ld_gemm_program_json": (c_int64, [POINTER(LdConfig), c_char_p, c_int64]),
    "ld_debug_read_plane": (c_int, [c_void_p, c_int32, c_int64, c_void_p]),
    "ld_plan_macs_per_row": (c_double, [c_void_p]),
    "ld_plan_gemm_macs_per_row": (c_double, [c_void_p]),
    "ld_kernel_launches": (c_int64, [c_void_p]),
    "ld_timing_enable": (c_int, [c_void_p, c_int32]),
    "ld_timing_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int32]),
    "ld_timing_read_convs": (c_int32, [c_void_p, POINTER(c_double), c_int32, c_int32]),
    "ld_train_create": (c_int, [c_void_p, c_int32]),
    "ld_train_table_json": (c_int64, [c_void_p, c_char_p, c_int64]),
    "ld_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, ctypes.c_float, c_void_p, c_void_p, c_void_p]),
    "ld_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ld_train_kernel_launches": (c_int64, [c_void_p]),
    "ld_clip_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                           ctypes.c_float, ctypes.c_float, c_int64, c_void_p, c_void_p]),
    "ld_clip_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                               ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_void_p]),
    "ld_train_debug_read": (c_int64, [c_void_p, c_int32, c_int32, c_void_p, POINTER(c_int32)]),
    "ld_train_debug_checksums": (c_int32, [c_void_p, POINTER(c_double), c_int32]),
    "ld_debug_gemm_counters": (c_int32, [c_void_p, POINTER(ctypes.c_uint64), c_int32, c_int32]),
    "ld_debug_gemm_sync_wait": (c_int32, [c_void_p, POINTER(ctypes.c_uint64), c_int32]),
    "ld_debug_gemm_cta_spread": (c_int32, [c_void_p, POINTER(c_double), POINTER(c_double), c_int32]),
    "ld_conv_pipeline_groups": (c_int32, [c_void_p, POINTER(c_int32), POINTER(c_int32), c_int32]),
}

_lib = None


def load_library():
    """dlopen the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m laughter_detection_icsi_b200.build). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise LdError(f"ld_b200 error {status}: {load_library().ld_last_error().decode()}")


def default_config(**overrides):
    cfg = LdConfig()
    load_library().ld_default_config(ctypes.byref(cfg))
    for k, v in overrides.items():
        if k == "filter_sizes":
            for i, f in enumerate(v):
                cfg.filter_sizes[i] = int(f)
        else:
            setattr(cfg, k, int(v))
    return cfg


def plan_json(cfg=None):
    """The streaming plan as a dict (planner only; runs without a GPU)."""
    lib = load_library()
    cfg = cfg if cfg is not None else default_config()
    need = lib.ld_plan_json(ctypes.byref(cfg), None, 0)
    if need < 0:
        raise LdError(lib.ld_last_error().decode())
    buf = ctypes.create_string_buffer(need)
    lib.ld_plan_json(ctypes.byref(cfg), buf, need)
    return json.loads(buf.value.decode())


def plan_plane_bytes_per_row(cfg=None, conv=None, groups=None):
    """Activation bytes the conv launches of the streaming plan move per sequence row if every plane crosses HBM once per
    launch that touches it: for each launch its distinct input, residual and output planes (fp16, wp x C per row).  This is
    the algorithmic traffic figure behind bench.py's HBM roofline.  `conv` restricts the sum to the
    launches of one conv layer (e.g. "block1.1.conv2").  `groups` (per conv: id of its layer-pipelined group or -1, see
    Engine.conv_pipeline_groups) makes the convs of a group ONE launch: a plane written by one role and read by the next
    is handed over through L2 and counts once."""
    plan = plan_json(cfg)
    size = {p["id"]: 2 * p["wp"] * p["C"] * (2 if p.get("split") else 1) for p in plan["planes"]}   # [hi | lo] planes count twice
    total = 0
    touched_of_group = {}
    for i, c in enumerate(plan["convs"]):
        if conv is not None and c["conv"] != conv:
            continue
        g = groups[i] if (groups is not None and conv is None and i < len(groups)) else -1
        touched = touched_of_group.setdefault(g, set()) if g >= 0 else set()
        for job in c["jobs"]:
            touched.update(p for p, _, _ in job["taps"])
            touched.update(x for x in (job["res"], job["out0"], job.get("out1", -1)) if x >= 0)
        if g < 0:
            total += sum(size[i] for i in touched)
    for touched in touched_of_group.values():
        total += sum(size[i] for i in touched)
    return float(total)


def gemm_program_json(cfg=None):
    """The tensor-core tap programs of the plan's conv launches as a dict (host builder only; runs without a GPU)."""
    lib = load_library()
    cfg = cfg if cfg is not None else default_config()
    need = lib.ld_gemm_program_json(ctypes.byref(cfg), None, 0)
    if need < 0:
        raise LdError(lib.ld_last_error().decode())
    buf = ctypes.create_string_buffer(need)
    lib.ld_gemm_program_json(ctypes.byref(cfg), buf, need)
    return json.loads(buf.value.decode())


def plan_gemm_smem_bytes_per_row(cfg=None):
    """Shared-memory traffic of the conv launches per sequence row: what the tensor core reads from shared memory for its SS-mode
    MMAs (a 128 x 16 A slab = 4 KB and an N x 16 B slab per MMA, straight from the tap programs) plus what the bulk copies write
    into the ring (every load group of every job once per 128-pixel tile).  At 128 B/clk per SM this is the highest of the three
    floors of the conv stack (HBM planes, tensor pipe, shared-memory operand bandwidth).
    Returns (operand_read_bytes, bulk_copy_write_bytes) per row."""
    prog = gemm_program_json(cfg)["convs"]
    wp = {c["conv"]: c["wp"] for c in plan_json(cfg)["convs"]}
    reads = writes = 0.0
    for c in prog:
        ks = c["cin"] // 16
        per_tile_r = per_tile_w = 0
        for j in c["jobs"]:
            for x, _, _, w in j["taps"]:
                k = ks // 2 if (x >> 27) & 1 else ks
                per_tile_r += k * (128 * 32 + ((w >> 17) & 63) * 8 * 32)
            per_tile_w += len(j["groups"]) * c.get("ext_copy", c["ext_alloc"]) * c["cin"] * 2
        reads += per_tile_r * wp[c["conv"]] / 128.0
        writes += per_tile_w * wp[c["conv"]] / 128.0
    return reads, writes


def i64_array(values):
    arr = (c_int64 * len(values))(*[int(v) for v in values])
    return arr


def f64_array(values):
    return (c_double * len(values))(*[float(v) for v in values])
