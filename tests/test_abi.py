"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/ld_b200.h declares, and its
host-side helpers (no GPU needed) agree with the reference arithmetic."""
import ctypes
import os
import re

import numpy as np
import scipy.signal

from laughter_detection_icsi_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "ld_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"^LD_API [^;(]*?\b(ld_[a-z0-9_]+)\(", text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ld_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == names, "ctypes binding and header disagree"
    assert b"sm_100a" in lib.ld_version()


def test_default_config_matches_reference_config_py():
    cfg = _native.default_config()
    assert cfg.struct_size == ctypes.sizeof(_native.LdConfig)
    assert (cfg.num_frames, cfg.num_filters, cfg.linear_layer_size) == (100, 44, 48)
    assert list(cfg.filter_sizes) == [64, 32, 16, 16]


def test_unsupported_config_is_rejected():
    lib = _native.load_library()
    cfg = _native.default_config(num_filters=40)
    assert lib.ld_plan_json(ctypes.byref(cfg), None, 0) < 0
    assert b"supported" in lib.ld_last_error()


def test_num_frames_helper():
    lib = _native.load_library()
    for n, t in ((399, 2), (400, 3), (16000, 100), (16037, 100), (160037, 1000), (9600000, 60000), (57600000, 360000)):
        assert lib.ld_fbank_num_frames(n) == t


def test_min_length_filter_rounds_like_python_floats():
    """end/fps - start/fps > min_l in float64: of the 200 placements of an exactly-20-frame span, 46 survive."""
    lib = _native.load_library()
    I32, F64 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    kept_total = 0
    for fps in (100.0, 3000 / 30.0037, 99.99983):
        starts = np.arange(0, 200, dtype=np.int32)
        ends = starts + 20
        os_, oe = np.empty(200), np.empty(200)
        kept = lib.ld_filter_min_length(starts.ctypes.data_as(I32), ends.ctypes.data_as(I32), 200, fps, 0.2,
                                        os_.ctypes.data_as(F64), oe.ctypes.data_as(F64))
        expect = [(s / fps, e / fps) for s, e in zip(starts.tolist(), ends.tolist()) if e / fps - s / fps > 0.2]
        assert kept == len(expect)
        assert list(zip(os_[:kept].tolist(), oe[:kept].tolist())) == expect
        if fps == 100.0:
            kept_total = kept
    assert kept_total == 46


def test_butter2_matches_scipy():
    lib = _native.load_library()
    for cutoff in (0.01, 0.05, 0.3):
        b, a = _native.f64_array([0, 0, 0]), _native.f64_array([0, 0, 0])
        lib.ld_butter2_lowpass(cutoff, b, a)
        rb, ra = scipy.signal.butter(2, cutoff, output="ba")
        assert np.allclose(list(b), rb, rtol=1e-12, atol=0) and np.allclose(list(a), ra, rtol=1e-12, atol=0)


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _native.load_library()
    h = ctypes.c_void_p()
    cfg = _native.default_config()
    rc = lib.ld_create(0, ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value and lib.ld_last_error()
    import pytest
    from laughter_detection_icsi_b200.engine import Engine
    with pytest.raises(_native.LdError):
        Engine(0)


def test_numpy_float64_filter_is_the_c_helper():
    """pipeline.instances vectorises the min-length filter in NumPy float64; it must agree bit for bit with
    ld_filter_min_length (and through it with the reference's Python floats, see the segmenter golden cases)."""
    lib = _native.load_library()
    rng = np.random.default_rng(0)
    I32, F64 = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    for fps in (100.0, 99.98766, 100.0123, 360000 / 3600.00123):
        s = rng.integers(0, 300000, 5000).astype(np.int32)
        e = (s + rng.integers(0, 40, 5000)).astype(np.int32)
        for ml in (0.0, 0.1, 0.2):
            os_, oe = np.empty(len(s)), np.empty(len(s))
            k = lib.ld_filter_min_length(s.ctypes.data_as(I32), e.ctypes.data_as(I32), len(s), fps, ml, os_.ctypes.data_as(F64),
                                         oe.ctypes.data_as(F64))
            ss, ee = s.astype(np.float64) / fps, e.astype(np.float64) / fps
            m = (ee - ss) > ml
            assert k == m.sum() and np.array_equal(os_[:k], ss[m]) and np.array_equal(oe[:k], ee[m])
