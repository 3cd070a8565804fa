"""B200-native laughter-detection hot path: drop-in mirror of the reference's Python surface
(models / config / load_data / datasets / laugh_segmenter / segment_laughter / utils) on top of
hand-written sm_100a CUDA kernels reached through the C ABI in include/ld_b200.h.

    from laughter_detection_icsi_b200 import models, config, laugh_segmenter, load_data

There is no CPU fallback: every compute path raises if the CUDA library or a B200 is missing.
"""
__version__ = "0.1.0"
