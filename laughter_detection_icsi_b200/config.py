"""Mirror of the reference's config.py (:7-31): MODEL_MAP and FEAT with the same keys and values.
``ANALYSIS`` (config.py:33-63) configures the transcript-analysis tooling, which is outside the hot path."""
from . import models

MODEL_MAP = {
    # the only configuration that fits the ICSI feature shape (100 x 44)  -- config.py:9-17
    "resnet_base": {
        "batch_size": 32,
        "model": models.ResNetBigger,
        "val_data_text_path": "./data/switchboard/val/switchboard_val_data.txt",
        "log_frequency": 900,
        "linear_layer_size": 48,
        "filter_sizes": [64, 32, 16, 16],
    },
    # kept for surface compatibility; sized for 128 x 44 inputs and fails on 100 x 44 in the reference too
    # (config.py:19-26)
    "resnet_with_augmentation": {
        "batch_size": 32,
        "model": models.ResNetBigger,
        "val_data_text_path": "./data/switchboard/val/switchboard_val_data.txt",
        "log_frequency": 200,
        "linear_layer_size": 128,
        "filter_sizes": [128, 64, 32, 32],
    },
}

FEAT = {"num_samples": 100, "num_filters": 44}
