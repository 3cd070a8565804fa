"""Mirror of the reference's datasets.py: InferenceDataset (:72-93) and LadDataset (:23-68)."""
import numpy as np
import torch

from . import config as cfg


class InferenceDataset(torch.utils.data.Dataset):
    """Item i = feats[i:i+n_frames], right-padded with zero rows at the tail (one item per frame)."""

    def __init__(self, feats, n_frames=cfg.FEAT["num_samples"]) -> None:
        super().__init__()
        self.feats = feats
        self.n_frames = n_frames

    def __len__(self):
        return len(self.feats)

    def __getitem__(self, index):
        ret = self.feats[index:index + self.n_frames]
        if ret.shape[0] != cfg.FEAT["num_samples"]:
            pad_amount = cfg.FEAT["num_samples"] - ret.shape[0]
            ret = np.pad(ret, ((0, pad_amount), (0, 0)))
        return ret


class FeatureCut:
    """Minimal stand-in for a lhotse cut with precomputed features: one (100, 44) window and its label."""

    class _Sup:
        def __init__(self, is_laugh):
            self.custom = {"is_laugh": int(is_laugh)}

    def __init__(self, feats, is_laugh, cut_id=None):
        self.features = np.asarray(feats, dtype=np.float32)
        self.supervisions = [FeatureCut._Sup(is_laugh)]
        self.id = cut_id

    def load_features(self):
        return self.features


def _precomputed_features(cuts):
    feats = [np.asarray(c.load_features(), dtype=np.float32) for c in cuts]
    lens = torch.tensor([f.shape[0] for f in feats], dtype=torch.int32)
    return torch.from_numpy(np.stack(feats)), lens


class LadDataset(torch.utils.data.Dataset):
    """Laugh-activity-detection batches: ``__getitem__(cuts)`` -> {'inputs': (B,T,F) float32, 'input_lens': (B,),
    'is_laugh': (B,) int32, 'cut': cuts}.  ``input_strategy`` defaults to reading each cut's precomputed
    features (lhotse's PrecomputedFeatures when a lhotse CutSet is passed, else ``cut.load_features()``)."""

    def __init__(self, input_strategy=None, cut_transforms=None, input_transforms=None) -> None:
        super().__init__()
        self.input_strategy = input_strategy if input_strategy is not None else _precomputed_features
        self.cut_transforms = cut_transforms or []
        self.input_transforms = input_transforms or []

    def __getitem__(self, cuts):
        for tfnm in self.cut_transforms:
            cuts = tfnm(cuts)
        inputs, input_lens = self.input_strategy(cuts)
        for tfnm in self.input_transforms:
            inputs = tfnm(inputs)
        is_laugh = [c.supervisions[0].custom["is_laugh"] for c in cuts]
        return {"inputs": inputs, "input_lens": input_lens, "is_laugh": torch.tensor(is_laugh, dtype=torch.int32), "cut": cuts}
