"""Mirror of the reference's load_data.py: ``create_inference_dataloader`` (:37-54) and
``create_training_dataloader`` (:12-34).  Features come from the fused B200 kernel (K1) instead of lhotse's CPU
Fbank; ``infer_audio_file`` is the fused fast path (PCM -> probabilities without materialising windows)."""
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import config as cfg
from . import engine as _engine
from .datasets import InferenceDataset, LadDataset
from .utils.audio_utils import load_wav_int16
from .utils.utils import get_feat_extractor


def compute_features_for_file(audio_path, device=0, mel="lhotse"):
    """Whole-file log-mel features as a CUDA tensor (T, 44): the ``cut.compute_features(extractor)`` call."""
    pcm, sr = load_wav_int16(audio_path)
    if sr != _engine.SAMPLE_RATE:
        raise ValueError(f"{audio_path}: sampling rate {sr} != {_engine.SAMPLE_RATE} (lhotse's Fbank asserts the same)")
    eng = _engine.get_engine(device)
    feats, _ = eng.fbank(torch.from_numpy(pcm).to(eng.device), mel=mel)
    return feats


def create_inference_dataloader(audio_path):
    """DataLoader over InferenceDataset(feats_all), batch_size=32: yields (<=32, 100, 44) float tensors in frame order."""
    get_feat_extractor(num_samples=cfg.FEAT['num_samples'], num_filters=cfg.FEAT['num_filters'])  # validates config
    feats_all = compute_features_for_file(audio_path).cpu().numpy()
    dataset = InferenceDataset(feats_all)
    return DataLoader(dataset, batch_size=32)


def infer_audio_file(audio_path, model, as_tensor=False):
    """Fused path: per-frame probabilities (numpy float32, length T) for one audio file; as_tensor=True leaves them on the
    GPU (a CUDA float32 tensor) so that the segmenter reads them without a host round trip."""
    feats = compute_features_for_file(audio_path, device=next(model.parameters()).device.index or 0)
    probs = model.infer_channel(feats)
    return probs if as_tensor else probs.cpu().numpy()


def create_training_dataloader(cutset_dir, split, shuffle=False):
    """Batches of 32 cuts as dicts {'inputs','input_lens','is_laugh','cut'} from '<split>_cutset_with_feats.jsonl'."""
    if split not in ['train', 'dev', 'test']:
        raise ValueError(
            f"Unexpected value for split. Needs to be one of 'train, dev, test'. Found {split}")
    store_manifest = os.path.join(cutset_dir, 'feats.jsonl')
    split_df = os.path.join(cutset_dir, f'{split}_df.csv')
    if os.path.exists(store_manifest) and os.path.exists(split_df):
        # lhotse-free path: raw GPU feature store + data frame (compute_features.FeatureStore)
        from . import compute_features as cf
        store = cf.FeatureStore.load(cutset_dir)
        cuts = cf.cuts_from_dataframe(cf.read_data_df(split_df), store, shuffle_seed=0 if shuffle else None)
        return list(cf.training_batches(cuts, max_cuts=32))
    try:
        from lhotse import CutSet
        from lhotse.dataset import SingleCutSampler
        from lhotse.dataset.input_strategies import PrecomputedFeatures
    except ImportError as e:
        raise ImportError("create_training_dataloader reads lhotse cut manifests; lhotse is not installed") from e
    cuts = CutSet.from_jsonl(os.path.join(cutset_dir, f'{split}_cutset_with_feats.jsonl'))
    if shuffle:
        cuts = cuts.shuffle()
    dataset = LadDataset(input_strategy=PrecomputedFeatures())
    sampler = SingleCutSampler(cuts, max_cuts=32)
    return DataLoader(dataset, sampler=sampler, batch_size=None)
