from .utils import get_feat_extractor  # noqa: F401
