"""Mirror of the reference's `dataset_process.extract_sample_features.FeatureExtractor` on the sm_100a extractor kernels."""
from .extract_sample_features import FeatureExtractor  # noqa: F401
