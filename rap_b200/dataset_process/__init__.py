"""Mirror of the reference's `dataset_process` modules on the hot path: the local feature extractor and the keypoint selection in
front of it (voxel down-sampling, outlier removal, keypoint budgets, batched FPS)."""
from .extract_sample_features import FeatureExtractor, SampleProcessor, remove_statistical_outlier  # noqa: F401
