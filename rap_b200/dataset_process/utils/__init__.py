"""Mirror of the keypoint-selection helpers of the reference's `dataset_process.utils` that sit in front of the extractor."""
from . import dataset_utils  # noqa: F401
from .point_sampling_utils import (allocate_fps_points, apply_batched_fps, calculate_adaptive_sample_count_per_part,  # noqa: F401
                                   calculate_voxel_coverage)
