from .metrics import (compute_cd, compute_overlap_ratio, compute_rigidity_rmse, compute_transform_errors, compute_transform_errors_direct,  # noqa: F401
                      fit_and_rigidity_sse, select_generation, trajectory_rigidity_rmse)
from .transforms_io import (apply_transform, format_transform, load_transform, relative_to_reference, relative_transform_matrix,  # noqa: F401
                            save_transformation_files)
