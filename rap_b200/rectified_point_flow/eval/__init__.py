from .metrics import compute_rigidity_rmse, fit_and_rigidity_sse, select_generation, trajectory_rigidity_rmse  # noqa: F401
from .transforms_io import (apply_transform, format_transform, load_transform, relative_to_reference, relative_transform_matrix,  # noqa: F401
                            save_transformation_files)
