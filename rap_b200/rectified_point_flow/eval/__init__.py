from .metrics import compute_rigidity_rmse  # noqa: F401
