"""Drop-in mirror of the reference's `rectified_point_flow` hot-path modules:
`sampler`, `procrustes`, `flow_model.PointCloudDiT`, `utils.point_clouds` -- same names, signatures,
return shapes and error behaviour, computed by the sm_100a library behind include/rap_b200.h."""
from . import procrustes, sampler  # noqa: F401
from .flow_model import PointCloudDiT  # noqa: F401
