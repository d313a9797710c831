"""``det3d.datasets.pipelines`` names of the steps this package provides on the device
(det3d/datasets/pipelines/preprocess.py:276-463 ``Voxelization``, :479-653 ``AssignLabel``)."""
from sparse2dense_b200.pipeline import AssignLabel, Voxelization  # noqa: F401
