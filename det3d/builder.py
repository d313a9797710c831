"""``det3d.builder`` names the reference configs import (configs/waymo/voxelnet/waymo_second_*.py:4)."""
from sparse2dense_b200.anchors import build_anchor_generator, build_box_coder  # noqa: F401
