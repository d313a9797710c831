"""sparse2dense_b200 -- B200 (sm_100a) implementation of the Sparse2Dense / CenterPoint hot path
(voxelize -> SpMiddleResNetFHD sparse 3-D conv backbone -> BEV) behind the reference's own
operator API.  See DESIGN.md and include/s2d_b200.h."""
from . import registry  # noqa: F401
from .registry import (BACKBONES, DETECTORS, HEADS, NECKS, READERS, build_backbone, build_detector,  # noqa: F401
                       build_from_cfg, build_head, build_neck, build_reader)
from . import readers, backbones, necks, bbox_heads, detectors, second_stage  # noqa: F401,E402  (populate the registries)
from .config import Config, ConfigDict, get_downsample_factor  # noqa: F401,E402
