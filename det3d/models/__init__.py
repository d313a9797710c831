from sparse2dense_b200.registry import (BACKBONES, DETECTORS, HEADS, LOSSES, NECKS, READERS, ROI_HEAD, SECOND_STAGE,  # noqa: F401
                                        build_backbone, build_detector, build_head, build_loss, build_neck,
                                        build_reader, build_roi_head, build_second_stage_module)
import sparse2dense_b200  # noqa: F401,E402  (populates the registries)
