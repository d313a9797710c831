"""Import-compatibility package: the names reference configs, tools and user code import from ``det3d`` resolve to the
B200 implementation in ``sparse2dense_b200`` (configs/waymo/**/*.py import ``det3d.utils.config_tool``; tools import
``det3d.torchie.Config`` and ``det3d.models.build_detector``).  Nothing here contains arithmetic."""
