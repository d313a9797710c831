from sparse2dense_b200.registry import Registry, build_from_cfg  # noqa: F401
