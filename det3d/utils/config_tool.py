from sparse2dense_b200.config import get_downsample_factor  # noqa: F401
