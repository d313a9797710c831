from sparse2dense_b200.config import Config, ConfigDict  # noqa: F401
