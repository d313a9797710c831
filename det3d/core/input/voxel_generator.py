from sparse2dense_b200.voxel_generator import VoxelGenerator  # noqa: F401
