from sparse2dense_b200.checkpoint import load_checkpoint, load_state_dict, save_checkpoint  # noqa: F401
