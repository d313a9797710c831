from sparse2dense_b200.checkpoint import load_checkpoint, load_state_dict, save_checkpoint  # noqa: F401
from sparse2dense_b200.trainer import DistillTrainer  # noqa: F401,E402  (TS_Trainer.batch_processor_inline + OptimizerHook)
