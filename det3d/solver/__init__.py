"""``det3d.solver`` names used by the reference's training entry points (det3d/solver/learning_schedules_fastai.py:77-95
``OneCycle``; det3d/solver/fastai_optim.py:118-174 ``OptimWrapper`` -> ``FlatAdam``, the flat-buffer equivalent)."""
from sparse2dense_b200.trainer import FlatAdam, OneCycle  # noqa: F401
