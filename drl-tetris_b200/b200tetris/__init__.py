"""b200tetris — B200-native batched Tetris engine, drop-in for DRL-Tetris' environment/game_backend as driven by
environment/tetris_environment_vector.py.  See include/b200tetris.h (C ABI), DESIGN.md and INTEGRATION.md."""
from ._lib import B2TError, LIB_PATH, VEC_LEN, load  # noqa: F401
from .engine import Engine, host_array, piece_map  # noqa: F401
from .vector_env import action, action_list, null_action, state, tetris_environment_vector  # noqa: F401
from .actor import ActorLoop  # noqa: F401
from . import sharding  # noqa: F401
from . import tetris_env  # noqa: F401  (per-env facade: the reference's pybind11 module surface)
