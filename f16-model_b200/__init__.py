"""f16-model_b200 -- B200-native batched F-16 longitudinal simulator.

A drop-in for the per-step NumPy path of lalapopa/F16-model (``F16model/model`` +
``F16model/env``): one fused sm_100a CUDA kernel advances N independent aircraft per
launch behind the reference's gym-style env interface.  The CUDA library
(``lib/libf16b200.so``, C ABI in ``include/f16_b200.h``) is mandatory: there is no CPU
fallback and importing the compute entry points without it raises.

    from f16_model_b200 import F16VecEnv, F16, get_trimmed_state_control
"""
from . import plane                                           # noqa: F401
from .task import ReferenceSignal, build_reference_bank        # noqa: F401
from .model import States, Control, F16model, rhs, model_step  # noqa: F401
from .env import (F16, F16VecEnv, get_trimmed_state_control, get_random_state, get_action,  # noqa: F401
                  find_correct_thrust_position)
from .policy import TensorCorePolicy, pack_policy_blob, pack_train_blob16   # noqa: F401
from .ppo import PPOConfig, PPOTrainer, gae, fused_rollout      # noqa: F401
from .trim import trim_grid, reference_starts                   # noqa: F401
from . import trim as find_trim_position                        # noqa: F401
from . import artifacts                                         # noqa: F401
from ._capi import build, load_library, library_path, F16Error   # noqa: F401

__all__ = ["F16", "F16VecEnv", "ReferenceSignal", "build_reference_bank", "States", "Control", "F16model", "rhs",
           "model_step", "get_trimmed_state_control", "get_random_state", "get_action", "find_correct_thrust_position", "plane",
           "TensorCorePolicy", "pack_policy_blob", "pack_train_blob16", "PPOConfig", "PPOTrainer", "gae", "fused_rollout", "trim_grid", "reference_starts", "find_trim_position", "artifacts", "build", "load_library", "library_path", "F16Error"]
