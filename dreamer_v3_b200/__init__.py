"""dreamer_v3_b200 -- B200-native (sm_100a) engine for the DreamerV3 training hot path behind the
Python API of sven1977/dreamer_v3. All compute runs in libdv3.so (hand-written CUDA); see DESIGN.md."""
from .spec import Box, Discrete, EngineConfig, init_params, param_spec  # noqa: F401

__all__ = ["Box", "Discrete", "EngineConfig", "init_params", "param_spec"]
