"""drone_rl_b200 -- B200-native (sm_100a) batched quadcopter environment + PPO hot path with the
Gym/SB3 surface of henryplas/drone_rl.  The compute lives in libdronecu.so (hand-written CUDA
behind the C ABI in include/dronecu.h); this package is the thin host side."""
from .core import DroneBatch, EnvConfig  # noqa: F401
from .envs import DroneGymEnv, DroneGymnasiumEnv, DroneVecEnv, VectorizedDroneGymEnv, make_sb3_vec_env  # noqa: F401
from ._lib import DronecuError  # noqa: F401

__all__ = ["DroneBatch", "EnvConfig", "DroneGymEnv", "DroneGymnasiumEnv", "DroneVecEnv", "make_sb3_vec_env",
           "VectorizedDroneGymEnv", "DronecuError"]
