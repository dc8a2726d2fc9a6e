"""multigrid_b200: B200-native batched engine for MultiGrid's step/observe hot path.

    from multigrid_b200.envs import make
    env = make('MultiGrid-Empty-8x8-v0', agents=4, num_envs=65536)
    obs, infos = env.reset(seed=0)
    obs, rewards, terminations, truncations, infos = env.step({0: 2, 1: 0, 2: 5, 3: 6})

Layers: `envs` (registry) -> `env.BatchedMultiGridEnv` (the reference's MultiGridEnv surface,
batched) -> `engine.StepEngine` (HBM-resident state + launches) -> `_cabi` (ctypes over
include/multigrid_b200.h) -> csrc/ (sm_100a kernels). See DESIGN.md.
"""
__version__ = "0.1.0"
