"""safe_mpc_b200 -- B200-native batched real-time-iteration MPC engine (drop-in backend for the closed-loop
simulation of idra-lab/safe-mpc).  The compute path is the CUDA library in ``csrc/`` behind the C ABI of
``include/safe_mpc_b200.h``; this package is the Python host layer that mirrors the reference's
``src/safe_mpc`` classes (parser, env_model, controller, cost_definition, safe_set, utils)."""
__version__ = '0.1.0'
