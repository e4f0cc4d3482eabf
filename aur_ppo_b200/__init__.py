"""aur_ppo_b200: the PPO hot path of biirving/aur_ppo (rollout step -> GAE ->
minibatch update) as hand-written sm_100a CUDA behind the reference's Python API.

Layout: csrc/ (kernels + C ABI -> libaurppo.so), _lib.py (ctypes binding),
kernels.py (tensor-level wrappers), ppo.py / run_ppo.py / models / nets (the
host-side mirror of the reference interface)."""
__version__ = "0.1.0"
