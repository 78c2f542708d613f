"""ibamr_b200 -- B200-native Lagrangian-Eulerian interaction (IB spread / interpolate) for IBAMR.

csrc/      CUDA kernels (sm_100a) + the C ABI of include/ibk.h, built into libibk.so
api.py     host-side mirror of the reference interface (LEInteractor, IBMethod-shaped level)
halo.py    multi-process halo exchange (torch.distributed / NCCL) around the device pack kernels
"""
__all__ = ["api", "build"]
