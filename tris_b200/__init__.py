"""tris_b200 -- B200-native (sm_100a) implementation of the TRIS Stage-1 hot path.

Drop-in for the reference's ``model.model_stage1.TRIS`` surface; all hot arithmetic runs in the
hand-written CUDA kernels of ``libtris_sm100.so`` (C ABI in include/tris_sm100.h).
"""
__version__ = "0.1.0"
