"""declip_b200 — B200-native (sm_100a) CLIP/DeCLIP dual-encoder training path.

Hand-written CUDA kernels (tcgen05/TMEM/TMA GEMMs, fused row kernels, attention, contrastive
head) behind a C ABI (`include/declip_b200.h`), mirrored on the Python side by modules with the
reference's `prototype.model` factory names, forward signatures and state_dict keys.
"""
__version__ = "0.1.0"
