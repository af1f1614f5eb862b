"""ctypes binding of declip_b200/_C.so (the C ABI declared in include/declip_b200.h).

There is NO fallback: if the shared library is missing, or the device is not an sm_100a part,
every entry point raises.  Nothing under `oracle/` is ever imported from here.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("DECLIP_B200_LIB") or os.path.join(_HERE, "_C.so")     # override: A/B builds of the same ABI

_lib = None
_lock = threading.Lock()
_inited_devices = set()
_missing = []

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
c_ll = ctypes.c_longlong
c_ull = ctypes.c_ulonglong


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int), ("a_mn_major", c_int),
        ("B", c_void_p), ("ldb", c_int), ("b_mn_major", c_int),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("epilogue", c_int), ("alpha", c_float),
        ("out", c_void_p), ("ldo", c_int),
        ("out2", c_void_p), ("ldo2", c_int),
        ("bias", c_void_p),
        ("aux", c_void_p), ("ldaux", c_int),
        ("splits", c_int), ("block_n", c_int),
        ("colsum", c_void_p),
        ("alpha_dev", c_void_p),
    ]


class TowerCfg(ctypes.Structure):
    _fields_ = [
        ("layers", c_int), ("width", c_int), ("heads", c_int), ("seq_len", c_int), ("causal", c_int),
        ("batch", c_int), ("embed_dim", c_int), ("res", c_int), ("patch", c_int), ("vocab", c_int),
    ]


class HeadArgs(ctypes.Structure):
    _fields_ = [("b", c_int), ("n", c_int), ("e", c_int), ("ld", c_int), ("n_src", c_int), ("row0", c_int),
                ("x_off", c_int * 2), ("y_off", c_int * 2), ("cross", c_int), ("ld_strip", c_int),
                ("x_base", c_void_p), ("y_src", c_void_p * 8), ("ws", c_void_p),
                ("strips", c_void_p * 2)]


class AdamWEntry(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("shadow", c_void_p), ("numel", c_ull), ("group", c_int), ("reserved", c_int)]


class CastEntry(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("numel", c_ull)]


# name -> (restype, argtypes); every symbol include/declip_b200.h declares.
SIGNATURES = {
    "dc_version": (c_int, []),
    "dc_last_error": (ctypes.c_char_p, []),
    "dc_init": (c_int, [c_int]),
    "dc_sm_count": (c_int, []),
    "dc_set_sm_reserve": (c_int, [c_int]),
    "dc_cast_bf16_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_float, c_int, c_void_p]),
    "dc_set_backward_progress_cb": (c_int, [c_void_p, c_void_p, c_int]),
    "dc_launch_count": (c_ll, []),
    "dc_gemm_bf16": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "dc_set_gemm_2cta": (c_int, [c_int]),
    "dc_gemm_choose_splits": (c_int, [c_int, c_int, c_int]),
    "dc_set_attention_tc": (None, [c_int]),
    "dc_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                 c_void_p]),
    "dc_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_colsum_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dc_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "dc_multi_cast_f32_bf16": (c_int, [c_void_p, c_int, c_ull, c_void_p]),
    "dc_adamw_multi": (c_int, [c_void_p, c_int, c_ull, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_int,
                              c_void_p]),
    "dc_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "dc_patchify": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_vit_assemble": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_text_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_text_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_scatter_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_eot_index": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_l2norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "dc_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "dc_ce_strip_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "dc_ce_strip_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p, c_int, c_int, c_void_p]),
    "dc_batchnorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_float, c_float, c_int, c_int, c_void_p]),
    "dc_batchnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_cosine_rows_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_cosine_rows_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_void_p]),
    "dc_argmax_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dc_gather_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_add_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_dot_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "dc_bpe_create": (c_void_p, [ctypes.c_char_p, c_ll]),
    "dc_bpe_destroy": (None, [c_void_p]),
    "dc_bpe_vocab_size": (c_int, [c_void_p]),
    "dc_bpe_token_id": (c_int, [c_void_p, ctypes.c_char_p]),
    "dc_bpe_encode": (c_ll, [c_void_p, ctypes.c_char_p, c_void_p, c_ll]),
    "dc_bpe_tokenize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int]),
    "dc_conv3x3_igemm_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "dc_conv3x3_igemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_conv3x3_wgrad_igemm_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "dc_conv3x3_wgrad_igemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_head_workspace_floats": (c_size_t, [c_int, c_int]),
    "dc_head_layout": (c_int, [c_int, c_int, c_void_p]),
    "dc_head_prepare": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "dc_head_prepare_push": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_ll, c_void_p, c_void_p,
                                     c_float, c_void_p]),
    "dc_head_forward": (c_int, [ctypes.POINTER(HeadArgs), c_void_p]),
    "dc_head_backward": (c_int, [ctypes.POINTER(HeadArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dc_bpe_tokenize_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int]),
    "dc_tower_workspace_bytes": (c_size_t, [ctypes.POINTER(TowerCfg)]),
    "dc_vit_forward": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "dc_vit_backward": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "dc_tower_pre_features": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_void_p, c_void_p]),
    "dc_vit_backward_pre": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "dc_token_scores": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dc_groupmax_mean_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "dc_groupmax_mean_bwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dc_groupmax_mean_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dc_add_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dc_im2col_stem": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_im2col3x3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_col2im3x3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_avgpool2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dc_bn2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_ll, c_int, c_float, c_float, c_int, c_int, c_void_p]),
    "dc_bn2d_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "dc_add_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dc_attnpool_assemble": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_attnpool_assemble_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dc_text_forward": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "dc_text_backward": (c_int, [ctypes.POINTER(TowerCfg), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
}


def so_path():
    return _SO


def load():
    """dlopen the library and attach the signatures.  Does not touch the GPU."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_SO):
            raise RuntimeError(
                "declip_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % _SO)
        lib = ctypes.CDLL(_SO)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                _missing.append(name)  # tests/test_abi.py asserts this list is empty
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def missing_symbols():
    load()
    return list(_missing)


def last_error():
    return load().dc_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("declip_b200 %s failed (rc=%d): %s" % (what, rc, last_error()))


def init(device_index):
    """dc_init once per device (checks compute capability 10.x; raises otherwise)."""
    lib = load()
    if device_index not in _inited_devices:
        check(lib.dc_init(int(device_index)), "dc_init")
        _inited_devices.add(device_index)
        if os.environ.get("DC_GEMM_2CTA") is not None:
            lib.dc_set_gemm_2cta(int(os.environ["DC_GEMM_2CTA"] != "0"))
    return lib


def set_attention_tc(enable):
    """Route dc_attention_fwd/bwd to the tcgen05 core (True, default) or the mma.sync core (False)."""
    load().dc_set_attention_tc(int(bool(enable)))


def set_gemm_2cta(enable):
    """Select the cta_group::2 GEMM (256x256 cluster tiles) where the problem is large enough; returns the old value."""
    return int(load().dc_set_gemm_2cta(int(bool(enable))))


def launch_count():
    return int(load().dc_launch_count())
