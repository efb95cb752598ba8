"""ctypes loader for libmucon_b200.so (C ABI in include/mucon_b200.h).

There is no CPU fallback: if the library is missing or a call fails this raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmucon_b200.so")

MUCON_OK = 0
UNIT_OK, UNIT_INFEASIBLE, UNIT_SHORT, UNIT_NONFINITE = 0, 1, 2, 3

# every symbol include/mucon_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "mucon_abi_version", "mucon_strerror", "mucon_last_cuda_error", "mucon_device_sm_count", "mucon_set_sm_limit",
    "mucon_viterbi_blockscores", "mucon_viterbi_decode", "mucon_viterbi_decode_generic", "mucon_viterbi_pack_lanes_h", "mucon_viterbi_decode_lanes", "mucon_viterbi_pack_h", "mucon_viterbi_align_fused", "mucon_viterbi_align_fused_tail", "mucon_viterbi_align_fused_pooled", "mucon_viterbi_select", "mucon_viterbi_select_ranked", "mucon_viterbi_labels",
    "mucon_poisson_params_h", "mucon_logfact_h", "mucon_sgemm_bias", "mucon_lstm_encoder", "mucon_seq_decoder", "mucon_class_mean_params", "mucon_single_create", "mucon_single_destroy", "mucon_single_decode_h", "mucon_peer_alloc", "mucon_peer_open", "mucon_peer_close", "mucon_peer_free",
    "mucon_masks_fwd", "mucon_masks_fwd_ws", "mucon_masks_bwd", "mucon_flint_fwd", "mucon_flint_fwd_ws", "mucon_flint_fwd_ws_words", "mucon_flint_bwd", "mucon_mask_template_h",
    "mucon_gemm_tf32_bias_act", "mucon_gemm_tf32_bias_act_bf16", "mucon_wavenet_layer_bf16", "mucon_wavenet_layer_bf16_ex", "mucon_conv_gemm_tf32", "mucon_conv_gemm_tf32_shifts", "mucon_conv_gemm_tf32_ex", "mucon_wgrad_tf32", "mucon_maxpool2_bwd", "mucon_groupnorm_relu_bwd", "mucon_expand_rows_bwd", "mucon_wavenet_layer_tf32", "mucon_wavenet_layer_tf32_pair", "mucon_conv1d", "mucon_maxpool2", "mucon_pool2", "mucon_groupnorm_relu", "mucon_logsoftmax_expand", "mucon_logsoftmax_rows", "mucon_tail_logprobs", "mucon_expand_rows",
    "mucon_vit_mof", "mucon_vit_segment_metrics", "mucon_vit_segment_metrics_ws_words",
]


class MuconError(RuntimeError):
    pass


class ViterbiBatch(C.Structure):
    """mirror of struct mucon_viterbi_batch"""
    _fields_ = [
        ("U", C.c_int32), ("C", C.c_int32), ("fs", C.c_int32), ("max_len", C.c_int32),
        ("bs_is_f64", C.c_int32), ("seg0_f32", C.c_int32), ("max_N", C.c_int32), ("max_K", C.c_int32),
        ("n_cta", C.c_int32), ("wpc", C.c_int32), ("lanes", C.c_int32), ("n_peers", C.c_int32),
        ("bs", C.c_void_p), ("vid_off", C.c_void_p), ("blk_off", C.c_void_p), ("unit_vid", C.c_void_p),
        ("tr", C.c_void_p), ("tr_off", C.c_void_p), ("len_rows", C.c_void_p), ("len_params", C.c_void_p),
        ("logfact", C.c_void_p), ("lab_off", C.c_void_p), ("bp_off", C.c_void_p), ("warp_unit", C.c_void_p),
        ("score", C.c_void_p), ("labels", C.c_void_p), ("seg_blocks", C.c_void_p), ("bp", C.c_void_p),
        ("final_j", C.c_void_p), ("status", C.c_void_p), ("peer_delta", C.c_int64 * 8),
    ]


class SHeadWeights(C.Structure):
    """mirror of struct mucon_shead_weights"""
    _fields_ = [(n, C.c_void_p) for n in ("hid_w", "hid_b", "cn_w", "cn_b", "l2_w", "l2_b", "att_v", "emb", "comb_w",
                                          "comb_b", "wih", "whh", "bih", "bhh", "t1_w", "t1_b", "t2_w", "t2_b", "n1_w",
                                          "n1_b", "n2_w", "n2_b")]


_lib = None


def lib():
    """Loads the shared library once; raises MuconError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MuconError(
                f"{LIB_PATH} not found: build it with `python -m mucon_b200.build` "
                "(there is no CPU fallback for the sm_100a kernels)")
        l = C.CDLL(LIB_PATH)
        l.mucon_strerror.restype = C.c_char_p
        l.mucon_last_cuda_error.restype = C.c_char_p
        for name in SYMBOLS:
            if name in ("mucon_strerror", "mucon_last_cuda_error"):
                continue
            getattr(l, name).restype = C.c_int64 if name.endswith("_ws_words") else C.c_int
        _lib = l
    return _lib


def check(rc, what):
    if rc != MUCON_OK:
        l = lib()
        msg = l.mucon_strerror(rc).decode()
        if rc == -3:
            msg += ": " + l.mucon_last_cuda_error().decode()
        raise MuconError(f"{what} failed: {msg} ({rc})")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
