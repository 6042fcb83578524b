"""ctypes binding of libzero_b200.so (include/zero_b200.h).  PyTorch tensors are only the buffer carrier:
every wrapper passes raw device pointers + the current CUDA stream through the C ABI.

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZB_LIB_PATH: load a differently-built copy of the same library (A/B experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("ZB_LIB_PATH") or os.path.join(_HERE, "libzero_b200.so")

ZB_BF16, ZB_F32 = 0, 1
ZB_K_MAJOR, ZB_MN_MAJOR = 0, 1
ZB_EPI_BIAS, ZB_EPI_RELU, ZB_EPI_ACCUM, ZB_EPI_RELU_MASK = 1, 2, 4, 8

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("d", vp), ("m", i64), ("n", i64), ("k", i64),
                ("lda", i64), ("ldb", i64), ("ldd", i64), ("a_layout", i32), ("b_layout", i32),
                ("d_dtype", i32), ("flags", i32), ("bias", vp), ("mask", vp), ("ldmask", i64),
                ("alpha", f32), ("split_k", i32)]


class ColsumArgs(C.Structure):
    _fields_ = [("x", vp), ("m", i64), ("n", i64), ("ld", i64), ("out", vp)]


class AttentionArgs(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp),
                ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64),
                ("bsq", i64), ("bsk", i64), ("bsv", i64), ("bso", i64),
                ("batch", i32), ("heads", i32), ("lq", i32), ("lk", i32), ("dh", i32),
                ("key_len", vp), ("causal", i32), ("q_offset", i32), ("scale", f32), ("inf_value", f32),
                ("lse", vp), ("rpr_k", vp), ("rpr_v", vp), ("max_rel", i32), ("relu_attn", i32),
                ("d_o", vp), ("dq", vp), ("dk", vp), ("dv", vp),
                ("lddo", i64), ("lddq", i64), ("lddk", i64), ("lddv", i64),
                ("bsdo", i64), ("bsdq", i64), ("bsdk", i64), ("bsdv", i64),
                ("d_rpr_k", vp), ("d_rpr_v", vp), ("delta", vp), ("kv_group", i32),
                ("dropout_rate", f32), ("dropout_site", C.c_uint32), ("dropout_seed", vp),
                ("workspace", vp), ("workspace_bytes", i64)]


class AddLnArgs(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("out", vp), ("mean", vp), ("rstd", vp), ("scale", vp), ("offset", vp),
                ("rows", i64), ("cols", i64), ("eps", f32),
                ("d_out", vp), ("d_out2", vp), ("ds", vp), ("dscale", vp), ("doffset", vp), ("dbias", vp),
                ("y32", vp), ("ybias", vp)]


class EmbedArgs(C.Structure):
    _fields_ = [("ids", vp), ("table", vp), ("bias", vp), ("out", vp),
                ("batch", i32), ("len", i32), ("dim", i32), ("vocab", i32),
                ("shift", i32), ("zero_if_all_pad", i32), ("time", i32), ("mult", f32),
                ("d_out", vp), ("d_table", vp), ("d_bias", vp), ("d_out2", vp)]


class CeArgs(C.Structure):
    _fields_ = [("logits", vp), ("ld", i64), ("labels", vp), ("batch", i32), ("seq_len", i32), ("nll", vp),
                ("d_logits", vp), ("ldd", i64), ("vocab", i32), ("smooth", f32), ("loss_scale", f32),
                ("per_sample", vp), ("loss", vp)]


class VocabCeArgs(C.Structure):
    _fields_ = [("feat", vp), ("ldf", i64), ("table", vp), ("ldt", i64), ("labels", vp),
                ("batch", i32), ("seq_len", i32), ("d", i32), ("vocab", i32), ("smooth", f32), ("loss_scale", f32),
                ("nll", vp), ("per_sample", vp), ("loss", vp), ("d_logits", vp), ("ldd", i64),
                ("workspace", vp), ("workspace_bytes", i64)]


class AdamArgs(C.Structure):
    _fields_ = [("param", vp), ("m", vp), ("v", vp), ("grad", vp), ("param_bf16", vp), ("n", i64),
                ("beta1", f32), ("beta2", f32), ("eps", f32), ("lr_t", f32), ("grad_scale", f32),
                ("clip_scale", vp), ("norms", vp)]


class BeamArgs(C.Structure):
    _fields_ = [("logits", vp), ("batch", i32), ("beam", i32), ("vocab", i32), ("time", i32),
                ("eos_id", i32), ("pad_id", i32), ("temperature", f32), ("inf_value", f32),
                ("length_penalty", f32), ("max_len", vp), ("max_penalty", vp), ("seq_cap", i32),
                ("alive_seq", vp), ("alive_logp", vp), ("alive_score", vp),
                ("fin_seq", vp), ("fin_score", vp), ("fin_flag", vp), ("parent", vp), ("tmp_seq", vp),
                ("active", vp), ("row_ws", vp), ("cand", vp)]


class VocabTopkArgs(C.Structure):
    _fields_ = [("feat", vp), ("ldf", i64), ("table", vp), ("ldt", i64), ("rows", i32), ("d", i32), ("vocab", i32),
                ("skip_col", i32), ("temperature", f32), ("workspace", vp), ("workspace_bytes", i64)]


SHARD_MAX_WORLD = 16
ZB_SHARD_UPDATE, ZB_SHARD_STORE_GRAD, ZB_SHARD_NORM_G, ZB_SHARD_NORM_P = 1, 2, 4, 8


class ShardAdamArgs(C.Structure):
    _fields_ = [("lo", i64), ("n", i64), ("world", i32), ("rank", i32), ("grad_sources", i32), ("flags", i32),
                ("grad_mc", vp), ("grad_peer", vp * SHARD_MAX_WORLD),
                ("param", vp), ("m", vp), ("v", vp),
                ("wide_mask", vp), ("param_mc", vp), ("param_peer", vp * SHARD_MAX_WORLD),
                ("mirror_mc", vp), ("mirror_peer", vp * SHARD_MAX_WORLD), ("grad_out", vp),
                ("beta1", f32), ("beta2", f32), ("eps", f32), ("lr_t", f32), ("grad_scale", f32),
                ("clip_scale", vp), ("norms", vp), ("norm_parts_peer", vp * SHARD_MAX_WORLD), ("done_counter", vp)]


# every symbol include/zero_b200.h declares (tests/test_abi.py checks the library exports all of them)
# order = the index zb_abi_struct_size() understands
STRUCTS = [GemmArgs, AttentionArgs, AddLnArgs, EmbedArgs, CeArgs, AdamArgs, BeamArgs, ColsumArgs, ShardAdamArgs,
           VocabCeArgs, VocabTopkArgs]

EXPORTS = [
    "zb_abi_version", "zb_last_error_string", "zb_launch_count", "zb_path_launch_count", "zb_abi_struct_size", "zb_dropout", "zb_gemm", "zb_attention_fwd",
    "zb_attention_bwd", "zb_add_ln_fwd", "zb_add_ln_bwd", "zb_embed_fwd", "zb_embed_bwd", "zb_softmax_ce",
    "zb_colsum", "zb_cast_f32_bf16", "zb_cast_bf16_f32", "zb_adam_tf", "zb_sumsq", "zb_beam_cond",
    "zb_beam_step", "zb_gather_rows", "zb_prefix_mean_fwd", "zb_prefix_mean_bwd", "zb_aan_step",
    "zb_aan_gate_fwd", "zb_aan_gate_bwd", "zb_gated_rms_fwd", "zb_gated_rms_bwd", "zb_add2d",
    "zb_gemm_grouped", "zb_colsum_grouped", "zb_aan_cat_step", "zb_aan_gate_ln", "zb_shard_adam", "zb_gumbel_add",
    "zb_attention_bwd_workspace_bytes", "zb_vocab_ce", "zb_vocab_ce_workspace_bytes", "zb_set_sm_reserve",
    "zb_vocab_topk", "zb_vocab_topk_workspace_bytes", "zb_vocab_topk_parts",
]

_lib = None


class ZeroB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZeroB200Error(
            "libzero_b200.so not found at %s — build it with `python -m zero_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.zb_abi_version.restype = C.c_int
    lib.zb_last_error_string.restype = C.c_char_p
    lib.zb_launch_count.restype = C.c_int64
    lib.zb_path_launch_count.argtypes = [i32]
    lib.zb_path_launch_count.restype = C.c_int64
    for name, argt in [
        ("zb_gemm", [C.POINTER(GemmArgs), vp]),
        ("zb_gemm_grouped", [C.POINTER(GemmArgs), i32, vp]),
        ("zb_colsum_grouped", [C.POINTER(ColsumArgs), i32, vp]),
        ("zb_attention_fwd", [C.POINTER(AttentionArgs), vp]),
        ("zb_attention_bwd", [C.POINTER(AttentionArgs), vp]),
        ("zb_add_ln_fwd", [C.POINTER(AddLnArgs), vp]),
        ("zb_add_ln_bwd", [C.POINTER(AddLnArgs), vp]),
        ("zb_embed_fwd", [C.POINTER(EmbedArgs), vp]),
        ("zb_embed_bwd", [C.POINTER(EmbedArgs), vp]),
        ("zb_softmax_ce", [C.POINTER(CeArgs), vp]),
        ("zb_vocab_ce", [C.POINTER(VocabCeArgs), vp]),
        ("zb_vocab_topk", [C.POINTER(VocabTopkArgs), vp]),
        ("zb_set_sm_reserve", [i32]),
        ("zb_colsum", [vp, i64, i64, i64, vp, vp]),
        ("zb_cast_f32_bf16", [vp, vp, i64, vp]),
        ("zb_cast_bf16_f32", [vp, vp, i64, vp]),
        ("zb_adam_tf", [C.POINTER(AdamArgs), vp]),
        ("zb_shard_adam", [C.POINTER(ShardAdamArgs), vp]),
        ("zb_sumsq", [vp, i64, vp, vp]),
        ("zb_beam_cond", [C.POINTER(BeamArgs), vp]),
        ("zb_beam_step", [C.POINTER(BeamArgs), vp]),
        ("zb_gather_rows", [vp, vp, vp, i64, i64, i64, vp]),
        ("zb_prefix_mean_fwd", [vp, vp, vp, i32, i32, i32, i32, vp]),
        ("zb_prefix_mean_bwd", [vp, vp, vp, i32, i32, i32, i32, vp]),
        ("zb_aan_step", [vp, vp, vp, i64, i32, vp]),
        ("zb_aan_cat_step", [vp, vp, vp, i64, vp, i64, i32, i32, vp]),
        ("zb_aan_gate_ln", [vp, vp, vp, vp, vp, vp, i64, i32, f32, vp]),
        ("zb_aan_gate_fwd", [vp, vp, vp, vp, i64, i32, vp]),
        ("zb_aan_gate_bwd", [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]),
        ("zb_gated_rms_fwd", [vp, vp, vp, vp, vp, i64, i64, f32, vp]),
        ("zb_gated_rms_bwd", [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]),
        ("zb_add2d", [vp, i64, vp, i64, vp, i64, i64, i64, vp]),
        ("zb_dropout", [vp, vp, vp, i64, f32, vp, C.c_uint32, vp]),
        ("zb_gumbel_add", [vp, i64, f32, vp, C.c_uint32, vp]),
    ]:
        fn = getattr(lib, name)
        fn.argtypes = argt
        fn.restype = C.c_int
    lib.zb_attention_bwd_workspace_bytes.argtypes = [C.POINTER(AttentionArgs)]
    lib.zb_attention_bwd_workspace_bytes.restype = C.c_int64
    lib.zb_vocab_ce_workspace_bytes.argtypes = [C.POINTER(VocabCeArgs)]
    lib.zb_vocab_ce_workspace_bytes.restype = C.c_int64
    lib.zb_vocab_topk_workspace_bytes.argtypes = [C.POINTER(VocabTopkArgs)]
    lib.zb_vocab_topk_workspace_bytes.restype = C.c_int64
    lib.zb_vocab_topk_parts.argtypes = [i32]
    lib.zb_vocab_topk_parts.restype = i32
    if lib.zb_abi_version() != 4:
        raise ZeroB200Error("libzero_b200.so ABI version mismatch")
    lib.zb_abi_struct_size.argtypes = [i32]
    lib.zb_abi_struct_size.restype = C.c_int64
    for which, cls in enumerate(STRUCTS):
        if lib.zb_abi_struct_size(which) != C.sizeof(cls):
            raise ZeroB200Error("libzero_b200.so: %s is %d bytes in the library, %d in the binding (stale build?)" % (
                cls.__name__, lib.zb_abi_struct_size(which), C.sizeof(cls)))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().zb_last_error_string()
        raise ZeroB200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


PATHS = {"gemm_tcgen05": 0, "gemm_pair": 1, "gemm_skinny": 2, "attn_mma": 3, "attn_generic": 4, "attn_decode": 5,
         "beam_sentence": 6, "beam_rows": 7, "beam_parts": 8, "gemm_bm64": 9, "attn_tc": 10, "beam_cand": 11}


def path_launch_count(name) -> int:
    """Launches served by one dispatch path (include/zero_b200.h zb_path)."""
    return int(load().zb_path_launch_count(PATHS[name]))


def launch_count() -> int:
    return int(load().zb_launch_count())
