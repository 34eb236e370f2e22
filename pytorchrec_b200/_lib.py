"""ctypes binding of ``libptrec_b200.so`` (the C ABI declared in ``include/ptrec_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C pytorchrec_b200/csrc``) every
product entry point raises.  The reference has no FFI of its own (its hot path is
``nn.Embedding`` + autograd + ``torch.optim``, torchrec/model/IModel.py:116-125); this file is the
binding a reference maintainer would add.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptrec_b200.so")

ABI_VERSION = 26

# enums (mirror include/ptrec_b200.h)
F32, BF16 = 0, 1
POOL_SUM, POOL_MEAN, POOL_SQRTN = 0, 1, 2
MASK_NONE, MASK_PAD, MASK_PAD_KEEP_FIRST, MASK_LENS = 0, 1, 2, 3
OPT_SGD, OPT_ADAGRAD, OPT_ROWWISE_ADAGRAD, OPT_LAZY_ADAM = 0, 1, 2, 3
FEAT_NEG_IS_PAD = 1

POOLING_NAMES = {"sum": POOL_SUM, "mean": POOL_MEAN, "sqrtn": POOL_SQRTN}
MASK_NAMES = {"none": MASK_NONE, "pad": MASK_PAD, "pad_keep_first": MASK_PAD_KEEP_FIRST, "lens": MASK_LENS}


class FeatureDesc(Structure):
    _fields_ = [
        ("table", c_int32),
        ("bag_len", c_int32),
        ("pooling", c_int32),
        ("mask_mode", c_int32),
        ("lens_col", c_int32),
        ("flags", c_int32),
        ("id_base", c_int64),
        ("out_col", c_int64),
    ]


class OptimArgs(Structure):
    _fields_ = [
        ("kind", c_int32),
        ("step", c_int32),
        ("lr", c_float),
        ("eps", c_float),
        ("beta1", c_float),
        ("beta2", c_float),
        ("weight_decay", c_float),
        ("lr_decay", c_float),
    ]


assert ctypes.sizeof(FeatureDesc) == 40
assert ctypes.sizeof(OptimArgs) == 32

_FD = POINTER(FeatureDesc)
_BWD_ARGS = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, _FD, _FD, c_int32, c_int64,
             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
             POINTER(OptimArgs), c_void_p, c_size_t, c_void_p]

# name -> (restype, argtypes); every symbol include/ptrec_b200.h declares
PROTOTYPES = {
    "ptrec_abi_version": (c_int, []),
    "ptrec_last_error": (c_char_p, []),
    "ptrec_launch_count": (c_int64, []),
    "ptrec_set_l2_fetch_granularity": (c_int, [c_int32]),
    "ptrec_get_l2_fetch_granularity": (c_int, []),
    "ptrec_index_prep_workspace_bytes": (c_size_t, [c_int64]),
    "ptrec_index_prep": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "ptrec_embedding_gather_pool_fwd": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, _FD, _FD,
                                                c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                                c_void_p, c_void_p, c_void_p]),
    "ptrec_embedding_gather_pool_fwd_sharded": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64,
                                                        c_int32, _FD, _FD, c_int32, c_void_p, c_int64, c_void_p,
                                                        c_int64, c_void_p, c_void_p]),
    "ptrec_set_smem_sort": (None, [c_int32]),
    "ptrec_smem_sort_enabled": (c_int32, []),
    "ptrec_set_one_sweep_sort": (None, [c_int32]),
    "ptrec_one_sweep_sort_enabled": (c_int32, []),
    "ptrec_sort_dedup_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "ptrec_sort_dedup": (c_int, [_FD, _FD, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "ptrec_embedding_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "ptrec_embedding_bwd_fused": (c_int, _BWD_ARGS),
    "ptrec_embedding_bwd_fused_sgd": (c_int, _BWD_ARGS),
    "ptrec_embedding_bwd_fused_adagrad": (c_int, _BWD_ARGS),
    "ptrec_embedding_bwd_fused_rowwise_adagrad": (c_int, _BWD_ARGS),
    "ptrec_embedding_bwd_fused_lazy_adam": (c_int, _BWD_ARGS),
    "ptrec_embedding_bwd_segment_sum": (c_int, [c_int32, c_int32, _FD, _FD, c_int32, c_int64, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                                c_void_p, c_void_p, c_void_p]),
    "ptrec_fm_head_supported": (c_int, [c_int32, c_int32, c_int32]),
    "ptrec_fm_head_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                  c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                  c_void_p]),
    "ptrec_fm_head_bwd_workspace_bytes": (c_size_t, [c_int32]),
    "ptrec_fm_head_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_fm_head_fwd_h2": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                     c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                     c_void_p, c_void_p]),
    "ptrec_rowdot_bwd_h2": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_bce_logits_workspace_bytes": (c_size_t, []),
    "ptrec_bce_logits_mean": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_din_attn_pool_fwd_ids": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                            c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                            c_int32] + [c_void_p] * 6 + [c_void_p, c_void_p, c_void_p]),
    "ptrec_din_attn_pool_bwd_ids": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                            c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32]
                                    + [c_void_p] * 6 + [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                                        c_size_t, c_void_p]),
    "ptrec_set_reduce_stream": (None, [c_void_p]),
    "ptrec_dcn_head_fwd": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "ptrec_dcn_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_set_dcn_2sm": (None, [c_int32]),
    "ptrec_dcn_2sm_enabled": (c_int32, []),
    "ptrec_rowdot_supported": (c_int, [c_int32]),
    "ptrec_rowdot_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "ptrec_rowdot_bwd_workspace_bytes": (c_size_t, [c_int32]),
    "ptrec_rowdot_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "ptrec_dense_optim_chunk": (c_int32, []),
    "ptrec_dense_optim_step": (c_int, [c_void_p, c_void_p, c_int32, c_int32, POINTER(OptimArgs), c_void_p]),
    "ptrec_peer_sync_max_ranks": (c_int32, []),
    "ptrec_peer_sync_slots": (c_int32, []),
    "ptrec_peer_barrier": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "ptrec_dense_pack": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_dense_optim_step_reduce": (c_int, [c_void_p, c_void_p, c_int32, c_int32, POINTER(OptimArgs), c_void_p,
                                              c_int32, c_float, c_void_p]),
    "ptrec_tc_set_2sm": (None, [c_int32]),
    "ptrec_tc_set_bk": (None, [c_int32]),
    "ptrec_tc_2sm_enabled": (c_int32, []),
    "ptrec_tc_split3_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ptrec_tc_split3": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_tc_gemm_split3_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "ptrec_tc_gemm_split3_default_splits": (c_int32, [c_int64, c_int64, c_int64]),
    "ptrec_tc_gemm_split3": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                     c_int32, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_size_t,
                                     c_void_p]),
    "ptrec_tc_gemm_split3_tn": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                        c_int64, c_int32, c_void_p, c_size_t, c_void_p]),
    "ptrec_tc_set_bn": (None, [c_int32]),
    "ptrec_tc_get_bn": (c_int32, []),
    "ptrec_tc_split2h_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ptrec_tc_split2h": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                 c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_tc_gemm_split2h": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                      c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_int32, c_void_p,
                                      c_size_t, c_void_p]),
    "ptrec_tc_gemm_split2h_tn": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                         c_int64, c_void_p, c_int64, c_int32, c_void_p, c_size_t, c_void_p]),
    "ptrec_tc_scale_roll": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "ptrec_tc_split2h_prescaled": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                           c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_tc_gemm_fused_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ptrec_tc_gemm_split2h_fused": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                            c_int64, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                            c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                            c_void_p]),
    "ptrec_a2a_pack_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "ptrec_a2a_pack_by_owner": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "ptrec_a2a_scatter_rows": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p,
                                       c_int64, c_void_p]),
    "ptrec_a2a_pack_by_owner_peer": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ptrec_a2a_scatter_rows_peer": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p,
                                            c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "ptrec_a2a_pack_by_owner_push": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_size_t, c_void_p]),
    "ptrec_a2a_scatter_rows_peer_ordered": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                                    c_float, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "ptrec_gather_push": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                  c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_a2a_scatter_rows_peer_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64,
                                                  c_int32, c_float, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                                  c_void_p]),
    "ptrec_din_attn_pool_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32,
                                        c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "ptrec_din_attn_pool_grad_floats": (c_int32, [c_int32, c_int32, c_int32]),
    "ptrec_din_attn_pool_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "ptrec_set_din_tc": (None, [c_int32]),
    "ptrec_din_tc_enabled": (c_int32, []),
    "ptrec_din_attn_pool_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32,
                                        c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "ptrec_dcn_cross_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p,
                                    c_void_p, c_void_p]),
    "ptrec_dcn_cross_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p,
                                      c_void_p, c_void_p]),
    "ptrec_dcn_cross_wgrad_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "ptrec_dcn_cross_wgrad": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "ptrec_dcn_prep_weight": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ptrec_dcn_pack_input": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_dcn_unpack": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_dcn_bwd_init": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "ptrec_dcn_bwd_layer_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "ptrec_dcn_bwd_layer": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "ptrec_dcn_bwd_final": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_fm2_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ptrec_fm2_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                              c_void_p, c_int64, c_void_p]),
}

_lib = None


class PtrecError(RuntimeError):
    """A libptrec_b200 call returned a negative PTREC_E* code."""


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the sm_100a extension first "
            "(`make -C pytorchrec_b200/csrc` or `__graft_entry__.build()`). "
            "pytorchrec_b200 has no CPU or library fallback for its hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    ver = lib.ptrec_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libptrec_b200.so ABI {ver} != binding ABI {ABI_VERSION}: rebuild")
    if os.environ.get("PTREC_ONE_SWEEP"):     # K2a: one-sweep radix sort on / off (A/B measurements)
        lib.ptrec_set_one_sweep_sort(int(os.environ["PTREC_ONE_SWEEP"]))
    if os.environ.get("PTREC_DIN_TC"):        # K4 forward: tensor-core build on / off
        lib.ptrec_set_din_tc(int(os.environ["PTREC_DIN_TC"]))
    if os.environ.get("PTREC_TC_BN"):  # K6 fp16 x 2 pair-tile width, 128 or 256 (A/B measurements)
        lib.ptrec_tc_set_bn(int(os.environ["PTREC_TC_BN"]))
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ptrec_last_error().decode("utf-8", "replace")
        raise PtrecError(f"{what or 'libptrec_b200'} failed ({rc}): {msg}")
