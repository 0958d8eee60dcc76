"""ctypes binding of libdrb200.so (include/drb200.h).  There is NO fallback: if the library is missing or a call
fails, a Python exception is raised (ValueError for DRB_ERR_INVALID, RuntimeError otherwise)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRB200_LIB", os.path.join(HERE, "libdrb200.so"))   # override: tuning builds only

DRB_OK, DRB_ERR_INVALID, DRB_ERR_CUDA, DRB_ERR_UNSUPPORTED = 0, -1, -2, -3
EPI_STORE, EPI_GELU, EPI_GATED_RESIDUAL = 0, 1, 2
TMODE_CAUSAL, TMODE_DOWN2, TMODE_UP2 = 0, 1, 2
RES_NONE, RES_SAME, RES_FRAME_UP2, RES_POOL_HW, RES_POOL_T, RES_NEAREST_UP_HW = 0, 1, 2, 3, 4, 5


class Conv3dArgs(Structure):
    """drb_conv3d_args of include/drb200.h"""
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("out", c_void_p), ("resid", c_void_p),
                ("stats", c_void_p)] + [(n, c_int) for n in (
                    "T_in", "H_in", "W_in", "Cin", "T_out", "H_out", "W_out", "Cout", "kt", "kh", "kw", "pad_h", "pad_w",
                    "stride_hw", "tmode", "out_scale", "out_off_h", "out_off_w", "resid_mode", "resid_H", "resid_W")]


class CpSync(Structure):
    """drb_cp_sync of include/drb200.h"""
    _fields_ = [("flag_ptrs", POINTER(c_void_p)), ("counter", c_void_p), ("world", c_int), ("rank", c_int), ("signal_slot", c_int),
                ("wait_slot", c_int), ("signal_epoch", c_uint32), ("wait_epoch", c_uint32), ("timeout_ms", c_uint32)]


# name -> argtypes, in the order of include/drb200.h (tests check that every one is exported)
PROTOTYPES = {
    "drb_gemm_bf16": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p,
                      c_int64, c_void_p, c_int, c_void_p],
    "drb_gemm_bf16_sync": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p,
                           c_int64, c_void_p, c_int, POINTER(CpSync), c_void_p],
    "drb_attention_bf16": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p],
    "drb_adaln_modulate": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "drb_qk_norm_rope": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "drb_gemv_bf16": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "drb_gemv_bf16_batched": [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int,
                              c_int, c_int, c_int, c_void_p],
    "drb_sigma_embedding": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "drb_scale_patchify": [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p],
    "drb_patchify_condition": [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "drb_unpatchify_euler": [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_void_p],
    "drb_edm_scale_input": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "drb_edm_euler_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
    "drb_postprocess_u8": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "drb_conv3d_cl": [POINTER(Conv3dArgs), c_void_p],
    "drb_haar_patch": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "drb_haar_unpatch": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "drb_haar_unpatch_u8": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "drb_frame_stats_cl": [c_void_p, c_void_p, c_int, c_int64, c_void_p],
    "drb_groupnorm_apply_cl": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p],
    "drb_softmax_rows": [c_void_p, c_int64, c_int, c_int, c_float, c_void_p],
    "drb_transpose_bf16": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p],
    "drb_spatial_attention_d512": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p],
    "drb_temporal_attention_cl": [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p],
    "drb_planar_to_cl": [c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_void_p],
    "drb_cl_to_planar": [c_void_p, c_void_p, c_int, c_int, c_int64, c_float, c_void_p],
    "drb_latent_normalize": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p],
    "drb_envmap_latlong_to_cubemap": [c_void_p, c_int, c_int, c_float, c_int, c_int, c_void_p, c_int, c_void_p],
    "drb_envmap_project": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "drb_envmap_tonemap": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "drb_attention_bf16_ring": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_void_p],
    "drb_attention_bf16_ring_bounded": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p],
    "drb_gemm_qkv_norm_rope": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, POINTER(c_void_p), c_int, c_int64, c_int, c_void_p],
    "drb_gemm_qkv_norm_rope_batched": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int64, c_int, c_int, c_int, POINTER(CpSync), c_void_p],
    "drb_attention_bf16_bounded": [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p],
    "drb_qk_logit_bound": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "drb_attention_bf16_cp_batched": [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_void_p), c_int, c_int64, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p, POINTER(CpSync), c_void_p],
    "drb_attention_bf16_cp": [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_void_p), c_int, c_int64, c_int, c_int, c_int, c_int,
                              c_int, c_void_p],
    "drb_cp_qk_norm_rope_scatter": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p), c_int,
                                    c_int64, c_int, c_void_p],
    "drb_cp_barrier": [POINTER(c_void_p), c_int, c_int, c_uint32, c_void_p],
    "drb_peer_alloc": [c_int64, POINTER(c_void_p)],
    "drb_peer_free": [c_void_p],
    "drb_peer_export": [c_void_p, c_void_p],
    "drb_peer_import": [c_void_p, POINTER(c_void_p)],
    "drb_peer_close": [c_void_p],
}
CP_MAX_RANKS = 8
CP_FLAG_SLOTS, CP_STATUS_WORD = 4, 63
CP_SLOT_BARRIER, CP_SLOT_QKV, CP_SLOT_ATTN = 0, 1, 2   # flag slots: stand-alone barrier, "q/k/v stored", "attention rows stored"
PEER_HANDLE_BYTES = 64


def ptr_array(ptrs):
    """ctypes void*[n] from a list of integer device pointers"""
    return (c_void_p * len(ptrs))(*ptrs)

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libdrb200.so (built by build.py) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The B200 path has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.drb_last_error.restype = c_char_p
    lib.drb_last_error.argtypes = []
    lib.drb_version.restype = c_int
    lib.drb_device_supported.restype = c_int
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == DRB_OK:
        return
    msg = load().drb_last_error().decode(errors="replace")
    if rc == DRB_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed ({rc}): {msg}")


LAUNCHES = 0   # C-ABI kernel entry points called so far (bench.py reports the count inside its timed region)


def call(name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += 1
    check(getattr(load(), name)(*args), name)
