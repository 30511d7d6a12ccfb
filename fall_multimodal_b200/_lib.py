"""ctypes loader for libfmm_b200.so (the C-ABI CUDA library declared in include/fmm_b200.h).

There is deliberately NO fallback: if the shared object is missing or the device is not an
sm_100 part, every op raises.  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libfmm_b200.so"
_lib = None

DT_BF16 = 0
DT_F32 = 1

c_void_p = C.c_void_p
c_int = C.c_int
c_ll = C.c_longlong
c_float = C.c_float


def dt_of(t: torch.dtype) -> int:
    if t == torch.bfloat16:
        return DT_BF16
    if t == torch.float32:
        return DT_F32
    raise TypeError(f"libfmm_b200 activations must be bf16 or fp32, got {t}")


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python -m fall_multimodal_b200.build` "
            "(there is no CPU/PyTorch fallback for the CUDA path)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.fmm_last_error.restype = C.c_char_p
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


# number of kernels each C entry point launches (default 1); used for the launch counter
_KERNELS_PER_CALL = {"se_fwd": 3, "se_bwd": 5, "conv1d_k5_bwd": 2, "head_ce_bwd": 2}
launch_count = 0


def check(status: int, what: str) -> None:
    global launch_count
    launch_count += _KERNELS_PER_CALL.get(what, 1)
    if status != 0:
        msg = load().fmm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libfmm_b200 {what} failed (status {status}): {msg}")


def ptr(t) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_device(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("libfmm_b200 ops need CUDA tensors (no CPU fallback)")


_P = c_void_p
_SIGNATURES = {
    "fmm_version": [],
    "fmm_debug_mma_probe": [c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "fmm_debug_wait_profile": [c_int, _P],
    "fmm_device_supported": [],
    "fmm_set_sm_limit": [c_int],
    "fmm_tapconv_bn": [c_int, c_int],
    "fmm_tapconv_packed_bytes": [c_int, c_int, c_int, c_int],
    "fmm_tapconv_pack": [_P, _P, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_int,
                         C.POINTER(c_int), c_int, _P],
    "fmm_tapconv": [_P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                    c_int, c_int, c_int, c_int, C.POINTER(c_int), c_int, _P, _P],
    "fmm_gcn_packed_bytes": [c_int, c_int, c_int],
    "fmm_gcn_pack": [_P, _P, c_int, c_int, c_int, _P],
    "fmm_gcn_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(c_int), _P, _P, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "fmm_gcn_wgrad": [_P, _P, _P, _P, _P, _P, C.POINTER(c_int), c_ll, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "fmm_gcn_packed_bwd_bytes": [c_int, c_int, c_int],
    "fmm_gcn_pack_bwd": [_P, _P, c_int, c_int, c_int, _P],
    "fmm_gcn_bwd": [_P] * 11 + [c_int, c_int, c_ll, c_int, c_int, c_int, c_int, _P, _P],
    "fmm_databn_stats": [_P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "fmm_databn_apply": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_databn_bwd": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_head_ce_fwd": [_P, _P],
    "fmm_head_ce_bwd": [_P, _P],
    "fmm_opt_chunk": [],
    "fmm_rmsprop_step": [_P, _P, _P, c_int, _P, c_float, c_float, c_float, _P, c_float, _P, _P],
    "fmm_grad_norm_sq": [_P, _P, _P, c_int, _P, _P],
    "fmm_gcn_prep_fwd": [_P] * 9 + [c_int, c_int, c_int, c_int, _P],
    "fmm_gcn_prep_bwd": [_P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "fmm_agg_fwd": [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_agg_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_agg_dcoef": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_colstats": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_affine_relu": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_block_out": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_blockout_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_bn2_bwd_apply": [_P] * 15 + [c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_bn1_bwd_reduce": [_P] * 6 + [c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_bn1_bwd_apply": [_P] * 9 + [c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_bn_finalize": [_P, _P, c_int, C.c_double, _P, _P, _P, _P, c_float, c_float, c_int, _P, _P, _P, _P, c_int, _P],
    "fmm_se_fwd": [_P, _P, _P, c_float, _P, _P, _P, _P, _P, _P, c_float, c_float, c_int] + [_P] * 11 +
                  [c_int, c_int, c_int, _P],
    "fmm_se_bwd": [_P] * 13 + [c_int] + [_P] * 11 + [c_int, c_int, c_int, _P],
    "fmm_se_bwd_params": [_P] * 8 + [c_int, c_int, c_int, _P],
    "fmm_bn2_bwd_coef": [_P] * 12 + [c_float, C.c_double, c_int] + [_P] * 10 + [c_int, c_int, _P],
    "fmm_bn1_bwd_coef": [_P, _P, c_int, _P, _P, _P, C.c_double, c_int] + [_P] * 5 + [c_int, _P],
    "fmm_lstm_fwd": [_P] * 8 + [c_int] * 5 + [_P],
    "fmm_lstm_infer": [_P] * 6 + [c_int] * 6 + [_P],
    "fmm_lstm_bwd": [_P] * 11 + [c_int] * 5 + [_P],
    "fmm_conv1d_k5_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "fmm_bn_relu_pool2_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, _P],
    "fmm_pool2_bwd": [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "fmm_conv1d_k5_bwd": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "fmm_wgrad": [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                  C.POINTER(c_int), c_int, c_ll, c_ll, c_ll, c_ll, c_int, _P, _P],
    "fmm_dwconv_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_dwconv_bwd_data": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_dwconv_bwd_weight": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_affine_act": [_P, _P, _P, _P, _P, c_ll, c_int, c_int, c_int, _P],
    "fmm_bn_act_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, c_ll, c_int, c_int, c_int, _P],
    "fmm_bn_act_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, C.c_double, c_int, _P, _P, c_ll, c_int, c_int, c_int, _P],
    "fmm_prep_frames": [_P, _P, _P, _P, _P, c_int, c_int, c_int, C.c_uint, c_int, c_int, _P],
    "fmm_prep_windows": [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_bgemm": [_P, _P],
    "fmm_tg_cell_fwd": [_P, c_int, _P],
    "fmm_tg_cell_bwd": [_P, c_int, _P],
    "fmm_gruscan_geometry": [c_int, _P, _P],
    "fmm_gruscan_max_clusters": [c_int],
    "fmm_gruscan": [_P, c_int, _P],
    "fmm_gruscan_export_xc": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_gruscan_export_dg": [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "fmm_gruscan_mix_dx": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_gruscan_export_fs": [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    "fmm_tattn": [_P, c_int, _P],
    "fmm_pn_dgrad": [_P, _P, _P, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P],
    "fmm_pn_ds_parts": [],
    "fmm_pn_ds": [_P, _P, _P, c_ll, c_int, c_int, _P],
    "fmm_pn_wgrad_chunks": [c_int, c_ll, c_int],
    "fmm_pn_wgrad": [_P, _P, _P, c_int, c_ll, c_int, c_int, c_int, _P],
    "fmm_tg_softmax_fwd": [_P, c_ll, c_int, c_int, c_int, _P],
    "fmm_tg_softmax_bwd": [_P, _P, c_ll, c_int, c_int, c_int, _P],
    "fmm_tg_ln_fwd": [_P, _P, _P, _P, _P, _P, _P, c_ll, c_int, c_float, c_int, _P],
    "fmm_tg_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_ll, c_int, c_int, _P],
    "fmm_tg_add_pe": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "fmm_tg_relu_mask": [_P, _P, c_ll, c_int, _P],
    "fmm_tg_transpose": [_P, _P, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_int, c_int, c_int, _P],
}
_RESTYPES = {"fmm_tapconv_packed_bytes": c_ll, "fmm_gcn_packed_bytes": c_ll, "fmm_gcn_packed_bwd_bytes": c_ll}


class BgemmDesc(C.Structure):
    """Mirror of ``fmm_bgemm_desc`` (include/fmm_b200.h)."""
    _fields_ = ([(n, c_void_p) for n in ("A", "B", "C", "bias_m", "bias_n")] +
                [(n, c_ll) for n in ("a_g1", "a_g2", "a_m", "a_k1", "a_k2", "a_k3", "b_g1", "b_g2", "b_n", "b_k1", "b_k2",
                                     "b_k3", "c_g1", "c_g2", "c_m", "c_n")] +
                [(n, c_int) for n in ("G1", "G2", "M", "N", "K1", "K2", "K3")] +
                [("alpha", c_float), ("beta", c_int), ("act", c_int), ("splitk", c_int), ("dtype", c_int), ("c_dtype", c_int)])


def _struct(name, spec):
    fields = []
    for kind, names in spec:
        fields += [(n, kind) for n in names.split()]
    return type(name, (C.Structure,), {"_fields_": fields})


# mirrors of fmm_cell_fwd_args / fmm_cell_bwd_args (include/fmm_b200.h, csrc/gru_cell.cu)
CellFwdArgs = _struct("CellFwdArgs", [
    (c_void_p, "x"), (c_ll, "xb xv"), (c_void_p, "hprev"), (c_ll, "hb hv"), (c_void_p, "S pre lin zr lg hc lu hout"),
    (c_ll, "ob ov"), (c_void_p, "xc0 xc1"), (c_int, "mode B V Din H Cp")])
CellBwdArgs = _struct("CellBwdArgs", [
    (c_void_p, "S carry dz dxc0 dxc1 dx"), (c_ll, "dxb dxv"), (c_void_p, "hprev"), (c_ll, "hb hv"),
    (c_void_p, "zr lg dpre_g dlin_g dH"), (c_ll, "db dv"), (c_void_p, "z1 hprev1"), (c_ll, "hb1 hv1"),
    (c_void_p, "hc1 lu1 dpre_u dlin_u"), (c_int, "mode dx_accum do_bwd1 B V Din H Cp")])


# mirror of fmm_gruscan_args (include/fmm_b200.h, csrc/gruscan.cu)
GruScanArgs = _struct("GruScanArgs", [
    (c_void_p, "xb px xcg xcu fs hout W Lw cs S bg bl dhout"), (c_ll, "dh_b dh_t dh_v"), (c_void_p, "dxu dxgz dxgr WT LT err prof"),
    (c_int, "B T V KS xb_slices xb_slot0 NC tsplit")])


# mirror of fmm_tattn_args (include/fmm_b200.h, csrc/tattn.cu)
TAttnArgs = _struct("TAttnArgs", [(c_void_p, "q k v out lse dout dq dk dv"), (c_int, "B V T Tp F"), (c_float, "scale"), (c_int, "v_btvc")])


# mirror of fmm_head_args (include/fmm_b200.h, csrc/head.cu)
HeadArgs = _struct("HeadArgs", [
    (c_void_p * 4, "feat dfeat"), (c_int * 4, "width"), (c_int, "nseg"), (c_void_p, "W bias target out prob prob2 loss gloss dz dW dbias"),
    (c_int, "N C F pre_softmax"), (c_float, "smoothing")])


def int_array(vals):
    arr = (c_int * len(vals))(*vals)
    return arr
