"""Python faces of the C-ABI kernels (raw pointers in, status out; see include/fmm_b200.h).

Activations are channels-last ``(N, T, V, C)`` contiguous CUDA tensors, bf16 or fp32.
Every function launches on ``torch.cuda.current_stream()`` and raises on a non-zero status.
"""
from __future__ import annotations

import torch

from . import _lib as L


_err_words: dict[int, torch.Tensor] = {}


def err_word(device: torch.device) -> torch.Tensor:
    """Per-device watchdog word the GEMM kernels write to before trapping on a stuck barrier."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _err_words:
        _err_words[idx] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_words[idx]


class PackedWeight:
    """Weights of one tap-conv, converted to the tensor-core shared-memory images."""

    __slots__ = ("buf", "cout", "cin", "ntaps", "dtype")

    def __init__(self, buf, cout, cin, ntaps, dtype):
        self.buf, self.cout, self.cin, self.ntaps, self.dtype = buf, cout, cin, ntaps, dtype


def tapconv_pack(w: torch.Tensor, cout: int, cin: int, n2: int, k2: int, sn1: int, sn2: int,
                 sk1: int, sk2: int, sm: int, tapmap, act_dtype: torch.dtype,
                 out: torch.Tensor | None = None) -> PackedWeight:
    """Pack fp32 weights ``w`` (any 2-level strided view, see csrc/tapconv.cu) for `tapconv`."""
    L.require_device(w)
    assert w.dtype == torch.float32
    lib = L.load()
    dt = L.dt_of(act_dtype)
    ntaps = len(tapmap)
    nbytes = lib.fmm_tapconv_packed_bytes(cin, cout, ntaps, dt)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    assert out.numel() >= nbytes
    st = lib.fmm_tapconv_pack(L.ptr(w), L.ptr(out), cout, cin, n2, k2, sn1, sn2, sk1, sk2, sm, ntaps,
                              L.int_array(list(tapmap)), dt, L.stream())
    L.check(st, "tapconv_pack")
    return PackedWeight(out, cout, cin, ntaps, act_dtype)


def tapconv(x: torch.Tensor, pw: PackedWeight, out: torch.Tensor, *, shifts, tj: int,
            istride: int = 1, ostride: int = 1, ooff: int = 0, in_scale=None, in_shift=None,
            in_relu: bool = False, bias=None) -> torch.Tensor:
    """out[n, j*ostride+ooff, v, :] = bias + sum_m f(x[n, j*istride+shifts[m], v, :]) @ W[m]^T."""
    L.require_device(x)
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype == pw.dtype
    N, Tin, V, Cin = x.shape
    No, Tout, Vo, Cout = out.shape
    assert (No, Vo) == (N, V) and Cin == pw.cin and Cout == pw.cout and len(shifts) == pw.ntaps
    lib = L.load()
    st = lib.fmm_tapconv(L.ptr(x), L.ptr(out), L.ptr(pw.buf), L.ptr(in_scale), L.ptr(in_shift),
                         int(in_relu), L.ptr(bias), N, V, Tin, Tout, Cin, Cout, tj, istride, ostride,
                         ooff, len(shifts), L.int_array(list(shifts)), L.dt_of(x.dtype),
                         L.ptr(err_word(x.device)), L.stream())
    L.check(st, "tapconv")
    return out


def wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, *, shifts, istride: int = 1,
          in_scale=None, in_shift=None, in_relu: bool = False, c2: int | None = None,
          s_m: int, s_c1: int = 0, s_c2: int, s_co: int) -> torch.Tensor:
    """dw[m, ci, co] += sum_{n,v,j} f(x[n, j*istride+shifts[m], v, ci]) * dy[n, j, v, co].

    ``dw`` is an fp32 buffer addressed as ``m*s_m + (ci//c2)*s_c1 + (ci%c2)*s_c2 + co*s_co`` and is
    accumulated into with atomics: zero it first.
    """
    L.require_device(x)
    assert x.is_contiguous() and dy.is_contiguous() and x.dtype == dy.dtype and dw.dtype == torch.float32
    N, Tin, V, Cin = x.shape
    Ny, Tj, Vy, Cout = dy.shape
    assert (Ny, Vy) == (N, V)
    lib = L.load()
    st = lib.fmm_wgrad(L.ptr(x), L.ptr(dy), L.ptr(dw), L.ptr(in_scale), L.ptr(in_shift), int(in_relu),
                       N, V, Tin, Tj, Cin, Cout, istride, len(shifts), L.int_array(list(shifts)),
                       c2 if c2 is not None else Cin, s_m, s_c1, s_c2, s_co, L.dt_of(x.dtype),
                       L.ptr(err_word(x.device)), L.stream())
    L.check(st, "wgrad")
    return dw
