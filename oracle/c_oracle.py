"""ctypes front end of the plain-C oracle (oracle/rqae_oracle.c).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import build as _build

_lib = None

# The fp32 summation order of the sm_100a forward kernel (rqae_b200/csrc/rq_forward.cuh): 256 interleaved
# per-thread partial sums, warp tree with lane strides 1,2,16,8,4, the 8 warp sums combined as a pairwise tree,
# bias folded into the out-projection fma chain, reconstruction emitted as x - r_final.  With these switches the C oracle is
# bit-identical to the kernel (codes AND reconstruction), which is what the GPU tests assert.
KERNEL_ORDER = dict(order_nt=256, tree=1, gtree=1, fold_bias=True, recon="x_minus_r")
# The opt-in D-split cluster variant of the kernel at Gemma-2-9B width (rqae_forward_variant(1), 2304 < D <= 3584): two
# CTAs own 7 x 256 elements of the d axis each, every slice is summed in the order above, the two slice sums are added,
# then the bias.
KERNEL_ORDER_9B = dict(KERNEL_ORDER, seg_blocks=7)


def lib():
    global _lib
    if _lib is None:
        so = _build.SO
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(_build.SRC):
            so = _build.build()
        _lib = ctypes.CDLL(so)
        _lib.rqo_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class CWeights:
    """Stacked fp32 weights in reference layouts: w_in (nq,cd,D), b_in (nq,cd),
    w_out (nq,D,cd), b_out (nq,D), codebook (nq|1,K,cd)."""

    def __init__(self, w_in, b_in, w_out, b_out, codebook, shared_codebook: bool):
        self.w_in, self.b_in, self.w_out, self.b_out = _f32(w_in), _f32(b_in), _f32(w_out), _f32(b_out)
        cb = _f32(codebook)
        if cb.ndim == 2:
            cb = cb[None]
        self.shared = bool(shared_codebook)
        self.codebook = np.ascontiguousarray(cb[:1]) if self.shared else cb
        self.nq, self.cd, self.D = self.w_in.shape
        self.K = self.codebook.shape[1]

    @classmethod
    def from_stacked(cls, w) -> "CWeights":
        """From oracle.rqae_oracle.StackedWeights."""
        shared = w.quantization_method in ("fsq", "round_fsq")
        return cls(w.w_in.numpy(), w.b_in.numpy(), w.w_out.numpy(), w.b_out.numpy(), w.codebook.numpy(), shared)


def num_threads() -> int:
    return lib().rqo_num_threads()


def forward_f32(w: CWeights, x, max_layers: Optional[int] = None, order_nt: int = 0, fold_bias: bool = False,
                recon: str = "accumulate", teacher=None, want_q: bool = True, tree: int = 0, gtree: int = 0,
                seg_blocks: int = 0):
    x = _f32(x)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, w.D)
    n = x2.shape[0]
    nq_run = w.nq if max_layers is None else max(0, min(int(max_layers), w.nq))
    codes = np.empty((n, nq_run), np.int32)
    q = np.empty((n, w.D), np.float32) if want_q else None
    t = None if teacher is None else np.ascontiguousarray(np.asarray(teacher, np.int32).reshape(n, nq_run))
    rc = lib().rqo_forward_f32(_p(w.w_in), _p(w.b_in), _p(w.w_out), _p(w.b_out), _p(w.codebook),
                               ctypes.c_int(int(w.shared)), ctypes.c_int(nq_run), ctypes.c_int(w.D),
                               ctypes.c_int(w.cd), ctypes.c_int(w.K), _p(x2), ctypes.c_long(n),
                               ctypes.c_int(order_nt), ctypes.c_int(int(tree)), ctypes.c_int(int(gtree)),
                               ctypes.c_int(int(seg_blocks)), ctypes.c_int(int(fold_bias)),
                               ctypes.c_int({"accumulate": 0, "x_minus_r": 1}[recon]), _p(t), _p(codes), _p(q))
    if rc:
        raise RuntimeError(f"rqo_forward_f32 failed: {rc}")
    return (None if q is None else q.reshape(*lead, w.D)), codes.reshape(*lead, nq_run)


def forward_f64(w: CWeights, x, max_layers: Optional[int] = None, teacher=None, want_q: bool = False):
    x = _f32(x)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, w.D)
    n = x2.shape[0]
    nq_run = w.nq if max_layers is None else max(0, min(int(max_layers), w.nq))
    codes = np.empty((n, nq_run), np.int32)
    margins = np.empty((n, nq_run), np.float32)
    q = np.empty((n, w.D), np.float64) if want_q else None
    t = None if teacher is None else np.ascontiguousarray(np.asarray(teacher, np.int32).reshape(n, nq_run))
    rc = lib().rqo_forward_f64(_p(w.w_in), _p(w.b_in), _p(w.w_out), _p(w.b_out), _p(w.codebook),
                               ctypes.c_int(int(w.shared)), ctypes.c_int(nq_run), ctypes.c_int(w.D),
                               ctypes.c_int(w.cd), ctypes.c_int(w.K), _p(x2), ctypes.c_long(n),
                               _p(t), _p(codes), _p(q), _p(margins))
    if rc:
        raise RuntimeError(f"rqo_forward_f64 failed: {rc}")
    return (None if q is None else q.reshape(*lead, w.D)), codes.reshape(*lead, nq_run), margins.reshape(*lead, nq_run)


def decode_f32(w: CWeights, codes=None, cv=None, layers: Optional[Sequence[int]] = None):
    if codes is not None:
        codes = np.ascontiguousarray(np.asarray(codes, np.int32))
        lead, nq = codes.shape[:-1], codes.shape[-1]
        n = int(np.prod(lead)) if lead else 1
    else:
        cv = _f32(cv)
        lead, nq = cv.shape[:-2], cv.shape[-2]
        n = int(np.prod(lead)) if lead else 1
    mask = None
    if layers is not None:
        mask = np.zeros(nq, np.uint8)
        for l in layers:
            if 0 <= l < nq:
                mask[l] = 1
    q = np.empty((n, w.D), np.float32)
    cb0 = np.ascontiguousarray(w.codebook[0])
    rc = lib().rqo_decode_f32(_p(w.w_out), _p(w.b_out), _p(cb0), ctypes.c_int(nq), ctypes.c_int(w.D),
                              ctypes.c_int(w.cd), ctypes.c_int(w.K), _p(codes), _p(cv), _p(mask),
                              ctypes.c_long(n), _p(q))
    if rc:
        raise RuntimeError(f"rqo_decode_f32 failed: {rc}")
    return q.reshape(*lead, w.D)
