"""ctypes wrapper over oracle/libg1s_oracle.so (the CPU restatement of the `diff` path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.

The class mirrors the three reference call sites on av1_grain::DiffGenerator
(/root/reference/src/main.rs:420-427 `new`, :442 `diff_frame`, :524 `finish`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from grav1synth_b200.abi import CFrame, CSegment, GrainTableSegment, frame_from_planes, segments_to_c

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libg1s_oracle.so")

GRAM_REF_ORDER = 0
GRAM_EXACT_INT = 1
EXP_LIBM = 0
EXP_FIXED = 1

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc -O3 -ffp-contract=off)."""
    src = os.path.join(_HERE, "g1s_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.g1s_oracle_new.restype = C.c_void_p
        L.g1s_oracle_new.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
        L.g1s_oracle_free.argtypes = [C.c_void_p]
        L.g1s_oracle_diff_frame.restype = C.c_int
        L.g1s_oracle_diff_frame.argtypes = [C.c_void_p, C.POINTER(CFrame), C.c_int, C.c_int, C.POINTER(CFrame),
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.g1s_oracle_finish.restype = C.c_int
        L.g1s_oracle_finish.argtypes = [C.c_void_p, C.POINTER(CSegment), C.c_size_t, C.POINTER(C.c_size_t)]
        L.g1s_oracle_last_error.restype = C.c_char_p
        L.g1s_oracle_last_error.argtypes = [C.c_void_p]
        L.g1s_oracle_last_status.argtypes = [C.c_void_p]
        L.g1s_oracle_last_num_flat.argtypes = [C.c_void_p]
        L.g1s_oracle_last_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.g1s_oracle_last_gram.restype = C.c_int64
        L.g1s_oracle_last_gram.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.g1s_oracle_get_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.g1s_oracle_exp_fixed.restype = C.c_double
        L.g1s_oracle_last_eqns.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.g1s_oracle_exp_fixed.argtypes = [C.c_double]
        L.g1s_oracle_write_table.argtypes = [C.POINTER(CSegment), C.c_size_t, C.c_char_p]
        _lib = L
    return _lib


class OracleDiffGenerator:
    """CPU oracle with the reference's DiffGenerator surface."""

    def __init__(self, fps_num: int, fps_den: int, source_bit_depth: int, denoised_bit_depth: int,
                 gram_mode: int = GRAM_EXACT_INT, exp_mode: int = EXP_FIXED, ss_x: int = 1, ss_y: int = 1):
        self._L = lib()
        self._h = self._L.g1s_oracle_new(fps_num, fps_den, source_bit_depth, denoised_bit_depth, gram_mode, exp_mode)
        self.ss_x, self.ss_y = ss_x, ss_y
        self._nb = 0

    def close(self):
        if self._h:
            self._L.g1s_oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def diff_frame(self, source: Sequence[Optional[np.ndarray]], denoised: Sequence[Optional[np.ndarray]]) -> None:
        sf, k1 = frame_from_planes(source)
        df, k2 = frame_from_planes(denoised)
        mono = len(source) < 3 or source[1] is None
        sh, sw = source[0].shape
        dh, dw = denoised[0].shape
        rc = self._L.g1s_oracle_diff_frame(self._h, C.byref(sf), sw, sh, C.byref(df), dw, dh, self.ss_x, self.ss_y,
                                           int(mono))
        self._nb = ((sw + 31) // 32) * ((sh + 31) // 32)
        if rc != 0:
            raise ValueError(self._L.g1s_oracle_last_error(self._h).decode())

    def finish(self) -> List[GrainTableSegment]:
        n = C.c_size_t(0)
        cap = 64
        while True:
            arr = (CSegment * cap)()
            rc = self._L.g1s_oracle_finish(self._h, arr, cap, C.byref(n))
            if rc == 0:
                break
            cap = n.value + 1
        return [GrainTableSegment.from_c(arr[i]) for i in range(n.value)]

    # ---- introspection (tests) ----
    @property
    def last_status(self) -> int:
        return self._L.g1s_oracle_last_status(self._h)

    @property
    def last_num_flat(self) -> int:
        return self._L.g1s_oracle_last_num_flat(self._h)

    def last_flat(self):
        flat = np.zeros(self._nb, np.uint8)
        scores = np.zeros(self._nb, np.float32)
        feat = np.zeros((self._nb, 5), np.float64)
        self._L.g1s_oracle_last_flat(self._h, flat.ctypes.data, scores.ctypes.data, feat.ctypes.data)
        return flat, scores, feat

    def last_gram(self, c: int):
        G = np.zeros((26, 26), np.int64)
        nobs = self._L.g1s_oracle_last_gram(self._h, c, G.ctypes.data)
        return G, int(nobs)

    def last_eqns(self, c: int):
        """(A [n][n], b [n]) of the latest frame's AR normal equations, in this handle's accumulation mode."""
        n = 24 if c == 0 else 25
        A = np.zeros((n, n))
        b = np.zeros(n)
        self._L.g1s_oracle_last_eqns(self._h, c, A.ctypes.data, b.ctypes.data)
        return A, b

    def state(self, which: int, c: int):
        n = 24 if c == 0 else 25
        x = np.zeros(n)
        gain = C.c_double(0)
        sx = np.zeros(20)
        nobs = C.c_int64(0)
        self._L.g1s_oracle_get_state(self._h, which, c, x.ctypes.data, C.byref(gain), sx.ctypes.data, C.byref(nobs))
        return x, gain.value, sx, nobs.value


def write_grain_table(segs: Sequence[GrainTableSegment], path: str) -> None:
    arr = segments_to_c(segs)
    rc = lib().g1s_oracle_write_table(arr, len(segs), path.encode())
    if rc != 0:
        raise OSError(f"oracle writer failed: {rc}")


def exp_fixed(x: float) -> float:
    return lib().g1s_oracle_exp_fixed(x)
