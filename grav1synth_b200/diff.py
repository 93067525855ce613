"""Host-side mirror of the reference's operator interface for the `diff` path.

`DiffGenerator` has the same three-call surface as `av1_grain::DiffGenerator`, which is all
grav1synth uses (/root/reference/src/main.rs:420-427 `new`, :442/462/482/502 `diff_frame`,
:524 `finish`), and `write_grain_table` is the `filmgrn1` writer of src/main.rs:525-530 +
631-696.  Everything goes through the C ABI of include/g1s.h (libg1s.so); there is no
Python or CPU implementation behind it — if the CUDA library or a GPU is missing the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import abi
from .abi import CDiffConfig, CFrame, CSegment, GrainTableSegment, frame_from_planes, segments_to_c

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libg1s.so")
_lib: Optional[C.CDLL] = None

RECORD_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t)

EXPORTS = [
    "g1s_abi_version", "g1s_diff_create", "g1s_diff_push_frame", "g1s_diff_push_frame_device", "g1s_diff_flush",
    "g1s_diff_finish", "g1s_diff_destroy", "g1s_diff_last_error", "g1s_diff_frames_pushed", "g1s_diff_batch_frames", "g1s_diff_model_on_device",
    "g1s_diff_frame_device",
    "g1s_diff_get_counters", "g1s_diff_mark", "g1s_diff_marks_elapsed_ms", "g1s_record_layout", "g1s_record_gramf_offset", "g1s_diff_record_bytes", "g1s_diff_set_record_tap",
    "g1s_diff_consume_record", "g1s_diff_consume_records", "g1s_digest_bytes", "g1s_diff_set_digest_sink",
    "g1s_diff_digest_count", "g1s_diff_wait_retired", "g1s_diff_consume_digests", "g1s_diff_consume_digests_borrowed", "g1s_diff_digest_from_record", "g1s_write_grain_table", "g1s_format_grain_table",
    "g1s_narrow_row", "g1s_narrow_row_with", "g1s_narrow_isa", "g1s_linsolve_probe", "g1s_diff_set_source_filters", "g1s_resize_table",
]


class G1SError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"g1s error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Loads libg1s.so (built in-tree by grav1synth_b200.build).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m grav1synth_b200.build` "
                "(there is no fallback implementation)")
        L = C.CDLL(LIB_PATH)
        L.g1s_diff_create.argtypes = [C.POINTER(CDiffConfig), C.POINTER(C.c_void_p)]
        L.g1s_diff_push_frame.argtypes = [C.c_void_p, C.POINTER(CFrame), C.POINTER(CFrame)]
        L.g1s_diff_push_frame_device.argtypes = [C.c_void_p, C.POINTER(CFrame), C.POINTER(CFrame)]
        L.g1s_diff_flush.argtypes = [C.c_void_p]
        L.g1s_diff_finish.argtypes = [C.c_void_p, C.POINTER(CSegment), C.c_size_t, C.POINTER(C.c_size_t)]
        L.g1s_diff_destroy.argtypes = [C.c_void_p]
        L.g1s_diff_destroy.restype = None
        L.g1s_diff_last_error.argtypes = [C.c_void_p]
        L.g1s_diff_last_error.restype = C.c_char_p
        L.g1s_diff_frames_pushed.argtypes = [C.c_void_p]
        L.g1s_diff_frames_pushed.restype = C.c_int64
        L.g1s_diff_set_source_filters.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32]
        L.g1s_resize_table.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.g1s_diff_batch_frames.argtypes = [C.c_void_p]
        L.g1s_diff_model_on_device.argtypes = [C.c_void_p]
        L.g1s_diff_frame_device.argtypes = [C.c_void_p, C.c_int64]
        L.g1s_diff_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_size_t]
        L.g1s_diff_mark.argtypes = [C.c_void_p, C.c_int]
        L.g1s_diff_marks_elapsed_ms.argtypes = [C.c_void_p]
        L.g1s_diff_marks_elapsed_ms.restype = C.c_double
        L.g1s_record_layout.argtypes = [C.c_int32, C.POINTER(C.c_size_t)]
        L.g1s_record_layout.restype = C.c_size_t
        L.g1s_record_gramf_offset.argtypes = [C.c_int32]
        L.g1s_record_gramf_offset.restype = C.c_size_t
        L.g1s_diff_record_bytes.argtypes = [C.c_void_p]
        L.g1s_diff_record_bytes.restype = C.c_size_t
        L.g1s_diff_set_record_tap.argtypes = [C.c_void_p, RECORD_FN, C.c_void_p]
        L.g1s_diff_consume_record.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.g1s_diff_consume_records.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        L.g1s_digest_bytes.restype = C.c_size_t
        L.g1s_diff_set_digest_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.g1s_diff_digest_count.argtypes = [C.c_void_p]
        L.g1s_diff_digest_count.restype = C.c_int64
        L.g1s_diff_wait_retired.argtypes = [C.c_void_p, C.c_int64]
        L.g1s_diff_consume_digests.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.g1s_diff_consume_digests_borrowed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.g1s_diff_digest_from_record.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.g1s_write_grain_table.argtypes = [C.POINTER(CSegment), C.c_size_t, C.c_char_p]
        L.g1s_format_grain_table.argtypes = [C.POINTER(CSegment), C.c_size_t, C.c_char_p, C.c_size_t]
        L.g1s_format_grain_table.restype = C.c_int64
        _lib = L
    return _lib


class RecordLayout:
    """Byte offsets of the arrays inside one per-frame record (g1s_record_layout)."""

    NAMES = ("gram", "nobs", "num_flat", "luma_sum", "rsum", "rsq", "score", "flat")

    def __init__(self, num_blocks: int):
        off = (C.c_size_t * 8)()
        self.bytes = int(lib().g1s_record_layout(num_blocks, off))
        self.nb = num_blocks
        self.off = dict(zip(self.NAMES, [int(o) for o in off]))
        self.off["gramf"] = int(lib().g1s_record_gramf_offset(num_blocks))

    def unpack(self, rec: np.ndarray) -> dict:
        nb, o = self.nb, self.off
        b = rec.view(np.uint8)
        return {
            "gram": b[o["gram"]:o["gram"] + 8 * 3 * 351].view(np.int64).reshape(3, 351),
            "nobs": b[o["nobs"]:o["nobs"] + 24].view(np.int64),
            "num_flat": int(b[o["num_flat"]:o["num_flat"] + 8].view(np.int64)[0]),
            "luma_sum": b[o["luma_sum"]:o["luma_sum"] + 4 * nb].view(np.uint32),
            "rsum": b[o["rsum"]:o["rsum"] + 12 * nb].view(np.int32).reshape(3, nb),
            "rsq": b[o["rsq"]:o["rsq"] + 12 * nb].view(np.uint32).reshape(3, nb),
            "score": b[o["score"]:o["score"] + 4 * nb].view(np.float32),
            "flat": b[o["flat"]:o["flat"] + nb],
            # strict mode only (zeros otherwise): reference-order f64 sums of product / 255^2 over the same tap pairs
            "gramf": b[o["gramf"]:o["gramf"] + 8 * 3 * 351].view(np.float64).reshape(3, 351),
        }

    def pack(self, gram, nobs, num_flat, luma_sum, rsum, rsq, score, flat, gramf=None) -> np.ndarray:
        rec = np.zeros(self.bytes, np.uint8)
        u = self.unpack(rec)
        u["gram"][:] = gram
        u["nobs"][:] = nobs
        rec[self.off["num_flat"]:self.off["num_flat"] + 8].view(np.int64)[0] = num_flat
        u["luma_sum"][:] = luma_sum
        u["rsum"][:] = rsum
        u["rsq"][:] = rsq
        u["score"][:] = score
        u["flat"][:] = flat
        if gramf is not None:
            u["gramf"][:] = gramf
        return rec


def gram_pairs_to_matrix(pairs: np.ndarray) -> np.ndarray:
    """351 upper-triangular sums -> symmetric 26x26 int64 matrix."""
    G = np.zeros((26, 26), np.int64)
    iu = np.triu_indices(26)
    G[iu] = pairs
    G[(iu[1], iu[0])] = pairs
    return G


class DiffGenerator:
    """B200 drop-in for av1_grain::DiffGenerator (new / diff_frame / finish).

    `width`/`height`/subsampling are needed up front because device buffers are sized at
    creation; the reference learns them from the first frame.
    """

    def __init__(self, fps_num: int, fps_den: int, source_bit_depth: int, denoised_bit_depth: int, width: int,
                 height: int, ss_x: int = 1, ss_y: int = 1, monochrome: bool = False, device: int = 0,
                 batch_frames: int = 0, mode: int = abi.MODE_FULL, gram_kernel: int = 0, host_threads: int = 0,
                 host_narrow: bool = False, gram_order: int = 0, devices: Optional[Sequence[int]] = None,
                 model_placement: int = 0):
        self._L = lib()
        cfg = CDiffConfig()
        cfg.fps_num, cfg.fps_den = fps_num, fps_den
        cfg.src_bit_depth, cfg.den_bit_depth = source_bit_depth, denoised_bit_depth
        cfg.width, cfg.height, cfg.ss_x, cfg.ss_y = width, height, ss_x, ss_y
        cfg.monochrome, cfg.device, cfg.batch_frames, cfg.mode = int(monochrome), device, batch_frames, mode
        cfg.gram_kernel = gram_kernel
        cfg.host_threads = host_threads
        cfg.host_narrow = int(host_narrow)
        cfg.gram_order = int(gram_order)  # abi.GRAM_EXACT_INT (fast) / abi.GRAM_REF_ORDER (strict: the reference's integers)
        cfg.model_placement = int(model_placement)  # abi.MODEL_AUTO / MODEL_HOST / MODEL_DEVICE
        if devices is not None and len(devices) > 1:
            cfg.n_devices = len(devices)
            for i, dv in enumerate(devices):
                cfg.device_ids[i] = int(dv)
        self.cfg = cfg
        h = C.c_void_p()
        rc = self._L.g1s_diff_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise G1SError(rc, self._L.g1s_diff_last_error(None).decode())
        self._h = h
        self._tap = None
        self.num_blocks = ((width + 31) // 32) * ((height + 31) // 32)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.g1s_diff_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._L.g1s_diff_last_error(self._h).decode()
            if rc == abi.G1S_E_DIMS:
                raise ValueError(msg)
            raise G1SError(rc, msg)

    # -- the reference's surface
    def diff_frame(self, source: Sequence[np.ndarray], denoised: Sequence[np.ndarray]) -> None:
        """diff_frame(&source, &denoised): planes are numpy arrays, borrowed for the call."""
        sf, k1 = frame_from_planes(source)
        df, k2 = frame_from_planes(denoised)
        self._check(self._L.g1s_diff_push_frame(self._h, C.byref(sf), C.byref(df)))

    def host_frame(self, planes: Sequence[np.ndarray]):
        """A g1s_frame over numpy planes, built once for a caller that reuses its buffers: (struct, arrays to keep alive)."""
        return frame_from_planes(planes)

    def diff_frames_prepared_host(self, source: CFrame, denoised: CFrame) -> None:
        """g1s_diff_push_frame on two g1s_frame structs built with host_frame (the planes are borrowed for the call)."""
        rc = self._L.g1s_diff_push_frame(self._h, C.byref(source), C.byref(denoised))
        if rc != 0:
            self._check(rc)

    def device_frame(self, ptrs, strides) -> CFrame:
        """A g1s_frame over planes already resident in HBM (raw device pointers + byte strides); reusable."""
        f = CFrame()
        f.width, f.height = self.cfg.width, self.cfg.height
        for i in range(3):
            f.plane[i] = ptrs[i] if i < len(ptrs) else None
            f.stride_bytes[i] = strides[i] if i < len(strides) else 0
        return f

    def diff_frame_device(self, source_ptrs, source_strides, denoised_ptrs, denoised_strides) -> None:
        """Same as diff_frame, for planes already resident in HBM (raw device pointers + byte strides)."""
        self.diff_frames_prepared(self.device_frame(source_ptrs, source_strides), self.device_frame(denoised_ptrs, denoised_strides))

    def diff_frames_prepared(self, source: CFrame, denoised: CFrame) -> None:
        """g1s_diff_push_frame_device on two g1s_frame structs built once with device_frame (a caller that cycles over
        resident frames pays one foreign call per frame and nothing else)."""
        rc = self._L.g1s_diff_push_frame_device(self._h, C.byref(source), C.byref(denoised))
        if rc != 0:
            self._check(rc)

    def set_source_filters(self, ops, source_width: int, source_height: int) -> None:
        """ops: FilterChain.filters (Crop / Resize objects) or ("crop", t, b, l, r) / ("resize", w, h, alg) tuples.  The
        SOURCE frames pushed afterwards have source_width x source_height; the chain (run on the device) must produce
        the handle's size.  FilterChain::apply of the reference (src/main.rs:621-624)."""
        arr = (abi.CFilterOp * max(1, len(ops)))()
        for i, op in enumerate(ops):
            if not isinstance(op, tuple):
                op = ("crop", op.top, op.bottom, op.left, op.right) if hasattr(op, "top") else ("resize", op.width, op.height, op.alg)
            if op[0] == "crop":
                arr[i] = abi.CFilterOp(abi.FILTER_CROP, op[1], op[2], op[3], op[4])
            else:
                arr[i] = abi.CFilterOp(abi.FILTER_RESIZE, op[1], op[2], abi.RESIZE_ALGS[op[3]], 0)
        rc = self._L.g1s_diff_set_source_filters(self._h, arr, len(ops), source_width, source_height)
        if rc == abi.G1S_E_ARG:
            raise ValueError(self._L.g1s_diff_last_error(self._h).decode())
        self._check(rc)

    def flush(self) -> None:
        self._check(self._L.g1s_diff_flush(self._h))

    def finish(self) -> List[GrainTableSegment]:
        n = C.c_size_t(0)
        cap = 64
        while True:
            arr = (CSegment * cap)()
            rc = self._L.g1s_diff_finish(self._h, arr, cap, C.byref(n))
            if rc == abi.G1S_E_STATE and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            break
        return [GrainTableSegment.from_c(arr[i]) for i in range(n.value)]

    # -- records (multi-GPU exchange unit, tests)
    @property
    def record_bytes(self) -> int:
        return int(self._L.g1s_diff_record_bytes(self._h))

    def set_record_tap(self, fn: Optional[Callable[[int, np.ndarray], None]]) -> None:
        """fn(frame_index, record_bytes_copy) is called once per frame, in frame order."""
        if fn is None:
            self._tap = None
            self._check(self._L.g1s_diff_set_record_tap(self._h, C.cast(None, RECORD_FN), None))
            return

        def _cb(_user, idx, ptr, nbytes):
            buf = (C.c_uint8 * nbytes).from_address(ptr)
            fn(int(idx), np.frombuffer(buf, np.uint8).copy())

        self._tap = RECORD_FN(_cb)
        self._check(self._L.g1s_diff_set_record_tap(self._h, self._tap, None))

    def consume_record(self, rec: np.ndarray) -> None:
        rec = np.ascontiguousarray(rec.view(np.uint8))
        self._check(self._L.g1s_diff_consume_record(self._h, rec.ctypes.data, rec.size))

    def consume_records(self, recs: np.ndarray) -> None:
        """recs: (count, record_bytes) uint8, consecutive frames in order."""
        recs = np.ascontiguousarray(recs.view(np.uint8))
        assert recs.ndim == 2
        self._check(self._L.g1s_diff_consume_records(self._h, recs.ctypes.data, recs.shape[0], recs.strides[0]))

    # -- digests (multi-GPU exchange unit)
    def set_digest_sink(self, ptr: int, capacity_frames: int) -> None:
        """ptr: address of a buffer of capacity_frames * digest_bytes() bytes (e.g. a pinned torch tensor)."""
        self._check(self._L.g1s_diff_set_digest_sink(self._h, ptr, capacity_frames))

    @property
    def digest_count(self) -> int:
        return int(self._L.g1s_diff_digest_count(self._h))

    def wait_retired(self, frames: int) -> None:
        self._check(self._L.g1s_diff_wait_retired(self._h, frames))

    def digest_from_record(self, rec: np.ndarray) -> np.ndarray:
        rec = np.ascontiguousarray(rec.view(np.uint8))
        out = np.zeros(digest_bytes() // 8, np.float64)
        self._check(self._L.g1s_diff_digest_from_record(self._h, rec.ctypes.data, rec.size, out.ctypes.data))
        return out

    def consume_digests(self, ptr: int, count: int, borrowed: bool = False) -> None:
        """borrowed: no copy; the memory must stay valid until flush() on this handle has returned."""
        fn = self._L.g1s_diff_consume_digests_borrowed if borrowed else self._L.g1s_diff_consume_digests
        self._check(fn(self._h, ptr, count))

    @property
    def batch_frames(self) -> int:
        return int(self._L.g1s_diff_batch_frames(self._h))

    @property
    def model_on_device(self) -> bool:
        """True when the per-frame half of the noise model runs on the GPU (latest_kernel) for this handle."""
        return bool(self._L.g1s_diff_model_on_device(self._h))

    def frame_device(self, frame_index: int) -> int:
        """CUDA ordinal that processes this frame (multi-device handles deal batches round-robin)."""
        return int(self._L.g1s_diff_frame_device(self._h, frame_index))

    @property
    def frames_pushed(self) -> int:
        return int(self._L.g1s_diff_frames_pushed(self._h))

    def mark(self, which: int) -> None:
        """CUDA event on the engine's kernel stream (benchmark timing)."""
        self._check(self._L.g1s_diff_mark(self._h, which))

    def marks_elapsed_ms(self) -> float:
        return float(self._L.g1s_diff_marks_elapsed_ms(self._h))

    def counters(self) -> dict:
        out = (C.c_double * 10)()
        self._check(self._L.g1s_diff_get_counters(self._h, out, 10))
        k = ("kernels_launched", "gram_ms", "gram_launches", "flat_ms", "flat_launches", "frames_done", "tma_batches",
             "residual_ms", "vector_batches", "strict_ms")
        return dict(zip(k, [float(v) for v in out]))


def resize_table(alg: str, src: int, dst: int):
    """(left [dst] int32, coef [dst, taps] float32): the weights of one resize axis exactly as the device uses them."""
    cap = 4096
    left = np.zeros(dst, np.int32)
    coef = np.zeros(dst * cap, np.float32)
    taps = lib().g1s_resize_table(abi.RESIZE_ALGS[alg], src, dst, left.ctypes.data, coef.ctypes.data, cap)
    if taps < 0:
        raise G1SError(taps, "g1s_resize_table")
    return left, coef[: dst * taps].reshape(dst, taps).copy()


def digest_bytes() -> int:
    return int(lib().g1s_digest_bytes())


def format_grain_table(segments: Sequence[GrainTableSegment]) -> str:
    arr = segments_to_c(segments)
    need = lib().g1s_format_grain_table(arr, len(segments), None, 0)
    buf = C.create_string_buffer(int(need) + 1)
    lib().g1s_format_grain_table(arr, len(segments), buf, need + 1)
    return buf.value.decode()


def write_grain_table(segments: Sequence[GrainTableSegment], path: str) -> None:
    """`filmgrn1` header + one write_film_grain_segment per segment (src/main.rs:525-530)."""
    rc = lib().g1s_write_grain_table(segments_to_c(segments), len(segments), os.fsencode(path))
    if rc != 0:
        raise OSError(f"g1s_write_grain_table failed with {rc}")
