"""Pin for the oracle: libaom's own noise_model.c, executed from a binary that ships in this image.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and by
tests/golden/make_aom_golden.py, never by the product package.

What this is.  The arithmetic behind grav1synth's `diff` lives in the crate av1-grain 0.4.2
(module `diff`), which is a Rust port of libaom `aom_dsp/noise_model.c` (+ `mathutils.h::linsolve`);
neither the crate nor a Rust toolchain is on this box.  The image does carry a
compiled libaom 3.13.1 -- bundled with opencv-python-headless as
`opencv_python_headless.libs/libaom-*.so.3.13.1` -- whose dynamic symbol table hides the noise-model
entry points but whose full `.symtab` was left in place.  This module resolves
`aom_flat_block_finder_{init,run,free}`, `aom_noise_model_{init,update,save_latest,get_grain_parameters,free}`
from that table and calls them in place, in the order of libaom's `examples/noise_model.c` main loop
-- which is the loop `av1_grain::DiffGenerator::diff_frame/finish` restates and the reference drives from
/root/reference/src/main.rs:420-524.  It is therefore an executable UPSTREAM of the reference's
dependency, not the reference itself: it pins the oracle's flat-block finder, AR normal equations,
strength solver, segment logic and `get_grain_parameters` for 8-bit input (the crate reduces every
input to 8 bits first, `frame_into_u8`), and leaves unpinned only what the crate adds around it
(the `>> (bd-8)` reduction, the timestamp rule, the fixed seed, error swallowing).

Struct layouts below are libaom's public headers (aom_dsp/noise_model.h, aom_dsp/grain_params.h).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BLOCK_SIZE = 32
AOM_NOISE_SHAPE_SQUARE = 1
STATUS_OK, STATUS_INVALID_ARGUMENT, STATUS_INSUFFICIENT_FLAT_BLOCKS, STATUS_DIFFERENT_NOISE_TYPE, \
    STATUS_INTERNAL_ERROR = range(5)

_WANTED = (
    "aom_flat_block_finder_init", "aom_flat_block_finder_run", "aom_flat_block_finder_free",
    "aom_noise_model_init", "aom_noise_model_update", "aom_noise_model_save_latest",
    "aom_noise_model_get_grain_parameters", "aom_noise_model_free",
)


def find_libaom() -> Optional[str]:
    """Path of the bundled libaom, or None (then every pin test skips with the reason)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("cv2")
        roots = [os.path.dirname(os.path.dirname(spec.origin))] if spec and spec.origin else []
    except Exception:
        roots = []
    import sysconfig
    roots.append(sysconfig.get_paths()["purelib"])
    for r in roots:
        hits = sorted(glob.glob(os.path.join(r, "opencv_python_headless.libs", "libaom-*.so*")))
        if hits:
            return hits[0]
    return None


def _elf_symtab(path: str, wanted: Sequence[str]) -> Dict[str, int]:
    """Minimal ELF64 .symtab reader: name -> st_value for FUNC symbols in `wanted`."""
    out: Dict[str, int] = {}
    with open(path, "rb") as f:
        eh = f.read(64)
        if eh[:4] != b"\x7fELF" or eh[4] != 2 or eh[5] != 1:
            raise OSError("not a little-endian ELF64 file")
        e_shoff, = struct.unpack_from("<Q", eh, 0x28)
        e_shentsize, e_shnum, _ = struct.unpack_from("<HHH", eh, 0x3A)
        f.seek(e_shoff)
        sh = f.read(e_shentsize * e_shnum)
        secs = [struct.unpack_from("<IIQQQQIIQQ", sh, i * e_shentsize) for i in range(e_shnum)]
        for (_, sh_type, _, _, sh_offset, sh_size, sh_link, _, _, sh_entsize) in secs:
            if sh_type != 2:  # SHT_SYMTAB
                continue
            str_off, str_size = secs[sh_link][4], secs[sh_link][5]
            f.seek(str_off)
            strtab = f.read(str_size)
            f.seek(sh_offset)
            sym = f.read(sh_size)
            want = {w.encode(): w for w in wanted}
            for i in range(sh_size // sh_entsize):
                st_name, st_info, _, st_shndx, st_value, _ = struct.unpack_from("<IBBHQQ", sym, i * sh_entsize)
                if (st_info & 0xF) != 2 or st_shndx == 0:  # STT_FUNC, defined
                    continue
                end = strtab.index(b"\0", st_name)
                nm = strtab[st_name:end]
                if nm in want:
                    out[want[nm]] = st_value
    return out


def _elf_symtab_objects(path: str, wanted: Sequence[str]) -> Dict[str, int]:
    """Like _elf_symtab, for data objects (STT_OBJECT)."""
    out: Dict[str, int] = {}
    with open(path, "rb") as f:
        eh = f.read(64)
        e_shoff, = struct.unpack_from("<Q", eh, 0x28)
        e_shentsize, e_shnum, _ = struct.unpack_from("<HHH", eh, 0x3A)
        f.seek(e_shoff)
        sh = f.read(e_shentsize * e_shnum)
        secs = [struct.unpack_from("<IIQQQQIIQQ", sh, i * e_shentsize) for i in range(e_shnum)]
        for (_, sh_type, _, _, sh_offset, sh_size, sh_link, _, _, sh_entsize) in secs:
            if sh_type != 2:
                continue
            f.seek(secs[sh_link][4])
            strtab = f.read(secs[sh_link][5])
            f.seek(sh_offset)
            sym = f.read(sh_size)
            want = {w.encode(): w for w in wanted}
            for i in range(sh_size // sh_entsize):
                st_name, st_info, _, st_shndx, st_value, _ = struct.unpack_from("<IBBHQQ", sym, i * sh_entsize)
                if (st_info & 0xF) != 1 or st_shndx == 0:  # STT_OBJECT, defined
                    continue
                nm = strtab[st_name:strtab.index(b"\0", st_name)]
                if nm in want:
                    out[want[nm]] = st_value
    return out


def _load_base(path: str) -> int:
    real = os.path.realpath(path)
    with open("/proc/self/maps") as m:
        for line in m:
            parts = line.split()
            if len(parts) >= 6 and os.path.realpath(parts[5]) == real and int(parts[2], 16) == 0:
                return int(parts[0].split("-")[0], 16)
    raise OSError("libaom is not mapped")


class _Params(C.Structure):  # aom_noise_model_params_t, passed by value
    _fields_ = [("shape", C.c_int), ("lag", C.c_int), ("bit_depth", C.c_int), ("use_highbd", C.c_int)]


class FilmGrain(C.Structure):  # aom_film_grain_t (aom_dsp/grain_params.h)
    _fields_ = [
        ("apply_grain", C.c_int), ("update_parameters", C.c_int),
        ("scaling_points_y", (C.c_int * 2) * 14), ("num_y_points", C.c_int),
        ("scaling_points_cb", (C.c_int * 2) * 10), ("num_cb_points", C.c_int),
        ("scaling_points_cr", (C.c_int * 2) * 10), ("num_cr_points", C.c_int),
        ("scaling_shift", C.c_int), ("ar_coeff_lag", C.c_int),
        ("ar_coeffs_y", C.c_int * 24), ("ar_coeffs_cb", C.c_int * 25), ("ar_coeffs_cr", C.c_int * 25),
        ("ar_coeff_shift", C.c_int),
        ("cb_mult", C.c_int), ("cb_luma_mult", C.c_int), ("cb_offset", C.c_int),
        ("cr_mult", C.c_int), ("cr_luma_mult", C.c_int), ("cr_offset", C.c_int),
        ("overlap_flag", C.c_int), ("clip_to_restricted_range", C.c_int), ("bit_depth", C.c_uint),
        ("chroma_scaling_from_luma", C.c_int), ("grain_scale_shift", C.c_int), ("random_seed", C.c_uint16),
        ("pad_", C.c_uint8 * 64),
    ]

    def as_dict(self) -> dict:
        ny, ncb, ncr = self.num_y_points, self.num_cb_points, self.num_cr_points
        return dict(
            scaling_points_y=[tuple(self.scaling_points_y[i]) for i in range(ny)],
            scaling_points_cb=[tuple(self.scaling_points_cb[i]) for i in range(ncb)],
            scaling_points_cr=[tuple(self.scaling_points_cr[i]) for i in range(ncr)],
            scaling_shift=self.scaling_shift, ar_coeff_lag=self.ar_coeff_lag, ar_coeff_shift=self.ar_coeff_shift,
            ar_coeffs_y=list(self.ar_coeffs_y), ar_coeffs_cb=list(self.ar_coeffs_cb),
            ar_coeffs_cr=list(self.ar_coeffs_cr),
            cb_mult=self.cb_mult, cb_luma_mult=self.cb_luma_mult, cb_offset=self.cb_offset,
            cr_mult=self.cr_mult, cr_luma_mult=self.cr_luma_mult, cr_offset=self.cr_offset,
            overlap_flag=self.overlap_flag, chroma_scaling_from_luma=self.chroma_scaling_from_luma,
            grain_scale_shift=self.grain_scale_shift,
        )


class _Lib:
    def __init__(self, path: str):
        self.path = path
        self.cdll = C.CDLL(path)  # maps the library; hidden functions are then reached by address
        syms = _elf_symtab(path, _WANTED)
        missing = [w for w in _WANTED if w not in syms]
        if missing:
            raise OSError("libaom .symtab lacks " + ", ".join(missing))
        base = _load_base(path)
        P, I = C.c_void_p, C.c_int
        proto = {
            "aom_flat_block_finder_init": C.CFUNCTYPE(I, P, I, I, I),
            "aom_flat_block_finder_run": C.CFUNCTYPE(I, P, P, I, I, I, P),
            "aom_flat_block_finder_free": C.CFUNCTYPE(None, P),
            "aom_noise_model_init": C.CFUNCTYPE(I, P, _Params),
            "aom_noise_model_update": C.CFUNCTYPE(I, P, P, P, I, I, P, P, P, I),
            "aom_noise_model_save_latest": C.CFUNCTYPE(None, P),
            "aom_noise_model_get_grain_parameters": C.CFUNCTYPE(I, P, C.POINTER(FilmGrain)),
            "aom_noise_model_free": C.CFUNCTYPE(None, P),
        }
        for name, ft in proto.items():
            setattr(self, name[4:], ft(base + syms[name]))


_lib: Optional[_Lib] = None
_lib_error: Optional[str] = None


def available() -> Tuple[bool, str]:
    """(True, path) when the bundled libaom and its hidden symbols can be used, else (False, reason)."""
    global _lib, _lib_error
    if _lib is not None:
        return True, _lib.path
    if _lib_error is not None:
        return False, _lib_error
    path = find_libaom()
    if path is None:
        _lib_error = "no opencv_python_headless.libs/libaom-*.so in this environment"
        return False, _lib_error
    try:
        _lib = _Lib(path)
    except OSError as e:
        _lib_error = str(e)
        return False, _lib_error
    return True, path


def to_u8(plane: np.ndarray, bit_depth: int) -> np.ndarray:
    """The crate's frame_into_u8 (truncating shift), applied before libaom sees the data."""
    if plane.dtype == np.uint8:
        return np.ascontiguousarray(plane)
    return np.ascontiguousarray((plane.astype(np.uint16) >> (bit_depth - 8)).astype(np.uint8))


class AomNoiseModel:
    """libaom's noise model driven like examples/noise_model.c (8-bit, lag 3, square shape, 32x32 blocks).

    update(source, denoised) -> status; on DIFFERENT_NOISE_TYPE the parameters of the finished segment are
    appended to `segments` (get_grain_parameters, then save_latest), exactly the example's loop; finish()
    appends the parameters of the last segment.  Planes are numpy uint8 arrays (3 of them, or 1 for monochrome
    is NOT supported by libaom's update: it always walks 3 channels)."""

    def __init__(self, ss_x: int = 1, ss_y: int = 1):
        ok, why = available()
        if not ok:
            raise OSError(why)
        self.L = _lib
        self.ss = (C.c_int * 2)(ss_x, ss_y)
        self.finder = C.create_string_buffer(256)
        self.model = C.create_string_buffer(8192)
        if not self.L.flat_block_finder_init(C.addressof(self.finder), BLOCK_SIZE, 8, 0):
            raise OSError("aom_flat_block_finder_init failed")
        if not self.L.noise_model_init(C.addressof(self.model), _Params(AOM_NOISE_SHAPE_SQUARE, 3, 8, 0)):
            raise OSError("aom_noise_model_init failed")
        self.segments: List[dict] = []
        self.segment_first_frame: List[int] = []
        self.statuses: List[int] = []
        self.flat: Optional[np.ndarray] = None
        self.num_flat = 0
        self._start = 0
        self._frames = 0
        self._closed = False

    def close(self):
        if not self._closed:
            self.L.noise_model_free(C.addressof(self.model))
            self.L.flat_block_finder_free(C.addressof(self.finder))
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def grain_parameters(self) -> dict:
        g = FilmGrain()
        if not self.L.noise_model_get_grain_parameters(C.addressof(self.model), C.byref(g)):
            raise OSError("aom_noise_model_get_grain_parameters failed")
        return g.as_dict()

    def state(self, which: str, c: int) -> dict:
        """Internal f64 state of channel c (`which` = "latest" | "combined"), read through libaom's public
        struct layout (aom_noise_model_t: params 16 B, combined_state[3], latest_state[3]; aom_noise_state_t =
        {eqns{A,b,x,n}, strength_solver{eqns, min, max, num_bins, num_equations, total}, num_observations, ar_gain},
        112 B)."""
        base = C.addressof(self.model) + 16 + (336 if which == "latest" else 0) + 112 * c
        buf = bytes((C.c_uint8 * 112).from_address(base))
        _, _, x, n = struct.unpack_from("<QQQi", buf, 0)
        _, _, sx, sn = struct.unpack_from("<QQQi", buf, 32)
        _, _, num_bins, num_eq, total = struct.unpack_from("<ddiid", buf, 64)
        nobs, = struct.unpack_from("<i", buf, 96)
        gain, = struct.unpack_from("<d", buf, 104)
        assert n in (24, 25) and sn == 20 and num_bins == 20, "unexpected aom_noise_model_t layout"
        return dict(x=np.ctypeslib.as_array((C.c_double * n).from_address(x)).copy(),
                    strength_x=np.ctypeslib.as_array((C.c_double * sn).from_address(sx)).copy(),
                    ar_gain=float(gain), num_observations=int(nobs), num_equations=int(num_eq), total=float(total))

    def update(self, source: Sequence[np.ndarray], denoised: Sequence[np.ndarray]) -> int:
        src = [np.ascontiguousarray(p) for p in source]
        den = [np.ascontiguousarray(p) for p in denoised]
        assert all(p.dtype == np.uint8 for p in src + den) and len(src) == 3 and len(den) == 3
        h, w = src[0].shape
        nb = ((w + BLOCK_SIZE - 1) // BLOCK_SIZE) * ((h + BLOCK_SIZE - 1) // BLOCK_SIZE)
        flat = np.zeros(nb, np.uint8)
        self.num_flat = self.L.flat_block_finder_run(C.addressof(self.finder), src[0].ctypes.data, w, h,
                                                     src[0].strides[0], flat.ctypes.data)
        self.flat = flat
        data = (C.c_void_p * 3)(*[p.ctypes.data for p in src])
        dd = (C.c_void_p * 3)(*[p.ctypes.data for p in den])
        # libaom uses strides[c] for both the source and the denoised plane of channel c
        for s, d in zip(src, den):
            assert s.strides[0] == d.strides[0]
        strides = (C.c_int * 3)(*[p.strides[0] for p in src])
        st = self.L.noise_model_update(C.addressof(self.model), data, dd, w, h, strides, self.ss,
                                       flat.ctypes.data, BLOCK_SIZE)
        self.statuses.append(st)
        if st == STATUS_DIFFERENT_NOISE_TYPE:
            self.segments.append(self.grain_parameters())
            self.segment_first_frame.append(self._start)
            self.L.noise_model_save_latest(C.addressof(self.model))
            self._start = self._frames
        self._frames += 1
        return st

    def finish(self) -> List[dict]:
        self.segments.append(self.grain_parameters())
        self.segment_first_frame.append(self._start)
        return self.segments


def segment_as_dict(seg) -> dict:
    """The same view of one of OUR GrainTableSegment objects (oracle or CUDA engine), for equality tests."""
    return dict(
        scaling_points_y=[tuple(p) for p in seg.scaling_points_y],
        scaling_points_cb=[tuple(p) for p in seg.scaling_points_cb],
        scaling_points_cr=[tuple(p) for p in seg.scaling_points_cr],
        scaling_shift=seg.scaling_shift, ar_coeff_lag=seg.ar_coeff_lag, ar_coeff_shift=seg.ar_coeff_shift,
        ar_coeffs_y=list(seg.ar_coeffs_y), ar_coeffs_cb=list(seg.ar_coeffs_cb), ar_coeffs_cr=list(seg.ar_coeffs_cr),
        cb_mult=seg.cb_mult, cb_luma_mult=seg.cb_luma_mult, cb_offset=seg.cb_offset,
        cr_mult=seg.cr_mult, cr_luma_mult=seg.cr_luma_mult, cr_offset=seg.cr_offset,
        overlap_flag=seg.overlap_flag, chroma_scaling_from_luma=seg.chroma_scaling_from_luma,
        grain_scale_shift=seg.grain_scale_shift,
    )
