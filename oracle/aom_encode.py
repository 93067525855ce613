"""Real AV1 streams for the `inspect` tests: libaom 3.13.1's own ENCODER (public API of the libaom bundled with
opencv-python-headless, see oracle/aom_pin.py) driven through ctypes.

TEST INFRASTRUCTURE ONLY.  `encode()` returns the temporal units libaom emits for a short synthetic clip, with film
grain signalled either from one of libaom's 16 built-in film grain test vectors (`film-grain-test`, table
`film_grain_test_vectors` in av1/encoder/grain_test_vectors.h, read back here from the binary's .rodata), from a
`filmgrn1` table file (`film-grain-table` -- the consumer side of what grav1synth writes), or from libaom's own
denoise-and-model pass (`denoise-noise-level`).  These are encoder-produced frame headers (key frames, inter frames,
hidden alt-refs, show_existing_frame, real tile / quantiser / loop-filter / CDEF / restoration / global-motion fields),
i.e. an independent witness for the header walk of csrc/g1s_obu.cpp.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import aom_pin

AOM_IMG_FMT_I420 = 0x102
AOM_CODEC_CX_FRAME_PKT = 0


class EncodeError(RuntimeError):
    pass


def _lib():
    ok, why = aom_pin.available()
    if not ok:
        raise OSError(why)
    L = aom_pin._lib.cdll
    L.aom_codec_av1_cx.restype = C.c_void_p
    L.aom_codec_enc_config_default.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.aom_codec_enc_init_ver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    L.aom_codec_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.aom_codec_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_ulong, C.c_long]
    L.aom_codec_get_cx_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.aom_codec_get_cx_data.restype = C.c_void_p
    L.aom_codec_destroy.argtypes = [C.c_void_p]
    L.aom_codec_error_detail.argtypes = [C.c_void_p]
    L.aom_codec_error_detail.restype = C.c_char_p
    L.aom_img_alloc.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_uint]
    L.aom_img_alloc.restype = C.c_void_p
    L.aom_img_free.argtypes = [C.c_void_p]
    return L


def test_vector(index: int) -> dict:
    """film_grain_test_vectors[index - 1] as libaom holds it (aom_film_grain_t), read from the mapped binary."""
    ok, why = aom_pin.available()
    if not ok:
        raise OSError(why)
    syms = aom_pin._elf_symtab_objects(aom_pin._lib.path, ("film_grain_test_vectors",))
    base = aom_pin._load_base(aom_pin._lib.path)
    size = C.sizeof(aom_pin.FilmGrain) - 64  # without our padding
    # sizeof(aom_film_grain_t): ints up to grain_scale_shift, then uint16 random_seed (+2 padding)
    g = aom_pin.FilmGrain.from_address(base + syms["film_grain_test_vectors"] + (index - 1) * size)
    d = g.as_dict()
    d["random_seed"] = int(g.random_seed)
    d["apply_grain"] = int(g.apply_grain)
    d["update_parameters"] = int(g.update_parameters)
    d["clip_to_restricted_range"] = int(g.clip_to_restricted_range)
    return d


def synthetic_frames(n: int, w: int, h: int, seed: int = 0):
    """Moving gradient + noise, I420 8-bit: (y, u, v) numpy planes per frame."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0, 1, w, dtype=np.float32)[None, :]
    ys = np.linspace(0, 1, h, dtype=np.float32)[:, None]
    out = []
    for k in range(n):
        y = 40 + 160 * ((xs + 0.01 * k) % 1.0) * (0.6 + 0.4 * ys) + rng.normal(0, 4, (h, w))
        u = 128 + 30 * np.sin(6.28 * (xs[:, ::2] + 0.02 * k)) + 0 * ys[::2]
        v = 128 + 30 * np.cos(6.28 * (ys[::2] + 0.01 * k)) + 0 * xs[:, ::2]
        out.append(tuple(np.clip(p, 0, 255).astype(np.uint8) for p in (y, u, v)))
    return out


# aom_codec_enc_cfg_t word offsets used below (aom/aom_encoder.h)
CFG_ERROR_RESILIENT, CFG_LAG_IN_FRAMES = 12, 14
CFG_SUPERRES_MODE, CFG_SUPERRES_DENOMINATOR, CFG_SUPERRES_KF_DENOMINATOR = 19, 20, 21


AOM_IMG_FMT_I444 = 0x106
AOM_IMG_FMT_HIGHBITDEPTH = 0x800
AOM_CODEC_USE_HIGHBITDEPTH = 0x40000


def encode(frames, w: int, h: int, options: Optional[Dict[str, str]] = None, fps: int = 24,
           lag_in_frames: Optional[int] = None, cfg_words: Optional[Dict[int, int]] = None, bit_depth: int = 8,
           chroma444: bool = False) -> List[bytes]:
    """Encodes planar frames with libaom (4:2:0, or 4:4:4 = profile 1; 8-bit planes, scaled up to `bit_depth` for a
    high-bit-depth encode); returns one bytes object per emitted temporal unit, in output order."""
    L = _lib()
    iface = L.aom_codec_av1_cx()
    cfg = C.create_string_buffer(4096)
    if L.aom_codec_enc_config_default(iface, cfg, 0) != 0:
        raise EncodeError("aom_codec_enc_config_default failed")
    u32 = (C.c_uint32 * 64).from_buffer(cfg)
    # aom_codec_enc_cfg_t starts: g_usage, g_threads, g_profile, g_w, g_h, g_limit, g_forced_max_frame_width,
    # g_forced_max_frame_height, g_bit_depth, g_input_bit_depth, g_timebase{num, den}, g_error_resilient, g_pass,
    # g_lag_in_frames
    assert u32[8] == 8 and u32[9] == 8, "unexpected aom_codec_enc_cfg_t layout (bit depths)"
    u32[3], u32[4] = w, h
    u32[10], u32[11] = 1, fps
    # ... g_lag_in_frames, rc_dropframe_thresh, rc_resize_{mode, denominator, kf_denominator},
    # rc_superres_{mode, denominator, kf_denominator, qthresh, kf_qthresh}
    assert (u32[17], u32[18], u32[20], u32[21]) == (8, 8, 8, 8), "unexpected aom_codec_enc_cfg_t layout (scaling)"
    if bit_depth > 8:
        u32[8] = u32[9] = bit_depth  # g_bit_depth, g_input_bit_depth
    if chroma444:
        u32[2] = 1                   # g_profile
    if lag_in_frames is not None:
        u32[CFG_LAG_IN_FRAMES] = lag_in_frames
    for idx, val in (cfg_words or {}).items():
        u32[idx] = val
    ctx = C.create_string_buffer(512)
    rc = 3
    for ver in range(20, 60):  # AOM_ENCODER_ABI_VERSION of this build (ABI_MISMATCH = 3 until it fits)
        rc = L.aom_codec_enc_init_ver(ctx, iface, cfg, AOM_CODEC_USE_HIGHBITDEPTH if bit_depth > 8 else 0, ver)
        if rc != 3:
            break
    if rc != 0:
        raise EncodeError(f"aom_codec_enc_init_ver failed: {rc}")
    try:
        opts = {"cpu-used": "8"}
        opts.update(options or {})
        for k, v in opts.items():
            if L.aom_codec_set_option(ctx, k.encode(), str(v).encode()) != 0:
                raise EncodeError(f"option {k}={v} rejected: {L.aom_codec_error_detail(ctx)}")
        fmt = (AOM_IMG_FMT_I444 if chroma444 else AOM_IMG_FMT_I420) | (AOM_IMG_FMT_HIGHBITDEPTH if bit_depth > 8 else 0)
        img = L.aom_img_alloc(None, fmt, w, h, 32)
        if not img:
            raise EncodeError("aom_img_alloc failed")
        hdr = (C.c_uint32 * 16).from_address(img)
        assert hdr[0] == fmt and hdr[7] == w and hdr[8] == h, "unexpected aom_image_t layout"
        planes = (C.c_void_p * 3).from_address(img + 64)
        strides = (C.c_int * 3).from_address(img + 88)
        packets: List[bytes] = []

        def drain():
            it = C.c_void_p(None)
            while True:
                pkt = L.aom_codec_get_cx_data(ctx, C.byref(it))
                if not pkt:
                    return
                if C.c_int.from_address(pkt).value != AOM_CODEC_CX_FRAME_PKT:
                    continue
                buf = C.c_void_p.from_address(pkt + 8).value
                sz = C.c_size_t.from_address(pkt + 16).value
                packets.append(C.string_at(buf, sz))

        for k, (y, u, v) in enumerate(frames):
            for i, p in enumerate((y, u, v)):
                if chroma444 and i > 0 and p.shape != y.shape:
                    p = np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)[: y.shape[0], : y.shape[1]]
                ph, pw = p.shape
                if bit_depth > 8:
                    dst = np.ctypeslib.as_array((C.c_uint16 * (strides[i] // 2 * ph)).from_address(planes[i]))
                    dst.reshape(ph, strides[i] // 2)[:, :pw] = p.astype(np.uint16) << (bit_depth - 8)
                else:
                    dst = np.ctypeslib.as_array((C.c_uint8 * (strides[i] * ph)).from_address(planes[i]))
                    dst.reshape(ph, strides[i])[:, :pw] = p
            if L.aom_codec_encode(ctx, img, k, 1, 0) != 0:
                raise EncodeError(f"aom_codec_encode failed: {L.aom_codec_error_detail(ctx)}")
            drain()
        while True:  # flush
            before = len(packets)
            if L.aom_codec_encode(ctx, None, -1, 1, 0) != 0:
                raise EncodeError("flush failed")
            drain()
            if len(packets) == before:
                break
        L.aom_img_free(img)
        return packets
    finally:
        L.aom_codec_destroy(ctx)


def decode(packets: Sequence[bytes]):
    """libaom's decoder over the temporal units: list of (y, u, v) numpy planes per output frame (film grain applied,
    as every AV1 decoder does by default).  Raises EncodeError when libaom rejects the stream."""
    L = _lib()
    L.aom_codec_av1_dx.restype = C.c_void_p
    L.aom_codec_dec_init_ver.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
    L.aom_codec_decode.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
    L.aom_codec_get_frame.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.aom_codec_get_frame.restype = C.c_void_p
    ctx = C.create_string_buffer(512)
    rc = 3
    for ver in range(10, 50):
        rc = L.aom_codec_dec_init_ver(ctx, L.aom_codec_av1_dx(), None, 0, ver)
        if rc != 3:
            break
    if rc != 0:
        raise EncodeError(f"aom_codec_dec_init_ver failed: {rc}")
    frames = []
    try:
        for k, pk in enumerate(packets):
            if L.aom_codec_decode(ctx, pk, len(pk), None) != 0:
                raise EncodeError(f"libaom cannot decode packet {k}: {L.aom_codec_error_detail(ctx)}")
            it = C.c_void_p(None)
            while True:
                img = L.aom_codec_get_frame(ctx, C.byref(it))
                if not img:
                    break
                hdr = (C.c_uint32 * 16).from_address(img)
                w, h = hdr[10], hdr[11]  # d_w, d_h
                planes = (C.c_void_p * 3).from_address(img + 64)
                strides = (C.c_int * 3).from_address(img + 88)
                out = []
                for i in range(3):
                    pw, ph = (w, h) if i == 0 else ((w + 1) // 2, (h + 1) // 2)
                    a = np.ctypeslib.as_array((C.c_uint8 * (strides[i] * ph)).from_address(planes[i]))
                    out.append(a.reshape(ph, strides[i])[:, :pw].copy())
                frames.append(tuple(out))
        return frames
    finally:
        L.aom_codec_destroy(ctx)
