"""Seeded synthetic source/denoised YUV frame pairs.

There is no network, no FFmpeg and no test clip in this environment, so every test
and benchmark input is generated: a smooth `denoised` picture (luma ramp spanning the
8-bit range so all 20 strength bins are populated, low-frequency chroma, an optional
fraction of high-contrast textured 32x32 blocks that must fail the flat-block test)
and `source = clip(denoised + grain)` where the grain is luma-dependent Gaussian noise
shaped by a small spatial filter (so the fitted AR coefficients are non-trivial) with
a luma-correlated chroma component.  Planar layout as produced by the reference's
decode_frame (/root/reference/src/reader.rs:172-212): uint8 for 8-bit, uint16 above.

torch is used so the same code generates on the CPU (tests) and directly in HBM
(benchmark); the generator is plumbing, not part of the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np
import torch


@dataclass
class SynthSpec:
    width: int
    height: int
    bit_depth: int = 8
    ss_x: int = 1
    ss_y: int = 1
    textured: float = 0.1      # fraction of 32x32 luma blocks given a texture
    sigma0: float = 2.0        # grain sigma at luma 0 (8-bit units)
    sigma1: float = 6.0        # additional sigma at luma 255
    ar_strength: float = 0.25  # spatial correlation of the grain
    chroma_scale: float = 0.5
    chroma_luma_corr: float = 0.3
    seed: int = 20260917


def _smooth_field(h, w, gen, device, lo, hi, fy=1.0, fx=1.0, phase=0.0):
    ys = torch.linspace(0, 1, h, device=device, dtype=torch.float64).view(h, 1)
    xs = torch.linspace(0, 1, w, device=device, dtype=torch.float64).view(1, w)
    v = 0.5 + 0.5 * torch.cos(2 * np.pi * (fy * ys + fx * xs) + phase)
    return lo + (hi - lo) * v


def make_pair(spec: SynthSpec, frame_index: int = 0, device: str = "cpu"):
    """Returns (source_planes, denoised_planes): lists of 3 torch tensors (uint8/int16 bit pattern
    of uint16), each HxW for luma and the ceil-subsampled size for chroma."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(spec.seed + frame_index)
    w, h = spec.width, spec.height
    cw, ch = (w + spec.ss_x) >> spec.ss_x, (h + spec.ss_y) >> spec.ss_y
    f64 = torch.float64

    xs = torch.linspace(0, 1, w, device=dev, dtype=f64).view(1, w)
    ys = torch.linspace(0, 1, h, device=dev, dtype=f64).view(h, 1)
    luma = 16.0 + (235.0 - 16.0) * xs + 6.0 * torch.cos(2 * np.pi * ys * 1.5 + 0.1 * frame_index)
    luma = luma.expand(h, w).clone()
    if spec.textured > 0:
        nbw, nbh = (w + 31) // 32, (h + 31) // 32
        pick = torch.rand((nbh, nbw), generator=gen, device=dev) < spec.textured
        mask = pick.repeat_interleave(32, 0).repeat_interleave(32, 1)[:h, :w]
        yy = torch.arange(h, device=dev).view(h, 1)
        xx = torch.arange(w, device=dev).view(1, w)
        checker = (((yy // 4) + (xx // 4)) % 2).to(f64) * 80.0 - 40.0
        luma = torch.where(mask, luma + checker, luma)
    luma = luma.clamp(0, 255)
    cb = _smooth_field(ch, cw, gen, dev, 108.0, 148.0, 0.7, 0.4, 0.3)
    cr = _smooth_field(ch, cw, gen, dev, 108.0, 148.0, 0.3, 0.9, 1.1)

    def shaped_noise(hh, ww):
        n = torch.randn((1, 1, hh + 4, ww + 4), generator=gen, device=dev, dtype=f64)
        a = spec.ar_strength
        k = torch.tensor([[0.0, a * 0.5, 0.0], [a, 1.0, a * 0.3], [0.0, a * 0.2, 0.0]], device=dev, dtype=f64)
        k = k / k.pow(2).sum().sqrt()
        out = torch.nn.functional.conv2d(n, k.view(1, 1, 3, 3))
        return out[0, 0, 1:hh + 1, 1:ww + 1]

    sigma = spec.sigma0 + spec.sigma1 * luma / 255.0
    gy = shaped_noise(h, w) * sigma
    # chroma grain: own component + a part correlated with the co-sited luma grain
    gy_pad = torch.nn.functional.pad(gy.view(1, 1, h, w), (0, cw * (1 << spec.ss_x) - w, 0, ch * (1 << spec.ss_y) - h),
                                     mode="replicate")
    gy_ds = torch.nn.functional.avg_pool2d(gy_pad, (1 << spec.ss_y, 1 << spec.ss_x))[0, 0]
    sig_c = spec.chroma_scale * (spec.sigma0 + 0.5 * spec.sigma1)
    gcb = shaped_noise(ch, cw) * sig_c + spec.chroma_luma_corr * gy_ds
    gcr = shaped_noise(ch, cw) * sig_c + spec.chroma_luma_corr * gy_ds

    scale = float(1 << (spec.bit_depth - 8))
    maxv = float((1 << spec.bit_depth) - 1)

    def quant(p, dither):
        v = p * scale
        if dither and spec.bit_depth > 8:
            v = v + torch.rand(p.shape, generator=gen, device=dev, dtype=f64) * (scale - 1)
        v = v.round().clamp(0, maxv)
        if spec.bit_depth == 8:
            return v.to(torch.uint8)
        return v.to(torch.int32).to(torch.int16)  # bit pattern of uint16 (values <= 65535 wrap into int16)

    den = [quant(luma, True), quant(cb, True), quant(cr, True)]
    src = [quant(luma + gy, False), quant(cb + gcb, False), quant(cr + gcr, False)]
    return src, den


def to_numpy(planes) -> List[np.ndarray]:
    out = []
    for p in planes:
        a = p.cpu().numpy()
        if a.dtype == np.int16:
            a = a.view(np.uint16)
        out.append(np.ascontiguousarray(a))
    return out


def make_pair_numpy(spec: SynthSpec, frame_index: int = 0) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    s, d = make_pair(spec, frame_index, "cpu")
    return to_numpy(s), to_numpy(d)


def frame_pair_bytes(width: int, height: int, ss_x: int, ss_y: int, src_bd: int, den_bd: int) -> int:
    """Algorithmic bytes per frame pair: one read of every sample of both frames."""
    cw, ch = (width + ss_x) >> ss_x, (height + ss_y) >> ss_y
    samples = width * height + 2 * cw * ch
    return samples * ((1 if src_bd == 8 else 2) + (1 if den_bd == 8 else 2))
