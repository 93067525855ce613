//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Rust toolchain).  `DiffGenerator` with the three
//! signatures grav1synth uses from `av1_grain::DiffGenerator` (src/main.rs:420-427, :442, :524), over
//! the C ABI of include/g1s.h.  Swapping `use av1_grain::DiffGenerator` for
//! `use grav1synth_cuda::DiffGenerator` in src/main.rs:18-21 is the whole integration.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use anyhow::{ensure, Result};
use arrayvec::ArrayVec;
use av1_grain::GrainTableSegment;
use num_rational::Rational64;
use v_frame::{frame::Frame, pixel::Pixel, plane::Plane};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct g1s_segment {
    pub start_time: u64,
    pub end_time: u64,
    pub num_y_points: u8,
    pub num_cb_points: u8,
    pub num_cr_points: u8,
    pub scaling_shift: u8,
    pub ar_coeff_lag: u8,
    pub ar_coeff_shift: u8,
    pub grain_scale_shift: u8,
    pub overlap_flag: u8,
    pub chroma_scaling_from_luma: u8,
    pub cb_mult: u8,
    pub cb_luma_mult: u8,
    pub cr_mult: u8,
    pub cr_luma_mult: u8,
    pub num_ar_coeffs_plus1: [u8; 3],
    pub cb_offset: u16,
    pub cr_offset: u16,
    pub random_seed: u16,
    pub clip_to_restricted_range: u16,
    pub scaling_points_y: [[u8; 2]; 14],
    pub scaling_points_cb: [[u8; 2]; 10],
    pub scaling_points_cr: [[u8; 2]; 10],
    pub ar_coeffs_y: [i8; 24],
    pub ar_coeffs_cb: [i8; 25],
    pub ar_coeffs_cr: [i8; 25],
}

#[repr(C)]
pub struct g1s_frame {
    pub plane: [*const c_void; 3],
    pub stride_bytes: [usize; 3],
    pub width: i32,
    pub height: i32,
}

#[repr(C)]
pub struct g1s_diff_config {
    pub fps_num: i64,
    pub fps_den: i64,
    pub src_bit_depth: i32,
    pub den_bit_depth: i32,
    pub width: i32,
    pub height: i32,
    pub ss_x: i32,
    pub ss_y: i32,
    pub monochrome: i32,
    pub device: i32,
    pub batch_frames: i32,
    pub mode: i32,
    pub gram_kernel: i32,
    pub host_threads: i32,
    pub host_narrow: i32,
    /// 0 = exact-integer Gram (fast), 1 = reference accumulation order (strict: av1-grain's integers on every stream)
    pub gram_order: i32,
    /// 0 / 1: one GPU (`device`); 2..8: this handle drives `device_ids[0..n_devices]`
    pub n_devices: i32,
    pub device_ids: [i32; 8],
    /// 0 = auto, 1 = per-frame model half on host threads, 2 = on the device (latest_kernel)
    pub model_placement: i32,
    pub reserved_: [i32; 3],
}

pub enum g1s_diff {}

#[link(name = "g1s")]
extern "C" {
    fn g1s_diff_create(cfg: *const g1s_diff_config, out: *mut *mut g1s_diff) -> c_int;
    fn g1s_diff_push_frame(d: *mut g1s_diff, src: *const g1s_frame, den: *const g1s_frame) -> c_int;
    fn g1s_diff_finish(d: *mut g1s_diff, out: *mut g1s_segment, cap: usize, n: *mut usize) -> c_int;
    fn g1s_diff_destroy(d: *mut g1s_diff);
    fn g1s_diff_last_error(d: *const g1s_diff) -> *const c_char;
}

/// The CPU-only bitstream side (`inspect`, `apply`, `remove`, `generate`): see INTEGRATION.md sections 6 and 7 for where
/// these replace `BitstreamParser::get_grain_headers` / `modify_grain_headers` and `aggregate_grain_headers`.
pub enum g1s_inspect {}
extern "C" {
    pub fn g1s_inspect_create(out: *mut *mut g1s_inspect) -> c_int;
    pub fn g1s_inspect_push_packet(h: *mut g1s_inspect, data: *const u8, size: usize) -> c_int;
    pub fn g1s_inspect_num_headers(h: *const g1s_inspect) -> usize;
    pub fn g1s_inspect_header(h: *const g1s_inspect, i: usize, kind: *mut i32, params: *mut g1s_segment) -> c_int;
    pub fn g1s_inspect_finish(h: *mut g1s_inspect, fps_num: i64, fps_den: i64, out: *mut g1s_segment, cap: usize,
                              n: *mut usize) -> c_int;
    pub fn g1s_inspect_last_error(h: *const g1s_inspect) -> *const c_char;
    pub fn g1s_inspect_destroy(h: *mut g1s_inspect);
    pub fn g1s_rewrite_create(table: *const g1s_segment, n: usize, apply: c_int, out: *mut *mut g1s_inspect) -> c_int;
    pub fn g1s_rewrite_packet(h: *mut g1s_inspect, data: *const u8, size: usize, packet_ts: u64,
                              out_size: *mut usize) -> c_int;
    pub fn g1s_rewrite_take(h: *mut g1s_inspect, out: *mut u8, cap: usize) -> c_int;
    pub fn g1s_generate_photon_noise(iso: u32, width: u32, height: u32, transfer: c_int, chroma_grain: c_int, full_range: c_int,
                                     random_seed: i32, start_time: u64, end_time: u64, out: *mut g1s_segment) -> c_int;
}

fn last_error(d: *const g1s_diff) -> String {
    unsafe { CStr::from_ptr(g1s_diff_last_error(d)) }.to_string_lossy().into_owned()
}

fn borrow_plane<T: Pixel>(p: &Plane<T>) -> (*const c_void, usize) {
    (p.data_origin().as_ptr().cast(), p.geometry().stride.get() * std::mem::size_of::<T>())
}

fn borrow<T: Pixel>(f: &Frame<T>) -> g1s_frame {
    let (y, ys) = borrow_plane(&f.y_plane);
    let (u, us) = f.u_plane.as_ref().map(borrow_plane).unwrap_or((std::ptr::null(), 0));
    let (v, vs) = f.v_plane.as_ref().map(borrow_plane).unwrap_or((std::ptr::null(), 0));
    g1s_frame {
        plane: [y, u, v],
        stride_bytes: [ys, us, vs],
        width: f.y_plane.width().get() as i32,
        height: f.y_plane.height().get() as i32,
    }
}

pub struct DiffGenerator {
    h: *mut g1s_diff,
    fps: Rational64,
    source_bit_depth: usize,
    denoised_bit_depth: usize,
}

impl DiffGenerator {
    /// src/main.rs:420-427.  The engine sizes its buffers from the first frame.
    #[must_use]
    pub fn new(fps: Rational64, source_bit_depth: usize, denoised_bit_depth: usize) -> Self {
        Self { h: std::ptr::null_mut(), fps, source_bit_depth, denoised_bit_depth }
    }

    /// src/main.rs:442 — `differ.diff_frame(&source_frame, &denoised_frame)?`
    pub fn diff_frame<T: Pixel, U: Pixel>(&mut self, source: &Frame<T>, denoised: &Frame<U>) -> Result<()> {
        if self.h.is_null() {
            let (ss_x, ss_y) = source.subsampling.subsample_ratio().map_or((0, 0), |(x, y)| (x.get() >> 1, y.get() >> 1));
            let cfg = g1s_diff_config {
                fps_num: *self.fps.numer(),
                fps_den: *self.fps.denom(),
                src_bit_depth: self.source_bit_depth as i32,
                den_bit_depth: self.denoised_bit_depth as i32,
                width: source.y_plane.width().get() as i32,
                height: source.y_plane.height().get() as i32,
                ss_x: ss_x as i32,
                ss_y: ss_y as i32,
                monochrome: i32::from(source.u_plane.is_none()),
                device: 0,
                batch_frames: 0,
                mode: 0,
                gram_kernel: 0,
                host_threads: 0,
                // v_frame planes are pageable: the staging threads reduce samples deeper than 8 bits to the 8 bits the
                // path reads (frame_into_u8), so half the bytes cross PCIe; identical tables
                host_narrow: i32::from(self.source_bit_depth > 8 || self.denoised_bit_depth > 8),
                gram_order: std::env::var("G1S_STRICT").map(|v| (v == "1") as i32).unwrap_or(0),
                n_devices: 0,
                device_ids: [0; 8],
                model_placement: 0,
                reserved_: [0; 3],
            };
            let rc = unsafe { g1s_diff_create(&cfg, &mut self.h) };
            ensure!(rc == 0, "g1s_diff_create failed: {}", last_error(std::ptr::null()));
        }
        let (fs, fd) = (borrow(source), borrow(denoised));
        let rc = unsafe { g1s_diff_push_frame(self.h, &fs, &fd) };
        ensure!(rc == 0, "{}", last_error(self.h)); // G1S_E_DIMS is verify_dimensions_match's error
        Ok(())
    }

    /// src/main.rs:524 — `differ.finish()`
    #[must_use]
    pub fn finish(self) -> Vec<GrainTableSegment> {
        let mut n = 0usize;
        let mut segs = vec![unsafe { std::mem::zeroed::<g1s_segment>() }; 64];
        let mut rc = unsafe { g1s_diff_finish(self.h, segs.as_mut_ptr(), segs.len(), &mut n) };
        if rc == -6 {
            segs.resize(n, segs[0]);
            rc = unsafe { g1s_diff_finish(self.h, segs.as_mut_ptr(), n, &mut n) };
        }
        assert_eq!(rc, 0, "{}", last_error(self.h));
        segs.truncate(n);
        segs.iter().map(to_av1_grain).collect()
    }
}

impl Drop for DiffGenerator {
    fn drop(&mut self) {
        unsafe { g1s_diff_destroy(self.h) }
    }
}

fn to_av1_grain(s: &g1s_segment) -> GrainTableSegment {
    GrainTableSegment {
        start_time: s.start_time,
        end_time: s.end_time,
        scaling_points_y: s.scaling_points_y[..s.num_y_points as usize].iter().copied().collect::<ArrayVec<_, 14>>(),
        scaling_points_cb: s.scaling_points_cb[..s.num_cb_points as usize].iter().copied().collect::<ArrayVec<_, 10>>(),
        scaling_points_cr: s.scaling_points_cr[..s.num_cr_points as usize].iter().copied().collect::<ArrayVec<_, 10>>(),
        scaling_shift: s.scaling_shift,
        ar_coeff_lag: s.ar_coeff_lag,
        ar_coeffs_y: s.ar_coeffs_y.iter().copied().collect(),
        ar_coeffs_cb: s.ar_coeffs_cb.iter().copied().collect(),
        ar_coeffs_cr: s.ar_coeffs_cr.iter().copied().collect(),
        ar_coeff_shift: s.ar_coeff_shift,
        cb_mult: s.cb_mult,
        cb_luma_mult: s.cb_luma_mult,
        cb_offset: s.cb_offset,
        cr_mult: s.cr_mult,
        cr_luma_mult: s.cr_luma_mult,
        cr_offset: s.cr_offset,
        overlap_flag: s.overlap_flag != 0,
        chroma_scaling_from_luma: s.chroma_scaling_from_luma != 0,
        grain_scale_shift: s.grain_scale_shift,
        random_seed: s.random_seed,
    }
}
