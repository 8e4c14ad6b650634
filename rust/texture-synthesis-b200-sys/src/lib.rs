//! Raw bindings of include/tsb200.h.  Uncompiled in this repository: no Rust toolchain in the image.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct tsb_params {            // GeneratorParams, ms.rs:18-42
    pub nearest_neighbors: u32, _pad0: u32,
    pub random_sample_locations: u64,
    pub cauchy_dispersion: f32, pub p: f32,
    pub p_stages: i32, pub alpha: f32,
    pub seed: u64, pub max_thread_count: u64,
    pub tiling_mode: i32, _pad1: i32,
}
#[repr(C)] pub struct tsb_image    { pub rgba: *const u8, pub width: u32, pub height: u32 }
#[repr(C)] pub struct tsb_pyramid  { pub levels: *const u8, pub width: u32, pub height: u32, pub n_levels: u32 }
#[repr(C)] pub struct tsb_sampling { pub kind: i32, _pad: i32, pub rgba: *const u8 }
#[repr(C)] pub struct tsb_guides   { pub target: tsb_pyramid, pub examples: *const tsb_pyramid, pub n_examples: u32, _pad: u32 }
#[repr(C)] pub struct tsb_generator_desc {
    pub out_width: u32, pub out_height: u32,
    pub inpaint_mask: *const u8, pub inpaint_color: *const u8,
    pub inpaint_example_index: u32, pub device: i32,
}
pub enum tsb_generator {}
pub type tsb_progress_fn = Option<unsafe extern "C" fn(*mut c_void, *const u8, u32, u32, u64, u64, u64, u64)>;

extern "C" {
    pub fn tsb_pyramid_build(rgba: *const u8, w: u32, h: u32, levels: u32, out: *mut u8) -> c_int;
    pub fn tsb_resize(rgba: *const u8, w: u32, h: u32, out: *mut u8, nw: u32, nh: u32, filter: c_int) -> c_int;
    pub fn tsb_generator_create(desc: *const tsb_generator_desc, out: *mut *mut tsb_generator) -> c_int;
    pub fn tsb_generator_destroy(g: *mut tsb_generator);
    pub fn tsb_generator_random_init(g: *mut tsb_generator, count: u64, top: *const tsb_image, n: u32, seed: u64) -> c_int;
    pub fn tsb_generator_resolve(g: *mut tsb_generator, p: *const tsb_params, ex: *const tsb_pyramid, n: u32,
                                 guides: *const tsb_guides, sampling: *const tsb_sampling,
                                 cb: tsb_progress_fn, user: *mut c_void) -> c_int;
    pub fn tsb_generator_read_color(g: *mut tsb_generator, rgba: *mut u8) -> c_int;
    pub fn tsb_generator_read_coord(g: *mut tsb_generator, xym: *mut u32) -> c_int;
    pub fn tsb_generator_read_id(g: *mut tsb_generator, patch_map: *mut u32) -> c_int;
    pub fn tsb_generator_resolved_count(g: *mut tsb_generator, n: *mut u64, locked: *mut u64) -> c_int;
    pub fn tsb_generator_read_resolved(g: *mut tsb_generator, flat: *mut u32, score: *mut f32) -> c_int;
    pub fn tsb_generator_read_uncertainty(g: *mut tsb_generator, rgba: *mut u8) -> c_int;
    pub fn tsb_generator_read_id_maps(g: *mut tsb_generator, patch: *mut u8, map: *mut u8) -> c_int;
    pub fn tsb_last_error() -> *const c_char;
}

extern "C" {
    pub fn tsb_guide_map(rgba: *const u8, w: u32, h: u32, sigma: f32, out: *mut u8) -> c_int;
    pub fn tsb_match_histograms(source: *const u8, sw: u32, sh: u32, target: *const u8, tw: u32, th: u32, out: *mut u8) -> c_int;
    pub fn tsb_generator_upload_inputs(g: *mut tsb_generator, ex: *const tsb_pyramid, n: u32, guides: *const tsb_guides,
                                       sampling: *const tsb_sampling) -> c_int;
    pub fn tsb_generator_resolve_resident(g: *mut tsb_generator, p: *const tsb_params, cb: tsb_progress_fn, user: *mut c_void) -> c_int;
    pub fn tsb_generator_reset(g: *mut tsb_generator) -> c_int;
    pub fn tsb_device_count() -> c_int;
}

impl tsb_sampling {
    // SamplingMethod, lib.rs:458-468
    pub fn all() -> Self { tsb_sampling { kind: 0, _pad: 0, rgba: std::ptr::null() } }
    pub fn ignore() -> Self { tsb_sampling { kind: 1, _pad: 0, rgba: std::ptr::null() } }
    pub fn image(rgba: *const u8) -> Self { tsb_sampling { kind: 2, _pad: 0, rgba } }
}
