//! `Generator` on the B200: same method names and argument meaning as `ms::Generator` (reference lib/src/ms.rs:207-293,
//! 427-445, 605-684, 702-1052), every method a thin wrapper over one entry point of include/tsb200.h.
use crate::{
    img_pyramid::ImagePyramid, session::GeneratorProgress, CoordinateTransform, Dims, SamplingMethod,
};
use std::os::raw::c_void;
use texture_synthesis_b200_sys as ffi;

pub(crate) use crate::ms::{GeneratorParams, GuidesPyramidStruct}; // plain data, unchanged

pub(crate) fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(ffi::tsb_last_error()) }
        .to_string_lossy()
        .into_owned()
}

/// `run()` is infallible in the reference (session.rs:37) and panics on poisoned locks (ms.rs:381,1036): same contract.
fn check(rc: i32) {
    if rc != 0 {
        panic!("tsb200: {}", last_error());
    }
}

pub struct Generator {
    handle: *mut ffi::tsb_generator,
    pub(crate) output_size: Dims,
    input_dimensions: Vec<Dims>,
    /// filled by `resolve` so that `GeneratedImage::{save, as_ref, into_image}` keep borrowing an `RgbaImage`
    pub(crate) color_map: image::RgbaImage,
}

impl Drop for Generator {
    fn drop(&mut self) {
        unsafe { ffi::tsb_generator_destroy(self.handle) }
    }
}

fn contiguous(p: &ImagePyramid) -> (Vec<u8>, u32, u32, u32) {
    let (w, h) = p.bottom().dimensions();
    let mut all = Vec::with_capacity(p.pyramid.len() * (w * h * 4) as usize);
    for lvl in &p.pyramid {
        all.extend_from_slice(lvl.as_raw());
    }
    (all, w, h, p.pyramid.len() as u32)
}

impl Generator {
    pub(crate) fn new(size: Dims) -> Self {
        Self::create(size, None, 0)
    }

    /// ms.rs:236-293: mask / colour are brought to `size` with the Triangle filter first (ms.rs:242-263 -> tsb_resize)
    pub(crate) fn new_from_inpaint(
        size: Dims,
        inpaint_map: image::RgbaImage,
        color_map: image::RgbaImage,
        color_map_index: usize,
    ) -> Self {
        let fit = |img: image::RgbaImage| -> image::RgbaImage {
            if img.width() == size.width && img.height() == size.height {
                return img;
            }
            let mut out = vec![0u8; (size.width * size.height * 4) as usize];
            check(unsafe {
                ffi::tsb_resize(img.as_raw().as_ptr(), img.width(), img.height(), out.as_mut_ptr(), size.width, size.height, 0)
            });
            image::RgbaImage::from_raw(size.width, size.height, out).unwrap()
        };
        Self::create(size, Some((fit(inpaint_map), fit(color_map))), color_map_index)
    }

    fn create(size: Dims, inpaint: Option<(image::RgbaImage, image::RgbaImage)>, index: usize) -> Self {
        let (mask_ptr, color_ptr) = match &inpaint {
            Some((m, c)) => (m.as_raw().as_ptr(), c.as_raw().as_ptr()),
            None => (std::ptr::null(), std::ptr::null()),
        };
        let desc = ffi::tsb_generator_desc {
            out_width: size.width,
            out_height: size.height,
            inpaint_mask: mask_ptr,
            inpaint_color: color_ptr,
            inpaint_example_index: index as u32,
            device: -1,
        };
        let mut handle = std::ptr::null_mut();
        check(unsafe { ffi::tsb_generator_create(&desc, &mut handle) });
        Self {
            handle,
            output_size: size,
            input_dimensions: Vec::new(),
            color_map: inpaint.map(|(_, c)| c).unwrap_or_else(|| image::RgbaImage::new(size.width, size.height)),
        }
    }

    /// ms.rs:427-445; `example_maps` = pyramid[len - 1] of EVERY example (session.rs:43-48)
    pub(crate) fn resolve_random_batch(&mut self, steps: usize, example_maps: &[&image::RgbaImage], seed: u64) {
        let top: Vec<ffi::tsb_image> = example_maps
            .iter()
            .map(|m| ffi::tsb_image { rgba: m.as_raw().as_ptr(), width: m.width(), height: m.height() })
            .collect();
        check(unsafe { ffi::tsb_generator_random_init(self.handle, steps as u64, top.as_ptr(), top.len() as u32, seed) });
    }

    /// ms.rs:702-1052.  Blocking; the callback runs on this thread only (`Box<dyn GeneratorProgress>` is not `Send`).
    pub(crate) fn resolve(
        &mut self,
        params: &GeneratorParams,
        example_maps_pyramid: &[ImagePyramid],
        mut progress: Option<Box<dyn GeneratorProgress>>,
        guides_pyramid: &Option<GuidesPyramidStruct>,
        valid_samples: &[SamplingMethod],
    ) {
        self.input_dimensions = example_maps_pyramid
            .iter()
            .map(|ip| Dims { width: ip.bottom().width(), height: ip.bottom().height() })
            .collect();
        let owned: Vec<_> = example_maps_pyramid.iter().map(contiguous).collect();
        let examples: Vec<ffi::tsb_pyramid> = owned
            .iter()
            .map(|(b, w, h, n)| ffi::tsb_pyramid { levels: b.as_ptr(), width: *w, height: *h, n_levels: *n })
            .collect();
        let sampling: Vec<ffi::tsb_sampling> = valid_samples
            .iter()
            .map(|s| match s {
                SamplingMethod::All => ffi::tsb_sampling::all(),
                SamplingMethod::Ignore => ffi::tsb_sampling::ignore(),
                SamplingMethod::Image(img) => ffi::tsb_sampling::image(img.as_raw().as_ptr()),
            })
            .collect();
        let g_owned = guides_pyramid.as_ref().map(|g| {
            (contiguous(&g.target_guide), g.example_guides.iter().map(contiguous).collect::<Vec<_>>())
        });
        let g_examples: Vec<ffi::tsb_pyramid> = g_owned
            .iter()
            .flat_map(|(_, ex)| ex.iter())
            .map(|(b, w, h, n)| ffi::tsb_pyramid { levels: b.as_ptr(), width: *w, height: *h, n_levels: *n })
            .collect();
        let guides = g_owned.as_ref().map(|((b, w, h, n), _)| ffi::tsb_guides {
            target: ffi::tsb_pyramid { levels: b.as_ptr(), width: *w, height: *h, n_levels: *n },
            examples: g_examples.as_ptr(),
            n_examples: g_examples.len() as u32,
            _pad: 0,
        });
        let ffi_params = ffi::tsb_params {
            nearest_neighbors: params.nearest_neighbors,
            _pad0: 0,
            random_sample_locations: params.random_sample_locations,
            cauchy_dispersion: params.cauchy_dispersion,
            p: params.p,
            p_stages: params.p_stages,
            alpha: params.alpha,
            seed: params.seed,
            max_thread_count: params.max_thread_count as u64,
            tiling_mode: params.tiling_mode as i32,
            _pad1: 0,
        };

        unsafe extern "C" fn trampoline(user: *mut c_void, rgba: *const u8, w: u32, h: u32, tc: u64, tt: u64, sc: u64, st: u64) {
            let progress = &mut *(user as *mut Box<dyn GeneratorProgress>);
            let bytes = std::slice::from_raw_parts(rgba, (w * h * 4) as usize);
            let image = image::RgbaImage::from_raw(w, h, bytes.to_vec()).unwrap();
            progress.update(crate::session::ProgressUpdate {
                image: &image,
                total: crate::session::ProgressStat { total: tt as usize, current: tc as usize },
                stage: crate::session::ProgressStat { total: st as usize, current: sc as usize },
            });
        }
        let (cb, user): (ffi::tsb_progress_fn, *mut c_void) = match progress.as_mut() {
            Some(p) => (Some(trampoline), p as *mut Box<dyn GeneratorProgress> as *mut c_void),
            None => (None, std::ptr::null_mut()),
        };
        check(unsafe {
            ffi::tsb_generator_resolve(
                self.handle,
                &ffi_params,
                examples.as_ptr(),
                examples.len() as u32,
                guides.as_ref().map_or(std::ptr::null(), |g| g as *const _),
                sampling.as_ptr(),
                cb,
                user,
            )
        });
        let mut color = vec![0u8; (self.output_size.width * self.output_size.height * 4) as usize];
        check(unsafe { ffi::tsb_generator_read_color(self.handle, color.as_mut_ptr()) });
        self.color_map = image::RgbaImage::from_raw(self.output_size.width, self.output_size.height, color).unwrap();
    }

    /// ms.rs:605-633: [patch_id_map, map_id_map]
    pub fn get_id_maps(&self) -> [image::RgbaImage; 2] {
        let (w, h) = (self.output_size.width, self.output_size.height);
        let mut a = vec![0u8; (w * h * 4) as usize];
        let mut b = vec![0u8; (w * h * 4) as usize];
        check(unsafe { ffi::tsb_generator_read_id_maps(self.handle, a.as_mut_ptr(), b.as_mut_ptr()) });
        [image::RgbaImage::from_raw(w, h, a).unwrap(), image::RgbaImage::from_raw(w, h, b).unwrap()]
    }

    /// ms.rs:635-653
    pub fn get_uncertainty_map(&self) -> image::RgbaImage {
        let (w, h) = (self.output_size.width, self.output_size.height);
        let mut a = vec![0u8; (w * h * 4) as usize];
        check(unsafe { ffi::tsb_generator_read_uncertainty(self.handle, a.as_mut_ptr()) });
        image::RgbaImage::from_raw(w, h, a).unwrap()
    }

    /// ms.rs:655-684: the device already returns `[x, y, map]` u32 triplets
    pub fn get_coord_transform(&self) -> CoordinateTransform {
        let n = (self.output_size.width * self.output_size.height) as usize;
        let mut buffer = vec![0u32; n * 3];
        check(unsafe { ffi::tsb_generator_read_coord(self.handle, buffer.as_mut_ptr()) });
        CoordinateTransform {
            buffer,
            output_size: Dims::new(self.output_size.width, self.output_size.height),
            original_maps: self.input_dimensions.clone(),
        }
    }
}
