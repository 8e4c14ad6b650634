//! Patched `lib/src/img_pyramid.rs`: the Gaussian pyramid is built by the CUDA library (`tsb_pyramid_build`,
//! include/tsb200.h) instead of `image::imageops::resize` (reference lines 20-37).  Public surface unchanged.
use texture_synthesis_b200_sys as ffi;

#[derive(Clone)]
pub struct ImagePyramid {
    pub pyramid: Vec<image::RgbaImage>,
}

impl ImagePyramid {
    pub fn new(in_img: image::RgbaImage, levels: Option<u32>) -> Self {
        let lvls = levels.unwrap_or_else(|| {
            let (dimx, dimy) = in_img.dimensions();
            (f64::from(dimx.max(dimy))).log2() as u32
        });
        Self {
            pyramid: Self::build_gaussian(lvls, in_img),
        }
    }

    /// level 0 = blurriest ... level `in_lvls - 1` = the input (reference: `for i in (1..in_lvls).rev()` + push of the input)
    fn build_gaussian(in_lvls: u32, in_img: image::RgbaImage) -> Vec<image::RgbaImage> {
        let (w, h) = in_img.dimensions();
        let n = in_lvls.max(1) as usize;
        let level_bytes = w as usize * h as usize * 4;
        let mut out = vec![0u8; n * level_bytes];
        let rc = unsafe { ffi::tsb_pyramid_build(in_img.as_raw().as_ptr(), w, h, in_lvls, out.as_mut_ptr()) };
        if rc != 0 {
            // the reference panics inside `image` on degenerate sizes; keep that contract
            panic!("tsb_pyramid_build failed: {}", crate::b200::last_error());
        }
        out.chunks_exact(level_bytes)
            .map(|lvl| image::RgbaImage::from_raw(w, h, lvl.to_vec()).expect("level size"))
            .collect()
    }

    pub fn bottom(&self) -> &image::RgbaImage {
        &self.pyramid[self.pyramid.len() - 1]
    }
}
