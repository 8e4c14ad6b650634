//! Runs against the UNPATCHED reference crate: dumps everything needed to diff the real `max_thread_count(1)` output against the
//! CUDA path on identical decoded inputs.  `cargo test --release --test dump_snapshots -- --nocapture`, then
//! `python tests/compare_rust_snapshots.py lib/target/tsb_snapshots` in the texture-synthesis-b200 repository.
//!
//! Per configuration of lib/tests/diff.rs:163-252 it writes into `target/tsb_snapshots/<name>/`:
//!   input_<i>.rgba (+ .dims), guide_<i>.rgba, target_guide.rgba, mask.rgba   -- the bytes the `image` crate decoded
//!   output.rgba                                                                -- GeneratedImage::into_image()
//!   transform.bin                                                              -- CoordinateTransform::write (lib.rs:212-249)
use std::{fs, io::Write, path::Path};
use texture_synthesis as ts;

fn dump_rgba(dir: &Path, name: &str, path: &str) {
    let img = image::open(path).expect("decode").to_rgba8();
    fs::write(dir.join(format!("{}.rgba", name)), img.as_raw()).unwrap();
    fs::write(dir.join(format!("{}.dims", name)), format!("{} {}", img.width(), img.height())).unwrap();
}

fn run(name: &str, inputs: &[(&str, &str)], builder: ts::SessionBuilder<'_>) {
    let dir = Path::new("target/tsb_snapshots").join(name);
    fs::create_dir_all(&dir).unwrap();
    for (tag, path) in inputs {
        dump_rgba(&dir, tag, path);
    }
    let generated = builder.max_thread_count(1).build().unwrap().run(None);
    let mut f = fs::File::create(dir.join("transform.bin")).unwrap();
    generated.get_coordinate_transform().write(&mut f).unwrap();
    f.flush().unwrap();
    let out = generated.into_image().to_rgba8();
    fs::write(dir.join("output.rgba"), out.as_raw()).unwrap();
    fs::write(dir.join("output.dims"), format!("{} {}", out.width(), out.height())).unwrap();
}

#[test]
fn dump_snapshots() {
    use ts::Dims;
    run("single_example", &[("input_0", "../imgs/1.jpg")],
        ts::Session::builder().add_example(&"../imgs/1.jpg").seed(120).output_size(Dims::square(100)));
    run("multi_example",
        &[("input_0", "../imgs/multiexample/1.jpg"), ("input_1", "../imgs/multiexample/2.jpg"),
          ("input_2", "../imgs/multiexample/3.jpg"), ("input_3", "../imgs/multiexample/4.jpg")],
        ts::Session::builder()
            .add_examples(&[&"../imgs/multiexample/1.jpg", &"../imgs/multiexample/2.jpg",
                            &"../imgs/multiexample/3.jpg", &"../imgs/multiexample/4.jpg"])
            .resize_input(Dims::square(100)).random_init(10).seed(211).output_size(Dims::square(100)));
    run("guided",
        &[("input_0", "../imgs/2.jpg"), ("guide_0", "../imgs/masks/2_example.jpg"), ("target_guide", "../imgs/masks/2_target.jpg")],
        ts::Session::builder()
            .add_example(ts::Example::builder(&"../imgs/2.jpg").with_guide(&"../imgs/masks/2_example.jpg"))
            .load_target_guide(&"../imgs/masks/2_target.jpg").output_size(Dims::square(100)));
    run("style_transfer", &[("input_0", "../imgs/multiexample/4.jpg"), ("target_guide", "../imgs/tom.jpg")],
        ts::Session::builder().add_example(&"../imgs/multiexample/4.jpg").load_target_guide(&"../imgs/tom.jpg")
            .output_size(Dims::square(100)));
    run("inpaint", &[("input_0", "../imgs/3.jpg"), ("mask", "../imgs/masks/3_inpaint.jpg")],
        ts::Session::builder().inpaint_example(
            &"../imgs/masks/3_inpaint.jpg",
            ts::Example::builder(&"../imgs/3.jpg").set_sample_method(&"../imgs/masks/3_inpaint.jpg"),
            Dims::square(100)));
    run("inpaint_channel", &[("input_0", "../imgs/bricks.png")],
        ts::Session::builder().inpaint_example_channel(ts::ChannelMask::A, &"../imgs/bricks.png", Dims::square(400)));
    run("tiling", &[("input_0", "../imgs/1.jpg"), ("mask", "../imgs/masks/1_tile.jpg")],
        ts::Session::builder()
            .inpaint_example(&"../imgs/masks/1_tile.jpg", ts::Example::new(&"../imgs/1.jpg"), Dims::square(100))
            .tiling_mode(true));
    run("sample_masks", &[("input_0", "../imgs/4.png"), ("mask", "../imgs/masks/4_sample_mask.png")],
        ts::Session::builder()
            .add_example(ts::Example::builder(&"../imgs/4.png").set_sample_method(&"../imgs/masks/4_sample_mask.png"))
            .seed(211).output_size(Dims::square(100)));
    run("sample_masks_ignore", &[("input_0", "../imgs/4.png"), ("input_1", "../imgs/5.png")],
        ts::Session::builder()
            .add_example(ts::Example::builder(&"../imgs/4.png").set_sample_method(ts::SampleMethod::Ignore))
            .add_example(ts::Example::builder(&"../imgs/5.png").set_sample_method(ts::SampleMethod::All))
            .seed(211).output_size(Dims::square(200)));
}
