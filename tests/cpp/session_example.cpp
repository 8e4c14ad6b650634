// C++ host-side usage of the Session mirror (include/tsb200_session.hpp) -- the same flows as the reference's
// lib/examples/01_single_example_synthesis.rs, 07_tiling_texture.rs and 04_style_transfer.rs, on raw RGBA files.
//   session_example <mode> <in.rgba> <w> <h> <out.rgba> <out_w> <out_h> [<target.rgba> <tw> <th>]
//   modes: single | tiling | style | validate
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "tsb200_session.hpp"

static tsb::Image read_rgba(const char* path, uint32_t w, uint32_t h) {
    tsb::Image im(w, h);
    FILE* f = fopen(path, "rb");
    if (!f || fread(im.rgba.data(), 1, im.rgba.size(), f) != im.rgba.size()) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return im;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "validate") {  // session.rs:450-524 without touching the GPU
        int ok = 0;
        tsb::Image img(16, 16);
        try { tsb::Session::builder().add_example(img).cauchy_dispersion(1.5f).build(); } catch (const tsb::Error& e) { ok += e.kind == tsb::Error::InvalidRange && e.name == "cauchy-dispersion"; }
        try { tsb::Session::builder().add_example(img).random_sample_locations(0).build(); } catch (const tsb::Error& e) { ok += e.kind == tsb::Error::InvalidRange && e.name == "m-rand"; }
        try { tsb::Session::builder().build(); } catch (const tsb::Error& e) { ok += e.kind == tsb::Error::NoExamples; }
        try { tsb::Session::builder().add_example(tsb::Example(img).with_guide(img)).add_example(img).build(); } catch (const tsb::Error& e) { ok += e.kind == tsb::Error::ExampleGuideMismatch; }
        printf("validate %d/4\n", ok);
        return ok == 4 ? 0 : 1;
    }
    if (argc < 8) { fprintf(stderr, "usage: see source\n"); return 2; }
    tsb::Image ex = read_rgba(argv[2], (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]));
    const tsb::Dims out{(uint32_t)atoi(argv[6]), (uint32_t)atoi(argv[7])};
    try {
        auto b = tsb::Session::builder();
        b.add_example(ex).seed(120).output_size(out).max_thread_count(1);
        if (mode == "tiling") b.tiling_mode(true);
        if (mode == "style") b.load_target_guide(read_rgba(argv[8], (uint32_t)atoi(argv[9]), (uint32_t)atoi(argv[10])));
        uint64_t last = 0;
        tsb::GeneratedImage gen = b.build().run([&](const uint8_t*, uint32_t, uint32_t, uint64_t cur, uint64_t total, uint64_t, uint64_t) { last = cur * 100 / total; });
        const tsb::Image& img = gen.as_image();
        tsb::CoordinateTransform ct = gen.get_coordinate_transform();
        tsb::Image again = ct.apply({ex});  // repeat_transform, lib/tests/diff.rs:254-284
        if (again.rgba != img.rgba) { fprintf(stderr, "coordinate transform does not reproduce the image\n"); return 1; }
        FILE* f = fopen(argv[5], "wb");
        fwrite(img.rgba.data(), 1, img.rgba.size(), f);
        fclose(f);
        printf("ok %s %ux%u progress %llu%%\n", mode.c_str(), img.width, img.height, (unsigned long long)last);
    } catch (const tsb::Error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
