// tsb200_session.hpp -- header-only C++17 mirror of the reference's public API for the hot path, on top of the C ABI
// (include/tsb200.h).  The reference is a Rust crate (`texture_synthesis::Session`, lib/src/session.rs:22-524;
// `Example`, `SampleMethod`, `GeneratedImage`, `CoordinateTransform`, `Dims`, `Error`, lib/src/lib.rs); no Rust toolchain
// exists in the build image, so the host side is mirrored here with the same names, defaults (lib.rs:343-359) and
// validation (session.rs:450-524).  Image decoding/encoding is out of scope: images are RGBA8 buffers.
//
//   auto session = tsb::Session::builder().add_example(img).seed(10).tiling_mode(true).build();
//   tsb::GeneratedImage generated = session.run();
//   const tsb::Image& out = generated.as_image();
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tsb200.h"

namespace tsb {

struct Dims {  // lib.rs:142-160
    uint32_t width = 0, height = 0;
    static Dims square(uint32_t s) { return Dims{s, s}; }
    bool operator==(const Dims& o) const { return width == o.width && height == o.height; }
};

struct Image {  // image::RgbaImage
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;
    Image() = default;
    Image(uint32_t w, uint32_t h) : width(w), height(h), rgba((size_t)w * h * 4) {}
    Dims dims() const { return Dims{width, height}; }
};

// texture_synthesis::Error (errors.rs:38-57)
struct Error : std::runtime_error {
    enum Kind { Image, InvalidRange, SizeMismatch, ExampleGuideMismatch, Io, UnsupportedOutputFormat, NoExamples, MapsCountMismatch };
    Kind kind;
    std::string name;  // InvalidRange: parameter name
    Error(Kind k, const std::string& msg, std::string n = "") : std::runtime_error(msg), kind(k), name(std::move(n)) {}
};

inline void check(int rc) {
    if (rc != 0) throw Error(Error::Io, std::string("tsb200: ") + tsb_last_error());  // Error::Io(Other, msg)
}

inline Image resize(const Image& img, Dims to, int filter) {
    Image out(to.width, to.height);
    check(tsb_resize(img.rgba.data(), img.width, img.height, out.rgba.data(), to.width, to.height, filter));
    return out;
}

// utils::load_image (utils.rs:55-80) for an already decoded image: CatmullRom resize if the size differs
inline Image load_image(const Image& img, const std::optional<Dims>& size) {
    if (size && !(img.dims() == *size)) return resize(img, *size, TSB_FILTER_CATMULLROM);
    return img;
}

struct ImagePyramid {  // img_pyramid.rs:2-42
    uint32_t width = 0, height = 0, levels = 0;
    std::vector<uint8_t> data;  // levels images back to back, level 0 = blurriest
    ImagePyramid() = default;
    ImagePyramid(const Image& img, uint32_t lv) : width(img.width), height(img.height), levels(lv == 0 ? 1 : lv) {
        data.resize((size_t)levels * width * height * 4);
        check(tsb_pyramid_build(img.rgba.data(), width, height, lv, data.data()));
    }
    Image bottom() const {
        Image im(width, height);
        std::copy(data.end() - (ptrdiff_t)im.rgba.size(), data.end(), im.rgba.begin());
        return im;
    }
    tsb_pyramid ffi() const { return tsb_pyramid{data.data(), width, height, levels}; }
};

struct SampleMethod {  // lib.rs:458-484
    enum Kind { All = TSB_SAMPLE_ALL, Ignore = TSB_SAMPLE_IGNORE, ImageMask = TSB_SAMPLE_IMAGE } kind = All;
    Image mask;
    static SampleMethod all() { return {}; }
    static SampleMethod ignore() { SampleMethod m; m.kind = Ignore; return m; }
    static SampleMethod image(Image im) { SampleMethod m; m.kind = ImageMask; m.mask = std::move(im); return m; }
};

struct Example {  // lib.rs:487-623
    Image img;
    std::optional<Image> guide;
    SampleMethod sample_method;
    Example(Image i) : img(std::move(i)) {}  // NOLINT: mirrors `impl From<IS> for Example`
    Example& with_guide(Image g) { guide = std::move(g); return *this; }
    Example& set_sample_method(SampleMethod m) { sample_method = std::move(m); return *this; }
};

struct CoordinateTransform {  // lib.rs:162-325
    std::vector<uint32_t> buffer;  // [x, y, map] per output pixel
    Dims output_size;
    std::vector<Dims> original_maps;
    Image apply(const std::vector<Image>& sources) const {  // lib.rs:178-210
        if (sources.size() != original_maps.size()) throw Error(Error::MapsCountMismatch, "maps count mismatch");
        std::vector<Image> imgs;
        for (size_t i = 0; i < sources.size(); ++i) imgs.push_back(load_image(sources[i], original_maps[i]));
        Image out(output_size.width, output_size.height);
        for (size_t p = 0; p < (size_t)output_size.width * output_size.height; ++p) {
            const Image& src = imgs[buffer[p * 3 + 2]];
            const uint8_t* s = &src.rgba[((size_t)buffer[p * 3 + 1] * src.width + buffer[p * 3]) * 4];
            std::copy(s, s + 4, &out.rgba[p * 4]);
        }
        return out;
    }
};

class GeneratedImage {  // lib.rs:378-455
  public:
    GeneratedImage(tsb_generator* g, Dims out, std::vector<Dims> inputs) : g_(g, tsb_generator_destroy), out_(out), inputs_(std::move(inputs)) {
        img_ = Image(out.width, out.height);
        check(tsb_generator_read_color(g_.get(), img_.rgba.data()));
    }
    const Image& as_image() const { return img_; }
    Image into_image() { return std::move(img_); }
    CoordinateTransform get_coordinate_transform() const {
        CoordinateTransform t;
        t.output_size = out_;
        t.original_maps = inputs_;
        t.buffer.resize((size_t)out_.width * out_.height * 3);
        check(tsb_generator_read_coord(g_.get(), t.buffer.data()));
        return t;
    }
    struct Debug { Image uncertainty, patch_id, map_id; };
    Debug debug_maps() const {  // what save_debug writes (lib.rs:407-419)
        Debug d{Image(out_.width, out_.height), Image(out_.width, out_.height), Image(out_.width, out_.height)};
        check(tsb_generator_read_uncertainty(g_.get(), d.uncertainty.rgba.data()));
        check(tsb_generator_read_id_maps(g_.get(), d.patch_id.rgba.data(), d.map_id.rgba.data()));
        return d;
    }

  private:
    std::shared_ptr<tsb_generator> g_;
    Dims out_;
    std::vector<Dims> inputs_;
    Image img_;
};

using ProgressFn = std::function<void(const uint8_t* rgba, uint32_t w, uint32_t h, uint64_t total_cur, uint64_t total, uint64_t stage_cur, uint64_t stage_total)>;

class Session {  // session.rs:22-66
  public:
    class Builder;
    static Builder builder();

    GeneratedImage run(ProgressFn progress = nullptr) {
        tsb_params p = params_;
        if (random_resolve_) {  // session.rs:42-52
            std::vector<Image> tops;
            std::vector<tsb_image> ffi;
            for (auto& e : examples_) tops.push_back(e.bottom());
            for (auto& t : tops) ffi.push_back(tsb_image{t.rgba.data(), t.width, t.height});
            check(tsb_generator_random_init(gen_, *random_resolve_, ffi.data(), (uint32_t)ffi.size(), p.seed));
        }
        std::vector<tsb_pyramid> ex;
        for (auto& e : examples_) ex.push_back(e.ffi());
        std::vector<tsb_sampling> sm;
        for (auto& m : methods_) sm.push_back(tsb_sampling{(int32_t)m.kind, 0, m.kind == SampleMethod::ImageMask ? m.mask.rgba.data() : nullptr});
        std::vector<tsb_pyramid> gex;
        tsb_guides guides{};
        const tsb_guides* gp = nullptr;
        if (target_guide_) {
            for (auto& e : guides_) gex.push_back(e.ffi());
            guides = tsb_guides{target_guide_->ffi(), gex.data(), (uint32_t)gex.size(), 0};
            gp = &guides;
        }
        auto tramp = [](void* user, const uint8_t* rgba, uint32_t w, uint32_t h, uint64_t tc, uint64_t tt, uint64_t sc, uint64_t st) {
            (*static_cast<ProgressFn*>(user))(rgba, w, h, tc, tt, sc, st);
        };
        check(tsb_generator_resolve(gen_, &p, ex.data(), (uint32_t)ex.size(), gp, sm.data(), progress ? +tramp : nullptr, progress ? &progress : nullptr));
        std::vector<Dims> dims;
        for (auto& e : examples_) dims.push_back(Dims{e.width, e.height});
        tsb_generator* g = gen_;
        gen_ = nullptr;
        return GeneratedImage(g, out_size_, dims);
    }
    ~Session() { if (gen_) tsb_generator_destroy(gen_); }
    Session(Session&& o) noexcept { *this = std::move(o); }
    Session& operator=(Session&& o) noexcept {
        std::swap(gen_, o.gen_); params_ = o.params_; out_size_ = o.out_size_; random_resolve_ = o.random_resolve_;
        examples_ = std::move(o.examples_); guides_ = std::move(o.guides_); target_guide_ = std::move(o.target_guide_); methods_ = std::move(o.methods_);
        return *this;
    }

  private:
    friend class Builder;
    Session() = default;
    tsb_generator* gen_ = nullptr;
    tsb_params params_{};
    Dims out_size_;
    std::optional<uint64_t> random_resolve_;
    std::vector<ImagePyramid> examples_, guides_;
    std::optional<ImagePyramid> target_guide_;
    std::vector<SampleMethod> methods_;
};

class Session::Builder {  // SessionBuilder, session.rs:72-524
  public:
    Builder& add_example(Example e) { examples_.push_back(std::move(e)); return *this; }
    Builder& add_examples(std::vector<Example> es) { for (auto& e : es) examples_.push_back(std::move(e)); return *this; }
    Builder& inpaint_example(Image mask, Example example, Dims size) {
        inpaint_ = Inpaint{std::move(mask), examples_.size(), size};
        return add_example(std::move(example));
    }
    Builder& load_target_guide(Image g) { target_guide_ = std::move(g); return *this; }
    Builder& resize_input(Dims d) { resize_input_ = d; return *this; }
    Builder& seed(uint64_t v) { seed_ = v; return *this; }
    Builder& tiling_mode(bool v) { tiling_ = v; return *this; }
    Builder& nearest_neighbors(uint32_t v) { k_ = v; return *this; }
    Builder& random_sample_locations(uint64_t v) { m_ = v; return *this; }
    Builder& random_init(uint64_t v) { random_resolve_ = v; return *this; }
    Builder& cauchy_dispersion(float v) { cauchy_ = v; return *this; }
    Builder& guide_alpha(float v) { alpha_ = v; return *this; }
    Builder& backtrack_percent(float v) { p_ = v; return *this; }
    Builder& backtrack_stages(uint32_t v) { stages_ = v; return *this; }
    Builder& output_size(Dims d) { out_size_ = d; return *this; }
    Builder& max_thread_count(size_t v) { threads_ = v; return *this; }

    Session build() {
        // check_parameters_validity, session.rs:450-499
        auto range = [](const char* name, float lo, float hi, float v) {
            if (v < lo || v > hi) throw Error(Error::InvalidRange, std::string("parameter '") + name + "' is out of range", name);
        };
        range("cauchy-dispersion", 0.f, 1.f, cauchy_);
        range("backtrack-percent", 0.f, 1.f, p_);
        range("guide-alpha", 0.f, 1.f, alpha_);
        if (threads_ && *threads_ == 0) throw Error(Error::InvalidRange, "max-thread-count must be >= 1", "max-thread-count");
        if (m_ == 0) throw Error(Error::InvalidRange, "m-rand must be >= 1", "m-rand");
        // check_images_validity, session.rs:501-524
        size_t usable = 0, n_guides = 0;
        for (auto& e : examples_) { usable += e.sample_method.kind != SampleMethod::Ignore; n_guides += e.guide.has_value(); }
        if (usable == 0) throw Error(Error::NoExamples, "at least 1 example that is not ignored is required");
        if (n_guides != 0 && n_guides != examples_.size()) throw Error(Error::ExampleGuideMismatch, "every example needs a guide");

        Session s;
        std::optional<Dims> in_size = resize_input_;
        s.out_size_ = out_size_;
        Image mask_img, color_img;
        if (inpaint_) {  // session.rs:346-382: output = input = inpaint dims
            s.out_size_ = inpaint_->size;
            in_size = inpaint_->size;
            mask_img = load_image(inpaint_->mask, inpaint_->size);
            color_img = load_image(examples_[inpaint_->example_index].img, inpaint_->size);
        }
        if (target_guide_) {  // session.rs:384-401
            Image tg = load_image(*target_guide_, s.out_size_);
            if (n_guides == 0) tg = guide_map(tg, 2.0f);
            s.target_guide_ = ImagePyramid(tg, stages_);
        }
        for (auto& e : examples_) {  // Example::resolve, lib.rs:564-603
            ImagePyramid pyr(load_image(e.img, in_size), stages_);
            if (s.target_guide_) {
                if (e.guide) s.guides_.emplace_back(load_image(*e.guide, in_size), stages_);
                else {
                    Image gm = guide_map(pyr.bottom(), 2.0f);
                    Image tb = s.target_guide_->bottom();
                    Image matched(gm.width, gm.height);
                    check(tsb_match_histograms(gm.rgba.data(), gm.width, gm.height, tb.rgba.data(), tb.width, tb.height, matched.rgba.data()));
                    s.guides_.emplace_back(matched, stages_);
                }
            }
            SampleMethod m = e.sample_method;
            if (m.kind == SampleMethod::ImageMask) m.mask = load_image(m.mask, in_size);
            s.methods_.push_back(std::move(m));
            s.examples_.push_back(std::move(pyr));
        }
        tsb_generator_desc d{s.out_size_.width, s.out_size_.height, inpaint_ ? mask_img.rgba.data() : nullptr,
                             inpaint_ ? color_img.rgba.data() : nullptr, inpaint_ ? (uint32_t)inpaint_->example_index : 0u, -1};
        check(tsb_generator_create(&d, &s.gen_));
        s.params_ = tsb_params{k_, 0, m_, cauchy_, p_, (int32_t)stages_, alpha_, seed_, threads_ ? (uint64_t)*threads_ : 1ull, tiling_ ? 1 : 0, 0};
        s.random_resolve_ = random_resolve_;
        return s;
    }

  private:
    static Image guide_map(const Image& img, float sigma) {  // utils::transform_to_guide_map
        Image out(img.width, img.height);
        check(tsb_guide_map(img.rgba.data(), img.width, img.height, sigma, out.rgba.data()));
        return out;
    }
    struct Inpaint { Image mask; size_t example_index; Dims size; };
    std::vector<Example> examples_;
    std::optional<Image> target_guide_;
    std::optional<Inpaint> inpaint_;
    // Parameters::default(), lib.rs:343-359
    bool tiling_ = false;
    uint32_t k_ = 50;
    uint64_t m_ = 50;
    float cauchy_ = 1.0f, p_ = 0.5f, alpha_ = 0.8f;
    uint32_t stages_ = 5;
    std::optional<Dims> resize_input_;
    Dims out_size_ = Dims::square(500);
    std::optional<uint64_t> random_resolve_;
    std::optional<size_t> threads_;
    uint64_t seed_ = 0;
};

inline Session::Builder Session::builder() { return Builder(); }

}  // namespace tsb
