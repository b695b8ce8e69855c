// raytracer.hpp — host-side C++ mirror of the reference's scene/camera builder API.
//
// The reference host is Rust (src/main.rs, src/scenes.rs over the `raytracer` crate); no
// Rust toolchain exists in this image, so the host side above the C ABI is C++ with the
// same names, argument meaning and error behaviour as the crate's public interface:
//   raytracer::scene::SceneBuilder{default,set_skybox,add,finalize}   scene/mod.rs:79-138
//   raytracer::geometry::{Sphere, xy_rect, yz_rect, xz_rect, RectBox} geometry/*.rs
//   raytracer::material::{Metal, Dielectric, Lambertian, DiffuseLight, FairyLight}
//   raytracer::material::texture::TextureLoader                        material/texture/loader.rs:17-61
//   raytracer::skybox::SkyBox                                          skybox/mod.rs:11-16
//   raytracer::camera::{CameraBuilder, Camera, CameraPosition, Dimmensions, AspectRatio}
// All values are f64, exactly as in the reference; `SceneBuilder::finalize` flattens to the
// f32 SoA arrays of include/b200rt.h (the only place precision is reduced).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../include/b200rt.h"

namespace raytracer {

// anyhow::Error stand-in: every fallible builder call throws this; the C shim converts it
// to a status code + message.
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

namespace core {
struct Vec3 {   // core/vec3.rs:29-32
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    Vec3 operator+(const Vec3& r) const { return {x + r.x, y + r.y, z + r.z}; }
    Vec3 operator-(const Vec3& r) const { return {x - r.x, y - r.y, z - r.z}; }
    Vec3 operator*(const Vec3& r) const { return {x * r.x, y * r.y, z * r.z}; }
    Vec3 scale(double s) const { return {x * s, y * s, z * s}; }
    double dot(const Vec3& r) const { return (x * r.x + y * r.y) + z * r.z; }
    double length() const { return std::sqrt(dot(*this)); }
    Vec3 cross(const Vec3& r) const { return {y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x}; }
    Vec3 unit() const { double n = length(); return {x / n, y / n, z / n}; }
    double unit_mut() { double n = length(); x /= n; y /= n; z /= n; return n; }   // vec3.rs:175-178
};
struct Point { Vec3 v; Point() = default; explicit Point(Vec3 a) : v(a) {} };     // vec3.rs:232
struct Color { Vec3 v; Color() = default; explicit Color(Vec3 a) : v(a) {} };     // color.rs:11
}  // namespace core

namespace geometry {
struct Sphere { core::Point center; double radius = 0; };                          // sphere.rs:12-15
struct Rect { uint32_t kind = B200RT_PRIM_RECT_XY; double d1_min = 0, d1_max = 0, d2_min = 0, d2_max = 0, offset = 0; };   // rect.rs:46-52
inline Rect xy_rect(double a, double b, double c, double d, double k) { return {B200RT_PRIM_RECT_XY, a, b, c, d, k}; }    // rect.rs:15
inline Rect yz_rect(double a, double b, double c, double d, double k) { return {B200RT_PRIM_RECT_YZ, a, b, c, d, k}; }    // rect.rs:25
inline Rect xz_rect(double a, double b, double c, double d, double k) { return {B200RT_PRIM_RECT_XZ, a, b, c, d, k}; }    // rect.rs:35
struct RectBox {                                                                    // rect.rs:102-129
    core::Point min, max;
    RectBox() = default;
    RectBox(core::Point p0, core::Point p1) : min(p0), max(p1) {}
    static RectBox create(core::Point p0, core::Point p1) { return RectBox(p0, p1); }   // RectBox::new
};
using GeometricObject = std::variant<Sphere, Rect, RectBox>;                        // object.rs:9-16
}  // namespace geometry

namespace material {
namespace texture {
// TextureLoader, loader.rs:17-28.  Equality/ordering are on f64 BIT patterns, like the
// reference's ColorSetting/ScalarSetting Hash+Eq (texture/mod.rs:42-84).
struct TextureLoader {
    enum Kind { Solid, ImagePath, Perlin, EarthBuiltin, Checker } kind = Solid;
    core::Color color;          // Solid
    std::string path;           // ImagePath
    double scalar = 0;          // Perlin scale / Checker size
    std::shared_ptr<TextureLoader> odd, even;   // Checker
    static TextureLoader solid(double r, double g, double b) { TextureLoader t; t.kind = Solid; t.color = core::Color({r, g, b}); return t; }
    static TextureLoader solid_from_vec(core::Vec3 v) { TextureLoader t; t.kind = Solid; t.color = core::Color(v); return t; }
    static TextureLoader checker(double size, TextureLoader o, TextureLoader e) {
        TextureLoader t; t.kind = Checker; t.scalar = size;
        t.odd = std::make_shared<TextureLoader>(std::move(o)); t.even = std::make_shared<TextureLoader>(std::move(e)); return t;
    }
    static TextureLoader noise(double scalar) { TextureLoader t; t.kind = Perlin; t.scalar = scalar; return t; }
    static TextureLoader earth_builtin() { TextureLoader t; t.kind = EarthBuiltin; return t; }
    static TextureLoader image_path(std::string p) { TextureLoader t; t.kind = ImagePath; t.path = std::move(p); return t; }
    std::string key() const;    // canonical dedup key (bit patterns)
};
}  // namespace texture

struct Metal {                                                                      // metal.rs:10-24
    core::Color albedo; double fuzz = 0;
    Metal() = default;
    Metal(core::Color a, std::optional<double> f) : albedo(a) { double z = f.value_or(0.0); if (z > 1.0) z = 1.0; fuzz = z; }
};
struct Dielectric { double ir = 1.0; };                                             // dielectric.rs:10-13
struct Lambertian { texture::TextureLoader albedo; explicit Lambertian(texture::TextureLoader t = {}) : albedo(std::move(t)) {} };
struct DiffuseLight { texture::TextureLoader albedo; explicit DiffuseLight(texture::TextureLoader t = {}) : albedo(std::move(t)) {} };
struct FairyLight { texture::TextureLoader albedo; explicit FairyLight(texture::TextureLoader t = {}) : albedo(std::move(t)) {} };
// MaterialType<TextureLoader>, material_type.rs:20-27 (same variant order)
using MaterialType = std::variant<Metal, Dielectric, Lambertian, DiffuseLight, FairyLight>;
}  // namespace material

namespace skybox {
struct SkyBox {                                                                     // skybox/mod.rs:11-16
    enum Kind { Above, Flat, None } kind = Above;
    core::Color color;
    static SkyBox above() { return {}; }
    static SkyBox flat(core::Color c) { SkyBox s; s.kind = Flat; s.color = c; return s; }
    static SkyBox none() { SkyBox s; s.kind = None; return s; }
};
}  // namespace skybox

namespace scene {

// Decoded images for TextureLoader::{EarthBuiltin, ImagePath}.  The reference embeds
// assets/earthmap.jpg (image_texture.rs:11) and decodes JPEG through the `image` crate; here
// the embedding host registers decoded RGB8 pixels by name ("EarthBuiltin" or the path).
struct ImageData { uint32_t width = 0, height = 0; std::vector<uint8_t> rgb; };
void register_image(const std::string& name, ImageData img);
bool lookup_image(const std::string& name, ImageData* out);
ImageData synthetic_earth(uint32_t width = 1024, uint32_t height = 512);   // stand-in of the same shape as earthmap.jpg
// Baseline JPEG -> RGB8 (jpeg_decoder.cpp), what `image::open` / `load_from_memory` do for the
// reference's textures (image_texture.rs:18-31).  Throw Error on malformed or unsupported files.
ImageData decode_jpeg(const uint8_t* data, size_t size);
ImageData load_jpeg_file(const std::string& path);
// PNG -> RGB8 (png_decoder.cpp) and the format-sniffing loader `image::open` corresponds to.
ImageData decode_png(const uint8_t* data, size_t size);
ImageData decode_bmp(const uint8_t* data, size_t size);     // uncompressed 1/4/8/24/32-bit
ImageData decode_pnm(const uint8_t* data, size_t size);     // P2 P3 P5 P6
ImageData decode_any(const uint8_t* data, size_t size);     // sniffs the format like image::load_from_memory
ImageData load_image_file(const std::string& path);

// The flattened scene: owns the SoA arrays `desc` points into.
struct Scene {
    std::vector<B200rtPrimRef> prims;
    std::vector<B200rtMaterial> materials;
    std::vector<B200rtSphere> spheres;
    std::vector<B200rtRect> rects;
    std::vector<B200rtBox> boxes;
    std::vector<B200rtTexture> textures;
    std::vector<ImageData> image_store;
    std::vector<B200rtImage> images;
    std::vector<B200rtPerlin> perlin;
    B200rtSceneDesc desc{};
    void seal();   // (re)point desc at the vectors
};

struct SceneLoadObject { geometry::GeometricObject geometry; material::MaterialType material; };   // scene/mod.rs:23-27

struct SceneBuilder {                                                               // scene/mod.rs:79-138
    skybox::SkyBox skybox;                       // Default: Above
    std::vector<SceneLoadObject> objects;
    SceneBuilder& set_skybox(skybox::SkyBox s) { skybox = s; return *this; }
    template <class G, class M> void add(G g, M m) { objects.push_back({geometry::GeometricObject(std::move(g)), material::MaterialType(std::move(m))}); }
    // finalize: load every material's texture through a dedup cache (loader.rs:113-131),
    // then flatten.  `perlin_seed` seeds the Perlin tables the reference draws from an
    // unseeded thread_rng (perlin/mod.rs:73-85).  Throws Error on a missing image.
    std::unique_ptr<Scene> finalize(uint64_t perlin_seed = 0x9E3779B97F4A7C15ull) const;
    std::string to_json() const;                 // serde_json::to_writer_pretty, src/scenes.rs:140-143
    static SceneBuilder from_json(const std::string& text);   // serde_json::from_reader, src/scenes.rs:128-131
};
}  // namespace scene

namespace camera {
struct AspectRatio {                                                                // camera/mod.rs:139-152
    bool rational = true; uint32_t n = 3, d = 2; double f = 1.5;
    static AspectRatio Rational(uint32_t n, uint32_t d) { AspectRatio a; a.rational = true; a.n = n; a.d = d; return a; }
    static AspectRatio Ratio(double f) { AspectRatio a; a.rational = false; a.f = f; return a; }
    double as_float() const { return rational ? (double)n / (double)d : f; }
};
struct Dimmensions { size_t width = 0, height = 0; };                               // camera/mod.rs:133-137
struct Camera { double height = 0, width = 0; std::optional<double> lens_radius; double focal_length = 1.0; Dimmensions dimm; };
struct CameraPosition {                                                             // camera/mod.rs:62-85
    core::Point origin; double focus_length = 1; core::Vec3 w, u, v;
    static CameraPosition look_at(core::Point camera, core::Point target, core::Vec3 up);
};
struct CameraBuilder {                                                              // camera/mod.rs:13-60
    std::optional<size_t> height_, width_; std::optional<double> focal_length_, aperture_, vfov_; std::optional<AspectRatio> ratio_;
    CameraBuilder& vfov(double v) { vfov_ = v; return *this; }
    CameraBuilder& aperture(double a) { aperture_ = a; return *this; }
    CameraBuilder& focal_length(double f) { focal_length_ = f; return *this; }
    CameraBuilder& aspect_ratio(AspectRatio r) { ratio_ = r; return *this; }
    CameraBuilder& width(size_t w) { width_ = w; return *this; }
    CameraBuilder& height(size_t h) { height_ = h; return *this; }
    Camera build() const;   // throws Error unless exactly two of (height, width, ratio) are set
};
B200rtCamera to_abi(const Camera& c, const CameraPosition& p);
}  // namespace camera

// src/scenes.rs scene factories (the host-side generators for the BASELINE configs)
namespace scenes {
struct HostRng {   // seeded stand-in for rand::thread_rng() (src/scenes.rs:137): splitmix64
    uint64_t s;
    explicit HostRng(uint64_t seed) : s(seed) {}
    uint64_t next_u64() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double gen() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }   // rand's f64: 53 bits in [0,1)
    double range(double lo, double hi) { return lo + (hi - lo) * gen(); }               // core/math.rs:23-25
};
scene::SceneBuilder random_scene(HostRng& rng, bool night);        // src/scenes.rs:281-429
scene::SceneBuilder create_earth_demo();                           // :81-93
scene::SceneBuilder create_perlin_demo();                          // :65-79
scene::SceneBuilder create_box_light();                            // :94-127
scene::SceneBuilder create_cornell_box();                          // :23-63
scene::SceneBuilder create_scene();                                // :431-483 (demo)
// BASELINE config 4: the Weekend distribution on a (2G)^2 grid, ~4 G^2 spheres
scene::SceneBuilder scaled_random_scene(HostRng& rng, uint32_t grid_half);
// criterion bench lattice (benches/my_benchmark.rs:35-60): (2s)^3 jittered spheres, log-normal radii
scene::SceneBuilder bench_lattice(HostRng& rng, uint32_t side_len);
struct CameraSettings { size_t width = 640; double camera_fov = 20.0, camera_focal_length = 1.0, camera_aperture = 0.001; uint32_t ratio_n = 3, ratio_d = 2; };   // src/argparse.rs:3-10,126-146
void default_camera(const CameraSettings& args, camera::Camera* cam, camera::CameraPosition* pos);   // src/scenes.rs:214-231
}  // namespace scenes

}  // namespace raytracer
