// scene.cpp — SceneBuilder::finalize (flatten to include/b200rt.h arrays), texture dedup,
// Perlin tables, image registry, camera builder.
#include <cstring>
#include <mutex>
#include <sstream>

#include <cstdlib>

#include <dlfcn.h>

#include "raytracer.hpp"

namespace raytracer {

// ---------------------------------------------------------------------------------------
// TextureLoader dedup key: the reference hashes f64 bit patterns (texture/mod.rs:42-84)
// ---------------------------------------------------------------------------------------
static std::string bits(double d) { uint64_t u; std::memcpy(&u, &d, 8); char b[20]; snprintf(b, sizeof b, "%016llx", (unsigned long long)u); return b; }
std::string material::texture::TextureLoader::key() const {
    switch (kind) {
        case Solid: return "S(" + bits(color.v.x) + "," + bits(color.v.y) + "," + bits(color.v.z) + ")";
        case ImagePath: return "I(" + path + ")";
        case Perlin: return "P(" + bits(scalar) + ")";
        case EarthBuiltin: return "E";
        default: return "C(" + bits(scalar) + "," + (odd ? odd->key() : "") + "," + (even ? even->key() : "") + ")";
    }
}

namespace scene {

// ---------------------------------------------------------------------------------------
// image registry
// ---------------------------------------------------------------------------------------
static std::mutex g_img_mu;
static std::map<std::string, ImageData>& registry() { static std::map<std::string, ImageData> r; return r; }
void register_image(const std::string& name, ImageData img) { std::lock_guard<std::mutex> lk(g_img_mu); registry()[name] = std::move(img); }
bool lookup_image(const std::string& name, ImageData* out) {
    std::lock_guard<std::mutex> lk(g_img_mu);
    auto it = registry().find(name);
    if (it == registry().end()) return false;
    *out = it->second;
    return true;
}

// A deterministic procedural stand-in with the shape of assets/earthmap.jpg (1024x512 RGB8):
// fBm "continents" over an ocean, ice caps at the poles.  Used when the embedding host has
// not registered the real decoded JPEG (the asset is not part of this repository).
// <directory of libb200rt.so>/assets/<file>: where the data the reference embeds in its binary ships here
static std::string builtin_asset_path(const char* file) {
    Dl_info info;
    std::string dir = ".";
    if (dladdr(reinterpret_cast<const void*>(&builtin_asset_path), &info) && info.dli_fname) {
        std::string so = info.dli_fname;
        size_t slash = so.find_last_of('/');
        dir = slash == std::string::npos ? "." : so.substr(0, slash);
    }
    return dir + "/assets/" + file;
}

ImageData synthetic_earth(uint32_t W, uint32_t H) {
    ImageData im; im.width = W; im.height = H; im.rgb.resize((size_t)W * H * 3);
    auto hash = [](int x, int y, int z) { uint32_t h = (uint32_t)x * 374761393u + (uint32_t)y * 668265263u + (uint32_t)z * 2147483647u; h = (h ^ (h >> 13)) * 1274126177u; return (double)((h ^ (h >> 16)) & 0xFFFFFF) / 16777216.0; };
    auto vnoise = [&](double x, double y, double z) {
        int xi = (int)std::floor(x), yi = (int)std::floor(y), zi = (int)std::floor(z);
        double fx = x - xi, fy = y - yi, fz = z - zi;
        fx = fx * fx * (3 - 2 * fx); fy = fy * fy * (3 - 2 * fy); fz = fz * fz * (3 - 2 * fz);
        double acc = 0;
        for (int dx = 0; dx < 2; ++dx) for (int dy = 0; dy < 2; ++dy) for (int dz = 0; dz < 2; ++dz)
            acc += (dx ? fx : 1 - fx) * (dy ? fy : 1 - fy) * (dz ? fz : 1 - fz) * hash(xi + dx, yi + dy, zi + dz);
        return acc;
    };
    const double PI = 3.14159265358979323846;
    for (uint32_t j = 0; j < H; ++j) {
        double lat = PI * ((j + 0.5) / H);            // 0 = north pole
        for (uint32_t i = 0; i < W; ++i) {
            double lon = 2 * PI * ((i + 0.5) / W);
            double x = std::sin(lat) * std::cos(lon), y = std::cos(lat), z = std::sin(lat) * std::sin(lon);
            double f = 0, amp = 0.5, fr = 1.6;
            for (int o = 0; o < 6; ++o) { f += amp * vnoise(x * fr + 11.3, y * fr + 4.7, z * fr + 7.9); amp *= 0.5; fr *= 2.0; }
            double r, g, b;
            if (f > 0.52) { double k = std::min(1.0, (f - 0.52) * 6.0); r = 0.25 + 0.45 * k; g = 0.45 + 0.10 * k; b = 0.15 + 0.10 * k; }   // land
            else { double k = f / 0.52; r = 0.02 + 0.05 * k; g = 0.10 + 0.20 * k; b = 0.35 + 0.35 * k; }                                  // ocean
            double ice = std::fabs(y) > 0.9 ? std::min(1.0, (std::fabs(y) - 0.9) * 14.0) : 0.0;
            r = r + (0.95 - r) * ice; g = g + (0.96 - g) * ice; b = b + (0.98 - b) * ice;
            uint8_t* px = &im.rgb[((size_t)j * W + i) * 3];
            px[0] = (uint8_t)(r * 255.0 + 0.5); px[1] = (uint8_t)(g * 255.0 + 0.5); px[2] = (uint8_t)(b * 255.0 + 0.5);
        }
    }
    return im;
}

void Scene::seal() {
    images.resize(image_store.size());
    for (size_t i = 0; i < image_store.size(); ++i) { images[i].width = image_store[i].width; images[i].height = image_store[i].height; images[i].rgb8 = image_store[i].rgb.data(); }
    desc.abi_version = B200RT_ABI_VERSION;
    desc.n_prims = (uint32_t)prims.size(); desc.prims = prims.data(); desc.materials = materials.data();
    desc.n_spheres = (uint32_t)spheres.size(); desc.spheres = spheres.data();
    desc.n_rects = (uint32_t)rects.size(); desc.rects = rects.data();
    desc.n_boxes = (uint32_t)boxes.size(); desc.boxes = boxes.data();
    desc.n_textures = (uint32_t)textures.size(); desc.textures = textures.data();
    desc.n_images = (uint32_t)images.size(); desc.images = images.data();
    desc.n_perlin = (uint32_t)perlin.size(); desc.perlin = perlin.data();
}

namespace {
using material::texture::TextureLoader;

// Perlin::new (perlin/mod.rs:73-85) with a seeded generator: 256 x U(-1,1)^3 gradients,
// then three Fisher-Yates permutations `for idx in (1..256).rev(): swap(idx, gen_range(0..=idx))`.
B200rtPerlin make_perlin(uint64_t seed) {
    scenes::HostRng rng(seed);
    B200rtPerlin p{};
    for (int i = 0; i < 256; ++i) for (int k = 0; k < 3; ++k) p.ranfloat[i][k] = (float)rng.range(-1.0, 1.0);
    auto perm = [&](uint8_t* out) {
        for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
        for (int idx = 255; idx >= 1; --idx) { int target = (int)(rng.next_u64() % (uint64_t)(idx + 1)); std::swap(out[idx], out[target]); }
    };
    perm(p.perm_x); perm(p.perm_y); perm(p.perm_z);
    return p;
}

struct TextureManager {   // loader.rs:108-131
    Scene& sc; uint64_t perlin_seed;
    std::map<std::string, int> cache;
    int load(const TextureLoader& t) {
        std::string k = t.key();
        auto it = cache.find(k);
        if (it != cache.end()) return it->second;
        B200rtTexture x{}; x.odd = x.even = x.image = -1;
        switch (t.kind) {
            case TextureLoader::Solid: x.kind = B200RT_TEX_SOLID; x.rgb[0] = (float)t.color.v.x; x.rgb[1] = (float)t.color.v.y; x.rgb[2] = (float)t.color.v.z; break;
            case TextureLoader::Perlin:
                x.kind = B200RT_TEX_PERLIN; x.scalar = (float)t.scalar; x.image = (int32_t)sc.perlin.size();
                sc.perlin.push_back(make_perlin(perlin_seed + 0x632BE59BD9B4E019ull * (uint64_t)(sc.perlin.size() + 1)));   // one table per distinct Perlin(scale)
                break;
            case TextureLoader::EarthBuiltin:
            case TextureLoader::ImagePath: {
                std::string name = t.kind == TextureLoader::EarthBuiltin ? "EarthBuiltin" : t.path;
                ImageData img;
                if (!lookup_image(name, &img)) {
                    if (t.kind == TextureLoader::EarthBuiltin) {
                        // the reference embeds assets/earthmap.jpg in its binary (image_texture.rs:11,18-20); here the
                        // same file ships next to the library (assets/earthmap.jpg) and is decoded on first use.
                        // B200RT_EARTHMAP overrides the path; B200RT_EARTHMAP=synthetic selects a procedural stand-in.
                        const char* env = getenv("B200RT_EARTHMAP");
                        if (env && !strcmp(env, "synthetic")) img = synthetic_earth();
                        else if (env && *env) img = load_image_file(env);
                        else img = load_image_file(builtin_asset_path("earthmap.jpg"));
                        register_image(name, img);      // decode once per process
                    } else {
                        img = load_image_file(t.path);   // image_texture.rs:23-26 `image::open(path)?` (JPEG and PNG here)
                    }
                }
                x.kind = B200RT_TEX_IMAGE; x.image = (int32_t)sc.image_store.size();
                sc.image_store.push_back(std::move(img));
                break;
            }
            default: {   // Checker: children first so they get lower indices
                if (!t.odd || !t.even) throw Error("checker texture without children");
                int o = load(*t.odd), e = load(*t.even);
                x.kind = B200RT_TEX_CHECKER; x.scalar = (float)t.scalar; x.odd = o; x.even = e;
            }
        }
        int idx = (int)sc.textures.size();
        sc.textures.push_back(x);
        cache[k] = idx;
        return idx;
    }
};
}  // namespace

std::unique_ptr<Scene> SceneBuilder::finalize(uint64_t perlin_seed) const {
    auto sc = std::make_unique<Scene>();
    TextureManager tm{*sc, perlin_seed, {}};
    for (const SceneLoadObject& o : objects) {
        B200rtPrimRef ref{};
        if (auto s = std::get_if<geometry::Sphere>(&o.geometry)) {
            ref.type = B200RT_PRIM_SPHERE; ref.index = (uint32_t)sc->spheres.size();
            sc->spheres.push_back({(float)s->center.v.x, (float)s->center.v.y, (float)s->center.v.z, (float)s->radius});
        } else if (auto r = std::get_if<geometry::Rect>(&o.geometry)) {
            ref.type = r->kind; ref.index = (uint32_t)sc->rects.size();
            B200rtRect q{}; q.d1_min = (float)r->d1_min; q.d1_max = (float)r->d1_max; q.d2_min = (float)r->d2_min; q.d2_max = (float)r->d2_max; q.offset = (float)r->offset; q.kind = r->kind;
            sc->rects.push_back(q);
        } else {
            const geometry::RectBox& b = std::get<geometry::RectBox>(o.geometry);
            ref.type = B200RT_PRIM_BOX; ref.index = (uint32_t)sc->boxes.size();
            B200rtBox q{}; q.min[0] = (float)b.min.v.x; q.min[1] = (float)b.min.v.y; q.min[2] = (float)b.min.v.z; q.max[0] = (float)b.max.v.x; q.max[1] = (float)b.max.v.y; q.max[2] = (float)b.max.v.z;
            sc->boxes.push_back(q);
        }
        sc->prims.push_back(ref);
        B200rtMaterial m{}; m.texture = -1;
        if (auto x = std::get_if<material::Metal>(&o.material)) {
            m.kind = B200RT_MAT_METAL; m.albedo[0] = (float)x->albedo.v.x; m.albedo[1] = (float)x->albedo.v.y; m.albedo[2] = (float)x->albedo.v.z; m.param = (float)x->fuzz;
        } else if (auto x2 = std::get_if<material::Dielectric>(&o.material)) {
            m.kind = B200RT_MAT_DIELECTRIC; m.param = (float)x2->ir;
        } else if (auto x3 = std::get_if<material::Lambertian>(&o.material)) {
            m.kind = B200RT_MAT_LAMBERTIAN; m.texture = tm.load(x3->albedo);
        } else if (auto x4 = std::get_if<material::DiffuseLight>(&o.material)) {
            m.kind = B200RT_MAT_DIFFUSE_LIGHT; m.texture = tm.load(x4->albedo);
        } else {
            m.kind = B200RT_MAT_FAIRY_LIGHT; m.texture = tm.load(std::get<material::FairyLight>(o.material).albedo);
        }
        sc->materials.push_back(m);
    }
    sc->desc.skybox.kind = skybox.kind == skybox::SkyBox::Above ? B200RT_SKY_ABOVE : (skybox.kind == skybox::SkyBox::Flat ? B200RT_SKY_FLAT : B200RT_SKY_NONE);
    sc->desc.skybox.rgb[0] = (float)skybox.color.v.x; sc->desc.skybox.rgb[1] = (float)skybox.color.v.y; sc->desc.skybox.rgb[2] = (float)skybox.color.v.z;
    sc->seal();
    return sc;
}
}  // namespace scene

// ---------------------------------------------------------------------------------------
// camera/mod.rs
// ---------------------------------------------------------------------------------------
namespace camera {
CameraPosition CameraPosition::look_at(core::Point cam, core::Point target, core::Vec3 up) {   // camera/mod.rs:73-85
    CameraPosition p;
    core::Vec3 w = cam.v - target.v;
    p.focus_length = w.unit_mut();
    core::Vec3 u = up.cross(w).unit();
    core::Vec3 v = w.cross(u);
    p.origin = cam; p.w = w; p.u = u; p.v = v;
    return p;
}

Camera CameraBuilder::build() const {   // camera/mod.rs:44-60 + Dimmensions::from_two_of_three :174-208
    Dimmensions dimm; AspectRatio ratio;
    if (!height_ && width_ && ratio_) { ratio = *ratio_; dimm.width = *width_; dimm.height = (size_t)((double)*width_ / ratio.as_float()); }
    else if (height_ && !width_ && ratio_) { ratio = *ratio_; dimm.height = *height_; dimm.width = (size_t)((double)*height_ * ratio.as_float()); }
    else if (height_ && width_ && !ratio_) { dimm.width = *width_; dimm.height = *height_; ratio = AspectRatio::Rational((uint32_t)*height_, (uint32_t)*width_); }   // sic: (h, w) in the reference
    else throw Error("could not construct dimm: require exactly 2 of (height, width, aspect ratio)");
    const double DEFAULT_FOCAL_LENGTH = 1.0;
    double theta = vfov_.value_or(DEFAULT_FOCAL_LENGTH) * 3.14159265358979323846 / 180.0;   // degrees_to_radians; the default really is DEFAULT_FOCAL_LENGTH (:46)
    double h = std::tan(theta / 2.0);
    Camera c;
    c.height = 2.0 * h;
    c.width = ratio.as_float() * c.height;
    if (aperture_) c.lens_radius = *aperture_ / 2.0;
    c.focal_length = focal_length_.value_or(DEFAULT_FOCAL_LENGTH);
    c.dimm = dimm;
    return c;
}

B200rtCamera to_abi(const Camera& c, const CameraPosition& p) {
    B200rtCamera o{};
    o.height = c.height; o.width = c.width; o.lens_radius = c.lens_radius ? *c.lens_radius : -1.0; o.focal_length = c.focal_length;
    o.image_width = (uint32_t)c.dimm.width; o.image_height = (uint32_t)c.dimm.height;
    o.origin[0] = p.origin.v.x; o.origin[1] = p.origin.v.y; o.origin[2] = p.origin.v.z;
    o.focus_length = p.focus_length;
    o.w[0] = p.w.x; o.w[1] = p.w.y; o.w[2] = p.w.z; o.u[0] = p.u.x; o.u[1] = p.u.y; o.u[2] = p.u.z; o.v[0] = p.v.x; o.v[1] = p.v.y; o.v[2] = p.v.z;
    return o;
}
}  // namespace camera

}  // namespace raytracer
