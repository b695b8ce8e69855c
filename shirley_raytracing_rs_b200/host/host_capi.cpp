// host_capi.cpp — C shim of include/b200rt_host.h over raytracer.hpp.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/b200rt_host.h"
#include "raytracer.hpp"

using namespace raytracer;

struct B200rtHostScene {
    scene::SceneBuilder builder;
    std::unique_ptr<scene::Scene> flat;
};

namespace {
thread_local std::string g_host_error;
int host_fail(int code, const std::string& msg) { g_host_error = msg; return code; }
}  // namespace

extern "C" {

// b200rt_last_error() (b200rt.cu) reports device-side errors; host-side errors are kept here.
const char* b200rt_host_last_error(void) { return g_host_error.c_str(); }

int b200rt_host_scene_from_json(const char* json, size_t len, uint64_t perlin_seed, B200rtHostScene** out) {
    if (!json || !out) return host_fail(B200RT_EINVAL, "NULL argument");
    *out = nullptr;
    try {
        auto hs = std::make_unique<B200rtHostScene>();
        hs->builder = scene::SceneBuilder::from_json(std::string(json, len));
        hs->flat = hs->builder.finalize(perlin_seed);
        *out = hs.release();
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

int b200rt_host_scene_to_json(const B200rtHostScene* sc, char** out_json, size_t* out_len) {
    if (!sc || !out_json) return host_fail(B200RT_EINVAL, "NULL argument");
    try {
        std::string s = sc->builder.to_json();
        char* buf = (char*)malloc(s.size() + 1);
        if (!buf) return host_fail(B200RT_ENOMEM, "out of memory");
        memcpy(buf, s.c_str(), s.size() + 1);
        *out_json = buf;
        if (out_len) *out_len = s.size();
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

int b200rt_host_scene_named(const char* name, uint64_t seed, uint32_t param, B200rtHostScene** out) {
    if (!name || !out) return host_fail(B200RT_EINVAL, "NULL argument");
    *out = nullptr;
    try {
        auto hs = std::make_unique<B200rtHostScene>();
        scenes::HostRng rng(seed);
        std::string n = name;
        if (n == "random") hs->builder = scenes::random_scene(rng, false);
        else if (n == "random-night") hs->builder = scenes::random_scene(rng, true);
        else if (n == "earth") hs->builder = scenes::create_earth_demo();
        else if (n == "perlin") hs->builder = scenes::create_perlin_demo();
        else if (n == "box-light") hs->builder = scenes::create_box_light();
        else if (n == "cornell") hs->builder = scenes::create_cornell_box();
        else if (n == "demo") hs->builder = scenes::create_scene();
        else if (n == "scaled") hs->builder = scenes::scaled_random_scene(rng, param ? param : 158);
        else if (n == "lattice") hs->builder = scenes::bench_lattice(rng, param ? param : 8);
        else return host_fail(B200RT_EINVAL, "unknown scene `" + n + "`");
        hs->flat = hs->builder.finalize(seed ^ 0xA5A5A5A55A5A5A5Aull);
        *out = hs.release();
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

void b200rt_host_scene_destroy(B200rtHostScene* sc) { delete sc; }

const B200rtSceneDesc* b200rt_host_scene_desc(const B200rtHostScene* sc) { return sc && sc->flat ? &sc->flat->desc : nullptr; }

int b200rt_host_register_image(const char* name, uint32_t width, uint32_t height, const uint8_t* rgb8) {
    if (!name || !rgb8 || width == 0 || height == 0) return host_fail(B200RT_EINVAL, "bad image");
    scene::ImageData img; img.width = width; img.height = height; img.rgb.assign(rgb8, rgb8 + (size_t)width * height * 3);
    scene::register_image(name, std::move(img));
    return B200RT_OK;
}

int b200rt_host_decode_jpeg(const uint8_t* data, size_t size, uint32_t* width, uint32_t* height, uint8_t** rgb8) {
    if (!data || !width || !height || !rgb8) return host_fail(B200RT_EINVAL, "NULL argument");
    *rgb8 = nullptr;
    try {
        // image::load_from_memory sniffs the format: JPEG or PNG here
        scene::ImageData img = scene::decode_any(data, size);
        uint8_t* buf = (uint8_t*)malloc(img.rgb.size() ? img.rgb.size() : 1);
        if (!buf) return host_fail(B200RT_ENOMEM, "out of memory");
        memcpy(buf, img.rgb.data(), img.rgb.size());
        *width = img.width; *height = img.height; *rgb8 = buf;
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

int b200rt_host_decode_image(const uint8_t* data, size_t size, uint32_t* width, uint32_t* height, uint8_t** rgb8) {
    return b200rt_host_decode_jpeg(data, size, width, height, rgb8);   // one sniffing implementation
}

int b200rt_host_checkpoint_save(const char* path, const float* accum, uint32_t W, uint32_t H, uint32_t samples_done, uint64_t seed) {
    if (!path || !accum || W == 0 || H == 0) return host_fail(B200RT_EINVAL, "bad argument");
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return host_fail(B200RT_EIO, "cannot create " + tmp);
    const size_t n = (size_t)W * H * 4;
    uint32_t hdr[4] = {1u, W, H, samples_done};
    uint32_t crc = (uint32_t)crc32_z(0L, reinterpret_cast<const unsigned char*>(accum), n * sizeof(float));
    bool ok = fwrite("B200RTAC", 1, 8, f) == 8 && fwrite(hdr, 4, 4, f) == 4 && fwrite(&seed, 8, 1, f) == 1 &&
              fwrite(accum, sizeof(float), n, f) == n && fwrite(&crc, 4, 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return host_fail(B200RT_EIO, std::string("cannot write ") + path); }   // atomic replace
    return B200RT_OK;
}

int b200rt_host_checkpoint_load(const char* path, uint32_t* W, uint32_t* H, uint32_t* samples_done, uint64_t* seed, float** accum) {
    if (!path || !W || !H || !samples_done || !seed || !accum) return host_fail(B200RT_EINVAL, "NULL argument");
    *accum = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return host_fail(B200RT_EIO, std::string("cannot open ") + path);
    char magic[8]; uint32_t hdr[4]; uint64_t sd = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "B200RTAC", 8) == 0 && fread(hdr, 4, 4, f) == 4 && fread(&sd, 8, 1, f) == 1;
    if (!ok || hdr[0] != 1u || hdr[1] == 0 || hdr[2] == 0 || (uint64_t)hdr[1] * hdr[2] > 0xFFFFFFFFull) { fclose(f); return host_fail(B200RT_EINVAL, std::string(path) + " is not a b200rt checkpoint (version 1)"); }
    const size_t n = (size_t)hdr[1] * hdr[2] * 4;
    float* buf = (float*)malloc(n * sizeof(float));
    if (!buf) { fclose(f); return host_fail(B200RT_ENOMEM, "out of memory"); }
    uint32_t crc = 0;
    ok = fread(buf, sizeof(float), n, f) == n && fread(&crc, 4, 1, f) == 1;
    fclose(f);
    if (!ok || crc != (uint32_t)crc32_z(0L, reinterpret_cast<const unsigned char*>(buf), n * sizeof(float))) { free(buf); return host_fail(B200RT_EIO, std::string(path) + ": truncated or corrupt checkpoint (CRC mismatch)"); }
    *W = hdr[1]; *H = hdr[2]; *samples_done = hdr[3]; *seed = sd; *accum = buf;
    return B200RT_OK;
}

int b200rt_host_camera(const double from[3], const double at[3], const double up[3], double vfov, double focal_length, double aperture,
                       uint32_t image_width, uint32_t image_height, uint32_t ratio_num, uint32_t ratio_den, double focus_length, B200rtCamera* out) {
    if (!from || !at || !up || !out) return host_fail(B200RT_EINVAL, "NULL argument");
    try {
        camera::CameraBuilder b;
        b.vfov(vfov).focal_length(focal_length);
        if (aperture >= 0) b.aperture(aperture);
        if (image_width) b.width(image_width);
        if (image_height) b.height(image_height);
        if (ratio_num && ratio_den) b.aspect_ratio(camera::AspectRatio::Rational(ratio_num, ratio_den));
        camera::Camera cam = b.build();
        camera::CameraPosition pos = camera::CameraPosition::look_at(core::Point({from[0], from[1], from[2]}), core::Point({at[0], at[1], at[2]}), core::Vec3(up[0], up[1], up[2]));
        if (focus_length > 0) pos.focus_length = focus_length;
        *out = camera::to_abi(cam, pos);
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

int b200rt_host_default_camera(uint32_t width, double vfov, double focal_length, double aperture, uint32_t rn, uint32_t rd, B200rtCamera* out) {
    if (!out) return host_fail(B200RT_EINVAL, "NULL argument");
    try {
        scenes::CameraSettings a; a.width = width; a.camera_fov = vfov; a.camera_focal_length = focal_length; a.camera_aperture = aperture; a.ratio_n = rn; a.ratio_d = rd;
        camera::Camera cam; camera::CameraPosition pos;
        scenes::default_camera(a, &cam, &pos);
        *out = camera::to_abi(cam, pos);
        return B200RT_OK;
    } catch (const std::exception& e) { return host_fail(B200RT_EINVAL, e.what()); }
}

int b200rt_host_render_scene(const B200rtHostScene* sc, const B200rtCamera* cam, uint32_t samples, uint32_t max_depth, uint64_t seed, int device,
                             const char* output_png, uint8_t* rgb8_out, B200rtStats* stats) {
    if (!sc || !cam) return host_fail(B200RT_EINVAL, "NULL argument");
    B200rtScene* dev = nullptr;
    int rc = b200rt_scene_create(&sc->flat->desc, device, &dev);
    if (rc) return host_fail(rc, b200rt_last_error());
    size_t px = (size_t)cam->image_width * cam->image_height;
    B200rtRenderParams p{};
    p.samples = samples; p.max_depth = max_depth; p.seed = seed; p.device = -1;
    std::vector<uint8_t> rgb;
    uint8_t* dst = rgb8_out;
    if (!dst) { rgb.resize(px * 3); dst = rgb.data(); }
    rc = b200rt_render_rgb8(dev, cam, &p, dst, nullptr, stats);   // image.samples = samples, main.rs:86
    if (rc) { std::string m = b200rt_last_error(); b200rt_scene_destroy(dev); return host_fail(rc, m); }
    b200rt_scene_destroy(dev);
    if (output_png) {
        rc = b200rt_write_png(output_png, dst, cam->image_width, cam->image_height);
        if (rc) return host_fail(rc, std::string("cannot write ") + output_png);   // image.rs:42-43 panics; here: an error code
    }
    return B200RT_OK;
}

}  // extern "C"
