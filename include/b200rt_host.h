/*
 * b200rt_host.h — C shim over the C++ host mirror (shirley_raytracing_rs_b200/host/raytracer.hpp)
 * so non-C++ hosts (the Python harness, a Rust host through bindgen) can drive the same
 * builder the reference exposes: SceneBuilder (scene/mod.rs:79-138), the scene factories of
 * src/scenes.rs, default_camera (src/scenes.rs:214-231) and render_scene
 * (src/main.rs:65-130).  Scenes cross this shim in the reference's own serde JSON wire
 * format (src/scenes.rs:128-134,140-143).  No compute happens here except
 * b200rt_host_render_scene, which calls the device API of b200rt.h.
 */
#ifndef B200RT_HOST_H
#define B200RT_HOST_H

#include "b200rt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct B200rtHostScene B200rtHostScene;   /* a finalized scene: SceneBuilder + flattened arrays */

/* Message of the last failing b200rt_host_* call on this thread. */
const char* b200rt_host_last_error(void);

/* serde_json::from_reader::<SceneBuilder> then SceneBuilder::finalize (src/scenes.rs:128-132).
 * perlin_seed seeds the Perlin tables (the reference draws them from thread_rng). */
int  b200rt_host_scene_from_json(const char* json, size_t len, uint64_t perlin_seed, B200rtHostScene** out);
/* serde_json::to_writer_pretty(&scene) (src/scenes.rs:140-143); release with b200rt_free. */
int  b200rt_host_scene_to_json(const B200rtHostScene* scene, char** out_json, size_t* out_len);
/* Scene factories of src/scenes.rs: name = "random" (:281-429, day), "random-night", "earth"
 * (:81-93), "perlin" (:65-79), "box-light" (:94-127), "cornell" (:23-63), "demo" (:431-483),
 * "scaled" (BASELINE config 4; `param` = grid half-width G), "lattice"
 * (benches/my_benchmark.rs:35-60; `param` = side_len).  `seed` replaces thread_rng. */
int  b200rt_host_scene_named(const char* name, uint64_t seed, uint32_t param, B200rtHostScene** out);
void b200rt_host_scene_destroy(B200rtHostScene* scene);
/* The flattened description (borrowed; valid until the scene is destroyed). */
const B200rtSceneDesc* b200rt_host_scene_desc(const B200rtHostScene* scene);

/* Decoded RGB8 pixels for TextureLoader::EarthBuiltin (name "EarthBuiltin") or
 * ImagePath(name) (image_texture.rs:18-31); a registration takes precedence over decoding. */
int  b200rt_host_register_image(const char* name, uint32_t width, uint32_t height, const uint8_t* rgb8);
/* image::load_from_memory for a JPEG, baseline or progressive (image_texture.rs:18-21,28-31): RGB8 pixels, top row
 * first, in a buffer to release with b200rt_free.  ImagePath textures that were not registered
 * are decoded with the same routine at finalize; EarthBuiltin is decoded from the file named by
 * the environment variable B200RT_EARTHMAP (a copy of the reference's assets/earthmap.jpg) when
 * set, else it is the procedural stand-in. */
int  b200rt_host_decode_jpeg(const uint8_t* data, size_t size, uint32_t* width, uint32_t* height, uint8_t** rgb8);
/* The same for any format decoded here — JPEG or PNG (8-bit grey / grey+alpha / RGB / RGBA / palette,
 * non-interlaced; alpha dropped) — sniffed from the first bytes like image::load_from_memory. */
int  b200rt_host_decode_image(const uint8_t* data, size_t size, uint32_t* width, uint32_t* height, uint8_t** rgb8);

/* CameraBuilder + CameraPosition::look_at (camera/mod.rs:23-85).  aperture < 0 = None;
 * focus_length <= 0 keeps look_at's |camera - target|.  Exactly two of
 * (image_width, image_height, ratio) must be non-zero, like Dimmensions::from_two_of_three. */
int  b200rt_host_camera(const double look_from[3], const double look_at[3], const double up[3],
                        double vfov_degrees, double focal_length, double aperture,
                        uint32_t image_width, uint32_t image_height, uint32_t ratio_num, uint32_t ratio_den,
                        double focus_length, B200rtCamera* out);
/* default_camera (src/scenes.rs:214-231): from (13,2,3) to the origin, focus_length 10. */
int  b200rt_host_default_camera(uint32_t width, double vfov_degrees, double focal_length, double aperture,
                                uint32_t ratio_num, uint32_t ratio_den, B200rtCamera* out);

/* Accumulation-buffer checkpoint for long progressive runs (BASELINE config 5: 4096 spp): the W*H float4
 * {sum r, g, b, n} buffer of b200rt_render plus what is needed to continue the SAME sample sequence —
 * the seed and the number of samples done (the next call's sample_offset).  Little-endian file:
 * "B200RTAC", u32 version, u32 W, u32 H, u32 samples_done, u64 seed, W*H*4 f32, u32 CRC-32 of the floats.
 * _load allocates *accum (release with b200rt_free) and fails on a bad magic, size or CRC. */
int  b200rt_host_checkpoint_save(const char* path, const float* accum, uint32_t width, uint32_t height,
                                 uint32_t samples_done, uint64_t seed);
int  b200rt_host_checkpoint_load(const char* path, uint32_t* width, uint32_t* height, uint32_t* samples_done,
                                 uint64_t* seed, float** accum);

/* render_scene (src/main.rs:65-130) end to end: upload, render, resolve, write the PNG.
 * output_png may be NULL (no file).  rgb8_out may be NULL, else W*H*3 bytes, top row first. */
int  b200rt_host_render_scene(const B200rtHostScene* scene, const B200rtCamera* camera,
                              uint32_t samples, uint32_t max_depth, uint64_t seed, int device,
                              const char* output_png, uint8_t* rgb8_out, B200rtStats* stats);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_HOST_H */
