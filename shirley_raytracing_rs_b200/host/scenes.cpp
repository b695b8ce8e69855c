// scenes.cpp — host-side scene factories mirroring src/scenes.rs (the generators for the
// BASELINE configs).  The reference draws from an unseeded thread_rng; these take a seeded
// generator and make the draws in the same order.
#include <algorithm>

#include "raytracer.hpp"

namespace raytracer {
namespace scenes {
using core::Color; using core::Point; using core::Vec3;
using geometry::RectBox; using geometry::Sphere; using geometry::xy_rect; using geometry::xz_rect; using geometry::yz_rect;
using material::Dielectric; using material::DiffuseLight; using material::FairyLight; using material::Lambertian; using material::Metal;
using material::texture::TextureLoader;
using scene::SceneBuilder; using skybox::SkyBox;

static void create_ground_checker(SceneBuilder& scene) {   // src/scenes.rs:233-249
    TextureLoader ground = TextureLoader::checker(10.0, TextureLoader::solid(0.2, 0.3, 0.1), TextureLoader::solid(0.9, 0.9, 0.9));
    double rect = 30.0;
    scene.add(xz_rect(-rect, rect, -rect, rect, -0.0001), Lambertian(ground));
}

static void create_fancy_ground(SceneBuilder& scene, double rect_size = 30.0) {   // src/scenes.rs:251-279
    const double TOP_COAT_DEPTH = 0.01, LAYER_SEP = 0.01;
    Lambertian lower(TextureLoader::checker(3.0, TextureLoader::noise(1.0), TextureLoader::solid(0.1, 0.1, 0.1)));
    scene.add(xz_rect(-rect_size, rect_size, -rect_size, rect_size, -TOP_COAT_DEPTH - LAYER_SEP), lower);
    scene.add(RectBox::create(Point({-rect_size, -TOP_COAT_DEPTH, -rect_size}), Point({rect_size, 0.0, rect_size})), Dielectric{1.0});
}

static Vec3 gen_vec3(HostRng& rng) { double x = rng.gen(), y = rng.gen(), z = rng.gen(); return {x, y, z}; }   // Standard: Distribution<Vec3>, vec3.rs:34-38

enum BallType { BColor, BSphereLight, BGlass, BMetal, BChecker, BMarble };
static BallType choose_weighted(HostRng& rng, const double w[6]) {   // SliceRandom::choose_weighted
    double total = 0; for (int i = 0; i < 6; ++i) total += w[i];
    double x = rng.gen() * total, acc = 0;
    for (int i = 0; i < 6; ++i) { acc += w[i]; if (x < acc) return (BallType)i; }
    return BColor;
}

static void add_ball(SceneBuilder& scene, HostRng& rng, BallType item, const Sphere& sphere, double radius) {   // src/scenes.rs:384-425
    switch (item) {
        case BColor: { Vec3 albedo = gen_vec3(rng) * gen_vec3(rng); scene.add(sphere, Lambertian(TextureLoader::solid_from_vec(albedo))); break; }
        case BSphereLight: { Vec3 albedo = (gen_vec3(rng) * gen_vec3(rng)).scale(5.0); scene.add(sphere, FairyLight(TextureLoader::solid_from_vec(albedo))); break; }
        case BGlass: scene.add(sphere, Dielectric{1.5}); break;
        case BMetal: {
            double r = rng.range(0.5, 1.0), g = rng.range(0.5, 1.0), b = rng.range(0.5, 1.0);
            double fuzz = rng.range(0.0, 0.5);
            scene.add(sphere, Metal(Color({r, g, b}), fuzz));
            break;
        }
        case BChecker: {
            Vec3 c = gen_vec3(rng) * gen_vec3(rng);
            scene.add(sphere, Lambertian(TextureLoader::checker(8.0 / radius, TextureLoader::solid_from_vec(c), TextureLoader::solid(0.9, 0.9, 0.9))));
            break;
        }
        default: scene.add(sphere, Lambertian(TextureLoader::noise(16.0))); break;
    }
}

SceneBuilder random_scene(HostRng& rng, bool night) {   // src/scenes.rs:281-429
    SceneBuilder scene;
    if (night) scene.set_skybox(SkyBox::none());
    if (night) create_ground_checker(scene); else create_fancy_ground(scene);

    std::vector<Sphere> balls;
    auto check_fit_ball = [&](Sphere s) {   // :295-306
        double orig = s.radius;
        for (const Sphere& other : balls) {
            double dist = (other.center.v - s.center.v).length();
            double rem = dist - other.radius;
            s.radius = std::min(s.radius, rem);
        }
        double delta = orig - s.radius;
        s.center = Point(s.center.v - Vec3(0.0, delta, 0.0));
        balls.push_back(s);
        return s;
    };
    scene.add(check_fit_ball(Sphere{Point({0.0, 1.0, 0.0}), 1.0}), Dielectric{1.5});
    if (night) scene.add(check_fit_ball(Sphere{Point({-4.0, 1.0, 0.0}), 1.0}), FairyLight(TextureLoader::solid_from_vec(Vec3(0.7, 0.6, 0.5).scale(1.3))));
    else scene.add(check_fit_ball(Sphere{Point({-4.0, 1.0, 0.0}), 1.0}), Lambertian(TextureLoader::solid(0.4, 0.2, 0.1)));
    scene.add(check_fit_ball(Sphere{Point({4.0, 1.0, 0.0}), 1.0}), Metal(Color({0.7, 0.6, 0.5}), std::nullopt));

    const double weights[6] = {4.0, night ? 4.0 : 0.0, 1.0, 4.0, 0.3, 0.0};   // :360-367
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            BallType item = choose_weighted(rng, weights);
            double radius = rng.range(0.05, 0.25);
            double cx = (double)a + 0.9 * rng.gen();
            double cz = (double)b + 0.9 * rng.gen();
            Point center({cx, radius, cz});
            Vec3 keepout(3.0, radius, 0.0);
            if ((center.v - keepout).length() <= 0.9) continue;
            Sphere sphere = check_fit_ball(Sphere{center, radius});
            add_ball(scene, rng, item, sphere, radius);
        }
    return scene;
}

// BASELINE config 4: the same per-cell distribution on a (2G)^2 grid.
// check_fit_ball is O(N^2) and is skipped; radius U[0.05,0.25) on a unit grid with 0.9
// jitter mostly avoids overlap already.
SceneBuilder scaled_random_scene(HostRng& rng, uint32_t G) {
    SceneBuilder scene;
    create_fancy_ground(scene, (double)G + 8.0);
    scene.add(Sphere{Point({0.0, 1.0, 0.0}), 1.0}, Dielectric{1.5});
    scene.add(Sphere{Point({-4.0, 1.0, 0.0}), 1.0}, Lambertian(TextureLoader::solid(0.4, 0.2, 0.1)));
    scene.add(Sphere{Point({4.0, 1.0, 0.0}), 1.0}, Metal(Color({0.7, 0.6, 0.5}), std::nullopt));
    const double weights[6] = {4.0, 0.0, 1.0, 4.0, 0.3, 0.0};
    int g = (int)G;
    scene.objects.reserve((size_t)4 * G * G + 8);
    for (int a = -g; a < g; ++a)
        for (int b = -g; b < g; ++b) {
            BallType item = choose_weighted(rng, weights);
            double radius = rng.range(0.05, 0.25);
            double cx = (double)a + 0.9 * rng.gen();
            double cz = (double)b + 0.9 * rng.gen();
            Point center({cx, radius, cz});
            // keep-outs around the three big spheres
            bool skip = false;
            for (double bx : {0.0, -4.0, 4.0}) if ((center.v - Vec3(bx, radius, 0.0)).length() <= 1.1) skip = true;
            if (skip) continue;
            add_ball(scene, rng, item, Sphere{center, radius}, radius);
        }
    return scene;
}

SceneBuilder bench_lattice(HostRng& rng, uint32_t side_len) {   // benches/my_benchmark.rs:35-60
    SceneBuilder scene;
    int s = (int)side_len;
    Lambertian grey(TextureLoader::solid(0.5, 0.5, 0.5));
    for (int x = -s; x < s; ++x) for (int y = -s; y < s; ++y) for (int z = -s; z < s; ++z) {
        Vec3 c((double)x + rng.range(-1, 1), (double)y + rng.range(-1, 1), (double)z + rng.range(-1, 1));
        double u1 = std::max(rng.gen(), 1e-300), u2 = rng.gen();
        double n = std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * 3.14159265358979323846 * u2);   // N(0,1)
        double radius = std::exp(0.5 + 0.5 * n);                                                    // LogNormal(0.5, 0.5)
        scene.add(Sphere{Point(c), radius}, grey);
    }
    return scene;
}

SceneBuilder create_earth_demo() {   // src/scenes.rs:81-93
    SceneBuilder scene;
    create_ground_checker(scene);
    scene.add(Sphere{Point({4.0, 1.0, 1.0}), 1.0}, Lambertian(TextureLoader::earth_builtin()));
    return scene;
}
SceneBuilder create_perlin_demo() {   // src/scenes.rs:65-79
    SceneBuilder scene;
    create_ground_checker(scene);
    scene.add(Sphere{Point({0.0, 2.0, -0.0}), 2.0}, Lambertian(TextureLoader::noise(4.0)));
    return scene;
}
SceneBuilder create_box_light() {   // src/scenes.rs:94-127
    SceneBuilder scene;
    scene.set_skybox(SkyBox::none());
    create_ground_checker(scene);
    scene.add(Sphere{Point({0.0, 2.0, -0.0}), 2.0}, Lambertian(TextureLoader::noise(4.0)));
    scene.add(xz_rect(3.0, 5.0, 1.0, 3.0, 3.5), DiffuseLight(TextureLoader::solid(4.0, 4.0, 4.0)));
    return scene;
}
SceneBuilder create_cornell_box() {   // src/scenes.rs:23-63
    SceneBuilder scene;
    scene.set_skybox(SkyBox::none());
    Lambertian red(TextureLoader::solid(0.65, 0.05, 0.05)), white(TextureLoader::solid(0.73, 0.73, 0.73)), green(TextureLoader::solid(0.12, 0.45, 0.15));
    FairyLight light(TextureLoader::solid(15.0, 15.0, 15.0));
    double box = 555.0;
    scene.add(yz_rect(0.0, box, 0.0, box, box), green);
    scene.add(yz_rect(0.0, box, 0.0, box, 0.0), red);
    scene.add(xz_rect(213.0, 343.0, 227.0, 332.0, 554.0), light);
    scene.add(xz_rect(0.0, box, 0.0, box, 0.0), white);
    scene.add(xz_rect(0.0, box, 0.0, box, box), white);
    scene.add(xy_rect(0.0, box, 0.0, box, box), white);
    scene.add(RectBox::create(Point({130.0, 0.0, 65.0}), Point({295.0, 165.0, 230.0})), white);
    scene.add(RectBox::create(Point({265.0, 0.0, 295.0}), Point({430.0, 330.0, 460.0})), white);
    return scene;
}
SceneBuilder create_scene() {   // src/scenes.rs:431-483 (`render demo`)
    SceneBuilder scene;
    scene.add(Sphere{Point({0.0, -100.5, -1.0}), 100.0}, Lambertian(TextureLoader::solid(0.8, 0.8, 0.0)));
    scene.add(Sphere{Point({0.0, 0.0, -1.0}), 0.5}, Lambertian(TextureLoader::solid(0.1, 0.2, 0.5)));
    scene.add(Sphere{Point({-1.0, 0.0, -1.0}), 0.5}, Dielectric{1.5});
    scene.add(Sphere{Point({-1.0, 0.0, -1.0}), -0.4}, Dielectric{1.5});   // negative radius: inverted bbox, never hit through the tree
    scene.add(Sphere{Point({1.0, 0.0, -1.0}), 0.5}, Metal(Color({0.8, 0.6, 0.2}), 0.0));
    return scene;
}

void default_camera(const CameraSettings& args, camera::Camera* cam, camera::CameraPosition* pos) {   // src/scenes.rs:214-231
    camera::CameraBuilder b;
    b.vfov(args.camera_fov).focal_length(args.camera_focal_length).aperture(args.camera_aperture).width(args.width)
        .aspect_ratio(camera::AspectRatio::Rational(args.ratio_n, args.ratio_d));
    *cam = b.build();
    *pos = camera::CameraPosition::look_at(Point({13.0, 2.0, 3.0}), Point({0.0, 0.0, 0.0}), Vec3(0.0, 1.0, 0.0));
    pos->focus_length = 10.0;
}

}  // namespace scenes
}  // namespace raytracer
