// json.cpp — the reference's scene wire format: serde_json of SceneBuilder
// (src/scenes.rs:128-134,140-143; derives at scene/mod.rs:23-27,79-83, geometry/object.rs:9-16,
// material/material_type.rs:20-27, material/texture/loader.rs:17-28, skybox/mod.rs:11-16).
// serde defaults: externally tagged enums, transparent newtypes, Vec3 = {"vec":[x,y,z]}
// (nalgebra serialises a Vector3 as a flat 3-array).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "raytracer.hpp"

namespace raytracer {
namespace {

struct JValue {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false; double n = 0; std::string s;
    std::vector<JValue> a;
    std::vector<std::pair<std::string, JValue>> o;
    const JValue* get(const std::string& k) const { for (auto& kv : o) if (kv.first == k) return &kv.second; return nullptr; }
    const JValue& at(const std::string& k) const { const JValue* v = get(k); if (!v) throw Error("scene json: missing field `" + k + "`"); return *v; }
};

struct Parser {
    const char* p; const char* e;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    [[noreturn]] void bad(const char* m) { throw Error(std::string("scene json: ") + m); }
    JValue value() {
        ws();
        if (p >= e) bad("unexpected end");
        JValue v;
        char c = *p;
        if (c == '{') {
            ++p; v.t = JValue::Obj; ws();
            if (p < e && *p == '}') { ++p; return v; }
            for (;;) {
                ws(); if (p >= e || *p != '"') bad("expected key");
                std::string k = str(); ws();
                if (p >= e || *p != ':') bad("expected ':'");
                ++p;
                v.o.push_back({k, value()}); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                bad("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++p; v.t = JValue::Arr; ws();
            if (p < e && *p == ']') { ++p; return v; }
            for (;;) {
                v.a.push_back(value()); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                bad("expected ',' or ']'");
            }
        } else if (c == '"') { v.t = JValue::Str; v.s = str(); }
        else if (!strncmp(p, "true", 4) && e - p >= 4) { v.t = JValue::Bool; v.b = true; p += 4; }
        else if (!strncmp(p, "false", 5) && e - p >= 5) { v.t = JValue::Bool; v.b = false; p += 5; }
        else if (!strncmp(p, "null", 4) && e - p >= 4) { p += 4; }
        else {
            char* end = nullptr;
            v.t = JValue::Num; v.n = strtod(p, &end);
            if (end == p) bad("bad token");
            p = end;
        }
        return v;
    }
    std::string str() {
        std::string out; ++p;
        while (p < e && *p != '"') {
            if (*p == '\\' && p + 1 < e) {
                ++p;
                switch (*p) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': { if (e - p < 5) bad("bad \\u"); unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); p += 4;
                        if (cp < 0x80) out += (char)cp; else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); } else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); } break; }
                    default: out += *p; }
                ++p;
            } else out += *p++;
        }
        if (p >= e) bad("unterminated string");
        ++p;
        return out;
    }
};

core::Vec3 vec3_of(const JValue& v) {
    const JValue& a = v.at("vec");
    if (a.t != JValue::Arr || a.a.size() != 3) throw Error("scene json: `vec` must be a 3-array");
    return {a.a[0].n, a.a[1].n, a.a[2].n};
}
// an externally tagged enum value: either "Unit" or {"Variant": payload}
void variant_of(const JValue& v, std::string* tag, const JValue** payload) {
    static const JValue null_value;
    if (v.t == JValue::Str) { *tag = v.s; *payload = &null_value; return; }
    if (v.t == JValue::Obj && v.o.size() == 1) { *tag = v.o[0].first; *payload = &v.o[0].second; return; }
    throw Error("scene json: expected an externally tagged enum");
}
using material::texture::TextureLoader;
TextureLoader texture_of(const JValue& v) {
    std::string tag; const JValue* p;
    variant_of(v, &tag, &p);
    if (tag == "Solid") return TextureLoader::solid_from_vec(vec3_of(*p));
    if (tag == "ImagePath") return TextureLoader::image_path(p->s);
    if (tag == "Perlin") return TextureLoader::noise(p->n);
    if (tag == "EarthBuiltin") return TextureLoader::earth_builtin();
    if (tag == "Checker") return TextureLoader::checker(p->at("size").n, texture_of(p->at("odd")), texture_of(p->at("even")));
    throw Error("scene json: unknown TextureLoader variant `" + tag + "`");
}
geometry::Rect rect_of(uint32_t kind, const JValue& p) {
    return {kind, p.at("d1_min").n, p.at("d1_max").n, p.at("d2_min").n, p.at("d2_max").n, p.at("offset").n};
}

void num(std::ostringstream& os, double d) { char b[40]; snprintf(b, sizeof b, "%.17g", d); std::string s = b; if (s.find_first_of(".eEn") == std::string::npos) s += ".0"; os << s; }
void vec(std::ostringstream& os, const core::Vec3& v) { os << "{\"vec\":["; num(os, v.x); os << ","; num(os, v.y); os << ","; num(os, v.z); os << "]}"; }
void str(std::ostringstream& os, const std::string& s) { os << '"'; for (char c : s) { if (c == '"' || c == '\\') os << '\\' << c; else if (c == '\n') os << "\\n"; else os << c; } os << '"'; }
void tex(std::ostringstream& os, const TextureLoader& t) {
    switch (t.kind) {
        case TextureLoader::Solid: os << "{\"Solid\":"; vec(os, t.color.v); os << "}"; break;
        case TextureLoader::ImagePath: os << "{\"ImagePath\":"; str(os, t.path); os << "}"; break;
        case TextureLoader::Perlin: os << "{\"Perlin\":"; num(os, t.scalar); os << "}"; break;
        case TextureLoader::EarthBuiltin: os << "\"EarthBuiltin\""; break;
        default: os << "{\"Checker\":{\"size\":"; num(os, t.scalar); os << ",\"odd\":"; tex(os, *t.odd); os << ",\"even\":"; tex(os, *t.even); os << "}}";
    }
}
void rect(std::ostringstream& os, const geometry::Rect& r) {
    os << "{\"d1_min\":"; num(os, r.d1_min); os << ",\"d1_max\":"; num(os, r.d1_max); os << ",\"d2_min\":"; num(os, r.d2_min); os << ",\"d2_max\":"; num(os, r.d2_max); os << ",\"offset\":"; num(os, r.offset); os << "}";
}
}  // namespace

namespace scene {

SceneBuilder SceneBuilder::from_json(const std::string& text) {
    Parser ps{text.data(), text.data() + text.size()};
    JValue root = ps.value();
    if (root.t != JValue::Obj) throw Error("scene json: top level must be an object");
    SceneBuilder sb;
    {
        std::string tag; const JValue* p;
        variant_of(root.at("skybox"), &tag, &p);
        if (tag == "Above") sb.skybox = skybox::SkyBox::above();
        else if (tag == "None") sb.skybox = skybox::SkyBox::none();
        else if (tag == "Flat") sb.skybox = skybox::SkyBox::flat(core::Color(vec3_of(*p)));
        else throw Error("scene json: unknown SkyBox variant `" + tag + "`");
    }
    const JValue& objs = root.at("objects");
    if (objs.t != JValue::Arr) throw Error("scene json: `objects` must be an array");
    for (const JValue& o : objs.a) {
        std::string gt, mt; const JValue *gp, *mp;
        variant_of(o.at("geometry"), &gt, &gp);
        variant_of(o.at("material"), &mt, &mp);
        geometry::GeometricObject g;
        if (gt == "Sphere") g = geometry::Sphere{core::Point(vec3_of(gp->at("center"))), gp->at("radius").n};
        else if (gt == "RectXY") g = rect_of(B200RT_PRIM_RECT_XY, *gp);
        else if (gt == "RectYZ") g = rect_of(B200RT_PRIM_RECT_YZ, *gp);
        else if (gt == "RectXZ") g = rect_of(B200RT_PRIM_RECT_XZ, *gp);
        else if (gt == "RectBox") g = geometry::RectBox(core::Point(vec3_of(gp->at("min"))), core::Point(vec3_of(gp->at("max"))));
        else throw Error("scene json: unknown GeometricObject variant `" + gt + "`");
        material::MaterialType m;
        if (mt == "Metal") { material::Metal x; x.albedo = core::Color(vec3_of(mp->at("albedo"))); x.fuzz = mp->at("fuzz").n; m = x; }   // deserialised verbatim (no clamp), like serde
        else if (mt == "Dielectric") m = material::Dielectric{mp->at("ir").n};
        else if (mt == "Lambertian") m = material::Lambertian(texture_of(mp->at("albedo")));
        else if (mt == "DiffuseLight") m = material::DiffuseLight(texture_of(mp->at("albedo")));
        else if (mt == "FairyLight") m = material::FairyLight(texture_of(mp->at("albedo")));
        else throw Error("scene json: unknown MaterialType variant `" + mt + "`");
        sb.objects.push_back({std::move(g), std::move(m)});
    }
    return sb;
}

std::string SceneBuilder::to_json() const {
    std::ostringstream os;
    os << "{\"skybox\":";
    if (skybox.kind == skybox::SkyBox::Above) os << "\"Above\""; else if (skybox.kind == skybox::SkyBox::None) os << "\"None\""; else { os << "{\"Flat\":"; vec(os, skybox.color.v); os << "}"; }
    os << ",\"objects\":[";
    bool first = true;
    for (const SceneLoadObject& o : objects) {
        if (!first) os << ",";
        first = false;
        os << "\n{\"geometry\":";
        if (auto s = std::get_if<geometry::Sphere>(&o.geometry)) { os << "{\"Sphere\":{\"center\":"; vec(os, s->center.v); os << ",\"radius\":"; num(os, s->radius); os << "}}"; }
        else if (auto r = std::get_if<geometry::Rect>(&o.geometry)) {
            os << (r->kind == B200RT_PRIM_RECT_XY ? "{\"RectXY\":" : (r->kind == B200RT_PRIM_RECT_YZ ? "{\"RectYZ\":" : "{\"RectXZ\":")); rect(os, *r); os << "}";
        } else {
            const geometry::RectBox& b = std::get<geometry::RectBox>(o.geometry);
            const core::Vec3 &p0 = b.min.v, &p1 = b.max.v;
            os << "{\"RectBox\":{\"min\":"; vec(os, p0); os << ",\"max\":"; vec(os, p1);
            os << ",\"xy_sides\":["; rect(os, geometry::xy_rect(p0.x, p1.x, p0.y, p1.y, p1.z)); os << ","; rect(os, geometry::xy_rect(p0.x, p1.x, p0.y, p1.y, p0.z));
            os << "],\"yz_sides\":["; rect(os, geometry::yz_rect(p0.y, p1.y, p0.z, p1.z, p1.x)); os << ","; rect(os, geometry::yz_rect(p0.y, p1.y, p0.z, p1.z, p0.x));
            os << "],\"xz_sides\":["; rect(os, geometry::xz_rect(p0.x, p1.x, p0.z, p1.z, p1.y)); os << ","; rect(os, geometry::xz_rect(p0.x, p1.x, p0.z, p1.z, p0.y));
            os << "]}}";
        }
        os << ",\"material\":";
        if (auto x = std::get_if<material::Metal>(&o.material)) { os << "{\"Metal\":{\"albedo\":"; vec(os, x->albedo.v); os << ",\"fuzz\":"; num(os, x->fuzz); os << "}}"; }
        else if (auto x2 = std::get_if<material::Dielectric>(&o.material)) { os << "{\"Dielectric\":{\"ir\":"; num(os, x2->ir); os << "}}"; }
        else if (auto x3 = std::get_if<material::Lambertian>(&o.material)) { os << "{\"Lambertian\":{\"albedo\":"; tex(os, x3->albedo); os << "}}"; }
        else if (auto x4 = std::get_if<material::DiffuseLight>(&o.material)) { os << "{\"DiffuseLight\":{\"albedo\":"; tex(os, x4->albedo); os << "}}"; }
        else { os << "{\"FairyLight\":{\"albedo\":"; tex(os, std::get<material::FairyLight>(o.material).albedo); os << "}}"; }
        os << "}";
    }
    os << "\n]}\n";
    return os.str();
}

}  // namespace scene
}  // namespace raytracer
