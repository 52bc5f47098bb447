// swr_gltf.hpp — glTF 2.0 / GLB -> swr_scene_desc, the step in front of the hot path.
//
// Mirrors what the reference's loader produces for the fields the renderer consumes:
//   Scene::from_gltf                 src/scene.rs:145-354   (meshes, materials, nodes flattened, node spheres, bounds)
//   Node::from_gltf / local xform    src/scene.rs:383-419
//   Primitive::from_gltf             src/scene.rs:440-502   (positions w=1, u32 indices, texcoords or zeros, auto normals / tangents)
//   compute_bounding_sphere          src/scene.rs:504-518
//   compute_smooth_normals           src/scene.rs:520-553
//   compute_tangents                 src/scene.rs:555-646
//   get_texture_and_sampler          src/scene.rs:648-708   (URI images only; sampler wrap modes, default Repeat)
//   Material::from_gltf              src/scene.rs:710-787
//   SceneCamera::from_gltf           src/scene.rs:789-818
//   TextureCache::load_texture       src/texture.rs:897-1010 (RGBA8, R in the MSB) + Texture::generate_mipmaps :45-128
// The reference parses with the `gltf` crate (1.4, Cargo.lock) and decodes images with `image` 0.25; neither exists here, so
// the container / accessor rules below follow the glTF 2.0 specification (what that crate implements): GLB chunks, data:
// URIs, bufferView strides, sparse accessors, normalised integer texcoords (u8 / 255, u16 / 65535), u8/u16/u32 indices.
// Images: PNG (zlib; 16-bit samples reduced as image::to_rgba8 does) and JPEG (swr_jpeg.hpp) are decoded here; anything else
// can be handed in decoded through register_image().
// The environment (sky cubemap, prefiltered cubemap, BRDF LUT, voxel grid) is not part of a glTF file: the caller supplies
// it (the reference bakes it from assets/cubemap.jpg).
// Errors are std::runtime_error with the reference's SceneError wording ("Missing data: No positions in primitive", ...).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>

#include "swr_host.hpp"
#include "swr_jpeg.hpp"

namespace swr {
namespace gltf {

// ------------------------------------------------------------------------------------------------------------------
// JSON (RFC 8259) — just enough for glTF: objects, arrays, strings with escapes, numbers, true/false/null
// ------------------------------------------------------------------------------------------------------------------
struct JVal {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal *get(const char *key) const {
        if (type != Obj) return nullptr;
        for (const auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char *key) const { return get(key) != nullptr; }
    double number(const char *key, double dflt) const {
        const JVal *v = get(key);
        return v && v->type == Num ? v->num : dflt;
    }
    static int64_t to_i64(double d) {  // saturating, NaN -> 0 (a plain cast of an out-of-range double is undefined)
        if (!(d == d)) return 0;
        if (d >= 9.0e18) return INT64_MAX;
        if (d <= -9.0e18) return INT64_MIN;
        return (int64_t)d;
    }
    int64_t integer(const char *key, int64_t dflt) const {
        const JVal *v = get(key);
        return v && v->type == Num ? to_i64(v->num) : dflt;
    }
    int64_t as_index() const { return type == Num ? to_i64(num) : -1; }
    std::string string(const char *key, const std::string &dflt = "") const {
        const JVal *v = get(key);
        return v && v->type == Str ? v->str : dflt;
    }
    size_t size() const { return type == Arr ? arr.size() : 0; }
    const JVal &operator[](size_t i) const { return arr[i]; }
};

class JsonParser {
   public:
    JsonParser(const char *p, size_t n) : p_(p), end_(p + n) {}
    JVal parse() {
        JVal v = value(0);
        ws();
        if (p_ != end_) fail("trailing characters after the JSON document");
        return v;
    }

   private:
    const char *p_, *end_;
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("Invalid data: JSON: ") + what); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++;
    }
    bool lit(const char *s) {
        size_t n = std::strlen(s);
        if ((size_t)(end_ - p_) >= n && std::memcmp(p_, s, n) == 0) {
            p_ += n;
            return true;
        }
        return false;
    }
    static void utf8(std::string &o, uint32_t cp) {
        if (cp < 0x80)
            o += (char)cp;
        else if (cp < 0x800) {
            o += (char)(0xC0 | (cp >> 6));
            o += (char)(0x80 | (cp & 0x3F));
        } else if (cp < 0x10000) {
            o += (char)(0xE0 | (cp >> 12));
            o += (char)(0x80 | ((cp >> 6) & 0x3F));
            o += (char)(0x80 | (cp & 0x3F));
        } else {
            o += (char)(0xF0 | (cp >> 18));
            o += (char)(0x80 | ((cp >> 12) & 0x3F));
            o += (char)(0x80 | ((cp >> 6) & 0x3F));
            o += (char)(0x80 | (cp & 0x3F));
        }
    }
    uint32_t hex4() {
        if (end_ - p_ < 4) fail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9')
                v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f')
                v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F')
                v |= (uint32_t)(c - 'A' + 10);
            else
                fail("bad \\u escape");
        }
        return v;
    }
    std::string string_() {
        if (p_ >= end_ || *p_ != '"') fail("expected a string");
        p_++;
        std::string o;
        while (true) {
            if (p_ >= end_) fail("unterminated string");
            char c = *p_++;
            if (c == '"') break;
            if ((unsigned char)c < 0x20) fail("control character in string");
            if (c != '\\') {
                o += c;
                continue;
            }
            if (p_ >= end_) fail("unterminated escape");
            char e = *p_++;
            switch (e) {
                case '"': o += '"'; break;
                case '\\': o += '\\'; break;
                case '/': o += '/'; break;
                case 'b': o += '\b'; break;
                case 'f': o += '\f'; break;
                case 'n': o += '\n'; break;
                case 'r': o += '\r'; break;
                case 't': o += '\t'; break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        uint32_t lo = hex4();
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(o, cp);
                    break;
                }
                default: fail("unknown escape");
            }
        }
        return o;
    }
    JVal value(int depth) {
        if (depth > 256) fail("nesting too deep");
        ws();
        if (p_ >= end_) fail("unexpected end of input");
        JVal v;
        char c = *p_;
        if (c == '{') {
            p_++;
            v.type = JVal::Obj;
            ws();
            if (p_ < end_ && *p_ == '}') {
                p_++;
                return v;
            }
            while (true) {
                ws();
                std::string k = string_();
                ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                p_++;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p_ < end_ && *p_ == ',') {
                    p_++;
                    continue;
                }
                if (p_ < end_ && *p_ == '}') {
                    p_++;
                    break;
                }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            p_++;
            v.type = JVal::Arr;
            ws();
            if (p_ < end_ && *p_ == ']') {
                p_++;
                return v;
            }
            while (true) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p_ < end_ && *p_ == ',') {
                    p_++;
                    continue;
                }
                if (p_ < end_ && *p_ == ']') {
                    p_++;
                    break;
                }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = JVal::Str;
            v.str = string_();
        } else if (lit("true")) {
            v.type = JVal::Bool;
            v.b = true;
        } else if (lit("false")) {
            v.type = JVal::Bool;
        } else if (lit("null")) {
        } else {
            const char *s = p_;
            if (p_ < end_ && *p_ == '-') p_++;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail("unexpected character");
            while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) p_++;
            v.type = JVal::Num;
            v.num = std::strtod(std::string(s, p_).c_str(), nullptr);  // correctly rounded, like serde_json's float parsing
        }
        return v;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// bytes: files, base64 data URIs, percent-decoding
// ------------------------------------------------------------------------------------------------------------------
inline std::vector<uint8_t> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Missing data: cannot open '" + path + "'");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline std::vector<uint8_t> base64_decode(const char *s, size_t n) {
    std::vector<uint8_t> out;
    out.reserve(n * 3 / 4);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; i++) {
        char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z')
            v = c - 'A';
        else if (c >= 'a' && c <= 'z')
            v = c - 'a' + 26;
        else if (c >= '0' && c <= '9')
            v = c - '0' + 52;
        else if (c == '+' || c == '-')
            v = 62;
        else if (c == '/' || c == '_')
            v = 63;
        else if (c == '=' || c == '\n' || c == '\r')
            continue;
        else
            throw std::runtime_error("Invalid data: bad base64 character in data URI");
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return out;
}
inline std::string percent_decode(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '%' && i + 2 < s.size() && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
            o += (char)std::strtol(s.substr(i + 1, 2).c_str(), nullptr, 16);
            i += 2;
        } else {
            o += s[i];
        }
    }
    return o;
}
inline std::string dir_of(const std::string &path) {
    size_t k = path.find_last_of("/\\");
    return k == std::string::npos ? std::string(".") : path.substr(0, k);
}

// ------------------------------------------------------------------------------------------------------------------
// PNG (every colour type and bit depth of the specification, tRNS, non-interlaced and Adam7) -> RGBA8 bytes
// ------------------------------------------------------------------------------------------------------------------
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;  // width * height * 4, as image::DynamicImage::to_rgba8
};

inline Image decode_png(const std::vector<uint8_t> &file, const std::string &name) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto bad = [&](const char *why) -> std::runtime_error { return std::runtime_error("Missing data: Could not load texture '" + name + "': " + why); };
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) throw bad("neither a PNG nor a JPEG file (register other formats decoded)");
    auto be32 = [&](size_t o) { return ((uint32_t)file[o] << 24) | ((uint32_t)file[o + 1] << 16) | ((uint32_t)file[o + 2] << 8) | file[o + 3]; };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t o = 8;
    while (o + 12 <= file.size()) {
        uint32_t len = be32(o);
        if (o + 12 + (size_t)len > file.size()) throw bad("truncated chunk");
        const uint8_t *ty = &file[o + 4], *d = &file[o + 8];
        if (!std::memcmp(ty, "IHDR", 4)) {
            if (len < 13) throw bad("bad IHDR");
            w = be32(o + 8);
            h = be32(o + 12);
            depth = d[8];
            ctype = d[9];
            interlace = d[12];
        } else if (!std::memcmp(ty, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!std::memcmp(ty, "tRNS", 4)) {
            trns.assign(d, d + len);
        } else if (!std::memcmp(ty, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(ty, "IEND", 4)) {
            break;
        }
        o += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || ctype < 0) throw bad("missing IHDR");
    // legal (colour type, bit depth) pairs of the PNG specification
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                          : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                       : (depth == 8 || depth == 16);
    if (!depth_ok) throw bad("illegal bit depth for the colour type");
    if (w > 32768 || h > 32768) throw bad("image larger than 32768 pixels on a side");
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw bad("unknown colour type");
    if (ctype == 3 && plte.empty()) throw bad("palette image without PLTE");
    // inflate
    struct Pass {
        uint32_t x0, y0, dx, dy;
    };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    std::vector<Pass> passes;
    if (interlace)
        passes.assign(adam7, adam7 + 7);
    else
        passes.push_back(Pass{0, 0, 1, 1});
    size_t need = 0;
    for (const Pass &p : passes) {
        uint32_t pw = w > p.x0 ? (w - p.x0 + p.dx - 1) / p.dx : 0, ph = h > p.y0 ? (h - p.y0 + p.dy - 1) / p.dy : 0;
        if (pw && ph) need += (size_t)ph * (((size_t)pw * ch * depth + 7) / 8 + 1);
    }
    // deflate cannot expand by more than ~1032:1: a header that asks for more than the data can hold is corrupt
    if (need > idat.size() * 1032 + 65536) throw bad("corrupt image data (header larger than the compressed stream)");
    std::vector<uint8_t> raw(need);
    uLongf got = (uLongf)need;
    int zr = uncompress(raw.data(), &got, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || got != need) throw bad("corrupt image data (inflate)");
    Image img;
    img.width = w;
    img.height = h;
    img.rgba.assign((size_t)w * h * 4, 255);
    size_t pos = 0;
    std::vector<uint8_t> prev, cur;
    for (const Pass &p : passes) {
        uint32_t pw = w > p.x0 ? (w - p.x0 + p.dx - 1) / p.dx : 0, ph = h > p.y0 ? (h - p.y0 + p.dy - 1) / p.dy : 0;
        if (!pw || !ph) continue;
        const size_t stride = ((size_t)pw * ch * depth + 7) / 8;  // bytes per scanline
        const size_t bpp = std::max<size_t>(1, (size_t)ch * depth / 8);  // filter distance in bytes
        prev.assign(stride, 0);
        cur.resize(stride);
        // sample k of the scanline as 8 bits: sub-byte grey is scaled to 0..255, 16 bits are reduced as image::to_rgba8
        // does ((v + 128) / 257, the rounded v * 255 / 65535); palette indices stay indices
        auto sample = [&](size_t k) -> uint32_t {
            if (depth == 8) return cur[k];
            if (depth == 16) return ((uint32_t)cur[2 * k] << 8) | cur[2 * k + 1];
            const size_t bit = k * (size_t)depth;
            return (uint32_t)(cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
        };
        auto to8 = [&](uint32_t v) -> uint8_t {
            if (depth == 8) return (uint8_t)v;
            if (depth == 16) return (uint8_t)((v + 128u) / 257u);
            return (uint8_t)(v * (255u / ((1u << depth) - 1u)));
        };
        auto trns16 = [&](size_t i) -> uint32_t { return ((uint32_t)trns[2 * i] << 8) | trns[2 * i + 1]; };
        for (uint32_t y = 0; y < ph; y++) {
            const uint8_t ft = raw[pos++];
            const uint8_t *src = &raw[pos];
            pos += stride;
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int v = src[i];
                switch (ft) {
                    case 0: break;
                    case 1: v += a; break;
                    case 2: v += b; break;
                    case 3: v += (a + b) >> 1; break;
                    case 4: {
                        const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                        v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        break;
                    }
                    default: throw bad("bad filter type");
                }
                cur[i] = (uint8_t)v;
            }
            for (uint32_t x = 0; x < pw; x++) {
                uint8_t *dst = &img.rgba[(((size_t)p.y0 + (size_t)y * p.dy) * w + p.x0 + (size_t)x * p.dx) * 4];
                const size_t k0 = (size_t)x * ch;
                switch (ctype) {
                    case 0: {
                        const uint32_t g = sample(k0);
                        dst[0] = dst[1] = dst[2] = to8(g);
                        if (trns.size() >= 2 && trns16(0) == g) dst[3] = 0;  // tRNS holds the transparent grey level in file precision
                        break;
                    }
                    case 2: {
                        const uint32_t r = sample(k0), g = sample(k0 + 1), bl = sample(k0 + 2);
                        dst[0] = to8(r), dst[1] = to8(g), dst[2] = to8(bl);
                        if (trns.size() >= 6 && trns16(0) == r && trns16(1) == g && trns16(2) == bl) dst[3] = 0;
                        break;
                    }
                    case 3: {
                        const size_t k = sample(k0);
                        if (k * 3 + 2 >= plte.size()) throw bad("palette index out of range");
                        dst[0] = plte[k * 3], dst[1] = plte[k * 3 + 1], dst[2] = plte[k * 3 + 2];
                        if (k < trns.size()) dst[3] = trns[k];
                        break;
                    }
                    case 4: dst[0] = dst[1] = dst[2] = to8(sample(k0)), dst[3] = to8(sample(k0 + 1)); break;
                    case 6: dst[0] = to8(sample(k0)), dst[1] = to8(sample(k0 + 1)), dst[2] = to8(sample(k0 + 2)), dst[3] = to8(sample(k0 + 3)); break;
                }
            }
            std::swap(prev, cur);
        }
    }
    return img;
}

// PNG or JPEG by signature (what image::open does for the two formats glTF allows)
inline Image decode_image(const std::vector<uint8_t> &file, const std::string &name) {
    if (file.size() >= 2 && file[0] == 0xFF && file[1] == 0xD8) {
        jpeg::DecodedImage j = jpeg::decode(file.data(), file.size(), name);
        Image img;
        img.width = j.width, img.height = j.height;
        img.rgba.swap(j.rgba);
        return img;
    }
    return decode_png(file, name);
}

// ------------------------------------------------------------------------------------------------------------------
// Textures: texture.rs:897-1010 (pack) and :45-128 (type-aware mip chain). util.rs:50-100 colour helpers.
// ------------------------------------------------------------------------------------------------------------------
inline float srgb_to_linear_scalar(float s) { return s <= 0.04045f ? s / 12.92f : std::pow((s + 0.055f) / 1.055f, 2.4f); }     // util.rs:50-56
inline float linear_to_srgb_scalar(float s) { return s <= 0.0031308f ? s * 12.92f : std::pow(s, 1.0f / 2.4f) * 1.055f - 0.055f; }  // util.rs:67-73
inline uint32_t f32_as_u32_saturating(float v) {  // Rust `as u32`: NaN -> 0, saturating
    if (!(v == v) || v <= 0.0f) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline uint32_t rgba8_pack_vec4(const float c[4]) {  // util.rs:91-96 (no clamp: each channel truncates and is OR-ed in)
    return (f32_as_u32_saturating(c[0] * 255.0f) << 24) | (f32_as_u32_saturating(c[1] * 255.0f) << 16) | (f32_as_u32_saturating(c[2] * 255.0f) << 8) |
           f32_as_u32_saturating(c[3] * 255.0f);
}
inline void rgba8_unpack_vec4(uint32_t p, float c[4]) {  // util.rs:83-89
    c[0] = (float)((p >> 24) & 0xFF) / 255.0f;
    c[1] = (float)((p >> 16) & 0xFF) / 255.0f;
    c[2] = (float)((p >> 8) & 0xFF) / 255.0f;
    c[3] = (float)(p & 0xFF) / 255.0f;
}

struct TextureData {
    std::vector<uint32_t> data, mip_offsets, mip_widths, mip_heights, array_stride;
    uint32_t width = 0, height = 0, type = SWR_TEX_SRGB;
    uint32_t max_mip_level() const { return (uint32_t)mip_offsets.size() - 1; }

    // TextureCache::load_texture, 2D branch (texture.rs:961-993)
    static TextureData from_rgba8(const Image &img, uint32_t type) {
        TextureData t;
        t.width = img.width;
        t.height = img.height;
        t.type = type;
        t.data.resize((size_t)img.width * img.height);
        for (size_t i = 0; i < t.data.size(); i++)
            t.data[i] = ((uint32_t)img.rgba[4 * i] << 24) | ((uint32_t)img.rgba[4 * i + 1] << 16) | ((uint32_t)img.rgba[4 * i + 2] << 8) | img.rgba[4 * i + 3];
        t.mip_offsets = {0};
        t.mip_widths = {img.width};
        t.mip_heights = {img.height};
        t.array_stride = {0};
        return t;
    }

    // Texture::generate_mipmaps (texture.rs:45-128); array_size 6 for cubemaps
    void generate_mipmaps() {
        if (max_mip_level() != 0) throw std::runtime_error("Texture already has mipmaps");
        const uint32_t num_mips = 1 + ilog2(std::max(width, height));
        const uint32_t array_size = type == SWR_TEX_CUBEMAP ? 6 : 1;
        for (uint32_t mip = 1; mip < num_mips; mip++) {
            mip_offsets.push_back((uint32_t)data.size());
            const uint32_t mw = std::max(width >> mip, 1u), mh = std::max(height >> mip, 1u);
            array_stride.push_back(mw * mh);
            const uint32_t poff = mip_offsets[mip - 1], pw = mip_widths[mip - 1], ph = mip_heights[mip - 1], pstride = array_stride[mip - 1];
            for (uint32_t slice = 0; slice < array_size; slice++) {
                const uint32_t so = poff + slice * pstride;
                for (uint32_t y = 0; y < mh; y++)
                    for (uint32_t x = 0; x < mw; x++) {
                        const uint32_t x0 = x * 2, y0 = y * 2, x1 = std::min(x0 + 1, pw - 1), y1 = std::min(y0 + 1, ph - 1);
                        float p00[4], p10[4], p01[4], p11[4], avg[4];
                        rgba8_unpack_vec4(data[so + y0 * pw + x0], p00);
                        rgba8_unpack_vec4(data[so + y0 * pw + x1], p10);
                        rgba8_unpack_vec4(data[so + y1 * pw + x0], p01);
                        rgba8_unpack_vec4(data[so + y1 * pw + x1], p11);
                        if (type == SWR_TEX_SRGB) {
                            for (int c = 0; c < 4; c++) {
                                const bool col = c < 3;
                                const float a = col ? srgb_to_linear_scalar(p00[c]) : p00[c], b = col ? srgb_to_linear_scalar(p10[c]) : p10[c];
                                const float cc = col ? srgb_to_linear_scalar(p01[c]) : p01[c], d = col ? srgb_to_linear_scalar(p11[c]) : p11[c];
                                const float m = (((a + b) + cc) + d) / 4.0f;
                                avg[c] = col ? linear_to_srgb_scalar(m) : m;
                            }
                        } else if (type == SWR_TEX_METALLIC_ROUGHNESS) {
                            const float rough = (((p00[1] * p00[1] + p10[1] * p10[1]) + p01[1] * p01[1]) + p11[1] * p11[1]) / 4.0f;
                            const float metal = (((p00[2] + p10[2]) + p01[2]) + p11[2]) / 4.0f;
                            avg[0] = p00[0];
                            avg[1] = std::sqrt(rough);
                            avg[2] = metal;
                            avg[3] = p00[3];
                        } else if (type == SWR_TEX_NORMAL) {
                            for (int c = 0; c < 4; c++) {
                                const float m = ((((p00[c] * 2.0f - 1.0f) + (p10[c] * 2.0f - 1.0f)) + (p01[c] * 2.0f - 1.0f)) + (p11[c] * 2.0f - 1.0f)) / 4.0f;
                                avg[c] = (m + 1.0f) / 2.0f;
                            }
                        } else {
                            for (int c = 0; c < 4; c++) avg[c] = (((p00[c] + p10[c]) + p01[c]) + p11[c]) / 4.0f;
                        }
                        data.push_back(rgba8_pack_vec4(avg));
                    }
            }
            mip_widths.push_back(mw);
            mip_heights.push_back(mh);
        }
    }

   private:
    static uint32_t ilog2(uint32_t v) {
        uint32_t r = 0;
        while (v >>= 1) r++;
        return r;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// geometry helpers (glam semantics: unfused f32, Vec3A dot = (x*x' + y*y') + z*z')
// ------------------------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) {  // glam: self * length_recip()
    const float r = 1.0f / length(a);
    return a * r;
}

struct PrimitiveData {
    std::vector<float> positions, normals, tangents, texcoords;  // 4, 4, 4, 2 floats per vertex (ABI layout)
    std::vector<uint32_t> indices;
    uint32_t material_index = 0;
    float bounding_sphere[4] = {0, 0, 0, 0};
    uint32_t nverts() const { return (uint32_t)(positions.size() / 4); }
};

// scene.rs:504-518
inline void compute_bounding_sphere(const std::vector<float> &pos4, float out[4]) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i + 3 < pos4.size(); i += 4)
        for (int c = 0; c < 3; c++) {
            mn[c] = std::fmin(mn[c], pos4[i + c]);
            mx[c] = std::fmax(mx[c], pos4[i + c]);
        }
    V3 d{mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    for (int c = 0; c < 3; c++) out[c] = (mn[c] + mx[c]) * 0.5f;  // Vec3A::midpoint
    out[3] = length(d) / 2.0f;
}

// scene.rs:520-553
inline std::vector<float> compute_smooth_normals(const std::vector<float> &pos4, const std::vector<uint32_t> &idx) {
    const size_t n = pos4.size() / 4;
    std::vector<V3> acc(n, V3{0, 0, 0});
    auto P = [&](uint32_t i) { return V3{pos4[4 * (size_t)i], pos4[4 * (size_t)i + 1], pos4[4 * (size_t)i + 2]}; };
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        const V3 v0 = P(idx[t]), e1 = P(idx[t + 1]) - v0, e2 = P(idx[t + 2]) - v0;
        const V3 fn = cross(e1, e2);
        for (int k = 0; k < 3; k++) acc[idx[t + k]] = acc[idx[t + k]] + fn;
    }
    std::vector<float> out(n * 4, 0.0f);
    for (size_t i = 0; i < n; i++) {
        V3 v = dot(acc[i], acc[i]) > 0.0f ? normalize(acc[i]) : V3{0.0f, 0.0f, 1.0f};
        out[4 * i] = v.x, out[4 * i + 1] = v.y, out[4 * i + 2] = v.z;
    }
    return out;
}

// scene.rs:555-646 (glam Vec3: plain scalar ops; dot = x*x' + y*y' + z*z' left to right)
inline std::vector<float> compute_tangents(const std::vector<float> &pos4, const std::vector<float> &uv2, const std::vector<float> &nrm4, const std::vector<uint32_t> &idx) {
    const size_t n = pos4.size() / 4;
    std::vector<V3> tan(n, V3{0, 0, 0}), bit(n, V3{0, 0, 0});
    auto P = [&](size_t i) { return V3{pos4[4 * i], pos4[4 * i + 1], pos4[4 * i + 2]}; };
    for (size_t t = 0; t + 2 < idx.size(); t += 3) {
        const size_t i0 = idx[t], i1 = idx[t + 1], i2 = idx[t + 2];
        const V3 e1 = P(i1) - P(i0), e2 = P(i2) - P(i0);
        const float du1 = uv2[2 * i1] - uv2[2 * i0], dv1 = uv2[2 * i1 + 1] - uv2[2 * i0 + 1], du2 = uv2[2 * i2] - uv2[2 * i0], dv2 = uv2[2 * i2 + 1] - uv2[2 * i0 + 1];
        const float det = du1 * dv2 - du2 * dv1;
        if (std::fabs(det) < 1e-6f) continue;
        const V3 tg = (e1 * dv2 - e2 * dv1) / det, bt = (e2 * du1 - e1 * du2) / det;
        for (size_t k : {i0, i1, i2}) {
            tan[k] = tan[k] + tg;
            bit[k] = bit[k] + bt;
        }
    }
    std::vector<float> out(n * 4, 0.0f);
    for (size_t i = 0; i < n; i++) {
        const V3 nn{nrm4[4 * i], nrm4[4 * i + 1], nrm4[4 * i + 2]}, t = tan[i];
        if (dot(t, t) < 1e-6f) {
            const V3 helper = std::fabs(nn.x) > 0.9f ? V3{0.0f, 1.0f, 0.0f} : V3{1.0f, 0.0f, 0.0f};
            const V3 tg = normalize(cross(nn, helper));
            out[4 * i] = tg.x, out[4 * i + 1] = tg.y, out[4 * i + 2] = tg.z, out[4 * i + 3] = 1.0f;
            continue;
        }
        const V3 tg = normalize(t - nn * dot(nn, t));
        const float hand = dot(cross(nn, tg), bit[i]) < 0.0f ? -1.0f : 1.0f;
        out[4 * i] = tg.x, out[4 * i + 1] = tg.y, out[4 * i + 2] = tg.z, out[4 * i + 3] = hand;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------------------------
// the document
// ------------------------------------------------------------------------------------------------------------------
struct CameraData {  // scene.rs:123-143, 789-818
    bool perspective = true;
    float fov_or_xmag = 0, aspect_or_ymag = 1, znear = 0, zfar = 100;
    float transform[16];  // world transform of the node that carries it (identity when none does)
};

struct Environment {  // supplied by the caller, copied
    const swr_texture_desc *cubemap = nullptr, *cubemap_specular = nullptr, *brdf_lut = nullptr;
    swr_voxel_grid_desc voxel_grid{};
    float light_direction[3] = {0, 1, 0}, light_color[3] = {5, 5, 4.75f};
};

class Document {
   public:
    std::vector<PrimitiveData> primitives;
    std::vector<swr_mesh_desc> meshes;
    std::vector<swr_node_desc> nodes;
    std::vector<swr_material_desc> materials;
    std::vector<std::shared_ptr<TextureData>> texture_data;  // per scene texture slot (shared between slots with equal URI)
    std::vector<std::pair<uint32_t, uint32_t>> texture_wrap;
    std::vector<std::string> texture_uri;
    std::vector<CameraData> cameras;
    float bounds_min[3], bounds_max[3], bounds_center[3], bounds_diagonal = 0;
    std::vector<float> voxels;

    // flat views (valid while the Document lives)
    std::vector<swr_primitive_desc> prim_descs;
    std::vector<swr_texture_desc> tex_descs;
    swr_scene_desc desc{};

    static std::map<std::string, Image> &registered_images() {
        static std::map<std::string, Image> m;
        return m;
    }
    static std::mutex &registered_images_mutex() {  // loads may run on any host thread (rayon workers in the reference)
        static std::mutex mx;
        return mx;
    }

    static std::unique_ptr<Document> load(const std::string &path, const Environment &env) {
        std::unique_ptr<Document> d(new Document());
        d->base_dir_ = dir_of(path);
        std::vector<uint8_t> file = read_file(path);
        std::vector<uint8_t> glb_bin;
        bool have_glb_bin = false;
        std::string json;
        if (file.size() >= 12 && std::memcmp(file.data(), "glTF", 4) == 0) {  // GLB container (spec 4.4)
            auto le32 = [&](size_t o) { return (uint32_t)file[o] | ((uint32_t)file[o + 1] << 8) | ((uint32_t)file[o + 2] << 16) | ((uint32_t)file[o + 3] << 24); };
            if (le32(4) != 2) throw std::runtime_error("Invalid data: unsupported GLB version");
            size_t o = 12;
            while (o + 8 <= file.size()) {
                const uint32_t len = le32(o), type = le32(o + 4);
                if (o + 8 + (size_t)len > file.size()) throw std::runtime_error("Invalid data: truncated GLB chunk");
                if (type == 0x4E4F534Au)
                    json.assign((const char *)&file[o + 8], len);
                else if (type == 0x004E4942u && !have_glb_bin) {
                    glb_bin.assign(&file[o + 8], &file[o + 8] + len);
                    have_glb_bin = true;
                }
                o += 8 + (size_t)((len + 3u) & ~3u);
            }
            if (json.empty()) throw std::runtime_error("Invalid data: GLB without a JSON chunk");
        } else {
            json.assign((const char *)file.data(), file.size());
        }
        d->root_ = JsonParser(json.data(), json.size()).parse();
        if (d->root_.type != JVal::Obj) throw std::runtime_error("Invalid data: glTF root is not an object");
        d->load_buffers(have_glb_bin ? &glb_bin : nullptr);
        d->build(env);
        return d;
    }

   private:
    JVal root_;
    std::string base_dir_;
    std::vector<std::vector<uint8_t>> buffers_;
    std::map<std::string, std::shared_ptr<TextureData>> texture_cache_;  // TextureCache: keyed by URI, first requested type wins
    std::map<std::string, int> slot_of_;                                 // (uri, wrap_s, wrap_t) -> scene texture slot

    const JVal &array(const char *key) const {
        static const JVal empty;
        const JVal *v = root_.get(key);
        return v && v->type == JVal::Arr ? *v : empty;
    }

    void load_buffers(const std::vector<uint8_t> *glb_bin) {
        const JVal &bs = array("buffers");
        for (size_t i = 0; i < bs.size(); i++) {
            const JVal &b = bs[i];
            const std::string uri = b.string("uri");
            if (uri.empty()) {
                if (i == 0 && glb_bin)
                    buffers_.push_back(*glb_bin);
                else
                    throw std::runtime_error("Missing data: buffer " + std::to_string(i) + " has no uri");
            } else if (uri.compare(0, 5, "data:") == 0) {
                size_t k = uri.find(";base64,");
                if (k == std::string::npos) throw std::runtime_error("Invalid data: data URI without base64 payload");
                buffers_.push_back(base64_decode(uri.data() + k + 8, uri.size() - k - 8));
            } else {
                buffers_.push_back(read_file(base_dir_ + "/" + percent_decode(uri)));
            }
            const int64_t want = b.integer("byteLength", -1);
            if (want >= 0 && buffers_.back().size() < (size_t)want) throw std::runtime_error("Invalid data: buffer " + std::to_string(i) + " is shorter than its byteLength");
        }
    }

    // ---- accessors -------------------------------------------------------------------------------------------
    struct View {
        const uint8_t *p = nullptr;
        size_t stride = 0, count = 0;
        int ctype = 0, ncomp = 0;
        bool normalized = false;
    };
    static int comp_size(int ctype) {
        switch (ctype) {
            case 5120: case 5121: return 1;
            case 5122: case 5123: return 2;
            case 5125: case 5126: return 4;
        }
        throw std::runtime_error("Invalid data: unknown accessor componentType " + std::to_string(ctype));
    }
    static int type_ncomp(const std::string &t) {
        if (t == "SCALAR") return 1;
        if (t == "VEC2") return 2;
        if (t == "VEC3") return 3;
        if (t == "VEC4") return 4;
        if (t == "MAT2") return 4;
        if (t == "MAT3") return 9;
        if (t == "MAT4") return 16;
        throw std::runtime_error("Invalid data: unknown accessor type '" + t + "'");
    }
    const uint8_t *view_bytes(int64_t view_index, size_t byte_offset, size_t need, size_t *stride_out) const {
        const JVal &views = array("bufferViews");
        if (view_index < 0 || (size_t)view_index >= views.size()) throw std::runtime_error("Invalid data: bufferView index out of range");
        const JVal &v = views[(size_t)view_index];
        const int64_t bi = v.integer("buffer", -1);
        if (bi < 0 || (size_t)bi >= buffers_.size()) throw std::runtime_error("Invalid data: buffer index out of range");
        const int64_t vo = v.integer("byteOffset", 0), vl = v.integer("byteLength", 0), st = v.integer("byteStride", 0);
        const size_t bsz = buffers_[(size_t)bi].size();
        if (vo < 0 || vl < 0 || st < 0 || st > 65536 || (uint64_t)vo > bsz || (uint64_t)vl > bsz - (uint64_t)vo)
            throw std::runtime_error("Invalid data: bufferView reaches outside its buffer");
        if (stride_out) *stride_out = (size_t)st;
        if (byte_offset > (size_t)vl || need > (size_t)vl - byte_offset) throw std::runtime_error("Invalid data: accessor reaches outside its bufferView");
        const size_t off = (size_t)vo + byte_offset;
        return buffers_[(size_t)bi].data() + off;
    }
    static double read_comp(const uint8_t *p, int ctype) {
        switch (ctype) {
            case 5120: return (double)*(const int8_t *)p;
            case 5121: return (double)*p;
            case 5122: { int16_t v; std::memcpy(&v, p, 2); return (double)v; }
            case 5123: { uint16_t v; std::memcpy(&v, p, 2); return (double)v; }
            case 5125: { uint32_t v; std::memcpy(&v, p, 4); return (double)v; }
            default: { float v; std::memcpy(&v, p, 4); return (double)v; }
        }
    }
    // Raw element-wise read of an accessor (sparse substitution applied): count * ncomp values as stored.
    void read_accessor(int64_t index, std::vector<double> &out, int &ctype, int &ncomp, bool &normalized) const {
        const JVal &accs = array("accessors");
        if (index < 0 || (size_t)index >= accs.size()) throw std::runtime_error("Invalid data: accessor index out of range");
        const JVal &a = accs[(size_t)index];
        ctype = (int)a.integer("componentType", 0);
        ncomp = type_ncomp(a.string("type"));
        const JVal *nv = a.get("normalized");
        normalized = nv && nv->type == JVal::Bool && nv->b;
        const int64_t count_i = a.integer("count", 0);
        if (count_i < 0 || count_i > (int64_t)1 << 28) throw std::runtime_error("Invalid data: accessor count out of range");
        const size_t count = (size_t)count_i, cs = (size_t)comp_size(ctype), es = cs * (size_t)ncomp;
        out.assign(count * (size_t)ncomp, 0.0);
        if (a.has("bufferView") && count) {
            size_t stride = 0;
            const size_t boff = (size_t)a.integer("byteOffset", 0);
            // probe the stride first, then check the whole extent
            view_bytes(a.integer("bufferView", -1), boff, es, &stride);
            if (stride == 0) stride = es;
            const uint8_t *p = view_bytes(a.integer("bufferView", -1), boff, stride * (count - 1) + es, nullptr);
            for (size_t i = 0; i < count; i++)
                for (int c = 0; c < ncomp; c++) out[i * (size_t)ncomp + (size_t)c] = read_comp(p + i * stride + (size_t)c * cs, ctype);
        }
        if (const JVal *sp = a.get("sparse")) {  // spec 3.6.2.3
            const int64_t n_i = sp->integer("count", 0);
            if (n_i < 0 || (uint64_t)n_i > count) throw std::runtime_error("Invalid data: sparse count out of range");
            const size_t n = (size_t)n_i;
            const JVal *si = sp->get("indices"), *sv = sp->get("values");
            if (!si || !sv) throw std::runtime_error("Invalid data: sparse accessor without indices/values");
            const int ict = (int)si->integer("componentType", 0);
            const size_t ics = (size_t)comp_size(ict);
            const uint8_t *ip = n ? view_bytes(si->integer("bufferView", -1), (size_t)si->integer("byteOffset", 0), ics * n, nullptr) : nullptr;
            const uint8_t *vp = n ? view_bytes(sv->integer("bufferView", -1), (size_t)sv->integer("byteOffset", 0), es * n, nullptr) : nullptr;
            for (size_t k = 0; k < n; k++) {
                const size_t at = (size_t)read_comp(ip + k * ics, ict);
                if (at >= count) throw std::runtime_error("Invalid data: sparse index out of range");
                for (int c = 0; c < ncomp; c++) out[at * (size_t)ncomp + (size_t)c] = read_comp(vp + k * es + (size_t)c * cs, ctype);
            }
        }
    }
    std::vector<float> read_f32(int64_t index, int want_ncomp, const char *what) const {
        std::vector<double> raw;
        int ct, nc;
        bool norm;
        read_accessor(index, raw, ct, nc, norm);
        if (ct != 5126 || nc != want_ncomp) throw std::runtime_error(std::string("Invalid data: ") + what + " must be float VEC" + std::to_string(want_ncomp));
        return std::vector<float>(raw.begin(), raw.end());
    }

    // ---- Primitive::from_gltf (scene.rs:440-502) ---------------------------------------------------------------
    PrimitiveData load_primitive(const JVal &p) const {
        PrimitiveData out;
        const JVal *attrs = p.get("attributes");
        if (!attrs || !attrs->has("POSITION")) throw std::runtime_error("Missing data: No positions in primitive");
        std::vector<float> pos3 = read_f32(attrs->integer("POSITION", -1), 3, "POSITION");
        const size_t n = pos3.size() / 3;
        out.positions.resize(n * 4);
        for (size_t i = 0; i < n; i++) {
            out.positions[4 * i] = pos3[3 * i], out.positions[4 * i + 1] = pos3[3 * i + 1], out.positions[4 * i + 2] = pos3[3 * i + 2];
            out.positions[4 * i + 3] = 1.0f;
        }
        if (!p.has("indices")) throw std::runtime_error("Missing data: No indices in primitive");
        {
            std::vector<double> raw;
            int ct, nc;
            bool norm;
            read_accessor(p.integer("indices", -1), raw, ct, nc, norm);
            if (nc != 1 || (ct != 5121 && ct != 5123 && ct != 5125)) throw std::runtime_error("Invalid data: indices must be unsigned SCALAR");
            out.indices.resize(raw.size());
            for (size_t i = 0; i < raw.size(); i++) {
                out.indices[i] = (uint32_t)raw[i];
                if (out.indices[i] >= n) throw std::runtime_error("Invalid data: vertex index out of range");
            }
        }
        out.texcoords.assign(n * 2, 0.0f);  // "dummy texcoords so the rasterizer doesn't complain" (scene.rs:463-466)
        if (attrs->has("TEXCOORD_0")) {
            std::vector<double> raw;
            int ct, nc;
            bool norm;
            read_accessor(attrs->integer("TEXCOORD_0", -1), raw, ct, nc, norm);
            if (nc != 2 || raw.size() != n * 2) throw std::runtime_error("Invalid data: TEXCOORD_0 must be VEC2 with one entry per vertex");
            for (size_t i = 0; i < raw.size(); i++) {  // gltf::mesh::util::ReadTexCoords::into_f32
                if (ct == 5126)
                    out.texcoords[i] = (float)raw[i];
                else if (ct == 5121)
                    out.texcoords[i] = (float)raw[i] / 255.0f;
                else if (ct == 5123)
                    out.texcoords[i] = (float)raw[i] / 65535.0f;
                else
                    throw std::runtime_error("Invalid data: TEXCOORD_0 component type");
            }
        }
        if (attrs->has("NORMAL")) {
            std::vector<float> n3 = read_f32(attrs->integer("NORMAL", -1), 3, "NORMAL");
            if (n3.size() != n * 3) throw std::runtime_error("Invalid data: NORMAL count differs from POSITION count");
            out.normals.assign(n * 4, 0.0f);
            for (size_t i = 0; i < n; i++) out.normals[4 * i] = n3[3 * i], out.normals[4 * i + 1] = n3[3 * i + 1], out.normals[4 * i + 2] = n3[3 * i + 2];
        } else {
            out.normals = compute_smooth_normals(out.positions, out.indices);
        }
        if (attrs->has("TANGENT")) {
            out.tangents = read_f32(attrs->integer("TANGENT", -1), 4, "TANGENT");
            if (out.tangents.size() != n * 4) throw std::runtime_error("Invalid data: TANGENT count differs from POSITION count");
        } else {
            out.tangents = compute_tangents(out.positions, out.texcoords, out.normals, out.indices);
        }
        compute_bounding_sphere(out.positions, out.bounding_sphere);
        out.material_index = (uint32_t)std::max<int64_t>(p.integer("material", 0), 0);  // unwrap_or(0), scene.rs:499
        return out;
    }

    // ---- textures (scene.rs:648-708, texture.rs:879-1010) -------------------------------------------------------
    int texture_slot(const JVal *info, uint32_t type) {
        if (!info) return -1;
        const JVal &texs = array("textures");
        const int64_t ti = info->integer("index", -1);
        if (ti < 0 || (size_t)ti >= texs.size()) return -1;
        const JVal &tex = texs[(size_t)ti];
        const JVal &imgs = array("images");
        const int64_t si = tex.integer("source", -1);
        if (si < 0 || (size_t)si >= imgs.size()) return -1;
        const std::string uri = imgs[(size_t)si].string("uri");
        if (uri.empty()) return -1;  // gltf::image::Source::View => None (scene.rs:706)
        uint32_t ws = SWR_WRAP_REPEAT, wt = SWR_WRAP_REPEAT;
        const int64_t sm = tex.integer("sampler", -1);
        const JVal &samplers = array("samplers");
        if (sm >= 0 && (size_t)sm < samplers.size()) {
            auto wrap = [](int64_t v) -> uint32_t { return v == 33071 ? SWR_WRAP_CLAMP_TO_EDGE : v == 33648 ? SWR_WRAP_MIRRORED_REPEAT : SWR_WRAP_REPEAT; };
            ws = wrap(samplers[(size_t)sm].integer("wrapS", 10497));
            wt = wrap(samplers[(size_t)sm].integer("wrapT", 10497));
        }
        std::shared_ptr<TextureData> td;
        auto it = texture_cache_.find(uri);
        if (it != texture_cache_.end()) {
            td = it->second;
        } else {
            Image img;
            bool registered = false;
            {
                std::lock_guard<std::mutex> lock(registered_images_mutex());
                auto reg = registered_images().find(uri);
                if (reg != registered_images().end()) {
                    img = reg->second;
                    registered = true;
                }
            }
            if (registered) {
            } else if (uri.compare(0, 5, "data:") == 0) {
                size_t k = uri.find(";base64,");
                if (k == std::string::npos) throw std::runtime_error("Invalid data: data URI without base64 payload");
                img = decode_image(base64_decode(uri.data() + k + 8, uri.size() - k - 8), "data:");
            } else {
                std::vector<uint8_t> bytes;
                try {
                    bytes = read_file(base_dir_ + "/" + percent_decode(uri));
                } catch (const std::exception &) {
                    throw std::runtime_error("Missing data: Could not load texture '" + uri + "' from path '" + base_dir_ + "/" + uri + "'");
                }
                img = decode_image(bytes, uri);
            }
            td = std::make_shared<TextureData>(TextureData::from_rgba8(img, type));
            td->generate_mipmaps();
            texture_cache_[uri] = td;
        }
        const std::string key = uri + "|" + std::to_string(ws) + "|" + std::to_string(wt);
        auto sl = slot_of_.find(key);
        if (sl != slot_of_.end()) return sl->second;
        const int slot = (int)texture_data.size();
        texture_data.push_back(td);
        texture_wrap.emplace_back(ws, wt);
        texture_uri.push_back(uri);
        slot_of_[key] = slot;
        return slot;
    }

    // ---- Material::from_gltf (scene.rs:710-787) ------------------------------------------------------------------
    swr_material_desc load_material(const JVal &m) {
        swr_material_desc d{};
        static const JVal empty_obj = [] {
            JVal v;
            v.type = JVal::Obj;
            return v;
        }();
        const JVal *pbr = m.get("pbrMetallicRoughness");
        if (!pbr) pbr = &empty_obj;
        const JVal *bc = pbr->get("baseColorFactor");
        for (int c = 0; c < 4; c++) d.base_color_factor[c] = bc && bc->size() == 4 ? (float)(*bc)[(size_t)c].num : 1.0f;
        d.base_color_texture = texture_slot(pbr->get("baseColorTexture"), SWR_TEX_SRGB);
        d.metallic_factor = (float)pbr->number("metallicFactor", 1.0);
        d.roughness_factor = (float)pbr->number("roughnessFactor", 1.0);
        d.metallic_roughness_texture = texture_slot(pbr->get("metallicRoughnessTexture"), SWR_TEX_METALLIC_ROUGHNESS);
        d.normal_texture = texture_slot(m.get("normalTexture"), SWR_TEX_NORMAL);
        const JVal *em = m.get("emissiveFactor");
        for (int c = 0; c < 3; c++) d.emissive_factor[c] = em && em->size() == 3 ? (float)(*em)[(size_t)c].num : 0.0f;
        d.emissive_texture = texture_slot(m.get("emissiveTexture"), SWR_TEX_SRGB);
        const JVal *occ = m.get("occlusionTexture");
        d.occlusion_texture = texture_slot(occ, SWR_TEX_LINEAR);
        d.occlusion_strength = occ ? (float)occ->number("strength", 1.0) : 1.0f;
        d.alpha_cutoff = (float)m.number("alphaCutoff", 0.5);
        const JVal *ext = m.get("extensions");
        const JVal *tr = ext ? ext->get("KHR_materials_transmission") : nullptr;
        d.transmission = tr ? (float)tr->number("transmissionFactor", 0.0) : 0.0f;
        d.transmission_texture = tr ? texture_slot(tr->get("transmissionTexture"), SWR_TEX_LINEAR) : -1;
        d.flags = 0;
        if (m.string("alphaMode", "OPAQUE") == "MASK" && d.base_color_texture >= 0) d.flags |= SWR_MAT_ALPHA_TESTED;  // scene.rs:716-717
        if (tr) d.flags |= SWR_MAT_TRANSLUCENT;                                                                        // scene.rs:719
        return d;
    }

    // ---- Node::get_local_transform (scene.rs:395-419) --------------------------------------------------------------
    static Mat4 local_transform(const JVal &n) {
        if (const JVal *mx = n.get("matrix")) {
            if (mx->size() == 16) {
                Mat4 r;
                for (int i = 0; i < 16; i++) r.m[i] = (float)(*mx)[(size_t)i].num;
                return r;
            }
        }
        float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        if (const JVal *v = n.get("translation"))
            for (size_t i = 0; i < 3 && i < v->size(); i++) t[i] = (float)(*v)[i].num;
        if (const JVal *v = n.get("rotation"))
            for (size_t i = 0; i < 4 && i < v->size(); i++) q[i] = (float)(*v)[i].num;
        if (const JVal *v = n.get("scale"))
            for (size_t i = 0; i < 3 && i < v->size(); i++) s[i] = (float)(*v)[i].num;
        Mat4 T = Mat4::identity(), S = Mat4::identity();
        T.m[12] = t[0], T.m[13] = t[1], T.m[14] = t[2];
        S.m[0] = s[0], S.m[5] = s[1], S.m[10] = s[2];
        return mul(mul(T, mat4_from_quat(Quat{q[0], q[1], q[2], q[3]})), S);  // translation * rotation * scale
    }

    void final_transforms(size_t node, const Mat4 &parent, const std::vector<Mat4> &local, std::vector<int> &guard, int depth = 0) {
        if (guard[node]++) throw std::runtime_error("Invalid data: node hierarchy is not a forest");
        if (depth > 2048) throw std::runtime_error("Invalid data: node hierarchy deeper than 2048 levels");
        const Mat4 fin = mul(parent, local[node]);  // scene.rs:365
        std::memcpy(nodes[node].transform, fin.m, 64);
        const JVal *ch = array("nodes")[node].get("children");
        for (size_t i = 0; ch && i < ch->size(); i++) {
            const int64_t ci = (*ch)[i].as_index();
            if (ci < 0 || (size_t)ci >= nodes.size()) throw std::runtime_error("Invalid data: child index out of range");
            const size_t c = (size_t)ci;
            final_transforms(c, fin, local, guard, depth + 1);
        }
    }

    // Mat4 * &BoundingSphere (scene.rs:53-63)
    static void transform_sphere(const float *m, const float *s, float out[4]) {
        auto len3 = [](const float *c) { return std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]); };
        const float max_scale = (len3(m) + len3(m + 4) + len3(m + 8)) / 3.0f;
        for (int r = 0; r < 3; r++) out[r] = ((m[r] * s[0] + m[4 + r] * s[1]) + m[8 + r] * s[2]) + m[12 + r];  // transform_point3a
        out[3] = s[3] * max_scale;
    }

    // ---- Scene::from_gltf (scene.rs:145-354) ---------------------------------------------------------------------
    void build(const Environment &env) {
        const JVal &jm = array("meshes");
        for (size_t i = 0; i < jm.size(); i++) {
            swr_mesh_desc md{(uint32_t)primitives.size(), 0};
            const JVal *ps = jm[i].get("primitives");
            for (size_t k = 0; ps && k < ps->size(); k++) primitives.push_back(load_primitive((*ps)[k]));
            md.num_primitives = (uint32_t)primitives.size() - md.first_primitive;
            meshes.push_back(md);
        }
        const JVal &jmat = array("materials");
        for (size_t i = 0; i < jmat.size(); i++) materials.push_back(load_material(jmat[i]));
        if (materials.empty() && !primitives.empty())
            throw std::runtime_error("Invalid data: no materials (the reference indexes materials[0] for such primitives and panics, scene.rs:499)");
        for (const PrimitiveData &p : primitives)
            if (p.material_index >= materials.size()) throw std::runtime_error("Invalid data: primitive material index out of range");
        const JVal &jc = array("cameras");
        for (size_t i = 0; i < jc.size(); i++) {
            CameraData c;
            std::memcpy(c.transform, Mat4::identity().m, 64);
            if (const JVal *p = jc[i].get("perspective")) {
                c.perspective = true;
                c.fov_or_xmag = (float)p->number("yfov", 0.0);
                c.aspect_or_ymag = (float)p->number("aspectRatio", 1.0);
                c.znear = (float)p->number("znear", 0.0);
                c.zfar = (float)p->number("zfar", 100.0);
            } else if (const JVal *o = jc[i].get("orthographic")) {
                c.perspective = false;
                c.fov_or_xmag = (float)o->number("xmag", 0.0);
                c.aspect_or_ymag = (float)o->number("ymag", 0.0);
                c.znear = (float)o->number("znear", 0.0);
                c.zfar = (float)o->number("zfar", 0.0);
            }
            cameras.push_back(c);
        }
        // nodes: flattened in document order, final transform = parent * local from the roots down
        const JVal &jn = array("nodes");
        std::vector<Mat4> local(jn.size());
        std::vector<char> is_child(jn.size(), 0);
        nodes.resize(jn.size());
        for (size_t i = 0; i < jn.size(); i++) {
            local[i] = local_transform(jn[i]);
            swr_node_desc nd{};
            std::memcpy(nd.transform, local[i].m, 64);
            const int64_t mi = jn[i].integer("mesh", -1);
            if (mi >= (int64_t)meshes.size()) throw std::runtime_error("Invalid data: node mesh index out of range");
            nd.mesh_index = (int32_t)mi;
            // Node::from_gltf: centre = LOCAL transform * origin, radius 0 (scene.rs:384-386)
            nd.bounding_sphere_world[0] = local[i].m[12], nd.bounding_sphere_world[1] = local[i].m[13], nd.bounding_sphere_world[2] = local[i].m[14];
            nd.bounding_sphere_world[3] = 0.0f;
            nodes[i] = nd;
            const JVal *ch = jn[i].get("children");
            for (size_t k = 0; ch && k < ch->size(); k++) {
                const int64_t c = (*ch)[k].as_index();
                if (c >= 0 && (size_t)c < is_child.size()) is_child[(size_t)c] = 1;
            }
        }
        std::vector<int> guard(jn.size(), 0);
        for (size_t i = 0; i < jn.size(); i++)
            if (!is_child[i]) final_transforms(i, Mat4::identity(), local, guard);
        for (size_t i = 0; i < jn.size(); i++) {
            const int64_t ci = jn[i].integer("camera", -1);
            if (ci >= 0 && (size_t)ci < cameras.size()) std::memcpy(cameras[(size_t)ci].transform, nodes[i].transform, 64);
        }
        // node spheres (scene.rs:318-329) and scene bounds (scene.rs:331-351)
        for (int c = 0; c < 3; c++) bounds_min[c] = INFINITY, bounds_max[c] = -INFINITY;
        for (swr_node_desc &nd : nodes) {
            if (nd.mesh_index < 0) continue;
            const swr_mesh_desc &md = meshes[(size_t)nd.mesh_index];
            for (uint32_t pi = md.first_primitive; pi < md.first_primitive + md.num_primitives; pi++) {
                const PrimitiveData &p = primitives[pi];
                float ts[4];
                transform_sphere(nd.transform, p.bounding_sphere, ts);
                const V3 dv{nd.bounding_sphere_world[0] - ts[0], nd.bounding_sphere_world[1] - ts[1], nd.bounding_sphere_world[2] - ts[2]};
                const float dist = length(dv);  // grow_by_sphere (scene.rs:44-49)
                if (dist + ts[3] > nd.bounding_sphere_world[3]) nd.bounding_sphere_world[3] = dist + ts[3];
                for (size_t v = 0; v + 3 < p.positions.size(); v += 4) {
                    float w[4];
                    mul_vec4(nd.transform, &p.positions[v], w);
                    for (int c = 0; c < 3; c++) {
                        bounds_min[c] = std::fmin(bounds_min[c], w[c]);
                        bounds_max[c] = std::fmax(bounds_max[c], w[c]);
                    }
                }
            }
        }
        for (int c = 0; c < 3; c++) bounds_center[c] = (bounds_min[c] + bounds_max[c]) * 0.5f;
        bounds_diagonal = length(V3{bounds_max[0] - bounds_min[0], bounds_max[1] - bounds_min[1], bounds_max[2] - bounds_min[2]});

        // environment: appended behind the file's textures, copied so the Document owns everything it points at
        int env_slot[3] = {-1, -1, -1};
        const swr_texture_desc *envs[3] = {env.cubemap, env.cubemap_specular, env.brdf_lut};
        for (int k = 0; k < 3; k++) {
            if (!envs[k]) continue;
            const swr_texture_desc &e = *envs[k];
            auto td = std::make_shared<TextureData>();
            td->width = e.width, td->height = e.height, td->type = e.texture_type;
            td->data.assign(e.data, e.data + e.ntexels);
            td->mip_offsets.assign(e.mip_offsets, e.mip_offsets + e.max_mip_level + 1);
            td->mip_widths.assign(e.mip_widths, e.mip_widths + e.max_mip_level + 1);
            td->mip_heights.assign(e.mip_heights, e.mip_heights + e.max_mip_level + 1);
            td->array_stride.assign(e.array_stride, e.array_stride + e.max_mip_level + 1);
            env_slot[k] = (int)texture_data.size();
            texture_data.push_back(td);
            texture_wrap.emplace_back(e.wrap_s, e.wrap_t);
            texture_uri.push_back(k == 0 ? "<cubemap>" : k == 1 ? "<cubemap_specular>" : "<brdf_lut>");
        }
        if (env.voxel_grid.gi_sh4) {
            const size_t nv = (size_t)env.voxel_grid.dims[0] * env.voxel_grid.dims[1] * env.voxel_grid.dims[2];
            voxels.assign(env.voxel_grid.gi_sh4, env.voxel_grid.gi_sh4 + nv * 16);
        }
        // flat descriptors
        prim_descs.resize(primitives.size());
        for (size_t i = 0; i < primitives.size(); i++) {
            const PrimitiveData &p = primitives[i];
            swr_primitive_desc &d = prim_descs[i];
            d.positions = p.positions.data(), d.normals = p.normals.data(), d.tangents = p.tangents.data(), d.texcoords = p.texcoords.data();
            d.indices = p.indices.data();
            d.nverts = p.nverts(), d.nindices = (uint32_t)p.indices.size(), d.material_index = p.material_index;
            std::memcpy(d.bounding_sphere, p.bounding_sphere, 16);
        }
        tex_descs.resize(texture_data.size());
        for (size_t i = 0; i < texture_data.size(); i++) {
            const TextureData &t = *texture_data[i];
            swr_texture_desc &d = tex_descs[i];
            d.data = t.data.data(), d.ntexels = (uint32_t)t.data.size(), d.width = t.width, d.height = t.height, d.texture_type = t.type;
            d.max_mip_level = t.max_mip_level();
            d.mip_offsets = t.mip_offsets.data(), d.mip_widths = t.mip_widths.data(), d.mip_heights = t.mip_heights.data(), d.array_stride = t.array_stride.data();
            d.wrap_s = texture_wrap[i].first, d.wrap_t = texture_wrap[i].second;
        }
        desc.primitives = prim_descs.data(), desc.nprimitives = (uint32_t)prim_descs.size();
        desc.meshes = meshes.data(), desc.nmeshes = (uint32_t)meshes.size();
        desc.nodes = nodes.data(), desc.nnodes = (uint32_t)nodes.size();
        desc.materials = materials.data(), desc.nmaterials = (uint32_t)materials.size();
        desc.textures = tex_descs.data(), desc.ntextures = (uint32_t)tex_descs.size();
        desc.voxel_grid = env.voxel_grid;
        if (std::memcmp(desc.voxel_grid.world_min, desc.voxel_grid.world_max, 12) == 0 && !nodes.empty()) {
            // no extent given: the grid spans the scene bounds, as in main.rs:228-235
            std::memcpy(desc.voxel_grid.world_min, bounds_min, 12);
            std::memcpy(desc.voxel_grid.world_max, bounds_max, 12);
        }
        desc.voxel_grid.gi_sh4 = voxels.empty() ? nullptr : voxels.data();
        desc.cubemap = env_slot[0], desc.cubemap_specular = env_slot[1], desc.brdf_lut = env_slot[2];
        std::memcpy(desc.light_direction, env.light_direction, 12);
        std::memcpy(desc.light_color, env.light_color, 12);
    }
};

}  // namespace gltf
}  // namespace swr
