// swr_cache.hpp — the reference's on-disk bake caches:
//   `.ggx`  prefiltered specular cubemap   src/texture.rs:12-17, 422-552   "GGX0", version 1, width, height, mips (LE u32), then
//                                                                         the brotli stream of width*height*6*mips u32 texels (LE)
//   `.gi`   voxel SH4 grid                 src/gi.rs:17-22, 30-122         "VGI0", version 8, w, h, d (LE u32), then the brotli
//                                                                         stream of w*h*d * 4 coefficients * (r, g, b, w) f32 (LE)
// The reference treats a missing file, a short file, another magic / version / size and a payload of the wrong length as
// "no cache" (it then bakes and writes one); only I/O and decoder failures are errors. Same here.
// Brotli itself is the system's libbrotlidec / libbrotlienc (the `brotli` crate is a port of the same format, RFC 7932),
// bound at run time with dlopen: the image carries the shared libraries but no headers. Without them the calls fail with a
// message — there is no second code path.
#pragma once
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace swr {
namespace cache {

struct Brotli {
    // BrotliDecoderDecompress(encoded_size, encoded, &decoded_size, decoded) -> 1 on success (BROTLI_DECODER_RESULT_SUCCESS)
    int (*decompress)(size_t, const uint8_t *, size_t *, uint8_t *) = nullptr;
    // BrotliEncoderCompress(quality, lgwin, mode, input_size, input, &encoded_size, encoded) -> BROTLI_TRUE
    int (*compress)(int, int, int, size_t, const uint8_t *, size_t *, uint8_t *) = nullptr;
    size_t (*max_compressed)(size_t) = nullptr;
    static const Brotli &get() {
        static Brotli b;
        static std::once_flag once;
        std::call_once(once, [] {
            if (void *d = dlopen("libbrotlidec.so.1", RTLD_NOW | RTLD_GLOBAL)) b.decompress = (decltype(b.decompress))dlsym(d, "BrotliDecoderDecompress");
            if (void *e = dlopen("libbrotlienc.so.1", RTLD_NOW | RTLD_GLOBAL)) {
                b.compress = (decltype(b.compress))dlsym(e, "BrotliEncoderCompress");
                b.max_compressed = (decltype(b.max_compressed))dlsym(e, "BrotliEncoderMaxCompressedSize");
            }
        });
        return b;
    }
};

inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void put32(std::vector<uint8_t> &v, uint32_t x) {
    for (int k = 0; k < 4; k++) v.push_back((uint8_t)(x >> (8 * k)));
}

inline bool read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;  // path.exists() == false -> Ok(None)
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) {
        std::fclose(f);
        throw std::runtime_error(std::string("cannot read ") + path);
    }
    out.resize((size_t)n);
    const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != (size_t)n) throw std::runtime_error(std::string("short read on ") + path);
    return true;
}

// header check shared by both formats: 20 bytes = magic, version, three u32
inline bool header_ok(const std::vector<uint8_t> &b, const char magic[4], uint32_t version, uint32_t a, uint32_t c, uint32_t d) {
    return b.size() >= 20 && std::memcmp(b.data(), magic, 4) == 0 && le32(&b[4]) == version && le32(&b[8]) == a && le32(&b[12]) == c && le32(&b[16]) == d;
}

// payload of exactly `bytes` bytes, or "no cache" (false) when the stream decodes to another length
inline bool inflate_exact(const std::vector<uint8_t> &file, size_t bytes, uint8_t *out) {
    const Brotli &br = Brotli::get();
    if (!br.decompress) throw std::runtime_error("libbrotlidec.so.1 (BrotliDecoderDecompress) is not available: the cache cannot be read");
    // one spare byte: a longer payload then fails with "output too small" instead of silently passing
    std::vector<uint8_t> tmp(bytes + 1);
    size_t n = tmp.size();
    const int rc = br.decompress(file.size() - 20, file.data() + 20, &n, tmp.data());
    if (rc != 1 || n != bytes) return false;
    std::memcpy(out, tmp.data(), bytes);
    return true;
}

inline void deflate_to_file(const char *path, std::vector<uint8_t> header, const uint8_t *payload, size_t bytes) {
    const Brotli &br = Brotli::get();
    if (!br.compress || !br.max_compressed) throw std::runtime_error("libbrotlienc.so.1 (BrotliEncoderCompress) is not available: the cache cannot be written");
    std::vector<uint8_t> z(br.max_compressed(bytes) + 64);
    size_t n = z.size();
    // quality 6, lgwin 22, generic mode: GGX_CACHE_BROTLI_* / GI_CACHE_BROTLI_* (texture.rs:15-17, gi.rs:20-22)
    if (!br.compress(6, 22, 0, bytes, payload, &n, z.data())) throw std::runtime_error("BrotliEncoderCompress failed");
    header.insert(header.end(), z.begin(), z.begin() + (long)n);
    FILE *f = std::fopen(path, "wb");
    if (!f) throw std::runtime_error(std::string("cannot write ") + path);
    const size_t put = std::fwrite(header.data(), 1, header.size(), f);
    if (std::fclose(f) != 0 || put != header.size()) throw std::runtime_error(std::string("short write on ") + path);
}

inline uint32_t ggx_mips(uint32_t w, uint32_t h) {  // 1 + ilog2(max(w, h))
    uint32_t n = 1;
    for (uint32_t m = w > h ? w : h; m >>= 1;) n++;
    return n;
}

// try_load_prefiltered_specular_cubemap_cache (texture.rs:426-514): texels of every mip at full face resolution
inline bool ggx_load(const char *path, uint32_t w, uint32_t h, std::vector<uint32_t> &texels) {
    std::vector<uint8_t> file;
    if (!w || !h || !read_file(path, file) || !header_ok(file, "GGX0", 1, w, h, ggx_mips(w, h))) return false;
    texels.assign((size_t)w * h * 6 * ggx_mips(w, h), 0u);
    std::vector<uint8_t> raw(texels.size() * 4);
    if (!inflate_exact(file, raw.size(), raw.data())) return false;
    for (size_t i = 0; i < texels.size(); i++) texels[i] = le32(&raw[4 * i]);
    return true;
}
// save_prefiltered_specular_cubemap_cache (texture.rs:516-552)
inline void ggx_save(const char *path, uint32_t w, uint32_t h, uint32_t mips, const uint32_t *texels, size_t ntexels) {
    if (ntexels != (size_t)w * h * 6 * mips) throw std::runtime_error("Invalid data: prefiltered cubemap texture has unexpected data size");
    std::vector<uint8_t> hdr(4), raw(ntexels * 4);
    std::memcpy(hdr.data(), "GGX0", 4);
    put32(hdr, 1), put32(hdr, w), put32(hdr, h), put32(hdr, mips);
    for (size_t i = 0; i < ntexels; i++)
        for (int k = 0; k < 4; k++) raw[4 * i + k] = (uint8_t)(texels[i] >> (8 * k));
    deflate_to_file(path, hdr, raw.data(), raw.size());
}

// try_load_gi_cache (gi.rs:30-83): w*h*d voxels x 4 coefficients x (r, g, b, w), the layout of swr_voxel_grid.gi_sh4
inline bool gi_load(const char *path, uint32_t w, uint32_t h, uint32_t d, float *gi_sh4) {
    std::vector<uint8_t> file;
    if (!read_file(path, file) || !header_ok(file, "VGI0", 8, w, h, d)) return false;
    const size_t nfloats = (size_t)w * h * d * 16;
    std::vector<uint8_t> raw(nfloats * 4);
    if (!inflate_exact(file, raw.size(), raw.data())) return false;
    for (size_t i = 0; i < nfloats; i++) {
        const uint32_t bits = le32(&raw[4 * i]);
        std::memcpy(&gi_sh4[i], &bits, 4);
    }
    return true;
}
// save_gi_cache (gi.rs:85-118)
inline void gi_save(const char *path, uint32_t w, uint32_t h, uint32_t d, const float *gi_sh4) {
    const size_t nfloats = (size_t)w * h * d * 16;
    std::vector<uint8_t> hdr(4), raw(nfloats * 4);
    std::memcpy(hdr.data(), "VGI0", 4);
    put32(hdr, 8), put32(hdr, w), put32(hdr, h), put32(hdr, d);
    for (size_t i = 0; i < nfloats; i++) {
        uint32_t bits;
        std::memcpy(&bits, &gi_sh4[i], 4);
        for (int k = 0; k < 4; k++) raw[4 * i + k] = (uint8_t)(bits >> (8 * k));
    }
    deflate_to_file(path, hdr, raw.data(), raw.size());
}

}  // namespace cache
}  // namespace swr
