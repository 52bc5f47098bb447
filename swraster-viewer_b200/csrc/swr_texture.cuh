// swr_texture.cuh — texel fetch and the reference's point sampler with its custom mip / wrap maths
// (texture.rs:577-589, 680-714, 851-863; util.rs:83-89). Shared by the shader and by the rasteriser's alpha test.
#pragma once
#include "swr_device.cuh"

// _mm_min_ps / _mm_max_ps semantics (second operand on NaN)
__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float sse_clamp(float a, float lo, float hi) { return sse_min(sse_max(a, lo), hi); }
// byte as f32 / 255.0 (util.rs:83-89) without the IEEE-division sequence: one multiply plus one exact-residual
// correction; bit-identical to the division for all 256 inputs (checked exhaustively, tests/test_gpu_parity.py).
__device__ __forceinline__ float unorm8(uint32_t b) {
    const float x = (float)b, rcp = 1.0f / 255.0f;
    const float q = __fmul_rn(x, rcp);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), rcp, q);
}
__device__ __forceinline__ float4 fetch_texel(const DevTex &t, uint32_t idx) {  // util.rs:83-89
    uint32_t p = tex1Dfetch<unsigned int>(t.obj, (int)idx);
    return make_float4(unorm8((p >> 24) & 0xFF), unorm8((p >> 16) & 0xFF), unorm8((p >> 8) & 0xFF), unorm8(p & 0xFF));
}
__device__ __forceinline__ float apply_wrap_mode(float texel, float dim, uint32_t mode) {  // texture.rs:578-589
    float t2 = texel;
    if (mode == 0u) {
        t2 = texel - floorf(texel / dim) * dim;
    } else if (mode == 1u) {
        float two = dim * 2.0f;
        float t = texel - floorf(texel / two) * two;
        t2 = sse_min(t, two - t);
    }
    return sse_min(t2, dim - 1.0f);
}
__device__ __forceinline__ uint32_t compute_mip_level(const DevTex &t, const float d[4]) {  // texture.rs:851-863
    float wf = (float)t.width, hf = (float)t.height;
    float d0 = d[0] * wf, d1 = d[1] * wf, d2 = d[2] * hf, d3 = d[3] * hf;
    float dx2 = d0 * d0 + d2 * d2;
    float dy2 = d1 * d1 + d3 * d3;
    float fp = (dx2 + dy2) * 0.5f;
    float fm = fp > 1.0f ? fp : 1.0f;  // f32::max(1.0)
    uint32_t mip = (31u - (uint32_t)__clz((int)__float2uint_rz(fm))) >> 1;
    return min(mip, t.max_mip);
}
__device__ __forceinline__ float4 sample4(const DevTex &t, float u, float v, const float du_dv[4]) {  // texture.rs:680-714
    uint32_t mip = compute_mip_level(t, du_dv);
    float wf = (float)t.mip_w[mip], hf = (float)t.mip_h[mip];
    uint32_t x = __float2uint_rz(apply_wrap_mode(floorf(u * wf), wf, t.wrap_s));
    uint32_t y = __float2uint_rz(apply_wrap_mode(floorf(v * hf), hf, t.wrap_t));
    return fetch_texel(t, t.mip_off[mip] + y * t.mip_w[mip] + x);
}
