// vr_procedural.h — the procedural density fields of the synthetic scenes, shared by the host scene builder
// (vr_scene.cpp, g++ -ffp-contract=off) and the device generator (vr_mipbuild.cu, nvcc --fmad=false): same expressions, same
// order, so the two produce identical voxels wherever no libm function is involved (every kind except the plume's sin / cos).
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/vrestir.h"

#ifdef __CUDACC__
#define VR_HD __host__ __device__
#else
#define VR_HD
#endif

namespace vr {
VR_HD inline float vr_fmin(float a, float b) { return a < b ? a : b; }
VR_HD inline float vr_fmax(float a, float b) { return a < b ? b : a; }

VR_HD inline uint32_t hash3(int x, int y, int z, uint32_t seed) {
    uint32_t h = seed * 0x9E3779B1u + 0x7F4A7C15u;
    h ^= (uint32_t)x * 0x85EBCA6Bu; h = (h << 13) | (h >> 19); h *= 0xC2B2AE35u;
    h ^= (uint32_t)y * 0x27D4EB2Fu; h = (h << 15) | (h >> 17); h *= 0x165667B1u;
    h ^= (uint32_t)z * 0x9E3779B1u; h = (h << 11) | (h >> 21); h *= 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
VR_HD inline float lattice(int x, int y, int z, uint32_t seed) { return (float)(hash3(x, y, z, seed) >> 8) * (1.0f / 16777216.0f); }
VR_HD inline float smooth(float t) { return t * t * (3.f - 2.f * t); }
VR_HD inline float valueNoise(float x, float y, float z, uint32_t seed) {
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float tx = smooth(x - fx), ty = smooth(y - fy), tz = smooth(z - fz);
    float c[8];
    for (int i = 0; i < 8; i++) c[i] = lattice(ix + (i & 1), iy + ((i >> 1) & 1), iz + (i >> 2), seed);
    float x00 = c[0] + tx * (c[1] - c[0]), x10 = c[2] + tx * (c[3] - c[2]), x01 = c[4] + tx * (c[5] - c[4]), x11 = c[6] + tx * (c[7] - c[6]);
    float y0 = x00 + ty * (x10 - x00), y1 = x01 + ty * (x11 - x01);
    return y0 + tz * (y1 - y0);
}
VR_HD inline float fbm(float x, float y, float z, uint32_t seed, int octaves = 5) {
    float sum = 0.f, amp = 0.5f, norm = 0.f;
    for (int o = 0; o < octaves; o++) {
        sum += amp * valueNoise(x, y, z, seed + 101u * (uint32_t)o);
        norm += amp; amp *= 0.5f; x *= 2.f; y *= 2.f; z *= 2.f;
    }
    return sum / norm;
}

// ------------------------------------------------------------------------------------------------ procedural fields
VR_HD inline float ellipsoid(float x, float y, float z, float cx, float cy, float cz, float rx, float ry, float rz) {
    float dx = (x - cx) / rx, dy = (y - cy) / ry, dz = (z - cz) / rz;
    return 1.f - sqrtf(dx * dx + dy * dy + dz * dz);   // > 0 inside
}
VR_HD inline float shapeDensity(const vrestir_scene_params& sp, float u, float v, float w, float* temperature, float vel[3]) {
    // (u,v,w) in [0,1]^3 over the grid; returns density in [0,1]
    const uint32_t seed = sp.seed;
    const float f = 6.f;
    switch (sp.kind) {
        case 0: {   // sphere (radius 0.375 of the box) x fBm, SURVEY.md 8(d) config 1
            float dx = u - 0.5f, dy = v - 0.5f, dz = w - 0.5f;
            float r = sqrtf(dx * dx + dy * dy + dz * dz);
            if (r > 0.375f) return 0.f;
            float n = fbm(u * f, v * f, w * f, seed);
            float edge = vr_fmin(1.f, (0.375f - r) * 16.f);
            return vr_fmax(0.f, n - 0.35f) / 0.65f * edge;
        }
        case 1: {   // bunny-cloud-like blob: body + head + two ears, eroded by fBm
            float s = -1.f;
            s = vr_fmax(s, ellipsoid(u, v, w, 0.50f, 0.36f, 0.52f, 0.36f, 0.30f, 0.34f));
            s = vr_fmax(s, ellipsoid(u, v, w, 0.30f, 0.62f, 0.50f, 0.20f, 0.19f, 0.21f));
            s = vr_fmax(s, ellipsoid(u, v, w, 0.27f, 0.85f, 0.40f, 0.065f, 0.16f, 0.08f));
            s = vr_fmax(s, ellipsoid(u, v, w, 0.30f, 0.85f, 0.61f, 0.065f, 0.16f, 0.08f));
            s = vr_fmax(s, ellipsoid(u, v, w, 0.82f, 0.30f, 0.52f, 0.10f, 0.10f, 0.10f));
            if (s < -0.25f) return 0.f;
            float n = fbm(u * 7.f, v * 7.f, w * 7.f, seed);
            float d = s * 2.2f + (n - 0.5f) * 1.1f;
            return vr_fmin(1.f, vr_fmax(0.f, d) * 2.5f);
        }
        case 2: {   // plume: rising turbulent column, advected upward with frame_time
            float t = sp.frame_time;
            float cx = 0.5f + 0.06f * sinf(6.f * v + 0.7f * t), cz = 0.5f + 0.06f * cosf(5.f * v + 0.9f * t);
            float rad = 0.07f + 0.22f * v;
            float dx = u - cx, dz = w - cz;
            float r = sqrtf(dx * dx + dz * dz) / rad;
            float rise = vr_fmin(1.f, 0.25f + 0.05f * t);
            float body = (r < 1.f && v < rise) ? (1.f - r) : 0.f;
            float n = fbm(u * 8.f, (v - 0.04f * t) * 8.f, w * 8.f, seed);
            float d = body * vr_fmax(0.f, n - 0.3f) * 2.4f * vr_fmin(1.f, (rise - v) * 12.f);
            d = vr_fmin(1.f, vr_fmax(0.f, d));
            if (temperature) *temperature = d > 0.f ? 2000.f * vr_fmax(0.f, 1.f - v / vr_fmax(rise, 1e-3f)) * vr_fmin(1.f, d * 3.f) : 0.f;
            if (vel) {   // curl-like swirl + rise, |v| <= 2 voxels / frame (index units of mip 0)
                float sw = 1.2f * (n - 0.5f);
                vel[0] = d > 0.f ? -dz / vr_fmax(rad, 1e-3f) * sw : 0.f;
                vel[1] = d > 0.f ? 1.5f * (1.f - 0.5f * r) : 0.f;
                vel[2] = d > 0.f ? dx / vr_fmax(rad, 1e-3f) * sw : 0.f;
            }
            return d;
        }
        case 3: {   // dense cloud filling most of the box
            float n = fbm(u * 5.f, v * 5.f, w * 5.f, seed);
            float bx = vr_fmin(vr_fmin(u, 1.f - u), vr_fmin(vr_fmin(v, 1.f - v), vr_fmin(w, 1.f - w)));
            float edge = vr_fmin(1.f, bx * 12.f);
            return vr_fmin(1.f, vr_fmax(0.f, n - 0.38f) * 3.0f) * edge;
        }
        default: {  // thin shells: three fBm-perturbed spherical shells, ~3 % of the voxels / ~5 % of the
                    // 8^3 bricks of a 2048^3 grid, so that the brick pools (GBs) are far larger than L2
            float dx = u - 0.5f, dy = v - 0.5f, dz = w - 0.5f;
            float r = sqrtf(dx * dx + dy * dy + dz * dz);
            if (r > 0.47f || r < 0.13f) return 0.f;
            float q = r + 0.03f * (fbm(u * 6.f, v * 6.f, w * 6.f, seed, 3) - 0.5f);
            const float h = 0.003f;
            float d = vr_fmin(vr_fmin(fabsf(q - 0.18f), fabsf(q - 0.30f)), fabsf(q - 0.42f));
            return vr_fmax(0.f, 1.f - d / h);
        }
    }
}


}  // namespace vr
