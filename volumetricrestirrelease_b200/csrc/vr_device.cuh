// vr_device.cuh — device-side library of the VolumetricReSTIR hot path for sm_100a.
//
// Layers (each function cites the reference logic it implements; VR/ = Source/RenderPasses/VolumetricReSTIR/,
// F/ = Source/Falcor/):
//   RNG               F/Utils/Sampling/UniformSampleGenerator.slang:49-75, SampleGenerator.slang:58-73
//   grid access       F/Scene/GVDB/{gvdbNodes,gvdb}.slang, VR/VolumeBase.slang:103-263  (brick pool instead of a 3-D atlas)
//   hierarchical DDA  F/Scene/GVDB/gvdbDda.slang:86-157, VR/VolumeUtils.slang:171-282
//   tracking adapters VR/VolumeTrackingAdapterGVDB.slang
//   lights            F/Experimental/Scene/Lights/*.slang, VR/VolumeUtils.slang:12-169,419-492
//   ReSTIR            VR/{Reservoir,ReSTIRHelper,ComputeInitialSample}.slang
//
// Numerics contract: this translation unit is compiled with --fmad=false, IEEE div/sqrt, no fast-math, so that the
// traversal geometry (cell sequence, RNG draw order) is reproducible against the CPU oracle bit for bit; only libm-level
// functions (expf/logf/sinf/cosf/atan2f/acosf/powf) may differ in the last ulp.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <float.h>
#include "../../include/vrestir.h"
#include "vr_types.h"

namespace vrd {
namespace cg = cooperative_groups;

#define VRD __device__ __forceinline__
#define VRD_NOINLINE static __device__ __noinline__

static __constant__ DPrevCam c_prev;   // same scheme as c_scene below
static __constant__ DScene c_scene;   // one private copy per translation unit (vr_kernels.cu, vr_wavefront.cu); uploadScene* fills each
// diagnostics: rays whose hierarchical DDA ran >= 1024 outer iterations (origin, dir, mip, iterations), first 64
static __device__ float g_dbgRays[64 * 8];
static __device__ unsigned g_dbgCount;


// ------------------------------------------------------------------------------------------------ math
constexpr float kRayTMax = FLT_MAX;
constexpr float kPi = 3.14159265358979323846f, k2Pi = 6.28318530717958647693f, k1_Pi = 0.318309886183790671538f;
constexpr float k1_2Pi = 0.159154943091895335769f, k1_4Pi = 0.079577471545947667884f, kPi_4 = 0.785398163397448309616f;
constexpr uint32_t ID_UNDEFL = 0xFFFFFFFFu;
constexpr int MAX_BRICK_STEPS = 128;
constexpr float kUnorm8 = 0.003921568859368563f;

VRD float3 f3(float a) { return make_float3(a, a, a); }
VRD float3 f3(float a, float b, float c) { return make_float3(a, b, c); }
VRD float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
VRD float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
VRD float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
VRD float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
VRD float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
VRD float3 operator*(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
VRD float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }
VRD float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
VRD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VRD float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
VRD float length(float3 a) { return sqrtf(dot(a, a)); }
VRD float3 normalize(float3 a) { return a / sqrtf(dot(a, a)); }
VRD bool any_gt0(float3 a) { return a.x > 0.f || a.y > 0.f || a.z > 0.f; }
VRD bool all_eq0(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
VRD float fsign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
VRD int f2i(float v) { return __float2int_rz(v); }            // saturating, NaN -> 0 (D3D ftoi)
VRD uint32_t f2u(float v) { return __float2uint_rz(v); }
VRD float luminance(float3 c) { return dot(c, f3(0.2126f, 0.7152f, 0.0722f)); }
VRD float lerpf(float a, float b, float t) { return __fmaf_rn(t, b - a, a); }   // pinned: one fused multiply-add (oracle: fmaf)
VRD float3 v3(const float* a) { return make_float3(a[0], a[1], a[2]); }
VRD float3 mulPoint(float3 p, const float* M) {
    return make_float3(p.x * M[0] + p.y * M[4] + p.z * M[8] + M[12], p.x * M[1] + p.y * M[5] + p.z * M[9] + M[13],
                       p.x * M[2] + p.y * M[6] + p.z * M[10] + M[14]);
}
VRD float3 mulVec(float3 v, const float* M) {
    return make_float3(v.x * M[0] + v.y * M[4] + v.z * M[8], v.x * M[1] + v.y * M[5] + v.z * M[9], v.x * M[2] + v.y * M[6] + v.z * M[10]);
}
VRD float3 mulVec3x3(float3 v, const float* M) {
    return make_float3(v.x * M[0] + v.y * M[3] + v.z * M[6], v.x * M[1] + v.y * M[4] + v.z * M[7], v.x * M[2] + v.y * M[5] + v.z * M[8]);
}

// ------------------------------------------------------------------------------------------------ RNG
VRD uint32_t interleave_32bit(uint32_t vx, uint32_t vy) {
    uint32_t x = vx & 0x0000ffffu, y = vy & 0x0000ffffu;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu; y = (y | (y << 4)) & 0x0F0F0F0Fu; y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}
struct SampleGenerator {
    uint32_t s0, s1, s2, s3;
    static VRD uint64_t splitmix(uint64_t& st) {
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    static VRD SampleGenerator create(uint32_t px, uint32_t py, uint32_t sampleNumber) {
        uint64_t st = ((uint64_t)sampleNumber << 32) | (uint64_t)interleave_32bit(px, py);
        uint64_t a = splitmix(st), b = splitmix(st);
        SampleGenerator g; g.s0 = (uint32_t)a; g.s1 = (uint32_t)(a >> 32); g.s2 = (uint32_t)b; g.s3 = (uint32_t)(b >> 32);
        return g;
    }
    VRD uint32_t next() {
        const uint32_t r = __funnelshift_l(s0 * 5u, s0 * 5u, 7) * 9u;
        const uint32_t t = s1 << 9;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = __funnelshift_l(s3, s3, 11);
        return r;
    }
};
VRD float sampleNext1D(SampleGenerator& sg) { return (float)(sg.next() >> 8) * 0x1p-24f; }
VRD float2 sampleNext2D(SampleGenerator& sg) { float2 r; r.x = sampleNext1D(sg); r.y = sampleNext1D(sg); return r; }

// ------------------------------------------------------------------------------------------------ records
struct Ray { float3 origin, dir; float tMin, tMax; VRD float3 at(float t) const { return origin + dir * t; } };
// p_partial: VERTEX_REUSE only (VR/HostDeviceSharedDefinitions.h:29-31) — the part of p-hat past the reuse vertex; plane p2 of a ResBuf
struct Reservoir { float runningSum, M, depth, p_y; float2 lightUV; int lightID, sampledPixel; int extraBounceStartId; float p_partial; };

VRD Reservoir createNewReservoir() { Reservoir r; r.runningSum = 0.f; r.M = 0.f; r.depth = FLT_MAX; r.p_y = 0.f; r.lightUV = make_float2(0, 0); r.lightID = 0; r.sampledPixel = 0; r.extraBounceStartId = 0; r.p_partial = 0.f; return r; }
VRD Reservoir loadReservoir(const ResBuf& b, int pixelId, int B) {
    float4 a = __ldg(&b.p0[pixelId]), c = __ldg(&b.p1[pixelId]);
    Reservoir r; r.runningSum = a.x; r.M = a.y; r.depth = a.z; r.p_y = a.w; r.lightUV = make_float2(c.x, c.y);
    r.lightID = __float_as_int(c.z); r.sampledPixel = __float_as_int(c.w); r.extraBounceStartId = B > 1 ? pixelId * (B - 1) : 0;
    r.p_partial = (B > 1 && b.p2) ? __ldg(&b.p2[pixelId]) : 0.f;
    return r;
}
VRD Reservoir loadReservoirRW(const ResBuf& b, int pixelId, int B) {   // buffer written by this kernel: no read-only path
    float4 a = b.p0[pixelId], c = b.p1[pixelId];
    Reservoir r; r.runningSum = a.x; r.M = a.y; r.depth = a.z; r.p_y = a.w; r.lightUV = make_float2(c.x, c.y);
    r.lightID = __float_as_int(c.z); r.sampledPixel = __float_as_int(c.w); r.extraBounceStartId = B > 1 ? pixelId * (B - 1) : 0;
    r.p_partial = (B > 1 && b.p2) ? b.p2[pixelId] : 0.f;
    return r;
}
VRD void storeReservoir(const ResBuf& b, int pixelId, const Reservoir& r) {
    b.p0[pixelId] = make_float4(r.runningSum, r.M, r.depth, r.p_y);
    b.p1[pixelId] = make_float4(r.lightUV.x, r.lightUV.y, __int_as_float(r.lightID), __int_as_float(r.sampledPixel));
    if (b.p2) b.p2[pixelId] = r.p_partial;
}

// ------------------------------------------------------------------------------------------------ tree access
VRD float3 nodePos(const int4& a) { return make_float3((float)a.x, (float)a.y, (float)a.z); }
struct NodeHead { int4 a; };   // (pos.xyz, link)
VRD NodeHead loadNodeHead(const DSlot& g, int lev, uint32_t n) { NodeHead h; h.a = __ldg((const int4*)&g.nodes[lev][n]); return h; }
VRD float4 loadNodeBounds(const DSlot& g, int lev, uint32_t n) { return __ldg(((const float4*)&g.nodes[lev][n]) + 1); }
VRD uint32_t getChild(const DSlot& g, uint32_t listid, int clev, int b) {
    if (listid == ID_UNDEFL) return ID_UNDEFL;
    const unsigned long long r = (unsigned long long)(g.res[clev] * g.res[clev] * g.res[clev]);
    return __ldg(&g.child[clev][(unsigned long long)listid * r + (unsigned long long)b]);
}
VRD bool outsideBox(float3 pos, float3 vmin, float ext) {
    return pos.x < vmin.x || pos.y < vmin.y || pos.z < vmin.z || pos.x >= vmin.x + ext || pos.y >= vmin.y + ext || pos.z >= vmin.z + ext;
}
// F/Scene/GVDB/gvdbNodes.slang:175-280; returns brick id (ID_UNDEFL if none) and the brick's index-space min corner
VRD uint32_t getBrickAtPoint(const DSlot& g, float3 pos, float3& vminOut) {
    NodeHead node; float3 vmin;
    if (g.top_lev == 2) {
        node = loadNodeHead(g, 2, 0); vmin = nodePos(node.a);
        if (outsideBox(pos, vmin, 4096.f)) return ID_UNDEFL;
        int px = f2i(pos.x - vmin.x) / 128, py = f2i(pos.y - vmin.y) / 128, pz = f2i(pos.z - vmin.z) / 128;
        uint32_t id = getChild(g, (uint32_t)node.a.w, 2, (((pz << 5) + py) << 5) + px);
        if (id == ID_UNDEFL) return ID_UNDEFL;
        node = loadNodeHead(g, 1, id); vmin = nodePos(node.a);
    } else {
        node = loadNodeHead(g, 1, 0); vmin = nodePos(node.a);
    }
    if (outsideBox(pos, vmin, 128.f)) return ID_UNDEFL;
    int px = f2i(pos.x - vmin.x) / 8, py = f2i(pos.y - vmin.y) / 8, pz = f2i(pos.z - vmin.z) / 8;
    uint32_t id = getChild(g, (uint32_t)node.a.w, 1, (((pz << 4) + py) << 4) + px);
    if (id == ID_UNDEFL) return ID_UNDEFL;
    node = loadNodeHead(g, 0, id);
    vminOut = nodePos(node.a);
    return (uint32_t)node.a.w;
}

// ---- brick-pool fetch ----
template <bool CHECKED>
VRD float atlasVoxelRaw(const DSlot& g, uint32_t brick, int ix, int iy, int iz, int ch = 0) {   // stored code, no UNORM scale
    if (CHECKED) { if ((unsigned)(ix + 1) > 9u || (unsigned)(iy + 1) > 9u || (unsigned)(iz + 1) > 9u) return 0.f; }
    const unsigned idx = (brick * (unsigned)g.channels + (unsigned)ch) * VRESTIR_BRICK_VOXELS + (unsigned)(((iz + 1) * 10 + (iy + 1)) * 10 + (ix + 1));
    if (g.format == VRESTIR_ATLAS_UNORM8) return (float)__ldg(&((const uint8_t*)g.atlas)[idx]);
    return __ldg(&((const float*)g.atlas)[idx]);
}
template <bool CHECKED>
VRD float atlasVoxel(const DSlot& g, uint32_t brick, int ix, int iy, int iz, int ch = 0) {
    const float r = atlasVoxelRaw<CHECKED>(g, brick, ix, iy, iz, ch);
    return g.format == VRESTIR_ATLAS_UNORM8 ? r * kUnorm8 : r;
}
template <bool CHECKED>
VRD float sampleBrickLinear(const DSlot& g, uint32_t brick, float3 p, int ch = 0) {
    float qx = p.x - 0.5f, qy = p.y - 0.5f, qz = p.z - 0.5f;
    float fx0 = floorf(qx), fy0 = floorf(qy), fz0 = floorf(qz);
    int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
    float fx = qx - fx0, fy = qy - fy0, fz = qz - fz0;
    float v000, v100, v010, v110, v001, v101, v011, v111;
    const bool u8 = g.format == VRESTIR_ATLAS_UNORM8;
    if (!CHECKED) {
        if (u8 && g.quads) {
            // 2 x LDG.32: each word holds the 2x2 xy-neighbourhood of one z plane (raw UNORM8 codes)
            const uint32_t* q = g.quads + (brick * 810u + (unsigned)(((iz + 1) * 9 + (iy + 1)) * 9 + (ix + 1)));
            const uint32_t w0 = __ldg(q), w1 = __ldg(q + 81);
            v000 = (float)(w0 & 0xffu); v100 = (float)((w0 >> 8) & 0xffu); v010 = (float)((w0 >> 16) & 0xffu); v110 = (float)(w0 >> 24);
            v001 = (float)(w1 & 0xffu); v101 = (float)((w1 >> 8) & 0xffu); v011 = (float)((w1 >> 16) & 0xffu); v111 = (float)(w1 >> 24);
        } else {
            const unsigned base = (brick * (unsigned)g.channels + (unsigned)ch) * VRESTIR_BRICK_VOXELS + (unsigned)(((iz + 1) * 10 + (iy + 1)) * 10 + (ix + 1));
            if (u8) {
                const uint8_t* a = (const uint8_t*)g.atlas + base;
                v000 = (float)__ldg(a); v100 = (float)__ldg(a + 1); v010 = (float)__ldg(a + 10); v110 = (float)__ldg(a + 11);
                v001 = (float)__ldg(a + 100); v101 = (float)__ldg(a + 101); v011 = (float)__ldg(a + 110); v111 = (float)__ldg(a + 111);
            } else {
                const float* a = (const float*)g.atlas + base;
                v000 = __ldg(a); v100 = __ldg(a + 1); v010 = __ldg(a + 10); v110 = __ldg(a + 11);
                v001 = __ldg(a + 100); v101 = __ldg(a + 101); v011 = __ldg(a + 110); v111 = __ldg(a + 111);
            }
        }
    } else {
        v000 = atlasVoxelRaw<true>(g, brick, ix, iy, iz, ch); v100 = atlasVoxelRaw<true>(g, brick, ix + 1, iy, iz, ch);
        v010 = atlasVoxelRaw<true>(g, brick, ix, iy + 1, iz, ch); v110 = atlasVoxelRaw<true>(g, brick, ix + 1, iy + 1, iz, ch);
        v001 = atlasVoxelRaw<true>(g, brick, ix, iy, iz + 1, ch); v101 = atlasVoxelRaw<true>(g, brick, ix + 1, iy, iz + 1, ch);
        v011 = atlasVoxelRaw<true>(g, brick, ix, iy + 1, iz + 1, ch); v111 = atlasVoxelRaw<true>(g, brick, ix + 1, iy + 1, iz + 1, ch);
    }
    // pinned filter: fp32 fma-lerp x, y, z on the stored codes; UNORM8 codes are scaled by fl(1/255) once, after filtering
    float c00 = lerpf(v000, v100, fx), c10 = lerpf(v010, v110, fx), c01 = lerpf(v001, v101, fx), c11 = lerpf(v011, v111, fx);
    float c0 = lerpf(c00, c10, fy), c1 = lerpf(c01, c11, fy);
    float r = lerpf(c0, c1, fz);
    return u8 ? r * kUnorm8 : r;
}
template <bool CHECKED>
VRD float sampleBrickPoint(const DSlot& g, uint32_t brick, float3 p, int ch = 0) {
    return atlasVoxel<CHECKED>(g, brick, (int)floorf(p.x), (int)floorf(p.y), (int)floorf(p.z), ch);
}
// F/Scene/GVDB/gvdb.slang:6-31
VRD float getValueAtPoint(const DSlot& g, float3 pos, bool linear, int ch = 0) {
    float3 vmin;
    uint32_t brick = getBrickAtPoint(g, pos, vmin);
    if (brick == ID_UNDEFL) return 0.f;
    float3 p_rel = pos - vmin;
    return linear ? sampleBrickLinear<false>(g, brick, p_rel, ch) : sampleBrickPoint<false>(g, brick, p_rel, ch);   // p_rel in [0,8)^3 by construction
}

// ------------------------------------------------------------------------------------------------ VolumeBase
VRD Ray WorldToMedium(const Ray& r, int mip) {
    const float* M = c_scene.slots[mip].w2m;
    Ray o; o.origin = mulPoint(r.origin, M); o.dir = mulVec(r.dir, M); o.tMin = r.tMin; o.tMax = r.tMax;
    return o;
}
VRD float3 WorldToMediumP(float3 pW, int mip) { return mulPoint(pW, c_scene.slots[mip].w2m); }

VRD bool IntersectP(float3 pMin, float3 pMax, const Ray& ray, float& hitt0, float& hitt1) {
    float t0 = 0, t1 = ray.tMax;
#define VRD_SLAB(c)                                                                 \
    {                                                                               \
        float invRayDir = 1 / ray.dir.c;                                            \
        float tNear = (pMin.c - ray.origin.c) * invRayDir;                          \
        float tFar = (pMax.c - ray.origin.c) * invRayDir;                           \
        if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }          \
        t0 = tNear > t0 ? tNear : t0;                                               \
        t1 = tFar < t1 ? tFar : t1;                                                 \
        if (t0 > t1) return false;                                                  \
    }
    VRD_SLAB(x) VRD_SLAB(y) VRD_SLAB(z)
#undef VRD_SLAB
    hitt0 = t0; hitt1 = t1;
    return true;
}
VRD bool IntersectVolumeBound(const Ray& ray, float& tMin, float& tMax, int mip, bool vertexCenter) {
    const DSlot& g = c_scene.slots[mip];
    float3 mn = v3(g.bmin), mx = v3(g.bmax);
    if (vertexCenter) { mn = mn - f3(0.5f); mx = mx - f3(0.5f); }
    return IntersectP(mn, mx, ray, tMin, tMax);
}
VRD float GetVolumeMaxDensity(int mip) { return c_scene.slots[mip].max_value * c_scene.vol.densityScaleFactorByScaling; }
VRD float Density(float3 p, int mip) {
    const DSlot& g = c_scene.slots[mip];
    if (p.x < g.bmin[0] || p.y < g.bmin[1] || p.z < g.bmin[2] || p.x >= g.bmax[0] || p.y >= g.bmax[1] || p.z >= g.bmax[2]) return 0.f;
    return getValueAtPoint(g, p, true) * c_scene.vol.densityScaleFactorByScaling;
}
VRD float DensityWorldSpace(float3 pW, int mip) { return Density(WorldToMediumP(pW, mip), mip); }
template <bool CHECKED>
VRD float DensityInAtlas(const DSlot& g, uint32_t brick, float3 p_local, bool linear) {
    float s = linear ? sampleBrickLinear<CHECKED>(g, brick, p_local) : sampleBrickPoint<CHECKED>(g, brick, p_local);
    return s * g.compress_scale * c_scene.vol.densityScaleFactorByScaling;
}
VRD void FetchEightVoxels(const DSlot& g, uint32_t brick, int3 cp, float v[8]) {
    const float s1 = g.compress_scale, s2 = c_scene.vol.densityScaleFactorByScaling;
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = atlasVoxel<true>(g, brick, cp.x + (i % 2), cp.y + (i % 4) / 2, cp.z + i / 4) * s1 * s2;
}

VRD float3 ConvertTempToColor(float temp) {
    const vrestir_volume_desc& vd = c_scene.vol;
    temp = fminf(6400.f, (temp - vd.temperatureCutOff) * vd.temperatureScale);
    float queryPoint = (temp - 25) / 6400;
    float3 rgb = f3(0.f);
    if (c_scene.lut) {
        float x = queryPoint * 128.f - 0.5f;
        float x0 = floorf(x); float fx = x - x0; int i0 = f2i(x0), i1 = i0 + 1;
        float3 a = f3(0.f), b = f3(0.f);
        if (i0 >= 0 && i0 <= 127) { float4 t = __ldg(&c_scene.lut[i0]); a = f3(t.x, t.y, t.z); }
        if (i1 >= 0 && i1 <= 127) { float4 t = __ldg(&c_scene.lut[i1]); b = f3(t.x, t.y, t.z); }
        rgb = f3(lerpf(a.x, b.x, fx), lerpf(a.y, b.y, fx), lerpf(a.z, b.z, fx));
    }
    return vd.LeScale * rgb;
}
VRD float3 EmissionWorldSpace(float3 pW, bool isLastFrame = false) {
    const vrestir_volume_desc& vd = c_scene.vol;
    if ((!isLastFrame && !vd.hasEmission) || (isLastFrame && !vd.lastFrameHasEmission)) return f3(0.f);
    int slot = isLastFrame ? VRESTIR_TEMPERATURE_GRID_ID + VRESTIR_PREV_EXTRA_GRID_OFFSET : VRESTIR_TEMPERATURE_GRID_ID;
    float3 pm = WorldToMediumP(pW, slot);
    return ConvertTempToColor(getValueAtPoint(c_scene.slots[slot], pm, true));
}
VRD float3 VelocityWorld(float3 pW, bool isLastFrame = false) {
    int slot = VRESTIR_VELOCITY_GRID_ID + (isLastFrame ? VRESTIR_PREV_EXTRA_GRID_OFFSET : 0);
    const DSlot& g = c_scene.slots[slot];
    float3 pm = WorldToMediumP(pW, slot);
    float3 v = f3(getValueAtPoint(g, pm, true, 0), getValueAtPoint(g, pm, true, 1), getValueAtPoint(g, pm, true, 2));
    return mulVec(v, c_scene.vol.externalModelToWorld);
}

VRD void CoordinateSystem(float3 v1, float3& v2, float3& v3_) {
    if (fabsf(v1.x) > fabsf(v1.y)) v2 = f3(-v1.z, 0, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else v2 = f3(0, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v3_ = cross(v1, v2);
}
VRD float PhaseHG(float cosTheta, float g) {
    float denom = 1 + g * g + 2 * g * cosTheta;
    const float Inv4Pi = 0.07957747154594766788444188168626f;
    return Inv4Pi * (1 - g * g) / (denom * sqrtf(denom));
}
struct MediumInteraction {
    float3 p, wo; float g; bool isValid;
    VRD float phaseFunction(float3 wo_, float3 wi) const { return PhaseHG(dot(wo_, wi), g); }
    VRD float Sample_p(float3 wo_, float3& wi, float2 u) const {
        float cosTheta;
        if (fabsf(g) < 1e-3f) cosTheta = 1 - 2 * u.x;
        else { float sqrTerm = (1 - g * g) / (1 + g - 2 * g * u.x); cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g); }
        float sinTheta = sqrtf(fmaxf(0.f, 1 - cosTheta * cosTheta));
        float phi = 2 * kPi * u.y;
        float3 v1, v2; CoordinateSystem(wo_, v1, v2);
        wi = sinTheta * cosf(phi) * v1 + sinTheta * sinf(phi) * v2 + cosTheta * wo_;
        return PhaseHG(cosTheta, g);
    }
};
VRD MediumInteraction makeMI(float3 p, float3 wo, bool valid) { MediumInteraction m; m.p = p; m.wo = wo; m.g = c_scene.vol.PhaseFunctionConstantG; m.isValid = valid; return m; }
VRD Ray makeRay(float3 o, float3 d, float tmin, float tmax) { Ray r; r.origin = o; r.dir = d; r.tMin = tmin; r.tMax = tmax; return r; }

// ------------------------------------------------------------------------------------------------ HDDA
struct HDDAState {
    // invDir = 1/dir once per ray.  Node spans are powers of two (8 / 128 voxels per child, 1 inside a brick), so
    // |vdel * invDir| and (x - vmin) * (1/vdel) are bit-identical to the reference's |vdel / dir| and (x - vmin) / vdel
    // (F/Scene/GVDB/gvdbDda.slang:121-135) while saving six IEEE divisions per level change.
    float3 pos, dir, invDir; float3 tDel; float tx, ty; int3 p; float3 tSide; int3 mask;
    VRD void SetFromRay(float3 startPos, float3 startDir, float t0) {
        pos = startPos; dir = startDir;
        invDir = make_float3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        tx = t0; ty = 0.f;
    }
    VRD float3 stepSign() const { return make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f); }
    VRD void Prepare(float3 vmin, float vdel, float invVdel) {
        tDel = make_float3(fabsf(vdel * invDir.x), fabsf(vdel * invDir.y), fabsf(vdel * invDir.z));
        float3 pFlt = (pos + tx * dir - vmin) * invVdel;
        float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        tSide = ((fl - pFlt + f3(0.5f)) * stepSign() + f3(0.5f)) * tDel + f3(tx);
        p = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
    }
    VRD void PrepareLeaf(float3 vmin) {
        tDel = make_float3(fabsf(invDir.x), fabsf(invDir.y), fabsf(invDir.z));
        float3 pFlt = pos + tx * dir - vmin;
        float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        tSide = ((fl - pFlt + f3(0.5f)) * stepSign() + f3(0.5f)) * tDel + f3(tx);
        p = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
    }
    VRD void Next() {
        mask.x = int((tSide.x < tSide.y) & (tSide.x <= tSide.z));
        mask.y = int((tSide.y < tSide.z) & (tSide.y <= tSide.x));
        mask.z = int((tSide.z < tSide.x) & (tSide.z <= tSide.y));
        ty = mask.x ? tSide.x : (mask.y ? tSide.y : tSide.z);
    }
    VRD void Step() {
        tx = ty;
        // select instead of gvdbDda.slang:153's mask*tDel: identical for finite tDel, and an exactly-zero direction component
        // (tDel = +inf) no longer turns tSide into NaN (0*inf), which made the traversal spin to the 4096-iteration cap
        tSide = make_float3(mask.x ? tSide.x + tDel.x : tSide.x, mask.y ? tSide.y + tDel.y : tSide.y, mask.z ? tSide.z + tDel.z : tSide.z);
        p = make_int3(p.x + (mask.x ? (dir.x >= 0 ? 1 : -1) : 0), p.y + (mask.y ? (dir.y >= 0 ? 1 : -1) : 0), p.z + (mask.z ? (dir.z >= 0 ? 1 : -1) : 0));
    }
};
VRD bool inRange(int3 p, int hiExclusive) { return (unsigned)p.x < (unsigned)hiExclusive && (unsigned)p.y < (unsigned)hiExclusive && (unsigned)p.z < (unsigned)hiExclusive; }

// VR/VolumeUtils.slang:171-282.  The adapter sees (dda, brick min corner, brick id, leaf node id).
// Per-level state (only levels 1 and 2 exist) lives in scalar registers, not in dynamically indexed local arrays.
template <class Adapter>
__device__ void VolumeTrackingGVDB(const Ray& rWorld, int mipLevel, SampleGenerator& sg, Adapter& adapter, bool vertexCenter) {
    const DSlot& g = c_scene.slots[mipLevel];
    const float epsilon = 0.01f;
    int lev = g.top_lev;
    const int topLev = lev;
    Ray ray = WorldToMedium(rWorld, mipLevel);
    if (vertexCenter) ray.origin = ray.origin - f3(0.5f);
    float tNear, tFar;
    if (!IntersectVolumeBound(ray, tNear, tFar, mipLevel, vertexCenter)) { adapter.ExecuteEndStep(); return; }
    adapter.SetRayInfo(tNear, tFar, ray);
    adapter.ExecuteStartStep();
    const int res1 = g.res[1], res2 = g.res[2], dim1 = g.dim[1], dim2 = g.dim[2];
    const float vdel1 = g.vdel[1], vdel2 = g.vdel[2], ivdel1 = 1.0f / vdel1, ivdel2 = 1.0f / vdel2;
    const unsigned cnt1 = g.childCount32[1], cnt2 = g.childCount32[2];
    uint32_t link1 = ID_UNDEFL, link2 = ID_UNDEFL; float3 vmin1 = f3(0.f), vmin2 = f3(0.f); float tMax1 = 0.f, tMax2 = 0.f;
    {
        NodeHead h = loadNodeHead(g, lev, 0);
        if (lev == 2) { link2 = (uint32_t)h.a.w; vmin2 = nodePos(h.a); tMax2 = tFar; }
        else { link1 = (uint32_t)h.a.w; vmin1 = nodePos(h.a); tMax1 = tFar; }
    }
    int iter = 0;
    HDDAState dda;
    dda.SetFromRay(ray.origin, ray.dir, tNear + epsilon);
    if (lev == 2) dda.Prepare(vmin2, vdel2, ivdel2); else dda.Prepare(vmin1, vdel1, ivdel1);
    if (vertexCenter) {
        const int r = lev == 2 ? res2 : res1;
        int it = 0;
        while (it++ < 3 && (dda.p.x < 0 || dda.p.y < 0 || dda.p.z < 0 || dda.p.x > r || dda.p.y > r || dda.p.z > r)) {
            dda.Next(); dda.Step(); dda.tx += epsilon;
        }
    }
    for (; iter < 4096 && lev > 0 && lev <= topLev && inRange(dda.p, (lev == 2 ? res2 : res1) + 1); iter++) {
        dda.Next();
        const int dm = lev == 2 ? dim2 : dim1;
        const int b = (((dda.p.z << dm) + dda.p.y) << dm) + dda.p.x;
        uint32_t childNodeId;
        {
            const uint32_t listid = lev == 2 ? link2 : link1;
            if (listid == ID_UNDEFL) childNodeId = ID_UNDEFL;
            else {
                // p == res passes the inclusive bound (VR/VolumeUtils.slang:231) and aliases into the list like the shader's
                // ByteAddressBuffer load; outside the whole list D3D returns 0
                const int r = lev == 2 ? res2 : res1;
                const long long idx = (long long)listid * (long long)(r * r * r) + (long long)b;
                childNodeId = (idx < 0 || idx >= (long long)(lev == 2 ? cnt2 : cnt1)) ? 0u : __ldg(&g.child[lev][idx]);
            }
        }
        if (childNodeId != ID_UNDEFL) {
            if (lev == 1) {
                float t = dda.tx - epsilon;
                NodeHead leaf = loadNodeHead(g, 0, childNodeId);
                bool shouldExit = adapter.ExecuteMainStep(g, dda, nodePos(leaf.a), (uint32_t)leaf.a.w, childNodeId, t, sg);
                if (shouldExit) return;
                dda.Step();
                dda.tx += epsilon;
            } else {
                lev = 1;
                NodeHead h = loadNodeHead(g, 1, childNodeId);
                link1 = (uint32_t)h.a.w; vmin1 = nodePos(h.a);
                tMax1 = dda.ty;
                dda.Prepare(vmin1, vdel1, ivdel1);
            }
        } else {
            dda.Step();
            dda.tx += epsilon;
        }
        while (lev <= topLev && dda.tx > (lev == 2 ? tMax2 : tMax1)) {
            lev++;
            if (lev <= topLev) dda.Prepare(vmin2, vdel2, ivdel2);
        }
    }
    if (iter >= 1024) {
        unsigned k = atomicAdd(&g_dbgCount, 1u);
        if (k < 64) {
            float* o = &g_dbgRays[k * 8];
            o[0] = rWorld.origin.x; o[1] = rWorld.origin.y; o[2] = rWorld.origin.z; o[3] = rWorld.dir.x; o[4] = rWorld.dir.y; o[5] = rWorld.dir.z;
            o[6] = (float)(mipLevel + (vertexCenter ? 100 : 0)); o[7] = (float)iter;
        }
    }
    adapter.ExecuteEndStep();
}

VRD void trilinearCubic(const float v[8], float sigma_t, float3 d, float3 p0, float& c3, float& c2, float& c1, float& c0) {
    float v_000 = v[0] * sigma_t, v_100 = v[1] * sigma_t, v_010 = v[2] * sigma_t, v_110 = v[3] * sigma_t;
    float v_001 = v[4] * sigma_t, v_101 = v[5] * sigma_t, v_011 = v[6] * sigma_t, v_111 = v[7] * sigma_t;
    float mxyz = v_111 - v_011 - v_101 - v_110 + v_100 + v_010 + v_001 - v_000;
    float mxy = v_000 - v_100 - v_010 + v_110;
    float mxz = v_000 - v_100 - v_001 + v_101;
    float myz = v_000 - v_010 - v_001 + v_011;
    float mx = v_100 - v_000, my = v_010 - v_000, mz = v_001 - v_000;
    c3 = mxyz * d.x * d.y * d.z;
    c2 = (p0.z * d.x * d.y + p0.y * d.x * d.z + p0.x * d.y * d.z) * mxyz + mxy * d.x * d.y + mxz * d.x * d.z + myz * d.y * d.z;
    c1 = (p0.y * p0.z * d.x + p0.x * p0.z * d.y + p0.x * p0.y * d.z) * mxyz + mx * d.x + my * d.y + mz * d.z +
         (p0.y * d.x + p0.x * d.y) * mxy + (p0.z * d.x + p0.x * d.z) * mxz + (p0.z * d.y + p0.y * d.z) * myz;
    c0 = p0.x * p0.y * p0.z * mxyz + p0.x * p0.y * mxy + p0.x * p0.z * mxz + p0.y * p0.z * myz + p0.x * mx + p0.y * my + p0.z * mz + v_000;
}

struct AdapterBase {
    float tNear, tFar; Ray ray;
    VRD void SetRayInfo(float a, float b, const Ray& r) { tNear = a; tFar = b; ray = r; }
};

// VR/VolumeTrackingAdapterGVDB.slang:20-136
struct MediumTrAnalyticAdapter : AdapterBase {
    float Tr; bool useLinearSampler;
    VRD void Init(bool linear) { Tr = 0.f; useLinearSampler = linear; }
    VRD void ExecuteStartStep() {}
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator&) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        const float sigma_t_ = c_scene.vol.sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, g.res[0]); iter++) {
            leaf.Next();
            float maxDeltaT = leaf.ty - t;
            if (useLinearSampler) {
                float v[8]; FetchEightVoxels(g, brick, leaf.p, v);
                float3 d = ray.dir;
                float3 p0 = leaf.pos + leaf.tx * leaf.dir - (make_float3((float)leaf.p.x, (float)leaf.p.y, (float)leaf.p.z) + vmin_leaf);
                float c3, c2, c1, c0; trilinearCubic(v, sigma_t_, d, p0, c3, c2, c1, c0);
                float t_dist = fminf(tFar - t, maxDeltaT);
                float t2 = t_dist * t_dist, t3 = t2 * t_dist, t4 = t2 * t2;
                Tr += -(c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t_dist);
            } else {
                float density = DensityInAtlas<true>(g, brick, make_float3((float)leaf.p.x, (float)leaf.p.y, (float)leaf.p.z) + f3(0.5f), false);
                float sigma_t = density * sigma_t_;
                Tr += -fminf(tFar - t, maxDeltaT) * sigma_t;
            }
            if (t + maxDeltaT >= tFar) { Tr = expf(Tr); return true; }
            t += maxDeltaT;
            leaf.Step();
        }
        return false;
    }
    VRD void ExecuteEndStep() { Tr = expf(Tr); }
};

// VR/VolumeTrackingAdapterGVDB.slang:140-208
struct MediumTrRayMarchingAdapter : AdapterBase {
    float Tr; bool useLinearSampler; float tStep; bool hasInitialized;
    VRD void Init(bool linear, float step) { Tr = 0.f; useLinearSampler = linear; tStep = step; hasInitialized = false; }
    VRD void ExecuteStartStep() { hasInitialized = true; }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator&) {
        t = tNear + (floorf((t - tNear) / tStep) + 0.5f) * tStep;
        if (t < dda.tx) t += tStep;
        const float tStepMultipler = 1.f;
        float3 wp = ray.origin + t * ray.dir;
        float3 p = wp - vmin_leaf;
        const float3 wpt = tStepMultipler * tStep * ray.dir;
        const float res = (float)g.res[0];
        const float sig = c_scene.vol.sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res && p.y < res && p.z < res; iter++) {
            if (t >= tFar) { Tr = expf(Tr); return true; }
            float density = DensityInAtlas<false>(g, brick, p, useLinearSampler);
            float sigma_t = density * sig;
            Tr += -sigma_t * (iter == 0 ? 1.f : tStepMultipler) * tStep;
            p = p + wpt;
            t += tStepMultipler * tStep;
        }
        return false;
    }
    VRD void ExecuteEndStep() { if (hasInitialized) Tr = expf(Tr); else Tr = 1.f; }
};

// Shared camera march: one traversal of a ray yields the ray-marched transmittance to up to 3 depths.
// Bit-identical to running MediumTrRayMarchingAdapter once per depth (same global sample phase, same partial sums in the
// same order; a depth's value is the running sum at the first sample with t >= min(boxFar, depth), or at the end of the
// ray).  Used by the task-parallel reuse kernels, where several stored samples are evaluated along the same pixel ray
// (VR/SpatialReuse.cs.slang:183-239 evaluates each tap at every other tap's ray).
struct MultiDepthRayMarchingAdapter : AdapterBase {
    float Tr; bool useLinearSampler; float tStep; bool hasInitialized;
    float thr[3], out[3]; int n; unsigned pending;
    VRD void Init(bool linear, float step, const float* depths, int count) {
        Tr = 0.f; useLinearSampler = linear; tStep = step; hasInitialized = false; n = count; pending = (1u << count) - 1u;
#pragma unroll
        for (int k = 0; k < 3; k++) { thr[k] = k < count ? depths[k] : 0.f; out[k] = 0.f; }
    }
    VRD void ExecuteStartStep() { hasInitialized = true; }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator&) {
        t = tNear + (floorf((t - tNear) / tStep) + 0.5f) * tStep;
        if (t < dda.tx) t += tStep;
        float3 wp = ray.origin + t * ray.dir;
        float3 p = wp - vmin_leaf;
        const float3 wpt = 1.f * tStep * ray.dir;
        const float res = (float)g.res[0];
        const float sig = c_scene.vol.sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res && p.y < res && p.z < res; iter++) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                if ((pending >> k) & 1u) { if (t >= fminf(tFar, thr[k])) { out[k] = Tr; pending &= ~(1u << k); } }
            if (!pending) return true;
            float density = DensityInAtlas<false>(g, brick, p, useLinearSampler);
            float sigma_t = density * sig;
            Tr += -sigma_t * 1.f * tStep;
            p = p + wpt;
            t += 1.f * tStep;
        }
        return false;
    }
    VRD void ExecuteEndStep() {
#pragma unroll
        for (int k = 0; k < 3; k++) if ((pending >> k) & 1u) out[k] = Tr;
        pending = 0;
    }
    VRD float result(int k) const { return hasInitialized ? expf(out[k]) : 1.f; }
};

// VR/VolumeTrackingAdapterGVDB.slang:212-436
struct SampleMediumAnalyticAdapter : AdapterBase {
    float hitDistances[4], outTr[4], pdf[4]; int numSamples; float opticalThickness; bool hasInitialized, useLinearSampler;
    VRD void Init(int n, bool linear) {
        numSamples = n; useLinearSampler = linear; opticalThickness = 0.f; hasInitialized = false;
#pragma unroll
        for (int i = 0; i < 4; i++) { hitDistances[i] = 0.f; outTr[i] = 0.f; pdf[i] = 0.f; }
    }
    VRD void ExecuteStartStep() {
#pragma unroll
        for (int i = 0; i < 4; i++) if (i < numSamples) hitDistances[i] = -1;
        hasInitialized = true;
    }
    static VRD float tauOf(float t, float c3, float c2, float c1, float c0) { float t2 = t * t, t3 = t2 * t, t4 = t2 * t2; return c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t; }
    static VRD float sigmaOf(float t, float c3, float c2, float c1, float c0) { float t2 = t * t, t3 = t2 * t; return c3 * t3 + c2 * t2 + c1 * t + c0; }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator& sg) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        const float sigma_t_ = c_scene.vol.sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, g.res[0]); iter++) {
            leaf.Next();
            float maxDeltaT = fminf(tFar - t, leaf.ty - t);
            float currentTMax = fminf(tFar, leaf.ty);
            int finishedCount = 0;
            float deltaThickness = 0.f;
            if (useLinearSampler) {
                float v[8]; FetchEightVoxels(g, brick, leaf.p, v);
                float3 d = ray.dir;
                float3 p0 = leaf.pos + leaf.tx * leaf.dir - (make_float3((float)leaf.p.x, (float)leaf.p.y, (float)leaf.p.z) + vmin_leaf);
                float c3, c2, c1, c0; trilinearCubic(v, sigma_t_, d, p0, c3, c2, c1, c0);
                deltaThickness = tauOf(maxDeltaT, c3, c2, c1, c0);
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    if (s >= numSamples) break;
                    if (hitDistances[s] == -1) {
                        if (opticalThickness + deltaThickness >= outTr[s]) {
                            float tau_target = outTr[s] - opticalThickness;
                            float t_low = 0, t_high = maxDeltaT, tau_low = 0, tau_high = deltaThickness, t_sol = 0;
                            int it = 0;
                            while (it++ < 32 && t_high - t_low > maxDeltaT * 0.001f) {
                                t_sol = t_low + (t_high - t_low) * (tau_target - tau_low) / (tau_high - tau_low);
                                float tau = tauOf(t_sol, c3, c2, c1, c0);
                                if (tau < tau_target) { t_low = t_sol; tau_low = tau; } else { t_high = t_sol; tau_high = tau; }
                            }
                            hitDistances[s] = t + t_sol;
                            outTr[s] = expf(-outTr[s]);
                            pdf[s] = sigmaOf(t_sol, c3, c2, c1, c0) * outTr[s];
                            finishedCount++;
                        }
                    } else finishedCount++;
                }
            } else {
                float density = DensityInAtlas<true>(g, brick, make_float3((float)leaf.p.x, (float)leaf.p.y, (float)leaf.p.z) + f3(0.5f), false);
                float sigma_t = density * sigma_t_;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    if (s >= numSamples) break;
                    if (hitDistances[s] == -1) {
                        // an empty cell (sigma_t == 0) can never be hit: -log(1-u)/0 is +inf or NaN -> kRayTMax; only the draw counts
                        if (sigma_t == 0.f) { (void)sg.next(); continue; }
                        float dT = -logf(1 - sampleNext1D(sg)) / sigma_t;
                        float curT = t + dT;
                        if (isnan(curT) || isinf(curT)) curT = kRayTMax;
                        if (curT < currentTMax) {
                            hitDistances[s] = curT;
                            outTr[s] = expf(-(dT * sigma_t + opticalThickness));
                            pdf[s] = sigma_t * outTr[s];
                            finishedCount++;
                        }
                    } else finishedCount++;
                }
                deltaThickness = maxDeltaT * sigma_t;
            }
            if (finishedCount == numSamples) return true;
            t = currentTMax;
            opticalThickness += deltaThickness;
            if (t >= tFar) { ExecuteEndStep(); return true; }
            leaf.Step();
        }
        return false;
    }
    VRD void ExecuteEndStep() {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (s >= numSamples) break;
            if (hasInitialized) { if (hitDistances[s] == -1) { hitDistances[s] = kRayTMax; outTr[s] = expf(-opticalThickness); pdf[s] = outTr[s]; } }
            else { hitDistances[s] = kRayTMax; outTr[s] = 1.f; pdf[s] = 1.f; }
        }
    }
};

// VR/VolumeTrackingAdapterGVDB.slang:439-515
struct SampleVolumeCellByDensityAdapter : AdapterBase {
    int densityBound; float2 selectedInterval; float runningSum, Tr;
    VRD void Init() { densityBound = 0; selectedInterval = make_float2(-1, -1); runningSum = 0.f; Tr = 0.f; }
    VRD void ExecuteStartStep() { selectedInterval = make_float2(-1, -1); }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator& sg) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, g.res[0]); iter++) {
            float density = DensityInAtlas<true>(g, brick, make_float3((float)leaf.p.x, (float)leaf.p.y, (float)leaf.p.z) + f3(0.5f), false);
            leaf.Next();
            float maxDeltaT = leaf.ty - t;
            float sigma_t = density * c_scene.vol.sigma_t;
            float weight = expf(Tr) * sigma_t;
            runningSum += weight;
            if (runningSum > 0.f && sampleNext1D(sg) < weight / runningSum) {
                densityBound = f2i(density);
                selectedInterval = make_float2(t, fminf(tFar, t + maxDeltaT));
            }
            Tr += -maxDeltaT * sigma_t;
            if (t + maxDeltaT >= tFar) return true;
            t += maxDeltaT;
            leaf.Step();
        }
        return false;
    }
    VRD void ExecuteEndStep() {}
};

// VR/VolumeTrackingAdapterGVDB.slang:518-602
struct ReservoirFeatureRayMarchingAdapter : AdapterBase {
    float accuTransmittance, Tr, tStep; bool useLinearSampler, hasInitialized;
    VRD void Init(bool linear, float step) { accuTransmittance = 1.f; Tr = 0.f; tStep = step; useLinearSampler = linear; hasInitialized = false; }
    VRD void ExecuteStartStep() { hasInitialized = true; }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t, float& t, SampleGenerator&) {
        t = tNear + (floorf((t - tNear) / tStep) + 0.5f) * tStep;
        if (t < dda.tx) t += tStep;
        float3 wp = ray.origin + t * ray.dir;
        float3 p = wp - vmin_leaf;
        const float3 wpt = tStep * ray.dir;
        const float res = (float)g.res[0];
        for (int iter = 0; iter < MAX_BRICK_STEPS && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res && p.y < res && p.z < res; iter++) {
            float curTransmittance = expf(Tr);
            if (curTransmittance < 0.01f) { ExecuteEndStep(); return true; }
            float density = DensityInAtlas<false>(g, brick, p, useLinearSampler);
            float sigma_t = density * c_scene.vol.sigma_t;
            float negOpticalLength = -sigma_t * tStep;
            Tr += negOpticalLength;
            p = p + wpt;
            t += tStep;
        }
        return false;
    }
    VRD void ExecuteEndStep() { if (hasInitialized) accuTransmittance = expf(Tr); }
};

// VR/VolumeTrackingAdapterGVDB.slang:606-707
struct DecompositionTrackingAdapter : AdapterBase {
    Ray rWorld; MediumInteraction mi;
    VRD void ExecuteStartStep() {}
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t leafId, float& t, SampleGenerator& sg) {
        const float sigma_t_ = c_scene.vol.sigma_t;
        const float4 bnd = loadNodeBounds(g, 0, leafId);
        const float s = c_scene.vol.densityScaleFactorByScaling;
        float minDensity = bnd.x * s, maxDensity = bnd.y * s;
        float currentTMax = fminf(tFar, dda.ty);
        float t_control;
        if (minDensity == 0.f) t_control = kRayTMax;
        else t_control = t - logf(1 - sampleNext1D(sg)) / (minDensity * sigma_t_);
        float invMaxDensity = 1.f / (maxDensity - minDensity);
        if (maxDensity - minDensity > 0.f) {
            while (true) {
                t -= logf(1 - sampleNext1D(sg)) * invMaxDensity / sigma_t_;
                if (t >= t_control || t >= currentTMax) break;
                float3 wp = ray.origin + t * ray.dir;
                float3 p = wp - vmin_leaf;
                float density = DensityInAtlas<true>(g, brick, p, true);
                if ((density - minDensity) * invMaxDensity > sampleNext1D(sg)) { mi = makeMI(rWorld.at(t), -rWorld.dir, true); return true; }
            }
            t = fminf(t_control, t);
            if (t < currentTMax) { mi = makeMI(rWorld.at(t), -rWorld.dir, true); return true; }
            t = currentTMax;
            if (t >= tFar) { ExecuteEndStep(); return true; }
        } else {
            if (t_control < currentTMax) { mi = makeMI(rWorld.at(t_control), -rWorld.dir, true); return true; }
            else { t = currentTMax; return false; }
        }
        return false;
    }
    VRD void ExecuteEndStep() { mi.isValid = false; }
};

// VR/VolumeTrackingAdapterGVDB.slang:710-794
struct ResidualRatioTrackingAdapter : AdapterBase {
    float Tr; bool useAnalogResidual, useGlobalMajorant; int mip;
    VRD void Init(bool analog, bool global, int mip_) { Tr = 1.f; useGlobalMajorant = global; useAnalogResidual = analog || global; mip = mip_; }
    VRD void ExecuteStartStep() { if (useGlobalMajorant) useAnalogResidual = true; }
    VRD bool ExecuteMainStep(const DSlot& g, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, uint32_t leafId, float& t, SampleGenerator& sg) {
        const float sigma_t_ = c_scene.vol.sigma_t;
        const float4 bnd = loadNodeBounds(g, 0, leafId);
        const float s = c_scene.vol.densityScaleFactorByScaling;
        float mu_min = useGlobalMajorant ? 0.f : (bnd.x * s) * sigma_t_;
        float mu_max = useGlobalMajorant ? GetVolumeMaxDensity(mip) * sigma_t_ : (bnd.y * s) * sigma_t_;
        float mu_avg = (bnd.z * s) * sigma_t_;
        float maxDeltaT = fminf(tFar - t, dda.ty - t);
        float currentTMax = fminf(tFar, dda.ty);
        float mu_r_temp = mu_max - mu_min;
        float D = c_scene.vol.superVoxelWorldSpaceDiagonalLength;
        // gamma = 2: pow(gamma, x) through double exp2 and one rounding, i.e. correctly rounded like the host's powf.  An ulp in mu_c
        // decides whether a collision in a saturated voxel (mu == mu_max) multiplies T_r by exactly 0 or by 6e-8, and with it whether
        // later segments of a p-hat are evaluated and how many numbers the pixel draws - CUDA's powf is within 2 ulp
        float mu_c = (mu_r_temp == 0.f || useAnalogResidual) ? mu_min : fminf(mu_avg, fmaxf(mu_min, mu_min + mu_r_temp * ((float)exp2((double)(1.f / (D * mu_r_temp))) - 1)));
        float mu_r = fmaxf(mu_c - mu_min, mu_max - mu_c);
        float inv_mu_r = 1.f / mu_r;
        float T_c = expf(-mu_c * fminf(tFar - t, maxDeltaT));
        float T_r = 1;
        if (mu_r > 0.f) {
            while (true) {
                t -= logf(1 - sampleNext1D(sg)) * inv_mu_r;
                if (t >= currentTMax) break;
                float3 wp = ray.origin + t * ray.dir;
                float3 p = wp - vmin_leaf;
                float density = DensityInAtlas<true>(g, brick, p, true);
                float mu = density * sigma_t_;
                T_r *= 1 - (mu - mu_c) * inv_mu_r;
            }
        }
        Tr *= T_c * T_r;
        t = currentTMax;
        if (t >= tFar) return true;
        return false;
    }
    VRD void ExecuteEndStep() {}
};

// ---- generic wrappers (VR/VolumeUtils.slang:284-378); __noinline__ keeps one copy of each traversal per kernel ----
VRD_NOINLINE float MediumTrAnalyticGeneric(const Ray& r, int mip, SampleGenerator& sg, bool linear) {
    MediumTrAnalyticAdapter a; a.Init(linear);
    VolumeTrackingGVDB(r, mip, sg, a, linear);
    return a.Tr;
}
VRD_NOINLINE float MediumTrResidualRatioTrackingGeneric(const Ray& r, int mip, SampleGenerator& sg, bool analog, bool global) {
    ResidualRatioTrackingAdapter a; a.Init(analog, global, mip);
    VolumeTrackingGVDB(r, mip, sg, a, false);
    return a.Tr;
}
VRD_NOINLINE float MediumTrRayMarchingGeneric(const Ray& r, int mip, bool linear, float tStepScale) {
    int eff = mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : mip;
    eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
    MediumTrRayMarchingAdapter a;
    a.Init(linear, c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * tStepScale * (eff + 1));
    SampleGenerator dummy; dummy.s0 = dummy.s1 = dummy.s2 = dummy.s3 = 0;
    VolumeTrackingGVDB(r, mip, dummy, a, false);
    return a.Tr;
}
// ray.tMax must be the largest of `depths`
VRD_NOINLINE void MediumTrRayMarchingMulti(const Ray& r, int mip, bool linear, float tStepScale, const float* depths, int count, float* outTr) {
    int eff = mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : mip;
    eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
    MultiDepthRayMarchingAdapter a;
    a.Init(linear, c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * tStepScale * (eff + 1), depths, count);
    SampleGenerator dummy; dummy.s0 = dummy.s1 = dummy.s2 = dummy.s3 = 0;
    VolumeTrackingGVDB(r, mip, dummy, a, false);
    for (int k = 0; k < count; k++) outTr[k] = a.result(k);
}
VRD_NOINLINE void SampleMediumAnalyticGeneric(const Ray& r, SampleGenerator& sg, bool linear, float hit[4], int mip, float pdf[4], float outTr[4], int numSamples) {
    SampleMediumAnalyticAdapter a; a.Init(numSamples, linear);
    if (linear) for (int i = 0; i < numSamples; i++) a.outTr[i] = -logf(1 - sampleNext1D(sg));
    VolumeTrackingGVDB(r, mip, sg, a, linear);
#pragma unroll
    for (int i = 0; i < 4; i++) { hit[i] = a.hitDistances[i]; outTr[i] = a.outTr[i]; pdf[i] = a.pdf[i]; }
}
VRD_NOINLINE void SampleMediumSuperVoxelGeneric(const Ray& r, SampleGenerator& sg, MediumInteraction& mi, int mip) {
    DecompositionTrackingAdapter a; a.rWorld = r; a.mi = mi;
    VolumeTrackingGVDB(r, mip, sg, a, false);
    mi = a.mi;
}
VRD float computeVisibility(const Ray& ray, SampleGenerator& sg, int visibilitySamples, int mip, bool linear, uint32_t method, float tStepScale) {
    float visibility = 0;
    for (int i = 0; i < visibilitySamples; i++) {
        if (method == VRESTIR_RAY_MARCHING) visibility += MediumTrRayMarchingGeneric(ray, mip, linear, tStepScale);
        else if (method == VRESTIR_ANALYTIC_TRACKING) visibility += MediumTrAnalyticGeneric(ray, mip, sg, linear);
        else if (method == VRESTIR_RATIO_TRACKING || method == VRESTIR_RESIDUAL_RATIO_TRACKING || method == VRESTIR_ANALOG_RESIDUAL_RATIO_TRACKING)
            visibility += MediumTrResidualRatioTrackingGeneric(ray, mip, sg, method == VRESTIR_ANALOG_RESIDUAL_RATIO_TRACKING, method == VRESTIR_RATIO_TRACKING);
    }
    visibility /= visibilitySamples;
    return visibility;
}
VRD_NOINLINE float RejectionSampleRandomPointByDensity(const Ray& r, SampleGenerator& sg, int mip) {
    SampleVolumeCellByDensityAdapter a; a.Init();
    VolumeTrackingGVDB(r, mip, sg, a, false);
    float2 sel = a.selectedInterval;
    if (sel.x == -1) return kRayTMax;
    float sampledDepth = sel.x + (sel.y - sel.x) * sampleNext1D(sg);
    (void)sampleNext1D(sg);   // sampledY draw (VR/VolumeUtils.slang:580), value unused
    return sampledDepth;
}

// ------------------------------------------------------------------------------------------------ lights
VRD float2 world_to_latlong_map(float3 dir) {
    float3 p = normalize(dir);
    float2 uv; uv.x = atan2f(p.x, -p.z) * k1_2Pi + 0.5f; uv.y = acosf(p.y) * k1_Pi;
    return uv;
}
VRD float3 oct_to_ndir_equal_area_unorm(float2 p) {
    p.x = p.x * 2.f - 1.f; p.y = p.y * 2.f - 1.f;
    float d = 1.f - (fabsf(p.x) + fabsf(p.y));
    float r = 1.f - fabsf(d);
    float phi = (r > 0.f) ? ((fabsf(p.y) - fabsf(p.x)) / r + 1.f) * kPi_4 : 0.f;
    float f = r * sqrtf(2.f - r * r);
    float x = f * fsign(p.x) * cosf(phi);
    float y = f * fsign(p.y) * sinf(phi);
    float z = fsign(d) * (1.f - r * r);
    return f3(x, y, z);
}
VRD float3 envTexel(int x, int y) { float4 t = __ldg(&c_scene.envTexels[(size_t)y * c_scene.envW + x]); return f3(t.x, t.y, t.z); }
VRD float3 envBilinear(float2 uv) {
    const int W = c_scene.envW, H = c_scene.envH;
    float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = f2i(x0f), y0 = f2i(y0f), x1 = x0 + 1, y1 = y0 + 1;
    x0 %= W; if (x0 < 0) x0 += W;
    x1 %= W; if (x1 < 0) x1 += W;
    y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
    float3 a = envTexel(x0, y0), b = envTexel(x1, y0), c = envTexel(x0, y1), d = envTexel(x1, y1);
    float3 top = f3(lerpf(a.x, b.x, fx), lerpf(a.y, b.y, fx), lerpf(a.z, b.z, fx));
    float3 bot = f3(lerpf(c.x, d.x, fx), lerpf(c.y, d.y, fx), lerpf(c.z, d.z, fx));
    return f3(lerpf(top.x, bot.x, fy), lerpf(top.y, bot.y, fy), lerpf(top.z, bot.z, fy));
}
VRD float3 envToLocal(float3 dir, bool last = false) { return mulVec3x3(dir, last ? c_scene.envPrevInvT : c_scene.envInvT); }
VRD float3 envToWorld(float3 dir, bool last = false) { return mulVec3x3(dir, last ? c_scene.envPrevT : c_scene.envT); }
VRD float3 envEval(float3 dir, bool last = false) {
    if (!c_scene.haveEnv) return f3(0.f);
    float2 uv = world_to_latlong_map(envToLocal(dir, last));
    return c_scene.envIntensity * c_scene.envTint * envBilinear(uv);
}
VRD float2 envEncodeLightUV(float3 dir, int& lightID) { dir = envToLocal(dir); lightID = dir.z < 0 ? -2 : -1; return make_float2(dir.x, dir.y); }
VRD float3 envDecodeLightUV(float2 uv, int lightID, bool last) {
    float3 dir = f3(uv.x, uv.y, 0);
    dir.z = sqrtf(1 - uv.x * uv.x - uv.y * uv.y);
    if (isnan(dir.z)) dir.z = 0.f;
    if (lightID == -2) dir.z = -dir.z;
    return envToWorld(dir, last);
}
VRD float impLoad(uint32_t x, uint32_t y, int mip) {
    int dim = c_scene.impDim >> mip;
    if ((int)x >= dim || (int)y >= dim) return 0.f;
    return __ldg(&c_scene.importance[c_scene.impOffset[mip] + (size_t)y * dim + x]);
}
// Top of the importance mip chain (mips with dim <= 32: 32^2 + 16^2 + ... + 1 = 1365 floats) staged in shared memory by
// kernels whose critical path is the 36-load dependent chain of the hierarchical warp (K1 step kernels).
constexpr int IMP_TOP_DIM = 32, IMP_TOP_FLOATS = 1365;
constexpr int IMP_TOP_BYTES = (IMP_TOP_FLOATS * 4 + 15) / 16 * 16;   // bulk copies move multiples of 16 B (the map is allocated with the slack)
// The 5.5 KB block is one contiguous piece of the importance chain: a single thread hands it to the bulk-copy engine
// (cp.async.bulk, global -> shared, completion on an mbarrier) and the CTA waits on the barrier — no per-thread load / store
// loop, no __syncthreads.  smem must be 16-byte aligned; every thread of the CTA calls and returns with the data visible.
VRD void stageImportanceTop(float* smem, uint64_t* bar) {
    if (c_scene.impDim < IMP_TOP_DIM) return;
    int first = 0; while ((c_scene.impDim >> first) > IMP_TOP_DIM) first++;
    const float* src = c_scene.importance + c_scene.impOffset[first];
    const unsigned barAddr = (unsigned)__cvta_generic_to_shared(bar), dstAddr = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // the barrier is initialised before anybody polls it
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"((unsigned)IMP_TOP_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dstAddr), "l"(src), "r"((unsigned)IMP_TOP_BYTES), "r"(barAddr) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(barAddr) : "memory");
}
VRD float impLoadT(uint32_t x, uint32_t y, int mip, const float* impTop) {
    const int dim = c_scene.impDim >> mip;
    if ((int)x >= dim || (int)y >= dim) return 0.f;
    if (impTop && dim <= IMP_TOP_DIM) {
        int first = 0; while ((c_scene.impDim >> first) > IMP_TOP_DIM) first++;
        return impTop[c_scene.impOffset[mip] - c_scene.impOffset[first] + (size_t)y * dim + x];
    }
    return __ldg(&c_scene.importance[c_scene.impOffset[mip] + (size_t)y * dim + x]);
}
struct EnvMapSample { float3 dir; float pdf; float3 Le; };
VRD_NOINLINE void envSample(float2 rnd, EnvMapSample& result, const float* impTop = nullptr) {
    float2 pp = rnd; uint32_t posx = 0, posy = 0;
    if (c_scene.envSamplerType == VRESTIR_ENV_SAMPLER_ALIAS && c_scene.envAliasCount) {
        const uint32_t count = c_scene.envAliasCount;
        float xs = rnd.x * (float)count;
        uint32_t index = min(count - 1, f2u(xs));
        float xi = xs - (float)index;
        float thr = __ldg(&c_scene.envAliasThr[index]);
        uint32_t texel;
        if (xi < thr) { texel = index; pp.x = xi / thr; } else { texel = __ldg(&c_scene.envAliasRedirect[index]); pp.x = (xi - thr) / (1.f - thr); }
        posx = texel % (uint32_t)c_scene.impDim; posy = texel / (uint32_t)c_scene.impDim;
    } else {
        for (int mip = c_scene.impBaseMip - 1; mip >= 0; mip--) {
            posx *= 2; posy *= 2;
            float w0 = impLoadT(posx, posy, mip, impTop), w1 = impLoadT(posx + 1, posy, mip, impTop), w2 = impLoadT(posx, posy + 1, mip, impTop), w3 = impLoadT(posx + 1, posy + 1, mip, impTop);
            float q0 = w0 + w2, q1 = w1 + w3;
            uint32_t offx, offy;
            float d = q0 / (q0 + q1);
            if (pp.x < d) { offx = 0; pp.x = pp.x / d; } else { offx = 1; pp.x = (pp.x - d) / (1.f - d); }
            float e = (offx ? w1 : w0) / (offx ? q1 : q0);
            if (pp.y < e) { offy = 0; pp.y = pp.y / e; } else { offy = 1; pp.y = (pp.y - e) / (1.f - e); }
            posx += offx; posy += offy;
        }
    }
    float invDim = 1.f / (float)c_scene.impDim;
    float2 uv = make_float2(((float)posx + pp.x) * invDim, ((float)posy + pp.y) * invDim);
    float3 dir = oct_to_ndir_equal_area_unorm(uv);
    float avg_w = impLoadT(0, 0, c_scene.impBaseMip, impTop);
    float pdf = impLoad(posx, posy, 0) / avg_w;
    result.dir = envToWorld(dir);
    result.pdf = pdf * k1_4Pi;
    result.Le = envEval(result.dir);
}

struct SceneLightSample { float3 dir; float distance; float3 Li; float pdf, pdfArea; float3 rayDir; float rayDistance; };
struct AnalyticLightSample { float3 dir; float distance; float3 Li; float pdf; };
VRD bool sampleLight(float3 shadingPosW, const vrestir_light* lp, AnalyticLightSample& ls) {
    const float kMinLightDistSqr = 1e-9f;
    const uint32_t type = __ldg(&lp->type);
    float3 I = f3(__ldg(&lp->intensity[0]), __ldg(&lp->intensity[1]), __ldg(&lp->intensity[2]));
    if (type == VRESTIR_LIGHT_POINT) {
        float3 posW = f3(__ldg(&lp->posW[0]), __ldg(&lp->posW[1]), __ldg(&lp->posW[2]));
        float3 toLight = posW - shadingPosW;
        float distSqr = fmaxf(dot(toLight, toLight), kMinLightDistSqr);
        ls.distance = sqrtf(distSqr);
        ls.dir = toLight / ls.distance;
        ls.Li = I / distSqr;
        ls.pdf = 0.f;
        return true;
    } else if (type == VRESTIR_LIGHT_DIRECTIONAL) {
        float3 dirW = f3(__ldg(&lp->dirW[0]), __ldg(&lp->dirW[1]), __ldg(&lp->dirW[2]));
        ls.distance = FLT_MAX; ls.dir = -dirW; ls.Li = I; ls.pdf = 0.f;
        return true;
    }
    ls.distance = 0.f; ls.dir = f3(0.f); ls.Li = f3(0.f); ls.pdf = 0.f;
    return false;
}
VRD float3 computeRayOrigin(float3 pos, float3 normal) {
    const float origin = 1.f / 32.f, fScale = 1.f / 65536.f, iScale = 256.f;
    float P[3] = {pos.x, pos.y, pos.z}, N[3] = {normal.x, normal.y, normal.z}, out[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int iOff = f2i(N[i] * iScale);
        int bits = __float_as_int(P[i]) + (P[i] < 0.f ? -iOff : iOff);
        float iPos = __int_as_float(bits);
        float fOff = N[i] * fScale;
        out[i] = fabsf(P[i]) < origin ? P[i] + fOff : iPos;
    }
    return f3(out[0], out[1], out[2]);
}
struct TriangleLightSample { uint32_t triangleIndex; float3 posW, normalW, dir; float distance; float3 Le; float pdf, pdfArea, cosTheta; float2 uv; };
VRD bool sampleTriangle(float3 posW, uint32_t triangleIndex, float2 u, TriangleLightSample& ls) {
    ls.triangleIndex = triangleIndex; ls.pdf = 0.f; ls.pdfArea = 0.f; ls.cosTheta = 0.f; ls.Le = f3(0.f); ls.distance = 0.f;
    const float* t = (const float*)&c_scene.tris[triangleIndex];   // posW[3][3], normal[3], area, Le[3]
    float su = sqrtf(u.x);
    float bx = 1.f - su, by = u.y * su;
    float3 bc = f3(1.f - bx - by, bx, by);
    ls.uv = u;
    float3 p0 = f3(__ldg(t + 0), __ldg(t + 1), __ldg(t + 2)), p1 = f3(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5)), p2 = f3(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8));
    float3 n = f3(__ldg(t + 9), __ldg(t + 10), __ldg(t + 11));
    const float area = __ldg(t + 12);
    ls.posW = p0 * bc.x + p1 * bc.y + p2 * bc.z;
    ls.posW = computeRayOrigin(ls.posW, n);
    float3 toLight = ls.posW - posW;
    const float distSqr = fmaxf(FLT_MIN, dot(toLight, toLight));
    ls.distance = sqrtf(distSqr);
    ls.dir = toLight / ls.distance;
    ls.normalW = n;
    float cosTheta = dot(ls.normalW, -ls.dir);
    if (cosTheta <= 0.f) return false;
    ls.Le = f3(__ldg(t + 13), __ldg(t + 14), __ldg(t + 15));
    float denom = fmaxf(FLT_MIN, cosTheta * area);
    ls.pdf = distSqr / denom;
    ls.cosTheta = cosTheta;
    ls.pdfArea = 1.f / area;
    return true;
}
VRD bool emissiveSampleLight(float3 posW, SampleGenerator& sg, TriangleLightSample& ls) {
    ls.triangleIndex = 0; ls.posW = f3(0.f); ls.normalW = f3(0.f); ls.dir = f3(0.f); ls.distance = 0.f; ls.Le = f3(0.f); ls.pdf = 0.f; ls.pdfArea = 0.f; ls.cosTheta = 0.f; ls.uv = make_float2(0, 0);
    if (c_scene.triCount == 0) return false;
    float2 rnd = sampleNext2D(sg);
    uint32_t count = (uint32_t)c_scene.triCount;
    uint32_t index = min(count - 1, f2u(rnd.x * (float)count));
    uint4 item = __ldg(&c_scene.alias[index]);
    uint32_t triangleIndex = rnd.y >= __uint_as_float(item.x) ? item.y : item.z;
    float triangleSelectionPdf = __ldg(&c_scene.aliasWeights[triangleIndex]) / c_scene.aliasWeightSum;
    float2 u = sampleNext2D(sg);
    if (!sampleTriangle(posW, triangleIndex, u, ls)) return false;
    ls.pdf *= triangleSelectionPdf;
    ls.pdfArea *= triangleSelectionPdf;
    return true;
}

// VR/VolumeUtils.slang:12-149
VRD_NOINLINE bool sampleSceneLights(float3 rayOrigin, bool kEnv, bool kAnalytic, bool kEmissive, SampleGenerator& sg, SceneLightSample& ls, int& outLightIndex, float2& outLightUV,
                                    const float* impTop = nullptr) {
    ls.dir = f3(0.f); ls.distance = 0.f; ls.Li = f3(0.f); ls.pdf = 0.f; ls.pdfArea = 0.f; ls.rayDir = f3(0.f); ls.rayDistance = 0.f;
    if (!kEnv && !kAnalytic && !kEmissive) return false;
    float p0 = kEnv ? 1.f : 0.f, p1 = kAnalytic ? 1.f : 0.f, p2 = kEmissive ? 1.f : 0.f;
    float sum = p0 + p1 + p2;
    if (sum == 0.f) return false;
    float invSum = 1.f / sum;
    p0 *= invSum; p1 *= invSum; p2 *= invSum;
    float u = sampleNext1D(sg);
    if (kEnv) {
        if (u < p0) {
            EnvMapSample lightSample;
            envSample(sampleNext2D(sg), lightSample, impTop);
            float pdf = p0 * lightSample.pdf;
            ls.rayDir = ls.dir = lightSample.dir;
            ls.rayDistance = ls.distance = kRayTMax;
            ls.pdf = pdf; ls.pdfArea = pdf;
            ls.Li = pdf > 0.f ? lightSample.Le / pdf : f3(0.f);
            outLightIndex = -1;
            outLightUV = envEncodeLightUV(ls.rayDir, outLightIndex);
            return !(isnan(ls.rayDir.x) || isnan(ls.rayDir.y) || isnan(ls.rayDir.z));
        }
        u -= p0;
    }
    if (kAnalytic) {
        if (u < p1) {
            u /= p1;
            uint32_t lightCount = (uint32_t)c_scene.lightCount;
            uint32_t lightIndex = min(f2u(u * (float)lightCount), lightCount - 1);
            float selectionPdf = p1 / (float)lightCount;
            AnalyticLightSample lightSample;
            sampleLight(rayOrigin, &c_scene.lights[lightIndex], lightSample);
            outLightIndex = (int)lightIndex;
            ls.rayDir = ls.dir = lightSample.dir;
            ls.rayDistance = ls.distance = lightSample.distance;
            if (lightSample.pdf == 0) lightSample.pdf = 1.f;
            ls.pdf = selectionPdf * lightSample.pdf;
            ls.pdfArea = ls.pdf;
            ls.Li = lightSample.Li / ls.pdf;
            outLightUV = make_float2(0, 0);
            return true;
        }
        u -= p1;
    }
    if (kEmissive) {
        if (u < p2) {
            TriangleLightSample lightSample;
            bool valid = emissiveSampleLight(rayOrigin, sg, lightSample);
            float pdf = p2 * lightSample.pdf;
            float pdfArea = p2 * lightSample.pdfArea;
            float3 offsetPos = computeRayOrigin(lightSample.posW, lightSample.normalW);
            float3 toLight = offsetPos - rayOrigin;
            ls.rayDistance = length(toLight);
            ls.rayDir = normalize(toLight);
            ls.dir = lightSample.dir; ls.distance = lightSample.distance;
            ls.pdf = pdf; ls.pdfArea = pdfArea;
            ls.Li = pdf > 0.f ? lightSample.Le * c_scene.emissiveMul / pdf : f3(0.f);
            outLightIndex = (int)lightSample.triangleIndex + c_scene.lightCount;
            outLightUV = lightSample.uv;
            return valid;
        }
    }
    return false;
}

// VR/VolumeUtils.slang:454-492
VRD float3 SampleDirectLighting(SampleGenerator& sg, float& pdf, const MediumInteraction& mi, const SamplingOptions& o, bool enableShadow, int& outLightIndex, float2& outLightUV) {
    pdf = 0.f;
    float3 Ld = f3(0.f);
    SceneLightSample ls;
    bool valid = sampleSceneLights(mi.p, o.useEnvironmentLights, o.useAnalyticLights, o.useEmissiveLights, sg, ls, outLightIndex, outLightUV);
    pdf = ls.pdfArea;
    if (!valid) { pdf = 0.f; return f3(0.f); }
    ls.pdf = 1.f;
    Ray shadowRay = makeRay(mi.p, ls.rayDir, 0, ls.rayDistance);
    if (enableShadow) {
        float visibility = computeVisibility(shadowRay, sg, o.lightSamples, o.lightingMipLevel, o.lightingUseLinearSampler, o.lightingTrackingMethod, o.lightingTStepScale);
        ls.Li = ls.Li * visibility;
    }
    float ph = mi.phaseFunction(mi.wo, ls.dir);
    Ld = Ld + ph * ls.Li / ls.pdf;
    return Ld;
}
// VR/VolumeUtils.slang:419-452 (reference path tracer NEE)
VRD float3 directLighting(SampleGenerator& sg, const MediumInteraction& mi, int nSamples, bool kEnv, bool kAnalytic, bool kEmissive, int mip, uint32_t method) {
    float3 Ld = f3(0.f);
    for (int i = 0; i < nSamples; i++) {
        SceneLightSample ls; int idx; float2 uv;
        bool valid = sampleSceneLights(mi.p, kEnv, kAnalytic, kEmissive, sg, ls, idx, uv);
        if (!valid) continue;
        ls.pdf = 1.f;
        Ray shadowRay = makeRay(mi.p, ls.rayDir, 0, ls.rayDistance);
        float vis = computeVisibility(shadowRay, sg, 1, mip, true, method, 1.f);
        ls.Li = ls.Li * vis;
        float ph = mi.phaseFunction(mi.wo, ls.dir);
        Ld = Ld + ph * ls.Li / ls.pdf;
    }
    Ld = Ld / (float)nSamples;
    return Ld;
}

// ------------------------------------------------------------------------------------------------ camera
VRD float3 camRayDirNN(float3 U, float3 V, float3 Wv, int px, int py, int W, int H) {
    float2 p = make_float2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H);
    float2 ndc = make_float2(2.f * p.x + -1.f, -2.f * p.y + 1.f);
    return ndc.x * U + ndc.y * V + Wv;
}

// ------------------------------------------------------------------------------------------------ ReSTIR helpers
VRD float3 decodeEmissivePosition(int lightID, float2 lightUV) { return f3(lightUV.x, lightUV.y, __int_as_float(lightID)); }
VRD void encodeEmissivePosition(float3 pos, int& lightID, float2& lightUV) { lightID = __float_as_int(pos.z); lightUV = make_float2(pos.x, pos.y); }
VRD float4 decodeWiDist(float3 in, bool reuseAsVertex = false) {
    if (reuseAsVertex) {   // VERTEX_REUSE (VR/ReSTIRHelper.slang:21-27): a world-space vertex (w = -1) or "left the medium"
        if (in.x == kRayTMax) return make_float4(0.f, 0.f, 0.f, kRayTMax);
        return make_float4(in.x, in.y, in.z, -1.f);
    }
    float3 wi; wi.x = in.x; wi.y = in.y;
    wi.z = sqrtf(1 - (in.x * in.x + in.y * in.y));
    if (isnan(wi.z)) wi.z = 0.f;
    if (in.z < 0) { wi.z = -wi.z; in.z = -in.z; }
    return make_float4(wi.x, wi.y, wi.z, in.z);
}
VRD float3 encodeWiDist(float4 in) { float3 wi; wi.x = in.x; wi.y = in.y; wi.z = in.w; if (in.z < 0) wi.z = -wi.z; return wi; }
template <int B> VRD int decodeMaxIndirectBounces(int storage) { return B == 1 ? 0 : storage >> 20; }
VRD int encodeMaxIndirectBounces(int storage, int bounce) { return (int)(((uint32_t)bounce << 20) | ((uint32_t)storage & 0xFFFFFu)); }
VRD int decodePathTag(int storage) { return (storage >> 16) & 0xF; }
VRD int encodePathTag(int storage, int tag) { return (int)(((uint32_t)tag << 16) | ((uint32_t)storage & 0xFFF0FFFFu)); }

template <int B> VRD void takeSample(const Reservoir& r, Reservoir& state, bool sel) {
    state.depth = sel ? r.depth : state.depth;
    state.p_y = sel ? r.p_y : state.p_y;
    state.lightUV = sel ? r.lightUV : state.lightUV;
    state.lightID = sel ? r.lightID : state.lightID;
    if (B > 1) state.extraBounceStartId = sel ? r.extraBounceStartId : state.extraBounceStartId;
    if (B > 1) state.p_partial = sel ? r.p_partial : state.p_partial;   // VERTEX_REUSE (VR/Reservoir.slang:46-47,76-77); 0 everywhere without it
    state.sampledPixel = sel ? r.sampledPixel : state.sampledPixel;
}
// VR/Reservoir.slang:26-87
template <int B> VRD bool simpleResampleStep(const Reservoir& reservoir, Reservoir& state, SampleGenerator& sg) {
    float sampleWeight = reservoir.runningSum;
    state.M += reservoir.M;
    if (sampleWeight <= 0.0f) return false;
    state.runningSum += sampleWeight;
    bool selectSample = sampleNext1D(sg) * state.runningSum < sampleWeight;
    takeSample<B>(reservoir, state, selectSample);
    return selectSample;
}
template <int B> VRD bool simpleResampleStepWithMaxM(const Reservoir& reservoir, float MThreshold, Reservoir& state, SampleGenerator& sg) {
    float correctedM = fminf(MThreshold, reservoir.M);
    float sampleWeight = correctedM == 0.0f ? 0.0f : correctedM / reservoir.M * reservoir.runningSum;
    state.M += correctedM;
    if (sampleWeight <= 0.0f) return false;
    state.runningSum += sampleWeight;
    bool selectSample = sampleNext1D(sg) * state.runningSum < sampleWeight;
    takeSample<B>(reservoir, state, selectSample);
    return selectSample;
}

// extra-bounce records: 12-byte packed float3 (global) or a per-thread register array (K1)
struct ExtraProvider {
    const float3* global; const float3* local;
    VRD float3 get(int i) const {
        if (local) return local[i];
        const float* p = (const float*)(global + i);
        return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
};
struct ExtraProviderRW {   // buffer written by the same kernel (K2's gCurExtraBounceReservoirs)
    const float3* global;
    VRD float3 get(int i) const { return global[i]; }
};

// VR/ReSTIRHelper.slang:435-496, first half: the shadow ray of a stored light sample and its unshadowed radiance * phase.
// Returns isValidSample.  Shared by the per-pixel kernels and the wavefront gather / combine kernels.
VRD bool lightRayAndLd(const MediumInteraction& mi, int lightID, float2 lightUV, bool isLastFrame, Ray& shadowRay, float3& Ld) {
    shadowRay = makeRay(mi.p, f3(0.f, 0.f, 1.f), 0, 0); Ld = f3(0.f); bool isValidSample = true;
    if (lightID < 0) {
        float3 wiWorld = envDecodeLightUV(lightUV, lightID, isLastFrame);
        shadowRay = makeRay(mi.p, wiWorld, 0, kRayTMax);
        Ld = envEval(wiWorld, isLastFrame) * mi.phaseFunction(mi.wo, wiWorld);
    } else if (lightID < c_scene.lightCount) {
        AnalyticLightSample ls;
        sampleLight(mi.p, &c_scene.lights[lightID], ls);
        shadowRay = makeRay(mi.p, ls.dir, 0, ls.distance);
        Ld = ls.Li * mi.phaseFunction(mi.wo, ls.dir);
    } else {
        TriangleLightSample ls;
        isValidSample = c_scene.triCount != 0 && sampleTriangle(mi.p, (uint32_t)(lightID - c_scene.lightCount), lightUV, ls);
        if (isValidSample) {
            shadowRay = makeRay(mi.p, ls.dir, 0, ls.distance);
            Ld = ls.Le * c_scene.emissiveMul * mi.phaseFunction(mi.wo, ls.dir) * ls.cosTheta / (ls.distance * ls.distance);
        }
    }
    return isValidSample;
}
// March provider of the target-function evaluation.  The transmittances an evaluation needs (camera ray to the first vertex,
// one scatter segment per indirect bounce, last vertex to the light) are independent of each other and of the random-number
// stream whenever the tracking method is deterministic, so the evaluation code is written once against this policy:
//   InlineMarch   computes each transmittance in place (per-pixel kernels, and every stochastic tracking method);
//   EmitMarch / ConsumeMarch (vr_wavefront.cu) run the SAME stage code twice: the first pass turns every march into a task
//   of a compacted stream and continues with a placeholder, the second pass reads the marched results — identical
//   arithmetic, different schedule.
// A march is identified by (evaluation id set with beginEval, slot): slot 0 camera, 1..3 scatter segment of bounce 0..2, 4 light.
enum { MARCH_SLOT_CAMERA = 0, MARCH_SLOT_SCATTER0 = 1, MARCH_SLOT_LIGHT = 4, MARCH_SLOTS = 5 };
struct InlineMarch {
    static constexpr bool kStore = true;    // the pass writes the stage's outputs
    VRD void beginEval(int) {}
    VRD float visibility(int, const Ray& ray, SampleGenerator& sg, int samples, int mip, bool linear, uint32_t method, float tStepScale) {
        return computeVisibility(ray, sg, samples, mip, linear, method, tStepScale);
    }
};

template <class MP>
VRD float3 evaluate_L_in_volume(const MediumInteraction& mi, int lightID, float2 lightUV, float& precomputedVisibility, SampleGenerator& sg, const SamplingOptions& o, bool lightVisibilityReuse,
                                bool isLastFrame, bool cullNonOpaqueGeometry, MP& mp) {
    Ray shadowRay; float3 Ld;
    const bool isValidSample = lightRayAndLd(mi, lightID, lightUV, isLastFrame, shadowRay, Ld);
    const bool useLastFrameGrid = c_scene.vol.usePrevGridForReproj && isLastFrame && c_scene.vol.hasAnimation;
    const int densityGridOffset = useLastFrameGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
    float Tr = 1.f;
    if (isValidSample) {
        if (lightVisibilityReuse) Tr = precomputedVisibility;   // VERTEX_REUSE: the shadow-ray transmittance the sample's own pixel stored
        else {
            Tr = mp.visibility(MARCH_SLOT_LIGHT, shadowRay, sg, o.lightSamples, cullNonOpaqueGeometry ? o.lightingMipLevel + densityGridOffset : 0, o.lightingUseLinearSampler,
                               o.lightingTrackingMethod, o.lightingTStepScale);
            precomputedVisibility = Tr;
        }
    }
    return Tr * Ld;
}

// VR/ReSTIRHelper.slang:91-423 (no SURFACE_SCENE).  VERTEX_REUSE is a run-time choice: o.vertexReuseStartBounce is the
// reference's S when mVertexReuse is on and VR_NO_VERTEX_REUSE (larger than any bounce count) otherwise, which turns every
// VERTEX_REUSE condition below into its #else branch.  `tap` is inout (REUSETYPE): unless spatialReuse reads it, the
// evaluation leaves the part of F past the reuse vertex in tap.p_partial.
template <int B, class Extra, class MP>
__device__ float3 evaluate_F_(Reservoir& tap, const Extra& extra, Ray ray, SampleGenerator& sg, const SamplingOptions& o, bool isLastFrame, bool noReuse, bool spatialReuse, bool isFinalShading, MP& mp) {
    const vrestir_volume_desc& vd = c_scene.vol;
    const int S = o.vertexReuseStartBounce;
    const bool useLastFrameGrid = vd.usePrevGridForReproj && isLastFrame && vd.hasAnimation;
    const int mipLevelOffset = useLastFrameGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
    bool isBackgroundSample = tap.depth == kRayTMax;
    ray.tMax = tap.depth;
    float3 F = f3(1.f);
    int maxIndirectBounces = 0; bool isSelfEmission;
    if (B > 1) { maxIndirectBounces = decodeMaxIndirectBounces<B>(tap.sampledPixel); isSelfEmission = maxIndirectBounces == 0 && tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID; }
    else isSelfEmission = tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID;
    float visibility = 1.f;
    float3 p_World = ray.at(ray.tMax);
    MediumInteraction mi = makeMI(ray.at(ray.tMax), -ray.dir, true);
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    {
        float density = (isBackgroundSample || noReuse) ? 1.f : DensityWorldSpace(p_World, mipLevelOffset);
        if (density == 0.f) return f3(0.f);
        if (!noReuse)
            visibility = mp.visibility(MARCH_SLOT_CAMERA, ray, sg, o.visibilitySamples, o.visibilityMipLevel + mipLevelOffset, o.visibilityUseLinearSampler, o.visibilityTrackingMethod, o.visibilityTStepScale);
        float3 sigma_s = isBackgroundSample ? f3(1.f) : (isSelfEmission ? sigA : sigS);
        if (noReuse && !isBackgroundSample) sigma_s = sigma_s / vd.sigma_t;
        F = F * (visibility * density * sigma_s);
    }
    float3 P_prefix = f3(1.f);
    int bounceId = 0;
    if (any_gt0(F)) {
        if (isBackgroundSample) {
            F = F * envEval(ray.dir, isLastFrame);
        } else if (isSelfEmission) {
            F = F * EmissionWorldSpace(p_World, useLastFrameGrid);
        } else {
            bool isScatterSelfEmission = false;
            if (B > 1 && maxIndirectBounces > 0) {
                Ray scatterRay = makeRay(p_World, f3(0.f, 0.f, 1.f), 0.f, 0.f);
                isScatterSelfEmission = decodePathTag(tap.sampledPixel) == 1;
                const int numIndirectBounces = maxIndirectBounces;
                for (; bounceId < numIndirectBounces; bounceId++) {
                    bool isCurrentVertexEmissive = isScatterSelfEmission && bounceId == numIndirectBounces - 1;
                    float4 wiDist = decodeWiDist(extra.get(tap.extraBounceStartId + bounceId), bounceId + 1 >= S);
                    if (isCurrentVertexEmissive && bounceId + 1 < S) { float3 e = decodeEmissivePosition(tap.lightID, tap.lightUV); wiDist = make_float4(e.x, e.y, e.z, -1.f); }
                    if (wiDist.w == kRayTMax) return f3(0.f);
                    float dist = 1.f;
                    if (wiDist.w == -1.f) {
                        float3 disp = f3(wiDist.x, wiDist.y, wiDist.z) - p_World;
                        dist = length(disp);
                        float3 dir = normalize(disp);
                        scatterRay = makeRay(p_World, dir, 0.f, dist);
                    } else {
                        scatterRay = makeRay(p_World, f3(wiDist.x, wiDist.y, wiDist.z), 0.f, wiDist.w);
                    }
                    float bsdf = mi.phaseFunction(mi.wo, scatterRay.dir);
                    F = F * bsdf;
                    if (all_eq0(F)) return f3(0.f);
                    if (bounceId == S) {   // :289-296, the bounce past the reuse vertex
                        if (spatialReuse) return F * tap.p_partial;
                        P_prefix = F;
                    }
                    if (wiDist.w == -1.f) p_World = f3(wiDist.x, wiDist.y, wiDist.z);
                    else p_World = scatterRay.at(scatterRay.tMax);
                    float3 sigma_s; float scatterDensity;
                    if (!noReuse) {
                        sigma_s = isCurrentVertexEmissive ? sigA : sigS;
                        scatterDensity = fmaxf(0.f, DensityWorldSpace(p_World, mipLevelOffset));
                    } else {
                        sigma_s = isCurrentVertexEmissive ? sigA / vd.sigma_t : sigS / vd.sigma_t;
                        scatterDensity = 1.f;
                    }
                    F = F * (scatterDensity * sigma_s);
                    if (bounceId + 1 == S || (isCurrentVertexEmissive && bounceId + 1 < S)) F = F * (1.f / (dist * dist));
                    if (all_eq0(F)) return f3(0.f);
                    float scatterVisibility = 1.f;
                    if (!noReuse)
                        scatterVisibility = mp.visibility(MARCH_SLOT_SCATTER0 + bounceId, scatterRay, sg, o.visibilitySamples, o.visibilityMipLevel + mipLevelOffset, o.visibilityUseLinearSampler,
                                                          o.visibilityTrackingMethod, o.visibilityTStepScale);
                    F = F * scatterVisibility;
                    mi.wo = -scatterRay.dir;
                    mi.p = p_World;
                    if (all_eq0(F)) return f3(0.f);
                }
                if (isScatterSelfEmission) F = F * EmissionWorldSpace(p_World, useLastFrameGrid);
                else mi = makeMI(scatterRay.at(scatterRay.tMax), -scatterRay.dir, true);
            }
            if (!isScatterSelfEmission && any_gt0(F)) {
                float precomputedVisibility = tap.p_partial;
                const bool lightVisibilityReuse = B > 1 && bounceId == S && spatialReuse;
                F = F * evaluate_L_in_volume(mi, tap.lightID, tap.lightUV, precomputedVisibility, sg, o, lightVisibilityReuse, isLastFrame, !isFinalShading, mp);
                if (B > 1 && bounceId == S && !spatialReuse) tap.p_partial = precomputedVisibility;
            }
        }
    }
    // :415-420, componentwise as written there
    if (B > 1 && bounceId > S && !spatialReuse) tap.p_partial = luminance(F / P_prefix);
    return F;
}
// `tap` is inout: VR/ReSTIRHelper.slang:426-441 under REUSETYPE = inout
template <int B, class Extra, class MP>
VRD float evaluate_P_hat(const Ray& ray, SampleGenerator& sg, const Extra& extra, const SamplingOptions& o, Reservoir& tap, bool isLastFrame, bool spatialReuse, MP& mp) {
    float3 F = evaluate_F_<B>(tap, extra, ray, sg, o, isLastFrame, false, spatialReuse, false, mp);
    return luminance(F);
}
// VR/ReSTIRHelper.slang:600-607: the reservoir travels by value, the caller's p_partial stays
template <int B, class Extra, class MP>
VRD float evaluatePHatReadOnly(const Ray& ray, SampleGenerator& sg, const Extra& extra, const SamplingOptions& o, Reservoir tap, bool isLastFrame, bool spatialReuse, MP& mp) {
    return evaluate_P_hat<B>(ray, sg, extra, o, tap, isLastFrame, spatialReuse, mp);
}
// VR/ReSTIRHelper.slang:560-597: resampleNeighbor (spatialReuse = false, overwrites tap.p_partial) / resampleNeighborSpatialReuse
template <int B, class Extra, class MP>
VRD void resampleNeighbor(Reservoir& tap, const Ray& ray, SampleGenerator& sg, const Extra& extra, const SamplingOptions& o, bool spatialReuse, MP& mp) {
    if (tap.runningSum == 0.f) return;
    float p_y_hat = evaluate_P_hat<B>(ray, sg, extra, o, tap, false, spatialReuse, mp);
    float weight = p_y_hat / tap.p_y;
    // an emit pass (placeholder transmittances of 1) must keep every sample the real pass can keep: with a denormal p_y the
    // placeholder ratio overflows where the real one is finite
    if (isinf(weight) || isnan(weight)) weight = MP::kStore ? 0.f : 1.f;
    tap.runningSum *= weight;
    tap.p_y = p_y_hat;
}

// ------------------------------------------------------------------------------------------------ march-task streams
// Explicit-origin tasks are PREPARED by the emitting kernel (where all 32 lanes are busy): the ray is already in the index
// space of the mip it will be marched through and clipped against the volume box (WorldToMedium + IntersectVolumeBound,
// VR/VolumeBase.slang:103-175); rays that miss the box get their result (transmittance 1) written at once and never
// become a task.  Layout, 3 x uint4: (pos.xyz, tNear) (dir.xyz, tFar) (result index, 0, 0, 0).
struct PreparedRay { float3 pos, dir; float tNear, tFar; };
VRD bool wfPrepare(const Ray& rW, int mip, bool vertexCenter, PreparedRay& o) {
    const DSlot& g = c_scene.slots[mip];
    Ray ray; ray.origin = mulPoint(rW.origin, g.w2m); ray.dir = mulVec(rW.dir, g.w2m); ray.tMin = rW.tMin; ray.tMax = rW.tMax;
    float3 mn = v3(g.bmin), mx = v3(g.bmax);
    if (vertexCenter) { ray.origin = ray.origin - f3(0.5f); mn = mn - f3(0.5f); mx = mx - f3(0.5f); }
    o.pos = ray.origin; o.dir = ray.dir;
    return IntersectP(mn, mx, ray, o.tNear, o.tFar);
}
// Append one prepared task per lane with `want` to a stream; every lane of the warp must call (one atomic per warp).
// missValue: what a ray that misses the volume box leaves in its result slot (transmittance 1; kRayTMax for a distance task)
VRD void wfEmitRay(const WfStream& s, bool want, const Ray& rW, int mip, bool vertexCenter, float* results, unsigned out, float missValue = 1.f) {
    PreparedRay pr;
    bool has = false;
    if (want) { has = wfPrepare(rW, mip, vertexCenter, pr); if (!has) results[out] = missValue; }
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(s.count, (unsigned)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (has) {
        const unsigned pos = base + __popc(bal & ((1u << lane) - 1u));
        if (pos < s.capacity) {
            uint4* q = s.tasks + 3 * (size_t)pos;
            q[0] = make_uint4(__float_as_uint(pr.pos.x), __float_as_uint(pr.pos.y), __float_as_uint(pr.pos.z), __float_as_uint(pr.tNear));
            q[1] = make_uint4(__float_as_uint(pr.dir.x), __float_as_uint(pr.dir.y), __float_as_uint(pr.dir.z), __float_as_uint(pr.tFar));
            q[2] = make_uint4(out, 0u, 0u, 0u);
        }
    }
}

// The same from divergent code (the emit pass of the generic task-stream path calls it from inside the evaluation): the lanes
// that happen to be converged here share one atomic (cooperative-groups coalesced group).  Task order in the stream depends on
// that grouping, results do not: every task carries its result index.
VRD void wfEmitRayAny(const WfStream& s, const Ray& rW, int mip, bool vertexCenter, float* results, unsigned out) {
    PreparedRay pr;
    if (!wfPrepare(rW, mip, vertexCenter, pr)) { results[out] = 1.f; return; }
    cg::coalesced_group grp = cg::coalesced_threads();
    unsigned base = 0;
    if (grp.thread_rank() == 0) base = atomicAdd(s.count, (unsigned)grp.size());
    base = grp.shfl(base, 0);
    const unsigned pos = base + grp.thread_rank();
    if (pos < s.capacity) {
        uint4* q = s.tasks + 3 * (size_t)pos;
        q[0] = make_uint4(__float_as_uint(pr.pos.x), __float_as_uint(pr.pos.y), __float_as_uint(pr.pos.z), __float_as_uint(pr.tNear));
        q[1] = make_uint4(__float_as_uint(pr.dir.x), __float_as_uint(pr.dir.y), __float_as_uint(pr.dir.z), __float_as_uint(pr.tFar));
        q[2] = make_uint4(out, 0u, 0u, 0u);
    }
}

// ------------------------------------------------------------------------------------------------ pixel mapping
// one thread per pixel; a warp covers an 8x4 pixel tile, a CTA of 4 warps 16x8 pixels
// A warp always covers an 8x4 pixel tile.  128-thread CTAs cover 16x8 pixels (2x2 tiles); kernels whose per-pixel run time
// varies a lot are launched with one warp per CTA (a CTA's registers are only released when its slowest warp retires).
VRD bool pixelOf(const FrameParams& fp, int& x, int& y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockDim.x == 32) {
        x = blockIdx.x * 8 + (lane & 7);
        y = fp.rowBegin + blockIdx.y * 4 + (lane >> 3);
    } else {
        x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
        y = fp.rowBegin + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    }
    return x < fp.W && y < fp.rowEnd;
}
VRD Ray primaryRay(const FrameParams& fp, int x, int y) {
    return makeRay(fp.camPos, normalize(camRayDirNN(fp.camU, fp.camV, fp.camW, x, y, fp.W, fp.H)), 0.f, kRayTMax);
}

}  // namespace vrd
