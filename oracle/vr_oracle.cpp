/*
 * vr_oracle.cpp — CPU oracle for the VolumetricReSTIR per-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see vr_oracle.h).  Function-for-function restatement of the reference's shader logic in
 * fp32 scalar C++; each function cites the reference file:line it follows.  Abbreviations:
 *   VR/ = Source/RenderPasses/VolumetricReSTIR/    F/ = Source/Falcor/
 *
 * PIN STATUS: parity unpinned against reference OUTPUT — the reference (Slang SM 6.5 / DXR / Falcor / D3D12 / Windows) cannot run
 * here and ships no golden vectors for this path; only the RNG core is checked against reference-held code (oracle/_ref, the
 * vendored xoshiro C) and its seeding KATs.  What stands in for reference output: independent Python restatements written from
 * the Slang, not from this file (oracle/march_witness.py, light_witness.py, stage_witness.py, alias_oracle.py, mip_oracle.py,
 * post_oracle.py), which every stage of this oracle is checked against in tests/ (DESIGN.md section 2 lists them).
 *
 * Build: g++ -O2 -ffp-contract=off -pthread (oracle/Makefile).  No fast-math, no FMA contraction, so every float
 * expression below is evaluated exactly as written, left to right.
 *
 * Hardware behaviours that are not in the reference source and are pinned here:
 *   - trilinear SampleLevel on the brick atlas: fp32 lerp x, then y, then z with lerp(a,b,t) = fma(t, b-a, a),
 *     texel centres at +0.5, fetched from the brick's own 10^3 apron-inclusive block (never crosses bricks);
 *   - coarse/conservative mips are 8-bit UNORM (the reference's ATLAS_COMPRESSION==1 variant, F/Scene/Scene.cpp:3164-3174,
 *     instead of BC4): point fetch = u8 * fl(1/255); trilinear = filter the codes, then * fl(1/255);
 *   - lat-long env lookup: bilinear, wrap U, clamp V, texel centres at +0.5;
 *   - importance-map mip chain: 2x2 box, ((a+b)+(c+d))*0.25;
 *   - structured-buffer reads out of range return 0; float->int conversions saturate, NaN -> 0;
 *   - normalize(v) = v / sqrt(dot(v,v)); dot = (x*x + y*y) + z*z.
 *   - world<->medium matrices are supplied pre-multiplied by the host (affine; the /w is dropped).
 */
#include "vr_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

// ------------------------------------------------------------------------------------------------ small vector math
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct int3 { int x, y, z; };
struct int2 { int x, y; };

inline float3 f3(float a) { return {a, a, a}; }
inline float3 f3(float a, float b, float c) { return {a, b, c}; }
inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, float3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float3 operator/(float3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(float3 a) { return sqrtf(dot(a, a)); }
inline float3 normalize(float3 a) { return a / sqrtf(dot(a, a)); }
inline float3 floor3(float3 a) { return {floorf(a.x), floorf(a.y), floorf(a.z)}; }
inline float3 abs3(float3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
inline float3 toF(int3 a) { return {(float)a.x, (float)a.y, (float)a.z}; }
inline bool any_gt0(float3 a) { return a.x > 0.f || a.y > 0.f || a.z > 0.f; }
inline bool all_eq0(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
inline float fsign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
// float -> int, D3D ftoi semantics: truncate, saturate, NaN -> 0
inline int f2i(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.f) return INT32_MAX;
    if (v <= -2147483648.f) return INT32_MIN;
    return (int)v;
}
inline uint32_t f2u(float v) {
    if (std::isnan(v) || v <= 0.f) return 0u;
    if (v >= 4294967296.f) return UINT32_MAX;
    return (uint32_t)v;
}
// F/Utils/Color/ColorHelpers.slang:39-42
inline float luminance(float3 rgb) { return dot(rgb, f3(0.2126f, 0.7152f, 0.0722f)); }

// row-vector * 4x4 (HLSL mul(float4(p,1), M)), row-major storage
inline float3 mulPoint(float3 p, const float* M) {
    return {p.x * M[0] + p.y * M[4] + p.z * M[8] + M[12], p.x * M[1] + p.y * M[5] + p.z * M[9] + M[13],
            p.x * M[2] + p.y * M[6] + p.z * M[10] + M[14]};
}
inline float3 mulVec(float3 v, const float* M) {
    return {v.x * M[0] + v.y * M[4] + v.z * M[8], v.x * M[1] + v.y * M[5] + v.z * M[9],
            v.x * M[2] + v.y * M[6] + v.z * M[10]};
}
inline float3 mulVec3x3(float3 v, const float* M) {  // mul(dir, (float3x3)M), M row-major 3x3
    return {v.x * M[0] + v.y * M[3] + v.z * M[6], v.x * M[1] + v.y * M[4] + v.z * M[7],
            v.x * M[2] + v.y * M[5] + v.z * M[8]};
}

constexpr float kRayTMax = FLT_MAX;  // VR/VolumeBase.slang:11
constexpr float M_PI_F = 3.14159265358979323846f;
constexpr float M_2PI_F = 6.28318530717958647693f;
constexpr float M_4PI_F = 12.5663706143591729539f;
constexpr float M_1_PI_F = 0.318309886183790671538f;
constexpr float M_1_2PI_F = 0.159154943091895335769f;
constexpr float M_1_4PI_F = 0.079577471545947667884f;
constexpr float M_PI_4_F = 0.785398163397448309616f;
constexpr uint32_t ID_UNDEFL = 0xFFFFFFFFu;
constexpr int MAX_BRICK_STEPS = 128;  // VR/VolumeTrackingAdapterGVDB.slang:4

// ------------------------------------------------------------------------------------------------ RNG
// F/Utils/Math/BitTricks.slang:45-61
inline uint32_t interleave_32bit(uint32_t vx, uint32_t vy) {
    uint32_t x = vx & 0x0000ffffu, y = vy & 0x0000ffffu;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu; y = (y | (y << 4)) & 0x0F0F0F0Fu;
    y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}

struct Counters {
    uint64_t taps = 0, voxels = 0, vbytes = 0, nodes = 0, rng = 0, marches = 0;
};
thread_local Counters tl_cnt;

// F/Utils/Sampling/UniformSampleGenerator.slang:49-75, Pseudorandom/SplitMix64.slang:54-78, Xoshiro.slang:52-66
struct SampleGenerator {
    uint32_t s[4];
    static uint64_t splitmix(uint64_t& state) {
        uint64_t z = (state += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    static SampleGenerator create(uint32_t px, uint32_t py, uint32_t sampleNumber) {
        uint64_t st = ((uint64_t)sampleNumber << 32) | (uint64_t)interleave_32bit(px, py);
        uint64_t s0 = splitmix(st), s1 = splitmix(st);
        SampleGenerator g;
        g.s[0] = (uint32_t)s0; g.s[1] = (uint32_t)(s0 >> 32); g.s[2] = (uint32_t)s1; g.s[3] = (uint32_t)(s1 >> 32);
        return g;
    }
    static uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    uint32_t next() {
        const uint32_t r = rotl(s[0] * 5u, 7) * 9u;
        const uint32_t t = s[1] << 9;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 11);
        return r;
    }
};
// F/Utils/Sampling/SampleGenerator.slang:58-73
inline float sampleNext1D(SampleGenerator& sg) { tl_cnt.rng++; return (float)(sg.next() >> 8) * 0x1p-24f; }
inline float2 sampleNext2D(SampleGenerator& sg) { float2 r; r.x = sampleNext1D(sg); r.y = sampleNext1D(sg); return r; }

// ------------------------------------------------------------------------------------------------ scene state
struct Ray { float3 origin, dir; float tMin, tMax; float3 at(float t) const { return origin + dir * t; } };

struct Reservoir {  // VR/HostDeviceSharedDefinitions.h:16-45 (+ extraBounceStartId, MAX_BOUNCES > 1)
    float runningSum, M, depth, p_y;
    float2 lightUV;
    int lightID, sampledPixel;
    int extraBounceStartId;
    float p_partial;   // VERTEX_REUSE (:29-31): the suffix of p-hat past the reuse vertex, kept for the spatial pass
};
struct ExtraBounce { float3 wi_dist; };          // VR/HostDeviceSharedDefinitions.h:53-56
struct Features { int noReflectiveSurface; float transmittance; };  // :67-71

struct SamplingOptions {  // VR/HostDeviceSharedDefinitions.h:82-138
    uint32_t visibilityTrackingMethod, lightingTrackingMethod;
    int lightSamples, lightingMipLevel, visibilitySamples, visibilityMipLevel;
    bool visibilityUseLinearSampler, lightingUseLinearSampler;
    float visibilityTStepScale, lightingTStepScale;
    bool useEnvironmentLights, useAnalyticLights, useEmissiveLights;
    int vertexReuseStartBounce;
};

struct AliasItem { float threshold; uint32_t indexA, indexB, pad; };

}  // namespace

struct vro_pass {
    vrestir_params P;
    // top-level dict keys (VR/VolumetricReSTIR.h:284-303)
    bool mOutputMotionVec = false, mFreezeFrame = false, mRandomizeFrameSeed = false;
    float volumeDensityScaleExtraControl = -1.f, volumeAlbedoExtraControl = -1.f, volumeAnisotropyExtraControl = -1.f;
    int envSamplerType = VRESTIR_ENV_SAMPLER_HIERARCHICAL;
    unsigned randState = 1;

    vrestir_volume_desc vol{};
    vrestir_volume_desc volBase{};
    vrestir_grid_slot slots[VRESTIR_MAX_SLOTS]{};
    std::vector<float> lut;  // 128 x float4
    bool haveVolume = false, haveCamera = false;

    vrestir_camera cam{};
    // env
    bool haveEnv = false;
    std::vector<float> envTexels; int envW = 0, envH = 0; float envIntensity = 1.f; float3 envTint{1, 1, 1};
    float envT[9], envInvT[9], envPrevT[9], envPrevInvT[9];
    std::vector<float> importance;  // mip chain, finest first
    std::vector<size_t> impOffset; int impBaseMip = 0; int impDim = 512;
    std::vector<float> envAliasThr, envAliasPdf; std::vector<uint32_t> envAliasRedirect;
    // lights
    std::vector<vrestir_light> lights;
    std::vector<vrestir_emissive_triangle> tris; std::vector<AliasItem> alias; std::vector<float> aliasWeights;
    float aliasWeightSum = 0.f, emissiveMul = 1.f;

    int W = 0, H = 0; int cx0 = 0, cy0 = 0, cx1 = 0, cy1 = 0;
    int threads = 0;
    // frame state (VR/VolumetricReSTIR.cpp:349-359,765-772)
    int mFrameCount = 0, mTemporalSampleAccumulated = 0; bool mOptionsChanged = true;
    float prevView[16], prevProj[16]; float3 prevU, prevV, prevW, prevPos;
    // buffers (VR/VolumetricReSTIR.h:95-108)
    std::vector<Reservoir> res[2], resT;
    std::vector<ExtraBounce> ext[2], extT;
    std::vector<Features> feat, featT;
    std::vector<uint32_t> k1Generator;   // per pixel: the generator after K1's candidate loop and after its p-hat evaluation (2 x 4 words), for the witnesses
    std::vector<float> refColor;  // gOutputColor of the mUseReference path
    int allocW = 0, allocH = 0, allocB = 0;
    int totalRoundId = 0;
    Counters cnt; vrestir_timings ms{};
    std::vector<int> dumpXYZ; std::vector<float> dumpT; bool dumping = false; int dumpMax = 0;
};

namespace {

using Pass = vro_pass;

// parallel-for over rows of the crop rectangle
template <class F> void forPixels(Pass& p, F f) {
    int nt = p.threads > 0 ? p.threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::atomic<int> nextRow{p.cy0};
    std::vector<std::thread> th;
    std::vector<Counters> cs(nt);
    auto body = [&](int tid) {
        tl_cnt = Counters{};
        for (;;) {
            int y = nextRow.fetch_add(1);
            if (y >= p.cy1) break;
            for (int x = p.cx0; x < p.cx1; x++) f(x, y);
        }
        cs[tid] = tl_cnt;
    };
    if (nt == 1) body(0);
    else { for (int i = 0; i < nt; i++) th.emplace_back(body, i); for (auto& t : th) t.join(); }
    for (auto& c : cs) {
        p.cnt.taps += c.taps; p.cnt.voxels += c.voxels; p.cnt.vbytes += c.vbytes; p.cnt.nodes += c.nodes;
        p.cnt.rng += c.rng; p.cnt.marches += c.marches;
    }
}

// ------------------------------------------------------------------------------------------------ GVDB tree access
// F/Scene/GVDB/gvdbNodes.slang:98-131
inline const vrestir_node& getNode(const vrestir_grid_slot& g, int lev, uint32_t n) { tl_cnt.nodes++; return g.nodes[lev][n]; }
inline uint32_t getChild(const vrestir_grid_slot& g, const vrestir_node& node, int clev, int b) {
    uint32_t listid = node.link;
    if (listid == ID_UNDEFL) return ID_UNDEFL;
    const uint64_t r = (uint64_t)g.res[clev] * g.res[clev] * g.res[clev];
    return g.childlist[clev][(uint64_t)listid * r + (uint64_t)b];
}
inline float3 nodePos(const vrestir_node& n) { return {(float)n.pos[0], (float)n.pos[1], (float)n.pos[2]}; }
inline bool outside(float3 pos, float3 vmin, float3 vmax) {
    return pos.x < vmin.x || pos.y < vmin.y || pos.z < vmin.z || pos.x >= vmax.x || pos.y >= vmax.y || pos.z >= vmax.z;
}
// F/Scene/GVDB/gvdbNodes.slang:175-280 (getNodeUnrolledThreeLevel / TwoLevel / getNodeAtPoint)
inline const vrestir_node* getNodeAtPoint(const vrestir_grid_slot& g, float3 pos, uint32_t& node_id) {
    node_id = ID_UNDEFL;
    const vrestir_node* node;
    float3 vmin;
    if (g.top_lev == 2) {
        node = &getNode(g, 2, 0); vmin = nodePos(*node);
        if (outside(pos, vmin, vmin + f3(4096.f))) return nullptr;
        int3 p = {f2i(pos.x - vmin.x) / 128, f2i(pos.y - vmin.y) / 128, f2i(pos.z - vmin.z) / 128};
        int b = (((p.z << 5) + p.y) << 5) + p.x;
        node_id = getChild(g, *node, 2, b);
        if (node_id == ID_UNDEFL) return nullptr;
        node = &getNode(g, 1, node_id); vmin = nodePos(*node);
    } else {
        node = &getNode(g, 1, 0); vmin = nodePos(*node);
    }
    {
        if (outside(pos, vmin, vmin + f3(128.f))) { node_id = ID_UNDEFL; return nullptr; }
        int3 p = {f2i(pos.x - vmin.x) / 8, f2i(pos.y - vmin.y) / 8, f2i(pos.z - vmin.z) / 8};
        int b = (((p.z << 4) + p.y) << 4) + p.x;
        node_id = getChild(g, *node, 1, b);
        if (node_id == ID_UNDEFL) return nullptr;
        node = &getNode(g, 0, node_id);
    }
    return node;
}

// ---- brick-pool voxel fetch (replaces the 3-D atlas texture; layout in include/vrestir.h) ----
inline float atlasVoxelRaw(const vrestir_grid_slot& g, uint32_t brick, int ix, int iy, int iz, int ch = 0) {
    // stored code (UNORM8 code as float, or the fp32 value); ix,iy,iz in [-1, 8]; outside the block = border colour 0
    if (ix < -1 || iy < -1 || iz < -1 || ix > 8 || iy > 8 || iz > 8) return 0.f;
    size_t idx = ((size_t)brick * g.atlas_channels + ch) * VRESTIR_BRICK_VOXELS + (size_t)((iz + 1) * 10 + (iy + 1)) * 10 + (ix + 1);
    tl_cnt.voxels++;
    if (g.atlas_format == VRESTIR_ATLAS_UNORM8) { tl_cnt.vbytes += 1; return (float)((const uint8_t*)g.atlas)[idx]; }
    tl_cnt.vbytes += 4;
    return ((const float*)g.atlas)[idx];
}
inline float atlasVoxel(const vrestir_grid_slot& g, uint32_t brick, int ix, int iy, int iz, int ch = 0) {
    float r = atlasVoxelRaw(g, brick, ix, iy, iz, ch);
    return g.atlas_format == VRESTIR_ATLAS_UNORM8 ? r * 0.003921568859368563f : r;
}
inline float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }   // pinned: one fused multiply-add
// SampleLevel with the linear border sampler at brick-local position p (voxel units, brick interior = [0,8)^3)
inline float sampleBrickLinear(const vrestir_grid_slot& g, uint32_t brick, float3 p, int ch = 0) {
    tl_cnt.taps++;
    float qx = p.x - 0.5f, qy = p.y - 0.5f, qz = p.z - 0.5f;
    float fx0 = floorf(qx), fy0 = floorf(qy), fz0 = floorf(qz);
    int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
    float fx = qx - fx0, fy = qy - fy0, fz = qz - fz0;
    float v000 = atlasVoxelRaw(g, brick, ix, iy, iz, ch), v100 = atlasVoxelRaw(g, brick, ix + 1, iy, iz, ch);
    float v010 = atlasVoxelRaw(g, brick, ix, iy + 1, iz, ch), v110 = atlasVoxelRaw(g, brick, ix + 1, iy + 1, iz, ch);
    float v001 = atlasVoxelRaw(g, brick, ix, iy, iz + 1, ch), v101 = atlasVoxelRaw(g, brick, ix + 1, iy, iz + 1, ch);
    float v011 = atlasVoxelRaw(g, brick, ix, iy + 1, iz + 1, ch), v111 = atlasVoxelRaw(g, brick, ix + 1, iy + 1, iz + 1, ch);
    // pinned filter: fp32 fma-lerp x, y, z on the stored codes; UNORM8 codes are scaled by fl(1/255) once, after filtering
    float c00 = lerpf(v000, v100, fx), c10 = lerpf(v010, v110, fx), c01 = lerpf(v001, v101, fx), c11 = lerpf(v011, v111, fx);
    float c0 = lerpf(c00, c10, fy), c1 = lerpf(c01, c11, fy);
    float r = lerpf(c0, c1, fz);
    return g.atlas_format == VRESTIR_ATLAS_UNORM8 ? r * 0.003921568859368563f : r;
}
inline float sampleBrickPoint(const vrestir_grid_slot& g, uint32_t brick, float3 p, int ch = 0) {
    tl_cnt.taps++;
    return atlasVoxel(g, brick, (int)floorf(p.x), (int)floorf(p.y), (int)floorf(p.z), ch);
}

// F/Scene/GVDB/gvdb.slang:6-31 getValueAtPoint (tree lookup + sample)
inline float getValueAtPoint(const vrestir_grid_slot& g, float3 pos, bool linear, int ch = 0) {
    uint32_t node_id;
    const vrestir_node* node = getNodeAtPoint(g, pos, node_id);
    if (node_id == ID_UNDEFL) return 0.f;
    float3 p_rel = pos - nodePos(*node);
    return linear ? sampleBrickLinear(g, node->link, p_rel, ch) : sampleBrickPoint(g, node->link, p_rel, ch);
}

// ------------------------------------------------------------------------------------------------ VolumeBase.slang
struct Ctx {  // everything a pixel needs
    const Pass& p;
    explicit Ctx(const Pass& pp) : p(pp) {}
    const vrestir_grid_slot& slot(int mip) const { return p.slots[mip]; }
    const vrestir_volume_desc& vd() const { return p.vol; }
};

// VR/VolumeBase.slang:103-116
inline Ray WorldToMedium(const Ctx& c, const Ray& r, int mip) {
    const float* M = c.slot(mip).world_to_medium;
    return {mulPoint(r.origin, M), mulVec(r.dir, M), r.tMin, r.tMax};
}
inline float3 WorldToMediumP(const Ctx& c, float3 pW, int mip) { return mulPoint(pW, c.slot(mip).world_to_medium); }

// VR/VolumeBase.slang:132-175
inline bool IntersectP(float3 pMin, float3 pMax, const Ray& ray, float& hitt0, float& hitt1) {
    float t0 = 0, t1 = ray.tMax;
    const float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z}, d[3] = {ray.dir.x, ray.dir.y, ray.dir.z};
    const float mn[3] = {pMin.x, pMin.y, pMin.z}, mx[3] = {pMax.x, pMax.y, pMax.z};
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1 / d[i];
        float tNear = (mn[i] - o[i]) * invRayDir;
        float tFar = (mx[i] - o[i]) * invRayDir;
        if (tNear > tFar) { float tmp = tNear; tNear = tFar; tFar = tmp; }
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return false;
    }
    hitt0 = t0; hitt1 = t1;
    return true;
}
inline bool IntersectVolumeBound(const Ctx& c, const Ray& ray, float& tMin, float& tMax, int mip, bool vertexCenter) {
    const auto& g = c.slot(mip);
    float3 mn = f3(g.bmin[0], g.bmin[1], g.bmin[2]), mx = f3(g.bmax[0], g.bmax[1], g.bmax[2]);
    if (vertexCenter) { mn = mn - f3(0.5f); mx = mx - f3(0.5f); }
    return IntersectP(mn, mx, ray, tMin, tMax);
}
// VR/VolumeBase.slang:177-180
inline float GetVolumeMaxDensity(const Ctx& c, int mip) { return c.slot(mip).max_value * c.vd().densityScaleFactorByScaling; }

// VR/VolumeBase.slang:234-243
inline float Density(const Ctx& c, float3 p, int mip) {
    const auto& g = c.slot(mip);
    if (p.x < g.bmin[0] || p.y < g.bmin[1] || p.z < g.bmin[2] || p.x >= g.bmax[0] || p.y >= g.bmax[1] || p.z >= g.bmax[2]) return 0.f;
    return getValueAtPoint(g, p, true) * c.vd().densityScaleFactorByScaling;
}
inline float DensityWorldSpace(const Ctx& c, float3 pW, int mip) { return Density(c, WorldToMediumP(c, pW, mip), mip); }

// VR/VolumeBase.slang:254-263 DensityInAtlas; p_local = brick-local voxel coordinate (p_atlas - o)
inline float DensityInAtlas(const Ctx& c, uint32_t brick, float3 p_local, int mip, bool linear) {
    const auto& g = c.slot(mip);
    float s = linear ? sampleBrickLinear(g, brick, p_local) : sampleBrickPoint(g, brick, p_local);
    return s * g.compress_scale * c.vd().densityScaleFactorByScaling;
}
// VR/VolumeBase.slang:245-252 FetchEightVoxelsInAtlas at cell `cp` (8 corner voxels cp + {0,1}^3)
inline void FetchEightVoxels(const Ctx& c, uint32_t brick, int3 cp, int mip, float v[8]) {
    const auto& g = c.slot(mip);
    tl_cnt.taps++;
    for (int i = 0; i < 8; i++)
        v[i] = atlasVoxel(g, brick, cp.x + (i % 2), cp.y + (i % 4) / 2, cp.z + i / 4) * g.compress_scale * c.vd().densityScaleFactorByScaling;
}

// VR/VolumeBase.slang:183-232 temperature / emission / velocity
inline float3 ConvertTempToColor(const Ctx& c, float temp) {
    const auto& vd = c.vd();
    temp = fminf(6400.f, (temp - vd.temperatureCutOff) * vd.temperatureScale);
    float queryPoint = (temp - 25) / 6400;
    // Texture1D<float4>, 128 texels, linear, border 0
    float3 rgb = f3(0.f);
    if (!c.p.lut.empty()) {
        float x = queryPoint * 128.f - 0.5f;
        float x0 = floorf(x); float fx = x - x0; int i0 = f2i(x0), i1 = i0 + 1;
        auto tex = [&](int i) -> float3 { if (i < 0 || i > 127) return f3(0.f); const float* t = &c.p.lut[(size_t)i * 4]; return f3(t[0], t[1], t[2]); };
        float3 a = tex(i0), b = tex(i1);
        rgb = f3(lerpf(a.x, b.x, fx), lerpf(a.y, b.y, fx), lerpf(a.z, b.z, fx));
    }
    return vd.LeScale * rgb;
}
inline float Temperature(const Ctx& c, float3 p, bool isLastFrame) {
    int off = isLastFrame ? VRESTIR_PREV_EXTRA_GRID_OFFSET : 0;
    return getValueAtPoint(c.slot(VRESTIR_TEMPERATURE_GRID_ID + off), p, true);
}
inline float3 EmissionWorldSpace(const Ctx& c, float3 pW, bool isLastFrame = false) {
    const auto& vd = c.vd();
    if ((!isLastFrame && !vd.hasEmission) || (isLastFrame && !vd.lastFrameHasEmission)) return f3(0.f);
    float3 pm = WorldToMediumP(c, pW, isLastFrame ? VRESTIR_TEMPERATURE_GRID_ID + VRESTIR_PREV_EXTRA_GRID_OFFSET : VRESTIR_TEMPERATURE_GRID_ID);
    return ConvertTempToColor(c, Temperature(c, pm, isLastFrame));
}
inline float3 VelocityWorld(const Ctx& c, float3 pW, bool isLastFrame = false) {
    int slotId = VRESTIR_VELOCITY_GRID_ID + (isLastFrame ? VRESTIR_PREV_EXTRA_GRID_OFFSET : 0);
    const auto& g = c.slot(slotId);
    float3 pm = WorldToMediumP(c, pW, slotId);
    float3 v = {getValueAtPoint(g, pm, true, 0), getValueAtPoint(g, pm, true, 1), getValueAtPoint(g, pm, true, 2)};
    return mulVec(v, c.vd().externalModelToWorld);
}

// VR/VolumeBase.slang:50-101
inline void CoordinateSystem(float3 v1, float3& v2, float3& v3) {
    if (fabsf(v1.x) > fabsf(v1.y)) v2 = f3(-v1.z, 0, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else v2 = f3(0, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v1, v2);
}
inline float3 SphericalDirection(float sinTheta, float cosTheta, float phi, float3 x, float3 y, float3 z) {
    return sinTheta * cosf(phi) * x + sinTheta * sinf(phi) * y + cosTheta * z;
}
inline float PhaseHG(float cosTheta, float g) {
    float denom = 1 + g * g + 2 * g * cosTheta;
    const float Inv4Pi = 0.07957747154594766788444188168626f;
    return Inv4Pi * (1 - g * g) / (denom * sqrtf(denom));
}
struct MediumInteraction {
    float3 p, wo; float g; bool isValid;
    float phaseFunction(float3 wo_, float3 wi) const { return PhaseHG(dot(wo_, wi), g); }
    float Sample_p(float3 wo_, float3& wi, float2 u) const {
        float cosTheta;
        if (fabsf(g) < 1e-3f) cosTheta = 1 - 2 * u.x;
        else { float sqrTerm = (1 - g * g) / (1 + g - 2 * g * u.x); cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g); }
        float sinTheta = sqrtf(fmaxf(0.f, 1 - cosTheta * cosTheta));
        float phi = 2 * M_PI_F * u.y;
        float3 v1, v2; CoordinateSystem(wo_, v1, v2);
        wi = SphericalDirection(sinTheta, cosTheta, phi, v1, v2, wo_);
        return PhaseHG(cosTheta, g);
    }
    Ray SpawnRay(float3 d) const { return {p, d, 0, kRayTMax}; }
};

// ------------------------------------------------------------------------------------------------ HDDA
// F/Scene/GVDB/gvdbDda.slang:86-157
struct HDDAState {
    float3 pos, dir; int3 pStep; float3 tDel; float3 t; int3 p; float3 tSide; int3 mask;
    void SetFromRay(float3 startPos, float3 startDir, float3 startT) {
        pos = startPos; dir = startDir;
        pStep = {dir.x >= 0 ? 1 : -1, dir.y >= 0 ? 1 : -1, dir.z >= 0 ? 1 : -1};
        t = startT;
    }
    void Prepare(float3 vmin, float vdel) {
        tDel = abs3(f3(vdel) / dir);
        float3 pFlt = (pos + t.x * dir - vmin) / f3(vdel);
        float3 fl = floor3(pFlt);
        tSide = ((fl - pFlt + f3(0.5f)) * toF(pStep) + f3(0.5f)) * tDel + f3(t.x);
        p = {(int)fl.x, (int)fl.y, (int)fl.z};
    }
    void PrepareLeaf(float3 vmin) {
        tDel = abs3(f3(1.0f) / dir);
        float3 pFlt = pos + t.x * dir - vmin;
        float3 fl = floor3(pFlt);
        tSide = ((fl - pFlt + f3(0.5f)) * toF(pStep) + f3(0.5f)) * tDel + f3(t.x);
        p = {(int)fl.x, (int)fl.y, (int)fl.z};
    }
    void Next() {
        mask.x = int((tSide.x < tSide.y) & (tSide.x <= tSide.z));
        mask.y = int((tSide.y < tSide.z) & (tSide.y <= tSide.x));
        mask.z = int((tSide.z < tSide.x) & (tSide.z <= tSide.y));
        t.y = mask.x ? tSide.x : (mask.y ? tSide.y : tSide.z);
    }
    void Step() {
        t.x = t.y;
        // Deliberate deviation from gvdbDda.slang:153 (`tSide += float3(mask) * tDel`): for a ray with an exactly-zero direction
        // component tDel is +inf and 0*inf poisons tSide with NaN, after which every mask is 0 and the traversal spins until
        // the 4096-iteration cap.  The select form is identical for finite tDel and traverses axis-parallel rays correctly.
        tSide = f3(mask.x ? tSide.x + tDel.x : tSide.x, mask.y ? tSide.y + tDel.y : tSide.y, mask.z ? tSide.z + tDel.z : tSide.z);
        p = {p.x + mask.x * pStep.x, p.y + mask.y * pStep.y, p.z + mask.z * pStep.z};
    }
};
inline bool inRange(int3 p, int lo, int hiExclusive) {
    return p.x >= lo && p.y >= lo && p.z >= lo && p.x < hiExclusive && p.y < hiExclusive && p.z < hiExclusive;
}

// VR/VolumeUtils.slang:171-282  VolumeTrackingGVDB<Adapter>
template <class Adapter>
void VolumeTrackingGVDB(const Ctx& c, const Ray& rWorld, int mipLevel, SampleGenerator& sg, Adapter& adapter, bool vertexCenter) {
    tl_cnt.marches++;
    const auto& g = c.slot(mipLevel);
    uint32_t nodeid[3]; float tMax[3]; int b;
    const float epsilon = 0.01f;
    float3 vmin;
    int lev = g.top_lev;
    const int topLev = lev;
    nodeid[lev] = 0;
    Ray ray = WorldToMedium(c, rWorld, mipLevel);
    if (vertexCenter) ray.origin = ray.origin - f3(0.5f);
    float tNear, tFar;
    if (!IntersectVolumeBound(c, ray, tNear, tFar, mipLevel, vertexCenter)) { adapter.ExecuteEndStep(); return; }
    adapter.SetRayInfo(tNear, tFar, ray);
    adapter.ExecuteStartStep();
    float3 tStart = f3(tNear, tFar, 0);
    const vrestir_node* node = &getNode(g, lev, nodeid[lev]); vmin = nodePos(*node);
    tStart.x += epsilon;
    tMax[lev] = tStart.y;
    int iter = 0;
    HDDAState dda;
    dda.SetFromRay(ray.origin, ray.dir, tStart);
    dda.Prepare(vmin, g.vdel[lev]);
    if (vertexCenter) {
        int it = 0;
        while (it++ < 3 && (dda.p.x < 0 || dda.p.y < 0 || dda.p.z < 0 || dda.p.x > g.res[lev] || dda.p.y > g.res[lev] || dda.p.z > g.res[lev])) {
            dda.Next(); dda.Step(); dda.t.x += epsilon;
        }
    }
    float t = tNear;
    for (; iter < 4096 && lev > 0 && lev <= topLev && inRange(dda.p, 0, g.res[lev] + 1); iter++) {
        dda.Next();
        b = (((dda.p.z << g.dim[lev]) + dda.p.y) << g.dim[lev]) + dda.p.x;
        // NOTE: the inclusive bound lets p == res through; the child index then aliases into the list like the shader's
        // ByteAddressBuffer load does.  Guard only against reading outside the whole list (D3D returns 0 there).
        uint32_t childNodeId;
        {
            uint32_t listid = node->link;
            if (listid == ID_UNDEFL) childNodeId = ID_UNDEFL;
            else {
                const uint64_t r3 = (uint64_t)g.res[lev] * g.res[lev] * g.res[lev];
                int64_t idx = (int64_t)listid * (int64_t)r3 + (int64_t)b;
                childNodeId = (idx < 0 || (uint64_t)idx >= g.childlist_count[lev]) ? 0u : g.childlist[lev][idx];
                tl_cnt.nodes++;
            }
        }
        if (childNodeId != ID_UNDEFL) {
            if (lev == 1) {
                nodeid[0] = childNodeId;
                t = dda.t.x - epsilon;
                const vrestir_node& leaf = getNode(g, 0, nodeid[0]);
                float3 vmin_leaf = nodePos(leaf);
                float s = c.vd().densityScaleFactorByScaling;
                float bounds[4] = {leaf.bounds[0] * s, leaf.bounds[1] * s, leaf.bounds[2] * s, leaf.bounds[3] * s};
                if (c.p.dumping && (int)c.p.dumpT.size() < c.p.dumpMax) {
                    auto& pp = const_cast<Pass&>(c.p);
                    pp.dumpXYZ.push_back(leaf.pos[0]); pp.dumpXYZ.push_back(leaf.pos[1]); pp.dumpXYZ.push_back(leaf.pos[2]);
                    pp.dumpT.push_back(dda.t.x);
                }
                bool shouldExit = adapter.ExecuteMainStep(c, dda, vmin_leaf, leaf.link, bounds, mipLevel, t, sg);
                if (shouldExit) return;
                dda.Step();
                dda.t.x += epsilon;
            } else {
                lev--;
                nodeid[lev] = childNodeId;
                node = &getNode(g, lev, nodeid[lev]); vmin = nodePos(*node);
                tMax[lev] = dda.t.y;
                dda.Prepare(vmin, g.vdel[lev]);
            }
        } else {
            dda.Step();
            dda.t.x += epsilon;
        }
        while (lev <= topLev && dda.t.x > tMax[lev]) {   // (shader evaluates tMax[lev] first; lev<=topLev guards the index)
            lev++;
            if (lev <= topLev) {
                node = &getNode(g, lev, nodeid[lev]); vmin = nodePos(*node);
                dda.Prepare(vmin, g.vdel[lev]);
            }
        }
    }
    adapter.ExecuteEndStep();
}

// cubic coefficients of sigma_t along the ray inside one trilinear cell (VR/VolumeTrackingAdapterGVDB.slang:64-97, 287-321)
inline void trilinearCubic(const float v[8], float sigma_t, float3 d, float3 p0, float& c3, float& c2, float& c1, float& c0) {
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
    float tNear = 0, tFar = 0; Ray ray{};
    void SetRayInfo(float a, float b, const Ray& r) { tNear = a; tFar = b; ray = r; }
};

// VR/VolumeTrackingAdapterGVDB.slang:20-136
struct MediumTrAnalyticAdapter : AdapterBase {
    float Tr = 0.f; bool useLinearSampler = false;
    void ExecuteStartStep() {}
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float*, int mip, float& t, SampleGenerator&) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        const float sigma_t_ = c.vd().sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, 0, c.slot(mip).res[0]); iter++) {
            leaf.Next();
            float maxDeltaT = leaf.t.y - t;
            if (useLinearSampler) {
                float v[8]; FetchEightVoxels(c, brick, leaf.p, mip, v);
                float3 d = ray.dir;
                float3 p0 = leaf.pos + leaf.t.x * leaf.dir - (toF(leaf.p) + vmin_leaf);
                float c3, c2, c1, c0; trilinearCubic(v, sigma_t_, d, p0, c3, c2, c1, c0);
                float t_dist = fminf(tFar - t, maxDeltaT);
                float t2 = t_dist * t_dist, t3 = t2 * t_dist, t4 = t2 * t2;
                Tr += -(c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t_dist);
            } else {
                float density = DensityInAtlas(c, brick, toF(leaf.p) + f3(0.5f), mip, false);
                float sigma_t = density * sigma_t_;
                Tr += -fminf(tFar - t, maxDeltaT) * sigma_t;
            }
            if (t + maxDeltaT >= tFar) { Tr = expf(Tr); return true; }
            t += maxDeltaT;
            leaf.Step();
        }
        return false;
    }
    void ExecuteEndStep() { Tr = expf(Tr); }
};

// VR/VolumeTrackingAdapterGVDB.slang:140-208
struct MediumTrRayMarchingAdapter : AdapterBase {
    float Tr = 0.f; bool useLinearSampler = false; float tStep = 0; bool hasInitialized = false;
    void ExecuteStartStep() { hasInitialized = true; }
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float*, int mip, float& t, SampleGenerator&) {
        t = tNear + (floorf((t - tNear) / tStep) + 0.5f) * tStep;
        if (t < dda.t.x) t += tStep;
        const float tStepMultipler = 1.f;
        float3 wp = ray.origin + t * ray.dir;
        float3 p = wp - vmin_leaf;
        const float3 wpt = tStepMultipler * tStep * ray.dir;
        const float res = (float)c.slot(mip).res[0];
        for (int iter = 0; iter < MAX_BRICK_STEPS && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res && p.y < res && p.z < res; iter++) {
            if (t >= tFar) { Tr = expf(Tr); return true; }
            float density = DensityInAtlas(c, brick, p, mip, useLinearSampler);
            float sigma_t = density * c.vd().sigma_t;
            Tr += -sigma_t * (iter == 0 ? 1.f : tStepMultipler) * tStep;
            p = p + wpt;
            t += tStepMultipler * tStep;
        }
        return false;
    }
    void ExecuteEndStep() { if (hasInitialized) Tr = expf(Tr); else Tr = 1.f; }
};

// VR/VolumeTrackingAdapterGVDB.slang:212-436
struct SampleMediumAnalyticAdapter : AdapterBase {
    float hitDistances[4] = {0, 0, 0, 0}, outTr[4] = {0, 0, 0, 0}, pdf[4] = {0, 0, 0, 0}; int numSamples = 0; float opticalThickness = 0.f;
    bool hasInitialized = false, useLinearSampler = false;
    void ExecuteStartStep() { for (int i = 0; i < numSamples; i++) hitDistances[i] = -1; hasInitialized = true; }
    static float tauOf(float t, float c3, float c2, float c1, float c0) {
        float t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
        return c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t;
    }
    static float sigmaOf(float t, float c3, float c2, float c1, float c0) {
        float t2 = t * t, t3 = t2 * t;
        return c3 * t3 + c2 * t2 + c1 * t + c0;
    }
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float*, int mip, float& t, SampleGenerator& sg) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        const float sigma_t_ = c.vd().sigma_t;
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, 0, c.slot(mip).res[0]); iter++) {
            leaf.Next();
            float maxDeltaT = fminf(tFar - t, leaf.t.y - t);
            float currentTMax = fminf(tFar, leaf.t.y);
            int finishedCount = 0;
            float deltaThickness = 0.f;
            if (useLinearSampler) {
                float v[8]; FetchEightVoxels(c, brick, leaf.p, mip, v);
                float3 d = ray.dir;
                float3 p0 = leaf.pos + leaf.t.x * leaf.dir - (toF(leaf.p) + vmin_leaf);
                float c3, c2, c1, c0; trilinearCubic(v, sigma_t_, d, p0, c3, c2, c1, c0);
                deltaThickness = tauOf(maxDeltaT, c3, c2, c1, c0);
                for (int s = 0; s < numSamples; s++) {
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
                float density = DensityInAtlas(c, brick, toF(leaf.p) + f3(0.5f), mip, false);
                float sigma_t = density * sigma_t_;
                for (int s = 0; s < numSamples; s++) {
                    if (hitDistances[s] == -1) {
                        float dT = -logf(1 - sampleNext1D(sg)) / sigma_t;
                        float curT = t + dT;
                        if (std::isnan(curT) || std::isinf(curT)) curT = kRayTMax;
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
    void ExecuteEndStep() {
        if (hasInitialized) {
            for (int s = 0; s < numSamples; s++)
                if (hitDistances[s] == -1) { hitDistances[s] = kRayTMax; outTr[s] = expf(-opticalThickness); pdf[s] = outTr[s]; }
        } else {
            for (int s = 0; s < numSamples; s++) { hitDistances[s] = kRayTMax; outTr[s] = 1.f; pdf[s] = 1.f; }
        }
    }
};

// VR/VolumeTrackingAdapterGVDB.slang:439-515
struct SampleVolumeCellByDensityAdapter : AdapterBase {
    int densityBound = 0; float2 selectedInterval = {-1, -1}; float runningSum = 0.f, Tr = 0.f;
    void ExecuteStartStep() { selectedInterval = {-1, -1}; }
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float*, int mip, float& t, SampleGenerator& sg) {
        HDDAState leaf = dda;
        leaf.PrepareLeaf(vmin_leaf);
        for (int iter = 0; iter < MAX_BRICK_STEPS && inRange(leaf.p, 0, c.slot(mip).res[0]); iter++) {
            float density = DensityInAtlas(c, brick, toF(leaf.p) + f3(0.5f), mip, false);
            leaf.Next();
            float maxDeltaT = leaf.t.y - t;
            float sigma_t = density * c.vd().sigma_t;
            float weight = expf(Tr) * sigma_t;
            runningSum += weight;
            if (runningSum > 0.f && sampleNext1D(sg) < weight / runningSum) {
                densityBound = f2i(density);
                selectedInterval = {t, fminf(tFar, t + maxDeltaT)};
            }
            Tr += -maxDeltaT * sigma_t;
            if (t + maxDeltaT >= tFar) return true;
            t += maxDeltaT;
            leaf.Step();
        }
        return false;
    }
    void ExecuteEndStep() {}
};

// VR/VolumeTrackingAdapterGVDB.slang:518-602
struct ReservoirFeatureRayMarchingAdapter : AdapterBase {
    float accuTransmittance = 1.f, Tr = 0.f, tStep = 0; bool useLinearSampler = true, hasInitialized = false;
    void ExecuteStartStep() { hasInitialized = true; }
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float*, int mip, float& t, SampleGenerator&) {
        t = tNear + (floorf((t - tNear) / tStep) + 0.5f) * tStep;
        if (t < dda.t.x) t += tStep;
        float3 wp = ray.origin + t * ray.dir;
        float3 p = wp - vmin_leaf;
        const float3 wpt = tStep * ray.dir;
        const float res = (float)c.slot(mip).res[0];
        for (int iter = 0; iter < MAX_BRICK_STEPS && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res && p.y < res && p.z < res; iter++) {
            float curTransmittance = expf(Tr);
            if (curTransmittance < 0.01f) { ExecuteEndStep(); return true; }
            float density = DensityInAtlas(c, brick, p, mip, useLinearSampler);
            float sigma_t = density * c.vd().sigma_t;
            float negOpticalLength = -sigma_t * tStep;
            Tr += negOpticalLength;
            p = p + wpt;
            t += tStep;
        }
        return false;
    }
    void ExecuteEndStep() { if (hasInitialized) accuTransmittance = expf(Tr); }
};

// VR/VolumeTrackingAdapterGVDB.slang:606-707
struct DecompositionTrackingAdapter : AdapterBase {
    Ray rWorld{}; MediumInteraction mi{};
    void ExecuteStartStep() {}
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float* bounds, int mip, float& t, SampleGenerator& sg) {
        const float sigma_t_ = c.vd().sigma_t, g = c.vd().PhaseFunctionConstantG;
        float minDensity = bounds[0], maxDensity = bounds[1];
        float currentTMax = fminf(tFar, dda.t.y);
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
                float density = DensityInAtlas(c, brick, p, mip, true);
                if ((density - minDensity) * invMaxDensity > sampleNext1D(sg)) {
                    mi = {rWorld.at(t), -rWorld.dir, g, true};
                    return true;
                }
            }
            t = fminf(t_control, t);
            if (t < currentTMax) { mi = {rWorld.at(t), -rWorld.dir, g, true}; return true; }
            t = currentTMax;
            if (t >= tFar) { ExecuteEndStep(); return true; }
        } else {
            if (t_control < currentTMax) { mi = {rWorld.at(t_control), -rWorld.dir, g, true}; return true; }
            else { t = currentTMax; return false; }
        }
        return false;
    }
    void ExecuteEndStep() { mi.isValid = false; }
};

// VR/VolumeTrackingAdapterGVDB.slang:710-794
struct ResidualRatioTrackingAdapter : AdapterBase {
    float Tr = 1.f; bool useAnalogResidual = false, useGlobalMajorant = false;
    void ExecuteStartStep() { if (useGlobalMajorant) useAnalogResidual = true; }
    bool ExecuteMainStep(const Ctx& c, const HDDAState& dda, float3 vmin_leaf, uint32_t brick, const float* bounds, int mip, float& t, SampleGenerator& sg) {
        const float sigma_t_ = c.vd().sigma_t;
        float mu_min = useGlobalMajorant ? 0.f : bounds[0] * sigma_t_;
        float mu_max = useGlobalMajorant ? GetVolumeMaxDensity(c, mip) * sigma_t_ : bounds[1] * sigma_t_;
        float mu_avg = bounds[2] * sigma_t_;
        float maxDeltaT = fminf(tFar - t, dda.t.y - t);
        float currentTMax = fminf(tFar, dda.t.y);
        float mu_r_temp = mu_max - mu_min;
        float D = c.vd().superVoxelWorldSpaceDiagonalLength;
        float gamma = 2;
        float mu_c = (mu_r_temp == 0.f || useAnalogResidual) ? mu_min
                     : fminf(mu_avg, fmaxf(mu_min, mu_min + mu_r_temp * (powf(gamma, 1.f / (D * mu_r_temp)) - 1)));
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
                float density = DensityInAtlas(c, brick, p, mip, true);
                float mu = density * sigma_t_;
                T_r *= 1 - (mu - mu_c) * inv_mu_r;
            }
        }
        Tr *= T_c * T_r;
        t = currentTMax;
        if (t >= tFar) return true;
        return false;
    }
    void ExecuteEndStep() {}
};

// ---- generic wrappers: VR/VolumeUtils.slang:284-378 ----
inline float MediumTrAnalyticGeneric(const Ctx& c, const Ray& r, int mip, SampleGenerator& sg, bool linear) {
    MediumTrAnalyticAdapter a; a.useLinearSampler = linear;
    VolumeTrackingGVDB(c, r, mip, sg, a, linear);
    return a.Tr;
}
inline void SampleMediumAnalyticGeneric(const Ctx& c, const Ray& r, SampleGenerator& sg, bool linear, float hit[4], int mip, float pdf[4], float outTr[4], int numSamples) {
    SampleMediumAnalyticAdapter a; a.numSamples = numSamples; a.useLinearSampler = linear;
    if (linear) for (int i = 0; i < numSamples; i++) a.outTr[i] = -logf(1 - sampleNext1D(sg));
    VolumeTrackingGVDB(c, r, mip, sg, a, linear);
    for (int i = 0; i < 4; i++) { hit[i] = a.hitDistances[i]; outTr[i] = a.outTr[i]; pdf[i] = a.pdf[i]; }
}
inline void SampleMediumSuperVoxelGeneric(const Ctx& c, const Ray& r, SampleGenerator& sg, MediumInteraction& mi, int mip) {
    DecompositionTrackingAdapter a; a.rWorld = r; a.mi = mi;
    VolumeTrackingGVDB(c, r, mip, sg, a, false);
    mi = a.mi;
}
inline float MediumTrResidualRatioTrackingGeneric(const Ctx& c, const Ray& r, int mip, SampleGenerator& sg, bool analog, bool global) {
    ResidualRatioTrackingAdapter a; a.useGlobalMajorant = global; a.useAnalogResidual = analog || global;
    VolumeTrackingGVDB(c, r, mip, sg, a, false);
    return a.Tr;
}
inline float MediumTrRayMarchingGeneric(const Ctx& c, const Ray& r, int mip, bool linear, float tStepScale, SampleGenerator& sg) {
    int eff = mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : mip;
    eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
    MediumTrRayMarchingAdapter a;
    a.tStep = c.vd().tStep * c.vd().volumeWorldScaling * tStepScale * (eff + 1);
    a.useLinearSampler = linear;
    VolumeTrackingGVDB(c, r, mip, sg, a, false);
    return a.Tr;
}
inline void ReservoirFeatureRayMarchingGeneric(const Ctx& c, const Ray& r, SampleGenerator& sg, int mip, bool linear, float tStepScale, float& accu) {
    ReservoirFeatureRayMarchingAdapter a;
    a.tStep = c.vd().tStep * c.vd().volumeWorldScaling * tStepScale;
    a.useLinearSampler = linear;
    VolumeTrackingGVDB(c, r, mip, sg, a, false);
    accu = a.accuTransmittance;
}
// VR/VolumeUtils.slang:553-571
inline float computeVisibility(const Ctx& c, const Ray& ray, SampleGenerator& sg, int visibilitySamples, int mip, bool linear, uint32_t method, float tStepScale = 1.f) {
    float visibility = 0;
    for (int i = 0; i < visibilitySamples; i++) {
        if (method == VRESTIR_RATIO_TRACKING || method == VRESTIR_RESIDUAL_RATIO_TRACKING || method == VRESTIR_ANALOG_RESIDUAL_RATIO_TRACKING)
            visibility += MediumTrResidualRatioTrackingGeneric(c, ray, mip, sg, method == VRESTIR_ANALOG_RESIDUAL_RATIO_TRACKING, method == VRESTIR_RATIO_TRACKING);
        else if (method == VRESTIR_RAY_MARCHING) visibility += MediumTrRayMarchingGeneric(c, ray, mip, linear, tStepScale, sg);
        else if (method == VRESTIR_ANALYTIC_TRACKING) visibility += MediumTrAnalyticGeneric(c, ray, mip, sg, linear);
    }
    visibility /= visibilitySamples;
    return visibility;
}
// VR/VolumeUtils.slang:342-348,573-582
inline float RejectionSampleRandomPointByDensity(const Ctx& c, const Ray& r, SampleGenerator& sg, int mip) {
    SampleVolumeCellByDensityAdapter a;
    VolumeTrackingGVDB(c, r, mip, sg, a, false);
    float2 sel = a.selectedInterval;
    if (sel.x == -1) return kRayTMax;
    float sampledDepth = sel.x + (sel.y - sel.x) * sampleNext1D(sg);
    float sampledY = sampleNext1D(sg) * (float)a.densityBound; (void)sampledY;
    return sampledDepth;
}

// ------------------------------------------------------------------------------------------------ lights
// F/Utils/Math/MathHelpers.slang:92-107
inline float2 world_to_latlong_map(float3 dir) {
    float3 p = normalize(dir);
    float2 uv; uv.x = atan2f(p.x, -p.z) * M_1_2PI_F + 0.5f; uv.y = acosf(p.y) * M_1_PI_F;
    return uv;
}
// F/Utils/Math/MathHelpers.slang:200-224
inline float3 oct_to_ndir_equal_area_unorm(float2 p) {
    p.x = p.x * 2.f - 1.f; p.y = p.y * 2.f - 1.f;
    float d = 1.f - (fabsf(p.x) + fabsf(p.y));
    float r = 1.f - fabsf(d);
    float phi = (r > 0.f) ? ((fabsf(p.y) - fabsf(p.x)) / r + 1.f) * M_PI_4_F : 0.f;
    float f = r * sqrtf(2.f - r * r);
    float x = f * fsign(p.x) * cosf(phi);
    float y = f * fsign(p.y) * sinf(phi);
    float z = fsign(d) * (1.f - r * r);
    return f3(x, y, z);
}
struct EnvMapSample { float3 dir; float pdf; float3 Le; };

inline float3 envTexel(const Pass& p, int x, int y) {
    const float* t = &p.envTexels[((size_t)y * p.envW + x) * 4];
    return f3(t[0], t[1], t[2]);
}
inline float3 envBilinear(const Pass& p, float2 uv) {  // wrap U, clamp V (F/.../EnvMap.cpp:105-110)
    float x = uv.x * (float)p.envW - 0.5f, y = uv.y * (float)p.envH - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = f2i(x0f), y0 = f2i(y0f), x1 = x0 + 1, y1 = y0 + 1;
    x0 %= p.envW; if (x0 < 0) x0 += p.envW;
    x1 %= p.envW; if (x1 < 0) x1 += p.envW;
    y0 = std::min(std::max(y0, 0), p.envH - 1); y1 = std::min(std::max(y1, 0), p.envH - 1);
    float3 a = envTexel(p, x0, y0), b = envTexel(p, x1, y0), cc = envTexel(p, x0, y1), d = envTexel(p, x1, y1);
    float3 top = f3(lerpf(a.x, b.x, fx), lerpf(a.y, b.y, fx), lerpf(a.z, b.z, fx));
    float3 bot = f3(lerpf(cc.x, d.x, fx), lerpf(cc.y, d.y, fx), lerpf(cc.z, d.z, fx));
    return f3(lerpf(top.x, bot.x, fy), lerpf(top.y, bot.y, fy), lerpf(top.z, bot.z, fy));
}
// F/Experimental/Scene/Lights/EnvMap.slang:42-78
inline float3 envToLocal(const Pass& p, float3 dir, bool last = false) { return mulVec3x3(dir, last ? p.envPrevInvT : p.envInvT); }
inline float3 envToWorld(const Pass& p, float3 dir, bool last = false) { return mulVec3x3(dir, last ? p.envPrevT : p.envT); }
inline float3 envEval(const Pass& p, float3 dir, bool last = false) {
    if (!p.haveEnv) return f3(0.f);
    float2 uv = world_to_latlong_map(envToLocal(p, dir, last));
    return p.envIntensity * p.envTint * envBilinear(p, uv);
}
// F/Experimental/Scene/Lights/EnvMapSampler.slang:75-90
inline float2 envEncodeLightUV(const Pass& p, float3 dir, int& lightID) {
    dir = envToLocal(p, dir);
    lightID = dir.z < 0 ? -2 : -1;
    return {dir.x, dir.y};
}
inline float3 envDecodeLightUV(const Pass& p, float2 uv, int lightID, bool last) {
    float3 dir = f3(uv.x, uv.y, 0);
    dir.z = sqrtf(1 - uv.x * uv.x - uv.y * uv.y);
    if (std::isnan(dir.z)) dir.z = 0.f;
    if (lightID == -2) dir.z = -dir.z;
    return envToWorld(p, dir, last);
}
inline float impLoad(const Pass& p, uint32_t x, uint32_t y, int mip) {
    int dim = p.impDim >> mip;
    if ((int)x >= dim || (int)y >= dim) return 0.f;
    return p.importance[p.impOffset[mip] + (size_t)y * dim + x];
}
// F/Experimental/Scene/Lights/EnvMapSampler.slang:94-167 (hierarchical warp) or alias table over the finest mip
inline bool envSample(const Pass& p, float2 rnd, EnvMapSample& result) {
    float2 pp = rnd; uint32_t posx = 0, posy = 0;
    if (p.envSamplerType == VRESTIR_ENV_SAMPLER_ALIAS && !p.envAliasThr.empty()) {
        // alias pick over dim^2 texels; the residual of rnd.x re-used as the sub-texel x
        const uint32_t count = (uint32_t)p.envAliasThr.size();
        float xs = rnd.x * (float)count;
        uint32_t index = std::min(count - 1, f2u(xs));
        float xi = xs - (float)index;
        float thr = p.envAliasThr[index];
        uint32_t texel;
        if (xi < thr) { texel = index; pp.x = xi / thr; } else { texel = p.envAliasRedirect[index]; pp.x = (xi - thr) / (1.f - thr); }
        posx = texel % (uint32_t)p.impDim; posy = texel / (uint32_t)p.impDim;
    } else {
        for (int mip = p.impBaseMip - 1; mip >= 0; mip--) {
            posx *= 2; posy *= 2;
            float w[4];
            w[0] = impLoad(p, posx, posy, mip); w[1] = impLoad(p, posx + 1, posy, mip);
            w[2] = impLoad(p, posx, posy + 1, mip); w[3] = impLoad(p, posx + 1, posy + 1, mip);
            float q[2]; q[0] = w[0] + w[2]; q[1] = w[1] + w[3];
            uint32_t offx, offy;
            float d = q[0] / (q[0] + q[1]);
            if (pp.x < d) { offx = 0; pp.x = pp.x / d; } else { offx = 1; pp.x = (pp.x - d) / (1.f - d); }
            float e = w[offx] / q[offx];
            if (pp.y < e) { offy = 0; pp.y = pp.y / e; } else { offy = 1; pp.y = (pp.y - e) / (1.f - e); }
            posx += offx; posy += offy;
        }
    }
    float invDim = 1.f / (float)p.impDim;
    float2 uv = {((float)posx + pp.x) * invDim, ((float)posy + pp.y) * invDim};
    float3 dir = oct_to_ndir_equal_area_unorm(uv);
    float avg_w = impLoad(p, 0, 0, p.impBaseMip);
    float pdf = impLoad(p, posx, posy, 0) / avg_w;
    result.dir = envToWorld(p, dir);
    result.pdf = pdf * M_1_4PI_F;
    result.Le = envEval(p, result.dir);
    return true;
}

struct SceneLightSample { float3 dir; float distance; float3 Li; float pdf, pdfArea; float3 rayDir; float rayDistance; };
struct AnalyticLightSample { float3 posW, normalW, dir; float distance; float3 Li; float pdf; };

// F/Experimental/Scene/Lights/LightHelpers.slang:196-274
inline bool sampleLight(float3 shadingPosW, const vrestir_light& light, SampleGenerator&, AnalyticLightSample& ls) {
    const float kMinLightDistSqr = 1e-9f;
    float3 I = f3(light.intensity[0], light.intensity[1], light.intensity[2]);
    float3 dirW = f3(light.dirW[0], light.dirW[1], light.dirW[2]);
    if (light.type == VRESTIR_LIGHT_POINT) {
        ls.posW = f3(light.posW[0], light.posW[1], light.posW[2]); ls.normalW = dirW;
        float3 toLight = ls.posW - shadingPosW;
        float distSqr = fmaxf(dot(toLight, toLight), kMinLightDistSqr);
        ls.distance = sqrtf(distSqr);
        ls.dir = toLight / ls.distance;
        ls.Li = I / distSqr;
        ls.pdf = 0.f;
        return true;
    } else if (light.type == VRESTIR_LIGHT_DIRECTIONAL) {
        ls.posW = f3(0.f); ls.normalW = dirW;
        ls.distance = FLT_MAX; ls.dir = -dirW; ls.Li = I; ls.pdf = 0.f;
        return true;
    }
    return false;
}

// F/Utils/Helpers.slang:62-103
inline float3 computeRayOrigin(float3 pos, float3 normal) {
    const float origin = 1.f / 32.f, fScale = 1.f / 65536.f, iScale = 256.f;
    int iOff[3] = {f2i(normal.x * iScale), f2i(normal.y * iScale), f2i(normal.z * iScale)};
    float P[3] = {pos.x, pos.y, pos.z}, N[3] = {normal.x, normal.y, normal.z}, out[3];
    for (int i = 0; i < 3; i++) {
        int32_t bits; memcpy(&bits, &P[i], 4);
        bits += (P[i] < 0.f ? -iOff[i] : iOff[i]);
        float iPos; memcpy(&iPos, &bits, 4);
        float fOff = N[i] * fScale;
        out[i] = fabsf(P[i]) < origin ? P[i] + fOff : iPos;
    }
    return f3(out[0], out[1], out[2]);
}
struct TriangleLightSample { uint32_t triangleIndex; float3 posW, normalW, dir; float distance; float3 Le; float pdf, pdfArea, cosTheta; float2 uv; };
// F/Utils/Math/MathHelpers.slang:242-250
inline float3 sample_triangle(float2 u) { float su = sqrtf(u.x); float2 b = {1.f - su, u.y * su}; return f3(1.f - b.x - b.y, b.x, b.y); }
// F/Experimental/Scene/Lights/EmissiveLightSamplerHelpers.slang:56-101
inline bool sampleTriangle(const Pass& p, float3 posW, uint32_t triangleIndex, float2 u, TriangleLightSample& ls) {
    ls = TriangleLightSample{};
    ls.triangleIndex = triangleIndex;
    const auto& tri = p.tris[triangleIndex];
    float3 bc = sample_triangle(u);
    ls.uv = u;
    float3 p0 = f3(tri.posW[0][0], tri.posW[0][1], tri.posW[0][2]), p1 = f3(tri.posW[1][0], tri.posW[1][1], tri.posW[1][2]), p2 = f3(tri.posW[2][0], tri.posW[2][1], tri.posW[2][2]);
    float3 n = f3(tri.normal[0], tri.normal[1], tri.normal[2]);
    ls.posW = p0 * bc.x + p1 * bc.y + p2 * bc.z;
    ls.posW = computeRayOrigin(ls.posW, n);
    float3 toLight = ls.posW - posW;
    const float distSqr = fmaxf(FLT_MIN, dot(toLight, toLight));
    ls.distance = sqrtf(distSqr);
    ls.dir = toLight / ls.distance;
    ls.normalW = n;
    float cosTheta = dot(ls.normalW, -ls.dir);
    if (cosTheta <= 0.f) return false;
    ls.Le = f3(tri.Le[0], tri.Le[1], tri.Le[2]);
    float denom = fmaxf(FLT_MIN, cosTheta * tri.area);
    ls.pdf = distSqr / denom;
    ls.cosTheta = cosTheta;
    ls.pdfArea = 1.f / tri.area;
    return true;
}
// F/Experimental/Scene/Lights/EmissivePowerSampler.slang:45-91, F/Utils/Sampling/AliasTable.slang:31-80
inline bool emissiveSampleLight(const Pass& p, float3 posW, SampleGenerator& sg, TriangleLightSample& ls) {
    if (p.tris.empty()) return false;
    float2 rnd = sampleNext2D(sg);
    uint32_t count = (uint32_t)p.alias.size();
    uint32_t index = std::min(count - 1, f2u(rnd.x * (float)count));
    const AliasItem& item = p.alias[index];
    uint32_t triangleIndex = rnd.y >= item.threshold ? item.indexA : item.indexB;
    float triangleSelectionPdf = p.aliasWeights[triangleIndex] / p.aliasWeightSum;
    float2 u = sampleNext2D(sg);
    if (!sampleTriangle(p, posW, triangleIndex, u, ls)) return false;
    ls.pdf *= triangleSelectionPdf;
    ls.pdfArea *= triangleSelectionPdf;
    return true;
}
inline bool getEmissiveLightSample(const Pass& p, float3 posW, int lightID, float2 uv, TriangleLightSample& ls) {
    if (p.tris.empty()) return false;
    return sampleTriangle(p, posW, (uint32_t)lightID, uv, ls);
}

// VR/VolumeUtils.slang:12-149
inline bool sampleSceneLights(const Ctx& c, float3 rayOrigin, bool kEnv, bool kAnalytic, bool kEmissive, SampleGenerator& sg,
                              SceneLightSample& ls, int& outLightIndex, float2& outLightUV) {
    const Pass& P = c.p;
    if (!kEnv && !kAnalytic && !kEmissive) return false;
    float p[3] = {kEnv ? 1.f : 0.f, kAnalytic ? 1.f : 0.f, kEmissive ? 1.f : 0.f};
    float sum = p[0] + p[1] + p[2];
    if (sum == 0.f) return false;
    float invSum = 1.f / sum;
    p[0] *= invSum; p[1] *= invSum; p[2] *= invSum;
    float u = sampleNext1D(sg);
    if (kEnv) {
        if (u < p[0]) {
            float selectionPdf = p[0];
            EnvMapSample lightSample;
            envSample(P, sampleNext2D(sg), lightSample);
            float pdf = selectionPdf * lightSample.pdf;
            ls.rayDir = ls.dir = lightSample.dir;
            ls.rayDistance = ls.distance = kRayTMax;
            ls.pdf = pdf; ls.pdfArea = pdf;
            ls.Li = pdf > 0.f ? lightSample.Le / pdf : f3(0.f);
            outLightIndex = -1;
            outLightUV = envEncodeLightUV(P, ls.rayDir, outLightIndex);
            return !(std::isnan(ls.rayDir.x) || std::isnan(ls.rayDir.y) || std::isnan(ls.rayDir.z));
        }
        u -= p[0];
    }
    if (kAnalytic) {
        if (u < p[1]) {
            u /= p[1];
            uint32_t lightCount = (uint32_t)P.lights.size();
            uint32_t lightIndex = std::min(f2u(u * (float)lightCount), lightCount - 1);
            float selectionPdf = p[1] / (float)lightCount;
            AnalyticLightSample lightSample{};
            sampleLight(rayOrigin, P.lights[lightIndex], sg, lightSample);
            outLightIndex = (int)lightIndex;
            ls.rayDir = ls.dir = lightSample.dir;
            ls.rayDistance = ls.distance = lightSample.distance;
            if (lightSample.pdf == 0) lightSample.pdf = 1.f;
            ls.pdf = selectionPdf * lightSample.pdf;
            ls.pdfArea = ls.pdf;
            ls.Li = lightSample.Li / ls.pdf;
            outLightUV = {0, 0};
            return true;
        }
        u -= p[1];
    }
    if (kEmissive) {
        if (u < p[2]) {
            float selectionPdf = p[2];
            TriangleLightSample lightSample{};
            bool valid = emissiveSampleLight(P, rayOrigin, sg, lightSample);
            float pdf = selectionPdf * lightSample.pdf;
            float pdfArea = selectionPdf * lightSample.pdfArea;
            float3 offsetPos = computeRayOrigin(lightSample.posW, lightSample.normalW);
            float3 toLight = offsetPos - rayOrigin;
            ls.rayDistance = length(toLight);
            ls.rayDir = normalize(toLight);
            ls.dir = lightSample.dir; ls.distance = lightSample.distance;
            ls.pdf = pdf; ls.pdfArea = pdfArea;
            ls.Li = pdf > 0.f ? lightSample.Le * P.emissiveMul / pdf : f3(0.f);
            outLightIndex = (int)lightSample.triangleIndex + (int)P.lights.size();
            outLightUV = lightSample.uv;
            if (!valid) return false;
            return true;
        }
        u -= p[2];
    }
    return false;
}
// VR/VolumeUtils.slang:151-169
inline SceneLightSample getAnalyticalLightSample(const Ctx& c, int lightIndex, float3 rayOrigin, SampleGenerator& sg) {
    AnalyticLightSample lightSample{};
    sampleLight(rayOrigin, c.p.lights[lightIndex], sg, lightSample);
    SceneLightSample ls{};
    ls.rayDir = ls.dir = lightSample.dir;
    ls.rayDistance = ls.distance = lightSample.distance;
    ls.pdf = lightSample.pdf; ls.Li = lightSample.Li;
    return ls;
}
// VR/VolumeUtils.slang:419-452 directLighting (used by the reference path tracer)
inline float3 directLighting(const Ctx& c, SampleGenerator& sg, const MediumInteraction& mi, int kLightSamplesPerVertex, bool kEnv, bool kAnalytic, bool kEmissive,
                             bool enableShadow, int mip, uint32_t trackingMethod) {
    float3 Ld = f3(0.f);
    for (int i = 0; i < kLightSamplesPerVertex; i++) {
        SceneLightSample ls{}; int idx; float2 uv;
        bool valid = sampleSceneLights(c, mi.p, kEnv, kAnalytic, kEmissive, sg, ls, idx, uv);
        if (!valid) continue;
        ls.pdf = 1.f;
        Ray shadowRay = {mi.p, ls.rayDir, 0, ls.rayDistance};
        if (enableShadow) { float vis = computeVisibility(c, shadowRay, sg, 1, mip, true, trackingMethod, 1.f); ls.Li *= vis; }
        float ph = mi.phaseFunction(mi.wo, ls.dir);
        Ld += ph * ls.Li / ls.pdf;
    }
    Ld = Ld / (float)kLightSamplesPerVertex;
    return Ld;
}
// VR/VolumeUtils.slang:454-492
inline float3 SampleDirectLighting(const Ctx& c, SampleGenerator& sg, float& pdf, const MediumInteraction& mi, bool kEnv, bool kAnalytic, bool kEmissive, bool enableShadow,
                                   int mip, uint32_t trackingMethod, bool linear, int lightVisibilitySamples, float tStepScale, int& outLightIndex, float2& outLightUV, float& outVisibility) {
    pdf = 0.f;
    float3 Ld = f3(0.f);
    SceneLightSample ls{};
    bool valid = sampleSceneLights(c, mi.p, kEnv, kAnalytic, kEmissive, sg, ls, outLightIndex, outLightUV);
    pdf = ls.pdfArea;
    if (!valid) { pdf = 0.f; return f3(0.f); }
    ls.pdf = 1.f;
    Ray shadowRay = {mi.p, ls.rayDir, 0, ls.rayDistance};
    if (enableShadow) {
        float visibility = computeVisibility(c, shadowRay, sg, lightVisibilitySamples, mip, linear, trackingMethod, tStepScale);
        outVisibility = visibility;
        ls.Li *= visibility;
    }
    float ph = mi.phaseFunction(mi.wo, ls.dir);
    Ld += ph * ls.Li / ls.pdf;
    return Ld;
}

// ------------------------------------------------------------------------------------------------ camera
// F/Scene/Camera/Camera.slang:160-228.  computeNonNormalizedRayDirPinhole (jitter = 0).  The "Scaled(1, 0.5)" spelling
// reduces to the same expression: pixelPos = pixel + 0.5, int2(pixelPos) = pixel, pixelPos - int2(pixelPos) = 0.5.
inline float3 camRayDirNN(float3 U, float3 V, float3 Wv, int px, int py, int W, int H) {
    float2 p = {((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H};
    float2 ndc = {2.f * p.x + -1.f, -2.f * p.y + 1.f};
    return ndc.x * U + ndc.y * V + Wv;
}
inline float3 v3(const float* a) { return f3(a[0], a[1], a[2]); }

// ------------------------------------------------------------------------------------------------ ReSTIRHelper.slang
// VR/ReSTIRHelper.slang:10-89
inline float3 decodeEmissivePosition(int lightID, float2 lightUV) { float z; memcpy(&z, &lightID, 4); return f3(lightUV.x, lightUV.y, z); }
inline void encodeEmissivePosition(float3 pos, int& lightID, float2& lightUV) { memcpy(&lightID, &pos.z, 4); lightUV = {pos.x, pos.y}; }
struct float4_ { float x, y, z, w; };
inline float4_ decodeWiDist(float3 in, bool reuseAsVertex = false) {
    if (reuseAsVertex) {   // VERTEX_REUSE: the record holds a world-space vertex (w = -1) or "left the medium"
        if (in.x == kRayTMax) return {0.f, 0.f, 0.f, kRayTMax};
        return {in.x, in.y, in.z, -1.f};
    }
    float3 wi; wi.x = in.x; wi.y = in.y;
    wi.z = sqrtf(1 - (in.x * in.x + in.y * in.y));
    if (std::isnan(wi.z)) wi.z = 0.f;
    if (in.z < 0) { wi.z = -wi.z; in.z = -in.z; }
    return {wi.x, wi.y, wi.z, in.z};
}
inline float3 encodeWiDist(float4_ in) { float3 wi; wi.x = in.x; wi.y = in.y; wi.z = in.w; if (in.z < 0) wi.z = -wi.z; return wi; }
inline int decodeMaxIndirectBounces(int storage, int maxBounces) { return maxBounces == 1 ? 0 : storage >> 20; }
inline int encodeMaxIndirectBounces(int storage, int bounce) { return (int)(((uint32_t)bounce << 20) | ((uint32_t)storage & 0xFFFFFu)); }
inline int decodePathTag(int storage) { return (storage >> 16) & 0xF; }
inline int encodePathTag(int storage, int tag) { return (int)(((uint32_t)tag << 16) | ((uint32_t)storage & 0xFFF0FFFFu)); }

// VR/Reservoir.slang:8-87
inline Reservoir createNewReservoir() { return {0.f, 0.f, FLT_MAX, 0.f, {0, 0}, 0, 0, 0, 0.f}; }
inline void takeSample(const Reservoir& r, Reservoir& state, bool sel, int maxBounces) {
    state.depth = sel ? r.depth : state.depth;
    state.p_y = sel ? r.p_y : state.p_y;
    state.lightUV = sel ? r.lightUV : state.lightUV;
    state.lightID = sel ? r.lightID : state.lightID;
    if (maxBounces > 1) state.extraBounceStartId = sel ? r.extraBounceStartId : state.extraBounceStartId;
    if (maxBounces > 1) state.p_partial = sel ? r.p_partial : state.p_partial;   // VERTEX_REUSE (Reservoir.slang:46-47,76-77); stays 0 without it
    state.sampledPixel = sel ? r.sampledPixel : state.sampledPixel;
}
inline bool simpleResampleStep(const Reservoir& reservoir, Reservoir& state, SampleGenerator& sg, int maxBounces) {
    float sampleWeight = reservoir.runningSum;
    state.M += reservoir.M;
    if (sampleWeight <= 0.0f) return false;
    state.runningSum += sampleWeight;
    bool selectSample = sampleNext1D(sg) * state.runningSum < sampleWeight;
    takeSample(reservoir, state, selectSample, maxBounces);
    return selectSample;
}
inline bool simpleResampleStepWithMaxM(const Reservoir& reservoir, float MThreshold, Reservoir& state, SampleGenerator& sg, int maxBounces) {
    float correctedM = fminf(MThreshold, reservoir.M);
    float sampleWeight = correctedM == 0.0f ? 0.0f : correctedM / reservoir.M * reservoir.runningSum;
    state.M += correctedM;
    if (sampleWeight <= 0.0f) return false;
    state.runningSum += sampleWeight;
    bool selectSample = sampleNext1D(sg) * state.runningSum < sampleWeight;
    takeSample(reservoir, state, selectSample, maxBounces);
    return selectSample;
}

// VR/ReSTIRHelper.slang:435-496
inline float3 evaluate_L_in_volume(const Ctx& c, const MediumInteraction& mi, int lightID, float2 lightUV, float& precomputedVisibility, SampleGenerator& sg,
                                   const SamplingOptions& options, bool lightVisibilityReuse, bool isLastFrame, bool cullNonOpaqueGeometry) {
    const Pass& P = c.p;
    Ray shadowRay{}; float3 Ld = f3(0.f); bool isValidSample = true;
    bool useLastFrameGrid = c.vd().usePrevGridForReproj && isLastFrame && c.vd().hasAnimation;
    int densityGridOffset = useLastFrameGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
    if (lightID < 0) {
        float3 wiWorld = envDecodeLightUV(P, lightUV, lightID, isLastFrame);
        shadowRay = {mi.p, wiWorld, 0, kRayTMax};
        Ld = envEval(P, wiWorld, isLastFrame) * mi.phaseFunction(mi.wo, wiWorld);
    } else if (lightID < (int)P.lights.size()) {
        SceneLightSample ls = getAnalyticalLightSample(c, lightID, mi.p, sg);
        shadowRay = {mi.p, ls.rayDir, 0, ls.rayDistance};
        Ld = ls.Li * mi.phaseFunction(mi.wo, ls.rayDir);
    } else {
        TriangleLightSample ls{};
        isValidSample = getEmissiveLightSample(P, mi.p, lightID - (int)P.lights.size(), lightUV, ls);
        if (isValidSample) {
            shadowRay = {mi.p, ls.dir, 0, ls.distance};
            Ld = ls.Le * P.emissiveMul * mi.phaseFunction(mi.wo, ls.dir) * ls.cosTheta / (ls.distance * ls.distance);
        }
    }
    float Tr = 1.f;
    if (isValidSample) {
        if (lightVisibilityReuse) Tr = precomputedVisibility;   // VERTEX_REUSE: the shadow-ray transmittance stored by the sample's own pixel
        else {
            Tr = computeVisibility(c, shadowRay, sg, options.lightSamples, cullNonOpaqueGeometry ? options.lightingMipLevel + densityGridOffset : 0,
                                   options.lightingUseLinearSampler, options.lightingTrackingMethod, options.lightingTStepScale);
            precomputedVisibility = Tr;
        }
    }
    return Tr * Ld;
}

// extra-bounce data provider (VR/ArrayDataProvider.slang)
struct ExtraProvider { const ExtraBounce* data; ExtraBounce get(int i) const { return data[i]; } };

// VR/ReSTIRHelper.slang:91-423 (MAX_BOUNCES and VERTEX_REUSE runtime, no SURFACE_SCENE).  `tap` is REUSETYPE = inout under
// VERTEX_REUSE: the evaluation leaves the suffix of F past the reuse vertex in tap.p_partial unless spatialReuse reads it.
inline float3 evaluate_F_(const Ctx& c, Reservoir& tap, const ExtraProvider& extra, Ray ray, SampleGenerator& sg, const SamplingOptions& options,
                          bool isLastFrame, bool noReuse, bool spatialReuse, bool isFinalShading) {
    const Pass& P = c.p; const auto& vd = c.vd();
    const int maxBounces = P.P.mMaxBounces;
    const bool vertexReuse = P.P.mVertexReuse && maxBounces > 1;
    const int S = options.vertexReuseStartBounce;
    bool useLastFrameGrid = vd.usePrevGridForReproj && isLastFrame && vd.hasAnimation;
    int mipLevelOffset = useLastFrameGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
    bool isBackgroundSample = tap.depth == kRayTMax;
    ray.tMax = tap.depth;
    float3 F = f3(1.f);
    int maxIndirectBounces = 0; bool isSelfEmission;
    if (maxBounces > 1) { maxIndirectBounces = decodeMaxIndirectBounces(tap.sampledPixel, maxBounces); isSelfEmission = maxIndirectBounces == 0 && tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID; }
    else isSelfEmission = tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID;
    float visibility = 1.f;
    float3 p_World = ray.at(ray.tMax);
    MediumInteraction mi = {ray.at(ray.tMax), -ray.dir, vd.PhaseFunctionConstantG, true};
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    {
        float3 sigma_s = f3(1.f);
        float density = (isBackgroundSample || noReuse) ? 1.f : DensityWorldSpace(c, p_World, mipLevelOffset);
        if (density == 0.f) return f3(0.f);
        if (!noReuse)
            visibility = computeVisibility(c, ray, sg, options.visibilitySamples, options.visibilityMipLevel + mipLevelOffset, options.visibilityUseLinearSampler,
                                           options.visibilityTrackingMethod, options.visibilityTStepScale);
        sigma_s = isBackgroundSample ? f3(1.f) : (isSelfEmission ? sigA : sigS);
        if (noReuse && !isBackgroundSample) sigma_s = sigma_s / vd.sigma_t;
        F *= visibility * density * sigma_s;
    }
    float3 P_prefix = f3(1.f);
    int bounceId = 0;
    if (any_gt0(F)) {
        if (isBackgroundSample) {
            F *= envEval(P, ray.dir, isLastFrame);
        } else if (isSelfEmission) {
            F *= EmissionWorldSpace(c, p_World, useLastFrameGrid);
        } else {
            bool isScatterSelfEmission = false;
            if (maxBounces > 1 && maxIndirectBounces > 0) {
                Ray scatterRay{};
                isScatterSelfEmission = decodePathTag(tap.sampledPixel) == 1;
                int numIndirectBounces = maxIndirectBounces;
                for (; bounceId < numIndirectBounces; bounceId++) {
                    bool isCurrentVertexEmissive = isScatterSelfEmission && bounceId == numIndirectBounces - 1;
                    float4_ wiDist = decodeWiDist(extra.get(tap.extraBounceStartId + bounceId).wi_dist, vertexReuse && bounceId + 1 >= S);
                    if (isCurrentVertexEmissive && !(vertexReuse && bounceId + 1 >= S)) {
                        float3 e = decodeEmissivePosition(tap.lightID, tap.lightUV);
                        wiDist = {e.x, e.y, e.z, -1.f};
                    }
                    if (wiDist.w == kRayTMax) return f3(0.f);
                    float dist = 1.f;
                    if (wiDist.w == -1.f) {
                        float3 disp = f3(wiDist.x, wiDist.y, wiDist.z) - p_World;
                        dist = length(disp);
                        float3 dir = normalize(disp);
                        scatterRay = {p_World, dir, 0.f, dist};
                    } else {
                        scatterRay = {p_World, f3(wiDist.x, wiDist.y, wiDist.z), 0.f, wiDist.w};
                    }
                    float bsdf = mi.phaseFunction(mi.wo, scatterRay.dir);
                    F *= bsdf;
                    if (all_eq0(F)) return f3(0.f);
                    if (vertexReuse && bounceId == S) {   // :289-296 (the bounce past the reuse vertex)
                        if (spatialReuse) return F * tap.p_partial;
                        P_prefix = F;
                    }
                    if (wiDist.w == -1.f) p_World = f3(wiDist.x, wiDist.y, wiDist.z);
                    else p_World = scatterRay.at(scatterRay.tMax);
                    float3 sigma_s; float scatterDensity;
                    if (!noReuse) {
                        sigma_s = isCurrentVertexEmissive ? sigA : sigS;
                        scatterDensity = fmaxf(0.f, DensityWorldSpace(c, p_World, mipLevelOffset));
                    } else {
                        sigma_s = isCurrentVertexEmissive ? sigA / vd.sigma_t : sigS / vd.sigma_t;
                        scatterDensity = 1.f;
                    }
                    F *= scatterDensity * sigma_s;
                    if (vertexReuse ? (bounceId + 1 == S || (isCurrentVertexEmissive && bounceId + 1 < S)) : isCurrentVertexEmissive) F *= 1.f / (dist * dist);
                    if (all_eq0(F)) return f3(0.f);
                    float scatterVisibility = 1.f;
                    if (!noReuse)
                        scatterVisibility = computeVisibility(c, scatterRay, sg, options.visibilitySamples, options.visibilityMipLevel + mipLevelOffset,
                                                              options.visibilityUseLinearSampler, options.visibilityTrackingMethod, options.visibilityTStepScale);
                    F *= scatterVisibility;
                    mi.wo = -scatterRay.dir;
                    mi.p = p_World;
                    if (all_eq0(F)) return f3(0.f);
                }
                if (isScatterSelfEmission) F *= EmissionWorldSpace(c, p_World, useLastFrameGrid);
                else mi = {scatterRay.at(scatterRay.tMax), -scatterRay.dir, vd.PhaseFunctionConstantG, true};
            }
            if (!isScatterSelfEmission && any_gt0(F)) {
                float precomputedVisibility = vertexReuse ? tap.p_partial : 1.f;
                const bool lightVisibilityReuse = vertexReuse && bounceId == S && spatialReuse;
                F *= evaluate_L_in_volume(c, mi, tap.lightID, tap.lightUV, precomputedVisibility, sg, options, lightVisibilityReuse, isLastFrame, !isFinalShading);
                if (vertexReuse && bounceId == S && !spatialReuse) tap.p_partial = precomputedVisibility;
            }
        }
    }
    // :415-420 — componentwise F / P_prefix as written there (a zero prefix component gives NaN like the shader)
    if (vertexReuse && bounceId > S && !spatialReuse) tap.p_partial = luminance(F / P_prefix);
    return F;
}
// VR/ReSTIRHelper.slang:426-441
inline float evaluate_P_hat(const Ctx& c, const Ray& ray, SampleGenerator& sg, const ExtraProvider& extra, const SamplingOptions& options, Reservoir& tap,
                            bool isLastFrame = false, bool nonBinary = true, bool spatialReuse = false) {
    float3 F = evaluate_F_(c, tap, extra, ray, sg, options, isLastFrame, false, spatialReuse, false);
    return (!nonBinary && any_gt0(F)) ? 1.f : luminance(F);
}
// VR/ReSTIRHelper.slang:600-607 (evaluatePHatReadOnly: the reservoir is passed by value, p_partial of the caller's copy stays)
inline float evaluatePHatReadOnly(const Ctx& c, const Ray& ray, SampleGenerator& sg, const ExtraProvider& extra, const SamplingOptions& options, Reservoir tap,
                                  bool isLastFrame, bool nonBinary, bool spatialReuse) {
    return evaluate_P_hat(c, ray, sg, extra, options, tap, isLastFrame, nonBinary, spatialReuse);
}
inline float3 evaluate_F(const Ctx& c, Reservoir tap, const ExtraProvider& extra, const Ray& ray, SampleGenerator& sg, const SamplingOptions& options, bool noReuse) {
    return evaluate_F_(c, tap, extra, ray, sg, options, false, noReuse, false, true);
}
// VR/ReSTIRHelper.slang:560-597 (resampleNeighbor / resampleNeighborSpatialReuse differ only in the spatialReuse flag; under VERTEX_REUSE
// the temporal variant overwrites tap.p_partial, the spatial one reads it)
inline bool resampleNeighbor(const Ctx& c, Reservoir& tap, const Ray& ray, SampleGenerator& sg, const ExtraProvider& extra, const SamplingOptions& options, bool spatial) {
    if (tap.runningSum == 0.f) return true;
    float p_y_hat = evaluate_P_hat(c, ray, sg, extra, options, tap, false, true, spatial);
    float weight = p_y_hat / tap.p_y;
    if (std::isinf(weight) || std::isnan(weight)) weight = 0.f;
    tap.runningSum *= weight;
    tap.p_y = p_y_hat;
    return true;
}

// ------------------------------------------------------------------------------------------------ ComputeInitialSample
// VR/ComputeInitialSample.slang:4-395
inline Reservoir ComputeInitialSample(const Ctx& c, const Ray& primaryRay, float precomputedHitDistance, float precomputedPdfDist, float precomputedTr, int maxBounces,
                                      SampleGenerator& sg, const SamplingOptions& options, bool useCoarserGridForIndirectBounce, bool useRussianRoulette, bool noReuse,
                                      ExtraBounce* extrabounceReservoir) {
    const Pass& P = c.p; const auto& vd = c.vd();
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    const bool vertexReuse = P.P.mVertexReuse && maxBounces > 1;
    float pathPdf = 1.f, pathPHat = 1.f;
    Ray ray = primaryRay;
    Reservoir combinedReservoir = createNewReservoir();
    Reservoir outReservoir = createNewReservoir();
    int bounce = 0;
    float primaryScatterDepth = 0;
    for (; bounce < maxBounces; bounce++) {
        outReservoir = createNewReservoir();
        outReservoir.M = 1;
        float curHitDist; float pdfDist = 0;
        MediumInteraction mi{};
        float Tr;
        if (bounce >= 1 || noReuse) {
            if (noReuse) {
                SampleMediumSuperVoxelGeneric(c, ray, sg, mi, 0);
                pdfDist = 1.f; Tr = 1.f;
                curHitDist = mi.isValid ? length(mi.p - ray.origin) : kRayTMax;
            } else {
                int curMip = options.visibilityMipLevel;
                if (useCoarserGridForIndirectBounce)
                    curMip = std::min((options.visibilityMipLevel >= VRESTIR_NUM_MAX_MIPS ? VRESTIR_NUM_MAX_MIPS : 0) + vd.numMips - 1, curMip + 1);
                float hd[4], pd[4], ot[4];
                SampleMediumAnalyticGeneric(c, ray, sg, options.visibilityUseLinearSampler, hd, curMip, pd, ot, 1);
                curHitDist = hd[0]; pdfDist = pd[0]; Tr = ot[0];
                mi = {ray.at(curHitDist), -ray.dir, vd.PhaseFunctionConstantG, curHitDist != kRayTMax};
            }
        } else {
            curHitDist = precomputedHitDistance; pdfDist = precomputedPdfDist; Tr = precomputedTr;
            mi = {ray.at(curHitDist), -ray.dir, vd.PhaseFunctionConstantG, curHitDist != kRayTMax};
        }
        pathPdf *= pdfDist;
        if (vertexReuse && bounce == options.vertexReuseStartBounce && curHitDist != kRayTMax) {   // :88-94 area measure at the reuse vertex
            pathPdf /= curHitDist * curHitDist;
            pathPHat /= curHitDist * curHitDist;
        }
        bool hitEmpty = false;
        float actualVolumeDensity = 0.f;
        if (bounce == 0) {
            if (maxBounces > 1) outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, 0);
            outReservoir.depth = mi.isValid ? curHitDist : kRayTMax;
            primaryScatterDepth = outReservoir.depth;
            outReservoir.p_y = pathPdf;
        } else {
            outReservoir.depth = primaryScatterDepth;
            if (maxBounces > 1) {
                outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, bounce);
                if (!vertexReuse || bounce < options.vertexReuseStartBounce)
                    extrabounceReservoir[bounce - 1].wi_dist = encodeWiDist({ray.dir.x, ray.dir.y, ray.dir.z, !mi.isValid ? kRayTMax : curHitDist});
                else   // :116-125 world-space vertex
                    extrabounceReservoir[bounce - 1].wi_dist = !mi.isValid ? f3(kRayTMax) : mi.p;
            }
            outReservoir.p_y = pathPdf;
        }
        if (mi.isValid) {
            if (noReuse) actualVolumeDensity = 1.f;
            else actualVolumeDensity = DensityWorldSpace(c, mi.p, 0);
        }
        if ((!mi.isValid && bounce > 0) || (mi.isValid && actualVolumeDensity == 0)) {
            outReservoir.p_y = 0.f; outReservoir.runningSum = 0.f; hitEmpty = true;
        }
        float pdfDir = 1.f;
        if (!hitEmpty) {
            if (mi.isValid) {
                float3 albedo = sigS / vd.sigma_t;
                outReservoir.lightID = -1;
                outReservoir.lightUV = {0, 0};
                float outLightPdf = 0.f, outVisibility = 1.f;
                float3 Ld = f3(0.f), Le = f3(0.f);
                float3 one_minus_albedo = f3(1.f) - albedo;
                if (vd.hasEmission && (actualVolumeDensity > 0.f)) Le = EmissionWorldSpace(c, mi.p);
                bool shouldComputeLightVisibility = options.lightSamples == 0 ? false : true;
                Ld = SampleDirectLighting(c, sg, outLightPdf, mi, options.useEnvironmentLights, options.useAnalyticLights, options.useEmissiveLights, shouldComputeLightVisibility,
                                          options.lightingMipLevel, options.lightingTrackingMethod, options.lightingUseLinearSampler, options.lightSamples, options.lightingTStepScale,
                                          outReservoir.lightID, outReservoir.lightUV, outVisibility);
                float3 wo = -ray.dir, wi = f3(0.f);
                if (maxBounces > 1) pdfDir = mi.Sample_p(wo, wi, sampleNext2D(sg));
                float p_src = outReservoir.p_y;
                {
                    float lumE = luminance(one_minus_albedo * Le);
                    float emissionRatio = lumE / (lumE + luminance(albedo * Ld));
                    if (std::isnan(emissionRatio)) emissionRatio = 0.f;
                    if (sampleNext1D(sg) < emissionRatio) { p_src *= emissionRatio; outReservoir.lightID = VRESTIR_SELF_EMISSION_LIGHT_ID; }
                    else p_src *= outLightPdf * (1 - emissionRatio);
                }
                outReservoir.runningSum = p_src == 0.f ? 0.f : 1.f;
                outReservoir.p_y = p_src;
                {
                    float p_y;
                    pathPHat *= Tr;
                    pathPHat *= actualVolumeDensity;
                    if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID) p_y = pathPHat * luminance(sigA * Le);
                    else p_y = pathPHat * luminance(sigS * Ld * outLightPdf);
                    pathPHat *= luminance(sigS) * pdfDir;
                    if (noReuse) { p_y /= vd.sigma_t; pathPHat /= vd.sigma_t; }
                    if (outReservoir.runningSum > 0.f) {
                        outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                        if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID && bounce > 0) {
                            if (!vertexReuse || bounce < options.vertexReuseStartBounce) {   // :324-331
                                encodeEmissivePosition(mi.p, outReservoir.lightID, outReservoir.lightUV);
                                p_y /= (curHitDist * curHitDist);
                            }
                            outReservoir.sampledPixel = encodePathTag(outReservoir.sampledPixel, 1);
                        }
                        outReservoir.p_y = p_y;
                    }
                }
                pathPdf *= pdfDir;
                if (maxBounces > 1 && bounce < maxBounces - 1) {
                    ray = mi.SpawnRay(wi);
                    if (useRussianRoulette && bounce >= 2) {
                        if (sampleNext1D(sg) < albedo.x) pathPdf *= albedo.x;
                        else { hitEmpty = true; combinedReservoir.M++; }
                    }
                }
            } else {
                float3 Le = envEval(P, ray.dir);
                pathPHat *= Tr;
                float p_y = pathPHat * luminance(Le);
                outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                outReservoir.p_y = p_y;
                hitEmpty = true;
            }
        }
        if (maxBounces > 1) simpleResampleStep(outReservoir, combinedReservoir, sg, maxBounces);
        if (hitEmpty) break;
    }
    if (maxBounces > 1) { combinedReservoir.M = 1; return combinedReservoir; }
    return outReservoir;
}

// VR/VolumePathTracingFunctions.slang:3-131 (mUseReference)
inline float3 IntegrateByVolumePathTracing(const Ctx& c, Ray ray, SampleGenerator& sg, int lightSamples, bool kEnv, bool kAnalytic, bool kEmissive, int mip, int maxBounces,
                                           bool useNEE, bool useRussianRoulette) {
    const Pass& P = c.p; const auto& vd = c.vd();
    MediumInteraction mi{};
    float3 beta = f3(1.f), L = f3(0.f);
    if (!useNEE) maxBounces += 1;
    for (int bounce = 0; bounce < maxBounces; bounce++) {
        mi.isValid = false;
        SampleMediumSuperVoxelGeneric(c, ray, sg, mi, mip);
        if (mi.isValid) {
            float3 albedo = v3(vd.sigma_s) / vd.sigma_t;
            float3 one_minus_albedo = f3(1.f) - albedo;
            { float3 Le = EmissionWorldSpace(c, mi.p); L += Le * one_minus_albedo * beta; }
            beta *= albedo;
            if (useNEE) {
                float3 Ld = directLighting(c, sg, mi, lightSamples, kEnv, kAnalytic, kEmissive, true, mip, VRESTIR_RESIDUAL_RATIO_TRACKING);
                L += beta * Ld;
            }
            float3 wo = -ray.dir, wi = f3(0.f);
            if (maxBounces > 1) mi.Sample_p(wo, wi, sampleNext2D(sg));
            if (bounce < maxBounces - 1) {
                ray = mi.SpawnRay(wi);
                if (useRussianRoulette && bounce >= 2) {
                    if (sampleNext1D(sg) < albedo.x) beta = beta / albedo.x;
                    else bounce = maxBounces;
                }
            }
        } else {
            float3 Le = envEval(P, ray.dir);
            if (!useNEE || bounce == 0) L += beta * Le;
            bounce = maxBounces;
        }
    }
    return L;
}

// ------------------------------------------------------------------------------------------------ host sequencing helpers
// VR/VolumetricReSTIR.cpp:452-496
struct FrameSetup { int numTotalRounds; SamplingOptions initial, spatial, fin; };
inline FrameSetup buildFrameSetup(const vrestir_params& m) {
    FrameSetup f;
    const int numInitialSamplingRounds = 1;
    f.numTotalRounds = (m.mEnableSpatialReuse ? m.mSpatialReuseRounds : 0) + (m.mEnableTemporalReuse ? 1 : 0) + 1 + numInitialSamplingRounds;
    f.initial = {VRESTIR_ANALYTIC_TRACKING, m.mInitialLightingTrackingMethod, m.mInitialLightSamples, m.mInitialLightingMipLevel, 1,
                 m.mInitialVisibilityUseLinearSampler ? m.mInitialBaseMipLevel : m.mInitialBaseMipLevel + VRESTIR_NUM_MAX_MIPS,
                 (bool)m.mInitialVisibilityUseLinearSampler, (bool)m.mInitialLightingUseLinearSampler, m.mInitialVisibilityTStepScale, m.mInitialLightingTStepScale,
                 (bool)m.mUseEnvironmentLights, (bool)m.mUseAnalyticLights, (bool)m.mUseEmissiveLights, m.mVertexReuseStartBounce};
    f.spatial = {m.mSpatialVisibilityTrackingMethod, m.mSpatialLightingTrackingMethod, 1, m.mSpatialLightingMipLevel, 1, m.mSpatialVisibilityMipLevel,
                 (bool)m.mSpatialVisibilityUseLinearSampler, (bool)m.mSpatialLightingUseLinearSampler, m.mSpatialVisibilityTStepScale, m.mSpatialLightingTStepScale,
                 (bool)m.mUseEnvironmentLights, (bool)m.mUseAnalyticLights, (bool)m.mUseEmissiveLights, m.mVertexReuseStartBounce};
    bool lightDet = m.mFinalLightTrackingMethod == VRESTIR_ANALYTIC_TRACKING || m.mFinalLightTrackingMethod == VRESTIR_RAY_MARCHING;
    bool visDet = m.mFinalVisibilityTrackingMethod == VRESTIR_ANALYTIC_TRACKING || m.mFinalVisibilityTrackingMethod == VRESTIR_RAY_MARCHING;
    f.fin = {m.mFinalVisibilityTrackingMethod, m.mFinalLightTrackingMethod, lightDet ? 1 : m.mFinalLightSamples, 0, visDet ? 1 : m.mFinalVisibilitySamples, 0, true, true,
             m.mFinalTStepScale, m.mFinalTStepScale, (bool)m.mUseEnvironmentLights, (bool)m.mUseAnalyticLights, (bool)m.mUseEmissiveLights, m.mVertexReuseStartBounce};
    return f;
}

// F/Utils/Math/MathHelpers.slang:178-197 sample_disk, :356-361 getHammersley; VR/SpatialReuse.cs.slang:64-81
inline float radicalInverse(uint32_t i) {
    i = (i & 0x55555555u) << 1 | (i & 0xAAAAAAAAu) >> 1; i = (i & 0x33333333u) << 2 | (i & 0xCCCCCCCCu) >> 2;
    i = (i & 0x0F0F0F0Fu) << 4 | (i & 0xF0F0F0F0u) >> 4; i = (i & 0x00FF00FFu) << 8 | (i & 0xFF00FF00u) >> 8;
    i = (i << 16) | (i >> 16);
    return (float)i * 2.3283064365386963e-10f;
}
inline int2 generateNeighborOffset(const vrestir_params& m, int sampleId, int sampleCount, float sampleRadius, int frameId) {
    float2 u;
    if (m.mRandomSamplerType == VRESTIR_SAMPLER_HAMMERSLEY) u = {(float)sampleId / (float)sampleCount, radicalInverse((uint32_t)sampleId)};
    else {
        double multiplier = (double)(frameId * m.mSpatialSampleCount + sampleId);
        if (sampleId == 0) u = {0, 0};
        else { double a = 0.754877669 * multiplier, b = 0.569840296 * multiplier; u = {(float)(a - floor(a)), (float)(b - floor(b))}; }
    }
    float r = sqrtf(u.x), phi = M_2PI_F * u.y;
    float2 d = {r * cosf(phi), r * sinf(phi)};
    return {f2i(sampleRadius * d.x), f2i(sampleRadius * d.y)};
}

inline bool IsWithinRange(int x, int y, int W, int H) { return x >= 0 && x < W && y >= 0 && y < H; }
inline int wrapMulAdd(int y, int W, int x) { return (int)((uint32_t)y * (uint32_t)W + (uint32_t)x); }

void ensureBuffers(Pass& p) {
    int B = p.P.mMaxBounces;
    if (p.allocW == p.W && p.allocH == p.H && p.allocB == B) return;
    size_t n = (size_t)p.W * p.H;
    Reservoir z; std::memset(&z, 0, sizeof(z));   // D3D buffers start zero-filled
    for (int i = 0; i < 2; i++) { p.res[i].assign(n, z); p.ext[i].assign(n * (size_t)std::max(0, B - 1), ExtraBounce{{0, 0, 0}}); }
    p.resT.assign(n, z); p.extT.assign(n * (size_t)std::max(0, B - 1), ExtraBounce{{0, 0, 0}});
    p.feat.assign(n, Features{0, 0.f}); p.featT.assign(n, Features{0, 0.f});
    p.refColor.assign(n * 4, 0.f);
    p.allocW = p.W; p.allocH = p.H; p.allocB = B;
}

// VR/VolumetricReSTIR.cpp:211-235 overrideVolumeDesc
void applyOverrides(Pass& p) {
    p.vol = p.volBase;
    if (p.volumeDensityScaleExtraControl > 0) p.vol.densityScaleFactor = p.volumeDensityScaleExtraControl;
    if (p.volumeAnisotropyExtraControl > 0) p.vol.PhaseFunctionConstantG = p.volumeAnisotropyExtraControl;
    if (p.volumeAlbedoExtraControl > 0) {
        for (int i = 0; i < 3; i++) { p.vol.sigma_s[i] = p.vol.sigma_t * p.volumeAlbedoExtraControl; p.vol.sigma_a[i] = p.vol.sigma_t - p.vol.sigma_s[i]; }
    }
    p.vol.usePrevGridForReproj = p.P.mUsePrevVolumeForReproj;
}

double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- K0 GenerateFeatures: VR/GenerateFeatures.cs.slang:57-102 ----
void stageFeatures(Pass& p, const FrameSetup& fs) {
    if (p.P.mUseReference) return;
    Ctx c(p);
    const float3 U = v3(p.cam.cameraU), V = v3(p.cam.cameraV), Wv = v3(p.cam.cameraW), pos = v3(p.cam.posW);
    forPixels(p, [&](int x, int y) {
        SampleGenerator sg{};  // unused
        Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, x, y, p.W, p.H)), 0, kRayTMax};
        float accu;
        ReservoirFeatureRayMarchingGeneric(c, ray, sg, 0, true, fs.initial.visibilityTStepScale, accu);
        p.feat[(size_t)y * p.W + x] = Features{1, accu};
    });
}

// ---- K1 TraceRays: VR/TraceRays.cs.slang:64-201 ----
void stageInitial(Pass& p, const FrameSetup& fs) {
    Ctx c(p);
    const vrestir_params& m = p.P;
    const int B = m.mMaxBounces;
    const float3 U = v3(p.cam.cameraU), V = v3(p.cam.cameraV), Wv = v3(p.cam.cameraW), pos = v3(p.cam.posW);
    const bool noReuse = !m.mEnableSpatialReuse && !m.mEnableTemporalReuse;
    forPixels(p, [&](int x, int y) {
        SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fs.numTotalRounds * p.mFrameCount));
        int reservoirId = y * p.W + x;
        Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, x, y, p.W, p.H)), 0.f, kRayTMax};
        if (m.mUseReference) {
            float3 avgL = f3(0.f);
            for (int r = 0; r < m.mBaselineSamplePerPixel; r++)
                avgL += IntegrateByVolumePathTracing(c, ray, sg, std::max(1, fs.initial.lightSamples), fs.initial.useEnvironmentLights, fs.initial.useAnalyticLights,
                                                     fs.initial.useEmissiveLights, 0, B, true, m.mInitialUseRussianRoulette);
            float3 o = avgL / (float)m.mBaselineSamplePerPixel;
            float* d = &p.refColor[(size_t)reservoirId * 4]; d[0] = o.x; d[1] = o.y; d[2] = o.z; d[3] = 1.f;
            return;
        }
        ExtraBounce finalExtra[3] = {};
        Reservoir finalReservoir = createNewReservoir();
        int rounds = (m.mInitialM + 3) / 4;
        for (int roundId = 0; roundId < rounds; roundId++) {
            float hd[4] = {0, 0, 0, 0}, pd[4] = {0, 0, 0, 0}, ot[4] = {0, 0, 0, 0};
            int roundSamples = roundId == rounds - 1 ? (m.mInitialM - 4 * (rounds - 1)) : 4;
            if (!noReuse) SampleMediumAnalyticGeneric(c, ray, sg, fs.initial.visibilityUseLinearSampler, hd, fs.initial.visibilityMipLevel, pd, ot, roundSamples);
            for (int s = 0; s < roundSamples; s++) {
                ExtraBounce extra[3] = {};
                Reservoir outReservoir = ComputeInitialSample(c, ray, hd[s], pd[s], ot[s], B, sg, fs.initial, m.mInitialUseCoarserGridForIndirectBounce,
                                                              m.mInitialUseRussianRoulette, noReuse, extra);
                bool isSelected = simpleResampleStep(outReservoir, finalReservoir, sg, B);
                if (isSelected && B > 1) {
                    int mib = decodeMaxIndirectBounces(finalReservoir.sampledPixel, B);
                    for (int b = 0; b < mib; b++) finalExtra[b] = extra[b];
                }
            }
        }
        ExtraProvider prov{finalExtra};
        Reservoir tapForEval = finalReservoir; tapForEval.extraBounceStartId = 0;
        uint32_t* gen = p.k1Generator.size() == (size_t)p.W * p.H * 8 ? &p.k1Generator[(size_t)reservoirId * 8] : nullptr;
        if (gen) for (int i = 0; i < 4; i++) gen[i] = sg.s[i];
        float p_hat = evaluate_P_hat(c, ray, sg, prov, fs.spatial, tapForEval, false, true, false);
        if (gen) for (int i = 0; i < 4; i++) gen[4 + i] = sg.s[i];
        finalReservoir.p_partial = tapForEval.p_partial;   // TraceRays.cs.slang:176-177 passes finalReservoir itself (inout under VERTEX_REUSE)
        if (finalReservoir.runningSum > 0.f) {
            finalReservoir.runningSum *= finalReservoir.p_y == 0.f ? 0.f : p_hat / finalReservoir.p_y;
            finalReservoir.p_y = p_hat;
        }
        finalReservoir.extraBounceStartId = B > 1 ? reservoirId * (B - 1) : 0;
        p.res[0][reservoirId] = finalReservoir;
        if (B > 1) {
            int mib = decodeMaxIndirectBounces(finalReservoir.sampledPixel, B);
            for (int b = 0; b < mib; b++) p.ext[0][(size_t)finalReservoir.extraBounceStartId + b] = finalExtra[b];
        }
    });
}

// ---- K2 TemporalReuse: VR/TemporalReuse.cs.slang:80-377 ----
void stageTemporal(Pass& p, const FrameSetup& fs, float* out_mvec) {
    Ctx c(p);
    const vrestir_params& m = p.P;
    const int B = m.mMaxBounces, W = p.W, H = p.H;
    const float3 U = v3(p.cam.cameraU), V = v3(p.cam.cameraV), Wv = v3(p.cam.cameraW), pos = v3(p.cam.posW);
    const SamplingOptions& opt = fs.spatial;
    std::vector<Reservoir>& cur = p.res[0];
    std::vector<ExtraBounce>& curExt = p.ext[0];
    const bool isFirstFrame = p.mTemporalSampleAccumulated == 0;
    if (isFirstFrame) return;
    const uint32_t mis = m.mTemporalMISMethod;
    // K2 writes gCurExtraBounceReservoirs in place while other pixels only read the *temporal* extra buffer: safe.
    forPixels(p, [&](int x, int y) {
        SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fs.numTotalRounds * p.mFrameCount + 1));
        int selectedId = -1;
        int pixelId = y * W + x;
        int numUsedReservoirs = 1;
        Reservoir taps[2];
        taps[0] = cur[pixelId];
        taps[1] = createNewReservoir();
        Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, x, y, W, H)), 0, kRayTMax};
        Reservoir output = mis == VRESTIR_MIS_TALBOT ? createNewReservoir() : taps[0];
        int centerExtraBounceStartId = taps[0].extraBounceStartId;
        float temporalOriginalDepth = 0.f;
        int2 reprojScreenPos = {0, 0};
        Features centerFeatures = p.feat[pixelId];
        bool isBackgroundReservoir = centerFeatures.transmittance == 1.f && centerFeatures.noReflectiveSurface;
        bool useFallbackReservoir = true;
        auto writeMvec = [&]() {
            if (p.mOutputMotionVec && out_mvec) {
                out_mvec[(size_t)pixelId * 2 + 0] = (float)(reprojScreenPos.x - x) / (float)W;
                out_mvec[(size_t)pixelId * 2 + 1] = (float)(reprojScreenPos.y - y) / (float)H;
            }
        };
        if (m.mTemporalReprojectionMode != VRESTIR_REPROJECTION_NONE) {
            float reprojDepth = taps[0].depth;
            if (reprojDepth == kRayTMax && m.mTemporalReprojectionMode != VRESTIR_REPROJECTION_NO_BACKGROUND && !isBackgroundReservoir)
                reprojDepth = RejectionSampleRandomPointByDensity(c, ray, sg, VRESTIR_NUM_MAX_MIPS + m.mTemporalReprojectionMipLevel);
            float3 pw = ray.origin + ray.dir * reprojDepth;
            if (c.vd().hasVelocity && c.vd().hasAnimation) {
                float3 v = VelocityWorld(c, pw) * c.vd().velocityScale;
                pw = pw - v;
            }
            // float4 viewPos = mul(float4(p,1), gPrevViewMat); float4 clipPos = mul(viewPos, gPrevProjMat);
            const float* Vm = p.prevView; const float* Pm = p.prevProj;
            float vp[4], cp[4];
            for (int j = 0; j < 4; j++) vp[j] = pw.x * Vm[0 + j] + pw.y * Vm[4 + j] + pw.z * Vm[8 + j] + 1.f * Vm[12 + j];
            for (int j = 0; j < 4; j++) cp[j] = vp[0] * Pm[0 + j] + vp[1] * Pm[4 + j] + vp[2] * Pm[8 + j] + vp[3] * Pm[12 + j];
            float2 scrPos = {cp[0] / cp[3], cp[1] / cp[3]};
            int2 scrPosI;
            if (reprojDepth == kRayTMax) { scrPos = {(float)x + 0.5f, (float)y + 0.5f}; scrPosI = {x, y}; }
            else {
                scrPos.x = 0.5f * scrPos.x + 0.5f; scrPos.y = -0.5f * scrPos.y + 0.5f;
                scrPos.x *= (float)W; scrPos.y *= (float)H;
                scrPosI = {f2i(scrPos.x), f2i(scrPos.y)};
            }
            {
                int id = wrapMulAdd(scrPosI.y, W, scrPosI.x);
                Features tapFeatures = (id >= 0 && id < W * H) ? p.featT[id] : Features{0, 0.f};
                bool isTapBackgroundReservoir = tapFeatures.transmittance == 1.f && tapFeatures.noReflectiveSurface;
                if (isBackgroundReservoir && !isTapBackgroundReservoir) { writeMvec(); return; }
            }
            {
                scrPosI = {f2i(scrPos.x), f2i(scrPos.y)};
                reprojScreenPos = scrPosI;
                if (IsWithinRange(scrPosI.x, scrPosI.y, W, H)) { numUsedReservoirs++; taps[1] = p.resT[scrPosI.y * W + scrPosI.x]; }
            }
            if (numUsedReservoirs > 1) useFallbackReservoir = false;
        }
        if (useFallbackReservoir) { numUsedReservoirs++; reprojScreenPos = {x, y}; taps[1] = p.resT[pixelId]; }
        writeMvec();
        float curM = taps[0].M;
        float MaxPrevM = m.mTemporalReuseMThreshold * curM;
        if (numUsedReservoirs == 2) {
            temporalOriginalDepth = taps[1].depth;
            if (taps[1].depth != kRayTMax) {
                float3 dir = normalize(camRayDirNN(p.prevU, p.prevV, p.prevW, reprojScreenPos.x, reprojScreenPos.y, W, H));
                float3 worldPos = p.prevPos + taps[1].depth * dir;
                taps[1].depth = length(worldPos - ray.origin);
            }
        }
        float centerPrevFrameDepth = taps[0].depth;
        if (centerPrevFrameDepth != kRayTMax) { float3 worldPos = ray.at(centerPrevFrameDepth); centerPrevFrameDepth = length(worldPos - p.prevPos); }
        bool hasSelection = output.runningSum > 0.f;
        int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
        if (startSampleId == 1) selectedId = 0;
        ExtraProvider curProv{curExt.data()}, tempProv{p.extT.data()};
        for (int i = startSampleId; i < numUsedReservoirs; i++) {
            float talbotMISWeight = 1.f;
            float neighbor_py = 0.f;
            if (taps[i].p_y > 0.f) {
                neighbor_py = taps[i].p_y;
                if (std::isnan(taps[i].runningSum) || std::isinf(taps[i].runningSum)) taps[i].runningSum = 0.f;
                if (i > 0) resampleNeighbor(c, taps[i], ray, sg, tempProv, opt, false);
            } else { taps[i].p_y = 0.f; taps[i].runningSum = 0.f; }
            if (mis == VRESTIR_MIS_TALBOT && taps[i].runningSum > 0.f) {
                float p_sum = 0, p_qi = 0, k = 0;
                for (int j = 0; j < numUsedReservoirs; j++) {
                    int2 tapPos2 = {j == 0 ? x : reprojScreenPos.x, j == 0 ? y : reprojScreenPos.y};
                    float correctedM = fminf(MaxPrevM, taps[j].M);
                    k += correctedM;
                    if (j == 0) { p_qi = taps[i].p_y; p_sum += taps[i].p_y * correctedM; }
                    else if (i == j) { p_qi = neighbor_py; p_sum += neighbor_py * correctedM; }
                    else {
                        // (i == 0, j == 1): evaluate the current sample from the previous frame's camera through the reprojected pixel
                        float3 nOrigin, nDir;
                        if (j == 0) { nOrigin = pos; nDir = normalize(camRayDirNN(U, V, Wv, tapPos2.x, tapPos2.y, W, H)); }
                        else { nOrigin = p.prevPos; nDir = normalize(camRayDirNN(p.prevU, p.prevV, p.prevW, tapPos2.x, tapPos2.y, W, H)); }
                        float usedDepth = j == 0 ? taps[i].depth : (i == 0 ? centerPrevFrameDepth : temporalOriginalDepth);
                        Ray neighborRay = {nOrigin, nDir, 0, usedDepth};
                        float backupDepth = taps[i].depth;
                        taps[i].depth = usedDepth;
                        float p_y = evaluatePHatReadOnly(c, neighborRay, sg, i == 0 ? curProv : tempProv, opt, taps[i], j > 0, true, false);
                        taps[i].depth = backupDepth;
                        if (std::isinf(p_y) || std::isnan(p_y)) p_y = 0.f;
                        p_sum += p_y * correctedM;
                    }
                }
                if (p_sum > 0) talbotMISWeight = p_qi * k / p_sum;
            }
            taps[i].runningSum *= talbotMISWeight;
            bool isCurrentSelected = simpleResampleStepWithMaxM(taps[i], MaxPrevM, output, sg, B);
            hasSelection |= isCurrentSelected;
            if (isCurrentSelected) selectedId = i;
        }
        if (B > 1) {
            if (hasSelection && selectedId > 0) {
                int mib = decodeMaxIndirectBounces(output.sampledPixel, B);
                for (int b = 0; b < mib; b++) curExt[(size_t)centerExtraBounceStartId + b] = p.extT[(size_t)output.extraBounceStartId + b];
            }
            output.extraBounceStartId = centerExtraBounceStartId;
        }
        cur[pixelId] = output;
    });
}

// ---- K3 SpatialReuse: VR/SpatialReuse.cs.slang:94-265 ----
void stageSpatial(Pass& p, const FrameSetup& fs, int roundIdIn, int inBuf) {
    Ctx c(p);
    const vrestir_params& m = p.P;
    const int B = m.mMaxBounces, W = p.W, H = p.H;
    const float3 U = v3(p.cam.cameraU), V = v3(p.cam.cameraV), Wv = v3(p.cam.cameraW), pos = v3(p.cam.posW);
    const SamplingOptions& opt = fs.spatial;
    const std::vector<Reservoir>& in = p.res[inBuf]; std::vector<Reservoir>& out = p.res[1 - inBuf];
    const std::vector<ExtraBounce>& inExt = p.ext[inBuf]; std::vector<ExtraBounce>& outExt = p.ext[1 - inBuf];
    const int gRoundOffset = (m.mEnableTemporalReuse ? 1 : 0) + 1;
    const int r2TimeSeed = ((m.mSpatialReuseRounds + 1) * p.mFrameCount + roundIdIn) % 16;
    const int numRounds = m.mSpatialReuseRounds + gRoundOffset + 1;
    const int roundId = roundIdIn + gRoundOffset;
    const uint32_t mis = m.mSpatialMISMethod;
    const int sampleCount = m.mSpatialSampleCount;
    std::vector<int2> offsets(sampleCount);
    for (int s = 0; s < sampleCount; s++) offsets[s] = generateNeighborOffset(m, s, sampleCount, m.mSampleRadius, r2TimeSeed);
    auto writeExtra = [&](const Reservoir& o, int outStart, int inStart) {
        int mib = decodeMaxIndirectBounces(o.sampledPixel, B);
        for (int b = 0; b < mib; b++) outExt[(size_t)outStart + b] = inExt[(size_t)inStart + b];
    };
    forPixels(p, [&](int x, int y) {
        SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(numRounds * p.mFrameCount + roundId));
        const int pixelId = y * W + x;
        Reservoir output = in[pixelId];
        int centerExtraBounceStartId = output.extraBounceStartId;
        Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, x, y, W, H)), 0, kRayTMax};
        if (mis == VRESTIR_MIS_TALBOT) output = createNewReservoir();
        bool hasSelection = output.runningSum > 0.f;
        Features centerFeatures = p.feat[pixelId];
        bool IsSelfBackground = !(centerFeatures.transmittance != 1.f);
        if (IsSelfBackground) {
            if (mis == VRESTIR_MIS_TALBOT) out[pixelId] = in[pixelId]; else out[pixelId] = output;
            if (B > 1) writeExtra(output, output.extraBounceStartId, output.extraBounceStartId);
            return;
        }
        ExtraProvider prov{inExt.data()};
        int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
        for (int sampleId = startSampleId; sampleId < sampleCount; sampleId++) {
            int tx = x + offsets[sampleId].x, ty = y + offsets[sampleId].y;
            if (!IsWithinRange(tx, ty, W, H)) continue;
            Reservoir tap = in[ty * W + tx];
            float MISWeight = 1.f;
            if (sampleId > 0) resampleNeighbor(c, tap, ray, sg, prov, opt, true);
            if (mis == VRESTIR_MIS_TALBOT && tap.runningSum > 0.f) {
                float p_sum = 0, p_qi = 0, k = 0;
                for (int j = 0; j < sampleCount; j++) {
                    int tx2 = x + offsets[j].x, ty2 = y + offsets[j].y;
                    if (!IsWithinRange(tx2, ty2, W, H)) continue;
                    const Reservoir& tap2 = in[ty2 * W + tx2];
                    k += tap2.M;
                    if (j == 0) { p_qi = tap.p_y; p_sum += tap.p_y * tap2.M; }
                    else if (sampleId == j) { p_qi = tap2.p_y; p_sum += tap2.p_y * tap2.M; }
                    else {
                        float3 neighborRayDir = normalize(camRayDirNN(U, V, Wv, tx2, ty2, W, H));
                        Ray neighborRay = {ray.origin, neighborRayDir, 0, tap.depth};
                        float p_y = evaluatePHatReadOnly(c, neighborRay, sg, prov, opt, tap, false, true, true);
                        if (std::isinf(p_y) || std::isnan(p_y)) p_y = 0.f;
                        p_sum += p_y * tap2.M;
                    }
                }
                if (p_sum > 0) MISWeight = p_qi * k / p_sum;
            }
            tap.runningSum *= MISWeight;
            bool isSelected = simpleResampleStep(tap, output, sg, B);
            if (isSelected) hasSelection = true;
        }
        if (B > 1) {
            if (hasSelection) writeExtra(output, centerExtraBounceStartId, output.extraBounceStartId);
            output.extraBounceStartId = centerExtraBounceStartId;
        }
        out[pixelId] = output;
    });
}

// ---- K4 CopyReservoirs: VR/CopyReservoirs.cs.slang:84-95 ----
void stageCopy(Pass& p, int srcBuf) {
    const int B = p.P.mMaxBounces;
    forPixels(p, [&](int x, int y) {
        int id = y * p.W + x;
        Reservoir r = p.res[srcBuf][id];
        p.resT[id] = r;
        if (B > 1) {
            int mib = decodeMaxIndirectBounces(r.sampledPixel, B);
            for (int b = 0; b < mib; b++) p.extT[(size_t)r.extraBounceStartId + b] = p.ext[srcBuf][(size_t)r.extraBounceStartId + b];
        }
    });
}

// ---- K5 FinalShading: VR/FinalShading.cs.slang:71-141 ----
void stageFinal(Pass& p, const FrameSetup& fs, int buf, float* out_color) {
    Ctx c(p);
    const vrestir_params& m = p.P;
    const float3 U = v3(p.cam.cameraU), V = v3(p.cam.cameraV), Wv = v3(p.cam.cameraW), pos = v3(p.cam.posW);
    const bool noReuse = !m.mEnableSpatialReuse && !m.mEnableTemporalReuse;
    const int frame = p.mFreezeFrame ? p.mFrameCount - 1 : p.mFrameCount;
    forPixels(p, [&](int x, int y) {
        size_t pixelId = (size_t)y * p.W + x;
        float3 outputColor = f3(0.f);
        SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fs.numTotalRounds * frame + fs.numTotalRounds - 1));
        if (m.mUseReference) {
            outputColor = f3(p.refColor[pixelId * 4], p.refColor[pixelId * 4 + 1], p.refColor[pixelId * 4 + 2]);
        } else if (m.mVisualizeTotalTransmittance) {
            outputColor = f3(powf(p.feat[pixelId].transmittance, 2.2f));
        } else {
            Reservoir cur = p.res[buf][pixelId];
            if (cur.runningSum > 0.f) {
                Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, x, y, p.W, p.H)), 0, cur.depth};
                ExtraProvider prov{p.ext[buf].data()};
                uint64_t taps0 = tl_cnt.taps;
                float3 col = evaluate_F(c, cur, prov, ray, sg, fs.fin, noReuse);
                if (getenv("VRO_DEBUG_HEAVY") && tl_cnt.taps - taps0 > 20000)
                    fprintf(stderr, "heavy pixel %d %d taps %llu depth %g lightID %d uv %.9g %.9g\n", x, y, (unsigned long long)(tl_cnt.taps - taps0), cur.depth, cur.lightID, cur.lightUV.x, cur.lightUV.y);
                float Wt = cur.p_y == 0.0f ? 1.f : cur.runningSum / (cur.p_y * cur.M);
                col *= Wt;
                outputColor += col;
            }
        }
        float o[4] = {outputColor.x, outputColor.y, outputColor.z, 1.f};
        bool bad = false;
        for (int i = 0; i < 4; i++) if (std::isnan(o[i]) || std::isinf(o[i])) bad = true;
        if (bad) o[0] = o[1] = o[2] = o[3] = 0.f;
        if (out_color) memcpy(&out_color[pixelId * 4], o, 16);
    });
}

// ---- importance map: F/Experimental/Scene/Lights/EnvMapSamplerSetup.cs.slang:48-75, EnvMapSampler.cpp:83-116 ----
void buildImportance(Pass& p, const float* given) {
    const int dim = 512, spp = 64;
    p.impDim = dim;
    int mips = 0; { int d = dim; while (d >= 1) { mips++; d >>= 1; } }
    p.impBaseMip = mips - 1;
    p.impOffset.resize(mips);
    size_t total = 0; for (int i = 0; i < mips; i++) { p.impOffset[i] = total; total += (size_t)(dim >> i) * (dim >> i); }
    p.importance.assign(total, 0.f);
    if (given) { memcpy(p.importance.data(), given, total * sizeof(float)); return; }
    const int sx = std::max(1, (int)std::sqrt((double)spp)), sy = spp / sx;
    const float invSamples = 1.f / (float)(sx * sy);
    const float dimSx = (float)(dim * sx), dimSy = (float)(dim * sy);
    int save_cx0 = p.cx0, save_cy0 = p.cy0, save_cx1 = p.cx1, save_cy1 = p.cy1;
    p.cx0 = 0; p.cy0 = 0; p.cx1 = dim; p.cy1 = dim;
    forPixels(p, [&](int px, int py) {
        float L = 0.f;
        for (int y = 0; y < sy; y++)
            for (int x = 0; x < sx; x++) {
                uint32_t spx = (uint32_t)px * sx + x, spy = (uint32_t)py * sy + y;
                float2 pp = {((float)spx + 0.5f) / dimSx, ((float)spy + 0.5f) / dimSy};
                float3 dir = oct_to_ndir_equal_area_unorm(pp);
                float2 uv = world_to_latlong_map(dir);
                L += luminance(envBilinear(p, uv));
            }
        p.importance[(size_t)py * dim + px] = L * invSamples;
    });
    p.cx0 = save_cx0; p.cy0 = save_cy0; p.cx1 = save_cx1; p.cy1 = save_cy1;
    for (int mip = 1; mip < mips; mip++) {
        int d = dim >> mip, dp = dim >> (mip - 1);
        const float* src = &p.importance[p.impOffset[mip - 1]]; float* dst = &p.importance[p.impOffset[mip]];
        for (int y = 0; y < d; y++)
            for (int x = 0; x < d; x++) {
                float a = src[(size_t)(2 * y) * dp + 2 * x], b = src[(size_t)(2 * y) * dp + 2 * x + 1];
                float cc = src[(size_t)(2 * y + 1) * dp + 2 * x], dd = src[(size_t)(2 * y + 1) * dp + 2 * x + 1];
                dst[(size_t)y * d + x] = ((a + b) + (cc + dd)) * 0.25f;
            }
    }
}

int finalBuffer(const Pass& p) {  // totalRoundId % 2 after the spatial rounds (VR/VolumetricReSTIR.cpp:643-697)
    int rounds = (p.P.mEnableSpatialReuse && !p.P.mUseReference) ? p.P.mSpatialReuseRounds : 0;
    return rounds % 2;
}

int runStage(Pass& p, int stage, int arg, float* out_color, float* out_mvec) {
    if (!p.haveVolume || !p.haveCamera || p.W <= 0) return fail(VRESTIR_ERR_NOT_READY, "volume/camera/frame not set");
    if (p.P.mUseSurfaceScene) return fail(VRESTIR_ERR_UNSUPPORTED, "surface scene out of scope");
    ensureBuffers(p);
    applyOverrides(p);
    FrameSetup fs = buildFrameSetup(p.P);
    const vrestir_params& m = p.P;
    {   // every mip the options name must be bound
        auto ok = [&](int slot) { return slot >= 0 && slot < VRESTIR_MAX_SLOTS && p.slots[slot].valid; };
        bool good = ok(0) && ok(fs.initial.lightingMipLevel) && (m.mEnableSpatialReuse || m.mEnableTemporalReuse ? ok(fs.initial.visibilityMipLevel) : true) &&
                    ok(fs.spatial.visibilityMipLevel) && ok(fs.spatial.lightingMipLevel) &&
                    (m.mEnableTemporalReuse && m.mTemporalReprojectionMode != VRESTIR_REPROJECTION_NONE ? ok(VRESTIR_NUM_MAX_MIPS + m.mTemporalReprojectionMipLevel) : true);
        if (!good) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "an option names a mip level that the volume does not have");
    }
    double t0 = nowMs();
    switch (stage) {
        case 0:
            if (p.mOptionsChanged) {   // VR/VolumetricReSTIR.cpp:349-359
                if (p.mRandomizeFrameSeed) p.mFrameCount = rand_r(&p.randState) % 65536; else p.mFrameCount = 0;
                p.mTemporalSampleAccumulated = 0; p.mOptionsChanged = false;
            }
            if (!p.mFreezeFrame) stageFeatures(p, fs);
            p.ms.features_ms = (float)(nowMs() - t0); break;
        case 1: if (!p.mFreezeFrame) stageInitial(p, fs); p.ms.initial_ms = (float)(nowMs() - t0); break;
        case 2:
            if (!p.mFreezeFrame && !m.mUseReference && m.mEnableTemporalReuse) {
                stageTemporal(p, fs, out_mvec);
                if (!m.mEnableSpatialReuse) { p.resT = p.res[0]; if (m.mMaxBounces > 1) p.extT = p.ext[0]; }   // :629-634
                p.featT = p.feat;                                                                               // :636
            }
            p.ms.temporal_ms = (float)(nowMs() - t0); break;
        case 3:
            if (!p.mFreezeFrame && !m.mUseReference && m.mEnableSpatialReuse) stageSpatial(p, fs, arg, arg % 2);
            if (arg == 0) p.ms.spatial_ms = 0;
            p.ms.spatial_ms += (float)(nowMs() - t0); break;
        case 4:
            if (!p.mFreezeFrame && !m.mUseReference && m.mEnableTemporalReuse) stageCopy(p, finalBuffer(p));
            p.ms.copy_ms = (float)(nowMs() - t0); break;
        case 5: stageFinal(p, fs, finalBuffer(p), out_color); p.ms.final_ms = (float)(nowMs() - t0); break;
        case 6:   // VR/VolumetricReSTIR.cpp:765-772
            p.mTemporalSampleAccumulated = 1;
            memcpy(p.prevView, p.cam.viewMat, 64); memcpy(p.prevProj, p.cam.projMat, 64);
            p.prevU = v3(p.cam.cameraU); p.prevV = v3(p.cam.cameraV); p.prevW = v3(p.cam.cameraW); p.prevPos = v3(p.cam.posW);
            if (!p.mFreezeFrame) p.mFrameCount++;
            break;
        default: return fail(VRESTIR_ERR_INVALID_ARGUMENT, "bad stage");
    }
    return VRESTIR_OK;
}

struct KeyDesc { const char* name; size_t off; int type; };  // 0 int32, 1 uint32, 2 float
#define K_I(f) {#f, offsetof(vrestir_params, f), 0}
#define K_U(f) {#f, offsetof(vrestir_params, f), 1}
#define K_F(f) {#f, offsetof(vrestir_params, f), 2}
const KeyDesc kKeys[] = {
    K_I(mMaxBounces), K_I(mEnableTemporalReuse), K_I(mEnableSpatialReuse), K_I(mVertexReuse), K_I(mVertexReuseStartBounce), K_I(mUseReference),
    K_I(mUseEnvironmentLights), K_I(mUseAnalyticLights), K_I(mUseEmissiveLights), K_I(mBaselineSamplePerPixel), K_I(mVisualizeTotalTransmittance),
    K_I(mUseSurfaceScene), K_I(mUsePrevVolumeForReproj), K_I(mInitialBaseMipLevel), K_I(mInitialM), K_I(mInitialLightSamples), K_I(mInitialLightingMipLevel),
    K_I(mInitialVisibilityUseLinearSampler), K_I(mInitialLightingUseLinearSampler), K_U(mInitialLightingTrackingMethod), K_F(mInitialVisibilityTStepScale),
    K_F(mInitialLightingTStepScale), K_I(mInitialUseRussianRoulette), K_I(mInitialUseCoarserGridForIndirectBounce), K_F(mTemporalReuseMThreshold),
    K_U(mTemporalReprojectionMode), K_U(mTemporalMISMethod), K_I(mTemporalReprojectionMipLevel), K_I(mSpatialReuseRounds), K_I(mSpatialVisibilityMipLevel),
    K_I(mSpatialLightingMipLevel), K_I(mSpatialVisibilityUseLinearSampler), K_I(mSpatialLightingUseLinearSampler), K_F(mSpatialVisibilityTStepScale),
    K_F(mSpatialLightingTStepScale), K_U(mSpatialVisibilityTrackingMethod), K_U(mSpatialLightingTrackingMethod), K_U(mRandomSamplerType), K_F(mSampleRadius),
    K_I(mSpatialSampleCount), K_I(mEnableVisibilitySimilarityRejection), K_U(mSpatialMISMethod), K_I(mFinalLightSamples), K_I(mFinalVisibilitySamples),
    K_U(mFinalVisibilityTrackingMethod), K_U(mFinalLightTrackingMethod), K_U(mFinalRandomSamplerType), K_F(mFinalTStepScale)};

}  // namespace

// ================================================================================================ C API
extern "C" {

const char* vro_last_error(void) { return g_err.c_str(); }

int vro_create(const vrestir_params* params, vro_pass** out) {
    if (!params || !out) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    auto* p = new vro_pass();
    p->P = *params;
    *out = p;
    return VRESTIR_OK;
}
int vro_destroy(vro_pass* p) { delete p; return VRESTIR_OK; }
int vro_set_threads(vro_pass* p, int threads) { p->threads = threads; return VRESTIR_OK; }
int vro_get_threads(const vro_pass* p) { return p->threads > 0 ? p->threads : (int)std::thread::hardware_concurrency(); }

int vro_set_volume(vro_pass* p, const vrestir_grid_desc* g) {
    if (!p || !g) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    p->volBase = g->volume;
    memcpy(p->slots, g->slots, sizeof(p->slots));
    if (g->blackbody_lut) p->lut.assign(g->blackbody_lut, g->blackbody_lut + 512); else p->lut.clear();
    p->haveVolume = true; p->mOptionsChanged = true;
    applyOverrides(*p);
    return VRESTIR_OK;
}
int vro_advance_volume(vro_pass* p, const vrestir_grid_desc* g) {
    if (!p || !g || !p->haveVolume) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "advance before set_volume");
    // prev-frame rebinding (F/Scene/Scene.cpp:3285-3295): density mips -> 19.., temperature -> 27, velocity -> 28
    for (int i = 0; i < VRESTIR_NUM_MAX_MIPS; i++) if (VRESTIR_PREV_DENSITY_GRID_OFFSET + i < VRESTIR_MAX_SLOTS) p->slots[VRESTIR_PREV_DENSITY_GRID_OFFSET + i] = p->slots[i];
    p->slots[VRESTIR_TEMPERATURE_GRID_ID + VRESTIR_PREV_EXTRA_GRID_OFFSET] = p->slots[VRESTIR_TEMPERATURE_GRID_ID];
    p->slots[VRESTIR_VELOCITY_GRID_ID + VRESTIR_PREV_EXTRA_GRID_OFFSET] = p->slots[VRESTIR_VELOCITY_GRID_ID];
    int lastHasEmission = p->volBase.hasEmission;
    for (int i = 0; i < VRESTIR_PREV_DENSITY_GRID_OFFSET - 1; i++) p->slots[i] = g->slots[i];
    p->volBase = g->volume;
    p->volBase.lastFrameHasEmission = lastHasEmission;
    p->volBase.hasAnimation = 1;
    applyOverrides(*p);
    return VRESTIR_OK;
}
int vro_set_camera(vro_pass* p, const vrestir_camera* cam) {
    if (!p || !cam) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    p->cam = *cam; p->haveCamera = true; return VRESTIR_OK;
}
int vro_set_envmap(vro_pass* p, const vrestir_envmap_desc* env, const float* importance) {
    if (!p || !env || !env->texels) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    p->envW = env->width; p->envH = env->height;
    p->envTexels.assign(env->texels, env->texels + (size_t)env->width * env->height * 4);
    p->envIntensity = env->intensity; p->envTint = v3(env->tint);
    memcpy(p->envT, env->transform, 36); memcpy(p->envInvT, env->invTransform, 36);
    memcpy(p->envPrevT, env->prevTransform, 36); memcpy(p->envPrevInvT, env->prevInvTransform, 36);
    p->haveEnv = true;
    buildImportance(*p, importance);
    p->mOptionsChanged = true;
    return VRESTIR_OK;
}
int vro_set_env_alias(vro_pass* p, const float* thr, const uint32_t* redirect, const float* pdf, int count) {
    if (!p || !thr || !redirect || count <= 0) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "bad alias table");
    p->envAliasThr.assign(thr, thr + count); p->envAliasRedirect.assign(redirect, redirect + count);
    if (pdf) p->envAliasPdf.assign(pdf, pdf + count);
    return VRESTIR_OK;
}
int vro_set_analytic_lights(vro_pass* p, const vrestir_light* lights, int count) {
    p->lights.assign(lights, lights + std::max(0, count)); p->mOptionsChanged = true; return VRESTIR_OK;
}
int vro_set_emissive_triangles(vro_pass* p, const vrestir_emissive_triangle* tris, int count, const uint32_t* items, const float* weights, float weight_sum, float mul) {
    p->tris.assign(tris, tris + std::max(0, count));
    p->alias.resize(std::max(0, count));
    if (count > 0) memcpy(p->alias.data(), items, (size_t)count * 16);
    p->aliasWeights.assign(weights, weights + std::max(0, count));
    p->aliasWeightSum = weight_sum; p->emissiveMul = mul; p->mOptionsChanged = true;
    return VRESTIR_OK;
}
int vro_set_frame(vro_pass* p, int w, int h) {
    if (w <= 0 || h <= 0) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "bad frame size");
    p->W = w; p->H = h; p->cx0 = 0; p->cy0 = 0; p->cx1 = w; p->cy1 = h; p->mOptionsChanged = true; return VRESTIR_OK;
}
int vro_set_crop(vro_pass* p, int x0, int y0, int x1, int y1) {
    p->cx0 = std::max(0, x0); p->cy0 = std::max(0, y0); p->cx1 = std::min(p->W, x1); p->cy1 = std::min(p->H, y1); return VRESTIR_OK;
}
int vro_update(vro_pass* p, const char* key, double value) {
    if (!p || !key) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    std::string k(key);
    if (k.rfind("mParams.", 0) == 0) k = k.substr(8);
    bool found = false;
    for (const auto& kd : kKeys)
        if (k == kd.name) {
            char* base = (char*)&p->P + kd.off;
            if (kd.type == 0) *(int32_t*)base = (int32_t)value; else if (kd.type == 1) *(uint32_t*)base = (uint32_t)value; else *(float*)base = (float)value;
            found = true; break;
        }
    if (!found) {
        found = true;
        if (k == "mOutputMotionVec") p->mOutputMotionVec = value != 0;
        else if (k == "mFreezeFrame") p->mFreezeFrame = value != 0;
        else if (k == "volumeDensityScaleExtraControl") p->volumeDensityScaleExtraControl = (float)value;
        else if (k == "volumeAlbedoExtraControl") p->volumeAlbedoExtraControl = (float)value;
        else if (k == "volumeAnisotropyExtraControl") p->volumeAnisotropyExtraControl = (float)value;
        else if (k == "mEnvSamplerType") p->envSamplerType = (int)value;
        else if (k == "randomizeFrameSeed") { if (!p->mRandomizeFrameSeed) p->randState = 123; p->mRandomizeFrameSeed = true; }
        else found = false;
    }
    p->mOptionsChanged = true;   // VR/VolumetricReSTIR.cpp:1339
    if (!found) return fail(VRESTIR_WARN_UNKNOWN_KEY, std::string("Unknown field '") + key + "'");
    return VRESTIR_OK;
}
int vro_set_params(vro_pass* p, const vrestir_params* params) { p->P = *params; p->mOptionsChanged = true; return VRESTIR_OK; }
int vro_get_params(const vro_pass* p, vrestir_params* out) { *out = p->P; return VRESTIR_OK; }
int vro_set_frame_count(vro_pass* p, int fc, int acc) { p->mFrameCount = fc; p->mTemporalSampleAccumulated = acc; p->mOptionsChanged = false; return VRESTIR_OK; }
int vro_get_frame_count(const vro_pass* p, int* fc) { *fc = p->mFrameCount; return VRESTIR_OK; }

int vro_execute_stage(vro_pass* p, int stage, int arg, float* out_color, float* out_mvec) { return runStage(*p, stage, arg, out_color, out_mvec); }
int vro_execute(vro_pass* p, float* out_color, float* out_mvec) {
    double t0 = nowMs();
    int rc;
    if ((rc = runStage(*p, 0, 0, out_color, out_mvec))) return rc;
    if ((rc = runStage(*p, 1, 0, out_color, out_mvec))) return rc;
    if ((rc = runStage(*p, 2, 0, out_color, out_mvec))) return rc;
    p->ms.spatial_ms = 0;
    if (p->P.mEnableSpatialReuse) for (int r = 0; r < p->P.mSpatialReuseRounds; r++) if ((rc = runStage(*p, 3, r, out_color, out_mvec))) return rc;
    if ((rc = runStage(*p, 4, 0, out_color, out_mvec))) return rc;
    if ((rc = runStage(*p, 5, 0, out_color, out_mvec))) return rc;
    if ((rc = runStage(*p, 6, 0, out_color, out_mvec))) return rc;
    p->ms.total_ms = (float)(nowMs() - t0);
    return VRESTIR_OK;
}

int vro_buffer_bytes(const vro_pass* p, int buffer, size_t* bytes) {
    size_t n = (size_t)p->W * p->H; int B = p->P.mMaxBounces;
    switch (buffer) {
        case VRESTIR_BUF_RESERVOIR_0: case VRESTIR_BUF_RESERVOIR_1: case VRESTIR_BUF_RESERVOIR_TEMPORAL: *bytes = n * sizeof(vrestir_reservoir); break;
        case VRESTIR_BUF_EXTRA_0: case VRESTIR_BUF_EXTRA_1: case VRESTIR_BUF_EXTRA_TEMPORAL: *bytes = n * (size_t)std::max(0, B - 1) * 12; break;
        case VRESTIR_BUF_FEATURES: case VRESTIR_BUF_FEATURES_TEMPORAL: *bytes = n * 8; break;
        case VRESTIR_BUF_ENV_IMPORTANCE: *bytes = p->importance.size() * 4; break;
        case VRESTIR_BUF_PPARTIAL_0: case VRESTIR_BUF_PPARTIAL_1: case VRESTIR_BUF_PPARTIAL_TEMPORAL: *bytes = n * 4; break;
        default: return fail(VRESTIR_ERR_INVALID_ARGUMENT, "bad buffer id");
    }
    return VRESTIR_OK;
}
static std::vector<Reservoir>* resBuf(vro_pass* p, int b) {
    return (b == VRESTIR_BUF_RESERVOIR_0 || b == VRESTIR_BUF_PPARTIAL_0) ? &p->res[0] : (b == VRESTIR_BUF_RESERVOIR_1 || b == VRESTIR_BUF_PPARTIAL_1) ? &p->res[1] : &p->resT;
}
static bool isPPartial(int b) { return b >= VRESTIR_BUF_PPARTIAL_0 && b <= VRESTIR_BUF_PPARTIAL_TEMPORAL; }
static std::vector<ExtraBounce>* extBuf(vro_pass* p, int b) { return b == VRESTIR_BUF_EXTRA_0 ? &p->ext[0] : b == VRESTIR_BUF_EXTRA_1 ? &p->ext[1] : &p->extT; }
int vro_get_buffer(vro_pass* p, int buffer, void* dst, size_t bytes) {
    ensureBuffers(*p);
    size_t need; int rc = vro_buffer_bytes(p, buffer, &need); if (rc) return rc;
    if (need != bytes) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "size mismatch");
    if (isPPartial(buffer)) { auto* v = resBuf(p, buffer); float* o = (float*)dst; for (size_t i = 0; i < v->size(); i++) o[i] = (*v)[i].p_partial; }
    else if (buffer <= VRESTIR_BUF_RESERVOIR_TEMPORAL) {
        auto* v = resBuf(p, buffer); auto* o = (vrestir_reservoir*)dst;
        for (size_t i = 0; i < v->size(); i++) { const Reservoir& r = (*v)[i]; o[i] = {r.runningSum, r.M, r.depth, r.p_y, {r.lightUV.x, r.lightUV.y}, r.lightID, r.sampledPixel}; }
    } else if (buffer <= VRESTIR_BUF_EXTRA_TEMPORAL) { if (bytes) memcpy(dst, extBuf(p, buffer)->data(), bytes); }
    else if (buffer == VRESTIR_BUF_FEATURES) memcpy(dst, p->feat.data(), bytes);
    else if (buffer == VRESTIR_BUF_FEATURES_TEMPORAL) memcpy(dst, p->featT.data(), bytes);
    else memcpy(dst, p->importance.data(), bytes);
    return VRESTIR_OK;
}
int vro_set_buffer(vro_pass* p, int buffer, const void* src, size_t bytes) {
    ensureBuffers(*p);
    size_t need; int rc = vro_buffer_bytes(p, buffer, &need); if (rc) return rc;
    if (need != bytes) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "size mismatch");
    int B = p->P.mMaxBounces;
    if (isPPartial(buffer)) { auto* v = resBuf(p, buffer); const float* o = (const float*)src; for (size_t i = 0; i < v->size(); i++) (*v)[i].p_partial = o[i]; }
    else if (buffer <= VRESTIR_BUF_RESERVOIR_TEMPORAL) {
        auto* v = resBuf(p, buffer); auto* o = (const vrestir_reservoir*)src;
        for (size_t i = 0; i < v->size(); i++)   // p_partial (its own buffer id) is kept
            (*v)[i] = {o[i].runningSum, o[i].M, o[i].depth, o[i].p_y, {o[i].lightUV[0], o[i].lightUV[1]}, o[i].lightID, o[i].sampledPixel, B > 1 ? (int)i * (B - 1) : 0, (*v)[i].p_partial};
    } else if (buffer <= VRESTIR_BUF_EXTRA_TEMPORAL) { if (bytes) memcpy(extBuf(p, buffer)->data(), src, bytes); }
    else if (buffer == VRESTIR_BUF_FEATURES) memcpy(p->feat.data(), src, bytes);
    else if (buffer == VRESTIR_BUF_FEATURES_TEMPORAL) memcpy(p->featT.data(), src, bytes);
    else memcpy(p->importance.data(), src, bytes);
    return VRESTIR_OK;
}
int vro_spatial_input_buffer(const vro_pass* p, int round, int* buffer) { (void)p; *buffer = round % 2; return VRESTIR_OK; }
int vro_get_counters(vro_pass* p, vro_counters* out, int reset) {
    *out = {p->cnt.taps, p->cnt.voxels, p->cnt.vbytes, p->cnt.nodes, p->cnt.rng, p->cnt.marches};
    if (reset) p->cnt = Counters{};
    return VRESTIR_OK;
}
int vro_get_stage_ms(vro_pass* p, vrestir_timings* out) { *out = p->ms; return VRESTIR_OK; }

// ---- unit hooks ----
void vro_rng_words(uint32_t px, uint32_t py, uint32_t n, int count, uint32_t* words, float* floats) {
    SampleGenerator sg = SampleGenerator::create(px, py, n);
    for (int i = 0; i < count; i++) { uint32_t w = sg.next(); if (words) words[i] = w; if (floats) floats[i] = (float)(w >> 8) * 0x1p-24f; }
}
uint32_t vro_morton(uint32_t x, uint32_t y) { return interleave_32bit(x, y); }
float vro_transmittance(vro_pass* p, const float o[3], const float d[3], float tmax, int method, int mip, int linear, float tss, uint32_t spx, uint32_t spy, uint32_t sn) {
    applyOverrides(*p);
    Ctx c(*p); SampleGenerator sg = SampleGenerator::create(spx, spy, sn);
    Ray r = {v3(o), v3(d), 0, tmax};
    return computeVisibility(c, r, sg, 1, mip, linear != 0, (uint32_t)method, tss);
}
// Records, from the next K1 on, the generator of every pixel after the candidate loop and after the p-hat evaluation (draw-count check)
int vro_record_k1_generator(vro_pass* p, int on) { if (on) p->k1Generator.assign((size_t)p->W * p->H * 8, 0u); else p->k1Generator.clear(); return VRESTIR_OK; }
int vro_get_k1_generator(vro_pass* p, int px, int py, uint32_t out8[8]) {
    if (p->k1Generator.size() != (size_t)p->W * p->H * 8 || px < 0 || py < 0 || px >= p->W || py >= p->H) return fail(VRESTIR_ERR_INVALID_ARGUMENT, "no K1 generator record");
    for (int i = 0; i < 8; i++) out8[i] = p->k1Generator[((size_t)py * p->W + px) * 8 + i];
    return VRESTIR_OK;
}
// computeVisibility with `samples` estimates; also returns the generator after the call (how many draws the walk consumed)
float vro_visibility_state(vro_pass* p, const float o[3], const float d[3], float tmax, int method, int mip, int linear, float tss, int samples,
                           uint32_t spx, uint32_t spy, uint32_t sn, uint32_t out_state[4]) {
    applyOverrides(*p);
    Ctx c(*p); SampleGenerator sg = SampleGenerator::create(spx, spy, sn);
    if (spx == 0xFFFFFFFFu && spy == 0xFFFFFFFFu) for (int i = 0; i < 4; i++) sg.s[i] = out_state[i];   // start from a given generator state
    Ray r = {v3(o), v3(d), 0, tmax};
    float v = computeVisibility(c, r, sg, samples, mip, linear != 0, (uint32_t)method, tss);
    for (int i = 0; i < 4; i++) out_state[i] = sg.s[i];
    return v;
}
// SampleMediumAnalyticGeneric along one ray with the generator of (pixel, sample number): out12 = hit distances | pdfs | transmittances
// (4 each), out_state = the generator after the call
void vro_sample_distances(vro_pass* p, const float o[3], const float d[3], int mip, int linear, int n, uint32_t spx, uint32_t spy, uint32_t sn, float out12[12], uint32_t out_state[4]) {
    applyOverrides(*p);
    Ctx c(*p); SampleGenerator sg = SampleGenerator::create(spx, spy, sn);
    Ray r = {v3(o), v3(d), 0, kRayTMax};
    float hd[4] = {0, 0, 0, 0}, pd[4] = {0, 0, 0, 0}, ot[4] = {0, 0, 0, 0};
    SampleMediumAnalyticGeneric(c, r, sg, linear != 0, hd, mip, pd, ot, n);
    for (int i = 0; i < 4; i++) { out12[i] = hd[i]; out12[4 + i] = pd[i]; out12[8 + i] = ot[i]; }
    out_state[0] = sg.s[0]; out_state[1] = sg.s[1]; out_state[2] = sg.s[2]; out_state[3] = sg.s[3];
}
// SampleMediumSuperVoxelGeneric along one ray: the interaction's distance from the origin (negative: the ray left the volume), and
// the generator after the call
float vro_sample_supervoxel(vro_pass* p, const float o[3], const float d[3], int mip, uint32_t spx, uint32_t spy, uint32_t sn, uint32_t out_state[4]) {
    applyOverrides(*p);
    Ctx c(*p); SampleGenerator sg = SampleGenerator::create(spx, spy, sn);
    Ray r = {v3(o), v3(d), 0, kRayTMax};
    MediumInteraction mi{}; mi.isValid = false;
    SampleMediumSuperVoxelGeneric(c, r, sg, mi, mip);
    for (int i = 0; i < 4; i++) out_state[i] = sg.s[i];
    return mi.isValid ? length(mi.p - r.origin) : -1.f;
}
// p-hat of a single-bounce reservoir (depth, lightUV, lightID) of pixel (px, py) under the spatial options: what K1's finish, K2 and K3 evaluate
float vro_p_hat(vro_pass* p, int px, int py, float depth, float uvx, float uvy, int lightID) {
    applyOverrides(*p);
    Ctx c(*p);
    const FrameSetup fs = buildFrameSetup(p->P);
    const float3 U = v3(p->cam.cameraU), V = v3(p->cam.cameraV), Wv = v3(p->cam.cameraW), pos = v3(p->cam.posW);
    Ray ray = {pos, normalize(camRayDirNN(U, V, Wv, px, py, p->W, p->H)), 0, kRayTMax};
    Reservoir tap = createNewReservoir();
    tap.runningSum = 1.f; tap.M = 1.f; tap.depth = depth; tap.p_y = 1.f; tap.lightUV = {uvx, uvy}; tap.lightID = lightID;
    SampleGenerator sg = SampleGenerator::create(0, 0, 0);
    ExtraProvider prov{nullptr};
    return evaluate_P_hat(c, ray, sg, prov, fs.spatial, tap, false, true, false);
}
float vro_density_world(vro_pass* p, const float pos[3], int mip) { applyOverrides(*p); Ctx c(*p); return DensityWorldSpace(c, v3(pos), mip); }
int vro_dump_brick_visits(vro_pass* p, const float o[3], const float d[3], int mip, int vertex_center, int max_cells, int32_t* out_xyz, float* out_t) {
    applyOverrides(*p);
    p->dumping = true; p->dumpMax = max_cells; p->dumpXYZ.clear(); p->dumpT.clear();
    Ctx c(*p); SampleGenerator sg{};
    Ray r = {v3(o), v3(d), 0, kRayTMax};
    MediumTrAnalyticAdapter a; a.useLinearSampler = vertex_center != 0;
    VolumeTrackingGVDB(c, r, mip, sg, a, vertex_center != 0);
    p->dumping = false;
    int n = (int)p->dumpT.size();
    for (int i = 0; i < n; i++) { out_xyz[3 * i] = p->dumpXYZ[3 * i]; out_xyz[3 * i + 1] = p->dumpXYZ[3 * i + 1]; out_xyz[3 * i + 2] = p->dumpXYZ[3 * i + 2]; out_t[i] = p->dumpT[i]; }
    return n;
}
int32_t vro_encode_max_indirect_bounces(int32_t s, int32_t b) { return encodeMaxIndirectBounces(s, b); }
int32_t vro_decode_max_indirect_bounces(int32_t s, int32_t mb) { return decodeMaxIndirectBounces(s, mb); }
int32_t vro_encode_path_tag(int32_t s, int32_t t) { return encodePathTag(s, t); }
int32_t vro_decode_path_tag(int32_t s) { return decodePathTag(s); }
void vro_encode_wi_dist(const float in4[4], float out3[3]) { float3 r = encodeWiDist({in4[0], in4[1], in4[2], in4[3]}); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
void vro_decode_wi_dist(const float in3[3], float out4[4]) { float4_ r = decodeWiDist(v3(in3)); out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w; }
void vro_env_eval(vro_pass* p, const float dir[3], float out_rgb[3]) { float3 r = envEval(*p, v3(dir)); out_rgb[0] = r.x; out_rgb[1] = r.y; out_rgb[2] = r.z; }
int vro_env_sample(vro_pass* p, float u0, float u1, float out_dir[3], float* out_pdf, float out_Le[3]) {
    EnvMapSample s{}; envSample(*p, {u0, u1}, s);
    out_dir[0] = s.dir.x; out_dir[1] = s.dir.y; out_dir[2] = s.dir.z; *out_pdf = s.pdf; out_Le[0] = s.Le.x; out_Le[1] = s.Le.y; out_Le[2] = s.Le.z;
    return 1;
}
float vro_phase_hg(float cos_theta, float g) { return PhaseHG(cos_theta, g); }
float vro_sample_phase(float g, const float wo[3], float u0, float u1, float out_wi[3]) {
    MediumInteraction mi{f3(0.f), v3(wo), g, true};
    float3 wi = f3(0.f);
    const float pdf = mi.Sample_p(v3(wo), wi, {u0, u1});
    out_wi[0] = wi.x; out_wi[1] = wi.y; out_wi[2] = wi.z;
    return pdf;
}
void vro_neighbor_offsets(vro_pass* p, int frame_count, int round, int32_t* out_xy) {
    const auto& m = p->P;
    int seed = ((m.mSpatialReuseRounds + 1) * frame_count + round) % 16;
    for (int s = 0; s < m.mSpatialSampleCount; s++) { int2 o = generateNeighborOffset(m, s, m.mSpatialSampleCount, m.mSampleRadius, seed); out_xy[2 * s] = o.x; out_xy[2 * s + 1] = o.y; }
}

}  // extern "C"
