// vr_scene.cpp — host-side scene data for the VolumetricReSTIR hot path: procedural sparse density grids, the
// mip / conservative-mip chain, the GVDB-style 5-4-3 tree + brick pool, VolumeDesc, camera / env-map / light helpers.
//
// This is the data contract of the reference's scene loader and offline converter, rebuilt for synthetic inputs:
//   tree + atlas + VolumeDesc      F/Scene/Scene.cpp:2898-3298 (addGVDBVolume), F/Scene/GVDB/gvdbNodes.slang:38-95
//   brick (min,max,avg) bounds      F/Scene/Scene.cpp:2981-3012 (10^3 apron-inclusive block, avg / 512)
//   8-bit coarse / conservative     F/Scene/Scene.cpp:3161-3174 (ATLAS_COMPRESSION == 1 variant), 1e-9 clamp :3154-3156
//   mip + conservative rule         gvdb-voxel-src/source/gvdb_library/src/gvdb_volume_gvdb.cpp:2703-2885
//   camera U,V,W + view/proj        F/Scene/Camera/Camera.cpp:150-189
// (F/ = Source/Falcor/).  Layout differences from the reference (brick pool instead of a 3-D atlas texture, explicit
// node.pos / node.link instead of 16-bit packed pos/value) are described in include/vrestir.h and DESIGN.md.
#include "../../include/vrestir.h"
#include "vr_host.h"
#include "vr_procedural.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace vr {

// ------------------------------------------------------------------------------------------------ helpers
template <class F> static void parallelFor(int n, F f) {
    int nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    if (n < 4) nt = 1;
    std::atomic<int> next{0};
    auto body = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; f(i); } };
    if (nt == 1) { body(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nt; i++) th.emplace_back(body);
    for (auto& t : th) t.join();
}

struct Dense {
    int nx = 0, ny = 0, nz = 0, ch = 1;
    std::vector<float> v;   // [ch][z][y][x]
    size_t n() const { return (size_t)nx * ny * nz; }
    float at(int x, int y, int z, int c = 0) const {
        if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return 0.f;
        return v[(size_t)c * n() + ((size_t)z * ny + y) * nx + x];
    }
};


// ------------------------------------------------------------------------------------------------ mip chain
// conservative mip 0: zero voxels take the mean of their positive 27-neighbourhood (GV/.../gvdb_volume_gvdb.cpp:2753-2801)
static Dense makeConservative0(const Dense& src) {
    Dense d = src;
    parallelFor(src.nz, [&](int z) {
        for (int y = 0; y < src.ny; y++)
            for (int x = 0; x < src.nx; x++) {
                float org = src.at(x, y, z);
                if (org != 0.f) continue;
                float avg = 0.f;
                for (int ii = -1; ii <= 1; ii++)
                    for (int jj = -1; jj <= 1; jj++)
                        for (int kk = -1; kk <= 1; kk++) { float t = src.at(x + ii, y + jj, z + kk); avg += t > 0.f ? t : 0.f; }
                avg /= 27.f;
                if (avg > 0.f) d.v[((size_t)z * src.ny + y) * src.nx + x] = avg;
            }
    });
    return d;
}
// mip k from mip k-1 of the same type: 2x box, 3-tap polyphase on odd axes (GV/.../gvdb_volume_gvdb.cpp:2803-2862)
static Dense downsample(const Dense& prev) {
    Dense d; d.nx = std::max(1, prev.nx / 2); d.ny = std::max(1, prev.ny / 2); d.nz = std::max(1, prev.nz / 2); d.ch = 1;
    d.v.assign(d.n(), 0.f);
    const int ni = prev.nx % 2 == 0 ? 2 : 3, nj = prev.ny % 2 == 0 ? 2 : 3, nk = prev.nz % 2 == 0 ? 2 : 3;
    auto weights = [](int n, int cur, int i, float w[3]) {
        if (n == 2) { w[0] = w[1] = w[2] = 0.5f; return; }
        float den = (float)(2 * cur + 1);
        w[0] = (float)(cur - i) / den; w[1] = (float)cur / den; w[2] = (float)(1 + i) / den;
    };
    parallelFor(d.nz, [&](int k) {
        float wi[3], wj[3], wk[3];
        weights(nk, d.nz, k, wk);
        for (int j = 0; j < d.ny; j++) {
            weights(nj, d.ny, j, wj);
            for (int i = 0; i < d.nx; i++) {
                weights(ni, d.nx, i, wi);
                float res = 0.f;
                for (int ii = 0; ii < ni; ii++)
                    for (int jj = 0; jj < nj; jj++)
                        for (int kk = 0; kk < nk; kk++) res += wi[ii] * wj[jj] * wk[kk] * prev.at(2 * i + ii, 2 * j + jj, 2 * k + kk);
                if (res > 0.f) d.v[((size_t)k * d.ny + j) * d.nx + i] = res;
            }
        }
    });
    return d;
}

// ------------------------------------------------------------------------------------------------ tree + brick pool
struct BuiltSlot {
    std::vector<vrestir_node> nodes[3];
    std::vector<uint32_t> child[3];
    std::vector<uint8_t> atlas;
    bool used = false;
};

// Tree over the active bricks (GVDB 5-4-3 layout): level-1 nodes per 128^3 cell that has an active brick (z-major), level-0
// nodes (bricks) in (node1, z, y, x) order, a level-2 root when more than one level-1 cell exists.  Shared by the host slot
// builder below and by the device-resident path (vrestir_set_volume_from_chain), which feeds it a GPU-computed activity map.
void buildTopology(std::vector<uint8_t>& active, int nx, int ny, int nz, Topology& out) {
    const int BX = (nx + 7) / 8, BY = (ny + 7) / 8, BZ = (nz + 7) / 8;
    // ensure at least one brick so the tree is well formed
    bool anyActive = false; for (auto a : active) anyActive |= a != 0;
    if (!anyActive) active[0] = 1;
    const int N1X = (nx + 127) / 128, N1Y = (ny + 127) / 128, N1Z = (nz + 127) / 128;
    const bool three = (N1X * N1Y * N1Z) > 1;
    out.topLev = three ? 2 : 1;
    std::vector<int32_t> n1id((size_t)N1X * N1Y * N1Z, -1);
    for (int bz = 0; bz < BZ; bz++) for (int by = 0; by < BY; by++) for (int bx = 0; bx < BX; bx++)
        if (active[((size_t)bz * BY + by) * BX + bx]) n1id[((size_t)(bz / 16) * N1Y + by / 16) * N1X + bx / 16] = 0;
    uint32_t n1count = 0;
    for (auto& v : n1id) if (v == 0) v = (int32_t)n1count++;
    out.n1count = (int)n1count;
    out.nodes[1].resize(n1count);
    out.child[1].assign((size_t)n1count * 4096, 0xFFFFFFFFu);
    uint32_t brickCount = 0;
    for (int z1 = 0; z1 < N1Z; z1++) for (int y1 = 0; y1 < N1Y; y1++) for (int x1 = 0; x1 < N1X; x1++) {
        int32_t id = n1id[((size_t)z1 * N1Y + y1) * N1X + x1];
        if (id < 0) continue;
        vrestir_node& n = out.nodes[1][id];
        n.pos[0] = x1 * 128; n.pos[1] = y1 * 128; n.pos[2] = z1 * 128; n.link = (uint32_t)id;
        n.bounds[0] = n.bounds[1] = n.bounds[2] = n.bounds[3] = 0.f;
        for (int cz = 0; cz < 16; cz++) for (int cy = 0; cy < 16; cy++) for (int cx = 0; cx < 16; cx++) {
            int bx = x1 * 16 + cx, by = y1 * 16 + cy, bz = z1 * 16 + cz;
            if (bx >= BX || by >= BY || bz >= BZ || !active[((size_t)bz * BY + by) * BX + bx]) continue;
            out.child[1][(size_t)id * 4096 + (((cz << 4) + cy) << 4) + cx] = brickCount++;
        }
    }
    out.brickCount = brickCount;
    out.nodes[0].resize(brickCount);
    for (uint32_t id = 0; id < n1count; id++)
        for (int b = 0; b < 4096; b++) {
            uint32_t c = out.child[1][(size_t)id * 4096 + b];
            if (c == 0xFFFFFFFFu) continue;
            const vrestir_node& n1 = out.nodes[1][id];
            vrestir_node& n = out.nodes[0][c];
            n.pos[0] = n1.pos[0] + (b & 15) * 8; n.pos[1] = n1.pos[1] + ((b >> 4) & 15) * 8; n.pos[2] = n1.pos[2] + (b >> 8) * 8; n.link = c;
            n.bounds[0] = n.bounds[1] = n.bounds[2] = n.bounds[3] = 0.f;
        }
    if (three) {
        out.nodes[2].resize(1);
        vrestir_node& r = out.nodes[2][0];
        r.pos[0] = r.pos[1] = r.pos[2] = 0; r.link = 0; r.bounds[0] = r.bounds[1] = r.bounds[2] = r.bounds[3] = 0.f;
        out.child[2].assign(32768, 0xFFFFFFFFu);
        for (int z1 = 0; z1 < N1Z; z1++) for (int y1 = 0; y1 < N1Y; y1++) for (int x1 = 0; x1 < N1X; x1++) {
            int32_t id = n1id[((size_t)z1 * N1Y + y1) * N1X + x1];
            if (id >= 0) out.child[2][(((z1 << 5) + y1) << 5) + x1] = (uint32_t)id;
        }
    } else if (n1count == 0) {
        out.nodes[1].resize(1);
    }
}

static void buildSlot(const Dense& src, int format, bool conservative, BuiltSlot& out, vrestir_grid_slot& g) {
    const int BX = (src.nx + 7) / 8, BY = (src.ny + 7) / 8, BZ = (src.nz + 7) / 8;
    const int ch = src.ch;
    float maxv = 0.f;
    for (size_t i = 0; i < src.n(); i++) maxv = std::max(maxv, std::fabs(src.v[i]));   // channel 0 (density / temperature / vx)
    if (ch > 1) for (size_t i = src.n(); i < src.v.size(); i++) maxv = std::max(maxv, std::fabs(src.v[i]));
    if (maxv <= 0.f) maxv = 1.f;
    // active bricks: any non-zero in the 10^3 apron-inclusive block (so every trilinear footprint lies in a brick)
    std::vector<uint8_t> active((size_t)BX * BY * BZ, 0);
    parallelFor(BZ, [&](int bz) {
        for (int by = 0; by < BY; by++)
            for (int bx = 0; bx < BX; bx++) {
                bool any = false;
                for (int c = 0; c < ch && !any; c++)
                    for (int z = bz * 8 - 1; z <= bz * 8 + 8 && !any; z++)
                        for (int y = by * 8 - 1; y <= by * 8 + 8 && !any; y++)
                            for (int x = bx * 8 - 1; x <= bx * 8 + 8; x++) if (src.at(x, y, z, c) != 0.f) { any = true; break; }
                active[((size_t)bz * BY + by) * BX + bx] = any;
            }
    });
    vr::Topology topo;
    vr::buildTopology(active, src.nx, src.ny, src.nz, topo);
    g = vrestir_grid_slot{};
    g.valid = 1; g.top_lev = topo.topLev;
    g.dim[0] = 3; g.dim[1] = 4; g.dim[2] = 5; g.res[0] = 8; g.res[1] = 16; g.res[2] = 32;
    g.vdel[0] = 1.f; g.vdel[1] = 8.f; g.vdel[2] = 128.f; g.noderange[0] = 8; g.noderange[1] = 128; g.noderange[2] = 4096;
    const uint32_t brickCount = topo.brickCount;
    for (int l = 0; l < 3; l++) { out.nodes[l] = std::move(topo.nodes[l]); out.child[l] = std::move(topo.child[l]); }
    const size_t bpv = format == VRESTIR_ATLAS_UNORM8 ? 1 : 4;
    out.atlas.assign((size_t)brickCount * ch * VRESTIR_BRICK_VOXELS * bpv, 0);
    // fill bricks (positions and links of the level-0 nodes come from the topology)
    parallelFor((int)brickCount, [&](int bi) {
        vrestir_node& n = out.nodes[0][bi];
        for (int c = 0; c < ch; c++) {
            size_t base = ((size_t)bi * ch + c) * VRESTIR_BRICK_VOXELS;
            for (int z = -1; z <= 8; z++) for (int y = -1; y <= 8; y++) for (int x = -1; x <= 8; x++) {
                float v = src.at(n.pos[0] + x, n.pos[1] + y, n.pos[2] + z, c);
                if (v / maxv < 1e-9f && v >= 0.f) v = 0.f;                            // F/Scene/Scene.cpp:3154-3156
                size_t idx = base + (size_t)((z + 1) * 10 + (y + 1)) * 10 + (x + 1);
                if (format == VRESTIR_ATLAS_UNORM8) {
                    int q = (int)std::lround(255.0 * (double)(v / maxv));
                    q = std::max(0, std::min(255, q));
                    if (q == 0 && v > 0.f && conservative) q = 1;                     // F/Scene/Scene.cpp:3171
                    out.atlas[idx] = (uint8_t)q;
                } else {
                    memcpy(&out.atlas[idx * 4], &v, 4);
                }
            }
        }
        // (min, max, avg) over the stored 10^3 block, x outermost like F/Scene/Scene.cpp:2989-3010; avg = sum / 512
        float mn = 3.402823466e+38f, mx = 0.f, sum = 0.f;
        size_t base = (size_t)bi * ch * VRESTIR_BRICK_VOXELS;
        for (int i = -1; i <= 8; i++) for (int j = -1; j <= 8; j++) for (int k = -1; k <= 8; k++) {
            size_t idx = base + (size_t)((k + 1) * 10 + (j + 1)) * 10 + (i + 1);
            float d;
            if (format == VRESTIR_ATLAS_UNORM8) d = (float)out.atlas[idx] * 0.003921568859368563f * maxv;
            else memcpy(&d, &out.atlas[idx * 4], 4);
            mn = std::min(mn, d); mx = std::max(mx, d); sum += d;
        }
        n.bounds[0] = mn; n.bounds[1] = mx; n.bounds[2] = sum / 512.f; n.bounds[3] = 0.f;
    });
    for (int l = 0; l < 3; l++) {
        g.node_count[l] = (uint32_t)out.nodes[l].size(); g.nodes[l] = out.nodes[l].empty() ? nullptr : out.nodes[l].data();
        g.childlist[l] = out.child[l].empty() ? nullptr : out.child[l].data(); g.childlist_count[l] = out.child[l].size();
    }
    g.bmin[0] = g.bmin[1] = g.bmin[2] = 0.f;
    g.bmax[0] = (float)src.nx; g.bmax[1] = (float)src.ny; g.bmax[2] = (float)src.nz;
    g.max_value = maxv;
    g.compress_scale = format == VRESTIR_ATLAS_UNORM8 ? maxv : 1.f;
    g.atlas_format = format; g.atlas_channels = ch; g.brick_count = brickCount; g.atlas = out.atlas.data();
    out.used = true;
}

static void mat4Identity(double* m) { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
static void mat4Mul(const double* a, const double* b, double* o) {   // o = a * b (row-vector convention: apply a, then b)
    double t[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += a[i * 4 + k] * b[k * 4 + j]; t[i * 4 + j] = s; }
    memcpy(o, t, sizeof(t));
}
static void toF(const double* m, float* o) { for (int i = 0; i < 16; i++) o[i] = (float)m[i]; }

}  // namespace vr

using namespace vr;

struct vrestir_scene {
    vrestir_grid_desc desc{};
    BuiltSlot slots[VRESTIR_MAX_SLOTS];
    std::vector<float> lut;
    vrestir_scene_params params{};
};

static void setTransforms(vrestir_scene& s, int slot, const int dim0[3], const int dimk[3]) {
    const vrestir_scene_params& p = s.params;
    double X[16], Xi[16], E[16], Ei[16], M[16];
    mat4Identity(X); mat4Identity(Xi); mat4Identity(E); mat4Identity(Ei);
    for (int a = 0; a < 3; a++) {
        double sc = (double)p.voxel_size * (double)dim0[a] / (double)dimk[a];   // per-axis prescale, gvdb_volume_gvdb.cpp:2731
        double org = -0.5 * (double)dim0[a] * (double)p.voxel_size;            // volume centred on the model origin
        X[a * 5] = sc; X[12 + a] = org;
        Xi[a * 5] = 1.0 / sc; Xi[12 + a] = -org / sc;
        E[a * 5] = p.world_scaling; E[12 + a] = p.world_translation[a];
        Ei[a * 5] = 1.0 / p.world_scaling; Ei[12 + a] = -p.world_translation[a] / p.world_scaling;
    }
    vrestir_grid_slot& g = s.desc.slots[slot];
    toF(X, g.xform); toF(Xi, g.invxform);
    mat4Mul(X, E, M); toF(M, g.medium_to_world);
    mat4Mul(Ei, Xi, M); toF(M, g.world_to_medium);
    if (slot == 0) {
        toF(E, s.desc.volume.externalModelToWorld); toF(Ei, s.desc.volume.externalWorldToModel);
    }
}

// VolumeDesc (F/Scene/Scene.cpp:3246-3298)
static void fillVolumeDesc(vrestir_scene& sc, int builtMips, bool temperature, bool velocity) {
    vrestir_scene* s = &sc;
    const vrestir_scene_params* p = &sc.params;
    vrestir_volume_desc& v = s->desc.volume;
    for (int i = 0; i < 3; i++) { v.sigma_a[i] = p->sigma_a[i]; v.sigma_s[i] = p->sigma_s[i]; }
    v.sigma_t = p->sigma_s[0] + p->sigma_a[0];
    v.PhaseFunctionConstantG = p->g;
    v.densityScaleFactor = p->density_scale;
    v.densityScaleFactorByScaling = p->density_scale / p->world_scaling;
    const float* X = s->desc.slots[0].xform;
    auto len3 = [](const float* r) { return std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); };
    v.tStep = (len3(X) + len3(X + 4) + len3(X + 8)) / 3.f;
    v.hasEmission = temperature ? 1 : 0; v.hasVelocity = velocity ? 1 : 0; v.hasAnimation = 0; v.lastFrameHasEmission = 0;
    v.LeScale = p->LeScale; v.temperatureCutOff = p->temperatureCutOff; v.temperatureScale = p->temperatureScale;
    v.velocityScale = 1.f; v.numMips = builtMips; v.usePrevGridForReproj = 0;
    v.volumeWorldScaling = p->world_scaling;
    v.superVoxelWorldSpaceDiagonalLength = 8.f * std::sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2] + X[4] * X[4] + X[5] * X[5] + X[6] * X[6] + X[8] * X[8] + X[9] * X[9] + X[10] * X[10]);
    if (temperature) { s->lut.resize(512); vrestir_make_blackbody_lut(s->lut.data()); s->desc.blackbody_lut = s->lut.data(); }
}

static int buildScene(const vrestir_scene_params* p, Dense&& density, Dense* temperature, Dense* velocity, vrestir_scene** out) {
    auto* s = new vrestir_scene();
    s->params = *p;
    const int dim0[3] = {density.nx, density.ny, density.nz};
    const int numMips = std::max(1, std::min(VRESTIR_NUM_MAX_MIPS, p->num_mips));
    // normal chain
    {
        Dense cur = std::move(density);
        Dense cons = makeConservative0(cur);
        for (int m = 0; m < numMips; m++) {
            const int dk[3] = {cur.nx, cur.ny, cur.nz};
            buildSlot(cur, m == 0 ? VRESTIR_ATLAS_F32 : VRESTIR_ATLAS_UNORM8, false, s->slots[m], s->desc.slots[m]);
            setTransforms(*s, m, dim0, dk);
            buildSlot(cons, VRESTIR_ATLAS_UNORM8, true, s->slots[VRESTIR_NUM_MAX_MIPS + m], s->desc.slots[VRESTIR_NUM_MAX_MIPS + m]);
            setTransforms(*s, VRESTIR_NUM_MAX_MIPS + m, dim0, dk);
            if (m + 1 < numMips) {
                if (cur.nx < 2 || cur.ny < 2 || cur.nz < 2) { s->params.num_mips = m + 1; break; }
                cur = downsample(cur); cons = downsample(cons);
            }
        }
    }
    int builtMips = 0; for (int m = 0; m < VRESTIR_NUM_MAX_MIPS; m++) if (s->desc.slots[m].valid) builtMips = m + 1;
    if (temperature) {
        buildSlot(*temperature, VRESTIR_ATLAS_F32, false, s->slots[VRESTIR_TEMPERATURE_GRID_ID], s->desc.slots[VRESTIR_TEMPERATURE_GRID_ID]);
        setTransforms(*s, VRESTIR_TEMPERATURE_GRID_ID, dim0, dim0);
    }
    if (velocity) {
        buildSlot(*velocity, VRESTIR_ATLAS_F32, false, s->slots[VRESTIR_VELOCITY_GRID_ID], s->desc.slots[VRESTIR_VELOCITY_GRID_ID]);
        setTransforms(*s, VRESTIR_VELOCITY_GRID_ID, dim0, dim0);
    }
    fillVolumeDesc(*s, builtMips, temperature != nullptr, velocity != nullptr);
    *out = s;
    return VRESTIR_OK;
}

extern "C" {

int vrestir_scene_create(const vrestir_scene_params* p, vrestir_scene** out) try {
    if (!p || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->dim[0] < 8 || p->dim[1] < 8 || p->dim[2] < 8 || p->dim[0] > 4096 || p->dim[1] > 4096 || p->dim[2] > 4096)
        return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "grid dimensions must be in [8, 4096]");
    if ((double)p->dim[0] * p->dim[1] * p->dim[2] > 1.2e9) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "dense procedural build limited to 1.2e9 voxels");
    Dense d; d.nx = p->dim[0]; d.ny = p->dim[1]; d.nz = p->dim[2]; d.v.assign(d.n(), 0.f);
    Dense T, V;
    const bool wt = p->with_temperature && p->kind == 2, wv = p->with_velocity && p->kind == 2;
    if (wt) { T = d; }
    if (wv) { V.nx = d.nx; V.ny = d.ny; V.nz = d.nz; V.ch = 3; V.v.assign(d.n() * 3, 0.f); }
    parallelFor(d.nz, [&](int z) {
        for (int y = 0; y < d.ny; y++)
            for (int x = 0; x < d.nx; x++) {
                float u = ((float)x + 0.5f) / (float)d.nx, v = ((float)y + 0.5f) / (float)d.ny, w = ((float)z + 0.5f) / (float)d.nz;
                float temp = 0.f, vel[3] = {0, 0, 0};
                float dens = shapeDensity(*p, u, v, w, wt ? &temp : nullptr, wv ? vel : nullptr);
                size_t i = ((size_t)z * d.ny + y) * d.nx + x;
                d.v[i] = dens;
                if (wt) T.v[i] = temp;
                if (wv) { V.v[i] = vel[0]; V.v[d.n() + i] = vel[1]; V.v[2 * d.n() + i] = vel[2]; }
            }
    });
    return buildScene(p, std::move(d), wt ? &T : nullptr, wv ? &V : nullptr, out);
} catch (...) { return vr::caughtException(); }

// A scene description without voxels: dimensions, formats, transforms and the volume description of every density level —
// the template vrestir_set_volume_from_chain needs when the grid itself only ever exists on the device.
int vrestir_scene_create_template(const vrestir_scene_params* p, vrestir_scene** out) try {
    if (!p || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->dim[0] < 8 || p->dim[1] < 8 || p->dim[2] < 8 || p->dim[0] > 4096 || p->dim[1] > 4096 || p->dim[2] > 4096)
        return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "grid dimensions must be in [8, 4096]");
    auto* s = new vrestir_scene();
    s->params = *p;
    const int dim0[3] = {p->dim[0], p->dim[1], p->dim[2]};
    const int numMips = std::max(1, std::min(VRESTIR_NUM_MAX_MIPS, p->num_mips));
    int d[3] = {dim0[0], dim0[1], dim0[2]};
    int built = 0;
    for (int m = 0; m < numMips; m++) {
        for (int c = 0; c < 2; c++) {
            const int slot = m + c * VRESTIR_NUM_MAX_MIPS;
            vrestir_grid_slot& g = s->desc.slots[slot];
            g = vrestir_grid_slot{};
            g.valid = 1; g.top_lev = ((d[0] + 127) / 128) * ((d[1] + 127) / 128) * ((d[2] + 127) / 128) > 1 ? 2 : 1;
            g.dim[0] = 3; g.dim[1] = 4; g.dim[2] = 5; g.res[0] = 8; g.res[1] = 16; g.res[2] = 32;
            g.vdel[0] = 1.f; g.vdel[1] = 8.f; g.vdel[2] = 128.f; g.noderange[0] = 8; g.noderange[1] = 128; g.noderange[2] = 4096;
            g.bmax[0] = (float)d[0]; g.bmax[1] = (float)d[1]; g.bmax[2] = (float)d[2];
            g.max_value = 1.f; g.compress_scale = 1.f;
            g.atlas_format = (m == 0 && c == 0) ? VRESTIR_ATLAS_F32 : VRESTIR_ATLAS_UNORM8; g.atlas_channels = 1;
            setTransforms(*s, slot, dim0, d);
        }
        built = m + 1;
        if (d[0] < 2 || d[1] < 2 || d[2] < 2) break;
        for (int a = 0; a < 3; a++) d[a] = std::max(1, d[a] / 2);
    }
    s->params.num_mips = built;
    fillVolumeDesc(*s, built, false, false);
    *out = s;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_scene_create_from_dense(const vrestir_scene_params* p, const float* density, const float* temperature, const float* velocity_xyz, vrestir_scene** out) try {
    if (!p || !density || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->dim[0] < 1 || p->dim[1] < 1 || p->dim[2] < 1 || p->dim[0] > 4096 || p->dim[1] > 4096 || p->dim[2] > 4096)
        return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "grid dimensions must be in [1, 4096]");
    Dense d; d.nx = p->dim[0]; d.ny = p->dim[1]; d.nz = p->dim[2]; d.v.assign(density, density + d.n());
    Dense T, V;
    if (temperature) { T.nx = d.nx; T.ny = d.ny; T.nz = d.nz; T.v.assign(temperature, temperature + d.n()); }
    if (velocity_xyz) {
        V.nx = d.nx; V.ny = d.ny; V.nz = d.nz; V.ch = 3; V.v.resize(d.n() * 3);
        for (size_t i = 0; i < d.n(); i++) for (int c = 0; c < 3; c++) V.v[(size_t)c * d.n() + i] = velocity_xyz[i * 3 + c];
    }
    return buildScene(p, std::move(d), temperature ? &T : nullptr, velocity_xyz ? &V : nullptr, out);
} catch (...) { return vr::caughtException(); }

}  // extern "C"
namespace vr {
// host copy of a device-resident volume (vrestir_download_volume in vr_pass.cu): the caller fills the vectors of each slot
vrestir_scene* newHostScene(const vrestir_volume_desc& vol) { auto* s = new vrestir_scene(); s->desc.volume = vol; return s; }
HostSlotVectors hostSlotVectors(vrestir_scene* s, int slot) {
    BuiltSlot& b = s->slots[slot];
    b.used = true;
    return HostSlotVectors{{&b.nodes[0], &b.nodes[1], &b.nodes[2]}, {&b.child[0], &b.child[1], &b.child[2]}, &b.atlas, &s->desc.slots[slot]};
}
void attachBlackbodyLut(vrestir_scene* s) { s->lut.resize(512); vrestir_make_blackbody_lut(s->lut.data()); s->desc.blackbody_lut = s->lut.data(); }
}  // namespace vr
extern "C" {

int vrestir_scene_destroy(vrestir_scene* s) { delete s; return VRESTIR_OK; }
const vrestir_grid_desc* vrestir_scene_grid(const vrestir_scene* s) { return s ? &s->desc : nullptr; }

int vrestir_scene_dense_mip(const vrestir_scene* s, int mip, int conservative, float* out, int32_t out_dim[3]) try {
    if (!s || mip < 0 || mip >= VRESTIR_NUM_MAX_MIPS) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad mip");
    const int slot = mip + (conservative ? VRESTIR_NUM_MAX_MIPS : 0);
    const vrestir_grid_slot& g = s->desc.slots[slot];
    if (!g.valid) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "slot not built");
    const int nx = (int)g.bmax[0], ny = (int)g.bmax[1], nz = (int)g.bmax[2];
    if (out_dim) { out_dim[0] = nx; out_dim[1] = ny; out_dim[2] = nz; }
    if (!out) return VRESTIR_OK;
    memset(out, 0, (size_t)nx * ny * nz * 4);
    for (uint32_t b = 0; b < g.brick_count; b++) {
        const vrestir_node& n = g.nodes[0][b];
        for (int z = 0; z < 8; z++) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
            int gx = n.pos[0] + x, gy = n.pos[1] + y, gz = n.pos[2] + z;
            if (gx >= nx || gy >= ny || gz >= nz) continue;
            size_t idx = (size_t)b * g.atlas_channels * VRESTIR_BRICK_VOXELS + (size_t)((z + 1) * 10 + (y + 1)) * 10 + (x + 1);
            float v;
            if (g.atlas_format == VRESTIR_ATLAS_UNORM8) v = (float)((const uint8_t*)g.atlas)[idx] * 0.003921568859368563f * g.compress_scale;
            else v = ((const float*)g.atlas)[idx];
            out[((size_t)gz * ny + gy) * nx + gx] = v;
        }
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_scene_stats(const vrestir_scene* s, int slot, uint32_t* bricks, uint64_t* atlas_bytes) try {
    if (!s || slot < 0 || slot >= VRESTIR_MAX_SLOTS) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad slot");
    const vrestir_grid_slot& g = s->desc.slots[slot];
    if (bricks) *bricks = g.valid ? g.brick_count : 0;
    if (atlas_bytes) *atlas_bytes = g.valid ? (uint64_t)s->slots[slot].atlas.size() : 0;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// F/Scene/Camera/Camera.cpp:150-189 (focalDistance = 10000 as in CameraData.slang:60; view = lookAt RH, proj = perspective RH)
int vrestir_camera_look_at(const float pos[3], const float target[3], const float up[3], float fovY, float aspect, float nearZ, float farZ, vrestir_camera* out) try {
    if (!pos || !target || !up || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    auto sub = [](const float* a, const float* b, float* o) { for (int i = 0; i < 3; i++) o[i] = a[i] - b[i]; };
    auto nrm = [](float* a) { float l = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); for (int i = 0; i < 3; i++) a[i] /= l; };
    auto crs = [](const float* a, const float* b, float* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    auto dt = [](const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const float focalDistance = 10000.f;
    float f[3], s[3], u[3];
    sub(target, pos, f); nrm(f);
    crs(f, up, s); nrm(s);
    crs(s, f, u);
    const float ulen = focalDistance * std::tan(fovY * 0.5f) * aspect, vlen = focalDistance * std::tan(fovY * 0.5f);
    float vv[3]; crs(s, f, vv); nrm(vv);   // cameraV = normalize(cross(U, W))
    for (int i = 0; i < 3; i++) { out->posW[i] = pos[i]; out->cameraW[i] = f[i] * focalDistance; out->cameraU[i] = s[i] * ulen; out->cameraV[i] = vv[i] * vlen; }
    // view (row-vector convention): columns are s, u, -f
    float* V = out->viewMat;
    V[0] = s[0]; V[1] = u[0]; V[2] = -f[0]; V[3] = 0; V[4] = s[1]; V[5] = u[1]; V[6] = -f[1]; V[7] = 0; V[8] = s[2]; V[9] = u[2]; V[10] = -f[2]; V[11] = 0;
    V[12] = -dt(s, pos); V[13] = -dt(u, pos); V[14] = dt(f, pos); V[15] = 1;
    float* P = out->projMat; memset(P, 0, 64);
    const float th = std::tan(fovY * 0.5f);
    P[0] = 1.f / (aspect * th); P[5] = 1.f / th; P[10] = farZ / (nearZ - farZ); P[11] = -1.f; P[14] = -(farZ * nearZ) / (farZ - nearZ);
    out->nearZ = nearZ; out->farZ = farZ;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_make_sky_envmap(int width, int height, uint32_t seed, float* out) try {
    if (width < 2 || height < 2 || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad env map size");
    const float sunDir[3] = {0.45f, 0.62f, -0.64f};
    parallelFor(height, [&](int y) {
        for (int x = 0; x < width; x++) {
            float u = ((float)x + 0.5f) / (float)width, v = ((float)y + 0.5f) / (float)height;
            float phi = (u - 0.5f) * 6.28318530718f, th = v * 3.14159265359f;
            float d[3] = {std::sin(phi) * std::sin(th), std::cos(th), -std::cos(phi) * std::sin(th)};
            float up = d[1];
            float horizon = std::exp(-std::fabs(up) * 4.f);
            float sky[3] = {0.25f + 0.55f * horizon, 0.42f + 0.45f * horizon, 0.85f + 0.1f * horizon};
            float gnd[3] = {0.18f, 0.16f, 0.13f};
            float t = std::min(1.f, std::max(0.f, up * 8.f + 0.5f));
            float cs = d[0] * sunDir[0] + d[1] * sunDir[1] + d[2] * sunDir[2];
            float sun = cs > 0.9995f ? 900.f : 0.f;
            float glow = std::pow(std::max(0.f, cs), 64.f) * 6.f + std::pow(std::max(0.f, cs), 8.f) * 0.6f;
            float cl = 0.85f + 0.3f * fbm(u * 12.f, v * 6.f, 0.5f, seed, 4);
            float* o = &out[((size_t)y * width + x) * 4];
            for (int c = 0; c < 3; c++) o[c] = (gnd[c] * (1.f - t) + sky[c] * t) * cl + (c == 2 ? 0.8f : 1.f) * (sun + glow);
            o[3] = 1.f;
        }
    });
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_make_emissive_shell(int count, uint32_t seed, const float center[3], float radius, vrestir_emissive_triangle* out) try {
    if (count < 0 || !out || !center) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad arguments");
    for (int i = 0; i < count; i++) {
        auto rnd = [&](int k) { return lattice(i, k, 17, seed); };
        float z = 1.f - 2.f * rnd(0), ph = 6.28318530718f * rnd(1), r = std::sqrt(std::max(0.f, 1.f - z * z));
        float n[3] = {r * std::cos(ph), z, r * std::sin(ph)};
        float c[3] = {center[0] + n[0] * radius, center[1] + n[1] * radius, center[2] + n[2] * radius};
        // tangent frame
        float a[3] = {std::fabs(n[0]) > 0.9f ? 0.f : 1.f, std::fabs(n[0]) > 0.9f ? 1.f : 0.f, 0.f};
        float t1[3] = {n[1] * a[2] - n[2] * a[1], n[2] * a[0] - n[0] * a[2], n[0] * a[1] - n[1] * a[0]};
        float l = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]); for (auto& v : t1) v /= l;
        float t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
        float size = radius * 0.02f * std::exp(0.6f * (rnd(2) + rnd(3) + rnd(4) - 1.5f) * 2.f);   // ~log-normal
        float rot = 6.28318530718f * rnd(5);
        vrestir_emissive_triangle& T = out[i];
        for (int v = 0; v < 3; v++) {
            float ang = rot + 2.09439510239f * (float)v;
            for (int k = 0; k < 3; k++) T.posW[v][k] = c[k] + size * (std::cos(ang) * t1[k] + std::sin(ang) * t2[k]);
        }
        // inward-facing normal (lights shine on the volume)
        float e1[3], e2[3];
        for (int k = 0; k < 3; k++) { e1[k] = T.posW[1][k] - T.posW[0][k]; e2[k] = T.posW[2][k] - T.posW[0][k]; }
        float cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        float cl = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        T.area = 0.5f * cl;
        float sgn = (cr[0] * n[0] + cr[1] * n[1] + cr[2] * n[2]) > 0.f ? -1.f : 1.f;
        if (sgn < 0.f) { for (int k = 0; k < 3; k++) std::swap(T.posW[1][k], T.posW[2][k]); }
        for (int k = 0; k < 3; k++) T.normal[k] = -n[k];
        float power = 40.f * std::exp(0.8f * (rnd(6) + rnd(7) + rnd(8) - 1.5f) * 2.f);
        float hue = rnd(9);
        T.Le[0] = power * (0.6f + 0.4f * hue); T.Le[1] = power * (0.6f + 0.4f * (1.f - std::fabs(2.f * hue - 1.f))); T.Le[2] = power * (1.f - 0.4f * hue);
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Planck-law spectrum -> linear sRGB via Gaussian fits of the CIE 1931 colour-matching functions (Wyman et al. 2013),
// normalised so that the hottest entry has max component 1; 128 entries for T = 50 K .. 6400 K (50 K steps).
int vrestir_make_blackbody_lut(float* out) try {
    if (!out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    auto g = [](double x, double mu, double s1, double s2) { double t = (x - mu) / (x < mu ? s1 : s2); return std::exp(-0.5 * t * t); };
    double rgb[128][3]; double mx = 0;
    for (int i = 0; i < 128; i++) {
        double T = 50.0 * (i + 1);
        double X = 0, Y = 0, Z = 0;
        for (double lam = 380; lam <= 780; lam += 5) {
            double l = lam * 1e-9;
            double B = 3.741771852e-16 / (std::pow(l, 5) * (std::exp(1.438776877e-2 / (l * T)) - 1.0));
            double xb = 1.056 * g(lam, 599.8, 37.9, 31.0) + 0.362 * g(lam, 442.0, 16.0, 26.7) - 0.065 * g(lam, 501.1, 20.4, 26.2);
            double yb = 0.821 * g(lam, 568.8, 46.9, 40.5) + 0.286 * g(lam, 530.9, 16.3, 31.1);
            double zb = 1.217 * g(lam, 437.0, 11.8, 36.0) + 0.681 * g(lam, 459.0, 26.0, 13.8);
            X += B * xb; Y += B * yb; Z += B * zb;
        }
        rgb[i][0] = std::max(0.0, 3.2406 * X - 1.5372 * Y - 0.4986 * Z);
        rgb[i][1] = std::max(0.0, -0.9689 * X + 1.8758 * Y + 0.0415 * Z);
        rgb[i][2] = std::max(0.0, 0.0557 * X - 0.2040 * Y + 1.0570 * Z);
        for (int c = 0; c < 3; c++) mx = std::max(mx, rgb[i][c]);
    }
    for (int i = 0; i < 128; i++) { for (int c = 0; c < 3; c++) out[i * 4 + c] = (float)(rgb[i][c] / mx); out[i * 4 + 3] = 0.f; }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ .vbx files
// GVDB .vbx reader / writer.  File layout: gvdb-voxel-src/GVDB_FILESPEC.txt as actually read by
// VolumeGVDB::LoadVBX (GV/src/gvdb_volume_gvdb.cpp:532-739) and written by SaveVBX (:1682-1831): version 1.12 (the
// reference's custom minor version carrying xform, inverse, effective voxel bounds and value range, :596-605), one grid,
// dtype 'f', 1 component, no compression, topology type 2, atlas layout, 64-byte nvdb::Node rows (GV/src/gvdb_node.h),
// 8-byte child ids grp | lev << 8 | ndx << 16 (GV/src/gvdb_allocator.h:73-76), one float atlas with a 1-voxel apron.
// File naming: <dir>/<name>_mip<k>[c].vbx, <name>_temperature.vbx, <name>_velocity_{x,y,z}.vbx (F/Scene/Scene.cpp:2806-2815).
// Loading repacks exactly like F/Scene/Scene.cpp:2932-3056 (32-byte nodes, 4-byte child ids) into the brick-pool layout;
// coarse / conservative mips are quantised to UNORM8 at load like F/Scene/Scene.cpp:3139-3240 (ATLAS_COMPRESSION == 1).
namespace {

#pragma pack(push, 1)
struct VbxNode {   // nvdb::Node, 64 bytes
    uint8_t lev, flags, priority, pad;
    int32_t pos[3];
    int32_t value[3];
    float vrange[3];
    uint64_t parent, childList, mask;
};
#pragma pack(pop)
static_assert(sizeof(VbxNode) == 64, "nvdb::Node is 64 bytes");

inline uint64_t vbxElem(unsigned grp, unsigned lev, uint64_t ndx) { return (uint64_t)grp | ((uint64_t)lev << 8) | (ndx << 16); }
const uint64_t kVbxUndef = 0xFFFFFFFFFFFFFFFFull;

struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
    template <class T> bool rd(T* v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }
    template <class T> bool wr(const T* v, size_t n = 1) { return fwrite(v, sizeof(T), n, f) == n; }
};

float slotVoxel(const vrestir_grid_slot& g, uint32_t brick, int ch, int x, int y, int z) {   // x,y,z in [-1, 8]
    const size_t idx = ((size_t)brick * g.atlas_channels + ch) * VRESTIR_BRICK_VOXELS + (size_t)((z + 1) * 10 + (y + 1)) * 10 + (x + 1);
    if (g.atlas_format == VRESTIR_ATLAS_UNORM8) return (float)((const uint8_t*)g.atlas)[idx] * 0.003921568859368563f * g.compress_scale;
    return ((const float*)g.atlas)[idx];
}

int saveSlotVbx(const vrestir_grid_slot& g, int channel, const std::string& path) {
    File fp; fp.f = fopen(path.c_str(), "wb");
    if (!fp.f) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "cannot open " + path + " for writing");
    const uint8_t major = 1, minor = 12;
    const float zero3[3] = {0, 0, 0}, one3[3] = {1, 1, 1};
    fp.wr(&major); fp.wr(&minor);
    fp.wr(zero3, 3); fp.wr(zero3, 3); fp.wr(one3, 3); fp.wr(zero3, 3);      // pretranslation, Euler angles, scale, translation
    fp.wr(g.xform, 16); fp.wr(g.invxform, 16);
    int32_t vmin[3], vmax[3];
    for (int i = 0; i < 3; i++) { vmin[i] = (int32_t)g.bmin[i]; vmax[i] = (int32_t)g.bmax[i]; }
    fp.wr(vmin, 3); fp.wr(vmax, 3);
    const float valMin = 0.f, valMax = g.max_value;
    fp.wr(&valMin); fp.wr(&valMax);
    const int32_t numGrids = 1; fp.wr(&numGrids);
    const long gridTable = ftell(fp.f);
    uint64_t gridOff = 0; fp.wr(&gridOff);
    gridOff = (uint64_t)ftell(fp.f);
    char name[256]; memset(name, 0, sizeof(name)); snprintf(name, sizeof(name), "density");
    fp.wr(name, 256);
    const uint8_t dtype = 'f', comps = 1, compress = 0, topo = 2, layout = 0;
    fp.wr(&dtype); fp.wr(&comps); fp.wr(&compress);
    fp.wr(one3, 3);
    const int32_t leafcnt = (int32_t)g.brick_count, leafdim[3] = {8, 8, 8}, apron = 1, numChan = 1;
    int32_t ac = 1; while ((int64_t)ac * ac * ac < leafcnt) ac++;
    const int32_t axiscnt[3] = {ac, ac, ac}, axisres[3] = {ac * 10, ac * 10, ac * 10};
    const uint64_t atlasSz = (uint64_t)axisres[0] * axisres[1] * axisres[2] * 4;
    const int32_t reuse = 0;
    fp.wr(&leafcnt); fp.wr(leafdim, 3); fp.wr(&apron); fp.wr(&numChan); fp.wr(&atlasSz); fp.wr(&topo); fp.wr(&reuse); fp.wr(&layout);
    fp.wr(axiscnt, 3); fp.wr(axisres, 3);
    const int32_t levels = g.top_lev + 1;
    const uint64_t root = vbxElem(0, (unsigned)g.top_lev, 0);
    fp.wr(&levels); fp.wr(&root);
    for (int n = 0; n < levels; n++) {
        const int32_t ld = g.dim[n], res = g.res[n], range[3] = {g.noderange[n], g.noderange[n], g.noderange[n]};
        const int32_t cnt0 = (int32_t)g.node_count[n], w0 = 64, cnt1 = n == 0 ? 0 : (int32_t)g.node_count[n], w1 = n == 0 ? 0 : res * res * res * 8;
        fp.wr(&ld); fp.wr(&res); fp.wr(range, 3); fp.wr(&cnt0); fp.wr(&w0); fp.wr(&cnt1); fp.wr(&w1);
    }
    // parents: one pass over the child lists
    std::vector<uint64_t> parent[3];
    for (int n = 0; n < levels; n++) parent[n].assign(g.node_count[n], kVbxUndef);
    for (int n = 1; n < levels; n++) {
        const size_t r3 = (size_t)g.res[n] * g.res[n] * g.res[n];
        for (uint32_t i = 0; i < g.node_count[n]; i++)
            for (size_t b = 0; b < r3; b++) {
                const uint32_t c = g.childlist[n][(size_t)g.nodes[n][i].link * r3 + b];
                if (c != 0xFFFFFFFFu && c < g.node_count[n - 1]) parent[n - 1][c] = vbxElem(0, (unsigned)n, i);
            }
    }
    for (int n = 0; n < levels; n++)
        for (uint32_t i = 0; i < g.node_count[n]; i++) {
            const vrestir_node& s = g.nodes[n][i];
            VbxNode o; memset(&o, 0, sizeof(o));
            o.lev = (uint8_t)n; o.flags = 1;
            for (int k = 0; k < 3; k++) { o.pos[k] = s.pos[k]; o.value[k] = -1; o.vrange[k] = s.bounds[k]; }
            if (n == 0) {
                const int32_t b = (int32_t)s.link;
                o.value[0] = (b % ac) * 10 + 1; o.value[1] = ((b / ac) % ac) * 10 + 1; o.value[2] = (b / (ac * ac)) * 10 + 1;
                o.childList = kVbxUndef;
            } else o.childList = vbxElem(1, (unsigned)n, s.link);
            o.parent = parent[n][i];
            fp.wr(&o);
        }
    for (int n = 1; n < levels; n++) {
        const size_t r3 = (size_t)g.res[n] * g.res[n] * g.res[n];
        std::vector<uint64_t> row(r3);
        // rows are indexed by the list id (= node.link); the builder uses link == node index
        for (uint32_t i = 0; i < g.node_count[n]; i++) {
            for (size_t b = 0; b < r3; b++) {
                const uint32_t c = g.childlist[n][(size_t)i * r3 + b];
                row[b] = c == 0xFFFFFFFFu ? kVbxUndef : vbxElem(0, (unsigned)(n - 1), c);
            }
            fp.wr(row.data(), r3);
        }
    }
    const int32_t chanType = 3 /* T_FLOAT */, chanStride = 4;
    fp.wr(&chanType); fp.wr(&chanStride);
    std::vector<float> slice((size_t)axisres[0] * axisres[1]);
    for (int z = 0; z < axisres[2]; z++) {
        std::fill(slice.begin(), slice.end(), 0.f);
        const int cz = z / 10, lz = z % 10 - 1;
        for (int cy = 0; cy < ac; cy++) for (int cx = 0; cx < ac; cx++) {
            const int64_t b = ((int64_t)cz * ac + cy) * ac + cx;
            if (b >= leafcnt) continue;
            for (int ly = -1; ly <= 8; ly++) for (int lx = -1; lx <= 8; lx++)
                slice[(size_t)(cy * 10 + ly + 1) * axisres[0] + (cx * 10 + lx + 1)] = slotVoxel(g, (uint32_t)b, channel, lx, ly, lz);
        }
        if (!fp.wr(slice.data(), slice.size())) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "short write to " + path);
    }
    fseek(fp.f, gridTable, SEEK_SET);
    fp.wr(&gridOff);
    return VRESTIR_OK;
}

struct LoadedVbx {
    BuiltSlot built; vrestir_grid_slot g{};
    std::vector<float> voxels;   // [brick][10*10*10] floats
    int32_t effMin[3] = {0, 0, 0}, effMax[3] = {0, 0, 0};   // mEffectiveVoxMin / mEffectiveVoxMax of the 1.12 header
};

int loadSlotVbx(const std::string& path, LoadedVbx& L) {
    File fp; fp.f = fopen(path.c_str(), "rb");
    if (!fp.f) return VRESTIR_ERR_NOT_READY;   // caller decides whether the file is optional
    auto bad = [&](const char* why) { return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, path + ": " + why); };
    // every size below is checked against the file size before it reaches an allocation: a malformed file must come back
    // as VRESTIR_ERR_INVALID_ARGUMENT, never as a bad_alloc / length_error unwinding through the C ABI
    if (fseek(fp.f, 0, SEEK_END) != 0) return bad("cannot seek");
    const int64_t fileSize = (int64_t)ftell(fp.f);
    if (fileSize < 0 || fseek(fp.f, 0, SEEK_SET) != 0) return bad("cannot seek");
    uint8_t major = 0, minor = 0;
    if (!fp.rd(&major) || !fp.rd(&minor)) return bad("truncated header");
    float xf[16], ixf[16]; int32_t vmin[3] = {0, 0, 0}, vmax[3] = {0, 0, 0}; float valMin = 0.f, valMax = 1.f;
    bool haveX = false;
    if ((major == 1 && minor >= 11) || major > 1) { float skip[12]; if (!fp.rd(skip, 12)) return bad("truncated transform"); }
    if (minor == 12) { if (!fp.rd(xf, 16) || !fp.rd(ixf, 16) || !fp.rd(vmin, 3) || !fp.rd(vmax, 3) || !fp.rd(&valMin) || !fp.rd(&valMax)) return bad("truncated 1.12 block"); haveX = true; }
    if (!haveX) return bad("only the reference's version 1.12 files (with xform / effective bounds) are supported");
    int32_t numGrids = 0; if (!fp.rd(&numGrids) || numGrids < 1) return bad("no grids");
    if ((int64_t)numGrids * 8 > fileSize) return bad("grid count exceeds the file size");
    if (major >= 2) { uint8_t masks; fp.rd(&masks); if (masks) return bad("bitmask topologies are not supported"); }
    std::vector<uint64_t> offs(numGrids); if (!fp.rd(offs.data(), offs.size())) return bad("truncated grid table");
    if (offs[0] >= (uint64_t)fileSize || fseek(fp.f, (long)offs[0], SEEK_SET) != 0) return bad("bad grid offset");
    char name[256]; uint8_t dtype, comps, compress, topo, layout; float vs[3];
    int32_t leafcnt, leafdim[3], apron, numChan, reuse, axiscnt[3], axisres[3]; uint64_t atlasSz;
    if (!fp.rd(name, 256) || !fp.rd(&dtype) || !fp.rd(&comps) || !fp.rd(&compress) || !fp.rd(vs, 3) || !fp.rd(&leafcnt) || !fp.rd(leafdim, 3) || !fp.rd(&apron) ||
        !fp.rd(&numChan) || !fp.rd(&atlasSz) || !fp.rd(&topo) || !fp.rd(&reuse) || !fp.rd(&layout) || !fp.rd(axiscnt, 3) || !fp.rd(axisres, 3)) return bad("truncated grid header");
    if (dtype != 'f' || comps != 1 || compress != 0 || topo != 2 || layout != 0) return bad("only float / 1 component / uncompressed / GVDB topology / atlas layout is supported");
    if (leafdim[0] != 8 || leafdim[1] != 8 || leafdim[2] != 8 || apron != 1) return bad("only 8^3 bricks with a 1-voxel apron are supported");
    int32_t levels = 0; uint64_t root = 0;
    if (!fp.rd(&levels) || !fp.rd(&root) || levels < 2 || levels > 3) return bad("only 2- or 3-level trees (5-4-3 configuration) are supported");
    int32_t ld[3], res[3], range[3][3], cnt0[3], w0[3], cnt1[3], w1[3];
    for (int n = 0; n < levels; n++)
        if (!fp.rd(&ld[n]) || !fp.rd(&res[n]) || !fp.rd(range[n], 3) || !fp.rd(&cnt0[n]) || !fp.rd(&w0[n]) || !fp.rd(&cnt1[n]) || !fp.rd(&w1[n])) return bad("truncated level table");
    static const int wantLd[3] = {3, 4, 5};
    for (int n = 0; n < levels; n++) if (ld[n] != wantLd[n] || res[n] != (1 << ld[n]) || w0[n] != 64) return bad("tree configuration is not <3,4,5> with 64-byte nodes");
    if (cnt0[0] != leafcnt) return bad("atlas does not hold every brick");
    if (leafcnt < 0) return bad("negative brick count");
    for (int n = 0; n < levels; n++) {
        if (cnt0[n] < 0 || cnt1[n] < 0 || w1[n] < 0) return bad("negative pool count");
        if ((int64_t)cnt0[n] * 64 > fileSize || (int64_t)cnt1[n] * (int64_t)w1[n] > fileSize) return bad("pool larger than the file");
        if (range[n][0] <= 0 || range[n][0] != range[n][1] || range[n][0] != range[n][2]) return bad("non-cubic node range");
    }
    if (cnt0[levels - 1] != 1) return bad("the top level must hold exactly one node");
    for (int a = 0; a < 3; a++) if (axisres[a] < 0 || axisres[a] > (1 << 20)) return bad("bad atlas resolution");
    if ((double)axisres[0] * (double)axisres[1] * (double)axisres[2] * 4.0 > (double)fileSize) return bad("atlas larger than the file");
    vrestir_grid_slot& g = L.g; g = vrestir_grid_slot{};
    BuiltSlot& B = L.built;
    std::vector<VbxNode> raw[3];
    for (int n = 0; n < levels; n++) { raw[n].resize(cnt0[n]); if (cnt0[n] && !fp.rd(raw[n].data(), raw[n].size())) return bad("truncated node pool"); }
    for (int n = 0; n < levels; n++) {
        const size_t r3 = (size_t)res[n] * res[n] * res[n];
        if (n == 0) { if (cnt1[0] * (int64_t)w1[0] != 0) fseek(fp.f, (long)((int64_t)cnt1[0] * w1[0]), SEEK_CUR); continue; }
        if ((size_t)w1[n] != r3 * 8) return bad("child-list width does not match res^3");
        std::vector<uint64_t> rows((size_t)cnt1[n] * r3);
        if (!rows.empty() && !fp.rd(rows.data(), rows.size())) return bad("truncated child lists");
        B.child[n].resize(rows.size());
        for (size_t i = 0; i < rows.size(); i++) {
            const uint32_t c = (uint32_t)(((rows[i] >> 32) & 0xFFFFull) << 16) | (uint32_t)((rows[i] >> 16) & 0xFFFFull);   // F/Scene/Scene.cpp:3037
            // the traversal dereferences child ids unchecked: an id must be "no child" or a node of the level below
            if (c != 0xFFFFFFFFu && c >= (uint32_t)cnt0[n - 1]) return bad("child id out of range");
            B.child[n][i] = c;
        }
    }
    int32_t chanType, chanStride;
    if (!fp.rd(&chanType) || !fp.rd(&chanStride) || chanStride != 4) return bad("atlas channel is not 4-byte float");
    std::vector<float> atlas((size_t)axisres[0] * axisres[1] * axisres[2]);
    if (!atlas.empty() && !fp.rd(atlas.data(), atlas.size())) return bad("truncated atlas");
    // repack (F/Scene/Scene.cpp:2932-3056)
    g.valid = 1; g.top_lev = 1;
    for (int n = 0; n < levels; n++) {
        g.dim[n] = ld[n]; g.res[n] = res[n]; g.noderange[n] = range[n][0]; g.vdel[n] = (float)range[n][0] / (float)res[n];
        if (cnt0[n] == 1) g.top_lev = n;
        B.nodes[n].resize(cnt0[n]);
        for (int i = 0; i < cnt0[n]; i++) {
            vrestir_node& o = B.nodes[n][i]; const VbxNode& s = raw[n][i];
            for (int k = 0; k < 3; k++) o.pos[k] = s.pos[k];
            o.bounds[0] = o.bounds[1] = o.bounds[2] = o.bounds[3] = 0.f;
            if (n == 0) o.link = (uint32_t)i;   // brick pool order = leaf order
            else {
                o.link = (uint32_t)(((s.childList >> 32) & 0xFFFFull) << 16) | (uint32_t)((s.childList >> 16) & 0xFFFFull);
                if (o.link != 0xFFFFFFFFu && o.link >= (uint32_t)cnt1[n]) return bad("child-list id out of range");
            }
        }
    }
    if (levels == 2) { g.dim[2] = 5; g.res[2] = 32; g.vdel[2] = 128.f; g.noderange[2] = 4096; }
    L.voxels.assign((size_t)leafcnt * VRESTIR_BRICK_VOXELS, 0.f);
    for (int b = 0; b < leafcnt; b++) {
        const VbxNode& s = raw[0][b];
        for (int z = -1; z <= 8; z++) for (int y = -1; y <= 8; y++) for (int x = -1; x <= 8; x++) {
            const int ax = s.value[0] + x, ay = s.value[1] + y, az = s.value[2] + z;
            if (ax < 0 || ay < 0 || az < 0 || ax >= axisres[0] || ay >= axisres[1] || az >= axisres[2]) return bad("brick lies outside the atlas");
            L.voxels[(size_t)b * VRESTIR_BRICK_VOXELS + (size_t)((z + 1) * 10 + (y + 1)) * 10 + (x + 1)] = atlas[((size_t)az * axisres[1] + ay) * axisres[0] + ax];
        }
    }
    for (int i = 0; i < 3; i++) { g.bmin[i] = (float)vmin[i]; g.bmax[i] = (float)vmax[i]; L.effMin[i] = vmin[i]; L.effMax[i] = vmax[i]; }
    memcpy(g.xform, xf, 64); memcpy(g.invxform, ixf, 64);
    g.max_value = valMax; g.brick_count = (uint32_t)leafcnt;
    return VRESTIR_OK;
}

// quantise / store the voxels of a loaded grid, brick bounds as in F/Scene/Scene.cpp:2981-3012
void finishLoadedSlot(LoadedVbx& L, int format, bool conservative, int channels, const std::vector<float>* extra1, const std::vector<float>* extra2) {
    vrestir_grid_slot& g = L.g; BuiltSlot& B = L.built;
    const float maxv = g.max_value > 0.f ? g.max_value : 1.f;
    const size_t bpv = format == VRESTIR_ATLAS_UNORM8 ? 1 : 4;
    B.atlas.assign((size_t)g.brick_count * channels * VRESTIR_BRICK_VOXELS * bpv, 0);
    const std::vector<float>* src[3] = {&L.voxels, extra1, extra2};
    for (uint32_t b = 0; b < g.brick_count; b++) {
        for (int c = 0; c < channels; c++)
            for (int i = 0; i < VRESTIR_BRICK_VOXELS; i++) {
                float v = (*src[c])[(size_t)b * VRESTIR_BRICK_VOXELS + i];
                if (v / maxv < 1e-9f && v >= 0.f) v = 0.f;
                const size_t idx = ((size_t)b * channels + c) * VRESTIR_BRICK_VOXELS + i;
                if (format == VRESTIR_ATLAS_UNORM8) {
                    int q = (int)std::lround(255.0 * (double)(v / maxv));
                    q = std::max(0, std::min(255, q));
                    if (q == 0 && v > 0.f && conservative) q = 1;
                    B.atlas[idx] = (uint8_t)q;
                } else memcpy(&B.atlas[idx * 4], &v, 4);
            }
        // F/Scene/Scene.cpp:2981-3012: (min, max, avg) of the RAW float densities of the 10^3 block (x outermost), voxels outside
        // the effective bounds count as 0, avg = sum / 512.  (The quantised values the kernels sample can exceed this max by
        // half a UNORM8 step, exactly as the reference's BC4 texels can; the residual-ratio trackers stay unbiased.)
        float mn = 3.402823466e+38f, mx = 0.f, sum = 0.f;
        vrestir_node& n = B.nodes[0][b];
        for (int i = -1; i <= 8; i++) for (int j = -1; j <= 8; j++) for (int k = -1; k <= 8; k++) {
            float d = L.voxels[(size_t)b * VRESTIR_BRICK_VOXELS + (size_t)((k + 1) * 10 + (j + 1)) * 10 + (i + 1)];
            if (n.pos[0] + i < L.effMin[0] || n.pos[0] + i > L.effMax[0] - 1 || n.pos[1] + j < L.effMin[1] || n.pos[1] + j > L.effMax[1] - 1 ||
                n.pos[2] + k < L.effMin[2] || n.pos[2] + k > L.effMax[2] - 1) d = 0.f;
            mn = std::min(mn, d); mx = std::max(mx, d); sum += d;
        }
        n.bounds[0] = mn; n.bounds[1] = mx; n.bounds[2] = sum / 512.f; n.bounds[3] = 0.f;
    }
    for (int l = 0; l < 3; l++) {
        g.node_count[l] = (uint32_t)B.nodes[l].size(); g.nodes[l] = B.nodes[l].empty() ? nullptr : B.nodes[l].data();
        g.childlist[l] = B.child[l].empty() ? nullptr : B.child[l].data(); g.childlist_count[l] = B.child[l].size();
    }
    g.compress_scale = format == VRESTIR_ATLAS_UNORM8 ? maxv : 1.f;
    g.atlas_format = format; g.atlas_channels = channels; g.atlas = B.atlas.data();
    B.used = true;
}

void externalTransforms(vrestir_scene& s, int slot) {   // VR/VolumeBase.slang:103-130 with the script's world translation / scaling
    const vrestir_scene_params& p = s.params;
    double X[16], Xi[16], E[16], Ei[16], M[16];
    vrestir_grid_slot& g = s.desc.slots[slot];
    for (int i = 0; i < 16; i++) { X[i] = g.xform[i]; Xi[i] = g.invxform[i]; }
    mat4Identity(E); mat4Identity(Ei);
    for (int a = 0; a < 3; a++) { E[a * 5] = p.world_scaling; E[12 + a] = p.world_translation[a]; Ei[a * 5] = 1.0 / p.world_scaling; Ei[12 + a] = -p.world_translation[a] / p.world_scaling; }
    mat4Mul(X, E, M); toF(M, g.medium_to_world);
    mat4Mul(Ei, Xi, M); toF(M, g.world_to_medium);
    if (slot == 0) { toF(E, s.desc.volume.externalModelToWorld); toF(Ei, s.desc.volume.externalWorldToModel); }
}

}  // namespace

extern "C" {

int vrestir_scene_save_vbx(const vrestir_scene* s, const char* dir_and_prefix) try {
    if (!s || !dir_and_prefix) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    const std::string base(dir_and_prefix);
    for (int m = 0; m < VRESTIR_NUM_MAX_MIPS; m++)
        for (int c = 0; c < 2; c++) {
            const vrestir_grid_slot& g = s->desc.slots[m + c * VRESTIR_NUM_MAX_MIPS];
            if (!g.valid) continue;
            int rc = saveSlotVbx(g, 0, base + "_mip" + std::to_string(m) + (c ? "c" : "") + ".vbx");
            if (rc) return rc;
        }
    const vrestir_grid_slot& T = s->desc.slots[VRESTIR_TEMPERATURE_GRID_ID];
    if (T.valid) { int rc = saveSlotVbx(T, 0, base + "_temperature.vbx"); if (rc) return rc; }
    const vrestir_grid_slot& V = s->desc.slots[VRESTIR_VELOCITY_GRID_ID];
    if (V.valid) for (int c = 0; c < 3; c++) { int rc = saveSlotVbx(V, c, base + "_velocity_" + "xyz"[c] + ".vbx"); if (rc) return rc; }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_scene_load_vbx(const char* dir_and_prefix, int num_mips, const vrestir_scene_params* p, vrestir_scene** out) try {
    if (!dir_and_prefix || !p || !out) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    const std::string base(dir_and_prefix);
    std::unique_ptr<vrestir_scene> s(new vrestir_scene());
    s->params = *p;
    const int numMips = std::max(1, std::min(VRESTIR_NUM_MAX_MIPS, num_mips));
    int built = 0;
    for (int m = 0; m < numMips; m++)
        for (int c = 0; c < 2; c++) {
            LoadedVbx L;
            const std::string path = base + "_mip" + std::to_string(m) + (c ? "c" : "") + ".vbx";
            int rc = loadSlotVbx(path, L);
            if (rc == VRESTIR_ERR_NOT_READY) {
                if (m == 0 && c == 0) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "cannot open " + path);
                continue;
            }
            if (rc) return rc;
            const int slot = m + c * VRESTIR_NUM_MAX_MIPS;
            finishLoadedSlot(L, (m == 0 && c == 0) ? VRESTIR_ATLAS_F32 : VRESTIR_ATLAS_UNORM8, c == 1, 1, nullptr, nullptr);
            s->slots[slot] = std::move(L.built);
            s->desc.slots[slot] = L.g;
            // pointers into the moved vectors
            vrestir_grid_slot& g = s->desc.slots[slot]; BuiltSlot& B = s->slots[slot];
            for (int l = 0; l < 3; l++) { g.nodes[l] = B.nodes[l].empty() ? nullptr : B.nodes[l].data(); g.childlist[l] = B.child[l].empty() ? nullptr : B.child[l].data(); }
            g.atlas = B.atlas.data();
            externalTransforms(*s, slot);
            if (c == 0) built = m + 1;
        }
    bool haveT = false, haveV = false;
    {
        LoadedVbx L;
        int rc = loadSlotVbx(base + "_temperature.vbx", L);
        if (rc == VRESTIR_OK) {
            finishLoadedSlot(L, VRESTIR_ATLAS_F32, false, 1, nullptr, nullptr);
            const int slot = VRESTIR_TEMPERATURE_GRID_ID;
            s->slots[slot] = std::move(L.built); s->desc.slots[slot] = L.g;
            vrestir_grid_slot& g = s->desc.slots[slot]; BuiltSlot& B = s->slots[slot];
            for (int l = 0; l < 3; l++) { g.nodes[l] = B.nodes[l].empty() ? nullptr : B.nodes[l].data(); g.childlist[l] = B.child[l].empty() ? nullptr : B.child[l].data(); }
            g.atlas = B.atlas.data();
            externalTransforms(*s, slot);
            haveT = true;
        } else if (rc != VRESTIR_ERR_NOT_READY) return rc;
    }
    {
        LoadedVbx Lx, Ly, Lz;
        int rx = loadSlotVbx(base + "_velocity_x.vbx", Lx);
        if (rx == VRESTIR_OK) {
            int ry = loadSlotVbx(base + "_velocity_y.vbx", Ly), rz = loadSlotVbx(base + "_velocity_z.vbx", Lz);
            if (ry || rz) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "velocity needs _velocity_x/_y/_z.vbx with one shared topology");
            if (Ly.g.brick_count != Lx.g.brick_count || Lz.g.brick_count != Lx.g.brick_count) return vr::setError(VRESTIR_ERR_INVALID_ARGUMENT, "velocity components have different topologies");
            Lx.g.max_value = std::max(Lx.g.max_value, std::max(Ly.g.max_value, Lz.g.max_value));
            finishLoadedSlot(Lx, VRESTIR_ATLAS_F32, false, 3, &Ly.voxels, &Lz.voxels);   // zipped to RGB like F/Scene/Scene.cpp:3180-3240
            const int slot = VRESTIR_VELOCITY_GRID_ID;
            s->slots[slot] = std::move(Lx.built); s->desc.slots[slot] = Lx.g;
            vrestir_grid_slot& g = s->desc.slots[slot]; BuiltSlot& B = s->slots[slot];
            for (int l = 0; l < 3; l++) { g.nodes[l] = B.nodes[l].empty() ? nullptr : B.nodes[l].data(); g.childlist[l] = B.child[l].empty() ? nullptr : B.child[l].data(); }
            g.atlas = B.atlas.data();
            externalTransforms(*s, slot);
            haveV = true;
        } else if (rx != VRESTIR_ERR_NOT_READY) return rx;
    }
    // VolumeDesc (F/Scene/Scene.cpp:3246-3298)
    vrestir_volume_desc& v = s->desc.volume;
    for (int i = 0; i < 3; i++) { v.sigma_a[i] = p->sigma_a[i]; v.sigma_s[i] = p->sigma_s[i]; }
    v.sigma_t = p->sigma_s[0] + p->sigma_a[0];
    v.PhaseFunctionConstantG = p->g;
    v.densityScaleFactor = p->density_scale;
    v.densityScaleFactorByScaling = p->density_scale / p->world_scaling;
    const float* X = s->desc.slots[0].xform;
    auto len3 = [](const float* r) { return std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); };
    v.tStep = (len3(X) + len3(X + 4) + len3(X + 8)) / 3.f;
    v.hasEmission = haveT ? 1 : 0; v.hasVelocity = haveV ? 1 : 0; v.hasAnimation = 0; v.lastFrameHasEmission = 0;
    v.LeScale = p->LeScale; v.temperatureCutOff = p->temperatureCutOff; v.temperatureScale = p->temperatureScale;
    v.velocityScale = 1.f; v.numMips = built; v.usePrevGridForReproj = 0;
    v.volumeWorldScaling = p->world_scaling;
    v.superVoxelWorldSpaceDiagonalLength = 8.f * std::sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2] + X[4] * X[4] + X[5] * X[5] + X[6] * X[6] + X[8] * X[8] + X[9] * X[9] + X[10] * X[10]);
    if (haveT) { s->lut.resize(512); vrestir_make_blackbody_lut(s->lut.data()); s->desc.blackbody_lut = s->lut.data(); }
    for (int i = 0; i < 3; i++) s->params.dim[i] = (int)s->desc.slots[0].bmax[i];
    s->params.num_mips = built;
    *out = s.release();
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

}  // extern "C"
