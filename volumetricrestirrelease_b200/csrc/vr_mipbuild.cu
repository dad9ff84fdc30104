// vr_mipbuild.cu — the mip / conservative-mip chain of a dense density grid, built on the GPU.
//
// Rule: gvdb-voxel-src/source/gvdb_library/src/gvdb_volume_gvdb.cpp:2703-2885 (conservative mip 0 :2753-2801, 2x box /
// 3-tap polyphase down-sampling :2803-2862) and the storage rules of F/Scene/Scene.cpp:3139-3174 (1e-9 flush, UNORM8 codes,
// conservative codes never round a positive value down to 0) — the same arithmetic, in the same order, as the host builder in
// vr_scene.cpp (makeConservative0 / downsample / buildSlot), so the two produce identical bits (tests/test_gpu_parity.py).
// This is the converter's job in the reference (one .vbx per mip, prepared offline); animated sequences need it per frame
// (200 frames x 8 grids).  All kernels are pure HBM streams: one thread per destination voxel, x fastest (coalesced).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vrestir.h"
#include "vr_mipbuild.h"
#include "vr_host.h"
#include "vr_procedural.h"

namespace vr { int setError(int code, const std::string& msg); }
using vr::setError;

#define CKM(x)                                                                                                   \
    do {                                                                                                         \
        cudaError_t e_ = (x);                                                                                    \
        if (e_ != cudaSuccess) return setError(VRESTIR_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct vrestir_mip_chain {
    int device = 0, numMips = 0;
    struct Level { void* data = nullptr; float* raw = nullptr; uint8_t* active = nullptr; int dim[3] = {0, 0, 0}; int format = 0; float maxValue = 1.f; size_t bytes = 0; };
    Level lev[2][VRESTIR_NUM_MAX_MIPS];   // [conservative][mip]
    unsigned* maxBits = nullptr;          // one per level: bits of max |v|
};

namespace {

struct Dim { int nx, ny, nz; };
__device__ __forceinline__ float at(const float* __restrict__ v, Dim d, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= d.nx || y >= d.ny || z >= d.nz) return 0.f;
    return v[((size_t)z * d.ny + y) * d.nx + x];
}

// conservative mip 0: a zero voxel takes the mean of its positive 27-neighbourhood (x offset outermost, z innermost)
__global__ void k_conservative0(const float* __restrict__ src, float* __restrict__ dst, Dim d) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= d.nx) return;
    const size_t o = ((size_t)z * d.ny + y) * d.nx + x;
    const float org = src[o];
    float out = org;
    if (org == 0.f) {
        float avg = 0.f;
        for (int ii = -1; ii <= 1; ii++)
            for (int jj = -1; jj <= 1; jj++)
                for (int kk = -1; kk <= 1; kk++) { const float t = at(src, d, x + ii, y + jj, z + kk); avg += t > 0.f ? t : 0.f; }
        avg /= 27.f;
        if (avg > 0.f) out = avg;
    }
    dst[o] = out;
}

__device__ __forceinline__ void weights(int n, int cur, int i, float w[3]) {
    if (n == 2) { w[0] = w[1] = w[2] = 0.5f; return; }
    const float den = (float)(2 * cur + 1);
    w[0] = (float)(cur - i) / den; w[1] = (float)cur / den; w[2] = (float)(1 + i) / den;
}
// mip k from mip k-1: 2x box on even axes, 3-tap polyphase on odd axes
__global__ void k_downsample(const float* __restrict__ prev, Dim pd, float* __restrict__ dst, Dim d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= d.nx) return;
    const int ni = pd.nx % 2 == 0 ? 2 : 3, nj = pd.ny % 2 == 0 ? 2 : 3, nk = pd.nz % 2 == 0 ? 2 : 3;
    float wi[3], wj[3], wk[3];
    weights(nk, d.nz, k, wk); weights(nj, d.ny, j, wj); weights(ni, d.nx, i, wi);
    float res = 0.f;
    for (int ii = 0; ii < ni; ii++)
        for (int jj = 0; jj < nj; jj++)
            for (int kk = 0; kk < nk; kk++) res += wi[ii] * wj[jj] * wk[kk] * at(prev, pd, 2 * i + ii, 2 * j + jj, 2 * k + kk);
    dst[((size_t)k * d.ny + j) * d.nx + i] = res > 0.f ? res : 0.f;
}

// max |v| (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ v, size_t n, unsigned* __restrict__ out) {
    float m = 0.f;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(v[p]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// storage rules of the brick pool: 1e-9 flush, then fp32 (mip 0) or UNORM8 codes
__global__ void k_store(const float* __restrict__ src, size_t n, const unsigned* __restrict__ maxBits, int format, int conservative, void* __restrict__ dst) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float maxv = __uint_as_float(*maxBits);
    if (maxv <= 0.f) maxv = 1.f;
    float v = src[p];
    if (v / maxv < 1e-9f && v >= 0.f) v = 0.f;
    if (format == VRESTIR_ATLAS_UNORM8) {
        long q = lround(255.0 * (double)(v / maxv));
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        if (q == 0 && v > 0.f && conservative) q = 1;
        ((uint8_t*)dst)[p] = (uint8_t)q;
    } else ((float*)dst)[p] = v;
}

// brick activity: any non-zero RAW value in the 10^3 apron-inclusive block (the host builder's rule, before the 1e-9 flush)
__global__ void k_activity(const float* __restrict__ raw, Dim d, int BX, int BY, int BZ, uint8_t* __restrict__ active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= BX * BY * BZ) return;
    const int bx = b % BX, by = (b / BX) % BY, bz = b / (BX * BY);
    bool any = false;
    for (int z = bz * 8 - 1; z <= bz * 8 + 8 && !any; z++)
        for (int y = by * 8 - 1; y <= by * 8 + 8 && !any; y++)
            for (int x = bx * 8 - 1; x <= bx * 8 + 8; x++) if (at(raw, d, x, y, z) != 0.f) { any = true; break; }
    active[b] = any ? 1 : 0;
}

__global__ void k_pack_bricks(const void* __restrict__ level, Dim d, int format, const vrestir_node* __restrict__ nodes0, size_t total, void* __restrict__ atlas) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const unsigned bi = (unsigned)(gid / VRESTIR_BRICK_VOXELS), r = (unsigned)(gid % VRESTIR_BRICK_VOXELS);
    const int lx = r % 10, ly = (r / 10) % 10, lz = r / 100;
    const int x = nodes0[bi].pos[0] + lx - 1, y = nodes0[bi].pos[1] + ly - 1, z = nodes0[bi].pos[2] + lz - 1;
    const bool in = x >= 0 && y >= 0 && z >= 0 && x < d.nx && y < d.ny && z < d.nz;
    const size_t o = in ? ((size_t)z * d.ny + y) * d.nx + x : 0;
    if (format == VRESTIR_ATLAS_UNORM8) ((uint8_t*)atlas)[gid] = in ? ((const uint8_t*)level)[o] : (uint8_t)0;
    else ((float*)atlas)[gid] = in ? ((const float*)level)[o] : 0.f;
}

__global__ void k_brick_bounds(const void* __restrict__ atlas, int format, float maxv, vrestir_node* __restrict__ nodes0, unsigned brickCount) {
    const unsigned bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= brickCount) return;
    const size_t base = (size_t)bi * VRESTIR_BRICK_VOXELS;
    float mn = 3.402823466e+38f, mx = 0.f, sum = 0.f;
    for (int i = -1; i <= 8; i++) for (int j = -1; j <= 8; j++) for (int k = -1; k <= 8; k++) {
        const size_t idx = base + (size_t)((k + 1) * 10 + (j + 1)) * 10 + (i + 1);
        const float v = format == VRESTIR_ATLAS_UNORM8 ? (float)((const uint8_t*)atlas)[idx] * 0.003921568859368563f * maxv : ((const float*)atlas)[idx];
        mn = fminf(mn, v); mx = fmaxf(mx, v); sum += v;
    }
    nodes0[bi].bounds[0] = mn; nodes0[bi].bounds[1] = mx; nodes0[bi].bounds[2] = sum / 512.f; nodes0[bi].bounds[3] = 0.f;
}

// per brick [10][9][9] words, word(z,y,x) = codes (x,y) (x+1,y) (x,y+1) (x+1,y+1) of plane z
__global__ void k_quad_repack(const uint8_t* __restrict__ atlas, size_t totalWords, uint32_t* __restrict__ quads) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= totalWords) return;
    const unsigned b = (unsigned)(gid / 810), r = (unsigned)(gid % 810);
    const int x = r % 9, y = (r / 9) % 9, z = r / 81;
    const uint8_t* c = atlas + (size_t)b * VRESTIR_BRICK_VOXELS + (z * 10 + y) * 10 + x;
    quads[gid] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[10] << 16) | ((uint32_t)c[11] << 24);
}

// the procedural density field of a synthetic scene, evaluated per voxel on the device (grids too large to build on the host:
// SURVEY.md 8d config 5, ~2048^3); same expressions as the host generator (vr_procedural.h)
__global__ void k_procedural(vrestir_scene_params sp, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= sp.dim[0]) return;
    const float u = ((float)x + 0.5f) / (float)sp.dim[0], v = ((float)y + 0.5f) / (float)sp.dim[1], w = ((float)z + 0.5f) / (float)sp.dim[2];
    dst[((size_t)z * sp.dim[1] + y) * sp.dim[0] + x] = vr::shapeDensity(sp, u, v, w, nullptr, nullptr);
}

void freeChain(vrestir_mip_chain* c) {
    for (auto& kind : c->lev) for (auto& l : kind) { if (l.data) cudaFree(l.data); if (l.raw) cudaFree(l.raw); if (l.active) cudaFree(l.active); l.data = nullptr; l.raw = nullptr; l.active = nullptr; }
    if (c->maxBits) cudaFree(c->maxBits);
    c->maxBits = nullptr;
}

dim3 gridOf(Dim d) { return dim3((d.nx + 127) / 128, d.ny, d.nz); }

}  // namespace

extern "C" {

int vrestir_mips_destroy(vrestir_mip_chain* c) try {
    if (!c) return VRESTIR_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    freeChain(c);
    delete c;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_mips_build_device(int device, const float* dense_mip0, const int32_t dim[3], int num_mips, vrestir_mip_chain** out, void* stream) try {
    if (!dense_mip0 || !dim || !out || dim[0] < 1 || dim[1] < 1 || dim[2] < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    if (dim[1] > 65535 || dim[2] > 65535) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "grid larger than 65535 in y or z");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return setError(VRESTIR_ERR_CUDA, std::string("no CUDA device: the GPU mip builder has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= count) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad device index");
    CKM(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int numMips = std::max(1, std::min((int)VRESTIR_NUM_MAX_MIPS, num_mips));
    auto* c = new vrestir_mip_chain();
    c->device = device;
    auto fail = [&](int rc) { cudaStreamSynchronize(st); freeChain(c); delete c; return rc; };
#define CKF(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(setError(VRESTIR_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_))); } while (0)
    CKF(cudaMalloc(&c->maxBits, 2 * VRESTIR_NUM_MAX_MIPS * sizeof(unsigned)));
    CKF(cudaMemsetAsync(c->maxBits, 0, 2 * VRESTIR_NUM_MAX_MIPS * sizeof(unsigned), st));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    Dim d{dim[0], dim[1], dim[2]};
    const float* cur = dense_mip0;      // raw chains: level m of the normal chain / the conservative chain
    const float* cons = nullptr;
    for (int m = 0; m < numMips; m++) {
        const size_t n = (size_t)d.nx * d.ny * d.nz;
        for (int k = 0; k < 2; k++) {
            vrestir_mip_chain::Level& L = c->lev[k][m];
            L.dim[0] = d.nx; L.dim[1] = d.ny; L.dim[2] = d.nz;
            L.format = (k == 0 && m == 0) ? VRESTIR_ATLAS_F32 : VRESTIR_ATLAS_UNORM8;
            L.bytes = n * (L.format == VRESTIR_ATLAS_F32 ? 4 : 1);
            CKF(cudaMalloc(&L.data, L.bytes));
        }
        if (m == 0) {   // conservative twin of mip 0
            CKF(cudaMalloc(&c->lev[1][0].raw, n * 4));
            k_conservative0<<<gridOf(d), 128, 0, st>>>(cur, c->lev[1][0].raw, d);
            cons = c->lev[1][0].raw;
        }
        const int BX = (d.nx + 7) / 8, BY = (d.ny + 7) / 8, BZ = (d.nz + 7) / 8;
        for (int k = 0; k < 2; k++) {
            const float* raw = k == 0 ? cur : cons;
            CKF(cudaMalloc(&c->lev[k][m].active, (size_t)BX * BY * BZ));
            k_activity<<<(BX * BY * BZ + 127) / 128, 128, 0, st>>>(raw, d, BX, BY, BZ, c->lev[k][m].active);
            unsigned* mb = c->maxBits + (k * VRESTIR_NUM_MAX_MIPS + m);
            k_absmax<<<std::min<size_t>((size_t)sms * 8, (n + 255) / 256), 256, 0, st>>>(raw, n, mb);
            k_store<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, n, mb, c->lev[k][m].format, k, c->lev[k][m].data);
        }
        CKF(cudaGetLastError());
        c->numMips = m + 1;
        if (m + 1 < numMips) {
            if (d.nx < 2 || d.ny < 2 || d.nz < 2) break;   // like the host builder: the chain ends here
            const Dim nd{std::max(1, d.nx / 2), std::max(1, d.ny / 2), std::max(1, d.nz / 2)};
            const size_t nn = (size_t)nd.nx * nd.ny * nd.nz;
            for (int k = 0; k < 2; k++) {
                CKF(cudaMalloc(&c->lev[k][m + 1].raw, nn * 4));
                k_downsample<<<gridOf(nd), 128, 0, st>>>(k == 0 ? cur : cons, d, c->lev[k][m + 1].raw, nd);
            }
            cur = c->lev[0][m + 1].raw; cons = c->lev[1][m + 1].raw;
            d = nd;
        }
    }
    CKF(cudaGetLastError());
    // max values to the host (the only synchronisation), raw chains released
    std::vector<unsigned> bits(2 * VRESTIR_NUM_MAX_MIPS);
    CKF(cudaMemcpyAsync(bits.data(), c->maxBits, bits.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CKF(cudaStreamSynchronize(st));
#undef CKF
    for (int k = 0; k < 2; k++)
        for (int m = 0; m < c->numMips; m++) {
            float mv; memcpy(&mv, &bits[k * VRESTIR_NUM_MAX_MIPS + m], 4);
            c->lev[k][m].maxValue = mv > 0.f ? mv : 1.f;
            if (c->lev[k][m].raw) { cudaFree(c->lev[k][m].raw); c->lev[k][m].raw = nullptr; }
        }
    *out = c;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_make_procedural_device(int device, const vrestir_scene_params* sp, float* dense_out, void* stream) try {
    if (!sp || !dense_out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (sp->dim[0] < 1 || sp->dim[1] < 1 || sp->dim[2] < 1 || sp->dim[1] > 65535 || sp->dim[2] > 65535) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad grid dimensions");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return setError(VRESTIR_ERR_CUDA, std::string("no CUDA device: the device generator has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= count) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad device index");
    CKM(cudaSetDevice(device));
    k_procedural<<<gridOf(Dim{sp->dim[0], sp->dim[1], sp->dim[2]}), 128, 0, (cudaStream_t)stream>>>(*sp, dense_out);
    CKM(cudaGetLastError());
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_mips_level(const vrestir_mip_chain* c, int mip, int conservative, vrestir_mip_level* out) try {
    if (!c || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (mip < 0 || mip >= c->numMips) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "mip level not built");
    const vrestir_mip_chain::Level& L = c->lev[conservative ? 1 : 0][mip];
    out->data = L.data; out->bytes = L.bytes; out->format = L.format; out->max_value = L.maxValue;
    for (int i = 0; i < 3; i++) out->dim[i] = L.dim[i];
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_mips_count(const vrestir_mip_chain* c, int* out) try {
    if (!c || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *out = c->numMips;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

}  // extern "C"

namespace vr {
int chainLevelView(const vrestir_mip_chain* c, int mip, int conservative, ChainLevelView& out) {
    if (!c || mip < 0 || mip >= c->numMips) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "mip level not built");
    const vrestir_mip_chain::Level& L = c->lev[conservative ? 1 : 0][mip];
    out.data = L.data; out.active = L.active; out.format = L.format; out.maxValue = L.maxValue;
    for (int i = 0; i < 3; i++) out.dim[i] = L.dim[i];
    return VRESTIR_OK;
}
int chainDevice(const vrestir_mip_chain* c) { return c ? c->device : -1; }
cudaError_t launchPackBricks(const void* level, const int dim[3], int format, const vrestir_node* nodes0, uint32_t brickCount, void* atlas, cudaStream_t st) {
    const size_t total = (size_t)brickCount * VRESTIR_BRICK_VOXELS;
    if (total) k_pack_bricks<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(level, Dim{dim[0], dim[1], dim[2]}, format, nodes0, total, atlas);
    return cudaGetLastError();
}
cudaError_t launchBrickBounds(const void* atlas, int format, float maxValue, vrestir_node* nodes0, uint32_t brickCount, cudaStream_t st) {
    if (brickCount) k_brick_bounds<<<(brickCount + 127) / 128, 128, 0, st>>>(atlas, format, maxValue, nodes0, brickCount);
    return cudaGetLastError();
}
cudaError_t launchQuadRepack(const uint8_t* atlas, uint32_t brickCount, uint32_t* quads, cudaStream_t st) {
    const size_t total = (size_t)brickCount * 810;
    if (total) k_quad_repack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(atlas, total, quads);
    return cudaGetLastError();
}
}  // namespace vr

