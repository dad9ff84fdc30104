// vr_post.cu — the two passes that sit behind VolumetricReSTIR.accumulated_color in the reference's scripts
// (graph: VR/Scripts/run_bunny_tree.py:9-21, run_bistro.py:15-24):
//
//   AccumulatePass     Source/RenderPasses/AccumulatePass/Accumulate.cs.slang:57-122 (three precision modes),
//                      AccumulatePass.cpp:128-205 (frame counter, auto reset, sub-frame count, pass-through when disabled)
//   ErrorMeasurePass   Source/RenderPasses/ErrorMeasurePass/ErrorMeasurer.cs.slang:41-61 (per-pixel difference),
//                      ErrorMeasurePass.cpp:217-259 (sum / pixel count, rgb + average)
//
// Both are pure HBM streams (16-48 B per pixel), one thread per pixel, float4 accesses.  The reduction of the error pass is a
// fixed-order two-level sum in double (the reference reduces float4 in an unspecified tree order): deterministic run to run.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vrestir.h"
#include "vr_host.h"

namespace vr { int setError(int code, const std::string& msg); }
using vr::setError;

#define CKP(x)                                                                                                   \
    do {                                                                                                         \
        cudaError_t e_ = (x);                                                                                    \
        if (e_ != cudaSuccess) return setError(VRESTIR_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct vrestir_accumulator {
    int device = 0, W = 0, H = 0;
    int precision = VRESTIR_ACCUM_DOUBLE;   // AccumulatePass.h:91-94 defaults
    bool enable = true, autoReset = true;
    int subFrameCount = 0;
    int frameCount = 0;
    float4* sum = nullptr; float4* corr = nullptr; double* dsum = nullptr;
};

namespace {

// count == 0: the reference clears the sum textures first (AccumulatePass.cpp prepareAccumulation), i.e. sum = 0 + cur
__global__ void k_accum_single(const float4* __restrict__ cur, float4* __restrict__ out, float4* __restrict__ sum, int W, int row0, int rows, unsigned count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * rows) return;
    const size_t p = (size_t)row0 * W + i;
    const float4 c = cur[p];
    const float4 s0 = count ? sum[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 s = make_float4(s0.x + c.x, s0.y + c.y, s0.z + c.z, s0.w + c.w);
    const float n = (float)(count + 1);
    sum[p] = s;
    out[p] = make_float4(s.x / n, s.y / n, s.z / n, s.w / n);
}

__device__ __forceinline__ void kahan(float cur, float s, float c, float n, float& sNext, float& cNext, float& o) {
    const float y = cur - c;
    sNext = s + y;
    o = sNext / n;
    cNext = (sNext - s) - y;
}
__global__ void k_accum_kahan(const float4* __restrict__ cur, float4* __restrict__ out, float4* __restrict__ sum, float4* __restrict__ corr, int W, int row0, int rows, unsigned count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * rows) return;
    const size_t p = (size_t)row0 * W + i;
    const float4 c = cur[p];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 s = count ? sum[p] : z, k = count ? corr[p] : z;
    const float n = (float)(count + 1);
    float4 sn, kn, o;
    kahan(c.x, s.x, k.x, n, sn.x, kn.x, o.x); kahan(c.y, s.y, k.y, n, sn.y, kn.y, o.y);
    kahan(c.z, s.z, k.z, n, sn.z, kn.z, o.z); kahan(c.w, s.w, k.w, n, sn.w, kn.w, o.w);
    sum[p] = sn; corr[p] = kn; out[p] = o;
}

__global__ void k_accum_double(const float4* __restrict__ cur, float4* __restrict__ out, double* __restrict__ dsum, int W, int row0, int rows, unsigned count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * rows) return;
    const size_t p = (size_t)row0 * W + i;
    const float4 c = cur[p];
    double2* d = (double2*)(dsum + 4 * p);
    double2 a = count ? d[0] : make_double2(0.0, 0.0), b = count ? d[1] : make_double2(0.0, 0.0);
    a.x += (double)c.x; a.y += (double)c.y; b.x += (double)c.z; b.y += (double)c.w;
    const double n = (double)(count + 1);
    d[0] = a; d[1] = b;
    out[p] = make_float4((float)(a.x / n), (float)(a.y / n), (float)(b.x / n), (float)(b.y / n));
}

// ErrorMeasurer.cs.slang:41-61 + first level of the sum: one partial (double3) per block, fixed order
__global__ void __launch_bounds__(256) k_error_partial(const float4* __restrict__ src, const float4* __restrict__ ref, const float4* __restrict__ worldPos, float4* __restrict__ diffOut,
                                                       size_t n, int ignoreBackground, int sqr, int average, double* __restrict__ partial) {
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const bool valid = !ignoreBackground || worldPos[p].w != 0.0f;
        const float4 s = src[p], r = ref[p];
        float dx = valid ? fabsf(s.x - r.x) : 0.f, dy = valid ? fabsf(s.y - r.y) : 0.f, dz = valid ? fabsf(s.z - r.z) : 0.f;
        if (sqr) { dx *= dx; dy *= dy; dz *= dz; }
        if (average) { const float a = (dx + dy + dz) / 3.f; dx = dy = dz = a; }
        if (diffOut) diffOut[p] = make_float4(dx, dy, dz, 0.f);
        ax += (double)dx; ay += (double)dy; az += (double)dz;
    }
    __shared__ double sh[3][256];
    sh[0][threadIdx.x] = ax; sh[1][threadIdx.x] = ay; sh[2][threadIdx.x] = az;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh[0][threadIdx.x] += sh[0][threadIdx.x + w]; sh[1][threadIdx.x] += sh[1][threadIdx.x + w]; sh[2][threadIdx.x] += sh[2][threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[3 * blockIdx.x] = sh[0][0]; partial[3 * blockIdx.x + 1] = sh[1][0]; partial[3 * blockIdx.x + 2] = sh[2][0]; }
}
__global__ void k_error_final(const double* __restrict__ partial, int blocks, double* __restrict__ out3) {
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int b = 0; b < blocks; b++) s += partial[3 * b + threadIdx.x];
        out3[threadIdx.x] = s;
    }
}

// ---- ToneMapper (Source/RenderPasses/ToneMapper/{ToneMapping,Luminance}.ps.slang) ----------------------------------------
__device__ __forceinline__ float tmLuminance(float3 c) { return c.x * 0.299f + c.y * 0.587f + c.z * 0.114f; }
// Luminance.ps.slang:37-42 rendered into the lower-power-of-two target of ToneMapper::createLuminanceFbo with the linear
// sampler (Falcor's default addressing: wrap): log2(max(1e-4, luminance(bilinear(src))))
__global__ void k_tm_luminance(const float4* __restrict__ src, int w, int h, float* __restrict__ dst, int w2, int h2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w2) return;
    const float u = ((float)i + 0.5f) / (float)w2 * (float)w - 0.5f, v = ((float)j + 0.5f) / (float)h2 * (float)h - 0.5f;
    const float u0 = floorf(u), v0 = floorf(v), fu = u - u0, fv = v - v0;
    int x0 = (int)u0 % w, y0 = (int)v0 % h; if (x0 < 0) x0 += w; if (y0 < 0) y0 += h;
    const int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
    const float4 a = src[(size_t)y0 * w + x0], b = src[(size_t)y0 * w + x1], c = src[(size_t)y1 * w + x0], d = src[(size_t)y1 * w + x1];
    float3 t, bt, r;
    t.x = __fmaf_rn(fu, b.x - a.x, a.x); t.y = __fmaf_rn(fu, b.y - a.y, a.y); t.z = __fmaf_rn(fu, b.z - a.z, a.z);
    bt.x = __fmaf_rn(fu, d.x - c.x, c.x); bt.y = __fmaf_rn(fu, d.y - c.y, c.y); bt.z = __fmaf_rn(fu, d.z - c.z, c.z);
    r.x = __fmaf_rn(fv, bt.x - t.x, t.x); r.y = __fmaf_rn(fv, bt.y - t.y, t.y); r.z = __fmaf_rn(fv, bt.z - t.z, t.z);
    dst[(size_t)j * w2 + i] = log2f(fmaxf(0.0001f, tmLuminance(r)));
}
// generateMips: one 2x2 box level, ((a + b) + (c + d)) * 0.25 (a 1-wide / 1-high level averages the two texels it has)
__global__ void k_tm_mip(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst, int dw, int dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= dw) return;
    const int x0 = min(2 * i, sw - 1), x1 = min(2 * i + 1, sw - 1), y0 = min(2 * j, sh - 1), y1 = min(2 * j + 1, sh - 1);
    dst[(size_t)j * dw + i] = ((src[(size_t)y0 * sw + x0] + src[(size_t)y0 * sw + x1]) + (src[(size_t)y1 * sw + x0] + src[(size_t)y1 * sw + x1])) * 0.25f;
}
__device__ __forceinline__ float3 tmUc2(float3 c) {
    const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f;
    float3 o;
    o.x = ((c.x * (A * c.x + C * B) + D * E) / (c.x * (A * c.x + B) + D * F)) - (E / F);
    o.y = ((c.y * (A * c.y + C * B) + D * E) / (c.y * (A * c.y + B) + D * F)) - (E / F);
    o.z = ((c.z * (A * c.z + C * B) + D * E) / (c.z * (A * c.z + B) + D * F)) - (E / F);
    return o;
}
// ToneMapping.ps.slang:142-167
__global__ void k_tonemap(const float4* __restrict__ src, float4* __restrict__ dst, size_t n, vrestir_tonemap_params P, const float* __restrict__ avgLogLum) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 color = src[p];
    float3 c = make_float3(color.x, color.y, color.z);
    if (P.autoExposure) {
        const float avgLuminance = exp2f(*avgLogLum);
        const float k = 0.042f / avgLuminance;
        c.x *= k; c.y *= k; c.z *= k;
    }
    const float* M = P.colorTransform;   // mul(color, (float3x3)M): row vector times matrix, row-major storage
    float3 t = make_float3(c.x * M[0] + c.y * M[3] + c.z * M[6], c.x * M[1] + c.y * M[4] + c.z * M[7], c.x * M[2] + c.y * M[5] + c.z * M[8]);
    switch (P.op) {
        case VRESTIR_TONEMAP_REINHARD: { const float l = tmLuminance(t), r = l / (l + 1.f), k = r / l; t.x *= k; t.y *= k; t.z *= k; break; }
        case VRESTIR_TONEMAP_REINHARD_MODIFIED: {
            const float l = tmLuminance(t), r = l * (1.f + l / (P.whiteMaxLuminance * P.whiteMaxLuminance)) * (1.f + l), k = r / l;
            t.x *= k; t.y *= k; t.z *= k; break;
        }
        case VRESTIR_TONEMAP_HEJI_HABLE_ALU: {
            t.x = fmaxf(0.f, t.x - 0.004f); t.y = fmaxf(0.f, t.y - 0.004f); t.z = fmaxf(0.f, t.z - 0.004f);
            t.x = (t.x * (6.2f * t.x + 0.5f)) / (t.x * (6.2f * t.x + 1.7f) + 0.06f);
            t.y = (t.y * (6.2f * t.y + 0.5f)) / (t.y * (6.2f * t.y + 1.7f) + 0.06f);
            t.z = (t.z * (6.2f * t.z + 0.5f)) / (t.z * (6.2f * t.z + 1.7f) + 0.06f);
            t.x = powf(t.x, 2.2f); t.y = powf(t.y, 2.2f); t.z = powf(t.z, 2.2f); break;
        }
        case VRESTIR_TONEMAP_HABLE_UC2: {
            t = tmUc2(make_float3(2.f * t.x, 2.f * t.y, 2.f * t.z));
            const float ws = 1.f / tmUc2(make_float3(P.whiteScale, P.whiteScale, P.whiteScale)).x;
            t.x *= ws; t.y *= ws; t.z *= ws; break;
        }
        case VRESTIR_TONEMAP_ACES: {
            t.x *= 0.6f; t.y *= 0.6f; t.z *= 0.6f;
            const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
            t.x = __saturatef((t.x * (A * t.x + B)) / (t.x * (C * t.x + D) + E));
            t.y = __saturatef((t.y * (A * t.y + B)) / (t.y * (C * t.y + D) + E));
            t.z = __saturatef((t.z * (A * t.z + B)) / (t.z * (C * t.z + D) + E)); break;
        }
        default: break;   // linear
    }
    if (P.clamp) { t.x = __saturatef(t.x); t.y = __saturatef(t.y); t.z = __saturatef(t.z); }
    dst[p] = make_float4(t.x, t.y, t.z, color.w);
}

void freeAccum(vrestir_accumulator* a) {
    if (a->sum) cudaFree(a->sum);
    if (a->corr) cudaFree(a->corr);
    if (a->dsum) cudaFree(a->dsum);
    a->sum = a->corr = nullptr; a->dsum = nullptr;
}

}  // namespace

extern "C" {

int vrestir_accum_create(int device, int width, int height, vrestir_accumulator** out) try {
    if (!out || width < 1 || height < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return setError(VRESTIR_ERR_CUDA, std::string("no CUDA device: the accumulate pass has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= count) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad device index");
    auto* a = new vrestir_accumulator();
    a->device = device; a->W = width; a->H = height;
    *out = a;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_destroy(vrestir_accumulator* a) try {
    if (!a) return VRESTIR_OK;
    cudaSetDevice(a->device);
    cudaDeviceSynchronize();
    freeAccum(a);
    delete a;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_update(vrestir_accumulator* a, const char* key, double value) try {
    if (!a || !key) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    const std::string k(key);
    if (k == "enableAccumulation") a->enable = value != 0;
    else if (k == "autoReset") a->autoReset = value != 0;
    else if (k == "subFrameCount") a->subFrameCount = (int)value;
    else if (k == "precisionMode") {
        const int m = (int)value;
        if (m < VRESTIR_ACCUM_DOUBLE || m > VRESTIR_ACCUM_SINGLE_COMPENSATED) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "precisionMode must be 0 (Double), 1 (Single) or 2 (SingleCompensated)");
        if (m != a->precision) { a->precision = m; a->frameCount = 0; }   // the sum buffers of the other mode hold nothing
    } else return setError(VRESTIR_WARN_UNKNOWN_KEY, "Unknown field '" + k + "' in an AccumulatePass dictionary");
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_reset(vrestir_accumulator* a) try {
    if (!a) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    a->frameCount = 0;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_resize(vrestir_accumulator* a, int width, int height) try {
    if (!a || width < 1 || height < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    if (width != a->W || height != a->H) {   // AccumulatePass.cpp:120-126
        CKP(cudaSetDevice(a->device));
        CKP(cudaDeviceSynchronize());
        freeAccum(a);
        a->W = width; a->H = height; a->frameCount = 0;
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_frame_count(const vrestir_accumulator* a, int* out) try {
    if (!a || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *out = a->frameCount;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_accum_execute(vrestir_accumulator* a, const float* input, float* output, int row_begin, int row_end, void* stream) try {
    if (!a || !input || !output) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (row_begin < 0 || row_end > a->H || row_begin >= row_end) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad row band");
    cudaStream_t st = (cudaStream_t)stream;
    CKP(cudaSetDevice(a->device));
    if (a->autoReset && a->subFrameCount > 0 && a->frameCount == a->subFrameCount) a->frameCount = -1;   // AccumulatePass.cpp:132-138
    if (a->autoReset && a->frameCount == -1) return VRESTIR_OK;                                          // :163 (output keeps the finished average)
    const int rows = row_end - row_begin;
    const size_t off = (size_t)row_begin * a->W, cnt = (size_t)rows * a->W;
    if (!a->enable) {   // :181-186 blit
        if (input != output) CKP(cudaMemcpyAsync(output + off * 4, input + off * 4, cnt * 16, cudaMemcpyDeviceToDevice, st));
        return VRESTIR_OK;
    }
    const size_t n = (size_t)a->W * a->H;
    if (a->precision == VRESTIR_ACCUM_DOUBLE) { if (!a->dsum) CKP(cudaMalloc(&a->dsum, n * 32)); }
    else {
        if (!a->sum) CKP(cudaMalloc(&a->sum, n * 16));
        if (a->precision == VRESTIR_ACCUM_SINGLE_COMPENSATED && !a->corr) CKP(cudaMalloc(&a->corr, n * 16));
    }
    const unsigned count = (unsigned)a->frameCount++;
    const int threads = 256, blocks = (int)((cnt + threads - 1) / threads);
    const float4* in4 = (const float4*)input; float4* out4 = (float4*)output;
    if (a->precision == VRESTIR_ACCUM_SINGLE) k_accum_single<<<blocks, threads, 0, st>>>(in4, out4, a->sum, a->W, row_begin, rows, count);
    else if (a->precision == VRESTIR_ACCUM_SINGLE_COMPENSATED) k_accum_kahan<<<blocks, threads, 0, st>>>(in4, out4, a->sum, a->corr, a->W, row_begin, rows, count);
    else k_accum_double<<<blocks, threads, 0, st>>>(in4, out4, a->dsum, a->W, row_begin, rows, count);
    CKP(cudaGetLastError());
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_error_measure(int device, const float* source, const float* reference, const float* world_position, int width, int height,
                          int ignore_background, int compute_squared_difference, int compute_average, float* difference_out,
                          float error_rgb_avg[4], void* stream) {
    if (!source || !reference || !error_rgb_avg || width < 1 || height < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CKP(cudaSetDevice(device));
    const size_t n = (size_t)width * height;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = (int)std::min<size_t>((size_t)sms * 8, (n + 255) / 256);
    double* scratch = nullptr;
    CKP(cudaMallocAsync(&scratch, ((size_t)blocks * 3 + 3) * sizeof(double), st));
    // an unbound world-position texture switches the background test off (ErrorMeasurePass.cpp:208-209)
    const int ignore = ignore_background && world_position;
    k_error_partial<<<blocks, 256, 0, st>>>((const float4*)source, (const float4*)reference, (const float4*)world_position, (float4*)difference_out, n, ignore,
                                            compute_squared_difference, compute_average, scratch);
    k_error_final<<<1, 32, 0, st>>>(scratch, blocks, scratch + (size_t)blocks * 3);
    CKP(cudaGetLastError());
    double sum[3];
    CKP(cudaMemcpyAsync(sum, scratch + (size_t)blocks * 3, sizeof(sum), cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    CKP(cudaFreeAsync(scratch, st));
    const float pixelCountf = (float)(width * height);   // ErrorMeasurePass.cpp:241-243
    for (int i = 0; i < 3; i++) error_rgb_avg[i] = (float)sum[i] / pixelCountf;
    error_rgb_avg[3] = (error_rgb_avg[0] + error_rgb_avg[1] + error_rgb_avg[2]) / 3.f;
    return VRESTIR_OK;
}


// ---- ToneMapper host side (Source/RenderPasses/ToneMapper/ToneMapper.cpp:502-522, F/Utils/Color/ColorUtils.h:60-216) -------
namespace {
void mat3mul(const double* a, const double* b, double* o) { double t[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j]; t[i * 3 + j] = s; } memcpy(o, t, sizeof(t)); }
// matrices below are written row-major for column vectors (c' = M c); the reference's glm initialisers list columns
const double kRGBtoXYZ[9] = {0.4123907992659595, 0.3575843393838780, 0.1804807884018343, 0.2126390058715104, 0.7151686787677559, 0.0721923153607337, 0.0193308187155918, 0.1191947797946259, 0.9505321522496608};
const double kXYZtoRGB[9] = {3.2409699419045213, -1.5373831775700935, -0.4986107602930033, -0.9692436362808798, 1.8759675015077206, 0.0415550574071756, 0.0556300796969936, -0.2039769588889765, 1.0569715142428784};
const double kXYZtoLMS[9] = {0.7328, 0.4296, -0.1624, -0.7036, 1.6975, 0.0061, 0.0030, 0.0136, 0.9834};
const double kLMStoXYZ[9] = {1.096123820835514, -0.278869000218287, 0.182745179382773, 0.454369041975359, 0.473533154307412, 0.072097803717229, -0.009627608738429, -0.005698031216113, 1.015325639954543};
void temperatureToXYZ(float T, float out[3]) {   // colorTemperatureToXYZ, Y = 1
    const double t = T, t2 = t * t, t3 = t * t * t;
    double xc = T < 4000.f ? -0.2661239e9 / t3 - 0.2343580e6 / t2 + 0.8776956e3 / t + 0.179910 : -3.0258469e9 / t3 + 2.1070379e6 / t2 + 0.2226347e3 / t + 0.240390;
    const double x = xc, x2 = x * x, x3 = x * x * x;
    double yc = T < 2222.f ? -1.1063814 * x3 - 1.34811020 * x2 + 2.18555832 * x - 0.20219683
              : (T < 4000.f ? -0.9549476 * x3 - 1.37418593 * x2 + 2.09137015 * x - 0.16748867 : 3.0817580 * x3 - 5.87338670 * x2 + 3.75112997 * x - 0.37001483);
    const float xf = (float)xc, yf = (float)yc;
    out[0] = xf * 1.f / yf; out[1] = 1.f; out[2] = (1.f - xf - yf) * 1.f / yf;
}
}  // namespace

void vrestir_tonemap_default_settings(vrestir_tonemap_settings* s) {
    if (!s) return;
    memset(s, 0, sizeof(*s));
    s->exposureCompensation = 0.f; s->autoExposure = 0; s->filmSpeed = 100.f; s->whiteBalance = 0; s->whitePoint = 6500.f;
    s->op = VRESTIR_TONEMAP_ACES; s->clamp = 1; s->whiteMaxLuminance = 1.f; s->whiteScale = 11.2f; s->fNumber = 1.f; s->shutter = 1.f;
}
int vrestir_tonemap_params_from_settings(const vrestir_tonemap_settings* s, vrestir_tonemap_params* out) try {
    if (!s || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (s->op > VRESTIR_TONEMAP_ACES) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "unknown tone-mapping operator");
    double wb[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (s->whiteBalance) {
        if (s->whitePoint < 1667.f || s->whitePoint > 25000.f) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "white point outside 1667 K .. 25000 K");
        double MA[9], invMA[9];
        mat3mul(kXYZtoLMS, kRGBtoXYZ, MA); mat3mul(kXYZtoRGB, kLMStoXYZ, invMA);
        float d65[3], src[3]; temperatureToXYZ(6500.f, d65); temperatureToXYZ(s->whitePoint, src);
        double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; i++) {
            float wd = 0.f, ws = 0.f;
            for (int k = 0; k < 3; k++) { wd += (float)kXYZtoLMS[i * 3 + k] * d65[k]; ws += (float)kXYZtoLMS[i * 3 + k] * src[k]; }
            D[i * 4] = (double)(wd / ws);
        }
        double t[9]; mat3mul(D, MA, t); mat3mul(invMA, t, wb);
    }
    const float exposureScale = powf(2.f, s->exposureCompensation);
    float manual = 1.f;
    if (!s->autoExposure) manual = ((1.f / 100.f) * s->filmSpeed) / (s->shutter * s->fNumber * s->fNumber);
    // the shader computes mul(color, (float3x3)colorTransform) with the glm (column-vector) matrix uploaded as is, i.e. color * M^T
    // in HLSL's row-vector reading = M * color: store M^T so that the kernel's row-vector product gives M * color
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out->colorTransform[j * 3 + i] = (float)wb[i * 3 + j] * exposureScale * manual;
    out->op = s->op; out->autoExposure = s->autoExposure; out->clamp = s->clamp; out->whiteScale = std::max(0.001f, s->whiteScale); out->whiteMaxLuminance = s->whiteMaxLuminance;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_tonemap_execute(int device, const vrestir_tonemap_params* P, const float* src, float* dst, int width, int height, float* avg_log_luminance_out, void* stream) try {
    if (!P || !src || !dst || width < 1 || height < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return setError(VRESTIR_ERR_CUDA, std::string("no CUDA device: the tone mapper has no CPU fallback (") + cudaGetErrorString(e) + ")");
    CKP(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    float* lum = nullptr;
    const float* avg = nullptr;
    if (P->autoExposure) {
        int w2 = 1, h2 = 1; while (w2 * 2 <= width) w2 *= 2; while (h2 * 2 <= height) h2 *= 2;     // getLowerPowerOf2
        size_t total = 0; for (int a = w2, b = h2;; a = std::max(1, a / 2), b = std::max(1, b / 2)) { total += (size_t)a * b; if (a == 1 && b == 1) break; }
        CKP(cudaMallocAsync((void**)&lum, total * sizeof(float), st));
        k_tm_luminance<<<dim3((w2 + 127) / 128, h2), 128, 0, st>>>((const float4*)src, width, height, lum, w2, h2);
        float* cur = lum; int cw = w2, ch = h2;
        while (cw > 1 || ch > 1) {
            const int nw = std::max(1, cw / 2), nh = std::max(1, ch / 2);
            float* nxt = cur + (size_t)cw * ch;
            k_tm_mip<<<dim3((nw + 127) / 128, nh), 128, 0, st>>>(cur, cw, ch, nxt, nw, nh);
            cur = nxt; cw = nw; ch = nh;
        }
        avg = cur;
        if (avg_log_luminance_out) CKP(cudaMemcpyAsync(avg_log_luminance_out, avg, 4, cudaMemcpyDeviceToHost, st));
    }
    const size_t n = (size_t)width * height;
    k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)src, (float4*)dst, n, *P, avg);
    CKP(cudaGetLastError());
    if (lum) CKP(cudaFreeAsync(lum, st));
    if (avg_log_luminance_out && P->autoExposure) CKP(cudaStreamSynchronize(st));
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

}  // extern "C"
