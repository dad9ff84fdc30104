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

}  // extern "C"
