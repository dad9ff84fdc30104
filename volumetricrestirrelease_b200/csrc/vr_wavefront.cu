// vr_wavefront.cu — wavefront (task-stream) form of the reuse stages for sm_100a.
//
// The per-pixel kernels of vr_kernels.cu run every transmittance march of a pixel serially inside one thread; ncu shows
// them issue-bound at ~10/32 active lanes (profiles/r01_ncu_full_k_spatial_k_initial_baseline.txt).  Here a stage is
// split into
//     gather  (one thread per pixel: load taps, point-query densities, emit march tasks into compacted global streams)
//  -> march   (vr_march.cuh: persistent lane pools over the task streams, lean registers, full occupancy)
//  -> combine (one thread per pixel: MIS weights + weighted reservoir streaming from the march results).
// All arithmetic per value is the same as in the per-pixel kernels (same helpers, same operation order), so the output
// is bit-identical to them (tests/test_gpu_parity.py::test_wavefront_equals_per_pixel) and carries their oracle parity.
//
// K3 spatial reuse (VR/SpatialReuse.cs.slang:94-265), Talbot MIS, S <= 4 taps: a pixel needs p-hat of tap i's sample
// seen from ray j for i != j (12 values).  p-hat = Tr(camera ray j -> depth_i) * density * sigma * Tr(point -> light) *
// Ld; the camera transmittances of one ray share ONE multi-depth march (4 camera tasks per pixel instead of 12
// marches), the light marches are 12 independent tasks.
#include "vr_march.cuh"
#include "vr_kernels.h"

#ifndef VR_MARCH_MINB
#define VR_MARCH_MINB 8
#endif

namespace vrd {

// ------------------------------------------------------------------------------------------------ march kernels
template <int NT, bool FAST>
__global__ void __launch_bounds__(128, VR_MARCH_MINB) k_march(const WfStream s, float* results, const MarchKind kind, const DSlot g) {
    marchPool<NT, FAST>(s.tasks, min(*s.count, s.capacity), s.cursor, results, kind, g);
}

// ------------------------------------------------------------------------------------------------ K3 gather
VRD bool tapInImage(const FrameParams& fp, int x, int y, int s, int& tx, int& ty) {
    tx = x + fp.offsets[s].x; ty = y + fp.offsets[s].y;
    return tx >= 0 && tx < fp.W && ty >= 0 && ty < fp.H;
}
VRD float3 tapRayDir(const FrameParams& fp, int tx, int ty) {
    return normalize(camRayDirNN(c_scene.camU, c_scene.camV, c_scene.camW, tx, ty, fp.W, fp.H));
}

__global__ void __launch_bounds__(128) k_spatial_gather(FrameParams fp, WfBufs wf) {
    int x, y;
    bool active = pixelOf(fp, x, y);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int W = fp.W, S = fp.sampleCount;
    const bool talbot = fp.spatialMIS == VRESTIR_MIS_TALBOT;
    const int pixelId = active ? y * W + x : fp.rowBegin * W;
    const unsigned blkBase = (unsigned)(pixelId - fp.rowBegin * W) * WF_BLOCK;
    float* blk = wf.results + blkBase;
    if (active) {
        const int2 cf = fp.features[pixelId];
        if (!(__int_as_float(cf.y) != 1.f)) {   // IsSelfBackground: pass through (VR/SpatialReuse.cs.slang:146-155)
            storeReservoir(fp.out, pixelId, loadReservoir(fp.cur, pixelId, 1));
            active = false;
        }
    }
    const float3 origin = c_scene.camPos;
    unsigned camBits = 0;     // bit j*3+k: camera ray j needs the transmittance to the depth of tap i, k = i - (i > j)
    unsigned lightBits = 0;   // bit i*4+j: light march from ray_j.at(depth_i)
    float depth[4] = {0.f, 0.f, 0.f, 0.f};
    // ---- pass 1: which evaluations exist, density point queries
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i >= S || (i == 0 && !talbot)) continue;
            int txi, tyi;
            if (!tapInImage(fp, x, y, i, txi, tyi)) continue;
            const Reservoir tap = loadReservoir(fp.cur, tyi * W + txi, 1);
            depth[i] = tap.depth;
            const bool wantResample = i > 0 && tap.runningSum != 0.f;                                    // resampleNeighbor
            const bool wantMIS = talbot && (i == 0 ? tap.runningSum > 0.f : tap.runningSum != 0.f);     // superset of "runningSum > 0 after resampling"
            if (!wantResample && !wantMIS) continue;
            const bool bg = tap.depth == kRayTMax;
            bool alive = true;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j >= S || j == i) continue;
                int txj = x, tyj = y;
                if (j == 0) { if (!wantResample) continue; }
                else { if (!wantMIS || !alive) continue; if (!tapInImage(fp, x, y, j, txj, tyj)) continue; }
                const float3 dir = tapRayDir(fp, txj, tyj);
                const Ray r = makeRay(origin, dir, 0.f, tap.depth);
                const float3 pW = r.at(r.tMax);
                const float density = bg ? 1.f : DensityWorldSpace(pW, 0);
                blk[WF_D + i * 4 + j] = density;
                if (density != 0.f) {
                    camBits |= 1u << (j * 3 + (i - (i > j ? 1 : 0)));
                    if (!bg && tap.lightID != VRESTIR_SELF_EMISSION_LIGHT_ID) {
                        Ray sh; float3 Ld;
                        if (lightRayAndLd(makeMI(pW, -dir, true), tap.lightID, tap.lightUV, false, sh, Ld)) lightBits |= 1u << (i * 4 + j);
                    }
                } else if (j == 0) alive = false;   // p-hat on the centre ray is 0: the tap is dropped, no MIS terms
            }
        }
    }
    // ---- pass 2: one reservation per stream per warp, then write the tasks (camera tasks grouped by ray index)
    unsigned camCnt = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) camCnt += ((camBits >> (3 * j)) & 7u) ? 1u : 0u;
    const unsigned camTot = __reduce_add_sync(FULL, camCnt), lightTot = __reduce_add_sync(FULL, (unsigned)__popc(lightBits));
    if (camTot == 0) return;
    unsigned camBase = 0, lightBase = 0;
    if (lane == 0) { camBase = atomicAdd(wf.cam.count, camTot); if (lightTot) lightBase = atomicAdd(wf.light.count, lightTot); }
    camBase = __shfl_sync(FULL, camBase, 0); lightBase = __shfl_sync(FULL, lightBase, 0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const unsigned m = (camBits >> (3 * j)) & 7u;
        const unsigned bal = __ballot_sync(FULL, m != 0);
        if (m) {
            const unsigned pos = camBase + __popc(bal & lt);
            int txj = x, tyj = y;
            if (j > 0) tapInImage(fp, x, y, j, txj, tyj);
            const float3 dir = tapRayDir(fp, txj, tyj);
            // threshold k of ray j is the depth of tap i = k + (k >= j)
            const float t0 = depth[j <= 0 ? 1 : 0], t1 = depth[j <= 1 ? 2 : 1], t2 = depth[j <= 2 ? 3 : 2];
            if (pos < wf.cam.capacity) {
                wf.cam.tasks[2 * (size_t)pos] = make_uint4(__float_as_uint(t0), __float_as_uint(t1), __float_as_uint(t2), m);
                wf.cam.tasks[2 * (size_t)pos + 1] = make_uint4(__float_as_uint(dir.x), __float_as_uint(dir.y), __float_as_uint(dir.z), blkBase + WF_C + j * 3);
            }
        }
        camBase += __popc(bal);
    }
    if (lightTot == 0) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned mi_ = (lightBits >> (4 * i)) & 15u;
        if (!__any_sync(FULL, mi_ != 0)) continue;
        Reservoir tap = createNewReservoir();
        if (mi_) { int txi, tyi; tapInImage(fp, x, y, i, txi, tyi); tap = loadReservoir(fp.cur, tyi * W + txi, 1); }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j == i) continue;
            const bool has = (mi_ >> j) & 1u;
            const unsigned bal = __ballot_sync(FULL, has);
            if (has) {
                const unsigned pos = lightBase + __popc(bal & lt);
                int txj = x, tyj = y;
                if (j > 0) tapInImage(fp, x, y, j, txj, tyj);
                const float3 dir = tapRayDir(fp, txj, tyj);
                const Ray r = makeRay(origin, dir, 0.f, tap.depth);
                const float3 pW = r.at(r.tMax);
                Ray sh; float3 Ld;
                lightRayAndLd(makeMI(pW, -dir, true), tap.lightID, tap.lightUV, false, sh, Ld);
                if (pos < wf.light.capacity) {
                    wf.light.tasks[2 * (size_t)pos] = make_uint4(__float_as_uint(sh.origin.x), __float_as_uint(sh.origin.y), __float_as_uint(sh.origin.z), __float_as_uint(sh.tMax));
                    wf.light.tasks[2 * (size_t)pos + 1] = make_uint4(__float_as_uint(sh.dir.x), __float_as_uint(sh.dir.y), __float_as_uint(sh.dir.z), blkBase + WF_L + i * 4 + j);
                }
            }
            lightBase += __popc(bal);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3 combine
// evaluate_F_ / evaluate_P_hat (VR/ReSTIRHelper.slang:91-423, B == 1, current frame) with the density and the two
// transmittances taken from the pixel's result block
VRD float wfPHat(const Reservoir& tap, float3 origin, float3 dir, const float* blk, int i, int j) {
    const vrestir_volume_desc& vd = c_scene.vol;
    Ray ray = makeRay(origin, dir, 0.f, tap.depth);
    const bool isBackgroundSample = tap.depth == kRayTMax;
    const bool isSelfEmission = tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID;
    float3 F = f3(1.f);
    const float3 p_World = ray.at(ray.tMax);
    const MediumInteraction mi = makeMI(p_World, -ray.dir, true);
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    const float density = blk[WF_D + i * 4 + j];
    if (density == 0.f) return luminance(f3(0.f));
    const float visibility = blk[WF_C + j * 3 + (i - (i > j ? 1 : 0))];
    const float3 sigma_s = isBackgroundSample ? f3(1.f) : (isSelfEmission ? sigA : sigS);
    F = F * (visibility * density * sigma_s);
    if (any_gt0(F)) {
        if (isBackgroundSample) F = F * envEval(ray.dir, false);
        else if (isSelfEmission) F = F * EmissionWorldSpace(p_World, false);
        else {
            Ray sh; float3 Ld;
            const bool valid = lightRayAndLd(mi, tap.lightID, tap.lightUV, false, sh, Ld);
            const float Tr = valid ? blk[WF_L + i * 4 + j] : 1.f;
            F = F * (Tr * Ld);
        }
    }
    return luminance(F);
}

__global__ void __launch_bounds__(128) k_spatial_combine(FrameParams fp, WfBufs wf) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int W = fp.W, S = fp.sampleCount;
    const int pixelId = y * W + x;
    const int2 cf = fp.features[pixelId];
    if (!(__int_as_float(cf.y) != 1.f)) return;   // passed through by the gather kernel
    const float* blk = wf.results + (size_t)(pixelId - fp.rowBegin * W) * WF_BLOCK;
    const int numRounds = fp.spatialRounds + fp.roundOffset + 1;
    const int roundId = fp.roundId + fp.roundOffset;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(numRounds * fp.frameCount + roundId));
    const uint32_t mis = fp.spatialMIS;
    Reservoir output = loadReservoir(fp.cur, pixelId, 1);
    if (mis == VRESTIR_MIS_TALBOT) output = createNewReservoir();
    const float3 origin = c_scene.camPos;
    const float3 dir0 = tapRayDir(fp, x, y);
    const int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
    for (int sampleId = startSampleId; sampleId < S; sampleId++) {
        int tx, ty;
        if (!tapInImage(fp, x, y, sampleId, tx, ty)) continue;
        Reservoir tap = loadReservoir(fp.cur, ty * W + tx, 1);
        float MISWeight = 1.f;
        if (sampleId > 0 && tap.runningSum != 0.f) {   // resampleNeighbor
            const float p_y_hat = wfPHat(tap, origin, dir0, blk, sampleId, 0);
            float weight = p_y_hat / tap.p_y;
            if (isinf(weight) || isnan(weight)) weight = 0.f;
            tap.runningSum *= weight;
            tap.p_y = p_y_hat;
        }
        if (mis == VRESTIR_MIS_TALBOT && tap.runningSum > 0.f) {
            float p_sum = 0, p_qi = 0, k = 0;
            for (int j = 0; j < S; j++) {
                int tx2, ty2;
                if (!tapInImage(fp, x, y, j, tx2, ty2)) continue;
                const float4 t2 = __ldg(&fp.cur.p0[ty2 * W + tx2]);   // (runningSum, M, depth, p_y)
                k += t2.y;
                if (j == 0) { p_qi = tap.p_y; p_sum += tap.p_y * t2.y; }
                else if (sampleId == j) { p_qi = t2.w; p_sum += t2.w * t2.y; }
                else {
                    float p_y = wfPHat(tap, origin, tapRayDir(fp, tx2, ty2), blk, sampleId, j);
                    if (isinf(p_y) || isnan(p_y)) p_y = 0.f;
                    p_sum += p_y * t2.y;
                }
            }
            if (p_sum > 0) MISWeight = p_qi * k / p_sum;
        }
        tap.runningSum *= MISWeight;
        simpleResampleStep<1>(tap, output, sg);
    }
    storeReservoir(fp.out, pixelId, output);
}

// ------------------------------------------------------------------------------------------------ K1 finish
// VR/TraceRays.cs.slang:176-183: p-hat of the pixel's own reservoir on its own ray under the spatial options
__global__ void __launch_bounds__(128) k_initial_finish(FrameParams fp, WfBufs wf) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int pixelId = y * fp.W + x;
    const float4 a = fp.cur.p0[pixelId];   // (runningSum, M, depth, p_y)
    if (!(a.x > 0.f)) return;
    Reservoir r = loadReservoirRW(fp.cur, pixelId, 1);
    const float* blk = wf.results + (size_t)(pixelId - fp.rowBegin * fp.W) * WF_BLOCK;
    const float p_hat = wfPHat(r, c_scene.camPos, tapRayDir(fp, x, y), blk, 0, 0);
    r.runningSum *= r.p_y == 0.f ? 0.f : p_hat / r.p_y;
    r.p_y = p_hat;
    fp.cur.p0[pixelId] = make_float4(r.runningSum, r.M, r.depth, r.p_y);
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 gridForWf(const FrameParams& fp) { return dim3((fp.W + 15) / 16, (fp.rowEnd - fp.rowBegin + 7) / 8); }

cudaError_t uploadSceneWavefront(const DScene& s, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_scene, &s, sizeof(DScene), 0, cudaMemcpyHostToDevice, st); }

int marchBlocksPerSM(int nt) {
    int n = 0;
    if (nt == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march<1, true>, 128, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march<3, true>, 128, 0);
    return n > 0 ? n : 1;
}
cudaError_t launchMarch(const WfStream& s, float* results, const MarchKind& kind, const DSlot& grid, int nt, int blocks, cudaStream_t st) {
    // FAST: trilinear sampler on a single-channel UNORM8 pool with the quad repack (every coarse / conservative mip)
    const bool fast = kind.linear && grid.format == VRESTIR_ATLAS_UNORM8 && grid.quads != nullptr;
    if (nt == 1) { if (fast) k_march<1, true><<<blocks, 128, 0, st>>>(s, results, kind, grid); else k_march<1, false><<<blocks, 128, 0, st>>>(s, results, kind, grid); }
    else { if (fast) k_march<3, true><<<blocks, 128, 0, st>>>(s, results, kind, grid); else k_march<3, false><<<blocks, 128, 0, st>>>(s, results, kind, grid); }
    return cudaGetLastError();
}
cudaError_t launchSpatialGather(const FrameParams& fp, const WfBufs& wf, cudaStream_t st) { k_spatial_gather<<<gridForWf(fp), 128, 0, st>>>(fp, wf); return cudaGetLastError(); }
cudaError_t launchInitialFinish(const FrameParams& fp, const WfBufs& wf, cudaStream_t st) { k_initial_finish<<<gridForWf(fp), 128, 0, st>>>(fp, wf); return cudaGetLastError(); }
cudaError_t launchSpatialCombine(const FrameParams& fp, const WfBufs& wf, cudaStream_t st) { k_spatial_combine<<<gridForWf(fp), 128, 0, st>>>(fp, wf); return cudaGetLastError(); }

}  // namespace vrd
