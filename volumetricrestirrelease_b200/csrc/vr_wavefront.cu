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
#include "vr_stages.cuh"
#include "vr_kernels.h"

#ifndef VR_MARCH_MINB
#define VR_MARCH_MINB 8
#endif

namespace vrd {

// ------------------------------------------------------------------------------------------------ march kernels
#ifndef VR_QUEUE_ENGINE
#define VR_QUEUE_ENGINE 1
#endif
template <int NT, bool FAST>
__global__ void __launch_bounds__(128, VR_MARCH_MINB) k_march(const WfStream s, float* results, const MarchKind kind, const DSlot g) {
#if VR_QUEUE_ENGINE
    __shared__ uint2 brickQueue[VR_QUEUE_DEPTH * 128];
    marchPoolQ<RayMarcherQ<NT, FAST>>(s.tasks, min(*s.count, s.capacity), s.cursor, results, kind, g, brickQueue);
#else
    marchPool<RayMarcher<NT, FAST>>(s.tasks, min(*s.count, s.capacity), s.cursor, results, kind, g);
#endif
}
#ifndef VR_ANALYTIC_MINB
#define VR_ANALYTIC_MINB 8
#endif
#ifndef VR_QUEUE_ANALYTIC
#define VR_QUEUE_ANALYTIC 1
#endif
__global__ void __launch_bounds__(128, VR_ANALYTIC_MINB) k_march_analytic(const WfStream s, float* results, const MarchKind kind, const DSlot g) {
#if VR_QUEUE_ANALYTIC
    __shared__ uint2 brickQueue[VR_QUEUE_DEPTH * 128];
    marchPoolQ<AnalyticMarcherQ>(s.tasks, min(*s.count, s.capacity), s.cursor, results, kind, g, brickQueue);
#else
    marchPool<AnalyticMarcher>(s.tasks, min(*s.count, s.capacity), s.cursor, results, kind, g);
#endif
}

// ------------------------------------------------------------------------------------------------ K3 gather
VRD bool tapInImage(const FrameParams& fp, int x, int y, int s, int& tx, int& ty) {
    tx = x + fp.offsets[s].x; ty = y + fp.offsets[s].y;
    return tx >= 0 && tx < fp.W && ty >= 0 && ty < fp.H;
}
VRD float3 tapRayDir(const FrameParams& fp, int tx, int ty) {
    return normalize(camRayDirNN(fp.camU, fp.camV, fp.camW, tx, ty, fp.W, fp.H));
}

// The tap loops are ROLLED (the four ray directions / tap depths live in shared memory): fully unrolled, the 12 inlined
// evaluations made the kernel instruction-fetch bound (ncu: stall_no_instruction 10.8 per issue).
__global__ void __launch_bounds__(128) k_spatial_gather(FrameParams fp, WfBufs wf) {
    __shared__ float s_dir[4][3][128];
    __shared__ float s_depth[4][128];
    int x, y;
    bool active = pixelOf(fp, x, y);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
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
    const float3 origin = fp.camPos;
    unsigned camBits = 0;     // bit j*3+k: camera ray j needs the transmittance to the depth of tap i, k = i - (i > j)
    unsigned lightBits = 0;   // bit i*4+j: light march from ray_j.at(depth_i)
    unsigned inImage = 0;     // bit s: tap s lies in the image
    // ---- pass 0: the four rays of the pixel (own ray + the primary rays of the tap pixels) and the tap depths
    if (active) {
#pragma unroll 1
        for (int s = 0; s < S; s++) {
            int tx, ty;
            if (!tapInImage(fp, x, y, s, tx, ty)) continue;
            inImage |= 1u << s;
            const float3 d = tapRayDir(fp, tx, ty);
            s_dir[s][0][tid] = d.x; s_dir[s][1][tid] = d.y; s_dir[s][2][tid] = d.z;
            s_depth[s][tid] = __ldg(&fp.cur.p0[ty * W + tx]).z;
        }
    }
    // ---- pass 1: which evaluations exist, density point queries
    if (active) {
#pragma unroll 1
        for (int i = talbot ? 0 : 1; i < S; i++) {
            if (!((inImage >> i) & 1u)) continue;
            int txi, tyi; tapInImage(fp, x, y, i, txi, tyi);
            const Reservoir tap = loadReservoir(fp.cur, tyi * W + txi, 1);
            const bool wantResample = i > 0 && tap.runningSum != 0.f;                                    // resampleNeighbor
            const bool wantMIS = talbot && (i == 0 ? tap.runningSum > 0.f : tap.runningSum != 0.f);     // superset of "runningSum > 0 after resampling"
            if (!wantResample && !wantMIS) continue;
            const bool bg = tap.depth == kRayTMax;
            bool alive = true;
#pragma unroll 1
            for (int j = 0; j < S; j++) {
                if (j == i) continue;
                if (j == 0) { if (!wantResample) continue; }
                else { if (!wantMIS || !alive) continue; if (!((inImage >> j) & 1u)) continue; }
                const float3 dir = f3(s_dir[j][0][tid], s_dir[j][1][tid], s_dir[j][2][tid]);
                const Ray r = makeRay(origin, dir, 0.f, tap.depth);
                const float3 pW = r.at(r.tMax);
                const float density = bg ? 1.f : DensityWorldSpace(pW, 0);
                blk[WF_D + i * 4 + j] = density;
                if (density != 0.f) {
                    camBits |= 1u << (j * 3 + (i - (i > j ? 1 : 0)));
                    if (!bg && tap.lightID != VRESTIR_SELF_EMISSION_LIGHT_ID) {
                        Ray sh; float3 Ld;
                        if (lightRayAndLd(makeMI(pW, -dir, true), tap.lightID, tap.lightUV, false, sh, Ld)) {
                            PreparedRay pr;   // rays that miss the volume box are resolved here (transmittance 1)
                            if (wfPrepare(sh, fp.spatial.lightingMipLevel, false, pr)) lightBits |= 1u << (i * 4 + j);
                            else blk[WF_L + i * 4 + j] = 1.f;
                        } else blk[WF_L + i * 4 + j] = 1.f;   // invalid light sample: the consumer loads the slot (wfLoadEval)
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
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        const unsigned m = (camBits >> (3 * j)) & 7u;
        const unsigned bal = __ballot_sync(FULL, m != 0);
        if (m) {
            const unsigned pos = camBase + __popc(bal & lt);
            // threshold k of ray j is the depth of tap i = k + (k >= j); unused thresholds are masked out by m
            const float t0 = s_depth[j <= 0 ? 1 : 0][tid], t1 = s_depth[j <= 1 ? 2 : 1][tid], t2 = s_depth[j <= 2 ? 3 : 2][tid];
            if (pos < wf.cam.capacity) {
                wf.cam.tasks[2 * (size_t)pos] = make_uint4(__float_as_uint(t0), __float_as_uint(t1), __float_as_uint(t2), m);
                wf.cam.tasks[2 * (size_t)pos + 1] = make_uint4(__float_as_uint(s_dir[j][0][tid]), __float_as_uint(s_dir[j][1][tid]), __float_as_uint(s_dir[j][2][tid]), blkBase + WF_C + j * 3);
            }
        }
        camBase += __popc(bal);
    }
    if (lightTot == 0) return;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const unsigned mi_ = (lightBits >> (4 * i)) & 15u;
        if (!__any_sync(FULL, mi_ != 0)) continue;
        Reservoir tap = createNewReservoir();
        if (mi_) { int txi, tyi; tapInImage(fp, x, y, i, txi, tyi); tap = loadReservoir(fp.cur, tyi * W + txi, 1); }
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            if (j == i) continue;
            const bool has = (mi_ >> j) & 1u;
            const unsigned bal = __ballot_sync(FULL, has);
            if (has) {
                const unsigned pos = lightBase + __popc(bal & lt);
                const float3 dir = f3(s_dir[j][0][tid], s_dir[j][1][tid], s_dir[j][2][tid]);
                const Ray r = makeRay(origin, dir, 0.f, tap.depth);
                const float3 pW = r.at(r.tMax);
                Ray sh; float3 Ld;
                lightRayAndLd(makeMI(pW, -dir, true), tap.lightID, tap.lightUV, false, sh, Ld);
                PreparedRay pr;
                wfPrepare(sh, fp.spatial.lightingMipLevel, false, pr);
                if (pos < wf.light.capacity) {
                    uint4* q = wf.light.tasks + 3 * (size_t)pos;
                    q[0] = make_uint4(__float_as_uint(pr.pos.x), __float_as_uint(pr.pos.y), __float_as_uint(pr.pos.z), __float_as_uint(pr.tNear));
                    q[1] = make_uint4(__float_as_uint(pr.dir.x), __float_as_uint(pr.dir.y), __float_as_uint(pr.dir.z), __float_as_uint(pr.tFar));
                    q[2] = make_uint4(blkBase + WF_L + i * 4 + j, 0u, 0u, 0u);
                }
            }
            lightBase += __popc(bal);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3 combine
// evaluate_F_ / evaluate_P_hat (VR/ReSTIRHelper.slang:91-423, B == 1) with the density at the sample point and the two
// transmittances supplied by the caller (gather kernel / march engine)
VRD float3 wfFV(const Reservoir& tap, float3 origin, float3 dir, bool isLastFrame, bool noReuse, float density, float visibility, float lightTr) {
    const vrestir_volume_desc& vd = c_scene.vol;
    const bool useLastFrameGrid = vd.usePrevGridForReproj && isLastFrame && vd.hasAnimation;
    Ray ray = makeRay(origin, dir, 0.f, tap.depth);
    const bool isBackgroundSample = tap.depth == kRayTMax;
    const bool isSelfEmission = tap.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID;
    float3 F = f3(1.f);
    const float3 p_World = ray.at(ray.tMax);
    const MediumInteraction mi = makeMI(p_World, -ray.dir, true);
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    if (density == 0.f) return f3(0.f);
    float3 sigma_s = isBackgroundSample ? f3(1.f) : (isSelfEmission ? sigA : sigS);
    if (noReuse && !isBackgroundSample) sigma_s = sigma_s / vd.sigma_t;
    F = F * (visibility * density * sigma_s);
    if (any_gt0(F)) {
        if (isBackgroundSample) F = F * envEval(ray.dir, isLastFrame);
        else if (isSelfEmission) F = F * EmissionWorldSpace(p_World, useLastFrameGrid);
        else {
            Ray sh; float3 Ld;
            const bool valid = lightRayAndLd(mi, tap.lightID, tap.lightUV, isLastFrame, sh, Ld);
            const float Tr = valid ? lightTr : 1.f;
            F = F * (Tr * Ld);
        }
    }
    return F;
}
VRD float wfPHatV(const Reservoir& tap, float3 origin, float3 dir, bool isLastFrame, float density, float visibility, float lightTr) {
    return luminance(wfFV(tap, origin, dir, isLastFrame, false, density, visibility, lightTr));
}
// The three results of one evaluation, loading only the slots its producer defined: the camera transmittance exists when the density
// is non-zero, the light transmittance when the sample is neither a background nor a self-emission sample (wfFV ignores the others)
VRD void wfLoadEval(const Reservoir& tap, const float* pd, const float* pc, const float* pl, float& density, float& vis, float& lightTr) {
    density = *pd; vis = 1.f; lightTr = 1.f;
    if (density != 0.f) {
        vis = *pc;
        if (tap.depth != kRayTMax && tap.lightID != VRESTIR_SELF_EMISSION_LIGHT_ID) lightTr = *pl;
    }
}
VRD float wfPHatE(const Reservoir& tap, float3 origin, float3 dir, bool isLastFrame, const float* e) {   // e: density, camera Tr, light Tr
    float density, vis, lightTr;
    wfLoadEval(tap, e, e + 1, e + 2, density, vis, lightTr);
    return wfPHatV(tap, origin, dir, isLastFrame, density, vis, lightTr);
}
VRD float wfPHat(const Reservoir& tap, float3 origin, float3 dir, const float* blk, int i, int j) {
    float density, vis, lightTr;
    wfLoadEval(tap, blk + WF_D + i * 4 + j, blk + WF_C + j * 3 + (i - (i > j ? 1 : 0)), blk + WF_L + i * 4 + j, density, vis, lightTr);
    return wfPHatV(tap, origin, dir, false, density, vis, lightTr);
}
// Gather side of one p-hat evaluation: density point query, then (when the sample can contribute) one camera march task
// (explicit origin, threshold = tap.depth) and one light march task.  Result slots: out+0 density, out+1 camera Tr, out+2 light Tr.
// Every lane of the warp must call; `want` masks the lanes that have an evaluation.
VRD void wfEmitEval(bool want, const Reservoir& tap, float3 origin, float3 dir, bool isLastFrame, float* results, unsigned out,
                    const WfStream& camStream, int camMip, const WfStream& lightStream, int lightMip) {
    bool hasCam = false, hasLight = false;
    Ray r = makeRay(origin, dir, 0.f, tap.depth), sh = r;
    if (want) {
        const vrestir_volume_desc& vd = c_scene.vol;
        const bool useLastFrameGrid = vd.usePrevGridForReproj && isLastFrame && vd.hasAnimation;
        const bool bg = tap.depth == kRayTMax;
        const float3 pW = r.at(r.tMax);
        const float density = bg ? 1.f : DensityWorldSpace(pW, useLastFrameGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0);
        results[out] = density;
        if (density != 0.f) {
            hasCam = true;
            if (!bg && tap.lightID != VRESTIR_SELF_EMISSION_LIGHT_ID) {
                float3 Ld;
                hasLight = lightRayAndLd(makeMI(pW, -dir, true), tap.lightID, tap.lightUV, isLastFrame, sh, Ld);
                if (!hasLight) results[out + 2] = 1.f;   // invalid light sample: no march, but the consumer loads the slot (wfLoadEval)
            }
        }
    }
    wfEmitRay(camStream, hasCam, r, camMip, false, results, out + 1);
    wfEmitRay(lightStream, hasLight, sh, lightMip, false, results, out + 2);
}

#ifndef VR_SCOMB_MINB
#define VR_SCOMB_MINB 8
#endif
__global__ void __launch_bounds__(128, VR_SCOMB_MINB) k_spatial_combine(FrameParams fp, WfBufs wf) {
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
    const float3 origin = fp.camPos;
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

// ------------------------------------------------------------------------------------------------ K1 wavefront
// VR/TraceRays.cs.slang:64-201 + VR/ComputeInitialSample.slang for B == 1 with reuse enabled and ray-marched light
// visibility.  The candidate loop is cut at the shadow march of SampleDirectLighting (VR/VolumeUtils.slang:454-492) and run
// as M + 1 lock-step kernels with one march launch between two of them:
//   k_initial_step(0)      traversal (<= 4 distance candidates) + light sample of candidate 0 -> one light-march task
//   k_march                the shadow marches of candidate s of all pixels
//   k_initial_step(s)      finishes candidate s-1 with its visibility, streams it through the reservoir (the WRS draw),
//                          then samples the light of candidate s;  the last step re-evaluates p-hat and stores.
// The random-number stream of candidate s+1 starts after the draws of candidate s, and whether candidate s draws its WRS
// number depends on its weight and hence on its march (VR/Reservoir.slang:29-32; a visibility that underflows to exactly 0
// is common in dense clouds: 40 % of the pixels of the bench frame have one), so the candidates of ONE pixel cannot be
// marched together; the candidates s of ALL pixels can.
struct K1Cand {
    unsigned flags;        // bit0 valid hit, bit1 hitEmpty, bit2 light sample valid, bit3 shadow march, bit4 speculated weight > 0
    float hd, pd, ot, density;
    float3 Li; float ph, outLightPdf;
    float3 Le; int lightID; float2 lightUV;
    float uEm, uWrs;   // uWrs unused (kept for the record layout)
};
VRD Reservoir k1Finish(const K1Cand& c, float vis, const FrameParams& fp) {
    const vrestir_volume_desc& vd = c_scene.vol;
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    Reservoir out = createNewReservoir();
    out.M = 1;
    const bool valid = c.flags & 1u;
    const float pathPdf = 1.f * c.pd;
    out.depth = valid ? c.hd : kRayTMax;
    out.p_y = pathPdf;
    if (c.flags & 2u) { out.p_y = 0.f; out.runningSum = 0.f; return out; }
    if (valid) {
        const float3 albedo = sigS / vd.sigma_t;
        out.lightID = c.lightID; out.lightUV = c.lightUV;
        float3 Ld = f3(0.f);
        if (c.flags & 4u) {
            float3 Li = c.Li;
            if (c.flags & 8u) Li = Li * vis;
            Ld = Ld + c.ph * Li / 1.f;
        }
        const float3 one_minus_albedo = f3(1.f) - albedo;
        float p_src = out.p_y;
        {
            float lumE = luminance(one_minus_albedo * c.Le);
            float emissionRatio = lumE / (lumE + luminance(albedo * Ld));
            if (isnan(emissionRatio)) emissionRatio = 0.f;
            if (c.uEm < emissionRatio) { p_src *= emissionRatio; out.lightID = VRESTIR_SELF_EMISSION_LIGHT_ID; }
            else p_src *= c.outLightPdf * (1 - emissionRatio);
        }
        out.runningSum = p_src == 0.f ? 0.f : 1.f;
        out.p_y = p_src;
        float pathPHat = 1.f;
        pathPHat *= c.ot;
        pathPHat *= c.density;
        float p_y;
        if (out.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID) p_y = pathPHat * luminance(sigA * c.Le);
        else p_y = pathPHat * luminance(sigS * Ld * c.outLightPdf);
        if (out.runningSum > 0.f) {
            out.runningSum = out.p_y == 0.f ? 0.f : p_y / out.p_y;
            out.p_y = p_y;
        }
    } else {
        float pathPHat = 1.f;
        pathPHat *= c.ot;
        const float p_y = pathPHat * luminance(c.Le);   // Le = envEval(ray.dir) for a candidate that left the volume
        out.runningSum = out.p_y == 0.f ? 0.f : p_y / out.p_y;
        out.p_y = p_y;
    }
    return out;
}
VRD void k1Store(float* r, const K1Cand& c) {
    float4* q = (float4*)r;
    q[0] = make_float4(__uint_as_float(c.flags), c.hd, c.pd, c.ot);
    q[1] = make_float4(c.density, c.Li.x, c.Li.y, c.Li.z);
    q[2] = make_float4(c.ph, c.outLightPdf, c.Le.x, c.Le.y);
    q[3] = make_float4(c.Le.z, __int_as_float(c.lightID), c.lightUV.x, c.lightUV.y);
    r[16] = c.uEm; r[17] = c.uWrs;
}
VRD K1Cand k1Load(const float* r) {
    const float4* q = (const float4*)r;
    const float4 a = q[0], b = q[1], c4 = q[2], d = q[3];
    K1Cand c;
    c.flags = __float_as_uint(a.x); c.hd = a.y; c.pd = a.z; c.ot = a.w;
    c.density = b.x; c.Li = f3(b.y, b.z, b.w);
    c.ph = c4.x; c.outLightPdf = c4.y; c.Le = f3(c4.z, c4.w, d.x);
    c.lightID = __float_as_int(d.y); c.lightUV = make_float2(d.z, d.w);
    c.uEm = r[16]; c.uWrs = r[17];
    return c;
}

// per-pixel state words (K1_STRIDE floats): [0,20) candidate record (+18 = its visibility, written by the march),
// [20,24) RNG state, [24,36) hd/pd/ot of the 4 distance candidates, [36,44) the reservoir being streamed
enum { K1_SG = 20, K1_HD = 24, K1_RES = 36 };

// The candidate traversal of K1 (SampleMediumAnalyticGeneric: <= 4 free-flight distances along the camera ray on the
// conservative mip, one random draw per voxel cell per pending sample) in a kernel of its own: with the candidate / p-hat
// code in the same kernel the hot loop missed the instruction cache (ncu: stall_no_instruction 2.9 per issue).
#ifndef VR_TRAV_MINB
#define VR_TRAV_MINB 6
#endif
__global__ void __launch_bounds__(128, VR_TRAV_MINB) k_initial_traverse(FrameParams fp, WfInitial wi) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int pixelId = y * fp.W + x;
    float* st = wi.state + (size_t)(pixelId - fp.rowBegin * fp.W) * K1_STRIDE;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount));
    const Ray ray = primaryRay(fp, x, y);
    float hds[4] = {0, 0, 0, 0}, pds[4] = {0, 0, 0, 0}, ots[4] = {0, 0, 0, 0};
    SampleMediumAnalyticGeneric(ray, sg, fp.initial.visibilityUseLinearSampler, hds, fp.initial.visibilityMipLevel, pds, ots, fp.initialM);
    ((float4*)(st + 24))[0] = make_float4(hds[0], hds[1], hds[2], hds[3]);
    ((float4*)(st + 24))[1] = make_float4(pds[0], pds[1], pds[2], pds[3]);
    ((float4*)(st + 24))[2] = make_float4(ots[0], ots[1], ots[2], ots[3]);
    ((float4*)(st + 20))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
}

// MODE 0: s == 0 (first candidate, after k_initial_traverse), 1: 0 < s < M, 2: s == M (last candidate's finish + p-hat); separate
// instantiations so that the light-weight middle steps do not carry the registers of the traversal / the p-hat marches
template <int MODE>
#ifndef VR_STEP0_MINB
#define VR_STEP0_MINB 5
#endif
#ifndef VR_STEP2_MINB
#define VR_STEP2_MINB 8
#endif
__global__ void __launch_bounds__(128, MODE == 1 ? 8 : (MODE == 0 ? VR_STEP0_MINB : VR_STEP2_MINB)) k_initial_step(FrameParams fp, WfInitial wi, int s) {
    int x, y;
    const bool inFrame = pixelOf(fp, x, y);
    const int pixelId = inFrame ? y * fp.W + x : fp.rowBegin * fp.W;
    const unsigned recBase = (unsigned)(pixelId - fp.rowBegin * fp.W) * K1_STRIDE;
    float* st = wi.state + recBase;
    const SamplingOptions& o = fp.initial;
    const int M = fp.initialM;
    bool hasTask = false, wantEval = false;
    Ray shadow = makeRay(f3(0.f), f3(0.f, 0.f, 1.f), 0.f, 0.f);
    Reservoir evalTap = createNewReservoir();
    float3 evalDir = f3(0.f, 0.f, 1.f);
    __shared__ __align__(16) float impTop[IMP_TOP_BYTES / 4];
    __shared__ uint64_t impBar;
    const bool stageImp = MODE != 2 && o.useEnvironmentLights && c_scene.haveEnv && c_scene.envSamplerType != VRESTIR_ENV_SAMPLER_ALIAS && c_scene.impDim >= IMP_TOP_DIM;
    if (stageImp) stageImportanceTop(impTop, &impBar);
    uint8_t* doneFlag = wi.done + (size_t)(pixelId - fp.rowBegin * fp.W);
    // pixels whose four distance candidates all left the volume were finished by step 0 (no light sample, no march)
    if (inFrame && (MODE == 0 || !*doneFlag)) {
        SampleGenerator sg;
        const Ray ray = primaryRay(fp, x, y);
        Reservoir finalReservoir = createNewReservoir();
        float hd = 0.f, pd = 0.f, ot = 0.f;
        bool finishNow = false;
        if (MODE == 0) {
            const float4 g4 = ((const float4*)(st + K1_SG))[0];
            sg.s0 = __float_as_uint(g4.x); sg.s1 = __float_as_uint(g4.y); sg.s2 = __float_as_uint(g4.z); sg.s3 = __float_as_uint(g4.w);
            const float4 h4 = ((const float4*)(st + K1_HD))[0], p4 = ((const float4*)(st + K1_HD))[1], o4 = ((const float4*)(st + K1_HD))[2];
            const float hds[4] = {h4.x, h4.y, h4.z, h4.w}, pds[4] = {p4.x, p4.y, p4.z, p4.w}, ots[4] = {o4.x, o4.y, o4.z, o4.w};
            hd = hds[0]; pd = pds[0]; ot = ots[0];
            bool allOut = true;
#pragma unroll
            for (int k = 0; k < 4; k++) if (k < M && hds[k] != kRayTMax) allOut = false;
            if (allOut) {
                // every candidate is a background candidate: its weight needs no march, so the whole candidate loop
                // (VR/TraceRays.cs.slang:111-175) runs here and the pixel skips the lock-step kernels
                const float3 LeBg = envEval(ray.dir);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k >= M) break;
                    K1Cand c;
                    c.flags = 0; c.hd = hds[k]; c.pd = pds[k]; c.ot = ots[k]; c.density = 0.f;
                    c.Li = f3(0.f); c.ph = 0.f; c.outLightPdf = 0.f; c.Le = LeBg; c.lightID = 0; c.lightUV = make_float2(0, 0); c.uEm = 0.f; c.uWrs = 0.f;
                    const Reservoir outReservoir = k1Finish(c, 1.f, fp);
                    simpleResampleStep<1>(outReservoir, finalReservoir, sg);
                }
                finishNow = true;
            }
            *doneFlag = allOut ? 1 : 0;
        } else {
            const float4 g4 = ((const float4*)(st + K1_SG))[0];
            sg.s0 = __float_as_uint(g4.x); sg.s1 = __float_as_uint(g4.y); sg.s2 = __float_as_uint(g4.z); sg.s3 = __float_as_uint(g4.w);
            const float4 r0 = ((const float4*)(st + K1_RES))[0], r1 = ((const float4*)(st + K1_RES))[1];
            finalReservoir.runningSum = r0.x; finalReservoir.M = r0.y; finalReservoir.depth = r0.z; finalReservoir.p_y = r0.w;
            finalReservoir.lightUV = make_float2(r1.x, r1.y); finalReservoir.lightID = __float_as_int(r1.z); finalReservoir.sampledPixel = __float_as_int(r1.w);
            // finish candidate s-1 and stream it through the reservoir
            const K1Cand c = k1Load(st);
            const float vis = (c.flags & 8u) ? st[18] : 1.f;
            const Reservoir outReservoir = k1Finish(c, vis, fp);
            simpleResampleStep<1>(outReservoir, finalReservoir, sg);
            if (MODE != 2) { hd = st[K1_HD + s]; pd = st[K1_HD + 4 + s]; ot = st[K1_HD + 8 + s]; }
        }
        if (MODE != 2 && !finishNow) {
            // candidate s up to its shadow march (VR/ComputeInitialSample.slang:60-284, bounce 0)
            K1Cand c;
            c.flags = 0; c.hd = hd; c.pd = pd; c.ot = ot; c.density = 0.f;
            c.Li = f3(0.f); c.ph = 0.f; c.outLightPdf = 0.f; c.Le = f3(0.f); c.lightID = 0; c.lightUV = make_float2(0, 0); c.uEm = 0.f; c.uWrs = 0.f;
            const bool valid = c.hd != kRayTMax;
            const MediumInteraction mi = makeMI(ray.at(c.hd), -ray.dir, valid);
            if (valid) { c.flags |= 1u; c.density = DensityWorldSpace(mi.p, 0); }
            const bool hitEmpty = valid && c.density == 0.f;
            if (hitEmpty) c.flags |= 2u;
            else if (valid) {
                c.lightID = -1;
                if (c_scene.vol.hasEmission && c.density > 0.f) c.Le = EmissionWorldSpace(mi.p);
                SceneLightSample ls;
                const bool lvalid = sampleSceneLights(mi.p, o.useEnvironmentLights, o.useAnalyticLights, o.useEmissiveLights, sg, ls, c.lightID, c.lightUV, stageImp ? impTop : nullptr);
                c.outLightPdf = lvalid ? ls.pdfArea : 0.f;
                if (lvalid) {
                    c.flags |= 4u;
                    c.Li = ls.Li;
                    c.ph = mi.phaseFunction(mi.wo, ls.dir);
                    if (o.lightSamples != 0) {
                        c.flags |= 8u; hasTask = true;
                        shadow = makeRay(mi.p, ls.rayDir, 0, ls.rayDistance);
                    }
                }
                c.uEm = sampleNext1D(sg);
            } else c.Le = envEval(ray.dir);
            k1Store(st, c);
            ((float4*)(st + K1_SG))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
            ((float4*)(st + K1_RES))[0] = make_float4(finalReservoir.runningSum, finalReservoir.M, finalReservoir.depth, finalReservoir.p_y);
            ((float4*)(st + K1_RES))[1] = make_float4(finalReservoir.lightUV.x, finalReservoir.lightUV.y, __int_as_float(finalReservoir.lightID), __int_as_float(finalReservoir.sampledPixel));
        } else if (MODE == 2 || finishNow) {
            // VR/TraceRays.cs.slang:176-183: the p-hat of the streamed reservoir under the spatial options (ray-marched: no
            // draws) becomes march tasks; k_initial_finish applies `runningSum *= p_hat / p_y`
            storeReservoir(fp.cur, pixelId, finalReservoir);
            wantEval = finalReservoir.runningSum > 0.f;
            evalTap = finalReservoir;
            evalDir = ray.dir;
        }
    }
    if (MODE != 2) wfEmitRay(wi.light, hasTask, shadow, o.lightingMipLevel, false, wi.state, recBase + 18);
    if (MODE != 1) wfEmitEval(wantEval, evalTap, fp.camPos, evalDir, false, wi.results, (unsigned)(pixelId - fp.rowBegin * fp.W) * K1_EVAL_BLOCK,
                              wi.evalCam, fp.spatial.visibilityMipLevel, wi.evalLight, fp.spatial.lightingMipLevel);
}

__global__ void __launch_bounds__(128) k_initial_finish(FrameParams fp, WfInitial wi) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int pixelId = y * fp.W + x;
    const float4 a = fp.cur.p0[pixelId];   // (runningSum, M, depth, p_y)
    if (!(a.x > 0.f)) return;
    Reservoir r = loadReservoirRW(fp.cur, pixelId, 1);
    const float* blk = wi.results + (size_t)(pixelId - fp.rowBegin * fp.W) * K1_EVAL_BLOCK;
    const float p_hat = wfPHatE(r, fp.camPos, tapRayDir(fp, x, y), false, blk);
    r.runningSum *= r.p_y == 0.f ? 0.f : p_hat / r.p_y;
    r.p_y = p_hat;
    fp.cur.p0[pixelId] = make_float4(r.runningSum, r.M, r.depth, r.p_y);
}

// ------------------------------------------------------------------------------------------------ K2 wavefront
// VR/TemporalReuse.cs.slang:80-377 for B == 1 and ray-marched p-hat: the kernel is cut at its two p-hat evaluations
// (E1: the history sample on the current ray = resampleNeighbor; E0: the current sample on the previous-frame ray = the
// Talbot MIS term).  Block slots: E0 = {0,1,2}, E1 = {3,4,5} (density, camera Tr, light Tr), 6 = state flag,
// 7/8 = reprojected pixel, 9..12 = RNG state after the reprojection-depth sampling.
enum { T2_E0 = 0, T2_E1 = 3, T2_FLAG = 6, T2_POS = 7, T2_SG = 9, K2_BLOCK = 16 };   // 13 floats used: one 64-byte block per pixel

VRD float3 prevRayDir(const FrameParams& fp, int px, int py) {
    return normalize(camRayDirNN(c_prev.prevU, c_prev.prevV, c_prev.prevW, px, py, fp.W, fp.H));
}

#ifndef VR_TGATHER_MINB
#define VR_TGATHER_MINB 5   // 96 registers without spills (103 unconstrained): the kernel is latency-bound (19 % warps active), 5 resident blocks instead of 4
#endif
#ifndef VR_TCOMB_MINB
#define VR_TCOMB_MINB 6
#endif
__global__ void __launch_bounds__(128, VR_TGATHER_MINB) k_temporal_gather(FrameParams fp, WfBufs4 wf) {
    int x, y;
    const bool inFrame = pixelOf(fp, x, y);
    const int W = fp.W, H = fp.H;
    const int pixelId = inFrame ? y * W + x : fp.rowBegin * W;
    const unsigned blkBase = (unsigned)(pixelId - fp.rowBegin * W) * K2_BLOCK;
    float* blk = wf.results + blkBase;
    bool wantE0 = false, wantE1 = false;
    Reservoir t0 = createNewReservoir(), t1 = createNewReservoir();
    float3 dirCur = f3(0.f), dirPrev = f3(0.f);
    if (inFrame) {
        SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount + 1));
        t0 = loadReservoir(fp.cur, pixelId, 1);
        const Ray ray = primaryRay(fp, x, y);
        dirCur = ray.dir;
        int2 reprojScreenPos = make_int2(0, 0);
        const int2 cf = fp.features[pixelId];
        const bool isBackgroundReservoir = __int_as_float(cf.y) == 1.f && cf.x;
        bool useFallbackReservoir = true, haveTap = false, skip = false;
        if (fp.reprojectionMode != VRESTIR_REPROJECTION_NONE) {
            float reprojDepth = t0.depth;
            if (reprojDepth == kRayTMax && fp.reprojectionMode != VRESTIR_REPROJECTION_NO_BACKGROUND && !isBackgroundReservoir)
                reprojDepth = RejectionSampleRandomPointByDensity(ray, sg, VRESTIR_NUM_MAX_MIPS + fp.reprojectionMip);
            float3 pw = ray.origin + ray.dir * reprojDepth;
            if (c_scene.vol.hasVelocity && c_scene.vol.hasAnimation) {
                float3 v = VelocityWorld(pw) * c_scene.vol.velocityScale;
                pw = pw - v;
            }
            const float* Vm = c_prev.prevView; const float* Pm = c_prev.prevProj;
            float vp[4], cp[4];
#pragma unroll
            for (int j = 0; j < 4; j++) vp[j] = pw.x * Vm[0 + j] + pw.y * Vm[4 + j] + pw.z * Vm[8 + j] + 1.f * Vm[12 + j];
#pragma unroll
            for (int j = 0; j < 4; j++) cp[j] = vp[0] * Pm[0 + j] + vp[1] * Pm[4 + j] + vp[2] * Pm[8 + j] + vp[3] * Pm[12 + j];
            float2 scrPos = make_float2(cp[0] / cp[3], cp[1] / cp[3]);
            int2 scrPosI;
            if (reprojDepth == kRayTMax) { scrPos = make_float2((float)x + 0.5f, (float)y + 0.5f); scrPosI = make_int2(x, y); }
            else {
                scrPos.x = 0.5f * scrPos.x + 0.5f; scrPos.y = -0.5f * scrPos.y + 0.5f;
                scrPos.x *= (float)W; scrPos.y *= (float)H;
                scrPosI = make_int2(f2i(scrPos.x), f2i(scrPos.y));
            }
            {
                const int id = (int)((uint32_t)scrPosI.y * (uint32_t)W + (uint32_t)scrPosI.x);
                int2 tf = make_int2(0, 0);
                if (id >= 0 && id < W * H) tf = __ldg(&fp.featuresTemporal[id]);
                const bool isTapBackgroundReservoir = __int_as_float(tf.y) == 1.f && tf.x;
                if (isBackgroundReservoir && !isTapBackgroundReservoir) skip = true;   // keeps K1's reservoir (VR/TemporalReuse.cs.slang:190-197)
            }
            if (!skip) {
                scrPosI = make_int2(f2i(scrPos.x), f2i(scrPos.y));
                reprojScreenPos = scrPosI;
                if (scrPosI.x >= 0 && scrPosI.x < W && scrPosI.y >= 0 && scrPosI.y < H) haveTap = true;
                if (haveTap) useFallbackReservoir = false;
            }
        }
        if (!skip && useFallbackReservoir) reprojScreenPos = make_int2(x, y);
        if (fp.outputMotionVec && fp.outMvec) fp.outMvec[pixelId] = make_float2((float)(reprojScreenPos.x - x) / (float)W, (float)(reprojScreenPos.y - y) / (float)H);
        blk[T2_FLAG] = skip ? 0.f : 1.f;
        if (!skip) {
            blk[T2_POS] = __int_as_float(reprojScreenPos.x); blk[T2_POS + 1] = __int_as_float(reprojScreenPos.y);
            blk[T2_SG] = __uint_as_float(sg.s0); blk[T2_SG + 1] = __uint_as_float(sg.s1); blk[T2_SG + 2] = __uint_as_float(sg.s2); blk[T2_SG + 3] = __uint_as_float(sg.s3);
            t1 = loadReservoir(fp.temporal, reprojScreenPos.y * W + reprojScreenPos.x, 1);
            dirPrev = prevRayDir(fp, reprojScreenPos.x, reprojScreenPos.y);
            if (t1.depth != kRayTMax) {
                float3 worldPos = c_prev.prevPos + t1.depth * dirPrev;
                t1.depth = length(worldPos - ray.origin);
            }
            float centerPrevFrameDepth = t0.depth;
            if (centerPrevFrameDepth != kRayTMax) { float3 worldPos = ray.at(centerPrevFrameDepth); centerPrevFrameDepth = length(worldPos - c_prev.prevPos); }
            // E1: resampleNeighbor(taps[1]) on the current ray
            if (t1.p_y > 0.f) {
                if (isnan(t1.runningSum) || isinf(t1.runningSum)) t1.runningSum = 0.f;
                wantE1 = t1.runningSum != 0.f;
            }
            // E0: Talbot term of taps[0] seen from the previous frame's ray
            if (fp.temporalMIS == VRESTIR_MIS_TALBOT && t0.p_y > 0.f) {
                if (isnan(t0.runningSum) || isinf(t0.runningSum)) t0.runningSum = 0.f;
                wantE0 = t0.runningSum > 0.f;
            }
            t0.depth = centerPrevFrameDepth;   // usedDepth of the (i = 0, j = 1) term
        }
    }
    wfEmitEval(wantE1, t1, fp.camPos, dirCur, false, wf.results, blkBase + T2_E1, wf.s[0], wf.mip[0], wf.s[1], wf.mip[1]);
    wfEmitEval(wantE0, t0, c_prev.prevPos, dirPrev, true, wf.results, blkBase + T2_E0, wf.s[2], wf.mip[2], wf.s[3], wf.mip[3]);
}

__global__ void __launch_bounds__(128, VR_TCOMB_MINB) k_temporal_combine(FrameParams fp, WfBufs4 wf) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int W = fp.W;
    const int pixelId = y * W + x;
    const float* blk = wf.results + (size_t)(pixelId - fp.rowBegin * W) * K2_BLOCK;
    if (blk[T2_FLAG] == 0.f) return;
    SampleGenerator sg;
    sg.s0 = __float_as_uint(blk[T2_SG]); sg.s1 = __float_as_uint(blk[T2_SG + 1]); sg.s2 = __float_as_uint(blk[T2_SG + 2]); sg.s3 = __float_as_uint(blk[T2_SG + 3]);
    const int2 reprojScreenPos = make_int2(__float_as_int(blk[T2_POS]), __float_as_int(blk[T2_POS + 1]));
    Reservoir taps[2];
    taps[0] = loadReservoirRW(fp.cur, pixelId, 1);
    taps[1] = loadReservoir(fp.temporal, reprojScreenPos.y * W + reprojScreenPos.x, 1);
    const Ray ray = primaryRay(fp, x, y);
    const uint32_t mis = fp.temporalMIS;
    Reservoir output = mis == VRESTIR_MIS_TALBOT ? createNewReservoir() : taps[0];
    const int numUsedReservoirs = 2;
    const float curM = taps[0].M;
    const float MaxPrevM = fp.temporalMThreshold * curM;
    const float3 dirPrev = prevRayDir(fp, reprojScreenPos.x, reprojScreenPos.y);
    if (taps[1].depth != kRayTMax) {
        float3 worldPos = c_prev.prevPos + taps[1].depth * dirPrev;
        taps[1].depth = length(worldPos - ray.origin);
    }
    float centerPrevFrameDepth = taps[0].depth;
    if (centerPrevFrameDepth != kRayTMax) { float3 worldPos = ray.at(centerPrevFrameDepth); centerPrevFrameDepth = length(worldPos - c_prev.prevPos); }
    const int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
    for (int i = startSampleId; i < numUsedReservoirs; i++) {
        float talbotMISWeight = 1.f;
        float neighbor_py = 0.f;
        if (taps[i].p_y > 0.f) {
            neighbor_py = taps[i].p_y;
            if (isnan(taps[i].runningSum) || isinf(taps[i].runningSum)) taps[i].runningSum = 0.f;
            if (i > 0 && taps[i].runningSum != 0.f) {   // resampleNeighbor
                const float p_y_hat = wfPHatE(taps[i], ray.origin, ray.dir, false, blk + T2_E1);
                float weight = p_y_hat / taps[i].p_y;
                if (isinf(weight) || isnan(weight)) weight = 0.f;
                taps[i].runningSum *= weight;
                taps[i].p_y = p_y_hat;
            }
        } else { taps[i].p_y = 0.f; taps[i].runningSum = 0.f; }
        if (mis == VRESTIR_MIS_TALBOT && taps[i].runningSum > 0.f) {
            float p_sum = 0, p_qi = 0, k = 0;
            for (int j = 0; j < numUsedReservoirs; j++) {
                const float correctedM = fminf(MaxPrevM, taps[j].M);
                k += correctedM;
                if (j == 0) { p_qi = taps[i].p_y; p_sum += taps[i].p_y * correctedM; }
                else if (i == j) { p_qi = neighbor_py; p_sum += neighbor_py * correctedM; }
                else {
                    // i == 0, j == 1: taps[0] at depth centerPrevFrameDepth on the previous frame's ray
                    Reservoir tp = taps[i]; tp.depth = centerPrevFrameDepth;
                    float p_y = wfPHatE(tp, c_prev.prevPos, dirPrev, true, blk + T2_E0);
                    if (isinf(p_y) || isnan(p_y)) p_y = 0.f;
                    p_sum += p_y * correctedM;
                }
            }
            if (p_sum > 0) talbotMISWeight = p_qi * k / p_sum;
        }
        taps[i].runningSum *= talbotMISWeight;
        simpleResampleStepWithMaxM<1>(taps[i], MaxPrevM, output, sg);
    }
    storeReservoir(fp.cur, pixelId, output);
}

// ------------------------------------------------------------------------------------------------ K5 wavefront
// VR/FinalShading.cs.slang:71-141 with analytic tracking for both transmittances (the default): the pixel's reservoir is
// shaded with the exact transmittance of the trilinear mip-0 interpolant along the camera ray and the light ray.
// Block slots: 0 density, 1 camera Tr, 2 light Tr.
__global__ void __launch_bounds__(128) k_final_gather(FrameParams fp, WfStream stream, float* results) {
    int x, y;
    const bool inFrame = pixelOf(fp, x, y);
    const int pixelId = inFrame ? y * fp.W + x : fp.rowBegin * fp.W;
    const unsigned out = (unsigned)(pixelId - fp.rowBegin * fp.W) * K5_BLOCK;
    bool hasCam = false, hasLight = false;
    Ray r = makeRay(f3(0.f), f3(0.f, 0.f, 1.f), 0.f, 0.f), sh = r;
    if (inFrame) {
        const Reservoir cur = loadReservoir(fp.cur, pixelId, 1);
        if (cur.runningSum > 0.f) {
            const bool noReuse = fp.noReuse != 0;
            const bool bg = cur.depth == kRayTMax;
            r = makeRay(fp.camPos, tapRayDir(fp, x, y), 0.f, cur.depth);
            const float3 pW = r.at(r.tMax);
            const float density = (bg || noReuse) ? 1.f : DensityWorldSpace(pW, 0);
            results[out] = density;
            if (noReuse) results[out + 1] = 1.f;   // no camera march without reuse, but the consumer loads the slot (wfLoadEval)
            if (density != 0.f) {
                hasCam = !noReuse;
                if (!bg && cur.lightID != VRESTIR_SELF_EMISSION_LIGHT_ID) {
                    float3 Ld;
                    hasLight = lightRayAndLd(makeMI(pW, -r.dir, true), cur.lightID, cur.lightUV, false, sh, Ld);
                    if (!hasLight) results[out + 2] = 1.f;   // (see wfEmitEval)
                }
            }
        }
    }
    // analytic tracking with the linear sampler traverses vertex-centred (VR/VolumeUtils.slang:284-292)
    wfEmitRay(stream, hasCam, r, 0, true, results, out + 1);
    wfEmitRay(stream, hasLight, sh, 0, true, results, out + 2);
}

__global__ void __launch_bounds__(128) k_final_combine(FrameParams fp, const float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int pixelId = y * fp.W + x;
    const float* blk = results + (size_t)(pixelId - fp.rowBegin * fp.W) * K5_BLOCK;
    float3 outputColor = f3(0.f);
    const Reservoir cur = loadReservoir(fp.cur, pixelId, 1);
    if (cur.runningSum > 0.f) {
        const bool noReuse = fp.noReuse != 0;
        float density, vis, lightTr;
        wfLoadEval(cur, blk, blk + 1, blk + 2, density, vis, lightTr);
        float3 col = wfFV(cur, fp.camPos, tapRayDir(fp, x, y), false, noReuse, density, noReuse ? 1.f : vis, lightTr);
        const float Wt = cur.p_y == 0.0f ? 1.f : cur.runningSum / (cur.p_y * cur.M);
        col = col * Wt;
        outputColor = outputColor + col;
    }
    float4 o = make_float4(outputColor.x, outputColor.y, outputColor.z, 1.f);
    if (isnan(o.x) || isinf(o.x) || isnan(o.y) || isinf(o.y) || isnan(o.z) || isinf(o.z)) o = make_float4(0.f, 0.f, 0.f, 0.f);
    fp.outColor[pixelId] = o;
}

// ------------------------------------------------------------------------------------------------ generic task-stream path
// Multi-bounce option sets (and every deterministic-tracking option set the specialised kernels above do not cover): the stage
// bodies of vr_stages.cuh — the very code of the per-pixel kernels — run twice around the march engine.
//   emit     every transmittance the evaluation asks for becomes a prepared task of the stream that holds its march
//            configuration and the evaluation continues with the placeholder 1 (a superset of the marches the real run
//            needs: a real transmittance can only cut the evaluation shorter); nothing is stored;
//   march    the persistent-lane engine (k_march / k_march_analytic) over each stream;
//   consume  the same body again, every transmittance read from the result block; stores the stage's outputs.
// Bit-identical to the per-pixel kernels by construction (same code, same operands), also for 2-4 bounces, emissive triangles
// and analytic lights.  In K3 the camera marches of one ray are still shared: slot-0 marches become thresholds of <= 4
// multi-threshold camera tasks per pixel.
struct EmitMarch {
    static constexpr bool kStore = false;
    const MarchStreams& ms; float* results; unsigned base; unsigned eval = 0; unsigned camBits = 0; bool sharedCamera;
    __device__ EmitMarch(const MarchStreams& m, float* r, unsigned b, bool shared) : ms(m), results(r), base(b), sharedCamera(shared) {}
    VRD void beginEval(int id) { eval = (unsigned)id; }
    VRD float visibility(int slot, const Ray& ray, SampleGenerator&, int, int mip, bool linear, uint32_t method, float tStepScale) {
        if (sharedCamera && slot == MARCH_SLOT_CAMERA) {   // K3: tap i seen from ray j = threshold k of camera task j
            const unsigned i = eval >> 2, j = eval & 3u;
            camBits |= 1u << (j * 3u + (i - (i > j ? 1u : 0u)));
            return 1.f;
        }
        const unsigned out = base + eval * MARCH_SLOTS + (unsigned)slot;
        const int analytic = method == VRESTIR_ANALYTIC_TRACKING ? 1 : 0;
        WfStream sel; sel.tasks = nullptr; sel.count = sel.cursor = nullptr; sel.capacity = 0;   // selected by predication: a dynamic index would spill the parameter block
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (!sel.tasks && q < ms.n && ms.mip[q] == mip && ms.linear[q] == (linear ? 1 : 0) && ms.analytic[q] == analytic && ms.scale[q] == tStepScale) sel = ms.s[q];
        if (!sel.tasks) { results[out] = __int_as_float(0x7fc00000); return 1.f; }   // no stream for this configuration (host bug): poison the result
        // analytic tracking with the linear sampler traverses vertex-centred (VR/VolumeUtils.slang:284-292)
        wfEmitRayAny(sel, ray, mip, analytic && linear, results, out);
        return 1.f;
    }
};
struct ConsumeMarch {
    static constexpr bool kStore = true;
    const float* results; unsigned base; unsigned eval = 0; bool sharedCamera;
    __device__ ConsumeMarch(const float* r, unsigned b, bool shared) : results(r), base(b), sharedCamera(shared) {}
    VRD void beginEval(int id) { eval = (unsigned)id; }
    VRD float visibility(int slot, const Ray&, SampleGenerator&, int, int, bool, uint32_t, float) {
        if (sharedCamera && slot == MARCH_SLOT_CAMERA) {
            const unsigned i = eval >> 2, j = eval & 3u;
            return results[base + MB_K3_CAM + j * 3u + (i - (i > j ? 1u : 0u))];
        }
        return results[base + eval * MARCH_SLOTS + (unsigned)slot];
    }
};

#ifndef VR_MB_MINB
#define VR_MB_MINB 4
#endif
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_temporal_emit(FrameParams fp, MarchStreams ms, float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    EmitMarch mp(ms, results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K2_STRIDE, false);
    temporalPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_temporal_consume(FrameParams fp, const float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    ConsumeMarch mp(results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K2_STRIDE, false);
    temporalPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_spatial_emit(FrameParams fp, MarchStreams ms, WfStream cam, float* results) {
    int x, y;
    const bool inFrame = pixelOf(fp, x, y);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int W = fp.W;
    const unsigned blkBase = inFrame ? (unsigned)(y * W + x - fp.rowBegin * W) * MB_K3_STRIDE : 0u;
    unsigned camBits = 0;
    if (inFrame) {
        EmitMarch mp(ms, results, blkBase, true);
        spatialPixel<B>(fp, x, y, mp);
        camBits = mp.camBits;
    }
    // the warp is converged again: camera tasks grouped by ray index, one reservation per warp (as k_spatial_gather)
    unsigned camCnt = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) camCnt += ((camBits >> (3 * j)) & 7u) ? 1u : 0u;
    const unsigned camTot = __reduce_add_sync(FULL, camCnt);
    if (camTot == 0) return;
    unsigned camBase = 0;
    if (lane == 0) camBase = atomicAdd(cam.count, camTot);
    camBase = __shfl_sync(FULL, camBase, 0);
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        const unsigned m = (camBits >> (3 * j)) & 7u;
        const unsigned bal = __ballot_sync(FULL, m != 0);
        if (m) {
            const unsigned pos = camBase + __popc(bal & lt);
            int tx, ty; tapInImage(fp, x, y, j, tx, ty);
            const float3 d = tapRayDir(fp, tx, ty);
            float thr[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 3; k++) {   // threshold k of ray j is the depth of tap i = k + (k >= j)
                if (!((m >> k) & 1u)) continue;
                const int i = k + (k >= j ? 1 : 0);
                int txi, tyi; tapInImage(fp, x, y, i, txi, tyi);
                thr[k] = __ldg(&fp.cur.p0[tyi * W + txi]).z;
            }
            if (pos < cam.capacity) {
                cam.tasks[2 * (size_t)pos] = make_uint4(__float_as_uint(thr[0]), __float_as_uint(thr[1]), __float_as_uint(thr[2]), m);
                cam.tasks[2 * (size_t)pos + 1] = make_uint4(__float_as_uint(d.x), __float_as_uint(d.y), __float_as_uint(d.z), blkBase + MB_K3_CAM + j * 3);
            }
        }
        camBase += __popc(bal);
    }
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_spatial_consume(FrameParams fp, const float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    ConsumeMarch mp(results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K3_STRIDE, true);
    spatialPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_final_emit(FrameParams fp, MarchStreams ms, float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    EmitMarch mp(ms, results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K5_STRIDE, false);
    finalPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_final_consume(FrameParams fp, const float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    ConsumeMarch mp(results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K5_STRIDE, false);
    finalPixel<B>(fp, x, y, mp);
}

// ------------------------------------------------------------------------------------------------ K1, 2-4 bounces, lock-step
// VR/TraceRays.cs.slang:64-201 + VR/ComputeInitialSample.slang:4-395 for B > 1 as a resumable per-pixel state machine: the
// candidate / bounce loop is cut at the shadow march of SampleDirectLighting, exactly where the single-bounce lock-step kernels
// above cut it.  One wave = every pixel advances to its NEXT shadow march (bounces that need none — the path left the volume,
// hit an empty voxel, drew an invalid light sample — are completed on the way), the march engine runs that wave's marches, the
// next wave consumes them.  A pixel's random-number stream is consumed in the reference's order because the pixel itself
// still runs its candidates and bounces one after the other; only the marches of DIFFERENT pixels are batched.  <= M * B waves.
// The final p-hat under the spatial options (VR/TraceRays.cs.slang:176-183) goes through the generic emit / consume passes.
struct MBPend {   // a bounce between its light sample and its shadow march
    unsigned flags;   // bit0 mi.isValid, bit1 light sample valid, bit2 shadow march pending
    float curHitDist, Tr, density;
    float3 Li; float ph, outLightPdf; float3 Le; int lightID; float2 lightUV; float3 mip;
};
VRD void mbStorePend(float* r, const MBPend& c) {
    float4* q = (float4*)r;
    q[0] = make_float4(__uint_as_float(c.flags), c.curHitDist, c.Tr, c.density);
    q[1] = make_float4(c.Li.x, c.Li.y, c.Li.z, c.ph);
    q[2] = make_float4(c.outLightPdf, c.Le.x, c.Le.y, c.Le.z);
    q[3] = make_float4(__int_as_float(c.lightID), c.lightUV.x, c.lightUV.y, c.mip.x);
    r[16] = c.mip.y; r[17] = c.mip.z;
}
VRD MBPend mbLoadPend(const float* r) {
    const float4* q = (const float4*)r;
    const float4 a = q[0], b = q[1], c4 = q[2], d = q[3];
    MBPend c;
    c.flags = __float_as_uint(a.x); c.curHitDist = a.y; c.Tr = a.z; c.density = a.w;
    c.Li = f3(b.x, b.y, b.z); c.ph = b.w; c.outLightPdf = c4.x; c.Le = f3(c4.y, c4.z, c4.w);
    c.lightID = __float_as_int(d.x); c.lightUV = make_float2(d.y, d.z); c.mip = f3(d.w, r[16], r[17]);
    return c;
}
VRD void mbStoreRes(float* r, const Reservoir& v) {
    ((float4*)r)[0] = make_float4(v.runningSum, v.M, v.depth, v.p_y);
    ((float4*)r)[1] = make_float4(v.lightUV.x, v.lightUV.y, __int_as_float(v.lightID), __int_as_float(v.sampledPixel));
}
VRD Reservoir mbLoadRes(const float* r) {
    const float4 a = ((const float4*)r)[0], b = ((const float4*)r)[1];
    Reservoir v = createNewReservoir();
    v.runningSum = a.x; v.M = a.y; v.depth = a.z; v.p_y = a.w; v.lightUV = make_float2(b.x, b.y); v.lightID = __float_as_int(b.z); v.sampledPixel = __float_as_int(b.w);
    return v;
}

#ifndef VR_MBSTEP_MINB
#define VR_MBSTEP_MINB 4
#endif
// One pixel advances to its next shadow march (or finishes).  first: the wave after k_initial_mb_traverse (the state block holds
// the RNG and the distance candidates only).  Returns whether a march was emitted (`shadow`), i.e. whether the pixel goes on.
// Suspension points: the shadow march of a bounce's light sample (march engine) and the free-flight sampling of an indirect
// bounce (k_initial_mb_bounce_traverse) — with that traversal inline the step kernel ran it at 3-4 lanes per instruction
// and missed the instruction cache (profiles/r02_ncu_full_k_initial_mb_step.txt).  Returns 0: finished, 1: shadow march
// emitted (`shadow`), 2: waiting for its bounce traversal.
template <int B>
__device__ __forceinline__ int mbAdvancePixel(const FrameParams& fp, const WfInitialMB& wi, bool first, bool active, int x, int y, unsigned local, Ray& shadow, const float* impTop) {
    const int pixelId = fp.rowBegin * fp.W + (int)local;
    const unsigned recBase = local * K1MB_STRIDE;
    float* st = wi.state + recBase;
    const SamplingOptions& options = fp.initial;
    const vrestir_volume_desc& vd = c_scene.vol;
    const int M = fp.initialM;
    bool hasTask = false, waitTrav = false;
    if (active) {
        const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
        const Ray primary = primaryRay(fp, x, y);
        SampleGenerator sg;
        { const float4 g4 = ((const float4*)(st + MBK_SG))[0]; sg.s0 = __float_as_uint(g4.x); sg.s1 = __float_as_uint(g4.y); sg.s2 = __float_as_uint(g4.z); sg.s3 = __float_as_uint(g4.w); }
        Reservoir finalReservoir, combinedReservoir;
        Ray ray = primary;
        float pathPdf = 1.f, pathPHat = 1.f, primaryScatterDepth = 0.f;
        int s = 0, bounce = 0, kind = 0;
        float3 extra[B - 1], finalExtra[B - 1];
        if (first) {
            finalReservoir = createNewReservoir(); combinedReservoir = createNewReservoir();
#pragma unroll
            for (int i = 0; i < B - 1; i++) { extra[i] = f3(0.f); finalExtra[i] = f3(0.f); }
        } else {
            finalReservoir = mbLoadRes(st + MBK_FIN); combinedReservoir = mbLoadRes(st + MBK_COMB);
            // PATH (9 floats) | CUR (2) | KIND, FINX and EXTRA (9 floats each, padded to 12): 16-byte loads
            float pc[12], fx[12], ex[12];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float4 a = ((const float4*)(st + MBK_PATH))[k], b = ((const float4*)(st + MBK_FINX))[k], c4 = ((const float4*)(st + MBK_EXTRA))[k];
                pc[4 * k] = a.x; pc[4 * k + 1] = a.y; pc[4 * k + 2] = a.z; pc[4 * k + 3] = a.w;
                fx[4 * k] = b.x; fx[4 * k + 1] = b.y; fx[4 * k + 2] = b.z; fx[4 * k + 3] = b.w;
                ex[4 * k] = c4.x; ex[4 * k + 1] = c4.y; ex[4 * k + 2] = c4.z; ex[4 * k + 3] = c4.w;
            }
            ray.origin = f3(pc[0], pc[1], pc[2]); ray.dir = f3(pc[3], pc[4], pc[5]);
            pathPdf = pc[6]; pathPHat = pc[7]; primaryScatterDepth = pc[8];
            s = __float_as_int(pc[MBK_CUR - MBK_PATH]); bounce = __float_as_int(pc[MBK_CUR + 1 - MBK_PATH]);
            kind = __float_as_int(pc[MBK_KIND - MBK_PATH]);
#pragma unroll
            for (int i = 0; i < B - 1; i++) {
                extra[i] = f3(ex[3 * i], ex[3 * i + 1], ex[3 * i + 2]);
                finalExtra[i] = f3(fx[3 * i], fx[3 * i + 1], fx[3 * i + 2]);
            }
        }
        bool resume = !first && kind == 0;      // after a shadow march
        bool resumeTrav = !first && !resume;                             // after a bounce traversal
        bool finished = false;
        MBPend c;
        for (int guard = 0; guard < 4 * B + 8 && !finished; guard++) {
            Reservoir outReservoir = createNewReservoir();
            outReservoir.M = 1;
            float vis = 1.f;
            if (!resume) {
                // ---- VR/ComputeInitialSample.slang:48-230: the bounce up to the shadow march of its light sample
                float curHitDist, pdfDist = 0.f, Tr;
                MediumInteraction mi;
                if (bounce >= 1) {
                    // SampleMediumAnalyticGeneric(ray, sg, ..., 1 sample) runs in k_initial_mb_bounce_traverse on the stored ray and RNG state
                    if (!resumeTrav) { waitTrav = true; shadow = ray; break; }
                    resumeTrav = false;
                    curHitDist = st[MBK_TRAV]; pdfDist = st[MBK_TRAV + 1]; Tr = st[MBK_TRAV + 2];
                } else {
                    curHitDist = st[MBK_HD + s]; pdfDist = st[MBK_HD + 4 + s]; Tr = st[MBK_HD + 8 + s];
                }
                mi = makeMI(ray.at(curHitDist), -ray.dir, curHitDist != kRayTMax);
                pathPdf *= pdfDist;
                if (bounce == options.vertexReuseStartBounce && curHitDist != kRayTMax) {   // VERTEX_REUSE :88-94
                    pathPdf /= curHitDist * curHitDist;
                    pathPHat /= curHitDist * curHitDist;
                }
                if (bounce == 0) primaryScatterDepth = mi.isValid ? curHitDist : kRayTMax;
                else if (bounce < options.vertexReuseStartBounce) extra[bounce - 1] = encodeWiDist(make_float4(ray.dir.x, ray.dir.y, ray.dir.z, !mi.isValid ? kRayTMax : curHitDist));
                else extra[bounce - 1] = !mi.isValid ? f3(kRayTMax) : mi.p;   // VERTEX_REUSE :116-125
                c.flags = mi.isValid ? 1u : 0u; c.curHitDist = curHitDist; c.Tr = Tr; c.density = 0.f;
                c.Li = f3(0.f); c.ph = 0.f; c.outLightPdf = 0.f; c.Le = f3(0.f); c.lightID = 0; c.lightUV = make_float2(0, 0); c.mip = mi.p;
                if (mi.isValid) c.density = DensityWorldSpace(mi.p, 0);
                const bool hitEmpty0 = (!mi.isValid && bounce > 0) || (mi.isValid && c.density == 0.f);
                if (!hitEmpty0 && mi.isValid) {
                    c.lightID = -1;
                    if (vd.hasEmission && c.density > 0.f) c.Le = EmissionWorldSpace(mi.p);
                    SceneLightSample ls;
                    const bool lvalid = sampleSceneLights(mi.p, options.useEnvironmentLights, options.useAnalyticLights, options.useEmissiveLights, sg, ls, c.lightID, c.lightUV, impTop);
                    c.outLightPdf = lvalid ? ls.pdfArea : 0.f;
                    if (lvalid) {
                        c.flags |= 2u;
                        c.Li = ls.Li;
                        c.ph = mi.phaseFunction(mi.wo, ls.dir);
                        if (options.lightSamples != 0) {
                            c.flags |= 4u;
                            hasTask = true;
                            shadow = makeRay(mi.p, ls.rayDir, 0, ls.rayDistance);
                            break;   // the wave ends here for this pixel: state is stored below, the march engine takes over
                        }
                    }
                }
            } else {
                c = mbLoadPend(st + MBK_PEND);
                vis = st[MBK_VIS];
                resume = false;
            }
            // ---- VR/ComputeInitialSample.slang:76-393: the rest of the bounce
            const bool valid = c.flags & 1u;
            const float curHitDist = c.curHitDist;
            bool hitEmpty = false;
            if (bounce == 0) {
                outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, 0);
                outReservoir.depth = valid ? curHitDist : kRayTMax;
                outReservoir.p_y = pathPdf;
            } else {
                outReservoir.depth = primaryScatterDepth;
                outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, bounce);
                outReservoir.p_y = pathPdf;
            }
            const float actualVolumeDensity = valid ? c.density : 0.f;
            if ((!valid && bounce > 0) || (valid && actualVolumeDensity == 0)) { outReservoir.p_y = 0.f; outReservoir.runningSum = 0.f; hitEmpty = true; }
            float pdfDir = 1.f;
            float3 wi_ = f3(0.f);
            if (!hitEmpty) {
                if (valid) {
                    const float3 albedo = sigS / vd.sigma_t;
                    outReservoir.lightID = c.lightID;
                    outReservoir.lightUV = c.lightUV;
                    const float outLightPdf = c.outLightPdf;
                    float3 Ld = f3(0.f);
                    const float3 Le = c.Le;
                    const float3 one_minus_albedo = f3(1.f) - albedo;
                    if (c.flags & 2u) {   // SampleDirectLighting (VR/VolumeUtils.slang:454-492) with the marched visibility
                        float3 Li = c.Li;
                        if (c.flags & 4u) Li = Li * vis;
                        Ld = Ld + c.ph * Li / 1.f;
                    }
                    const MediumInteraction mi = makeMI(c.mip, -ray.dir, true);
                    const float3 wo = -ray.dir;
                    pdfDir = mi.Sample_p(wo, wi_, sampleNext2D(sg));
                    float p_src = outReservoir.p_y;
                    {
                        float lumE = luminance(one_minus_albedo * Le);
                        float emissionRatio = lumE / (lumE + luminance(albedo * Ld));
                        if (isnan(emissionRatio)) emissionRatio = 0.f;
                        if (sampleNext1D(sg) < emissionRatio) { p_src *= emissionRatio; outReservoir.lightID = VRESTIR_SELF_EMISSION_LIGHT_ID; }
                        else p_src *= outLightPdf * (1 - emissionRatio);
                    }
                    outReservoir.runningSum = p_src == 0.f ? 0.f : 1.f;
                    outReservoir.p_y = p_src;
                    {
                        float p_y;
                        pathPHat *= c.Tr;
                        pathPHat *= actualVolumeDensity;
                        if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID) p_y = pathPHat * luminance(sigA * Le);
                        else p_y = pathPHat * luminance(sigS * Ld * outLightPdf);
                        pathPHat *= luminance(sigS) * pdfDir;
                        if (outReservoir.runningSum > 0.f) {
                            outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                            if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID && bounce > 0) {
                                if (bounce < options.vertexReuseStartBounce) {   // VERTEX_REUSE :324-331
                                    encodeEmissivePosition(c.mip, outReservoir.lightID, outReservoir.lightUV);
                                    p_y /= (curHitDist * curHitDist);
                                }
                                outReservoir.sampledPixel = encodePathTag(outReservoir.sampledPixel, 1);
                            }
                            outReservoir.p_y = p_y;
                        }
                    }
                    pathPdf *= pdfDir;
                    if (bounce < B - 1) {
                        ray = makeRay(c.mip, wi_, 0, kRayTMax);
                        if (fp.useRussianRoulette && bounce >= 2) {
                            if (sampleNext1D(sg) < albedo.x) pathPdf *= albedo.x;
                            else { hitEmpty = true; combinedReservoir.M++; }
                        }
                    }
                } else {
                    const float3 Le = envEval(ray.dir);
                    pathPHat *= c.Tr;
                    const float p_y = pathPHat * luminance(Le);
                    outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                    outReservoir.p_y = p_y;
                    hitEmpty = true;
                }
            }
            simpleResampleStep<B>(outReservoir, combinedReservoir, sg);
            if (hitEmpty || bounce == B - 1) {
                // ---- the candidate is complete (VR/ComputeInitialSample.slang:395, VR/TraceRays.cs.slang:153-172)
                combinedReservoir.M = 1;
                const bool isSelected = simpleResampleStep<B>(combinedReservoir, finalReservoir, sg);
                if (isSelected) {
                    const int mib = decodeMaxIndirectBounces<B>(finalReservoir.sampledPixel);
#pragma unroll
                    for (int b = 0; b < B - 1; b++) if (b < mib) finalExtra[b] = extra[b];
                }
                s++; bounce = 0;
                ray = primary; pathPdf = 1.f; pathPHat = 1.f; primaryScatterDepth = 0.f;
                combinedReservoir = createNewReservoir();
                if (s >= M) finished = true;
            } else bounce++;
        }
        if (finished) {
            // the streamed reservoir; its p-hat under the spatial options is evaluated by the emit / consume passes that follow
            storeReservoir(fp.cur, pixelId, finalReservoir);
            const int mib = decodeMaxIndirectBounces<B>(finalReservoir.sampledPixel);
#pragma unroll
            for (int b = 0; b < B - 1; b++) if (b < mib) fp.extCur[(size_t)pixelId * (B - 1) + b] = finalExtra[b];
        } else {
            ((float4*)(st + MBK_SG))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
            mbStoreRes(st + MBK_FIN, finalReservoir); mbStoreRes(st + MBK_COMB, combinedReservoir);
            ((float4*)(st + MBK_PATH))[0] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x);
            ((float4*)(st + MBK_PATH))[1] = make_float4(ray.dir.y, ray.dir.z, pathPdf, pathPHat);
            ((float4*)(st + MBK_PATH))[2] = make_float4(primaryScatterDepth, __int_as_float(s), __int_as_float(bounce), __int_as_float(waitTrav ? 1 : 0));
            float fx[12], ex[12];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const bool have = i < B - 1;
                ex[3 * i] = have ? extra[i < B - 1 ? i : 0].x : 0.f; ex[3 * i + 1] = have ? extra[i < B - 1 ? i : 0].y : 0.f; ex[3 * i + 2] = have ? extra[i < B - 1 ? i : 0].z : 0.f;
                fx[3 * i] = have ? finalExtra[i < B - 1 ? i : 0].x : 0.f; fx[3 * i + 1] = have ? finalExtra[i < B - 1 ? i : 0].y : 0.f; fx[3 * i + 2] = have ? finalExtra[i < B - 1 ? i : 0].z : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                ((float4*)(st + MBK_EXTRA))[k] = make_float4(ex[4 * k], ex[4 * k + 1], ex[4 * k + 2], ex[4 * k + 3]);
                ((float4*)(st + MBK_FINX))[k] = make_float4(fx[4 * k], fx[4 * k + 1], fx[4 * k + 2], fx[4 * k + 3]);
            }
            if (!waitTrav) mbStorePend(st + MBK_PEND, c);
        }
    }
    return hasTask ? 1 : (waitTrav ? 2 : 0);
}
// the mip an indirect bounce samples its free-flight distance on (VR/ComputeInitialSample.slang:60-66)
VRD int mbBounceMip(const FrameParams& fp) {
    int curMip = fp.initial.visibilityMipLevel;
    if (fp.useCoarserGrid) curMip = min((fp.initial.visibilityMipLevel >= VRESTIR_NUM_MAX_MIPS ? VRESTIR_NUM_MAX_MIPS : 0) + c_scene.vol.numMips - 1, curMip + 1);
    return curMip;
}
// append the pixels waiting for a bounce traversal to the wave's list (one atomic per warp; every lane of the warp calls)
VRD void mbEmitTrav(const WfInitialMB& wi, bool want, unsigned local) {
    const unsigned bal = __ballot_sync(0xffffffffu, want);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == __ffs(bal) - 1) base = atomicAdd(wi.travCount, (unsigned)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (want) wi.travList[base + __popc(bal & ((1u << lane) - 1u))] = local;
}
// A pixel that waits for its bounce traversal: entry in the wave's list (the next step walks it) and, with the point sampler, the
// free-flight sampling as a prepared task of the march engine (DistanceMarcher; a ray that misses the volume box is final here:
// hit distance kRayTMax, pdf 1, transmittance 1 — SampleMediumAnalyticAdapter::ExecuteEndStep before ExecuteStartStep)
VRD void mbEmitWaiting(const FrameParams& fp, const WfInitialMB& wi, bool want, unsigned local, const Ray& bounceRay) {
    mbEmitTrav(wi, want, local);
    if (fp.initial.visibilityUseLinearSampler) return;   // k_initial_mb_bounce_traverse walks the list instead
    if (want) { wi.state[local * K1MB_STRIDE + MBK_TRAV + 1] = 1.f; wi.state[local * K1MB_STRIDE + MBK_TRAV + 2] = 1.f; }
    wfEmitRay(wi.trav, want, bounceRay, mbBounceMip(fp), false, wi.state, local * K1MB_STRIDE + MBK_TRAV, kRayTMax);
}
// first != 0: one thread per pixel of the band (8x4 tiles); afterwards: grid-stride over the previous wave's task list
template <int B>
__global__ void __launch_bounds__(128, VR_MBSTEP_MINB) k_initial_mb_step(FrameParams fp, WfInitialMB wi, int first) {
    // top levels of the importance map in shared memory (one bulk copy per CTA): the hierarchical descent of every env-light sample
    // starts with IMP_TOP levels of dependent loads that otherwise each wait for L2
    __shared__ __align__(16) float impTopS[IMP_TOP_BYTES / 4];
    __shared__ uint64_t impBar;
    const bool stageImp = fp.initial.useEnvironmentLights && c_scene.haveEnv && c_scene.envSamplerType != VRESTIR_ENV_SAMPLER_ALIAS && c_scene.impDim >= IMP_TOP_DIM;
    if (stageImp) stageImportanceTop(impTopS, &impBar);
    const float* impTop = stageImp ? impTopS : nullptr;
    if (first) {
        int x, y;
        const bool inFrame = pixelOf(fp, x, y);
        const unsigned local = inFrame ? (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) : 0u;
        Ray shadow = makeRay(f3(0.f), f3(0.f, 0.f, 1.f), 0.f, 0.f);
        const int r = mbAdvancePixel<B>(fp, wi, true, inFrame, x, y, local, shadow, impTop);
        wfEmitRay(wi.light, r == 1, shadow, fp.initial.lightingMipLevel, false, wi.state, local * K1MB_STRIDE + MBK_VIS);
        mbEmitWaiting(fp, wi, r == 2, local, shadow);
        return;
    }
    // the pixels still running = the previous wave's march tasks + its traversal list
    const unsigned nMarch = min(*wi.prev.count, wi.prev.capacity), total = nMarch + *wi.prevTravCount;
    for (unsigned t0 = blockIdx.x * blockDim.x; t0 < total; t0 += gridDim.x * blockDim.x) {   // block-uniform trip count: the emits are warp-collective
        const unsigned t = t0 + threadIdx.x;
        const bool active = t < total;
        unsigned local = 0;
        if (active) local = t < nMarch ? __ldg(&wi.prev.tasks[3 * (size_t)t + 2]).x / K1MB_STRIDE   // the task's result slot identifies its pixel
                                       : __ldg(&wi.prevTravList[t - nMarch]);
        const int pixelId = fp.rowBegin * fp.W + (int)local;
        Ray shadow = makeRay(f3(0.f), f3(0.f, 0.f, 1.f), 0.f, 0.f);
        const int r = mbAdvancePixel<B>(fp, wi, false, active, pixelId % fp.W, pixelId / fp.W, local, shadow, impTop);
        wfEmitRay(wi.light, r == 1, shadow, fp.initial.lightingMipLevel, false, wi.state, local * K1MB_STRIDE + MBK_VIS);
        mbEmitWaiting(fp, wi, r == 2, local, shadow);
    }
}

// K1's primary free-flight sampling on the decoupled engine (point sampler): implicit tasks, one per pixel of the band
#ifndef VR_PRIMARY_MINB
#define VR_PRIMARY_MINB 6
#endif
__global__ void __launch_bounds__(128, VR_PRIMARY_MINB) k_march_primary_distance(const PrimaryDistanceCtx c, unsigned* cursor, const MarchKind kind, const DSlot g) {
    __shared__ uint2 brickQueue[VR_QUEUE_DEPTH * 128];
    const unsigned tilesX = (unsigned)(c.fp.W + 7) / 8u, tilesY = (unsigned)(c.fp.rowEnd - c.fp.rowBegin + 3) / 4u;
    marchPoolQ<DistanceMarcherQ, PrimaryDistanceCtx>(nullptr, tilesX * tilesY * 32u, cursor, c.state, kind, g, brickQueue, &c);
}

// the engine form of the same (point sampler): one DistanceMarcher task per waiting pixel
__global__ void __launch_bounds__(128, VR_ANALYTIC_MINB) k_march_distance(const WfStream s, float* state, const MarchKind kind, const DSlot g) {
    marchPool<DistanceMarcher>(s.tasks, min(*s.count, s.capacity), s.cursor, state, kind, g);
}

// Free-flight sampling of one indirect bounce per listed pixel (VR/ComputeInitialSample.slang:60-72: one analytic-tracking sample
// along the scattered ray on the visibility mip or its coarser neighbour), with the pixel's own RNG stream: ray and generator
// come from the state block, hit distance / pdf / transmittance and the advanced generator go back into it.
__global__ void __launch_bounds__(128, VR_TRAV_MINB) k_initial_mb_bounce_traverse(FrameParams fp, WfInitialMB wi) {
    const unsigned total = *wi.prevTravCount;
    const SamplingOptions& options = fp.initial;
    const int curMip = mbBounceMip(fp);
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        float* st = wi.state + (size_t)__ldg(&wi.prevTravList[t]) * K1MB_STRIDE;
        SampleGenerator sg;
        { const float4 g4 = ((const float4*)(st + MBK_SG))[0]; sg.s0 = __float_as_uint(g4.x); sg.s1 = __float_as_uint(g4.y); sg.s2 = __float_as_uint(g4.z); sg.s3 = __float_as_uint(g4.w); }
        const Ray ray = makeRay(f3(st[MBK_PATH], st[MBK_PATH + 1], st[MBK_PATH + 2]), f3(st[MBK_PATH + 3], st[MBK_PATH + 4], st[MBK_PATH + 5]), 0, kRayTMax);
        float hd[4], pd[4], ot[4];
        SampleMediumAnalyticGeneric(ray, sg, options.visibilityUseLinearSampler, hd, curMip, pd, ot, 1);
        st[MBK_TRAV] = hd[0]; st[MBK_TRAV + 1] = pd[0]; st[MBK_TRAV + 2] = ot[0];
        ((float4*)(st + MBK_SG))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
    }
}


// distance candidates of the primary ray into the multi-bounce state block (same traversal as k_initial_traverse)
__global__ void __launch_bounds__(128, VR_TRAV_MINB) k_initial_mb_traverse(FrameParams fp, WfInitialMB wi) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    const int pixelId = y * fp.W + x;
    float* st = wi.state + (size_t)(pixelId - fp.rowBegin * fp.W) * K1MB_STRIDE;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount));
    const Ray ray = primaryRay(fp, x, y);
    float hds[4] = {0, 0, 0, 0}, pds[4] = {0, 0, 0, 0}, ots[4] = {0, 0, 0, 0};
    SampleMediumAnalyticGeneric(ray, sg, fp.initial.visibilityUseLinearSampler, hds, fp.initial.visibilityMipLevel, pds, ots, fp.initialM);
    ((float4*)(st + MBK_HD))[0] = make_float4(hds[0], hds[1], hds[2], hds[3]);
    ((float4*)(st + MBK_HD))[1] = make_float4(pds[0], pds[1], pds[2], pds[3]);
    ((float4*)(st + MBK_HD))[2] = make_float4(ots[0], ots[1], ots[2], ots[3]);
    ((float4*)(st + MBK_SG))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
}

// VR/TraceRays.cs.slang:176-183 for the streamed reservoir of a pixel: p-hat under the spatial options, then
// runningSum *= p_hat / p_y.  Run as an emit pass and a consume pass (march provider) like the reuse stages.
template <int B, class MP>
__device__ __forceinline__ void initialFinishPixel(const FrameParams& fp, int x, int y, MP& mp) {
    const int pixelId = y * fp.W + x;
    Reservoir r = loadReservoirRW(fp.cur, pixelId, B);
    // the reference evaluates p-hat unconditionally and uses it only when runningSum > 0; ray-marched p-hat draws no random numbers
    // (under VERTEX_REUSE the evaluation also leaves p_partial behind, so it is not skipped)
    if (!(r.runningSum > 0.f) && !fp.cur.p2) return;
    ExtraProviderRW prov; prov.global = fp.extCur;
    const Ray ray = primaryRay(fp, x, y);
    mp.beginEval(0);
    SampleGenerator none; none.s0 = none.s1 = none.s2 = none.s3 = 0;   // deterministic tracking: never drawn from
    const float p_hat = evaluate_P_hat<B>(ray, none, prov, fp.spatial, r, false, false, mp);
    if (!MP::kStore) return;
    if (r.runningSum > 0.f) {
        r.runningSum *= r.p_y == 0.f ? 0.f : p_hat / r.p_y;
        r.p_y = p_hat;
        fp.cur.p0[pixelId] = make_float4(r.runningSum, r.M, r.depth, r.p_y);
    }
    if (fp.cur.p2) fp.cur.p2[pixelId] = r.p_partial;
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_initial_finish_emit(FrameParams fp, MarchStreams ms, float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    EmitMarch mp(ms, results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K1_EVAL_STRIDE, false);
    initialFinishPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MB_MINB) k_initial_finish_consume(FrameParams fp, const float* results) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    ConsumeMarch mp(results, (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * MB_K1_EVAL_STRIDE, false);
    initialFinishPixel<B>(fp, x, y, mp);
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 gridForWf(const FrameParams& fp) { return dim3((fp.W + 15) / 16, (fp.rowEnd - fp.rowBegin + 7) / 8); }
// one warp (8x4 tile) per CTA
static dim3 gridForWarp(const FrameParams& fp) { return dim3((fp.W + 7) / 8, (fp.rowEnd - fp.rowBegin + 3) / 4); }
#ifndef VR_TRAV_BLOCK
#define VR_TRAV_BLOCK 128
#endif
#ifndef VR_TGATHER_BLOCK
#define VR_TGATHER_BLOCK 128
#endif

cudaError_t uploadSceneWavefront(const DScene& s, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_scene, &s, sizeof(DScene), 0, cudaMemcpyHostToDevice, st); }
cudaError_t uploadPrevCamWavefront(const DPrevCam& s, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_prev, &s, sizeof(DPrevCam), 0, cudaMemcpyHostToDevice, st); }

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
int analyticBlocksPerSM() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march_analytic, 128, 0); return n > 0 ? n : 1; }
cudaError_t launchMarchAnalytic(const WfStream& s, float* results, const MarchKind& kind, const DSlot& grid, int blocks, cudaStream_t st) { k_march_analytic<<<blocks, 128, 0, st>>>(s, results, kind, grid); return cudaGetLastError(); }
cudaError_t launchFinalGather(const FrameParams& fp, const WfStream& s, float* results, cudaStream_t st) { k_final_gather<<<gridForWf(fp), 128, 0, st>>>(fp, s, results); return cudaGetLastError(); }
cudaError_t launchFinalCombine(const FrameParams& fp, const float* results, cudaStream_t st) { k_final_combine<<<gridForWf(fp), 128, 0, st>>>(fp, results); return cudaGetLastError(); }
cudaError_t launchInitialStep(const FrameParams& fp, const WfInitial& wi, int s, cudaStream_t st) {
    if (s == 0) { k_initial_traverse<<<VR_TRAV_BLOCK == 32 ? gridForWarp(fp) : gridForWf(fp), VR_TRAV_BLOCK, 0, st>>>(fp, wi); k_initial_step<0><<<gridForWf(fp), 128, 0, st>>>(fp, wi, s); }
    else if (s < fp.initialM) k_initial_step<1><<<gridForWf(fp), 128, 0, st>>>(fp, wi, s);
    else k_initial_step<2><<<gridForWf(fp), 128, 0, st>>>(fp, wi, s);
    return cudaGetLastError();
}
// step 0 without the per-pixel traversal kernel (the primary distances came from launchPrimaryDistance)
cudaError_t launchInitialStepOnly(const FrameParams& fp, const WfInitial& wi, int s, cudaStream_t st) { k_initial_step<0><<<gridForWf(fp), 128, 0, st>>>(fp, wi, s); return cudaGetLastError(); }
cudaError_t launchInitialFinish(const FrameParams& fp, const WfInitial& wi, cudaStream_t st) { k_initial_finish<<<gridForWf(fp), 128, 0, st>>>(fp, wi); return cudaGetLastError(); }
cudaError_t launchTemporalGather(const FrameParams& fp, const WfBufs4& wf, cudaStream_t st) { k_temporal_gather<<<VR_TGATHER_BLOCK == 32 ? gridForWarp(fp) : gridForWf(fp), VR_TGATHER_BLOCK, 0, st>>>(fp, wf); return cudaGetLastError(); }
cudaError_t launchTemporalCombine(const FrameParams& fp, const WfBufs4& wf, cudaStream_t st) { k_temporal_combine<<<gridForWf(fp), 128, 0, st>>>(fp, wf); return cudaGetLastError(); }
cudaError_t launchSpatialCombine(const FrameParams& fp, const WfBufs& wf, cudaStream_t st) { k_spatial_combine<<<gridForWf(fp), 128, 0, st>>>(fp, wf); return cudaGetLastError(); }

#define VR_DISPATCH_MB(kern, fp, st, ...)                                                        \
    switch ((fp).maxBounces) {                                                                   \
        case 1: kern<1><<<gridForWf(fp), 128, 0, st>>>(__VA_ARGS__); break;                      \
        case 2: kern<2><<<gridForWf(fp), 128, 0, st>>>(__VA_ARGS__); break;                      \
        case 3: kern<3><<<gridForWf(fp), 128, 0, st>>>(__VA_ARGS__); break;                      \
        default: kern<4><<<gridForWf(fp), 128, 0, st>>>(__VA_ARGS__); break;                     \
    }
cudaError_t launchInitialMBTraverse(const FrameParams& fp, const WfInitialMB& wi, cudaStream_t st) { k_initial_mb_traverse<<<gridForWf(fp), 128, 0, st>>>(fp, wi); return cudaGetLastError(); }
int primaryDistanceBlocksPerSM() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march_primary_distance, 128, 0); return n > 0 ? n : 1; }
// state: per-pixel blocks of `stride` floats; the 12 result floats go to `hdOffset`, the generator 4 floats before them
cudaError_t launchPrimaryDistance(const FrameParams& fp, float* state, unsigned stride, unsigned hdOffset, unsigned* cursor, const DSlot& grid, int blocks, cudaStream_t st) {
    PrimaryDistanceCtx c; c.fp = fp; c.state = state; c.stride = stride; c.hdOffset = hdOffset; c.numSamples = fp.initialM;
    const MarchKind kind = {fp.initial.visibilityMipLevel, 0, 1.f, 0, {0.f, 0.f, 0.f}};
    k_march_primary_distance<<<blocks, 128, 0, st>>>(c, cursor, kind, grid);
    return cudaGetLastError();
}
int distanceBlocksPerSM() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march_distance, 128, 0); return n > 0 ? n : 1; }
cudaError_t launchMarchDistance(const WfStream& s, float* state, const MarchKind& kind, const DSlot& grid, int blocks, cudaStream_t st) { k_march_distance<<<blocks, 128, 0, st>>>(s, state, kind, grid); return cudaGetLastError(); }
int initialMBBounceTraverseBlocksPerSM() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_initial_mb_bounce_traverse, 128, 0); return n > 0 ? n : 1; }
cudaError_t launchInitialMBBounceTraverse(const FrameParams& fp, const WfInitialMB& wi, int blocks, cudaStream_t st) {
    k_initial_mb_bounce_traverse<<<blocks, 128, 0, st>>>(fp, wi); return cudaGetLastError();
}
int initialMBStepBlocksPerSM() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_initial_mb_step<4>, 128, 0); return n > 0 ? n : 1; }
// first: one thread per pixel of the band; afterwards `blocks` CTAs stride over the previous wave's task list
cudaError_t launchInitialMBStep(const FrameParams& fp, const WfInitialMB& wi, int first, int blocks, cudaStream_t st) {
    const dim3 grid = first ? gridForWf(fp) : dim3((unsigned)blocks);
    switch (fp.maxBounces) {
        case 2: k_initial_mb_step<2><<<grid, 128, 0, st>>>(fp, wi, first); break;
        case 3: k_initial_mb_step<3><<<grid, 128, 0, st>>>(fp, wi, first); break;
        default: k_initial_mb_step<4><<<grid, 128, 0, st>>>(fp, wi, first); break;
    }
    return cudaGetLastError();
}
// stage: 1 = K1's final p-hat, 2 temporal, 3 spatial, 5 final
cudaError_t launchStageEmit(int stage, const FrameParams& fp, const MarchStreams& ms, const WfStream& cam, float* results, cudaStream_t st) {
    if (stage == 1) { VR_DISPATCH_MB(k_initial_finish_emit, fp, st, fp, ms, results); }
    else if (stage == 2) { VR_DISPATCH_MB(k_temporal_emit, fp, st, fp, ms, results); }
    else if (stage == 3) { VR_DISPATCH_MB(k_spatial_emit, fp, st, fp, ms, cam, results); }
    else { VR_DISPATCH_MB(k_final_emit, fp, st, fp, ms, results); }
    return cudaGetLastError();
}
cudaError_t launchStageConsume(int stage, const FrameParams& fp, const float* results, cudaStream_t st) {
    if (stage == 1) { VR_DISPATCH_MB(k_initial_finish_consume, fp, st, fp, results); }
    else if (stage == 2) { VR_DISPATCH_MB(k_temporal_consume, fp, st, fp, results); }
    else if (stage == 3) { VR_DISPATCH_MB(k_spatial_consume, fp, st, fp, results); }
    else { VR_DISPATCH_MB(k_final_consume, fp, st, fp, results); }
    return cudaGetLastError();
}

}  // namespace vrd
