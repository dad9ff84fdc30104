// vr_kernels.cu — the per-frame kernels of the VolumetricReSTIR pass for sm_100a (K0..K6 of SURVEY.md section 2.2).
//
//   k_features   K0  VR/GenerateFeatures.cs.slang:57-102
//   k_initial    K1  VR/TraceRays.cs.slang:64-201 (+ VR/ComputeInitialSample.slang, VR/VolumePathTracingFunctions.slang)
//   k_temporal   K2  VR/TemporalReuse.cs.slang:80-377
//   k_spatial    K3  VR/SpatialReuse.cs.slang:94-265
//   (K4 CopyReservoirs is a buffer rotation on the host: zero bytes moved, see vr_pass.cu)
//   k_final      K5  VR/FinalShading.cs.slang:71-141
//   k_importance K6  F/Experimental/Scene/Lights/EnvMapSamplerSetup.cs.slang:48-75 + mip chain
//
// Thread mapping: one thread per pixel; a warp covers an 8x4 pixel tile (coherent brick paths), a CTA of 4 warps covers
// 16x8 pixels.  Reservoirs are SoA float4 planes: every load/store is a coalesced 16-byte access.
#include "vr_stages.cuh"
#include "vr_kernels.h"

#ifndef VR_MINB
#define VR_MINB 4
#endif

namespace vrd {

// ------------------------------------------------------------------------------------------------ K0
__global__ void __launch_bounds__(128, VR_MINB) k_features(FrameParams fp) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    Ray ray = primaryRay(fp, x, y);
    ReservoirFeatureRayMarchingAdapter a;
    a.Init(true, c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * fp.initial.visibilityTStepScale);
    SampleGenerator sg; sg.s0 = sg.s1 = sg.s2 = sg.s3 = 0;
    VolumeTrackingGVDB(ray, 0, sg, a, false);
    fp.features[y * fp.W + x] = make_int2(1, __float_as_int(a.accuTransmittance));
}

// ------------------------------------------------------------------------------------------------ K1
// VR/ComputeInitialSample.slang:4-395
template <int B>
__device__ Reservoir ComputeInitialSample(const Ray& primaryRay_, float precomputedHitDistance, float precomputedPdfDist, float precomputedTr, SampleGenerator& sg,
                                          const FrameParams& fp, float3* extrabounceReservoir) {
    const vrestir_volume_desc& vd = c_scene.vol;
    const SamplingOptions& options = fp.initial;
    const bool noReuse = fp.noReuse != 0;
    const float3 sigA = v3(vd.sigma_a), sigS = v3(vd.sigma_s);
    float pathPdf = 1.f, pathPHat = 1.f;
    Ray ray = primaryRay_;
    Reservoir combinedReservoir = createNewReservoir();
    Reservoir outReservoir = createNewReservoir();
    float primaryScatterDepth = 0;
    for (int bounce = 0; bounce < B; bounce++) {
        outReservoir = createNewReservoir();
        outReservoir.M = 1;
        float curHitDist; float pdfDist = 0;
        MediumInteraction mi = makeMI(f3(0.f), f3(0.f), false);
        float Tr;
        if (bounce >= 1 || noReuse) {
            if (noReuse) {
                SampleMediumSuperVoxelGeneric(ray, sg, mi, 0);
                pdfDist = 1.f; Tr = 1.f;
                curHitDist = mi.isValid ? length(mi.p - ray.origin) : kRayTMax;
            } else {
                int curMip = options.visibilityMipLevel;
                if (fp.useCoarserGrid) curMip = min((options.visibilityMipLevel >= VRESTIR_NUM_MAX_MIPS ? VRESTIR_NUM_MAX_MIPS : 0) + vd.numMips - 1, curMip + 1);
                float hd[4], pd[4], ot[4];
                SampleMediumAnalyticGeneric(ray, sg, options.visibilityUseLinearSampler, hd, curMip, pd, ot, 1);
                curHitDist = hd[0]; pdfDist = pd[0]; Tr = ot[0];
                mi = makeMI(ray.at(curHitDist), -ray.dir, curHitDist != kRayTMax);
            }
        } else {
            curHitDist = precomputedHitDistance; pdfDist = precomputedPdfDist; Tr = precomputedTr;
            mi = makeMI(ray.at(curHitDist), -ray.dir, curHitDist != kRayTMax);
        }
        pathPdf *= pdfDist;
        if (B > 1 && bounce == options.vertexReuseStartBounce && curHitDist != kRayTMax) {   // VERTEX_REUSE :88-94, area measure at the reuse vertex
            pathPdf /= curHitDist * curHitDist;
            pathPHat /= curHitDist * curHitDist;
        }
        bool hitEmpty = false;
        float actualVolumeDensity = 0.f;
        if (bounce == 0) {
            if (B > 1) outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, 0);
            outReservoir.depth = mi.isValid ? curHitDist : kRayTMax;
            primaryScatterDepth = outReservoir.depth;
            outReservoir.p_y = pathPdf;
        } else {
            outReservoir.depth = primaryScatterDepth;
            if (B > 1) {
                outReservoir.sampledPixel = encodeMaxIndirectBounces(outReservoir.sampledPixel, bounce);
                if (bounce < options.vertexReuseStartBounce) extrabounceReservoir[bounce - 1] = encodeWiDist(make_float4(ray.dir.x, ray.dir.y, ray.dir.z, !mi.isValid ? kRayTMax : curHitDist));
                else extrabounceReservoir[bounce - 1] = !mi.isValid ? f3(kRayTMax) : mi.p;   // VERTEX_REUSE :116-125, world-space vertex
            }
            outReservoir.p_y = pathPdf;
        }
        if (mi.isValid) {
            if (noReuse) actualVolumeDensity = 1.f;
            else actualVolumeDensity = DensityWorldSpace(mi.p, 0);
        }
        if ((!mi.isValid && bounce > 0) || (mi.isValid && actualVolumeDensity == 0)) { outReservoir.p_y = 0.f; outReservoir.runningSum = 0.f; hitEmpty = true; }
        float pdfDir = 1.f;
        if (!hitEmpty) {
            if (mi.isValid) {
                float3 albedo = sigS / vd.sigma_t;
                outReservoir.lightID = -1;
                outReservoir.lightUV = make_float2(0, 0);
                float outLightPdf = 0.f;
                float3 Ld = f3(0.f), Le = f3(0.f);
                float3 one_minus_albedo = f3(1.f) - albedo;
                if (vd.hasEmission && (actualVolumeDensity > 0.f)) Le = EmissionWorldSpace(mi.p);
                const bool shouldComputeLightVisibility = options.lightSamples == 0 ? false : true;
                Ld = SampleDirectLighting(sg, outLightPdf, mi, options, shouldComputeLightVisibility, outReservoir.lightID, outReservoir.lightUV);
                float3 wo = -ray.dir, wi = f3(0.f);
                if (B > 1) pdfDir = mi.Sample_p(wo, wi, sampleNext2D(sg));
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
                    pathPHat *= Tr;
                    pathPHat *= actualVolumeDensity;
                    if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID) p_y = pathPHat * luminance(sigA * Le);
                    else p_y = pathPHat * luminance(sigS * Ld * outLightPdf);
                    pathPHat *= luminance(sigS) * pdfDir;
                    if (noReuse) { p_y /= vd.sigma_t; pathPHat /= vd.sigma_t; }
                    if (outReservoir.runningSum > 0.f) {
                        outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                        if (outReservoir.lightID == VRESTIR_SELF_EMISSION_LIGHT_ID && bounce > 0) {
                            if (bounce < options.vertexReuseStartBounce) {   // VERTEX_REUSE :324-331
                                encodeEmissivePosition(mi.p, outReservoir.lightID, outReservoir.lightUV);
                                p_y /= (curHitDist * curHitDist);
                            }
                            outReservoir.sampledPixel = encodePathTag(outReservoir.sampledPixel, 1);
                        }
                        outReservoir.p_y = p_y;
                    }
                }
                pathPdf *= pdfDir;
                if (B > 1 && bounce < B - 1) {
                    ray = makeRay(mi.p, wi, 0, kRayTMax);
                    if (fp.useRussianRoulette && bounce >= 2) {
                        if (sampleNext1D(sg) < albedo.x) pathPdf *= albedo.x;
                        else { hitEmpty = true; combinedReservoir.M++; }
                    }
                }
            } else {
                float3 Le = envEval(ray.dir);
                pathPHat *= Tr;
                float p_y = pathPHat * luminance(Le);
                outReservoir.runningSum = outReservoir.p_y == 0.f ? 0.f : p_y / outReservoir.p_y;
                outReservoir.p_y = p_y;
                hitEmpty = true;
            }
        }
        if (B > 1) simpleResampleStep<B>(outReservoir, combinedReservoir, sg);
        if (hitEmpty) break;
    }
    if (B > 1) { combinedReservoir.M = 1; return combinedReservoir; }
    return outReservoir;
}

// VR/VolumePathTracingFunctions.slang:3-131 (mUseReference)
__device__ float3 IntegrateByVolumePathTracing(Ray ray, SampleGenerator& sg, const FrameParams& fp) {
    const vrestir_volume_desc& vd = c_scene.vol;
    const SamplingOptions& o = fp.initial;
    const int lightSamples = max(1, o.lightSamples);
    const int maxBouncesIn = fp.maxBounces;
    MediumInteraction mi = makeMI(f3(0.f), f3(0.f), false);
    float3 beta = f3(1.f), L = f3(0.f);
    int maxBounces = maxBouncesIn;
    for (int bounce = 0; bounce < maxBounces; bounce++) {
        mi.isValid = false;
        SampleMediumSuperVoxelGeneric(ray, sg, mi, 0);
        if (mi.isValid) {
            float3 albedo = v3(vd.sigma_s) / vd.sigma_t;
            float3 one_minus_albedo = f3(1.f) - albedo;
            { float3 Le = EmissionWorldSpace(mi.p); L = L + Le * one_minus_albedo * beta; }
            beta = beta * albedo;
            {
                float3 Ld = directLighting(sg, mi, lightSamples, o.useEnvironmentLights, o.useAnalyticLights, o.useEmissiveLights, 0, VRESTIR_RESIDUAL_RATIO_TRACKING);
                L = L + beta * Ld;
            }
            float3 wo = -ray.dir, wi = f3(0.f);
            if (maxBounces > 1) mi.Sample_p(wo, wi, sampleNext2D(sg));
            if (bounce < maxBounces - 1) {
                ray = makeRay(mi.p, wi, 0, kRayTMax);
                if (fp.useRussianRoulette && bounce >= 2) {
                    if (sampleNext1D(sg) < albedo.x) beta = beta / albedo.x;
                    else bounce = maxBounces;
                }
            }
        } else {
            float3 Le = envEval(ray.dir);
            if (bounce == 0) L = L + beta * Le;
            bounce = maxBounces;
        }
    }
    return L;
}

template <int B>
__global__ void __launch_bounds__(128, VR_MINB) k_initial(FrameParams fp) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount));
    const int reservoirId = y * fp.W + x;
    Ray ray = primaryRay(fp, x, y);
    if (fp.useReference) {
        float3 avgL = f3(0.f);
        for (int r = 0; r < fp.baselineSpp; r++) avgL = avgL + IntegrateByVolumePathTracing(ray, sg, fp);
        float3 o = avgL / (float)fp.baselineSpp;
        fp.refColor[reservoirId] = make_float4(o.x, o.y, o.z, 1.f);
        return;
    }
    float3 finalExtra[B > 1 ? B - 1 : 1];
    float3 extra[B > 1 ? B - 1 : 1];
#pragma unroll
    for (int i = 0; i < (B > 1 ? B - 1 : 1); i++) { finalExtra[i] = f3(0.f); extra[i] = f3(0.f); }
    Reservoir finalReservoir = createNewReservoir();
    const int rounds = (fp.initialM + 3) / 4;
    for (int roundId = 0; roundId < rounds; roundId++) {
        float hd[4] = {0, 0, 0, 0}, pd[4] = {0, 0, 0, 0}, ot[4] = {0, 0, 0, 0};
        const int roundSamples = roundId == rounds - 1 ? (fp.initialM - 4 * (rounds - 1)) : 4;
        if (!fp.noReuse) SampleMediumAnalyticGeneric(ray, sg, fp.initial.visibilityUseLinearSampler, hd, fp.initial.visibilityMipLevel, pd, ot, roundSamples);
        for (int s = 0; s < roundSamples; s++) {
            Reservoir outReservoir = ComputeInitialSample<B>(ray, hd[s], pd[s], ot[s], sg, fp, extra);
            bool isSelected = simpleResampleStep<B>(outReservoir, finalReservoir, sg);
            if (B > 1 && isSelected) {
                const int mib = decodeMaxIndirectBounces<B>(finalReservoir.sampledPixel);
                for (int b = 0; b < mib && b < B - 1; b++) finalExtra[b] = extra[b];
            }
        }
    }
    ExtraProvider prov; prov.global = nullptr; prov.local = finalExtra;
    Reservoir tapForEval = finalReservoir; tapForEval.extraBounceStartId = 0;
    InlineMarch mp;
    float p_hat = evaluate_P_hat<B>(ray, sg, prov, fp.spatial, tapForEval, false, false, mp);
    finalReservoir.p_partial = tapForEval.p_partial;   // TraceRays.cs.slang:176-177: finalReservoir itself is the inout argument
    if (finalReservoir.runningSum > 0.f) {
        finalReservoir.runningSum *= finalReservoir.p_y == 0.f ? 0.f : p_hat / finalReservoir.p_y;
        finalReservoir.p_y = p_hat;
    }
    storeReservoir(fp.cur, reservoirId, finalReservoir);
    if (B > 1) {
        const int mib = decodeMaxIndirectBounces<B>(finalReservoir.sampledPixel);
        for (int b = 0; b < mib && b < B - 1; b++) fp.extCur[(size_t)reservoirId * (B - 1) + b] = finalExtra[b];
    }
}

// ------------------------------------------------------------------------------------------------ K2 / K3 / K5
// bodies in vr_stages.cuh; here with every transmittance marched in place
template <int B>
__global__ void __launch_bounds__(128, VR_MINB) k_temporal(FrameParams fp) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    InlineMarch mp;
    temporalPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MINB) k_spatial(FrameParams fp) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    InlineMarch mp;
    spatialPixel<B>(fp, x, y, mp);
}
template <int B>
__global__ void __launch_bounds__(128, VR_MINB) k_final(FrameParams fp) {
    int x, y;
    if (!pixelOf(fp, x, y)) return;
    InlineMarch mp;
    finalPixel<B>(fp, x, y, mp);
}

// ------------------------------------------------------------------------------------------------ K6
__global__ void __launch_bounds__(256) k_importance(float* importance, int dim, int sx, int sy) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= dim || py >= dim) return;
    const float invSamples = 1.f / (float)(sx * sy);
    const float dimSx = (float)(dim * sx), dimSy = (float)(dim * sy);
    float L = 0.f;
    for (int yy = 0; yy < sy; yy++)
        for (int xx = 0; xx < sx; xx++) {
            uint32_t spx = (uint32_t)px * sx + xx, spy = (uint32_t)py * sy + yy;
            float2 pp = make_float2(((float)spx + 0.5f) / dimSx, ((float)spy + 0.5f) / dimSy);
            float3 dir = oct_to_ndir_equal_area_unorm(pp);
            float2 uv = world_to_latlong_map(dir);
            L += luminance(envBilinear(uv));
        }
    importance[(size_t)py * dim + px] = L * invSamples;
}
__global__ void __launch_bounds__(256) k_importance_mip(const float* src, float* dst, int d) {
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= d || y >= d) return;
    const int dp = d * 2;
    float a = src[(size_t)(2 * y) * dp + 2 * x], b = src[(size_t)(2 * y) * dp + 2 * x + 1];
    float c = src[(size_t)(2 * y + 1) * dp + 2 * x], e = src[(size_t)(2 * y + 1) * dp + 2 * x + 1];
    dst[(size_t)y * d + x] = ((a + b) + (c + e)) * 0.25f;
}

// AoS <-> SoA converters for get/set_buffer
__global__ void k_res_to_aos(ResBuf b, vrestir_reservoir* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    float4 a = b.p0[i], c = b.p1[i];
    vrestir_reservoir r; r.runningSum = a.x; r.M = a.y; r.depth = a.z; r.p_y = a.w; r.lightUV[0] = c.x; r.lightUV[1] = c.y; r.lightID = __float_as_int(c.z); r.sampledPixel = __float_as_int(c.w);
    out[i] = r;
}
__global__ void k_res_from_aos(ResBuf b, const vrestir_reservoir* in, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    vrestir_reservoir r = in[i];
    b.p0[i] = make_float4(r.runningSum, r.M, r.depth, r.p_y);
    b.p1[i] = make_float4(r.lightUV[0], r.lightUV[1], __int_as_float(r.lightID), __int_as_float(r.sampledPixel));
}

// ------------------------------------------------------------------------------------------------ bandwidth probe
// every thread streams the buffer with 16-byte loads (grid-stride), `iters` times; the xor keeps the loads alive
__global__ void __launch_bounds__(256) k_read_bandwidth(const uint4* __restrict__ buf, size_t n16, int iters, unsigned* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) *sink = acc.x;
}
cudaError_t launchReadBandwidth(const void* buf, size_t bytes, int iters, int blocks, unsigned* sink, cudaStream_t st) {
    k_read_bandwidth<<<blocks, 256, 0, st>>>((const uint4*)buf, bytes / 16, iters, sink);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 gridFor(const FrameParams& fp) { return dim3((fp.W + 15) / 16, (fp.rowEnd - fp.rowBegin + 7) / 8); }

cudaError_t readDebugRays(float* out64x8, unsigned* count) {
    cudaError_t e = cudaMemcpyFromSymbol(count, g_dbgCount, 4);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out64x8, g_dbgRays, sizeof(float) * 64 * 8);
    unsigned zero = 0;
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_dbgCount, &zero, 4);
    return e;
}
cudaError_t uploadScene(const DScene& s, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_scene, &s, sizeof(DScene), 0, cudaMemcpyHostToDevice, st); }
cudaError_t uploadPrevCam(const DPrevCam& s, cudaStream_t st) { return cudaMemcpyToSymbolAsync(c_prev, &s, sizeof(DPrevCam), 0, cudaMemcpyHostToDevice, st); }

#define VR_DISPATCH_B(kern, fp, st)                                                              \
    switch ((fp).maxBounces) {                                                                   \
        case 1: kern<1><<<gridFor(fp), 128, 0, st>>>(fp); break;                                 \
        case 2: kern<2><<<gridFor(fp), 128, 0, st>>>(fp); break;                                 \
        case 3: kern<3><<<gridFor(fp), 128, 0, st>>>(fp); break;                                 \
        default: kern<4><<<gridFor(fp), 128, 0, st>>>(fp); break;                                \
    }

cudaError_t launchFeatures(const FrameParams& fp, cudaStream_t st) { k_features<<<gridFor(fp), 128, 0, st>>>(fp); return cudaGetLastError(); }
cudaError_t launchInitial(const FrameParams& fp, cudaStream_t st) { VR_DISPATCH_B(k_initial, fp, st); return cudaGetLastError(); }
cudaError_t launchTemporal(const FrameParams& fp, cudaStream_t st) { VR_DISPATCH_B(k_temporal, fp, st); return cudaGetLastError(); }
cudaError_t launchSpatial(const FrameParams& fp, cudaStream_t st) { VR_DISPATCH_B(k_spatial, fp, st); return cudaGetLastError(); }
cudaError_t launchFinal(const FrameParams& fp, cudaStream_t st) { VR_DISPATCH_B(k_final, fp, st); return cudaGetLastError(); }
cudaError_t launchImportance(float* importance, int dim, int sx, int sy, cudaStream_t st) {
    k_importance<<<dim3((dim + 15) / 16, (dim + 15) / 16), 256, 0, st>>>(importance, dim, sx, sy);
    return cudaGetLastError();
}
cudaError_t launchImportanceMip(const float* src, float* dst, int d, cudaStream_t st) {
    k_importance_mip<<<dim3((d + 15) / 16, (d + 15) / 16), 256, 0, st>>>(src, dst, d);
    return cudaGetLastError();
}
cudaError_t launchResToAos(ResBuf b, vrestir_reservoir* out, int n, cudaStream_t st) { k_res_to_aos<<<(n + 255) / 256, 256, 0, st>>>(b, out, n); return cudaGetLastError(); }
cudaError_t launchResFromAos(ResBuf b, const vrestir_reservoir* in, int n, cudaStream_t st) { k_res_from_aos<<<(n + 255) / 256, 256, 0, st>>>(b, in, n); return cudaGetLastError(); }

}  // namespace vrd
