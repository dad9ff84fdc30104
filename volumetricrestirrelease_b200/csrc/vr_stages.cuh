// vr_stages.cuh — the per-pixel bodies of the reuse stages (K2 temporal, K3 spatial, K5 final), written once against a march
// provider (vr_device.cuh: InlineMarch / the task-stream providers of vr_wavefront.cu).  vr_kernels.cu instantiates them with
// InlineMarch as the per-pixel kernels; vr_wavefront.cu runs the same bodies as an emit pass and a consume pass around the
// march engine (multi-bounce option sets).
#pragma once
#include "vr_device.cuh"

namespace vrd {

// ------------------------------------------------------------------------------------------------ K2
// VR/TemporalReuse.cs.slang:80-377 for one pixel; MP = march provider (vr_device.cuh)
template <int B, class MP>
__device__ __forceinline__ void temporalPixel(const FrameParams& fp, int x, int y, MP& mp) {
    const int W = fp.W, H = fp.H;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount + 1));
    int selectedId = -1;
    const int pixelId = y * W + x;
    int numUsedReservoirs = 1;
    Reservoir taps[2];
    taps[0] = loadReservoirRW(fp.cur, pixelId, B);
    taps[1] = createNewReservoir();
    Ray ray = primaryRay(fp, x, y);
    const uint32_t mis = fp.temporalMIS;
    Reservoir output = mis == VRESTIR_MIS_TALBOT ? createNewReservoir() : taps[0];
    const int centerExtraBounceStartId = taps[0].extraBounceStartId;
    float temporalOriginalDepth = 0.f;
    int2 reprojScreenPos = make_int2(0, 0);
    const int2 cf = fp.features[pixelId];
    const bool isBackgroundReservoir = __int_as_float(cf.y) == 1.f && cf.x;
    bool useFallbackReservoir = true;
    if (fp.reprojectionMode != VRESTIR_REPROJECTION_NONE) {
        float reprojDepth = taps[0].depth;
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
            if (isBackgroundReservoir && !isTapBackgroundReservoir) {
                if (MP::kStore && fp.outputMotionVec && fp.outMvec) fp.outMvec[pixelId] = make_float2((float)(reprojScreenPos.x - x) / (float)W, (float)(reprojScreenPos.y - y) / (float)H);
                return;
            }
        }
        {
            scrPosI = make_int2(f2i(scrPos.x), f2i(scrPos.y));
            reprojScreenPos = scrPosI;
            if (scrPosI.x >= 0 && scrPosI.x < W && scrPosI.y >= 0 && scrPosI.y < H) { numUsedReservoirs++; taps[1] = loadReservoir(fp.temporal, scrPosI.y * W + scrPosI.x, B); }
        }
        if (numUsedReservoirs > 1) useFallbackReservoir = false;
    }
    if (useFallbackReservoir) { numUsedReservoirs++; reprojScreenPos = make_int2(x, y); taps[1] = loadReservoir(fp.temporal, pixelId, B); }
    if (MP::kStore && fp.outputMotionVec && fp.outMvec) fp.outMvec[pixelId] = make_float2((float)(reprojScreenPos.x - x) / (float)W, (float)(reprojScreenPos.y - y) / (float)H);
    const float curM = taps[0].M;
    const float MaxPrevM = fp.temporalMThreshold * curM;
    if (numUsedReservoirs == 2) {
        temporalOriginalDepth = taps[1].depth;
        if (taps[1].depth != kRayTMax) {
            float3 dir = normalize(camRayDirNN(c_prev.prevU, c_prev.prevV, c_prev.prevW, reprojScreenPos.x, reprojScreenPos.y, W, H));
            float3 worldPos = c_prev.prevPos + taps[1].depth * dir;
            taps[1].depth = length(worldPos - ray.origin);
        }
    }
    float centerPrevFrameDepth = taps[0].depth;
    if (centerPrevFrameDepth != kRayTMax) { float3 worldPos = ray.at(centerPrevFrameDepth); centerPrevFrameDepth = length(worldPos - c_prev.prevPos); }
    bool hasSelection = output.runningSum > 0.f;
    const int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
    if (startSampleId == 1) selectedId = 0;
    ExtraProviderRW curProv; curProv.global = fp.extCur;
    ExtraProvider tempProv; tempProv.global = fp.extTemporal; tempProv.local = nullptr;
    for (int i = startSampleId; i < numUsedReservoirs; i++) {
        float talbotMISWeight = 1.f;
        float neighbor_py = 0.f;
        if (taps[i].p_y > 0.f) {
            neighbor_py = taps[i].p_y;
            if (isnan(taps[i].runningSum) || isinf(taps[i].runningSum)) taps[i].runningSum = 0.f;
            if (i > 0) { mp.beginEval(i * 2); resampleNeighbor<B>(taps[i], ray, sg, tempProv, fp.spatial, false, mp); }
        } else { taps[i].p_y = 0.f; taps[i].runningSum = 0.f; }
        if (mis == VRESTIR_MIS_TALBOT && taps[i].runningSum > 0.f) {
            float p_sum = 0, p_qi = 0, k = 0;
            for (int j = 0; j < numUsedReservoirs; j++) {
                const int2 tapPos2 = make_int2(j == 0 ? x : reprojScreenPos.x, j == 0 ? y : reprojScreenPos.y);
                const float correctedM = fminf(MaxPrevM, taps[j].M);
                k += correctedM;
                if (j == 0) { p_qi = taps[i].p_y; p_sum += taps[i].p_y * correctedM; }
                else if (i == j) { p_qi = neighbor_py; p_sum += neighbor_py * correctedM; }
                else {
                    float3 nOrigin, nDir;
                    if (j == 0) { nOrigin = fp.camPos; nDir = normalize(camRayDirNN(fp.camU, fp.camV, fp.camW, tapPos2.x, tapPos2.y, W, H)); }
                    else { nOrigin = c_prev.prevPos; nDir = normalize(camRayDirNN(c_prev.prevU, c_prev.prevV, c_prev.prevW, tapPos2.x, tapPos2.y, W, H)); }
                    const float usedDepth = j == 0 ? taps[i].depth : (i == 0 ? centerPrevFrameDepth : temporalOriginalDepth);
                    Ray neighborRay = makeRay(nOrigin, nDir, 0, usedDepth);
                    const float backupDepth = taps[i].depth;
                    taps[i].depth = usedDepth;
                    float p_y;
                    mp.beginEval(i * 2 + j);
                    if (i == 0) p_y = evaluatePHatReadOnly<B>(neighborRay, sg, curProv, fp.spatial, taps[i], j > 0, false, mp);
                    else p_y = evaluatePHatReadOnly<B>(neighborRay, sg, tempProv, fp.spatial, taps[i], j > 0, false, mp);
                    taps[i].depth = backupDepth;
                    if (isinf(p_y) || isnan(p_y)) p_y = 0.f;
                    p_sum += p_y * correctedM;
                }
            }
            if (p_sum > 0) talbotMISWeight = p_qi * k / p_sum;
        }
        taps[i].runningSum *= talbotMISWeight;
        const bool isCurrentSelected = simpleResampleStepWithMaxM<B>(taps[i], MaxPrevM, output, sg);
        hasSelection |= isCurrentSelected;
        if (isCurrentSelected) selectedId = i;
    }
    if (!MP::kStore) return;
    if (B > 1) {
        if (hasSelection && selectedId > 0) {
            const int mib = decodeMaxIndirectBounces<B>(output.sampledPixel);
            for (int b = 0; b < mib && b < B - 1; b++) fp.extCur[(size_t)centerExtraBounceStartId + b] = tempProv.get(output.extraBounceStartId + b);
        }
        output.extraBounceStartId = centerExtraBounceStartId;
    }
    storeReservoir(fp.cur, pixelId, output);
}

// ------------------------------------------------------------------------------------------------ K3
// VR/SpatialReuse.cs.slang:94-265 for one pixel; evaluation id = tap * 4 + ray (ray 0 = the pixel's own)
template <int B, class MP>
__device__ __forceinline__ void spatialPixel(const FrameParams& fp, int x, int y, MP& mp) {
    const int W = fp.W, H = fp.H;
    const int numRounds = fp.spatialRounds + fp.roundOffset + 1;
    const int roundId = fp.roundId + fp.roundOffset;
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(numRounds * fp.frameCount + roundId));
    const int pixelId = y * W + x;
    const Reservoir centerIn = loadReservoir(fp.cur, pixelId, B);
    Reservoir output = centerIn;
    const int centerExtraBounceStartId = output.extraBounceStartId;
    Ray ray = primaryRay(fp, x, y);
    const uint32_t mis = fp.spatialMIS;
    if (mis == VRESTIR_MIS_TALBOT) output = createNewReservoir();
    bool hasSelection = output.runningSum > 0.f;
    const int2 cf = fp.features[pixelId];
    const bool IsSelfBackground = !(__int_as_float(cf.y) != 1.f);
    ExtraProvider prov; prov.global = fp.extCur; prov.local = nullptr;
    if (IsSelfBackground) {
        if (!MP::kStore) return;
        const Reservoir& w = mis == VRESTIR_MIS_TALBOT ? centerIn : output;
        storeReservoir(fp.out, pixelId, w);
        if (B > 1) {
            const int mib = decodeMaxIndirectBounces<B>(output.sampledPixel);
            for (int b = 0; b < mib && b < B - 1; b++) fp.extOut[(size_t)output.extraBounceStartId + b] = prov.get(output.extraBounceStartId + b);
        }
        return;
    }
    const int startSampleId = mis == VRESTIR_MIS_TALBOT ? 0 : 1;
    for (int sampleId = startSampleId; sampleId < fp.sampleCount; sampleId++) {
        const int tx = x + fp.offsets[sampleId].x, ty = y + fp.offsets[sampleId].y;
        if (!(tx >= 0 && tx < W && ty >= 0 && ty < H)) continue;
        Reservoir tap = loadReservoir(fp.cur, ty * W + tx, B);
        float MISWeight = 1.f;
        if (sampleId > 0) { mp.beginEval(sampleId * 4); resampleNeighbor<B>(tap, ray, sg, prov, fp.spatial, true, mp); }
        if (mis == VRESTIR_MIS_TALBOT && tap.runningSum > 0.f) {
            float p_sum = 0, p_qi = 0, k = 0;
            for (int j = 0; j < fp.sampleCount; j++) {
                const int tx2 = x + fp.offsets[j].x, ty2 = y + fp.offsets[j].y;
                if (!(tx2 >= 0 && tx2 < W && ty2 >= 0 && ty2 < H)) continue;
                const float4 t2 = __ldg(&fp.cur.p0[ty2 * W + tx2]);   // (runningSum, M, depth, p_y)
                k += t2.y;
                if (j == 0) { p_qi = tap.p_y; p_sum += tap.p_y * t2.y; }
                else if (sampleId == j) { p_qi = t2.w; p_sum += t2.w * t2.y; }
                else {
                    float3 neighborRayDir = normalize(camRayDirNN(fp.camU, fp.camV, fp.camW, tx2, ty2, W, H));
                    Ray neighborRay = makeRay(ray.origin, neighborRayDir, 0, tap.depth);
                    mp.beginEval(sampleId * 4 + j);
                    float p_y = evaluatePHatReadOnly<B>(neighborRay, sg, prov, fp.spatial, tap, false, true, mp);
                    if (isinf(p_y) || isnan(p_y)) p_y = 0.f;
                    p_sum += p_y * t2.y;
                }
            }
            if (p_sum > 0) MISWeight = p_qi * k / p_sum;
        }
        tap.runningSum *= MISWeight;
        if (simpleResampleStep<B>(tap, output, sg)) hasSelection = true;
    }
    if (!MP::kStore) return;
    if (B > 1) {
        if (hasSelection) {
            const int mib = decodeMaxIndirectBounces<B>(output.sampledPixel);
            for (int b = 0; b < mib && b < B - 1; b++) fp.extOut[(size_t)centerExtraBounceStartId + b] = prov.get(output.extraBounceStartId + b);
        }
        output.extraBounceStartId = centerExtraBounceStartId;
    }
    storeReservoir(fp.out, pixelId, output);
}

// ------------------------------------------------------------------------------------------------ K5
// VR/FinalShading.cs.slang:71-141 for one pixel
template <int B, class MP>
__device__ __forceinline__ void finalPixel(const FrameParams& fp, int x, int y, MP& mp) {
    const int pixelId = y * fp.W + x;
    float3 outputColor = f3(0.f);
    SampleGenerator sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount + fp.numTotalRounds - 1));
    if (fp.useReference) {
        float4 c = fp.refColor[pixelId];
        outputColor = f3(c.x, c.y, c.z);
    } else if (fp.visualizeTransmittance) {
        outputColor = f3(powf(__int_as_float(fp.features[pixelId].y), 2.2f));
    } else {
        Reservoir cur = loadReservoir(fp.cur, pixelId, B);
        if (cur.runningSum > 0.f) {
            Ray ray = primaryRay(fp, x, y);
            ray.tMax = cur.depth;
            ExtraProvider prov; prov.global = fp.extCur; prov.local = nullptr;
            mp.beginEval(0);
            float3 col = evaluate_F_<B>(cur, prov, ray, sg, fp.fin, false, fp.noReuse != 0, false, true, mp);
            float Wt = cur.p_y == 0.0f ? 1.f : cur.runningSum / (cur.p_y * cur.M);
            col = col * Wt;
            outputColor = outputColor + col;
        }
    }
    if (!MP::kStore) return;
    float4 o = make_float4(outputColor.x, outputColor.y, outputColor.z, 1.f);
    if (isnan(o.x) || isinf(o.x) || isnan(o.y) || isinf(o.y) || isnan(o.z) || isinf(o.z)) o = make_float4(0.f, 0.f, 0.f, 0.f);
    fp.outColor[pixelId] = o;
}

}  // namespace vrd
