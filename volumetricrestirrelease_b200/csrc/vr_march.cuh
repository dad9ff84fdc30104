// vr_march.cuh — the march engine: ray-marched transmittance for a stream of independent march tasks.
//
// One task = one ray through one density mip with up to NT depth thresholds; the result for threshold k is
// exp(-sum sigma_t * step) over the ray-march samples with t < min(boxFar, thr[k]) — bit-identical to running
// VolumeTrackingGVDB<MediumTrRayMarchingAdapter> (VR/VolumeUtils.slang:171-282,350-362 +
// VR/VolumeTrackingAdapterGVDB.slang:140-208) once per threshold: same hierarchical-DDA arithmetic, same global sample
// phase, same partial sums in the same order (see MultiDepthRayMarchingAdapter in vr_device.cuh).
//
// What is different from the per-pixel kernels is only the scheduling:
//   * tasks come from a compacted global stream (background pixels and dead taps never occupy a lane);
//   * a warp is a pool of 32 persistent lanes: a lane that finishes its ray pulls the next task from the stream
//     (refill when >= VR_REFILL_MIN lanes are idle), so short and long rays do not wait for each other;
//   * inside the pool the two inner loops of the reference traversal (empty-space DDA stepping and in-brick sampling)
//     are run as two alternating phases chosen by majority vote over the lanes, so each issued instruction serves at
//     least half of the busy lanes instead of the ~7/32 measured for the nested loops (profiles/r01_ncu_full_*).
#pragma once
#include "vr_device.cuh"

namespace vrd {

#ifndef VR_REFILL_MIN
#define VR_REFILL_MIN 8
#endif

enum { MARCH_IDLE = 0, MARCH_TRAV = 1, MARCH_BRICK = 2 };

template <int NT>
struct Marcher {
    // ray (index space of the mip) and adapter state
    HDDAState dda;                 // dda.pos / dda.dir are the medium-space ray
    float tNear, tFar, tStep;
    float thr[NT], out[NT];
    float Tr;
    unsigned pending, todo;        // thresholds not yet resolved / thresholds this task asked for
    unsigned outIdx;
    // traversal (VR/VolumeUtils.slang:198-279); only levels 1 and 2 exist
    int lev, iter;
    uint32_t link1, link2;
    float3 vmin1, vmin2;
    float tMax1, tMax2;
    // in-brick sampling (VR/VolumeTrackingAdapterGVDB.slang:165-201)
    float3 pb; float t; uint32_t brick; int biter;
    int phase;

    VRD void finish(float* results, bool initialized) {
#pragma unroll
        for (int k = 0; k < NT; k++)
            if ((todo >> k) & 1u) results[outIdx + k] = initialized ? expf(((pending >> k) & 1u) ? Tr : out[k]) : 1.f;
        phase = MARCH_IDLE;
    }

    // originMode: 0 = explicit origin + one threshold (a = origin.xyz, tMax), 1 / 2 = camera / previous camera origin with
    // up to 3 thresholds (a = thr0, thr1, thr2, mask)
    VRD void setup(const uint4 a, const uint4 b, const MarchKind& kind, float* results) {
        Ray rW;
        rW.dir = make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z));
        outIdx = b.w;
        rW.tMin = 0.f;
        if (kind.originMode == 0) {
            rW.origin = make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z));
            thr[0] = __uint_as_float(a.w);
#pragma unroll
            for (int k = 1; k < NT; k++) thr[k] = 0.f;
            todo = 1u;
            rW.tMax = thr[0];
        } else {
            rW.origin = kind.originMode == 1 ? c_scene.camPos : c_scene.prevPos;
            const float v[3] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)};
            todo = a.w & ((1u << NT) - 1u);
            float mx = 0.f;
#pragma unroll
            for (int k = 0; k < NT; k++) { thr[k] = k < 3 ? v[k] : 0.f; if ((todo >> k) & 1u) mx = fmaxf(mx, thr[k]); }
            rW.tMax = mx;
        }
        pending = todo;
#pragma unroll
        for (int k = 0; k < NT; k++) out[k] = 0.f;
        Tr = 0.f;
        const int mip = kind.mip;
        const DSlot& g = c_scene.slots[mip];
        int eff = mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : mip;
        eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
        tStep = c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * kind.tStepScale * (eff + 1);
        Ray ray = WorldToMedium(rW, mip);
        if (!IntersectVolumeBound(ray, tNear, tFar, mip, false)) { finish(results, false); return; }
        lev = g.top_lev;
        link1 = link2 = ID_UNDEFL; vmin1 = vmin2 = f3(0.f); tMax1 = tMax2 = 0.f;
        {
            NodeHead h = loadNodeHead(g, lev, 0);
            if (lev == 2) { link2 = (uint32_t)h.a.w; vmin2 = nodePos(h.a); tMax2 = tFar; }
            else { link1 = (uint32_t)h.a.w; vmin1 = nodePos(h.a); tMax1 = tFar; }
        }
        iter = 0;
        dda.SetFromRay(ray.origin, ray.dir, tNear + 0.01f);
        if (lev == 2) dda.Prepare(vmin2, g.vdel[2], 1.0f / g.vdel[2]); else dda.Prepare(vmin1, g.vdel[1], 1.0f / g.vdel[1]);
        phase = MARCH_TRAV;
    }

    VRD void ascend(const DSlot& g, int topLev) {
        while (lev <= topLev && dda.tx > (lev == 2 ? tMax2 : tMax1)) {
            lev++;
            if (lev <= topLev) dda.Prepare(vmin2, g.vdel[2], 1.0f / g.vdel[2]);
        }
    }

    // one iteration of the outer loop of VolumeTrackingGVDB, up to (not including) the adapter call
    VRD void travStep(const DSlot& g, float* results) {
        const int topLev = g.top_lev;
        const int r = lev == 2 ? g.res[2] : g.res[1];
        if (!(iter < 4096 && lev > 0 && lev <= topLev && inRange(dda.p, r + 1))) { finish(results, true); return; }
        iter++;
        dda.Next();
        const int dm = lev == 2 ? g.dim[2] : g.dim[1];
        const int b = (((dda.p.z << dm) + dda.p.y) << dm) + dda.p.x;
        uint32_t childNodeId;
        {
            const uint32_t listid = lev == 2 ? link2 : link1;
            if (listid == ID_UNDEFL) childNodeId = ID_UNDEFL;
            else {
                const long long idx = (long long)listid * (long long)(r * r * r) + (long long)b;
                childNodeId = (idx < 0 || idx >= (long long)(lev == 2 ? g.childCount32[2] : g.childCount32[1])) ? 0u : __ldg(&g.child[lev][idx]);
            }
        }
        if (childNodeId != ID_UNDEFL) {
            if (lev == 1) {
                // brick entry: MultiDepthRayMarchingAdapter::ExecuteMainStep prologue
                float tt = dda.tx - 0.01f;
                NodeHead leaf = loadNodeHead(g, 0, childNodeId);
                brick = (uint32_t)leaf.a.w;
                tt = tNear + (floorf((tt - tNear) / tStep) + 0.5f) * tStep;
                if (tt < dda.tx) tt += tStep;
                t = tt;
                float3 wp = dda.pos + tt * dda.dir;
                pb = wp - nodePos(leaf.a);
                biter = 0;
                phase = MARCH_BRICK;
                return;
            }
            lev = 1;
            NodeHead h = loadNodeHead(g, 1, childNodeId);
            link1 = (uint32_t)h.a.w; vmin1 = nodePos(h.a);
            tMax1 = dda.ty;
            dda.Prepare(vmin1, g.vdel[1], 1.0f / g.vdel[1]);
        } else {
            dda.Step();
            dda.tx += 0.01f;
        }
        ascend(g, topLev);
    }

    // one iteration of the in-brick sampling loop; on leaving the brick performs the tail of the outer iteration
    VRD void sampleStep(const DSlot& g, bool linear, float* results) {
        const float res = (float)g.res[0];
        if (!(biter < MAX_BRICK_STEPS && pb.x >= 0 && pb.y >= 0 && pb.z >= 0 && pb.x < res && pb.y < res && pb.z < res)) {
            dda.Step();
            dda.tx += 0.01f;
            phase = MARCH_TRAV;
            ascend(g, g.top_lev);
            return;
        }
#pragma unroll
        for (int k = 0; k < NT; k++)
            if ((pending >> k) & 1u) { if (t >= fminf(tFar, thr[k])) { out[k] = Tr; pending &= ~(1u << k); } }
        if (!pending) { finish(results, true); return; }
        float density = DensityInAtlas<false>(g, brick, pb, linear);
        float sigma_t = density * c_scene.vol.sigma_t;
        Tr += -sigma_t * 1.f * tStep;
        const float3 wpt = 1.f * tStep * dda.dir;
        pb = pb + wpt;
        t += 1.f * tStep;
        biter++;
    }
};

// Persistent-lane pool over one task stream.  tasks: 2 x uint4 per task; count: tasks in the stream; cursor: next unclaimed.
template <int NT>
__device__ __forceinline__ void marchPool(const uint4* __restrict__ tasks, unsigned total, unsigned* cursor, float* results, const MarchKind kind) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const DSlot& g = c_scene.slots[kind.mip];
    const bool linear = kind.linear != 0;
    Marcher<NT> m;
    m.phase = MARCH_IDLE;
    bool drained = false;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, m.phase == MARCH_IDLE);
        if (!drained && __popc(idle) >= VR_REFILL_MIN) {
            const unsigned n = __popc(idle);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(cursor, n);
            base = __shfl_sync(FULL, base, 0);
            if (m.phase == MARCH_IDLE) {
                const unsigned idx = base + __popc(idle & ltMask);
                if (idx < total) {
                    const uint4 a = __ldcs(&tasks[2 * (size_t)idx]), b = __ldcs(&tasks[2 * (size_t)idx + 1]);
                    m.setup(a, b, kind, results);
                }
            }
            if (base + n >= total) drained = true;
            idle = __ballot_sync(FULL, m.phase == MARCH_IDLE);
        }
        if (idle == FULL) { if (drained) break; continue; }
        const unsigned tm = __ballot_sync(FULL, m.phase == MARCH_TRAV);
        if (2 * __popc(tm) >= 32 - __popc(idle)) { if (m.phase == MARCH_TRAV) m.travStep(g, results); }
        else { if (m.phase == MARCH_BRICK) m.sampleStep(g, linear, results); }
    }
}

}  // namespace vrd
