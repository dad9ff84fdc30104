// vr_march.cuh — the march engine: ray-marched transmittance for a stream of independent march tasks.
//
// One task = one ray through one density mip with up to NT depth thresholds; the result for threshold k is
// exp(-sum sigma_t * step) over the ray-march samples with t < min(boxFar, thr[k]) — bit-identical to running
// VolumeTrackingGVDB<MediumTrRayMarchingAdapter> (VR/VolumeUtils.slang:171-282,350-362 +
// VR/VolumeTrackingAdapterGVDB.slang:140-208) once per threshold: same hierarchical-DDA arithmetic (F/Scene/GVDB/
// gvdbDda.slang:86-157), same global sample phase, same partial sums in the same order.
//
// What is different from the per-pixel kernels is only the scheduling and the instruction diet:
//   * tasks come from a compacted global stream (background pixels and dead taps never occupy a lane);
//   * a warp is a pool of 32 persistent lanes: a lane that finishes its ray parks its result and pulls the next task from
//     the stream when >= VR_REFILL_MIN lanes are parked (result write-out, task fetch and ray setup then run on >= 8 lanes);
//   * the two inner loops of the reference traversal (empty-space DDA stepping and in-brick sampling) are two alternating
//     phases chosen by majority vote over the busy lanes: every issued step serves at least half of them, instead of the
//     ~7/32 lanes ncu measured for the nested loops (profiles/r01_ncu_full_k_spatial_k_initial_baseline.txt);
//   * the grid descriptor travels as a kernel parameter (constant bank 0 operands instead of indexed c[3] loads), the
//     level-2 state lives in constants (the root node), float->int / floor / UNORM8->float conversions use exact
//     full-rate FADD/LOP forms instead of the quarter-rate F2I / FRND / I2F units.
#pragma once
#include "vr_device.cuh"

namespace vrd {

#ifndef VR_REFILL_MIN
#define VR_REFILL_MIN 16
#endif
#ifndef VR_SLOW_WEIGHT
#define VR_SLOW_WEIGHT 2
#endif
#ifndef VR_STEPS_PER_VOTE
#define VR_STEPS_PER_VOTE 3
#endif
#ifndef VR_PREFETCH_CHILD
#define VR_PREFETCH_CHILD 1
#endif
// FAST sampler, optional: the two texel words of a sample are loaded one step ahead (when the lane steps to the position, not
// when it samples it).  Meant to shorten the tail of a launch (a lone long ray pays the load latency every step).  Measured
// neutral at 1080p and on a 640x360 frame (floor regime), slower with the extra registers: off (profiles/r02_negative_results.txt)
#ifndef VR_SAMPLE_PREFETCH
#define VR_SAMPLE_PREFETCH 0
#endif

// Lane states; state >> 1 is the group the majority vote counts: 0 parked, 1 level-1 stepping, 2 in-brick sampling, 3 slow
// events.  The events of a ray (brick entry / exit, leaving or entering a level-1 node, steps at the root level) are not
// executed where they are detected (1-5 lanes, ncu) but deferred to the head of the phase they lead to, where the lanes
// that hit the same event since the last switch are processed together.
enum { MARCH_IDLE = 0, MARCH_DONE = 1, MARCH_TRAV = 2, MARCH_EXIT = 3, MARCH_BRICK = 4, MARCH_ENTER = 5, MARCH_ASCEND = 6, MARCH_ROOT = 7 };

// exact floor for |x| < 2^22: round-down add of 1.5 * 2^23 leaves floor(x) in the low mantissa bits
VRD float fastFloor(float x, int& i) {
    const float m = __fadd_rd(x, 12582912.f);
    i = __float_as_int(m) - 0x4B400000;
    return m - 12582912.f;
}
// exact UNORM8 code -> float: 2^23 + b, minus 2^23
VRD float launder(float x) { asm volatile("" : "+f"(x)); return x; }
VRD float byteToMagic(uint32_t w, int sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + sel)); }   // 2^23 + byte
VRD float byteToFloat(uint32_t w, int sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + sel) ) - 8388608.f; }

// ---- traversal shared by all marchers: the hierarchical DDA of VolumeTrackingGVDB (VR/VolumeUtils.slang:171-282) as a
// per-lane state machine.  The adapter-specific parts (in-brick work, results) live in the derived marchers.
struct MarchTrav {
    // medium-space ray
    float3 pos, dir, invDir;
    int3 stepI;                    // +1 / -1 per axis (dir >= 0 ? 1 : -1)
    // DDA state at the current level
    float3 tDel, tSide; int3 p; float tx, ty; int mask;   // mask bits 0..2 = x, y, z
    float tNear, tFar;
    // level 1 node in registers, level 2 is the root (constants of the launch)
    int iter;
    uint32_t link1; float3 vmin1; float tMax1;
    uint32_t brick;                // leaf node id while MARCH_ENTER is pending, brick id afterwards
    int phase;

    VRD void prepare(float3 vmin, float vdel, float ivdel) {   // HDDAState::Prepare
        tDel = make_float3(fabsf(vdel * invDir.x), fabsf(vdel * invDir.y), fabsf(vdel * invDir.z));
        float3 pFlt = (pos + tx * dir - vmin) * ivdel;
        float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        const float3 sgn = make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f);
        tSide = ((fl - pFlt + f3(0.5f)) * sgn + f3(0.5f)) * tDel + f3(tx);
        p = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
    }
    VRD void next() {   // HDDAState::Next
        const bool mx = (tSide.x < tSide.y) & (tSide.x <= tSide.z);
        const bool my = (tSide.y < tSide.z) & (tSide.y <= tSide.x);
        const bool mz = (tSide.z < tSide.x) & (tSide.z <= tSide.y);
        mask = (mx ? 1 : 0) | (my ? 2 : 0) | (mz ? 4 : 0);
        ty = mx ? tSide.x : (my ? tSide.y : tSide.z);
    }
    VRD void step() {   // HDDAState::Step (select form) followed by the traversal's `t.x += epsilon`
        if (mask & 1) { tSide.x += tDel.x; p.x += stepI.x; }
        if (mask & 2) { tSide.y += tDel.y; p.y += stepI.y; }
        if (mask & 4) { tSide.z += tDel.z; p.z += stepI.z; }
        tx = ty + 0.01f;
    }
    // WorldToMedium + IntersectVolumeBound + the prologue of VolumeTrackingGVDB (VR/VolumeBase.slang:103-175,
    // VR/VolumeUtils.slang:183-225).  Returns false when the ray misses the box.
    VRD bool beginTraversal(const Ray& rW, const DSlot& g, bool vertexCenter) {
        Ray ray; ray.origin = mulPoint(rW.origin, g.w2m); ray.dir = mulVec(rW.dir, g.w2m); ray.tMin = rW.tMin; ray.tMax = rW.tMax;
        float3 mn = v3(g.bmin), mx = v3(g.bmax);
        if (vertexCenter) { ray.origin = ray.origin - f3(0.5f); mn = mn - f3(0.5f); mx = mx - f3(0.5f); }
        if (!IntersectP(mn, mx, ray, tNear, tFar)) return false;
        beginPrepared(ray.origin, ray.dir, g, vertexCenter);
        return true;
    }
    // the part after the box test, for rays prepared by the emitting kernel (tNear / tFar already set)
    VRD void beginPrepared(float3 rayOrigin, float3 rayDir, const DSlot& g, bool vertexCenter) {
        pos = rayOrigin; dir = rayDir;
        invDir = make_float3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        stepI = make_int3(dir.x >= 0 ? 1 : -1, dir.y >= 0 ? 1 : -1, dir.z >= 0 ? 1 : -1);
        tx = tNear + 0.01f; ty = 0.f; mask = 0;
        iter = 0;
        const float3 rootPos = make_float3((float)g.rootPos[0], (float)g.rootPos[1], (float)g.rootPos[2]);
        const bool top2 = g.top_lev == 2;
        if (top2) { link1 = ID_UNDEFL; vmin1 = f3(0.f); tMax1 = 0.f; prepare(rootPos, g.vdel[2], g.ivdel[2]); phase = MARCH_ROOT; }
        else { link1 = g.rootLink; vmin1 = rootPos; tMax1 = tFar; prepare(rootPos, g.vdel[1], g.ivdel[1]); phase = MARCH_TRAV; }
        if (vertexCenter) {   // VR/VolumeUtils.slang:216-225: step until the start cell is inside the node
            const int r = top2 ? g.res[2] : g.res[1];
            int it = 0;
            while (it++ < 3 && (p.x < 0 || p.y < 0 || p.z < 0 || p.x > r || p.y > r || p.z > r)) { next(); step(); }
        }
        if (phase == MARCH_TRAV) fetchChild(g);
    }

    // ---- slow events (group 3) -------------------------------------------------------------------------------------
    // MARCH_ASCEND: `while (lev <= topLev && t.x > tMax[lev]) { lev++; if (lev <= topLev) Prepare(root) }` on leaving a level-1
    // node; MARCH_ROOT: one iteration of the outer loop of VolumeTrackingGVDB at the root level, including the descent.
    VRD void slowStep(const DSlot& g) {
        tx = launder(tx);
        if (phase == MARCH_ASCEND) {
            if (g.top_lev == 1) { phase = MARCH_DONE; return; }
            prepare(make_float3((float)g.rootPos[0], (float)g.rootPos[1], (float)g.rootPos[2]), g.vdel[2], g.ivdel[2]);
            if (tx > tFar) { phase = MARCH_DONE; return; }   // tMax[2] = tFar
            phase = MARCH_ROOT;
        }
        if (!(iter < 4096 && inRange(p, g.res[2] + 1))) { phase = MARCH_DONE; return; }
        iter++;
        next();
        const unsigned b = (unsigned)((((p.z << g.dim[2]) + p.y) << g.dim[2]) + p.x);
        uint32_t child = ID_UNDEFL;
        if (g.rootLink != ID_UNDEFL) {
            const unsigned idx = g.rootLink * g.res3[2] + b;
            child = idx >= g.childCount32[2] ? 0u : __ldg(g.child[2] + idx);
        }
        if (child == ID_UNDEFL) { step(); if (tx > tFar) phase = MARCH_DONE; return; }
        const int4 h = __ldg((const int4*)&g.nodes[1][child]);
        link1 = (uint32_t)h.w; vmin1 = nodePos(h);
        tMax1 = ty;
        prepare(vmin1, g.vdel[1], g.ivdel[1]);
        if (tx > tMax1) phase = MARCH_ASCEND; else { phase = MARCH_TRAV; fetchChild(g); }
    }

    // ---- level-1 stepping (group 1) ---------------------------------------------------------------------------------
    // MARCH_EXIT: tail of the outer iteration after the adapter returned (`dda.Step(); t.x += epsilon;` + level check)
    VRD void exitBrick(const DSlot& g) {
        step();
        if (tx > tMax1) phase = MARCH_ASCEND; else { phase = MARCH_TRAV; fetchChild(g); }
    }
    // one iteration of the outer loop of VolumeTrackingGVDB inside a level-1 node, up to (not including) the adapter call
    // child id of the level-1 cell p (ID_UNDEFL = empty)
    VRD uint32_t childOf(const DSlot& g) const {
        const unsigned b = (unsigned)((((p.z << g.dim[1]) + p.y) << g.dim[1]) + p.x);
        if (link1 == ID_UNDEFL) return ID_UNDEFL;
        // p == res passes the inclusive bound (VR/VolumeUtils.slang:231) and aliases into the list like the shader's
        // ByteAddressBuffer load; outside the whole list D3D returns 0.  link1 < node count and the list holds < 2^31
        // entries (checked at upload), so 32-bit arithmetic cannot wrap.
        const unsigned idx = link1 * g.res3[1] + b;
        return idx >= g.childCount32[1] ? 0u : __ldg(g.child[1] + idx);
    }
#if VR_PREFETCH_CHILD
    // the child id of the cell a lane stands in is loaded when it steps INTO the cell (`brick` holds it while the lane is in
    // MARCH_TRAV), so the load latency overlaps the vote and the other lanes' work instead of stalling the next step
    VRD void fetchChild(const DSlot& g) { brick = inRange(p, g.res[1] + 1) ? childOf(g) : ID_UNDEFL; }
#else
    VRD void fetchChild(const DSlot&) {}
#endif
    VRD void travStep(const DSlot& g) {
        if (!(iter < 4096 && inRange(p, g.res[1] + 1))) { phase = MARCH_DONE; return; }
        iter++;
        next();
#if VR_PREFETCH_CHILD
        const uint32_t child = brick;
#else
        const uint32_t child = childOf(g);
#endif
        if (child == ID_UNDEFL) { step(); if (tx > tMax1) phase = MARCH_ASCEND; else fetchChild(g); return; }
        brick = child;   // leaf node id until enterBrick() replaces it with the brick id
        phase = MARCH_ENTER;
    }
};

// ---- ray-marched transmittance (MediumTrRayMarchingAdapter) with up to NT depth thresholds.
// FAST: the slot is a single-channel UNORM8 pool with the quad repack and the sampler is trilinear (every reuse mip)
template <int NT, bool FAST>
struct RayMarcher : MarchTrav {
    float tStep;
    float thrEff[NT], out[NT];     // thrEff = min(tFar, thr)
    float Tr;
    unsigned pending, todo, outIdx;
    bool initialized;
    float3 pb; float t; int biter;   // in-brick sampling
#if VR_SAMPLE_PREFETCH
    uint32_t pw0, pw1; float pfx, pfy, pfz;   // texel words and filter fractions of the sample at pb (FAST sampler)
    VRD void prefetchSample(const DSlot& g) {
        if (!FAST) return;
        if (!(pb.x >= 0 && pb.y >= 0 && pb.z >= 0 && pb.x < 8.f && pb.y < 8.f && pb.z < 8.f)) return;   // the step that would sample it leaves the brick instead
        int ix, iy, iz;
        const float qx = pb.x - 0.5f, qy = pb.y - 0.5f, qz = pb.z - 0.5f;
        const float fx0 = fastFloor(qx, ix), fy0 = fastFloor(qy, iy), fz0 = fastFloor(qz, iz);
        pfx = qx - fx0; pfy = qy - fy0; pfz = qz - fz0;
        const uint32_t* q = g.quads + (brick * 810u + (unsigned)(((iz + 1) * 9 + (iy + 1)) * 9 + (ix + 1)));
        pw0 = __ldg(q); pw1 = __ldg(q + 81);
    }
#endif

    VRD void writeOut(float* results) {
#pragma unroll
        for (int k = 0; k < NT; k++)
            if ((todo >> k) & 1u) results[outIdx + k] = initialized ? expf(((pending >> k) & 1u) ? Tr : out[k]) : 1.f;
        phase = MARCH_IDLE;
    }

    VRD void setup(const uint4* q, const MarchKind& kind, const DSlot& g, const float*) {
        const int mip = kind.mip;
        int eff = mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : mip;
        eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
        tStep = c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * kind.tStepScale * (eff + 1);
        Tr = 0.f;
#pragma unroll
        for (int k = 0; k < NT; k++) out[k] = 0.f;
        if (kind.originMode == 0) {
            // prepared task: the ray hits the box; its single threshold is ray.tMax, and min(tFar, tMax) == tFar
            const uint4 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);
            outIdx = c.x;
            todo = pending = 1u;
            initialized = true;
            tNear = __uint_as_float(a.w); tFar = __uint_as_float(b.w);
            thrEff[0] = tFar;
#pragma unroll
            for (int k = 1; k < NT; k++) thrEff[k] = 0.f;
            beginPrepared(make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)),
                          make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)), g, false);
            return;
        }
        const uint4 a = __ldcs(q), b = __ldcs(q + 1);
        Ray rW;
        rW.dir = make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z));
        outIdx = b.w;
        rW.tMin = 0.f;
        rW.origin = make_float3(kind.origin[0], kind.origin[1], kind.origin[2]);
        const float v[3] = {__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)};
        float thr[NT];
        todo = a.w & ((1u << NT) - 1u);
        float mx = 0.f;
#pragma unroll
        for (int k = 0; k < NT; k++) { thr[k] = k < 3 ? v[k] : 0.f; if ((todo >> k) & 1u) mx = fmaxf(mx, thr[k]); }
        rW.tMax = mx;
        pending = todo;
        initialized = beginTraversal(rW, g, false);
        if (!initialized) { phase = MARCH_DONE; return; }
#pragma unroll
        for (int k = 0; k < NT; k++) thrEff[k] = fminf(tFar, thr[k]);
    }

    // ---- in-brick sampling (group 2) --------------------------------------------------------------------------------
    // MARCH_ENTER: prologue of MediumTrRayMarchingAdapter::ExecuteMainStep
    VRD void enterBrick(const DSlot& g) {
        const int4 leaf = __ldg((const int4*)&g.nodes[0][brick]);
        brick = (uint32_t)leaf.w;
        float tt = tx - 0.01f;
        tt = tNear + (floorf((tt - tNear) / tStep) + 0.5f) * tStep;
        if (tt < tx) tt += tStep;
        t = tt;
        const float3 wp = pos + tt * dir;
        pb = wp - nodePos(leaf);
        biter = 0;
        phase = MARCH_BRICK;
#if VR_SAMPLE_PREFETCH
        prefetchSample(g);
#endif
    }

    VRD float sampleFast(const DSlot& g) {   // sampleBrickLinear<false> on the quad repack, same filter arithmetic
#if VR_SAMPLE_PREFETCH
        const float fx = pfx, fy = pfy, fz = pfz;
        const uint32_t w0 = pw0, w1 = pw1;
#else
        int ix, iy, iz;
        const float qx = pb.x - 0.5f, qy = pb.y - 0.5f, qz = pb.z - 0.5f;
        const float fx0 = fastFloor(qx, ix), fy0 = fastFloor(qy, iy), fz0 = fastFloor(qz, iz);
        const float fx = qx - fx0, fy = qy - fy0, fz = qz - fz0;
        const uint32_t* q = g.quads + (brick * 810u + (unsigned)(((iz + 1) * 9 + (iy + 1)) * 9 + (ix + 1)));
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 81);
#endif
        // x-lerps: the codes arrive as 2^23 + b; (2^23 + b1) - (2^23 + b0) == b1 - b0 exactly, so only the base corner of each
        // pair is converted (lerpf(a, b, t) = fma(t, b - a, a))
        const float m000 = byteToMagic(w0, 0), m100 = byteToMagic(w0, 1), m010 = byteToMagic(w0, 2), m110 = byteToMagic(w0, 3);
        const float m001 = byteToMagic(w1, 0), m101 = byteToMagic(w1, 1), m011 = byteToMagic(w1, 2), m111 = byteToMagic(w1, 3);
        const float c00 = __fmaf_rn(fx, m100 - m000, m000 - 8388608.f), c10 = __fmaf_rn(fx, m110 - m010, m010 - 8388608.f);
        const float c01 = __fmaf_rn(fx, m101 - m001, m001 - 8388608.f), c11 = __fmaf_rn(fx, m111 - m011, m011 - 8388608.f);
        const float c0 = lerpf(c00, c10, fy), c1 = lerpf(c01, c11, fy);
        return lerpf(c0, c1, fz) * kUnorm8;
    }

    // one iteration of the in-brick sampling loop; on leaving the brick performs the tail of the outer iteration
    VRD void sampleStep(const DSlot& g, bool linear) {
        const float res = 8.f;   // bricks are 8^3 (VRESTIR_BRICK_VOXELS = 10^3 with apron)
        if (!(biter < MAX_BRICK_STEPS && pb.x >= 0 && pb.y >= 0 && pb.z >= 0 && pb.x < res && pb.y < res && pb.z < res)) {
            phase = MARCH_EXIT;
            return;
        }
#pragma unroll
        for (int k = 0; k < NT; k++)
            if ((pending >> k) & 1u) { if (t >= thrEff[k]) { out[k] = Tr; pending &= ~(1u << k); } }
        // every further sample can only lower Tr and expf(x) == 0 for x <= -110: the remaining thresholds are exactly 0
        if (!pending || Tr < -110.f) { phase = MARCH_DONE; return; }
        float density;
        if (FAST) density = sampleFast(g) * g.compress_scale * c_scene.vol.densityScaleFactorByScaling;
        else density = DensityInAtlas<false>(g, brick, pb, linear);
        const float sigma_t = density * c_scene.vol.sigma_t;
        Tr += -sigma_t * 1.f * tStep;
        const float3 wpt = 1.f * tStep * dir;
        pb = pb + wpt;
        t += 1.f * tStep;
        biter++;
#if VR_SAMPLE_PREFETCH
        if (biter < MAX_BRICK_STEPS) prefetchSample(g);
#endif
    }
};

// ---- the same marcher with the traversal DECOUPLED from the sampling (marchPoolQ).  In the reference's loop a ray alternates
// "find the next brick" and "sample it"; the lanes of a warp drift apart and the vote of marchPool leaves the minority phase
// idle (13-15 of 32 lanes per instruction, ncu).  Nothing in the sampling feeds back into the traversal except "stop": the brick
// sequence of a ray and the DDA time at which each brick is entered depend on the ray alone.  So every lane keeps a small FIFO of
// (leaf node, entry time) in shared memory: its traversal runs ahead and pushes, its sampler pops and accumulates in the same
// order (same partial sums).  Most busy lanes can then take part in EITHER phase, whichever the warp runs.
#ifndef VR_QUEUE_DEPTH
#define VR_QUEUE_DEPTH 8
#endif
// phase choice: 0 = every round the phase more lanes can take part in; k > 0 = hysteresis, a phase is kept until fewer than k/8 of
// the busy lanes can take part.  Measured: hysteresis is slower for every k, so are weighted votes and popping the next brick in
// the middle of a sampling round (profiles/r02_queue_engine.txt)
// steps per vote of the decoupled pool, per phase.  Traversal steps are cheap (43 instructions) and a traversing lane rarely runs out
// of work inside a round (the FIFO absorbs what it finds), so long traversal rounds amortise the vote; a brick yields only a few
// samples, so long sampling rounds idle (sweep in profiles/r02_queue_engine.txt)
// parked lanes at which the decoupled pool refills (the plain pool: VR_REFILL_MIN = 16).  Its busy lanes lose less to a late
// refill than to refilling with few lanes: 20-24 is the flat optimum on the sparse headline scene, 16-20 on the dense / huge grids
// of configs 4 and 5; 20 with 6 traversal steps is the compromise
#ifndef VR_Q_REFILL_MIN
#define VR_Q_REFILL_MIN 20
#endif
#ifndef VR_Q_STEPS_TRAV
#define VR_Q_STEPS_TRAV 6
#endif
#ifndef VR_Q_STEPS_SAMPLE
#define VR_Q_STEPS_SAMPLE 3
#endif
#ifndef VR_QUEUE_KEEP
#define VR_QUEUE_KEEP 0
#endif
template <int NT, bool FAST>
struct RayMarcherQ : RayMarcher<NT, FAST> {
    using Base = RayMarcher<NT, FAST>;
    uint32_t cbrick;        // brick being sampled
    unsigned qh, qt;        // FIFO read / write counters
    bool live, inBrick, cdone;

    // MarchTrav::travStep with the adapter call replaced by a push; the tail of the outer iteration (exitBrick) follows at once
    VRD void travStepQ(const DSlot& g, uint2* q) {
        if (!(this->iter < 4096 && inRange(this->p, g.res[1] + 1))) { this->phase = MARCH_DONE; return; }
        this->iter++;
        this->next();
#if VR_PREFETCH_CHILD
        const uint32_t child = this->brick;
#else
        const uint32_t child = this->childOf(g);
#endif
        if (child != ID_UNDEFL) { q[(qt % VR_QUEUE_DEPTH) * 128] = make_uint2(child, __float_as_uint(this->tx)); qt++; }
        this->step();
        if (this->tx > this->tMax1) this->phase = MARCH_ASCEND; else this->fetchChild(g);
    }
    // prologue of MediumTrRayMarchingAdapter::ExecuteMainStep for the oldest queued brick
    VRD void enterBrickQ(const DSlot& g, const uint2* q) {
        const uint2 e = q[(qh % VR_QUEUE_DEPTH) * 128]; qh++;
        const int4 leaf = __ldg((const int4*)&g.nodes[0][e.x]);
        cbrick = (uint32_t)leaf.w;
        const float txe = __uint_as_float(e.y);
        float tt = txe - 0.01f;
        tt = this->tNear + (floorf((tt - this->tNear) / this->tStep) + 0.5f) * this->tStep;
        if (tt < txe) tt += this->tStep;
        this->t = tt;
        const float3 wp = this->pos + tt * this->dir;
        this->pb = wp - nodePos(leaf);
        this->biter = 0;
        inBrick = true;
    }
    VRD float sampleFastQ(const DSlot& g) {   // RayMarcher::sampleFast on cbrick
        int ix, iy, iz;
        const float qx = this->pb.x - 0.5f, qy = this->pb.y - 0.5f, qz = this->pb.z - 0.5f;
        const float fx0 = fastFloor(qx, ix), fy0 = fastFloor(qy, iy), fz0 = fastFloor(qz, iz);
        const float fx = qx - fx0, fy = qy - fy0, fz = qz - fz0;
        const uint32_t* q = g.quads + (cbrick * 810u + (unsigned)(((iz + 1) * 9 + (iy + 1)) * 9 + (ix + 1)));
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 81);
        const float m000 = byteToMagic(w0, 0), m100 = byteToMagic(w0, 1), m010 = byteToMagic(w0, 2), m110 = byteToMagic(w0, 3);
        const float m001 = byteToMagic(w1, 0), m101 = byteToMagic(w1, 1), m011 = byteToMagic(w1, 2), m111 = byteToMagic(w1, 3);
        const float c00 = __fmaf_rn(fx, m100 - m000, m000 - 8388608.f), c10 = __fmaf_rn(fx, m110 - m010, m010 - 8388608.f);
        const float c01 = __fmaf_rn(fx, m101 - m001, m001 - 8388608.f), c11 = __fmaf_rn(fx, m111 - m011, m011 - 8388608.f);
        const float c0 = lerpf(c00, c10, fy), c1 = lerpf(c01, c11, fy);
        return lerpf(c0, c1, fz) * kUnorm8;
    }
    VRD void sampleStepQ(const DSlot& g, bool linear) {
        const float res = 8.f;
        if (!(this->biter < MAX_BRICK_STEPS && this->pb.x >= 0 && this->pb.y >= 0 && this->pb.z >= 0 && this->pb.x < res && this->pb.y < res && this->pb.z < res)) {
            inBrick = false;
            return;
        }
#pragma unroll
        for (int k = 0; k < NT; k++)
            if ((this->pending >> k) & 1u) { if (this->t >= this->thrEff[k]) { this->out[k] = this->Tr; this->pending &= ~(1u << k); } }
        if (!this->pending || this->Tr < -110.f) { cdone = true; inBrick = false; return; }
        float density;
        if (FAST) density = sampleFastQ(g) * g.compress_scale * c_scene.vol.densityScaleFactorByScaling;
        else density = DensityInAtlas<false>(g, cbrick, this->pb, linear);
        const float sigma_t = density * c_scene.vol.sigma_t;
        this->Tr += -sigma_t * 1.f * this->tStep;
        const float3 wpt = 1.f * this->tStep * this->dir;
        this->pb = this->pb + wpt;
        this->t += 1.f * this->tStep;
        this->biter++;
    }
};

// ---- exact transmittance of the trilinear interpolant (MediumTrAnalyticAdapter with the linear sampler,
// VR/VolumeTrackingAdapterGVDB.slang:20-136; vertex-centred traversal): the in-brick phase walks the 8^3 voxel cells of the
// brick with a leaf DDA and integrates the cubic sigma(t) of every cell in closed form.  One explicit-origin task, one result.
struct AnalyticMarcher : MarchTrav {
    float Tr; unsigned outIdx;
    float3 lSide; int3 lp; float ltx, lty;   // leaf DDA (HDDAState leaf = dda; leaf.PrepareLeaf(vmin_leaf))
    float3 vminLeaf;
    float t; int biter;

    VRD void writeOut(float* results) { results[outIdx] = expf(Tr); phase = MARCH_IDLE; }

    VRD void setup(const uint4* q, const MarchKind& kind, const DSlot& g, const float*) {
        const uint4 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);   // prepared task (vertex-centred box)
        outIdx = c.x;
        Tr = 0.f;
        tNear = __uint_as_float(a.w); tFar = __uint_as_float(b.w);
        beginPrepared(make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)),
                      make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)), g, true);
    }

    VRD void enterBrick(const DSlot& g) {
        const int4 leaf = __ldg((const int4*)&g.nodes[0][brick]);
        brick = (uint32_t)leaf.w;
        vminLeaf = nodePos(leaf);
        t = tx - 0.01f;
        // HDDAState::PrepareLeaf on a copy of the level-1 state
        const float3 tDelL = make_float3(fabsf(invDir.x), fabsf(invDir.y), fabsf(invDir.z));
        const float3 pFlt = pos + tx * dir - vminLeaf;
        const float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        const float3 sgn = make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f);
        lSide = ((fl - pFlt + f3(0.5f)) * sgn + f3(0.5f)) * tDelL + f3(tx);
        lp = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
        ltx = tx; lty = ty;
        biter = 0;
        phase = MARCH_BRICK;
    }

    // one voxel cell of the brick
    VRD void sampleStep(const DSlot& g, bool) {
        if (!(biter < MAX_BRICK_STEPS && inRange(lp, 8))) { phase = MARCH_EXIT; return; }
        // leaf.Next()
        const bool mx = (lSide.x < lSide.y) & (lSide.x <= lSide.z);
        const bool my = (lSide.y < lSide.z) & (lSide.y <= lSide.x);
        const bool mz = (lSide.z < lSide.x) & (lSide.z <= lSide.y);
        lty = mx ? lSide.x : (my ? lSide.y : lSide.z);
        const float maxDeltaT = lty - t;
        float v[8];
        if (g.format == VRESTIR_ATLAS_F32 && g.channels == 1) {
            // corners lp + {0,1}^3 lie inside the 10^3 apron block for every cell of the brick: unchecked loads
            const float* a = (const float*)g.atlas + (brick * (unsigned)VRESTIR_BRICK_VOXELS + (unsigned)(((lp.z + 1) * 10 + (lp.y + 1)) * 10 + (lp.x + 1)));
            const float s1 = g.compress_scale, s2 = c_scene.vol.densityScaleFactorByScaling;
            v[0] = __ldg(a) * s1 * s2; v[1] = __ldg(a + 1) * s1 * s2; v[2] = __ldg(a + 10) * s1 * s2; v[3] = __ldg(a + 11) * s1 * s2;
            v[4] = __ldg(a + 100) * s1 * s2; v[5] = __ldg(a + 101) * s1 * s2; v[6] = __ldg(a + 110) * s1 * s2; v[7] = __ldg(a + 111) * s1 * s2;
        } else FetchEightVoxels(g, brick, lp, v);
        const float3 p0 = pos + ltx * dir - (make_float3((float)lp.x, (float)lp.y, (float)lp.z) + vminLeaf);
        float c3, c2, c1, c0;
        trilinearCubic(v, c_scene.vol.sigma_t, dir, p0, c3, c2, c1, c0);
        const float t_dist = fminf(tFar - t, maxDeltaT);
        const float t2 = t_dist * t_dist, t3 = t2 * t_dist, t4 = t2 * t2;
        Tr += -(c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t_dist);
        // past tFar, or expf(Tr) already underflows to exactly 0 and can only stay there
        if (t + maxDeltaT >= tFar || Tr < -110.f) { phase = MARCH_DONE; return; }
        t += maxDeltaT;
        // leaf.Step()
        ltx = lty;
        if (mx) { lSide.x += fabsf(invDir.x); lp.x += stepI.x; }
        if (my) { lSide.y += fabsf(invDir.y); lp.y += stepI.y; }
        if (mz) { lSide.z += fabsf(invDir.z); lp.z += stepI.z; }
        biter++;
    }
};

// AnalyticMarcher with the traversal decoupled from the cell walk (see RayMarcherQ): the FIFO holds (leaf node, DDA entry time)
struct AnalyticMarcherQ : AnalyticMarcher {
    uint32_t cbrick;
    unsigned qh, qt;
    bool live, inBrick, cdone;
    VRD void travStepQ(const DSlot& g, uint2* q) {
        if (!(iter < 4096 && inRange(p, g.res[1] + 1))) { phase = MARCH_DONE; return; }
        iter++;
        next();
#if VR_PREFETCH_CHILD
        const uint32_t child = brick;
#else
        const uint32_t child = childOf(g);
#endif
        if (child != ID_UNDEFL) { q[(qt % VR_QUEUE_DEPTH) * 128] = make_uint2(child, __float_as_uint(tx)); qt++; }
        step();
        if (tx > tMax1) phase = MARCH_ASCEND; else fetchChild(g);
    }
    VRD void enterBrickQ(const DSlot& g, const uint2* q) {   // AnalyticMarcher::enterBrick for the oldest queued brick
        const uint2 e = q[(qh % VR_QUEUE_DEPTH) * 128]; qh++;
        const int4 leaf = __ldg((const int4*)&g.nodes[0][e.x]);
        cbrick = (uint32_t)leaf.w;
        vminLeaf = nodePos(leaf);
        const float txe = __uint_as_float(e.y);
        t = txe - 0.01f;
        const float3 tDelL = make_float3(fabsf(invDir.x), fabsf(invDir.y), fabsf(invDir.z));
        const float3 pFlt = pos + txe * dir - vminLeaf;
        const float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        const float3 sgn = make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f);
        lSide = ((fl - pFlt + f3(0.5f)) * sgn + f3(0.5f)) * tDelL + f3(txe);
        lp = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
        ltx = txe;
        biter = 0;
        inBrick = true;
    }
    VRD void sampleStepQ(const DSlot& g, bool) {   // AnalyticMarcher::sampleStep on cbrick
        if (!(biter < MAX_BRICK_STEPS && inRange(lp, 8))) { inBrick = false; return; }
        const bool mx = (lSide.x < lSide.y) & (lSide.x <= lSide.z);
        const bool my = (lSide.y < lSide.z) & (lSide.y <= lSide.x);
        const bool mz = (lSide.z < lSide.x) & (lSide.z <= lSide.y);
        lty = mx ? lSide.x : (my ? lSide.y : lSide.z);
        const float maxDeltaT = lty - t;
        float v[8];
        if (g.format == VRESTIR_ATLAS_F32 && g.channels == 1) {
            const float* a = (const float*)g.atlas + (cbrick * (unsigned)VRESTIR_BRICK_VOXELS + (unsigned)(((lp.z + 1) * 10 + (lp.y + 1)) * 10 + (lp.x + 1)));
            const float s1 = g.compress_scale, s2 = c_scene.vol.densityScaleFactorByScaling;
            v[0] = __ldg(a) * s1 * s2; v[1] = __ldg(a + 1) * s1 * s2; v[2] = __ldg(a + 10) * s1 * s2; v[3] = __ldg(a + 11) * s1 * s2;
            v[4] = __ldg(a + 100) * s1 * s2; v[5] = __ldg(a + 101) * s1 * s2; v[6] = __ldg(a + 110) * s1 * s2; v[7] = __ldg(a + 111) * s1 * s2;
        } else FetchEightVoxels(g, cbrick, lp, v);
        const float3 p0 = pos + ltx * dir - (make_float3((float)lp.x, (float)lp.y, (float)lp.z) + vminLeaf);
        float c3, c2, c1, c0;
        trilinearCubic(v, c_scene.vol.sigma_t, dir, p0, c3, c2, c1, c0);
        const float t_dist = fminf(tFar - t, maxDeltaT);
        const float t2 = t_dist * t_dist, t3 = t2 * t_dist, t4 = t2 * t2;
        Tr += -(c3 * t4 / 4 + c2 * t3 / 3 + c1 * t2 / 2 + c0 * t_dist);
        if (t + maxDeltaT >= tFar || Tr < -110.f) { cdone = true; inBrick = false; return; }
        t += maxDeltaT;
        ltx = lty;
        if (mx) { lSide.x += fabsf(invDir.x); lp.x += stepI.x; }
        if (my) { lSide.y += fabsf(invDir.y); lp.y += stepI.y; }
        if (mz) { lSide.z += fabsf(invDir.z); lp.z += stepI.z; }
        biter++;
    }
};

// ---- free-flight distance sampling (SampleMediumAnalyticAdapter with the point sampler and ONE sample,
// VR/VolumeTrackingAdapterGVDB.slang:210-437): the in-brick phase walks the voxel cells of the brick with a leaf DDA and draws one
// exponential step per non-empty cell from the task's OWN random-number stream, exactly as the per-pixel traversal does.  Used
// for the indirect bounces of multi-bounce K1: the task's result slot lies in the pixel's state block (MBK_TRAV: hit distance,
// pdf, transmittance), the generator is read from and written back to MBK_SG of the same block.
struct DistanceMarcher : MarchTrav {
    unsigned outIdx;
    float3 lSide; int3 lp; float lty;   // leaf DDA
    float t; int biter;
    float opticalThickness, hit, outTr, pdf;
    SampleGenerator sg;

    VRD void writeOut(float* results) {
        if (hit == -1.f) { hit = kRayTMax; outTr = expf(-opticalThickness); pdf = outTr; }   // ExecuteEndStep
        results[outIdx] = hit; results[outIdx + 1] = pdf; results[outIdx + 2] = outTr;
        ((float4*)(results + (outIdx - MBK_TRAV + MBK_SG)))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
        phase = MARCH_IDLE;
    }
    VRD void setup(const uint4* q, const MarchKind& kind, const DSlot& g, const float* results) {
        const uint4 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);   // prepared task
        outIdx = c.x;
        const float4 g4 = ((const float4*)(results + (outIdx - MBK_TRAV + MBK_SG)))[0];
        sg.s0 = __float_as_uint(g4.x); sg.s1 = __float_as_uint(g4.y); sg.s2 = __float_as_uint(g4.z); sg.s3 = __float_as_uint(g4.w);
        opticalThickness = 0.f; hit = -1.f; outTr = 0.f; pdf = 0.f;
        tNear = __uint_as_float(a.w); tFar = __uint_as_float(b.w);
        beginPrepared(make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)),
                      make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)), g, false);
    }
    VRD void enterBrick(const DSlot& g) {
        const int4 leaf = __ldg((const int4*)&g.nodes[0][brick]);
        brick = (uint32_t)leaf.w;
        const float3 vminLeaf = nodePos(leaf);
        t = tx - 0.01f;
        // HDDAState::PrepareLeaf on a copy of the level-1 state
        const float3 tDelL = make_float3(fabsf(invDir.x), fabsf(invDir.y), fabsf(invDir.z));
        const float3 pFlt = pos + tx * dir - vminLeaf;
        const float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        const float3 sgn = make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f);
        lSide = ((fl - pFlt + f3(0.5f)) * sgn + f3(0.5f)) * tDelL + f3(tx);
        lp = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
        biter = 0;
        phase = MARCH_BRICK;
    }
    // one voxel cell of the brick
    VRD void sampleStep(const DSlot& g, bool) {
        if (!(biter < MAX_BRICK_STEPS && inRange(lp, 8))) { phase = MARCH_EXIT; return; }
        // leaf.Next()
        const bool mx = (lSide.x < lSide.y) & (lSide.x <= lSide.z);
        const bool my = (lSide.y < lSide.z) & (lSide.y <= lSide.x);
        const bool mz = (lSide.z < lSide.x) & (lSide.z <= lSide.y);
        lty = mx ? lSide.x : (my ? lSide.y : lSide.z);
        const float maxDeltaT = fminf(tFar - t, lty - t);
        const float currentTMax = fminf(tFar, lty);
        const float density = DensityInAtlas<true>(g, brick, make_float3((float)lp.x, (float)lp.y, (float)lp.z) + f3(0.5f), false);
        const float sigma_t = density * c_scene.vol.sigma_t;
        // an empty cell can never be hit (-log(1-u)/0 is +inf or NaN -> kRayTMax); only the draw counts
        if (sigma_t == 0.f) (void)sg.next();
        else {
            const float dT = -logf(1 - sampleNext1D(sg)) / sigma_t;
            float curT = t + dT;
            if (isnan(curT) || isinf(curT)) curT = kRayTMax;
            if (curT < currentTMax) {
                hit = curT;
                outTr = expf(-(dT * sigma_t + opticalThickness));
                pdf = sigma_t * outTr;
                phase = MARCH_DONE;
                return;
            }
        }
        t = currentTMax;
        opticalThickness += maxDeltaT * sigma_t;
        if (t >= tFar) { phase = MARCH_DONE; return; }   // ExecuteEndStep in writeOut
        // leaf.Step()
        if (mx) { lSide.x += fabsf(invDir.x); lp.x += stepI.x; }
        if (my) { lSide.y += fabsf(invDir.y); lp.y += stepI.y; }
        if (mz) { lSide.z += fabsf(invDir.z); lp.z += stepI.z; }
        biter++;
    }
};

// Persistent-lane pool over one task stream.  tasks: 3 (explicit, prepared) or 2 (camera) x uint4 per task; total: tasks in
// the stream; cursor: next unclaimed.
template <class M>
__device__ __forceinline__ void marchPool(const uint4* __restrict__ tasks, unsigned total, unsigned* cursor, float* results, const MarchKind& kind, const DSlot& g) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const bool linear = kind.linear != 0;
    M m;
    m.phase = MARCH_IDLE;
    bool drained = false;
    for (;;) {
        unsigned parked = __ballot_sync(FULL, m.phase < 2);
        if (__popc(parked) >= VR_REFILL_MIN) {
            if (m.phase == MARCH_DONE) m.writeOut(results);
            if (!drained) {
                const unsigned n = __popc(parked);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(cursor, n);
                base = __shfl_sync(FULL, base, 0);
                if (m.phase < 2) {
                    const unsigned idx = base + __popc(parked & ltMask);
                    if (idx < total) m.setup(tasks + (size_t)(kind.originMode == 0 ? 3 : 2) * idx, kind, g, results);
                }
                if (base + n >= total) drained = true;
            }
            parked = __ballot_sync(FULL, m.phase < 2);
            if (parked == FULL) {
                if (__ballot_sync(FULL, m.phase == MARCH_DONE)) continue;   // box misses of this refill: write them out first
                if (drained) break;
                continue;
            }
        }
        const int grp = m.phase >> 1;
        const int nT = __popc(__ballot_sync(FULL, grp == 1)), nS = __popc(__ballot_sync(FULL, grp == 2));
        const int nSlow = 32 - __popc(parked) - nT - nS;
        if (VR_SLOW_WEIGHT * nSlow >= max(nT, nS)) {
            if (grp == 3) m.slowStep(g);
        } else if (nT >= nS) {
            if (m.phase == MARCH_EXIT) m.exitBrick(g);
#pragma unroll 1
            for (int rep = 0; rep < VR_STEPS_PER_VOTE; rep++) if (m.phase == MARCH_TRAV) m.travStep(g);
        } else {
            if (m.phase == MARCH_ENTER) m.enterBrick(g);
#pragma unroll 1
            for (int rep = 0; rep < VR_STEPS_PER_VOTE; rep++) if (m.phase == MARCH_BRICK) m.sampleStep(g, linear);
        }
    }
}


// ---- K1's primary free-flight sampling (SampleMediumAnalyticGeneric with the point sampler, up to 4 samples along the camera ray
// of a pixel, VR/TraceRays.cs.slang:111-128 + VR/VolumeTrackingAdapterGVDB.slang:210-437) on the decoupled pool.  Tasks are
// IMPLICIT: task i is pixel i of the band in 8x4-tile order; setup seeds the pixel's generator, builds its camera ray and clips
// it.  A sample is written to the pixel's state block the moment it is found (hit distances | pdfs | transmittances, 4 floats
// each, at `out`; the generator 4 floats before them), so only the mask of pending samples lives in registers.
struct PrimaryDistanceCtx { FrameParams fp; float* state; unsigned stride, hdOffset; int numSamples; };
struct DistanceMarcherQ : MarchTrav {
    static constexpr bool kImplicitTasks = true;
    unsigned outIdx;
    float3 lSide; int3 lp;
    float t; int biter;
    float opticalThickness;
    unsigned pendingSamples;
    SampleGenerator sg;
    uint32_t cbrick;
    unsigned qh, qt;
    bool live, inBrick, cdone;

    VRD void writeOut(float* results) {
        // ExecuteEndStep for the samples still pending (the traversal left the volume), then the generator
        const float tr = expf(-opticalThickness);
#pragma unroll
        for (int s = 0; s < 4; s++)
            if ((pendingSamples >> s) & 1u) { results[outIdx + s] = kRayTMax; results[outIdx + 4 + s] = tr; results[outIdx + 8 + s] = tr; }
        ((float4*)(results + outIdx - 4))[0] = make_float4(__uint_as_float(sg.s0), __uint_as_float(sg.s1), __uint_as_float(sg.s2), __uint_as_float(sg.s3));
        phase = MARCH_IDLE;
    }
    // returns false for a slot of the tile grid outside the frame
    VRD bool setupIndex(unsigned idx, const MarchKind& kind, const DSlot& g, const PrimaryDistanceCtx& c) {
        const FrameParams& fp = c.fp;
        const unsigned tilesX = (unsigned)(fp.W + 7) / 8u, tile = idx >> 5, within = idx & 31u;
        const int x = (int)((tile % tilesX) * 8u + (within & 7u)), y = fp.rowBegin + (int)((tile / tilesX) * 4u + (within >> 3));
        if (x >= fp.W || y >= fp.rowEnd) return false;
        outIdx = (unsigned)(y * fp.W + x - fp.rowBegin * fp.W) * c.stride + c.hdOffset;
        sg = SampleGenerator::create((uint32_t)x, (uint32_t)y, (uint32_t)(fp.numTotalRounds * fp.frameCount));
        float* out = c.state + outIdx;
        ((float4*)out)[0] = make_float4(0.f, 0.f, 0.f, 0.f); ((float4*)out)[1] = make_float4(0.f, 0.f, 0.f, 0.f); ((float4*)out)[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        opticalThickness = 0.f;
        pendingSamples = (1u << c.numSamples) - 1u;
        const Ray rW = primaryRay(fp, x, y);
        if (!beginTraversal(rW, g, false)) {
            // the camera ray misses the volume box: ExecuteEndStep before ExecuteStartStep
#pragma unroll
            for (int s = 0; s < 4; s++) if (s < c.numSamples) { out[s] = kRayTMax; out[4 + s] = 1.f; out[8 + s] = 1.f; }
            pendingSamples = 0u;
            phase = MARCH_DONE;
        }
        return true;
    }
    VRD void travStepQ(const DSlot& g, uint2* q) {
        if (!(iter < 4096 && inRange(p, g.res[1] + 1))) { phase = MARCH_DONE; return; }
        iter++;
        next();
#if VR_PREFETCH_CHILD
        const uint32_t child = brick;
#else
        const uint32_t child = childOf(g);
#endif
        if (child != ID_UNDEFL) { q[(qt % VR_QUEUE_DEPTH) * 128] = make_uint2(child, __float_as_uint(tx)); qt++; }
        step();
        if (tx > tMax1) phase = MARCH_ASCEND; else fetchChild(g);
    }
    VRD void enterBrickQ(const DSlot& g, const uint2* q) {
        const uint2 e = q[(qh % VR_QUEUE_DEPTH) * 128]; qh++;
        const int4 leaf = __ldg((const int4*)&g.nodes[0][e.x]);
        cbrick = (uint32_t)leaf.w;
        const float3 vminLeaf = nodePos(leaf);
        const float txe = __uint_as_float(e.y);
        t = txe - 0.01f;
        const float3 tDelL = make_float3(fabsf(invDir.x), fabsf(invDir.y), fabsf(invDir.z));
        const float3 pFlt = pos + txe * dir - vminLeaf;
        const float3 fl = make_float3(floorf(pFlt.x), floorf(pFlt.y), floorf(pFlt.z));
        const float3 sgn = make_float3(dir.x >= 0 ? 1.f : -1.f, dir.y >= 0 ? 1.f : -1.f, dir.z >= 0 ? 1.f : -1.f);
        lSide = ((fl - pFlt + f3(0.5f)) * sgn + f3(0.5f)) * tDelL + f3(txe);
        lp = make_int3((int)fl.x, (int)fl.y, (int)fl.z);
        biter = 0;
        inBrick = true;
    }
    // one voxel cell: every pending sample draws (VR/VolumeTrackingAdapterGVDB.slang:300-361, point sampler)
    VRD void sampleStepQ(const DSlot& g, bool, float* results) {
        if (!(biter < MAX_BRICK_STEPS && inRange(lp, 8))) { inBrick = false; return; }
        const bool mx = (lSide.x < lSide.y) & (lSide.x <= lSide.z);
        const bool my = (lSide.y < lSide.z) & (lSide.y <= lSide.x);
        const bool mz = (lSide.z < lSide.x) & (lSide.z <= lSide.y);
        const float lty = mx ? lSide.x : (my ? lSide.y : lSide.z);
        const float maxDeltaT = fminf(tFar - t, lty - t);
        const float currentTMax = fminf(tFar, lty);
        const float density = DensityInAtlas<true>(g, cbrick, make_float3((float)lp.x, (float)lp.y, (float)lp.z) + f3(0.5f), false);
        const float sigma_t = density * c_scene.vol.sigma_t;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (!((pendingSamples >> s) & 1u)) continue;
            if (sigma_t == 0.f) { (void)sg.next(); continue; }   // an empty cell can never be hit; only the draw counts
            const float dT = -logf(1 - sampleNext1D(sg)) / sigma_t;
            float curT = t + dT;
            if (isnan(curT) || isinf(curT)) curT = kRayTMax;
            if (curT < currentTMax) {
                const float tr = expf(-(dT * sigma_t + opticalThickness));
                results[outIdx + s] = curT; results[outIdx + 4 + s] = sigma_t * tr; results[outIdx + 8 + s] = tr;
                pendingSamples &= ~(1u << s);
            }
        }
        if (!pendingSamples) { cdone = true; inBrick = false; return; }
        t = currentTMax;
        opticalThickness += maxDeltaT * sigma_t;
        if (t >= tFar) { cdone = true; inBrick = false; return; }   // ExecuteEndStep in writeOut
        if (mx) { lSide.x += fabsf(invDir.x); lp.x += stepI.x; }
        if (my) { lSide.y += fabsf(invDir.y); lp.y += stepI.y; }
        if (mz) { lSide.z += fabsf(invDir.z); lp.z += stepI.z; }
        biter++;
    }
};

// Pool of 32 persistent lanes over one task stream for RayMarcherQ (see there).  q: VR_QUEUE_DEPTH x 128 entries of shared memory
// per CTA, [slot][thread].
template <class T> struct MarcherTraits { static constexpr bool kImplicit = false; };
template <> struct MarcherTraits<DistanceMarcherQ> { static constexpr bool kImplicit = true; };
template <class MQ, class Ctx = int>
__device__ __forceinline__ void marchPoolQ(const uint4* __restrict__ tasks, unsigned total, unsigned* cursor, float* results, const MarchKind& kind, const DSlot& g,
                                           uint2* qBase, const Ctx* ctx = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const bool linear = kind.linear != 0;
    uint2* const q = qBase + threadIdx.x;
    MQ m;
    m.phase = MARCH_IDLE; m.live = false; m.inBrick = false; m.cdone = false; m.qh = m.qt = 0;
    bool drained = false;
    bool sampling = false;   // warp-uniform: the phase the pool is in (hysteresis: it stays until fewer than VR_QUEUE_KEEP / 8 of the busy lanes can take part)
    for (;;) {
        // finished: the sampler stopped (thresholds passed / transmittance underflowed) or the traversal ended and everything queued is sampled
        bool fin = m.live && (m.cdone || (m.phase == MARCH_DONE && !m.inBrick && m.qh == m.qt));
        unsigned parked = __ballot_sync(FULL, !m.live || fin);
        if (__popc(parked) >= VR_Q_REFILL_MIN) {
            if (fin) { m.writeOut(results); m.live = false; fin = false; }
            if (!drained) {
                const unsigned n = __popc(parked);
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(cursor, n);
                base = __shfl_sync(FULL, base, 0);
                if (!m.live) {
                    const unsigned idx = base + __popc(parked & ltMask);
                    if (idx < total) {
                        bool have = true;
                        if constexpr (MarcherTraits<MQ>::kImplicit) have = m.setupIndex(idx, kind, g, *ctx);
                        else m.setup(tasks + (size_t)(kind.originMode == 0 ? 3 : 2) * idx, kind, g, results);
                        if (have) { m.live = true; m.inBrick = false; m.cdone = false; m.qh = m.qt = 0; }
                    }
                }
                if (base + n >= total) drained = true;
            }
            fin = m.live && (m.cdone || (m.phase == MARCH_DONE && !m.inBrick && m.qh == m.qt));   // box misses of this refill
            parked = __ballot_sync(FULL, !m.live || fin);
            if (parked == FULL) {
                if (__ballot_sync(FULL, fin)) continue;
                if (drained) break;
                continue;
            }
        }
        const bool canS = m.live && !fin && (m.inBrick || m.qh != m.qt);
        const bool canT = m.live && !fin && m.phase == MARCH_TRAV && (m.qt - m.qh) < (unsigned)VR_QUEUE_DEPTH;
        const bool canK = m.live && !fin && (m.phase == MARCH_ASCEND || m.phase == MARCH_ROOT);
        const int nT = __popc(__ballot_sync(FULL, canT)), nS = __popc(__ballot_sync(FULL, canS)), nK = __popc(__ballot_sync(FULL, canK));
#if VR_QUEUE_KEEP == 0
        sampling = nT < nS;   // the phase more lanes can take part in
#else
        const int nBusy = 32 - __popc(parked);
        if (sampling) { if (8 * nS < VR_QUEUE_KEEP * nBusy && nT > nS) sampling = false; }
        else if (8 * nT < VR_QUEUE_KEEP * nBusy && nS > nT) sampling = true;
#endif
        if (nK && VR_SLOW_WEIGHT * nK >= max(nT, nS)) {
            if (canK) m.slowStep(g);
        } else if (!sampling) {
#pragma unroll 1
            for (int rep = 0; rep < VR_Q_STEPS_TRAV; rep++)
                if (m.live && !fin && m.phase == MARCH_TRAV && (m.qt - m.qh) < (unsigned)VR_QUEUE_DEPTH) m.travStepQ(g, q);
        } else {
            if (canS && !m.inBrick) m.enterBrickQ(g, q);
#pragma unroll 1
            for (int rep = 0; rep < VR_Q_STEPS_SAMPLE; rep++) {
                if constexpr (MarcherTraits<MQ>::kImplicit) { if (m.inBrick) m.sampleStepQ(g, linear, results); }
                else { if (m.inBrick) m.sampleStepQ(g, linear); }
            }
        }
    }
}

}  // namespace vrd
