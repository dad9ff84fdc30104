// vr_march_pair.cuh — phase-specialised form of the ray-march engine (single-threshold, prepared tasks, FAST sampler).
//
// vr_march.cuh runs both phases of a ray (empty-space DDA stepping, in-brick sampling) in every warp and picks one per
// vote: each issued step serves about half of the busy lanes (ncu: 13.6 of 32 lanes).  Here the two phases run in
// DIFFERENT warps.  Warps come in pairs that share a pool of PAIR_RAYS ray slots in shared memory:
//
//     T-warp  owns the traversal: fetches tasks, steps the hierarchical DDA (registers), handles level changes; when a ray
//             reaches a brick it writes the ray's traversal state back to its slot and posts the slot in the T->S ring;
//     S-warp  owns the sampling: pops a slot, runs the in-brick sampling loop (MediumTrRayMarchingAdapter), then posts the
//             slot back in the S->T ring (ray left the brick: the T-warp resumes it) or finishes the ray (threshold reached /
//             exp underflow) and returns the slot as free.
//
// Both rings are single-producer single-consumer (one warp each side, warp-synchronous), sized for every slot, so a post
// never blocks and neither warp ever waits on the other while it has rays of its own; an idle warp sleeps (__nanosleep).
// Per-value arithmetic is that of vr_march.cuh (same MarchTrav code for the traversal), so results stay bit-identical.
#pragma once
#include "vr_march.cuh"

namespace vrd {

#ifndef VR_PAIR_RAYS
#define VR_PAIR_RAYS 80
#endif
#ifndef VR_PAIR_SLEEP
#define VR_PAIR_SLEEP 40
#endif
#ifndef VR_PAIR_STEPS
#define VR_PAIR_STEPS 4
#endif
constexpr int PAIR_RAYS = VR_PAIR_RAYS;
constexpr int PAIR_RING = 128;          // >= PAIR_RAYS, power of two
constexpr unsigned PAIR_FREE = 0x8000u;

struct PairShared {
    // ray slots (SoA).  Static per ray: pos, dir, inv, tNear, tFar, outIdx.  Traversal state: the rest.  Tr: adapter sum.
    float pos[3][PAIR_RAYS], dir[3][PAIR_RAYS], inv[3][PAIR_RAYS];
    float tSide[3][PAIR_RAYS]; int p[3][PAIR_RAYS]; float tx[PAIR_RAYS], ty[PAIR_RAYS]; int maskIter[PAIR_RAYS];
    float tNear[PAIR_RAYS], tFar[PAIR_RAYS], tMax1[PAIR_RAYS]; unsigned link1[PAIR_RAYS]; float vmin1[3][PAIR_RAYS];
    float Tr[PAIR_RAYS]; unsigned outIdx[PAIR_RAYS], leaf[PAIR_RAYS];
    unsigned short t2s[PAIR_RING], s2t[PAIR_RING];
    unsigned t2sTail, s2tTail;     // published producer positions
    int inflight, quit;
};

// watchdog diagnostics: a warp that spins more than VR_PAIR_WATCHDOG loop iterations records its state and gives up
#ifndef VR_PAIR_WATCHDOG
#define VR_PAIR_WATCHDOG 4000000u
#endif
static __device__ unsigned g_pairDbg[64 * 16];
static __device__ unsigned g_pairDbgCount;
VRD void pairWatchdog(unsigned role, unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g_, unsigned h) {
    const unsigned k = atomicAdd(&g_pairDbgCount, 1u);
    if (k < 64) { unsigned* o = &g_pairDbg[k * 16]; o[0] = role; o[1] = blockIdx.x; o[2] = threadIdx.x >> 5; o[3] = a; o[4] = b; o[5] = c; o[6] = d; o[7] = e; o[8] = f; o[9] = g_; o[10] = h; }
}
VRD unsigned ldVolatile(const unsigned* p) { return *(const volatile unsigned*)p; }
VRD int ldVolatile(const int* p) { return *(const volatile int*)p; }

// ------------------------------------------------------------------------------------------------ T-warp
VRD void pairTraverse(PairShared& S, const uint4* __restrict__ tasks, unsigned total, unsigned* cursor, float* results, const DSlot& g) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    MarchTrav m;
    m.phase = MARCH_IDLE;
    bool has = false;
    int slot = lane;                      // slots 0..31 start in the hands of the T lanes, the rest as FREE ring entries
    unsigned head = 0, tailLocal = 0;     // S->T consumer position, T->S producer position (warp-uniform)
    bool drained = false;
    unsigned spins = 0;
    for (;;) {
        if (++spins > VR_PAIR_WATCHDOG) {
            const unsigned hm = __ballot_sync(FULL, has), sm_ = __ballot_sync(FULL, slot >= 0);
            if (lane == 0) pairWatchdog(1, hm, sm_, head, ldVolatile(&S.s2tTail), tailLocal, (unsigned)ldVolatile(&S.inflight), drained, (unsigned)m.phase);
            if (lane == 0) *(volatile int*)&S.quit = 2;
            break;
        }
        // ---- (1) refill: resume rays that left a brick, start new tasks in free slots
        if (drained && !has) slot = -1;   // no more tasks: a free slot in hand is retired, the lane serves returning rays
        const unsigned noslot = __ballot_sync(FULL, !has && slot < 0);
        if (noslot) {
            unsigned tail = 0;
            if (lane == 0) tail = ldVolatile(&S.s2tTail);
            tail = __shfl_sync(FULL, tail, 0);
            const unsigned take = min((unsigned)__popc(noslot), tail - head);
            if (take) {
                __threadfence_block();
                const unsigned rank = __popc(noslot & lt);
                if (!has && slot < 0 && rank < take) {
                    const unsigned e = S.s2t[(head + rank) & (PAIR_RING - 1)];
                    slot = (int)(e & 0x7fffu);
                    if (!(e & PAIR_FREE)) {
                        // resume after the brick: load the traversal state, then the tail of the outer iteration
                        m.pos = f3(S.pos[0][slot], S.pos[1][slot], S.pos[2][slot]);
                        m.dir = f3(S.dir[0][slot], S.dir[1][slot], S.dir[2][slot]);
                        m.invDir = f3(S.inv[0][slot], S.inv[1][slot], S.inv[2][slot]);
                        m.stepI = make_int3(m.dir.x >= 0 ? 1 : -1, m.dir.y >= 0 ? 1 : -1, m.dir.z >= 0 ? 1 : -1);
                        m.tSide = f3(S.tSide[0][slot], S.tSide[1][slot], S.tSide[2][slot]);
                        m.p = make_int3(S.p[0][slot], S.p[1][slot], S.p[2][slot]);
                        m.tx = S.tx[slot]; m.ty = S.ty[slot];
                        const int mi = S.maskIter[slot];
                        m.mask = mi & 7; m.iter = mi >> 3;
                        m.tNear = S.tNear[slot]; m.tFar = S.tFar[slot]; m.tMax1 = S.tMax1[slot]; m.link1 = S.link1[slot];
                        m.vmin1 = f3(S.vmin1[0][slot], S.vmin1[1][slot], S.vmin1[2][slot]);
                        m.tDel = make_float3(fabsf(g.vdel[1] * m.invDir.x), fabsf(g.vdel[1] * m.invDir.y), fabsf(g.vdel[1] * m.invDir.z));   // level-1 tDel (bricks are entered from level 1)
                        has = true;
                        m.exitBrick(g);
                    }
                }
                head += take;
            }
        }
        const unsigned want = __ballot_sync(FULL, !has && slot >= 0);
        if (want && !drained) {
            const unsigned n = __popc(want);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(cursor, n);
            base = __shfl_sync(FULL, base, 0);
            bool started = false;
            if (!has && slot >= 0) {
                const unsigned idx = base + __popc(want & lt);
                if (idx < total) {
                    const uint4* q = tasks + 3 * (size_t)idx;
                    const uint4 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);
                    m.tNear = __uint_as_float(a.w); m.tFar = __uint_as_float(b.w);
                    m.beginPrepared(make_float3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)),
                                    make_float3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)), g, false);
                    S.pos[0][slot] = m.pos.x; S.pos[1][slot] = m.pos.y; S.pos[2][slot] = m.pos.z;
                    S.dir[0][slot] = m.dir.x; S.dir[1][slot] = m.dir.y; S.dir[2][slot] = m.dir.z;
                    S.inv[0][slot] = m.invDir.x; S.inv[1][slot] = m.invDir.y; S.inv[2][slot] = m.invDir.z;
                    S.tNear[slot] = m.tNear; S.tFar[slot] = m.tFar; S.outIdx[slot] = c.x; S.Tr[slot] = 0.f;
                    has = true; started = true;
                }
            }
            if (base + n >= total) drained = true;
            const unsigned sm = __ballot_sync(FULL, started);
            if (lane == 0 && sm) atomicAdd(&S.inflight, __popc(sm));
        }
        // ---- (2) nothing in hand: finished, or wait for rays to come back from the S-warp
        if (!__ballot_sync(FULL, has)) {
            int fl = 0;
            if (lane == 0) fl = ldVolatile(&S.inflight);
            fl = __shfl_sync(FULL, fl, 0);
            if (drained && fl == 0) { if (lane == 0) { *(volatile int*)&S.quit = 1; } break; }
            __nanosleep(VR_PAIR_SLEEP);
            continue;
        }
        // ---- (3) steps: level-1 stepping vs slow events by majority
#pragma unroll 1
        for (int rep = 0; rep < VR_PAIR_STEPS; rep++) {
            const unsigned nf = __ballot_sync(FULL, has && m.phase == MARCH_TRAV);
            const unsigned ns = __ballot_sync(FULL, has && (m.phase == MARCH_ASCEND || m.phase == MARCH_ROOT));
            if (!(nf | ns)) break;
            if (__popc(nf) >= __popc(ns)) { if (has && m.phase == MARCH_TRAV) m.travStep(g); }
            else { if (has && (m.phase == MARCH_ASCEND || m.phase == MARCH_ROOT)) m.slowStep(g); }
        }
        // ---- (4) hand-off
        const unsigned doneM = __ballot_sync(FULL, has && m.phase == MARCH_DONE);
        if (doneM) {
            if (has && m.phase == MARCH_DONE) { results[S.outIdx[slot]] = expf(S.Tr[slot]); has = false; m.phase = MARCH_IDLE; }   // slot stays in hand
            if (lane == 0) atomicSub(&S.inflight, __popc(doneM));
        }
        const unsigned enterM = __ballot_sync(FULL, has && m.phase == MARCH_ENTER);
        if (enterM) {
            if (has && m.phase == MARCH_ENTER) {
                S.tSide[0][slot] = m.tSide.x; S.tSide[1][slot] = m.tSide.y; S.tSide[2][slot] = m.tSide.z;
                S.p[0][slot] = m.p.x; S.p[1][slot] = m.p.y; S.p[2][slot] = m.p.z;
                S.tx[slot] = m.tx; S.ty[slot] = m.ty; S.maskIter[slot] = m.mask | (m.iter << 3);
                S.tMax1[slot] = m.tMax1; S.link1[slot] = m.link1;
                S.vmin1[0][slot] = m.vmin1.x; S.vmin1[1][slot] = m.vmin1.y; S.vmin1[2][slot] = m.vmin1.z;
                S.leaf[slot] = m.brick;
                S.t2s[(tailLocal + __popc(enterM & lt)) & (PAIR_RING - 1)] = (unsigned short)slot;
                has = false; slot = -1; m.phase = MARCH_IDLE;
            }
            tailLocal += __popc(enterM);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *(volatile unsigned*)&S.t2sTail = tailLocal;
        }
    }
}

// ------------------------------------------------------------------------------------------------ S-warp
VRD void pairSample(PairShared& S, float* results, const MarchKind& kind, const DSlot& g) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int eff = kind.mip >= VRESTIR_PREV_DENSITY_GRID_OFFSET ? kind.mip - VRESTIR_PREV_DENSITY_GRID_OFFSET : kind.mip;
    eff = eff >= VRESTIR_NUM_MAX_MIPS ? eff - VRESTIR_NUM_MAX_MIPS : eff;
    const float tStep = c_scene.vol.tStep * c_scene.vol.volumeWorldScaling * kind.tStepScale * (eff + 1);
    RayMarcher<1, true> r;   // only the sampling half of its state is used here
    r.tStep = tStep;
    r.phase = MARCH_IDLE;
    int slot = -1;
    unsigned head = 0, tailLocal = PAIR_RAYS - 32;   // the ring starts with the FREE entries of slots 32..PAIR_RAYS-1
    unsigned spins = 0;
    for (;;) {
        if (++spins > VR_PAIR_WATCHDOG) {
            const unsigned bm = __ballot_sync(FULL, r.phase == MARCH_BRICK);
            if (lane == 0) pairWatchdog(2, bm, 0, head, ldVolatile(&S.t2sTail), tailLocal, (unsigned)ldVolatile(&S.inflight), (unsigned)ldVolatile(&S.quit), (unsigned)r.phase);
            break;
        }
        // ---- (1) pop rays that reached a brick: prologue of MediumTrRayMarchingAdapter::ExecuteMainStep
        const unsigned needy = __ballot_sync(FULL, r.phase != MARCH_BRICK);
        if (needy) {
            unsigned tail = 0;
            if (lane == 0) tail = ldVolatile(&S.t2sTail);
            tail = __shfl_sync(FULL, tail, 0);
            const unsigned take = min((unsigned)__popc(needy), tail - head);
            if (take) {
                __threadfence_block();
                const unsigned rank = __popc(needy & lt);
                if (r.phase != MARCH_BRICK && rank < take) {
                    slot = S.t2s[(head + rank) & (PAIR_RING - 1)];
                    r.pos = f3(S.pos[0][slot], S.pos[1][slot], S.pos[2][slot]);
                    r.dir = f3(S.dir[0][slot], S.dir[1][slot], S.dir[2][slot]);
                    r.tx = S.tx[slot]; r.tNear = S.tNear[slot]; r.tFar = S.tFar[slot];
                    r.thrEff[0] = r.tFar; r.pending = r.todo = 1u; r.out[0] = 0.f;
                    r.Tr = S.Tr[slot];
                    r.brick = S.leaf[slot];
                    r.enterBrick(g);
                }
                head += take;
            }
        }
        if (!__ballot_sync(FULL, r.phase == MARCH_BRICK)) {
            int q = 0;
            if (lane == 0) q = ldVolatile(&S.quit);
            q = __shfl_sync(FULL, q, 0);
            if (q) break;
            __nanosleep(VR_PAIR_SLEEP);
            continue;
        }
        // ---- (2) sampling
#pragma unroll 1
        for (int rep = 0; rep < VR_PAIR_STEPS; rep++) if (r.phase == MARCH_BRICK) r.sampleStep(g, true);
        // ---- (3) hand-off: back to the T-warp (left the brick) or finished (threshold reached / exp underflow)
        const unsigned exitM = __ballot_sync(FULL, r.phase == MARCH_EXIT), doneM = __ballot_sync(FULL, r.phase == MARCH_DONE);
        const unsigned post = exitM | doneM;
        if (post) {
            if (r.phase == MARCH_EXIT || r.phase == MARCH_DONE) {
                unsigned e = (unsigned)slot;
                if (r.phase == MARCH_EXIT) S.Tr[slot] = r.Tr;
                else { results[S.outIdx[slot]] = expf(((r.pending & 1u) ? r.Tr : r.out[0])); e |= PAIR_FREE; }
                S.s2t[(tailLocal + __popc(post & lt)) & (PAIR_RING - 1)] = (unsigned short)e;
                r.phase = MARCH_IDLE; slot = -1;
            }
            tailLocal += __popc(post);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                *(volatile unsigned*)&S.s2tTail = tailLocal;
                if (doneM) atomicSub(&S.inflight, __popc(doneM));
            }
        }
    }
}

// CTA = 128 threads = 2 pairs; warp 2k traverses, warp 2k+1 samples
__device__ __forceinline__ void marchPairs(const uint4* __restrict__ tasks, unsigned total, unsigned* cursor, float* results, const MarchKind& kind, const DSlot& g) {
    __shared__ PairShared sh[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PairShared& S = sh[warp >> 1];
    if ((warp & 1) == 0) {
        if (lane == 0) { S.t2sTail = 0; S.s2tTail = PAIR_RAYS - 32; S.inflight = 0; S.quit = 0; }
        for (int i = lane; i < PAIR_RAYS - 32; i += 32) S.s2t[i] = (unsigned short)((32 + i) | PAIR_FREE);
    }
    __syncthreads();
    if ((warp & 1) == 0) pairTraverse(S, tasks, total, cursor, results, g);
    else pairSample(S, results, kind, g);
}

}  // namespace vrd
