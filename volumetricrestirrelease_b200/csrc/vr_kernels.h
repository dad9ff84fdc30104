// vr_kernels.h — host-callable launchers of the kernels in vr_kernels.cu.
#pragma once
#include "vr_types.h"

namespace vrd {
cudaError_t readDebugRays(float* out64x8, unsigned* count);
cudaError_t uploadScene(const DScene& s, cudaStream_t st);
cudaError_t uploadPrevCam(const DPrevCam& s, cudaStream_t st);
cudaError_t launchFeatures(const FrameParams& fp, cudaStream_t st);
cudaError_t launchInitial(const FrameParams& fp, cudaStream_t st);
cudaError_t launchTemporal(const FrameParams& fp, cudaStream_t st);
cudaError_t launchSpatial(const FrameParams& fp, cudaStream_t st);
cudaError_t launchFinal(const FrameParams& fp, cudaStream_t st);
cudaError_t launchImportance(float* importance, int dim, int sx, int sy, cudaStream_t st);
cudaError_t launchImportanceMip(const float* src, float* dst, int d, cudaStream_t st);
// wavefront path (vr_wavefront.cu)
cudaError_t uploadSceneWavefront(const DScene& s, cudaStream_t st);
cudaError_t uploadPrevCamWavefront(const DPrevCam& s, cudaStream_t st);
int marchBlocksPerSM(int nt);
cudaError_t launchMarch(const WfStream& s, float* results, const MarchKind& kind, const DSlot& grid, int nt, int blocks, cudaStream_t st);
cudaError_t launchSpatialGather(const FrameParams& fp, const WfBufs& wf, cudaStream_t st);
int analyticBlocksPerSM();
cudaError_t launchMarchAnalytic(const WfStream& s, float* results, const MarchKind& kind, const DSlot& grid, int blocks, cudaStream_t st);
cudaError_t launchFinalGather(const FrameParams& fp, const WfStream& s, float* results, cudaStream_t st);
cudaError_t launchFinalCombine(const FrameParams& fp, const float* results, cudaStream_t st);
cudaError_t launchInitialFinish(const FrameParams& fp, const WfInitial& wi, cudaStream_t st);
cudaError_t launchInitialStep(const FrameParams& fp, const WfInitial& wi, int s, cudaStream_t st);
cudaError_t launchTemporalGather(const FrameParams& fp, const WfBufs4& wf, cudaStream_t st);
cudaError_t launchTemporalCombine(const FrameParams& fp, const WfBufs4& wf, cudaStream_t st);
cudaError_t launchSpatialCombine(const FrameParams& fp, const WfBufs& wf, cudaStream_t st);
cudaError_t launchInitialMBTraverse(const FrameParams& fp, const WfInitialMB& wi, cudaStream_t st);
int initialMBStepBlocksPerSM();
int initialMBBounceTraverseBlocksPerSM();
int distanceBlocksPerSM();
int primaryDistanceBlocksPerSM();
cudaError_t launchPrimaryDistance(const FrameParams& fp, float* state, unsigned stride, unsigned hdOffset, unsigned* cursor, const DSlot& grid, int blocks, cudaStream_t st);
cudaError_t launchInitialStepOnly(const FrameParams& fp, const WfInitial& wi, int s, cudaStream_t st);
cudaError_t launchMarchDistance(const WfStream& s, float* state, const MarchKind& kind, const DSlot& grid, int blocks, cudaStream_t st);
cudaError_t launchInitialMBBounceTraverse(const FrameParams& fp, const WfInitialMB& wi, int blocks, cudaStream_t st);
cudaError_t launchInitialMBStep(const FrameParams& fp, const WfInitialMB& wi, int first, int blocks, cudaStream_t st);
// generic task-stream path (stage = 1 K1's final p-hat, 2 temporal, 3 spatial, 5 final): the stage body as an emit pass / a consume pass
cudaError_t launchStageEmit(int stage, const FrameParams& fp, const MarchStreams& ms, const WfStream& cam, float* results, cudaStream_t st);
cudaError_t launchStageConsume(int stage, const FrameParams& fp, const float* results, cudaStream_t st);
cudaError_t launchReadBandwidth(const void* buf, size_t bytes, int iters, int blocks, unsigned* sink, cudaStream_t st);
cudaError_t launchResToAos(ResBuf b, vrestir_reservoir* out, int n, cudaStream_t st);
cudaError_t launchResFromAos(ResBuf b, const vrestir_reservoir* in, int n, cudaStream_t st);
}
