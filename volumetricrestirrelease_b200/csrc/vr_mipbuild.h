// vr_mipbuild.h — internal interface between the GPU mip builder (vr_mipbuild.cu) and the pass (vr_pass.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/vrestir.h"

namespace vr {
// one stored level of a chain + the brick-activity map computed from its raw values (1 byte per 8^3 brick, z-major)
struct ChainLevelView { const void* data; const uint8_t* active; int dim[3]; int format; float maxValue; };
int chainLevelView(const vrestir_mip_chain* chain, int mip, int conservative, ChainLevelView& out);
int chainDevice(const vrestir_mip_chain* chain);
// brick pool from a dense stored level: 10^3 voxels per brick (1-voxel apron, zero outside the grid), brick positions from nodes0
cudaError_t launchPackBricks(const void* level, const int dim[3], int format, const vrestir_node* nodes0, uint32_t brickCount, void* atlas, cudaStream_t st);
// (min, max, avg) of every brick's stored 10^3 block into nodes0[].bounds (F/Scene/Scene.cpp:2989-3010: x outermost, avg = sum / 512)
cudaError_t launchBrickBounds(const void* atlas, int format, float maxValue, vrestir_node* nodes0, uint32_t brickCount, cudaStream_t st);
// quad repack of a single-channel UNORM8 pool for the trilinear fetch (see uploadSlot in vr_pass.cu)
cudaError_t launchQuadRepack(const uint8_t* atlas, uint32_t brickCount, uint32_t* quads, cudaStream_t st);
}  // namespace vr
