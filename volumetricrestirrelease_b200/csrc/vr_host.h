// vr_host.h — internal host-side declarations shared by the C-ABI translation units (not installed).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/vrestir.h"

namespace vr {
int setError(int code, const std::string& msg);   // records vrestir_last_error() for this thread, returns code
// call from a catch (...) handler of a C entry point: no C++ exception may unwind through the C ABI
int caughtException();

// tree over a brick-activity map (vr_scene.cpp)
struct Topology {
    std::vector<vrestir_node> nodes[3];   // [0] = bricks (pos, link = brick id; bounds left 0), [1] level-1 nodes, [2] root if topLev == 2
    std::vector<uint32_t> child[3];
    uint32_t brickCount = 0; int n1count = 0; int topLev = 1;
};
struct HostSlotVectors { std::vector<vrestir_node>* nodes[3]; std::vector<uint32_t>* child[3]; std::vector<uint8_t>* atlas; vrestir_grid_slot* desc; };
}
struct vrestir_scene;
namespace vr {
vrestir_scene* newHostScene(const vrestir_volume_desc& vol);
HostSlotVectors hostSlotVectors(vrestir_scene* s, int slot);
void attachBlackbodyLut(vrestir_scene* s);
void buildTopology(std::vector<uint8_t>& active /* (nz+7)/8 x (ny+7)/8 x (nx+7)/8, may get its first entry set */, int nx, int ny, int nz, Topology& out);
}
