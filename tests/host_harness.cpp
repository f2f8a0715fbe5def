// host_harness.cpp -- TEST INFRASTRUCTURE.  Compiles the product's host/device-neutral logic
// (quadtree core) as a serial CPU emulation so its control flow can be checked against the oracle
// on a box without a GPU.  Nothing here ships in libswm_orb.so.
#define SWM_OCTREE_HOST 1
#include "../swarmmap_b200/csrc/octree_core.cuh"

#include <vector>

extern "C" int hh_octree(const uint32_t* pts, int n, int W, int H, int N, uint32_t* out, int out_cap) {
  static swm::OtState<swm::ot_max_live(2048)> S;
  std::vector<uint16_t> pnode(n + 1);
  std::vector<uint8_t> pchild(n + 1);
  return swm::ot_distribute(S, pts, pnode.data(), pchild.data(), n, W, H, N, out, out_cap);
}

#include "../swarmmap_b200/csrc/swm_core.cuh"

// Per-pixel helpers of the level kernel, run serially over an image (same arithmetic as the device).
extern "C" void hh_fast_score_map(const uint8_t* img, int w, int h, int stride, int th, uint8_t* out) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      out[(size_t)y * w + x] =
          (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) ? (uint8_t)swm::fast_score(img + (size_t)y * stride + x, stride, th) : 0;
}
