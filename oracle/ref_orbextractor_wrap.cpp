// ref_orbextractor_wrap.cpp -- C entry points around the REFERENCE's own ORBextractor.cc (compiled unmodified from
// /root/reference/code/src by `make -C oracle ref`, against oracle/ref_shim's stand-ins for the OpenCV headers).
// Only the parts that are plain C++ are exercised: the constructor (scale tables, per-level feature quotas, umax,
// the rBRIEF pattern it hands to the GPU) and ORBextractor::DistributeOctTree with ExtractorNode::DivideNode.  The
// image-processing calls and the reference's CUDA helper classes are defined here as stubs: constructors that do
// nothing, loaders that record what they are given, everything else aborts (it is never reached).
// Test infrastructure only (tests/test_ref_orbextractor.py); nothing in the product links or loads this.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define protected public  // the quota table and DistributeOctTree are protected members
#include <ORBextractor.h>
#undef protected
#include <cuda/Allocator.hpp>

// ---- non-recycling allocator for the duration of a DistributeOctTree call.
// The reference orders equal-sized nodes by std::sort on pair<int, ExtractorNode*>, i.e. by the ADDRESS of the list
// node (ORBextractor.cc:603-604): with glibc malloc, chunks freed by lNodes.erase() are reused by later push_front()
// calls, so that order depends on allocator internals.  The oracle freezes it as "later-created node first"
// (DESIGN.md section 2).  With this bump allocator addresses grow with creation order, and the reference's own code
// then realises exactly that definition -- which is what tests/test_ref_orbextractor.py compares against.
#include <new>
#include <sys/mman.h>
namespace {
char* g_arena = nullptr;
size_t g_arena_cap = 0, g_arena_used = 0;
bool g_arena_on = false;
void arena_begin() {
  if (!g_arena) {
    g_arena_cap = (size_t)1 << 30;
    g_arena = (char*)mmap(nullptr, g_arena_cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_arena == MAP_FAILED) std::abort();
  }
  g_arena_used = 0;
  g_arena_on = true;
}
void arena_end() { g_arena_on = false; }
}  // namespace
void* operator new(size_t n) {
  if (g_arena_on) {
    const size_t off = (g_arena_used + 15) & ~(size_t)15;
    if (off + n > g_arena_cap) std::abort();
    g_arena_used = off + n;
    return g_arena + off;
  }
  void* p = std::malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete(void* p) noexcept {
  if (p >= (void*)g_arena && p < (void*)(g_arena + g_arena_cap)) return;  // arena memory is never recycled
  std::free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }

namespace {
[[noreturn]] void unreachable(const char* what) {
  std::fprintf(stderr, "ref_orbextractor_wrap: %s is a stub and must not be reached\n", what);
  std::abort();
}
std::vector<int> g_umax;
std::vector<cv::Point> g_pattern;
}  // namespace

namespace cv {
void resize(InputArray, OutputArray, Size, double, double, int) { unreachable("cv::resize"); }
void copyMakeBorder(InputArray, OutputArray, int, int, int, int, int, const Scalar&) { unreachable("cv::copyMakeBorder"); }
void GaussianBlur(InputArray, OutputArray, Size, double, double, int) { unreachable("cv::GaussianBlur"); }
void FAST(InputArray, std::vector<KeyPoint>&, int, bool) { unreachable("cv::FAST"); }
float fastAtan2(float, float) { unreachable("cv::fastAtan2"); }
namespace cuda {
void resize(InputArray, OutputArray, Size, double, double, int, Stream&) { unreachable("cv::cuda::resize"); }
void copyMakeBorder(InputArray, OutputArray, int, int, int, int, int, Scalar, Stream&) { unreachable("cv::cuda::copyMakeBorder"); }
Ptr<Filter> createGaussianFilter(int, int, Size, double, double, int, int) { unreachable("cv::cuda::createGaussianFilter"); }
}  // namespace cuda
}  // namespace cv

namespace ORB_SLAM2 { namespace cuda {
cv::cuda::GpuMat::Allocator* gpu_mat_allocator = nullptr;
GpuFast::GpuFast(int, int, int) {}
GpuFast::~GpuFast() {}
void GpuFast::detectAsync(InputArray) { unreachable("GpuFast::detectAsync"); }
void GpuFast::joinDetectAsync(std::vector<KeyPoint>&) { unreachable("GpuFast::joinDetectAsync"); }
IC_Angle::IC_Angle(unsigned int) {}
IC_Angle::~IC_Angle() {}
void IC_Angle::launch_async(InputArray, KeyPoint*, int, int, int, int, int, int) { unreachable("IC_Angle::launch_async"); }
void IC_Angle::join(KeyPoint*, int) { unreachable("IC_Angle::join"); }
void IC_Angle::loadUMax(const int* u_max, int count) { g_umax.assign(u_max, u_max + count); }
GpuOrb::GpuOrb(int) {}
GpuOrb::~GpuOrb() {}
void GpuOrb::launch_async(InputArray, const KeyPoint*, const int) { unreachable("GpuOrb::launch_async"); }
void GpuOrb::join(Mat&) { unreachable("GpuOrb::join"); }
void GpuOrb::loadPattern(const Point* p) { g_pattern.assign(p, p + 512); }
} }  // namespace ORB_SLAM2::cuda

extern "C" {

void* ref_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  return new ORB_SLAM2::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void ref_orb_destroy(void* e) { delete static_cast<ORB_SLAM2::ORBextractor*>(e); }

// tables built by the reference constructor (ORBextractor.cc:346-404)
void ref_orb_tables(void* ep, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2, int32_t* quotas, int32_t* umax16,
                    int8_t* pattern1024) {
  ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(ep);
  const int nl = e->GetLevels();
  for (int i = 0; i < nl; i++) {
    sf[i] = e->mvScaleFactor[i];
    inv_sf[i] = e->mvInvScaleFactor[i];
    sigma2[i] = e->mvLevelSigma2[i];
    inv_sigma2[i] = e->mvInvLevelSigma2[i];
    quotas[i] = e->mnFeaturesPerLevel[i];
  }
  for (int i = 0; i < 16; i++) umax16[i] = i < (int)g_umax.size() ? g_umax[i] : -1;
  for (int i = 0; i < 512; i++) {
    pattern1024[2 * i] = (int8_t)g_pattern[i].x;
    pattern1024[2 * i + 1] = (int8_t)g_pattern[i].y;
  }
}

// ORBextractor::DistributeOctTree (ORBextractor.cc:407-689).  pts: n x (x, y, response) in ROI coordinates as the
// caller (ComputeKeyPointsOctTree) passes them; out: the selected keypoints in the returned order.
int ref_orb_distribute(void* ep, const float* pts, int n, int min_x, int max_x, int min_y, int max_y, int n_features, int level,
                       float* out, int out_cap, int monotonic_allocator) {
  ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(ep);
  struct Scope {
    bool on;
    explicit Scope(bool o) : on(o) { if (on) arena_begin(); }
    ~Scope() { if (on) arena_end(); }
  } scope(monotonic_allocator != 0);
  std::vector<cv::KeyPoint> in(n);
  for (int i = 0; i < n; i++) {
    in[i].pt.x = pts[3 * i];
    in[i].pt.y = pts[3 * i + 1];
    in[i].response = pts[3 * i + 2];
  }
  std::vector<cv::KeyPoint> res = e->DistributeOctTree(in, min_x, max_x, min_y, max_y, n_features, level);
  const int m = (int)res.size() < out_cap ? (int)res.size() : out_cap;
  for (int i = 0; i < m; i++) {
    out[3 * i] = res[i].pt.x;
    out[3 * i + 1] = res[i].pt.y;
    out[3 * i + 2] = res[i].response;
  }
  return (int)res.size();
}

}  // extern "C"
