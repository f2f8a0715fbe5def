// opencv2/core/core.hpp -- a stand-in for the handful of OpenCV core types the reference's vendored DBoW2
// (code/Thirdparty/DBoW2) touches, so that its UNMODIFIED sources compile here without OpenCV (test infrastructure
// only: oracle/Makefile target `ref`, output oracle/_ref/libdbow2_ref.so, used by tests/test_ref_dbow2.py to pin
// the oracle's restatement against the reference's own code).  cv::Mat here is a reference-counted row-major
// byte / float container with the members DBoW2 uses; FileStorage / FileNode exist only so that the YAML
// save()/load() members (virtual, hence instantiated) compile -- they are never called.
#pragma once
// the real header brings these in transitively and the DBoW2 sources rely on that
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    const size_t bytes = (size_t)r * c * elemSize();
    buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
    data = buf_.get();
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.data, 0, (size_t)r * c * m.elemSize());
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    if (data) std::memcpy(m.data, data, (size_t)rows * cols * elemSize());
    return m;
  }
  void release() { rows = cols = 0; data = nullptr; buf_.reset(); }
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elemSize()); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

 private:
  int type_ = CV_8U;
  std::shared_ptr<unsigned char> buf_;
};

// compile-only stand-ins (see the header comment)
class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0; }
  operator float() const { return 0; }
  operator std::string() const { return std::string(); }
};

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T>
inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
