// Stand-in for the reference's scanner/util/opencv.h, used ONLY by oracle/Makefile to compile the reference's
// tests/test_ops.cpp *unmodified* into oracle/_ref/libref_test_ops.so (test infrastructure, like everything
// under oracle/).  OpenCV's C++ headers do not exist in this image, so the handful of cv:: names that file
// mentions are declared here and abort when called: the kernels whose arithmetic is written out in that file
// (Blur, TestIncrement*, Sleep) are what the library is built for -- they pin oracle.blur and the engine's
// state / warm-up semantics to the reference's own compiled code.  Histogram / Resize / OpticalFlow of that
// file call OpenCV and are pinned through cv2 golden vectors instead (oracle/make_golden.py).
#pragma once
#include <unistd.h>  // the reference file calls sleep(); upstream it arrives through OpenCV / glog headers

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "scanner/api/frame.h"
#include "scanner/api/kernel.h"
#include "scanner/util/common.h"

#define CV_32SC1 4
#define CV_8UC1 0
namespace cv {
[[noreturn]] inline void no_opencv(const char* what) {
  fprintf(stderr, "oracle/_ref/libref_test_ops.so: %s needs OpenCV, which this build does not have\n", what);
  abort();
}
struct Size {
  Size(int w = 0, int h = 0) : width(w), height(h) {}
  int width, height;
};
class Mat {
 public:
  Mat() {}
  Mat(int, int, int) {}
  Mat(int, int, int, void*) {}
  void convertTo(Mat&, int) const { no_opencv("cv::Mat::convertTo"); }
};
inline void calcHist(const Mat*, int, const int*, const Mat&, Mat&, int, const int*, const float**) {
  no_opencv("cv::calcHist");
}
inline void resize(const Mat&, Mat&, Size) { no_opencv("cv::resize"); }
enum { COLOR_BGR2GRAY = 6 };
inline void cvtColor(const Mat&, Mat&, int) { no_opencv("cv::cvtColor"); }
template <typename T>
class Ptr {
 public:
  Ptr() {}
  T* operator->() const { no_opencv("cv::Ptr"); }
};
class DenseOpticalFlow {
 public:
  void calc(const Mat&, const Mat&, Mat&) { no_opencv("cv::DenseOpticalFlow::calc"); }
};
class FarnebackOpticalFlow : public DenseOpticalFlow {
 public:
  static Ptr<DenseOpticalFlow> create(int, double, bool, int, int, int, double, int) { return Ptr<DenseOpticalFlow>(); }
};
}  // namespace cv

namespace scanner {
inline cv::Mat frame_to_mat(const Frame*) { return cv::Mat(); }
inline cv::Mat frame_to_mat(Frame*) { return cv::Mat(); }
}  // namespace scanner
