// opencv2/core.hpp STAND-IN (test infrastructure): the cv::Mat subset the reference's NID path touches.
// Mat(rows, cols, type, Scalar) fills with saturate_cast<T>(value), which for float is a plain (float) conversion
// (DBL_MAX -> +inf), as in OpenCV.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_64FC1 6

namespace cv {

struct Scalar {
  double v[4];
  static Scalar all(double x) { return Scalar{{x, x, x, x}}; }
};

class Mat {
public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;

  Mat() {}
  Mat(int rows, int cols, int type, const Scalar& s) : rows(rows), cols(cols) {
    const size_t es = elem_size(type);
    step = es * cols;
    own = std::make_shared<std::vector<unsigned char>>(step * rows);
    data = own->data();
    for (int r = 0; r < rows; r++) {
      for (int c = 0; c < cols; c++) {
        unsigned char* p = data + r * step + c * es;
        if (type == CV_32FC1) {
          const float f = static_cast<float>(s.v[0]);
          std::memcpy(p, &f, 4);
        } else if (type == CV_32SC1) {
          const int i = static_cast<int>(s.v[0]);
          std::memcpy(p, &i, 4);
        } else if (type == CV_64FC1) {
          const double d = s.v[0];
          std::memcpy(p, &d, 8);
        } else {
          *p = static_cast<unsigned char>(s.v[0]);
        }
      }
    }
  }
  // header over caller-owned pixels
  Mat(int rows, int cols, int /*type*/, void* ext, size_t step) : rows(rows), cols(cols), data(static_cast<unsigned char*>(ext)), step(step) {}

  // u8 -> f64 with a scale (the one conversion the reference's BFGS branch asks for)
  void convertTo(Mat& out, int type, double scale) const {
    out = Mat(rows, cols, type, Scalar::all(0.0));
    for (int r = 0; r < rows; r++) {
      for (int c = 0; c < cols; c++) out.at<double>(r, c) = at<unsigned char>(r, c) * scale;
    }
  }
  Mat clone() const {
    Mat o;
    o.rows = rows, o.cols = cols, o.step = step;
    o.own = std::make_shared<std::vector<unsigned char>>(data, data + step * rows);
    o.data = o.own->data();
    return o;
  }
  template <class T>
  T& at(int r, int c) {
    return *reinterpret_cast<T*>(data + r * step + c * sizeof(T));
  }
  template <class T>
  const T& at(int r, int c) const {
    return *reinterpret_cast<const T*>(data + r * step + c * sizeof(T));
  }

private:
  static size_t elem_size(int type) { return type == CV_8UC1 ? 1 : (type == CV_64FC1 ? 8 : 4); }
  std::shared_ptr<std::vector<unsigned char>> own;
};

}  // namespace cv
