// cv_bridge/cv_bridge.h — stand-in (TEST INFRASTRUCTURE) for the two things plan_env/src/grid_map.cpp takes from
// OpenCV / cv_bridge: a 16-bit single-channel image (cv::Mat with rows, cols, ptr<uint16_t>(row), copyTo, convertTo)
// and toCvCopy of a sensor_msgs::Image.  Written here; nothing of OpenCV is copied.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <sensor_msgs/Image.h>

#define CV_16UC1 2
#define CV_32FC1 5

namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  int type_ = CV_16UC1;
  std::shared_ptr<std::vector<uint8_t>> buf;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    buf = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(r) * c * elem(), 0);
  }
  size_t elem() const { return type_ == CV_32FC1 ? 4 : 2; }
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return type_; }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf->data() + static_cast<size_t>(r) * cols * elem()); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf->data() + static_cast<size_t>(r) * cols * elem()); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  void copyTo(Mat& o) const {
    o.rows = rows; o.cols = cols; o.type_ = type_;
    o.buf = buf ? std::make_shared<std::vector<uint8_t>>(*buf) : nullptr;
  }
  // the one conversion grid_map.cpp asks for: 32-bit float metres -> 16-bit with a scale (saturate_cast rounding)
  void convertTo(Mat& o, int type, double alpha = 1.0) const {
    Mat r(rows, cols, type);
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) {
        const double v = (type_ == CV_32FC1 ? static_cast<double>(at<float>(i, j)) : static_cast<double>(at<uint16_t>(i, j))) * alpha;
        if (type == CV_16UC1) {
          const double q = std::nearbyint(v);
          r.at<uint16_t>(i, j) = static_cast<uint16_t>(q < 0 ? 0 : q > 65535 ? 65535 : q);
        } else {
          r.at<float>(i, j) = static_cast<float>(v);
        }
      }
    o = r;
  }
};
}  // namespace cv

namespace sensor_msgs { namespace image_encodings {
static const std::string TYPE_32FC1 = "32FC1";
static const std::string TYPE_16UC1 = "16UC1";
} }

namespace cv_bridge {
struct CvImage {
  cv::Mat image;
  std::string encoding;
};
typedef std::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& img, const std::string& = std::string()) {
  CvImagePtr p = std::make_shared<CvImage>();
  p->encoding = img->encoding;
  p->image.create(static_cast<int>(img->height), static_cast<int>(img->width),
                  img->encoding == sensor_msgs::image_encodings::TYPE_32FC1 ? CV_32FC1 : CV_16UC1);
  std::copy(img->data.begin(), img->data.end(), p->image.buf->begin());
  return p;
}
}  // namespace cv_bridge
