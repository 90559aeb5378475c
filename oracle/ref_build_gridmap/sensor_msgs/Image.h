// sensor_msgs/Image.h — stand-in (TEST INFRASTRUCTURE): the fields grid_map.cpp reads.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <std_msgs/Header.h>

namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint32_t step = 0;
  std::vector<uint8_t> data;
  typedef std::shared_ptr<Image> Ptr;
  typedef std::shared_ptr<const Image> ConstPtr;
};
typedef std::shared_ptr<Image> ImagePtr;
typedef std::shared_ptr<const Image> ImageConstPtr;
}  // namespace sensor_msgs
