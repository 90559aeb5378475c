// visualization_msgs/MarkerArray.h — a STAND-IN.  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_MARKER_ARRAY_STANDIN_H
#define POP_MARKER_ARRAY_STANDIN_H
#include <visualization_msgs/Marker.h>
#include <memory>
namespace visualization_msgs { struct MarkerArray { std::vector<Marker> markers; typedef std::shared_ptr<const MarkerArray> ConstPtr; }; }
#endif
