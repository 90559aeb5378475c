// std_msgs/ColorRGBA.h — a STAND-IN.  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_STD_COLOR_STANDIN_H
#define POP_STD_COLOR_STANDIN_H
namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
#endif
