// std_msgs/Header.h — a STAND-IN (plain struct with the message's fields).  TEST INFRASTRUCTURE (oracle/_ref only).
#ifndef POP_STD_HEADER_STANDIN_H
#define POP_STD_HEADER_STANDIN_H
#include <geometry_msgs/PolygonStamped.h>
namespace std_msgs { typedef geometry_msgs::Header Header; }
#endif
