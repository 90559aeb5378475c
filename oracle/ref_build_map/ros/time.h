// stand-in: see ros/ros.h
#include <ros/ros.h>
