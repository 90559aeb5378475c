// ros/ros.h for the grid_map.cpp build — the shared stand-in plus boost::bind, which roscpp's headers bring in and
// grid_map.cpp uses without including it.  TEST INFRASTRUCTURE.
#pragma once
#include_next <ros/ros.h>
#include <boost/bind.hpp>
