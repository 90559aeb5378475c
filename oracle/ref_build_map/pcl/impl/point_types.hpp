// stand-in: see pcl/point_types.h
#include <pcl/point_types.h>
