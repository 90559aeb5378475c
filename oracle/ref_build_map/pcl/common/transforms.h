// stand-in: named by plan_env/map.h, nothing of it is used on the path
#include <pcl/point_cloud.h>
