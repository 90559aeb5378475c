// std_msgs/Float32MultiArray.h — a STAND-IN (the fields RiskBase::futureRiskCallback reads).  TEST INFRASTRUCTURE.
#ifndef POP_STD_F32MA_STANDIN_H
#define POP_STD_F32MA_STANDIN_H
#include <memory>
#include <string>
#include <vector>
namespace std_msgs {
struct MultiArrayDimension { std::string label; unsigned size = 0, stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned data_offset = 0; };
struct Float32MultiArray { MultiArrayLayout layout; std::vector<float> data; };
typedef std::shared_ptr<const Float32MultiArray> Float32MultiArrayConstPtr;
}  // namespace std_msgs
#endif
