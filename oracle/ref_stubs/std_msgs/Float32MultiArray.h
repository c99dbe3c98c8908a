// ORACLE — TEST INFRASTRUCTURE ONLY: plain-struct stand-in for the ROS message.
#pragma once
#include <string>
#include <vector>
namespace std_msgs {
struct MultiArrayDimension { std::string label; unsigned size = 0, stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned data_offset = 0; };
struct Float32MultiArray { MultiArrayLayout layout; std::vector<float> data; };
}
