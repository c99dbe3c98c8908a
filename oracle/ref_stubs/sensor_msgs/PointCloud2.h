// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for the ROS message — a flat list of (x, y, z, intensity).
#pragma once
#include <vector>
#include "visualization_msgs/Marker.h"
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; std::vector<float> xyzi; }; }
