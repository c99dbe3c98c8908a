// ORACLE — TEST INFRASTRUCTURE ONLY: plain-struct stand-in for the ROS message (see ros/ros.h here).
#pragma once
#include "visualization_msgs/Marker.h"
namespace visualization_msgs {
struct MarkerArray { std::vector<Marker> markers; };
}
