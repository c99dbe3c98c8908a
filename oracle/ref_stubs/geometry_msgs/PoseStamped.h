// ORACLE — TEST INFRASTRUCTURE ONLY: plain-struct stand-in for the ROS message.
#pragma once
#include "visualization_msgs/Marker.h"
namespace geometry_msgs { struct PoseStamped { std_msgs::Header header; Pose pose; }; }
