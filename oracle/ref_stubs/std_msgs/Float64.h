// ORACLE — TEST INFRASTRUCTURE ONLY: plain-struct stand-in for the ROS message.
#pragma once
namespace std_msgs { struct Float64 { double data = 0.0; }; }
