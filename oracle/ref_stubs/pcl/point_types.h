// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for the two PCL point types the reference uses.
#pragma once
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; PointXYZ() {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
struct PointXYZINormal { float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
}
