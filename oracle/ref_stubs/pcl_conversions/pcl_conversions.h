// ORACLE — TEST INFRASTRUCTURE ONLY: cloud <-> message copies for the stand-in types.
#pragma once
#include "pcl/point_cloud.h"
#include "pcl/point_types.h"
#include "sensor_msgs/PointCloud2.h"
namespace pcl {
inline float stub_intensity(const PointXYZ&) { return 0.f; }
inline float stub_intensity(const PointXYZI& p) { return p.intensity; }
inline void stub_set_intensity(PointXYZ&, float) {}
inline void stub_set_intensity(PointXYZI& p, float v) { p.intensity = v; }
template <class P>
void toROSMsg(const PointCloud<P>& pc, sensor_msgs::PointCloud2& msg) {
    msg.xyzi.clear();
    for (const P& p : pc.points) { msg.xyzi.push_back(p.x); msg.xyzi.push_back(p.y); msg.xyzi.push_back(p.z); msg.xyzi.push_back(stub_intensity(p)); }
}
template <class P>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<P>& pc) {
    pc.clear();
    for (std::size_t i = 0; i + 3 < msg.xyzi.size(); i += 4) {
        P p; p.x = msg.xyzi[i]; p.y = msg.xyzi[i + 1]; p.z = msg.xyzi[i + 2]; stub_set_intensity(p, msg.xyzi[i + 3]);
        pc.push_back(p);
    }
}
}
