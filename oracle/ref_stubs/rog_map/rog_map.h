// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// GridMap (src/map) holds a rog_map::ROGMap for its `use_rog: true` branches. The oracle/_ref build
// compiles GridMap with `use_rog: false` (params/grid_map.yaml:3, the benchmark configuration), so this
// stand-in only has to satisfy the compiler: every method aborts if it is ever reached.
#pragma once
#include <cstdlib>
#include <memory>
#include <Eigen/Eigen>
#include "ros/ros.h"
#include "sensor_msgs/PointCloud2.h"
namespace rog_map {
struct ESDFMapUnavailable {
    [[noreturn]] static void die() { std::abort(); }
    void evaluateEDT(const Eigen::Vector3d&, double&) { die(); }
    void evaluateFirstGrad(const Eigen::Vector3d&, Eigen::Vector3d&) { die(); }
    void getValueGrad(const Eigen::Vector3d&, double&, Eigen::Vector3d&) const { die(); }
    void getCriticalValueGrad(const Eigen::Vector3d&, double&, Eigen::Vector3d&) const { die(); }
    void getValueGrad2d(const Eigen::Vector3d&, double&, Eigen::Vector3d&) const { die(); }
    bool isLineFree2d(const Eigen::Vector2d&, const Eigen::Vector2d&, double = 0.0) const { die(); }
    Eigen::Vector3i getHalfMapSize() const { die(); }
    void posToGlobalIndex(const Eigen::Vector3d&, Eigen::Vector3i&) const { die(); }
    void globalIndexToPos(const Eigen::Vector3i&, Eigen::Vector3d&) const { die(); }
    void globalIndexToLocalIndex(const Eigen::Vector3i&, Eigen::Vector3i&) const { die(); }
    void localIndexToPos(const Eigen::Vector3i&, Eigen::Vector3d&) const { die(); }
};
class ROGMap {
public:
    typedef std::shared_ptr<ROGMap> Ptr;
    explicit ROGMap(ros::NodeHandle&) : esdf_map_(new ESDFMapUnavailable()) {}
    std::shared_ptr<ESDFMapUnavailable> esdf_map_;
    sensor_msgs::PointCloud2 public_occ, public_sdf;
    double getResolution() const { ESDFMapUnavailable::die(); }
    Eigen::Vector3d getLocalMapOrigin() const { ESDFMapUnavailable::die(); }
};
}
