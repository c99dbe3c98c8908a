// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for pcl::PointCloud — a std::vector of points with PCL's member names.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader { std::string frame_id; std::uint64_t stamp = 0; };
template <class P>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
    std::vector<P> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;
    PCLHeader header;
    void push_back(const P& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void resize(std::size_t n) { points.resize(n); }
    void reserve(std::size_t n) { points.reserve(n); }
    P& operator[](std::size_t i) { return points[i]; }
    const P& operator[](std::size_t i) const { return points[i]; }
    P& at(std::size_t i) { return points.at(i); }
    const P& at(std::size_t i) const { return points.at(i); }
    typename std::vector<P>::iterator begin() { return points.begin(); }
    typename std::vector<P>::iterator end() { return points.end(); }
    typename std::vector<P>::const_iterator begin() const { return points.begin(); }
    typename std::vector<P>::const_iterator end() const { return points.end(); }
    PointCloud& operator+=(const PointCloud& o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (std::uint32_t)points.size();
        height = 1;
        return *this;
    }
    PointCloud operator+(const PointCloud& o) const { PointCloud r = *this; r += o; return r; }
};
namespace console { enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE }; inline void setVerbosityLevel(int) {} }
}
