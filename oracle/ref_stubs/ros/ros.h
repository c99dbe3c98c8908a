// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Just enough of roscpp's surface for the reference's hot-path headers to compile without ROS
// (oracle/Makefile, target _ref). Nothing here talks to a ROS master: parameters come from a
// process-local table the driver fills, time is the steady clock.
#pragma once
#include <chrono>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace ros {
struct Duration {
    double s = 0.0;
    Duration() {}
    explicit Duration(double v) : s(v) {}
    double toSec() const { return s; }
};
struct Time {
    double s = 0.0;
    static Time now() {
        Time t;
        t.s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        return t;
    }
    double toSec() const { return s; }
    Duration operator-(const Time& o) const { return Duration(s - o.s); }
};
inline bool ok() { return true; }

// parameter table: name -> list of doubles (scalars are one-element lists)
inline std::map<std::string, std::vector<double>>& stub_params() {
    static std::map<std::string, std::vector<double>> t;
    return t;
}
template <class M>
struct Publisher {
    void publish(const M&) const {}
};
struct AnyPublisher {
    template <class M> void publish(const M&) const {}
    int getNumSubscribers() const { return 0; }
};
class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string&) {}
    template <class T>
    bool getParam(const std::string& name, T& out) const {
        auto it = stub_params().find(name);
        if (it == stub_params().end() || it->second.empty()) return false;
        out = static_cast<T>(it->second[0]);
        return true;
    }
    template <class T>
    bool getParam(const std::string& name, std::vector<T>& out) const {
        auto it = stub_params().find(name);
        if (it == stub_params().end()) return false;
        out.assign(it->second.begin(), it->second.end());
        return true;
    }
    bool getParam(const std::string&, std::string&) const { return false; }
    template <class T>
    bool param(const std::string& name, T& out, const T& dflt) const {
        if (getParam(name, out)) return true;
        out = dflt;
        return false;
    }
    template <class M>
    AnyPublisher advertise(const std::string&, int, bool = false) { return AnyPublisher(); }
};
}  // namespace ros

#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
