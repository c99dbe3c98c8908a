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
// test knobs (see oracle/ref_driver_full.cpp): a frozen clock and a per-thread budget of ros::ok() calls
inline bool& stub_clock_frozen() {
    static bool frozen = false;
    return frozen;
}
inline int& stub_ok_budget() {
    static thread_local int budget = -1;   // < 0: unlimited
    return budget;
}
struct Time {
    double s = 0.0;
    static Time now() {
        Time t;
        if (!stub_clock_frozen())
            t.s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        return t;
    }
    double toSec() const { return s; }
    Duration operator-(const Time& o) const { return Duration(s - o.s); }
};
inline bool ok() {
    int& b = stub_ok_budget();
    if (b < 0) return true;
    if (b == 0) return false;
    b--;
    return true;
}
inline void spinOnce() {}

// parameter table: name -> list of doubles (scalars are one-element lists)
inline std::map<std::string, std::vector<double>>& stub_params() {
    static std::map<std::string, std::vector<double>> t;
    return t;
}
struct Publisher {
    template <class M> void publish(const M&) const {}
    int getNumSubscribers() const { return 0; }
};
struct Subscriber {};
struct Timer {};
struct TimerEvent {};
inline std::map<std::string, std::string>& stub_string_params() {
    static std::map<std::string, std::string> t;
    return t;
}
class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string&) {}
    std::string getNamespace() const { return ""; }
    bool hasParam(const std::string& name) const { return stub_params().count(name) || stub_string_params().count(name); }
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
    bool getParam(const std::string& name, std::string& out) const {
        auto it = stub_string_params().find(name);
        if (it == stub_string_params().end()) return false;
        out = it->second;
        return true;
    }
    template <class T>
    bool param(const std::string& name, T& out, const T& dflt) const {
        if (getParam(name, out)) return true;
        out = dflt;
        return false;
    }
    template <class M>
    Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
    template <class... A>
    Subscriber subscribe(const std::string&, int, A&&...) { return Subscriber(); }
    template <class... A>
    Timer createTimer(Duration, A&&...) { return Timer(); }
};
}  // namespace ros

#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
