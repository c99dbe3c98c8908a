// ORACLE — TEST INFRASTRUCTURE ONLY: stand-in for roscpp's package lookup (see ros/ros.h here).
#pragma once
#include <string>
namespace ros { namespace package { inline std::string getPath(const std::string&) { return "/tmp"; } } }
