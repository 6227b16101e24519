// Stand-in for <ros/ros.h>: map_merging.cpp reads parameters through NodeHandle::getParam; here no parameter is ever set.
#pragma once
#include <string>
namespace ros
{
class NodeHandle
{
public:
  template <typename T>
  bool getParam(const std::string&, T&) const
  {
    return false;
  }
};
}  // namespace ros
