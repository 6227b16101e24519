// Stand-in for <ros/ros.h>: map_merging.h only names ros::NodeHandle in a declaration.
#pragma once
namespace ros
{
class NodeHandle;
}
