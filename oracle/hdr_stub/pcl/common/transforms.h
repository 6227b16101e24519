// Stand-in for <pcl/common/transforms.h>: declared here, defined in oracle/mapmerging_ref_shim.cpp with the CPU checker's
// point transform.
#pragma once
#include <map_merge_3d/typedefs.h>
namespace pcl
{
void transformPointCloud(const map_merge_3d::PointCloud& in, map_merge_3d::PointCloud& out, const Eigen::Matrix4f& t);
}
