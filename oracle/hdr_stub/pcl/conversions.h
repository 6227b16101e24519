// Stand-in for <pcl/conversions.h>: LocalDescriptors (the PCLPointCloud2 stand-in) -> typed descriptor cloud.
#pragma once
#include <map_merge_3d/typedefs.h>
namespace pcl
{
template <typename T>
void fromPCLPointCloud2(const map_merge_3d::LocalDescriptors& in, PointCloud<T>& out)
{
  const int D = desc_dim<T>::value;
  const size_t n = in.dim > 0 ? in.v.size() / (size_t)in.dim : 0;
  out.points.assign(n, T());
  for (size_t i = 0; i < n; ++i) {
    std::memset(&out.points[i], 0, sizeof(T));
    std::memcpy(&out.points[i], in.v.data() + i * (size_t)in.dim, sizeof(float) * (size_t)std::min(D, in.dim));
  }
}
}  // namespace pcl
