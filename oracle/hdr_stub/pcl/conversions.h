// Stand-in for <pcl/conversions.h>: typed descriptor cloud <-> LocalDescriptors (the PCLPointCloud2 stand-in).
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
template <typename T>
void toPCLPointCloud2(const PointCloud<T>& in, map_merge_3d::LocalDescriptors& out)
{
  const int D = desc_dim<T>::value;
  out.fields.clear();
  out.fields.push_back(map_merge_3d::PCLPointField{desc_field<T>::name()});
  out.dim = D;
  out.v.resize(in.points.size() * (size_t)D);
  for (size_t i = 0; i < in.points.size(); ++i) std::memcpy(&out.v[i * (size_t)D], &in.points[i], sizeof(float) * (size_t)D);
}
}  // namespace pcl
