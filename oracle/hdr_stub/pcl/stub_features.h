// Stand-ins for the PCL filter / keypoint / feature classes features.cpp configures and runs.  They record the settings;
// the work is done by functions of the CPU checker (declared here, defined in oracle/mapmerging_ref_shim.cpp).
// Test infrastructure only.
#pragma once
#include <cassert>  // the real PCL headers pull it in; features.cpp relies on that
#include <cmath>

#include <map_merge_3d/typedefs.h>
namespace pcl
{
namespace stub
{
typedef PointCloud<PointXYZRGB> Cloud;
typedef PointCloud<Normal> Normals;
void voxel_grid(const Cloud& in, float lx, float ly, float lz, Cloud& out);
void radius_outlier_removal(const Cloud& in, double radius, int min_neighbours, Cloud& out);
void normal_estimation(const Cloud& in, double radius, Normals& out);
void sift_keypoints(const Cloud& in, float min_scale, int nr_octaves, int nr_scales_per_octave, float min_contrast, PointCloud<PointWithScale>& out);
void harris_keypoints(const Cloud& in, const Normals& normals, bool non_max, bool refine, float threshold, float radius, PointCloud<PointXYZI>& out);
// kind: 0 PFH, 1 PFHRGB, 2 FPFH, 3 RSD, 4 SHOT colour, 5 3DSC.  out = keypoints.size() x dim, rows PCL could not compute are NaN
void descriptors(int kind, const Cloud& surface, const Normals& normals, const Cloud& keypoints, double radius, int dim, std::vector<float>& out);
}  // namespace stub

template <typename P>
class VoxelGrid
{
public:
  void setLeafSize(float x, float y, float z) { lx_ = x; ly_ = y; lz_ = z; }
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void filter(PointCloud<P>& out) { stub::voxel_grid(*in_, lx_, ly_, lz_, out); }

private:
  float lx_ = 0, ly_ = 0, lz_ = 0;
  typename PointCloud<P>::ConstPtr in_;
};
template <typename P>
class RadiusOutlierRemoval
{
public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void setRadiusSearch(double r) { r_ = r; }
  void setMinNeighborsInRadius(int m) { m_ = m; }
  void filter(PointCloud<P>& out) { stub::radius_outlier_removal(*in_, r_, m_, out); }

private:
  double r_ = 0;
  int m_ = 1;
  typename PointCloud<P>::ConstPtr in_;
};
template <typename P, typename N>
class NormalEstimation
{
public:
  void setRadiusSearch(double r) { r_ = r; }
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void compute(PointCloud<N>& out) { stub::normal_estimation(*in_, r_, out); }

private:
  double r_ = 0;
  typename PointCloud<P>::ConstPtr in_;
};
template <typename P, typename K>
class SIFTKeypoint
{
public:
  void setScales(float min_scale, int nr_octaves, int nr_scales_per_octave) { s_ = min_scale; o_ = nr_octaves; n_ = nr_scales_per_octave; }
  void setMinimumContrast(float c) { c_ = c; }
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void compute(PointCloud<K>& out) { stub::sift_keypoints(*in_, s_, o_, n_, c_, out); }

private:
  float s_ = 0, c_ = 0;
  int o_ = 0, n_ = 0;
  typename PointCloud<P>::ConstPtr in_;
};
template <typename P, typename K>
class HarrisKeypoint3D
{
public:
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void setNormals(const typename PointCloud<Normal>::ConstPtr& n) { nm_ = n; }
  void setNonMaxSupression(bool v) { nms_ = v; }
  void setRefine(bool v) { refine_ = v; }
  void setThreshold(float v) { thr_ = v; }
  void setRadius(float v) { rad_ = v; }
  void compute(PointCloud<K>& out) { stub::harris_keypoints(*in_, *nm_, nms_, refine_, thr_, rad_, out); }

private:
  bool nms_ = false, refine_ = true;
  float thr_ = 0, rad_ = 0;
  typename PointCloud<P>::ConstPtr in_;
  typename PointCloud<Normal>::ConstPtr nm_;
};
// copyPointCloud between point types: the common fields (x, y, z) are copied, the rest is default-constructed
// (pcl::PointXYZRGB's default colour is opaque black: rgba = 0xff000000)
template <typename A>
void copyPointCloud(const PointCloud<A>& in, PointCloud<PointXYZRGB>& out)
{
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); ++i) {
    out.points[i].x = in.points[i].x;
    out.points[i].y = in.points[i].y;
    out.points[i].z = in.points[i].z;
    out.points[i].rgba = 0xff000000u;
  }
}

template <typename T>
class DefaultPointRepresentation
{
public:
  bool isValid(const T& p) const
  {
    const float* f = reinterpret_cast<const float*>(&p);
    for (int i = 0; i < desc_dim<T>::value; ++i)
      if (!std::isfinite(f[i])) return false;
    return true;
  }
};
template <typename T>
class ExtractIndices
{
public:
  void setInputCloud(const typename PointCloud<T>::ConstPtr& c) { in_ = c; }
  void setIndices(const IndicesPtr& i) { idx_ = i; }
  void setNegative(bool n) { neg_ = n; }
  void filter(PointCloud<T>& out)
  {
    std::vector<char> mark(in_->points.size(), 0);
    for (int i : *idx_) mark[i] = 1;
    std::vector<T> keep;
    for (size_t i = 0; i < in_->points.size(); ++i)
      if ((mark[i] != 0) != neg_) keep.push_back(in_->points[i]);
    out.points.swap(keep);  // in-place use (input == output) is how features.cpp calls it
  }

private:
  typename PointCloud<T>::ConstPtr in_;
  IndicesPtr idx_;
  bool neg_ = false;
};

namespace stub
{
template <typename In, typename N, typename Out, int KIND>
class DescriptorEstimator
{
public:
  void setRadiusSearch(double r) { r_ = r; }
  void setSearchSurface(const typename PointCloud<In>::ConstPtr& c) { surface_ = c; }
  void setInputNormals(const typename PointCloud<N>::ConstPtr& n) { normals_ = n; }
  void setInputCloud(const typename PointCloud<In>::ConstPtr& c) { in_ = c; }
  void compute(PointCloud<Out>& out)
  {
    const int D = desc_dim<Out>::value;
    std::vector<float> v;
    descriptors(KIND, *surface_, *normals_, *in_, r_, D, v);
    out.points.assign(in_->points.size(), Out());
    for (size_t i = 0; i < in_->points.size(); ++i) {
      std::memset(&out.points[i], 0, sizeof(Out));
      std::memcpy(&out.points[i], &v[i * (size_t)D], sizeof(float) * (size_t)D);
    }
  }

private:
  double r_ = 0;
  typename PointCloud<In>::ConstPtr surface_, in_;
  typename PointCloud<N>::ConstPtr normals_;
};
}  // namespace stub
template <typename In, typename N, typename Out> class PFHEstimation : public stub::DescriptorEstimator<In, N, Out, 0> {};
template <typename In, typename N, typename Out> class PFHRGBEstimation : public stub::DescriptorEstimator<In, N, Out, 1> {};
template <typename In, typename N, typename Out> class FPFHEstimation : public stub::DescriptorEstimator<In, N, Out, 2> {};
template <typename In, typename N, typename Out> class RSDEstimation : public stub::DescriptorEstimator<In, N, Out, 3> {};
template <typename In, typename N, typename Out> class SHOTColorEstimation : public stub::DescriptorEstimator<In, N, Out, 4> {};
template <typename In, typename N, typename Out> class ShapeContext3DEstimation : public stub::DescriptorEstimator<In, N, Out, 5> {};
}  // namespace pcl
