// Stand-ins for the PCL types the reference's matching.cpp / dispatch_descriptors.h name.  Descriptor structs have the
// layout of PCL's (the histogram first), estimators are only declared, the kd-tree is a brute-force search with the CPU
// checker's arithmetic (sequential float accumulation, ties to the lower index, k clamped to the cloud size like
// pcl::KdTreeFLANN).  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

namespace pcl
{
struct PFHSignature125 { float histogram[125]; };
struct PFHRGBSignature250 { float histogram[250]; };
struct FPFHSignature33 { float histogram[33]; };
struct PrincipalRadiiRSD { float r_min, r_max; };
struct SHOT1344 { float descriptor[1344]; float rf[9]; };
struct ShapeContext1980 { float descriptor[1980]; float rf[9]; };
// dimensions of pcl::DefaultPointRepresentation<T>
template <typename T> struct desc_dim;
template <> struct desc_dim<PFHSignature125> { static const int value = 125; };
template <> struct desc_dim<PFHRGBSignature250> { static const int value = 250; };
template <> struct desc_dim<FPFHSignature33> { static const int value = 33; };
template <> struct desc_dim<PrincipalRadiiRSD> { static const int value = 2; };
template <> struct desc_dim<SHOT1344> { static const int value = 1344; };
template <> struct desc_dim<ShapeContext1980> { static const int value = 1980; };

// PointCloud2 field name of each descriptor type (what pcl::toPCLPointCloud2 writes as fields[0].name)
template <typename T> struct desc_field;
template <> struct desc_field<PFHSignature125> { static const char* name() { return "pfh"; } };
template <> struct desc_field<PFHRGBSignature250> { static const char* name() { return "pfhrgb"; } };
template <> struct desc_field<FPFHSignature33> { static const char* name() { return "fpfh"; } };
template <> struct desc_field<PrincipalRadiiRSD> { static const char* name() { return "r_min"; } };
template <> struct desc_field<SHOT1344> { static const char* name() { return "shot"; } };
template <> struct desc_field<ShapeContext1980> { static const char* name() { return "shape_context"; } };

template <typename T>
struct PointCloud {
  std::vector<T> points;
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
  size_t size() const { return points.size(); }
  typename std::vector<T>::iterator begin() { return points.begin(); }
  typename std::vector<T>::iterator end() { return points.end(); }
  typename std::vector<T>::const_iterator begin() const { return points.begin(); }
  typename std::vector<T>::const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& o)
  {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
typedef std::shared_ptr<std::vector<int>> IndicesPtr;

// the point types of the path (layouts: x, y, z, packed colour / normal + curvature — 16 bytes like the CPU checker's)
struct PointXYZRGB {
  float x, y, z;
  uint32_t rgba;
};
struct Normal {
  float normal_x, normal_y, normal_z, curvature;
};
struct PointWithScale {
  float x, y, z, scale;
};
struct PointXYZI {
  float x, y, z, intensity;
};

struct Correspondence {
  int index_query = 0, index_match = -1;
  float distance = 0.f;
  Correspondence() {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;

namespace search
{
template <typename T>
class KdTree
{
public:
  void setInputCloud(const typename PointCloud<T>::Ptr& c) { cloud_ = c; }
  void setSortedResults(bool) {}
  int nearestKSearch(const PointCloud<T>& query, int index, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const
  {
    const int D = desc_dim<T>::value;
    const int n = (int)cloud_->points.size();
    if (k > n) k = n;  // pcl::KdTreeFLANN::nearestKSearch
    k_indices.resize(k);
    k_sqr_distances.resize(k);
    const float* a = reinterpret_cast<const float*>(&query.points[index]);
    std::vector<std::pair<float, int>> best;
    for (int j = 0; j < n; ++j) {
      const float* b = reinterpret_cast<const float*>(&cloud_->points[j]);
      float r = 0.f;
      for (int t = 0; t < D; ++t) {
        const float diff = a[t] - b[t];
        r += diff * diff;
      }
      const std::pair<float, int> cand(r, j);
      if ((int)best.size() < k) best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
      else if (k > 0 && cand < best.back()) {
        best.pop_back();
        best.insert(std::upper_bound(best.begin(), best.end(), cand), cand);
      }
    }
    for (int t = 0; t < k; ++t) {
      k_indices[t] = best[t].second;
      k_sqr_distances[t] = best[t].first;
    }
    return k;
  }

private:
  typename PointCloud<T>::Ptr cloud_;
};
}  // namespace search
}  // namespace pcl
