// Stand-in for the reference's <map_merge_3d/typedefs.h> (which needs PCL) so that the REFERENCE's own public headers, its
// map_merging.cpp and its matching.cpp compile here unmodified (see ../params_ref_shim.cpp, ../mapmerging_ref_shim.cpp).
// Plain containers with the few members that code touches.  Test infrastructure only.
#ifndef MAP_MERGE_TYPEDEFS_H_
#define MAP_MERGE_TYPEDEFS_H_
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <pcl/stub_types.h>

namespace map_merge_3d
{
struct PointT {
  float x, y, z;
  uint32_t rgba;
};
struct NormalT {
  float normal_x, normal_y, normal_z, curvature;
};
struct PointCloud {
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  PointCloud& operator+=(const PointCloud& o)
  {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
typedef std::shared_ptr<PointCloud> PointCloudPtr;
typedef std::shared_ptr<const PointCloud> PointCloudConstPtr;
struct SurfaceNormals {
  std::vector<float> v;  // n x (nx, ny, nz, curvature)
};
typedef std::shared_ptr<SurfaceNormals> SurfaceNormalsPtr;
typedef std::shared_ptr<const SurfaceNormals> SurfaceNormalsConstPtr;
struct PCLPointField {
  std::string name;
};
struct LocalDescriptors {          // pcl::PCLPointCloud2: fields[0].name selects the descriptor type
  std::vector<PCLPointField> fields;
  std::vector<float> v;            // n x dim
  int dim = 0;
};
typedef std::shared_ptr<LocalDescriptors> LocalDescriptorsPtr;
typedef std::shared_ptr<const LocalDescriptors> LocalDescriptorsConstPtr;
using pcl::Correspondences;
using pcl::CorrespondencesPtr;
}  // namespace map_merge_3d
#endif
