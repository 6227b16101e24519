// Stand-in for the reference's <map_merge_3d/typedefs.h> (which needs PCL) so that the REFERENCE's own public headers and
// its features.cpp, matching.cpp and map_merging.cpp compile here unmodified (see ../params_ref_shim.cpp,
// ../mapmerging_ref_shim.cpp).  Same typedef names over the stand-in containers of <pcl/stub_types.h>.
// Test infrastructure only.
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
typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> PointCloud;
typedef pcl::PointCloud<PointT>::Ptr PointCloudPtr;
typedef pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;
typedef pcl::Normal NormalT;
typedef pcl::PointCloud<NormalT> SurfaceNormals;
typedef pcl::PointCloud<NormalT>::Ptr SurfaceNormalsPtr;
typedef pcl::PointCloud<NormalT>::ConstPtr SurfaceNormalsConstPtr;
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
