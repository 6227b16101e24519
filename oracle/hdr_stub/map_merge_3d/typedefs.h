// Stand-in for the reference's <map_merge_3d/typedefs.h> (which needs PCL) so that the REFERENCE's own public headers
// — enum.h, features.h, matching.h, map_merging.h — compile here unmodified (see ../params_ref_shim.cpp).  Only opaque
// handles: nothing in those headers looks inside the types.  Test infrastructure only.
#ifndef MAP_MERGE_TYPEDEFS_H_
#define MAP_MERGE_TYPEDEFS_H_
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
namespace Eigen
{
struct Matrix4f {
  float m[16];
};
}  // namespace Eigen
namespace map_merge_3d
{
struct PointCloud {};
typedef std::shared_ptr<PointCloud> PointCloudPtr;
typedef std::shared_ptr<const PointCloud> PointCloudConstPtr;
struct SurfaceNormals {};
typedef std::shared_ptr<SurfaceNormals> SurfaceNormalsPtr;
typedef std::shared_ptr<const SurfaceNormals> SurfaceNormalsConstPtr;
struct LocalDescriptors {};
typedef std::shared_ptr<LocalDescriptors> LocalDescriptorsPtr;
typedef std::shared_ptr<const LocalDescriptors> LocalDescriptorsConstPtr;
struct Correspondences {};
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;
}  // namespace map_merge_3d
#endif
