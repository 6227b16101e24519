// features.h — low-level feature interface, same signatures as
// map_merge_3d/include/map_merge_3d/features.h:20-98, implemented on libmm3d (CUDA, no CPU fallback).
#ifndef MM3D_SHIM_FEATURES_H_
#define MM3D_SHIM_FEATURES_H_

#include <map_merge_3d/enum.h>
#include <map_merge_3d/typedefs.h>

namespace map_merge_3d
{
#define MM3D_DESCRIPTOR_LIST(X) X(PFH) X(PFHRGB) X(FPFH) X(RSD) X(SHOT) X(SC3D)
MM3D_ENUM_CLASS(Descriptor, MM3D_DESCRIPTOR_LIST)
#define MM3D_KEYPOINT_LIST(X) X(SIFT) X(HARRIS)
MM3D_ENUM_CLASS(Keypoint, MM3D_KEYPOINT_LIST)

PointCloudPtr downSample(const PointCloudConstPtr& input, double resolution);
PointCloudPtr removeOutliers(const PointCloudConstPtr& input, double radius, int min_neighbours);
PointCloudPtr detectKeypoints(const PointCloudConstPtr& points, const SurfaceNormalsPtr& normals, Keypoint type, double threshold,
                              double radius, double resolution);
/// keypoints whose descriptor cannot be computed are removed from `keypoints` (features.cpp:119-141)
LocalDescriptorsPtr computeLocalDescriptors(const PointCloudConstPtr& points, const SurfaceNormalsPtr& normals,
                                            const PointCloudPtr& keypoints, Descriptor descriptor, double feature_radius);
SurfaceNormalsPtr computeSurfaceNormals(const PointCloudConstPtr& input, double radius);
}  // namespace map_merge_3d

#endif  // MM3D_SHIM_FEATURES_H_
