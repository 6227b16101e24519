// matching.h — low-level pair-matching interface, same signatures as
// map_merge_3d/include/map_merge_3d/matching.h:26-152, implemented on libmm3d.
#ifndef MM3D_SHIM_MATCHING_H_
#define MM3D_SHIM_MATCHING_H_

#include <map_merge_3d/enum.h>
#include <map_merge_3d/typedefs.h>

namespace map_merge_3d
{
CorrespondencesPtr findFeatureCorrespondences(const LocalDescriptorsPtr& source_descriptors, const LocalDescriptorsPtr& target_descriptors,
                                              size_t k = 5);
/// zero matrix (and empty inliers) when no transformation could be estimated (matching.h:41-42)
Eigen::Matrix4f estimateTransformFromCorrespondences(const PointCloudPtr& source_keypoints, const PointCloudPtr& target_keypoints,
                                                     const CorrespondencesPtr& correspondences, CorrespondencesPtr& inliers,
                                                     double inlier_threshold);
Eigen::Matrix4f estimateTransformFromDescriptorsSets(const PointCloudPtr& source_keypoints, const LocalDescriptorsPtr& source_descriptors,
                                                     const PointCloudPtr& target_keypoints, const LocalDescriptorsPtr& target_descriptors,
                                                     double min_sample_distance, double max_correspondence_distance, int max_iterations);
Eigen::Matrix4f estimateTransformICP(const PointCloudPtr& source_points, const PointCloudPtr& target_points,
                                     const Eigen::Matrix4f& initial_guess, double max_correspondence_distance,
                                     double outlier_rejection_threshold, int max_iterations = 100, double transformation_epsilon = 0.0);

#define MM3D_METHOD_LIST(X) X(MATCHING) X(SAC_IA)
MM3D_ENUM_CLASS(EstimationMethod, MM3D_METHOD_LIST)

Eigen::Matrix4f estimateTransform(const PointCloudPtr& source_points, const PointCloudPtr& source_keypoints,
                                  const LocalDescriptorsPtr& source_descriptors, const PointCloudPtr& target_points,
                                  const PointCloudPtr& target_keypoints, const LocalDescriptorsPtr& target_descriptors,
                                  EstimationMethod method, bool refine, double inlier_threshold, double max_correspondence_distance,
                                  int max_iterations, size_t matching_k, double transform_epsilon);
double transformScore(const PointCloudPtr& source_points, const PointCloudPtr& target_points, const Eigen::Matrix4f& transform,
                      double max_distance);
}  // namespace map_merge_3d

#endif  // MM3D_SHIM_MATCHING_H_
