// map_merging.h — high-level interface, same surface as
// map_merge_3d/include/map_merge_3d/map_merging.h:28-101 (MapMergingParams, estimateMapsTransforms,
// composeMaps), implemented on libmm3d.  fromROSNode is declared only when ROS headers exist.
#ifndef MM3D_SHIM_MAP_MERGING_H_
#define MM3D_SHIM_MAP_MERGING_H_

#include <ostream>
#include <vector>

#include <map_merge_3d/features.h>
#include <map_merge_3d/matching.h>
#include <map_merge_3d/typedefs.h>

#if defined(__has_include)
#if __has_include(<ros/ros.h>)
#include <ros/ros.h>
#define MM3D_HAVE_ROS 1
#endif
#endif

namespace map_merge_3d
{
struct MapMergingParams {
  // dependent defaults are evaluated once, at resolution 0.1 (map_merging.h:29-39)
  double resolution = 0.1;
  double descriptor_radius = resolution * 8.0;
  int outliers_min_neighbours = 50;
  double normal_radius = resolution * 6.0;
  Keypoint keypoint_type = Keypoint::SIFT;
  double keypoint_threshold = 5.0;
  Descriptor descriptor_type = Descriptor::PFH;
  EstimationMethod estimation_method = EstimationMethod::MATCHING;
  bool refine_transform = true;
  double inlier_threshold = resolution * 5.0;
  double max_correspondence_distance = inlier_threshold * 2.0;
  int max_iterations = 500;
  size_t matching_k = 5;
  double transform_epsilon = 1e-2;
  double confidence_threshold = 0.0;
  double output_resolution = 0.05;

  /// `--param_name <value>`; unknown flags are ignored, `--name=value` is not supported (doc/wiki.txt:192)
  static MapMergingParams fromCommandLine(int argc, char** argv);
#ifdef MM3D_HAVE_ROS
  static MapMergingParams fromROSNode(const ros::NodeHandle& node);
#endif
};
std::ostream& operator<<(std::ostream& stream, const MapMergingParams& params);

/// transforms cloud -> reference frame; zero matrix where it could not be estimated
std::vector<Eigen::Matrix4f> estimateMapsTransforms(const std::vector<PointCloudConstPtr>& clouds, const MapMergingParams& params);

/// nullptr for empty input; clouds with a zero transform are skipped
PointCloudPtr composeMaps(const std::vector<PointCloudConstPtr>& clouds, const std::vector<Eigen::Matrix4f>& transforms, double resolution);
}  // namespace map_merge_3d

#endif  // MM3D_SHIM_MAP_MERGING_H_
