// The REFERENCE's own driver — src/map_merging.cpp (estimateMapsTransforms, computeGlobalTransforms, composeMaps,
// MapMergingParams::fromCommandLine, operator<<) and src/graph.cpp — compiled unmodified where it lies under
// /root/reference, on top of the CPU checker's stage functions.  PCL / Eigen / ROS are replaced by the stand-in headers of
// hdr_stub/ and eigen_stub/; this file supplies what those headers only declare:
//   * map_merge_3d::downSample ... transformScore  (features.h / matching.h)  -> the checker's restatements
//   * Eigen::Matrix4f arithmetic, pcl::transformPointCloud                      -> the checker's 4x4 routines
// So the control flow (stage order, pair generation, confidence = 1 / score, pose graph, chaining with inverses, the
// composeMaps skip / exception rules, the command-line flag table and the params printout) is the reference's own code,
// and only the PCL-backed arithmetic underneath is restated.  Output: oracle/_ref/libmapmerging_ref.so, and (with
// -DMAPMERGING_REF_MAIN) oracle/_ref/mapmerging_params, which prints the parsed command line.
// Test infrastructure only.
#include "mm3d_oracle.cpp"  // the checker, as one translation unit (its functions are internal)

#include <cstring>
#include <sstream>

#include <map_merge_3d/map_merging.h>  // the reference's header (-I/root/reference/map_merge_3d/include)

// ---- Eigen / PCL stand-ins ------------------------------------------------------------------------------------------------
namespace Eigen
{
static orc::Mat4 to_orc(const Matrix4f& a)
{
  orc::Mat4 r;
  memcpy(r.m, a.m, sizeof(r.m));
  return r;
}
static Matrix4f from_orc(const orc::Mat4& a)
{
  Matrix4f r;
  memcpy(r.m, a.m, sizeof(r.m));
  return r;
}
Matrix4f Matrix4f::Zero() { return from_orc(orc::Mat4::zero()); }
Matrix4f Matrix4f::Identity() { return from_orc(orc::Mat4::identity()); }
Matrix4f Matrix4f::inverse() const { return from_orc(orc::inverse(to_orc(*this))); }
Matrix4f Matrix4f::operator*(const Matrix4f& o) const { return from_orc(orc::mul(to_orc(*this), to_orc(o))); }
bool Matrix4f::isZero(float prec) const
{
  for (int i = 0; i < 16; ++i)
    if (!(std::fabs(m[i]) <= prec)) return false;
  return true;
}
}  // namespace Eigen

static orc::Cloud to_orc_cloud(const map_merge_3d::PointCloud& c)
{
  orc::Cloud o(c.points.size());
  static_assert(sizeof(map_merge_3d::PointT) == sizeof(orc::P4), "point layouts differ");
  if (!o.empty()) memcpy(o.data(), c.points.data(), o.size() * sizeof(orc::P4));
  return o;
}
static map_merge_3d::PointCloudPtr from_orc_cloud(const orc::Cloud& c)
{
  map_merge_3d::PointCloudPtr p(new map_merge_3d::PointCloud);
  p->points.resize(c.size());
  if (!c.empty()) memcpy(p->points.data(), c.data(), c.size() * sizeof(orc::P4));
  return p;
}

namespace pcl
{
void transformPointCloud(const map_merge_3d::PointCloud& in, map_merge_3d::PointCloud& out, const Eigen::Matrix4f& t)
{
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); ++i) {
    const map_merge_3d::PointT& p = in.points[i];
    map_merge_3d::PointT q = p;
    mm3d::em::transform_point(t.m, p.x, p.y, p.z, &q.x, &q.y, &q.z);
    out.points[i] = q;
  }
}
}  // namespace pcl

// ---- features.h / matching.h on the checker -----------------------------------------------------------------------------------
namespace map_merge_3d
{
PointCloudPtr downSample(const PointCloudConstPtr& input, double resolution)
{
  return from_orc_cloud(orc::voxel_grid(to_orc_cloud(*input), (float)resolution));
}
PointCloudPtr removeOutliers(const PointCloudConstPtr& input, double radius, int min_neighbours)
{
  return from_orc_cloud(orc::radius_outlier_removal(to_orc_cloud(*input), radius, min_neighbours, nullptr, nullptr));
}
static orc::Normals to_orc_normals(const SurfaceNormals& n)
{
  orc::Normals o(n.v.size() / 4);
  if (!o.empty()) memcpy(o.data(), n.v.data(), n.v.size() * 4);
  return o;
}
SurfaceNormalsPtr computeSurfaceNormals(const PointCloudConstPtr& input, double radius)
{
  const orc::Normals n = orc::surface_normals(to_orc_cloud(*input), radius);
  SurfaceNormalsPtr p(new SurfaceNormals);
  p->v.resize(n.size() * 4);
  if (!n.empty()) memcpy(p->v.data(), n.data(), n.size() * sizeof(orc::N4));
  return p;
}
PointCloudPtr detectKeypoints(const PointCloudConstPtr& points, const SurfaceNormalsPtr& normals, Keypoint type, double threshold, double radius,
                              double resolution)
{
  const orc::Cloud c = to_orc_cloud(*points);
  if (type == Keypoint::HARRIS) return from_orc_cloud(orc::harris_keypoints(c, to_orc_normals(*normals), (float)threshold, (float)radius));
  return from_orc_cloud(orc::sift_keypoints(c, (float)resolution, 3, 3, (float)threshold, 0));
}
LocalDescriptorsPtr computeLocalDescriptors(const PointCloudConstPtr& points, const SurfaceNormalsPtr& normals, const PointCloudPtr& keypoints,
                                            Descriptor descriptor, double feature_radius)
{
  const orc::Cloud c = to_orc_cloud(*points);
  const orc::Normals n = to_orc_normals(*normals);
  orc::Cloud kp = to_orc_cloud(*keypoints);
  LocalDescriptorsPtr d(new LocalDescriptors);
  switch (descriptor) {
    case Descriptor::PFH: d->v = orc::pfh_descriptors(c, n, kp, feature_radius); d->dim = 125; break;
    case Descriptor::PFHRGB: d->v = orc::pfhrgb_descriptors(c, n, kp, feature_radius); d->dim = 250; break;
    case Descriptor::FPFH: d->v = orc::fpfh_descriptors(c, n, kp, feature_radius); d->dim = 33; break;
    case Descriptor::RSD: d->v = orc::rsd_descriptors(c, n, kp, feature_radius); d->dim = 2; break;
    case Descriptor::SHOT: d->v = orc::shot_descriptors(c, n, kp, feature_radius); d->dim = 1344; break;
    case Descriptor::SC3D: d->v = orc::sc3d_descriptors(c, n, kp, feature_radius); d->dim = 1980; break;
  }
  keypoints->points = from_orc_cloud(kp)->points;  // the reference filters the keypoints in place (features.cpp:137-141)
  return d;
}
Eigen::Matrix4f estimateTransform(const PointCloudPtr& source_points, const PointCloudPtr& source_keypoints,
                                  const LocalDescriptorsPtr& source_descriptors, const PointCloudPtr& target_points,
                                  const PointCloudPtr& target_keypoints, const LocalDescriptorsPtr& target_descriptors, EstimationMethod method,
                                  bool refine, double inlier_threshold, double max_correspondence_distance, int max_iterations, size_t matching_k,
                                  double transform_epsilon)
{
  // the glue of matching.cpp:223-257 (a PCL translation unit), restated
  const orc::Cloud skp = to_orc_cloud(*source_keypoints), tkp = to_orc_cloud(*target_keypoints);
  orc::Mat4 t;
  if (method == EstimationMethod::SAC_IA) {
    t = orc::sac_ia_transform(skp, source_descriptors->v, tkp, target_descriptors->v, source_descriptors->dim, inlier_threshold,
                              max_correspondence_distance, max_iterations, orc::g_pipeline_rand);
  } else {
    const std::vector<orc::Corr> corr = orc::find_correspondences(source_descriptors->v.data(), skp.size(), target_descriptors->v.data(), tkp.size(),
                                                                  source_descriptors->dim, matching_k);
    std::vector<int> inl;
    t = orc::ransac_transform(skp, tkp, corr, inlier_threshold, inl);
  }
  if (refine)
    t = orc::icp_refine(to_orc_cloud(*source_points), to_orc_cloud(*target_points), t, max_correspondence_distance, max_iterations, transform_epsilon);
  return Eigen::from_orc(t);
}
double transformScore(const PointCloudPtr& source_points, const PointCloudPtr& target_points, const Eigen::Matrix4f& transform, double max_distance)
{
  return orc::transform_score(to_orc_cloud(*source_points), to_orc_cloud(*target_points), Eigen::to_orc(transform), max_distance);
}
}  // namespace map_merge_3d

// ---- C entry points -----------------------------------------------------------------------------------------------------------
static map_merge_3d::MapMergingParams to_ref_params(const Params& p)
{
  using namespace map_merge_3d;
  MapMergingParams r;
  r.resolution = p.resolution;
  r.descriptor_radius = p.descriptor_radius;
  r.outliers_min_neighbours = p.outliers_min_neighbours;
  r.normal_radius = p.normal_radius;
  r.keypoint_type = static_cast<Keypoint>(p.keypoint_type);
  r.keypoint_threshold = p.keypoint_threshold;
  r.descriptor_type = static_cast<Descriptor>(p.descriptor_type);
  r.estimation_method = static_cast<EstimationMethod>(p.estimation_method);
  r.refine_transform = p.refine_transform != 0;
  r.inlier_threshold = p.inlier_threshold;
  r.max_correspondence_distance = p.max_correspondence_distance;
  r.max_iterations = p.max_iterations;
  r.matching_k = (size_t)p.matching_k;
  r.transform_epsilon = p.transform_epsilon;
  r.confidence_threshold = p.confidence_threshold;
  r.output_resolution = p.output_resolution;
  return r;
}

extern "C" int ref_estimate_maps_transforms(int n_maps, const float* const* clouds, const uint64_t* n_points, const Params* p, float* out_transforms,
                                            int* n_out)
{
  std::vector<map_merge_3d::PointCloudConstPtr> in;
  for (int i = 0; i < n_maps; ++i) in.push_back(from_orc_cloud(to_cloud(clouds[i], n_points[i])));
  orc::g_pipeline_rand.seed(1);  // a fresh process: rand() has never been called
  const std::vector<Eigen::Matrix4f> t = map_merge_3d::estimateMapsTransforms(in, to_ref_params(*p));
  for (size_t i = 0; i < t.size(); ++i) to_colmajor(Eigen::to_orc(t[i]), out_transforms + 16 * i);
  *n_out = (int)t.size();
  return 0;
}

// returns 0 ok, 1 nullptr result (empty input), 2 the reference threw
extern "C" int ref_compose_maps(int n_maps, const float* const* clouds, const uint64_t* n_points, int n_transforms, const float* transforms,
                                double resolution, float** out, uint64_t* n_out)
{
  std::vector<map_merge_3d::PointCloudConstPtr> in;
  for (int i = 0; i < n_maps; ++i) in.push_back(from_orc_cloud(to_cloud(clouds[i], n_points[i])));
  std::vector<Eigen::Matrix4f> t(n_transforms);
  for (int i = 0; i < n_transforms; ++i)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) t[i].m[r * 4 + c] = transforms[16 * i + c * 4 + r];
  *out = nullptr;
  *n_out = 0;
  try {
    map_merge_3d::PointCloudPtr res = map_merge_3d::composeMaps(in, t, resolution);
    if (!res) return 1;
    *out = dup_f(res->points.data(), res->points.size() * sizeof(map_merge_3d::PointT));
    *n_out = res->points.size();
    return 0;
  } catch (std::runtime_error* e) {  // "throw new std::runtime_error" (map_merging.cpp:285)
    delete e;
    return 2;
  } catch (...) {
    return 2;
  }
}

#ifdef MAPMERGING_REF_MAIN
// oracle/_ref/mapmerging_params: MapMergingParams::fromCommandLine + operator<< on this process's command line.  A
// separate executable because iostreams inside a ctypes-loaded library clash with the libstdc++ numpy brings along.
#include <iostream>
int main(int argc, char** argv)
{
  try {
    std::cout << map_merge_3d::MapMergingParams::fromCommandLine(argc, argv);
  } catch (const std::exception& e) {
    std::cout << "EXCEPTION: " << e.what();
  }
  return 0;
}
#endif
