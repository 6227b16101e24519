// The REFERENCE's own code — src/features.cpp (downSample, removeOutliers, detectKeypoints, computeLocalDescriptors with its
// invalid-descriptor filtering, computeSurfaceNormals), src/map_merging.cpp (estimateMapsTransforms, computeGlobalTransforms,
// composeMaps, MapMergingParams::fromCommandLine, operator<<), src/matching.cpp (findFeatureCorrespondences with its reciprocal
// cross-match, estimateTransformFromCorrespondences, estimateTransformFromDescriptorsSets, estimateTransformICP,
// estimateTransform, transformScore) and src/graph.cpp — compiled unmodified where it lies under /root/reference, on top of
// the CPU checker's stage functions.  PCL / Eigen / ROS are replaced by the stand-in headers of hdr_stub/ and
// eigen_stub/; this file supplies what those headers only declare:
//   * the PCL filter / keypoint / feature classes features.cpp configures (VoxelGrid, RadiusOutlierRemoval,
//     NormalEstimation, SIFTKeypoint, HarrisKeypoint3D, the six descriptor estimators)                     -> the checker
//   * the PCL registration classes matching.cpp configures (RANSAC rejector, SVD, SAC-IA, ICP, validation) -> the checker
//   * Eigen::Matrix4f arithmetic, pcl::transformPointCloud                      -> the checker's 4x4 routines
// So the control flow (stage order, how features.cpp configures every PCL object — leaf sizes, SIFT scales, Harris flags,
// search surface vs. input cloud — the NaN filtering of descriptors and keypoints, pair generation, the k-NN cross-match,
// which parameter configures which PCL registration object, the
// "identity means RANSAC failed" rule, final * initial_guess, confidence = 1 / score, pose graph, chaining with inverses,
// the composeMaps skip / exception rules, the command-line flag table and the params printout) is the reference's own
// code, and only the PCL-internal arithmetic underneath is restated.  Output: oracle/_ref/libmapmerging_ref.so, and (with
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
bool Matrix4f::isIdentity(float) const { return orc::is_identity(to_orc(*this)); }
void Matrix4f::setZero() { *this = Zero(); }
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

// ---- the PCL registration classes matching.cpp drives -----------------------------------------------------------------------
#include <pcl/registration/stub_registration.h>
namespace pcl
{
namespace stub
{
static std::vector<orc::Corr> to_orc_corr(const Correspondences& c)
{
  std::vector<orc::Corr> o(c.size());
  for (size_t i = 0; i < c.size(); ++i) o[i] = orc::Corr{c[i].index_query, c[i].index_match, c[i].distance};
  return o;
}
void RansacRejector::getCorrespondences(Correspondences& remaining)
{
  const std::vector<orc::Corr> oc = to_orc_corr(*corr);
  std::vector<int> inl;
  orc::Mat4 b;
  if (orc::ransac_reject(to_orc_cloud(*src), to_orc_cloud(*tgt), oc, thr, inl, b)) {
    remaining.clear();
    for (int i : inl) remaining.push_back((*corr)[i]);
    best = Eigen::from_orc(b);
  } else {  // PCL keeps the original correspondences and an identity transformation
    remaining = *corr;
    best = Eigen::Matrix4f::Identity();
  }
}
void SvdEstimator::estimateRigidTransformation(const map_merge_3d::PointCloud& src, const map_merge_3d::PointCloud& tgt, const Correspondences& corr,
                                               Eigen::Matrix4f& out) const
{
  const std::vector<orc::Corr> oc = to_orc_corr(corr);
  std::vector<int> all(oc.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  out = Eigen::from_orc(orc::svd_transform(to_orc_cloud(src), to_orc_cloud(tgt), oc, all));
}
void Icp::align(map_merge_3d::PointCloud&)
{
  // the source arrives already transformed by the initial guess (matching.cpp:208-210)
  final_t = Eigen::from_orc(orc::icp_refine(to_orc_cloud(*src), to_orc_cloud(*tgt), orc::Mat4::identity(), max_corr, max_it, eps));
}
double Validator::validateTransformation(const map_merge_3d::PointCloudPtr& src, const map_merge_3d::PointCloudPtr& tgt, const Eigen::Matrix4f& t) const
{
  return orc::transform_score(to_orc_cloud(*src), to_orc_cloud(*tgt), Eigen::to_orc(t), max_range);
}
Eigen::Matrix4f sac_ia_run(const map_merge_3d::PointCloud& skp, const float* sdesc, const map_merge_3d::PointCloud& tkp, const float* tdesc, int dim,
                           double min_sample_distance, double max_corr, int max_it)
{
  const std::vector<float> s(sdesc, sdesc + skp.points.size() * (size_t)dim), t(tdesc, tdesc + tkp.points.size() * (size_t)dim);
  return Eigen::from_orc(orc::sac_ia_transform(to_orc_cloud(skp), s, to_orc_cloud(tkp), t, dim, min_sample_distance, max_corr, max_it,
                                               orc::g_pipeline_rand));
}
}  // namespace stub
}  // namespace pcl

// ---- the PCL filter / keypoint / feature classes features.cpp drives ---------------------------------------------------------
#include <pcl/stub_features.h>
namespace pcl
{
namespace stub
{
static orc::Normals to_orc_normals(const Normals& n)
{
  orc::Normals o(n.points.size());
  static_assert(sizeof(pcl::Normal) == sizeof(orc::N4), "normal layouts differ");
  if (!o.empty()) memcpy(o.data(), n.points.data(), o.size() * sizeof(orc::N4));
  return o;
}
void voxel_grid(const Cloud& in, float lx, float ly, float lz, Cloud& out)
{
  if (lx != ly || lx != lz) throw std::runtime_error("stub VoxelGrid: cubic leaves only");
  out.points = from_orc_cloud(orc::voxel_grid(to_orc_cloud(in), lx))->points;
}
void radius_outlier_removal(const Cloud& in, double radius, int min_neighbours, Cloud& out)
{
  out.points = from_orc_cloud(orc::radius_outlier_removal(to_orc_cloud(in), radius, min_neighbours, nullptr, nullptr))->points;
}
void normal_estimation(const Cloud& in, double radius, Normals& out)
{
  const orc::Normals n = orc::surface_normals(to_orc_cloud(in), radius);
  out.points.resize(n.size());
  if (!n.empty()) memcpy(out.points.data(), n.data(), n.size() * sizeof(orc::N4));
}
void sift_keypoints(const Cloud& in, float min_scale, int nr_octaves, int nr_scales_per_octave, float min_contrast, PointCloud<PointWithScale>& out)
{
  std::vector<float> scales;
  const orc::Cloud k = orc::sift_keypoints(to_orc_cloud(in), min_scale, nr_octaves, nr_scales_per_octave, min_contrast, 0, nullptr, &scales);
  out.points.resize(k.size());
  for (size_t i = 0; i < k.size(); ++i) out.points[i] = PointWithScale{k[i].x, k[i].y, k[i].z, i < scales.size() ? scales[i] : 0.f};
}
void harris_keypoints(const Cloud& in, const Normals& normals, bool non_max, bool refine, float threshold, float radius, PointCloud<PointXYZI>& out)
{
  if (!non_max || !refine) throw std::runtime_error("stub HarrisKeypoint3D: only the configuration features.cpp uses (NMS + refine)");
  const orc::Cloud k = orc::harris_keypoints(to_orc_cloud(in), to_orc_normals(normals), threshold, radius);
  out.points.resize(k.size());
  for (size_t i = 0; i < k.size(); ++i) out.points[i] = PointXYZI{k[i].x, k[i].y, k[i].z, 0.f};
}
void descriptors(int kind, const Cloud& surface, const Normals& normals, const Cloud& keypoints, double radius, int dim, std::vector<float>& out)
{
  const orc::Cloud c = to_orc_cloud(surface);
  const orc::Normals n = to_orc_normals(normals);
  orc::Cloud kp = to_orc_cloud(keypoints);  // the checker drops the keypoints PCL marks invalid (NaN rows) ...
  std::vector<float> v;
  switch (kind) {
    case 0: v = orc::pfh_descriptors(c, n, kp, radius); break;
    case 1: v = orc::pfhrgb_descriptors(c, n, kp, radius); break;
    case 2: v = orc::fpfh_descriptors(c, n, kp, radius); break;
    case 3: v = orc::rsd_descriptors(c, n, kp, radius); break;
    case 4: v = orc::shot_descriptors(c, n, kp, radius); break;
    default: v = orc::sc3d_descriptors(c, n, kp, radius); break;
  }
  // ... so they are put back as NaN rows here, and the REFERENCE's own filtering (features.cpp:119-141) removes them
  const size_t nk = keypoints.points.size();
  out.assign(nk * (size_t)dim, std::numeric_limits<float>::quiet_NaN());
  size_t j = 0;
  for (size_t i = 0; i < nk && j < kp.size(); ++i) {
    const PointXYZRGB& p = keypoints.points[i];
    if (memcmp(&p, &kp[j], sizeof(orc::P4)) == 0) {  // kept keypoints are a subsequence of the input
      memcpy(&out[i * (size_t)dim], &v[j * (size_t)dim], sizeof(float) * (size_t)dim);
      ++j;
    }
  }
  if (j != kp.size()) throw std::runtime_error("stub descriptors: kept keypoints are not a subsequence of the input");
}
}  // namespace stub
}  // namespace pcl

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

// findFeatureCorrespondences (matching.cpp:31-108, the reference's reciprocal k-NN cross-match) on plain descriptor arrays;
// dim selects the descriptor type the way the PointCloud2 field name does.  pairs / dist are malloc'ed (n_corr x 2, n_corr).
extern "C" int ref_find_correspondences(const float* ds, uint64_t ns, const float* dt, uint64_t nt, int dim, uint64_t k, int32_t** pairs,
                                        float** dist, uint64_t* n_corr)
{
  const char* field = dim == 125 ? "pfh" : dim == 250 ? "pfhrgb" : dim == 33 ? "fpfh" : dim == 2 ? "r_min" : dim == 1344 ? "shot" : "shape_context";
  map_merge_3d::LocalDescriptorsPtr s(new map_merge_3d::LocalDescriptors), t(new map_merge_3d::LocalDescriptors);
  s->fields.push_back(map_merge_3d::PCLPointField{field});
  t->fields.push_back(map_merge_3d::PCLPointField{field});
  s->dim = t->dim = dim;
  s->v.assign(ds, ds + ns * (size_t)dim);
  t->v.assign(dt, dt + nt * (size_t)dim);
  const map_merge_3d::CorrespondencesPtr c = map_merge_3d::findFeatureCorrespondences(s, t, (size_t)k);
  *n_corr = c->size();
  *pairs = (int32_t*)malloc(std::max<size_t>(c->size(), 1) * 8);
  *dist = (float*)malloc(std::max<size_t>(c->size(), 1) * 4);
  for (size_t i = 0; i < c->size(); ++i) {
    (*pairs)[2 * i] = (*c)[i].index_query;
    (*pairs)[2 * i + 1] = (*c)[i].index_match;
    (*dist)[i] = (*c)[i].distance;
  }
  return 0;
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
