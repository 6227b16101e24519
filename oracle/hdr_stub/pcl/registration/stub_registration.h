// Stand-ins for the PCL registration classes matching.cpp configures and runs.  They only record the settings; the work is
// done by functions of the CPU checker (defined in oracle/mapmerging_ref_shim.cpp).  Test infrastructure only.
#pragma once
#include <map_merge_3d/typedefs.h>
#include <pcl/common/transforms.h>  // the real registration headers pull it in; matching.cpp relies on that
namespace pcl
{
namespace stub
{
struct RansacRejector {
  map_merge_3d::PointCloudPtr src, tgt;
  CorrespondencesPtr corr;
  double thr = 0.05;
  Eigen::Matrix4f best = Eigen::Matrix4f::Identity();
  void setInputSource(const map_merge_3d::PointCloudPtr& c) { src = c; }
  void setInputTarget(const map_merge_3d::PointCloudPtr& c) { tgt = c; }
  void setInputCorrespondences(const CorrespondencesPtr& c) { corr = c; }
  void setInlierThreshold(double t) { thr = t; }
  void getCorrespondences(Correspondences& remaining);  // runs the rejection
  Eigen::Matrix4f getBestTransformation() { return best; }
};
struct SvdEstimator {
  void estimateRigidTransformation(const map_merge_3d::PointCloud& src, const map_merge_3d::PointCloud& tgt, const Correspondences& corr,
                                   Eigen::Matrix4f& out) const;
};
struct Icp {
  double max_corr = 0, ransac_thr = 0, eps = 0;
  int max_it = 10;
  map_merge_3d::PointCloudPtr src, tgt;
  Eigen::Matrix4f final_t = Eigen::Matrix4f::Identity();
  void setMaxCorrespondenceDistance(double v) { max_corr = v; }
  void setRANSACOutlierRejectionThreshold(double v) { ransac_thr = v; }  // unused by pcl::IterativeClosestPoint without a rejector
  void setTransformationEpsilon(double v) { eps = v; }
  void setMaximumIterations(int v) { max_it = v; }
  void setInputSource(const map_merge_3d::PointCloudPtr& c) { src = c; }
  void setInputTarget(const map_merge_3d::PointCloudPtr& c) { tgt = c; }
  void align(map_merge_3d::PointCloud& output);
  Eigen::Matrix4f getFinalTransformation() { return final_t; }
};
struct Validator {
  double max_range = 0;
  void setMaxRange(double v) { max_range = v; }
  double validateTransformation(const map_merge_3d::PointCloudPtr& src, const map_merge_3d::PointCloudPtr& tgt, const Eigen::Matrix4f& t) const;
};
Eigen::Matrix4f sac_ia_run(const map_merge_3d::PointCloud& skp, const float* sdesc, const map_merge_3d::PointCloud& tkp, const float* tdesc, int dim,
                           double min_sample_distance, double max_corr, int max_it);
}  // namespace stub

namespace registration
{
template <typename P> using CorrespondenceRejectorSampleConsensus = stub::RansacRejector;
template <typename P, typename Q> using TransformationEstimationSVD = stub::SvdEstimator;
template <typename P, typename Q> using TransformationValidationEuclidean = stub::Validator;
}  // namespace registration
template <typename P, typename Q> using IterativeClosestPoint = stub::Icp;

template <typename P, typename Q, typename F>
class SampleConsensusInitialAlignment
{
public:
  void setMinSampleDistance(double v) { min_d_ = v; }
  void setMaxCorrespondenceDistance(double v) { max_c_ = v; }
  void setMaximumIterations(int v) { max_it_ = v; }
  void setInputSource(const map_merge_3d::PointCloudPtr& c) { src_ = c; }
  void setInputTarget(const map_merge_3d::PointCloudPtr& c) { tgt_ = c; }
  void setSourceFeatures(const typename PointCloud<F>::Ptr& f) { sf_ = f; }
  void setTargetFeatures(const typename PointCloud<F>::Ptr& f) { tf_ = f; }
  void align(map_merge_3d::PointCloud&)
  {
    const int D = desc_dim<F>::value;
    std::vector<float> s(sf_->points.size() * D), t(tf_->points.size() * D);
    for (size_t i = 0; i < sf_->points.size(); ++i) std::memcpy(&s[i * D], &sf_->points[i], sizeof(float) * D);
    for (size_t i = 0; i < tf_->points.size(); ++i) std::memcpy(&t[i * D], &tf_->points[i], sizeof(float) * D);
    final_ = stub::sac_ia_run(*src_, s.data(), *tgt_, t.data(), D, min_d_, max_c_, max_it_);
  }
  Eigen::Matrix4f getFinalTransformation() { return final_; }

private:
  double min_d_ = 0, max_c_ = 0;
  int max_it_ = 0;
  map_merge_3d::PointCloudPtr src_, tgt_;
  typename PointCloud<F>::Ptr sf_, tf_;
  Eigen::Matrix4f final_ = Eigen::Matrix4f::Identity();
};
}  // namespace pcl
