// The reference's five gtest cases (map_merge_3d/test/test_map_merging.cpp:9-40) re-expressed against the
// shim, plus API-surface checks (enum spellings, parameter defaults and parsing).  No gtest, no GPU needed:
// the degenerate cases never reach the device.
#include <cstdio>
#include <iostream>
#include <sstream>

#include <map_merge_3d/map_merging.h>

using Eigen::Matrix4f;
using namespace map_merge_3d;

static int failures = 0, passed = 0;
#define EXPECT_TRUE(x)                                              \
  do {                                                              \
    if (!(x)) {                                                     \
      std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #x);  \
      ++failures;                                                   \
    }                                                               \
  } while (0)
#define TEST(name) static void name()
#define RUN(name)                     \
  do {                                \
    const int before = failures;      \
    name();                           \
    if (failures == before) ++passed; \
    std::printf("%s %s\n", failures == before ? "[ OK ]" : "[FAIL]", #name); \
  } while (0)

TEST(estimateMapsTransforms_empty)
{
  std::vector<Matrix4f> result = estimateMapsTransforms({}, MapMergingParams());
  EXPECT_TRUE(result.empty());
}

TEST(estimateMapsTransforms_one)
{
  std::vector<Matrix4f> result = estimateMapsTransforms({PointCloudConstPtr(new PointCloud)}, MapMergingParams());
  EXPECT_TRUE(result.size() == 1);
  EXPECT_TRUE(result[0] == Matrix4f::Identity());
}

TEST(composeMaps_empty)
{
  PointCloudPtr result = composeMaps({}, {}, 0.0);
  EXPECT_TRUE(result == nullptr);
}

TEST(composeMaps_wrongSizes)
{
  bool thrown = false;
  try {
    composeMaps({nullptr}, {}, 0.0);
  } catch (...) {
    thrown = true;
  }
  EXPECT_TRUE(thrown);
}

TEST(composeMaps_one)
{
  PointCloudPtr result = composeMaps({PointCloudConstPtr(new PointCloud)}, {Matrix4f::Identity()}, 0.0);
  EXPECT_TRUE(result != nullptr);
  EXPECT_TRUE(result->size() == 0);
}

static void api_surface()
{
  // enum spellings are CLI / ROS parameter values (features.h:20-24,49; matching.h:103)
  EXPECT_TRUE(std::string(enums::to_string(Descriptor::FPFH)) == "FPFH");
  EXPECT_TRUE(enums::from_string<Descriptor>("SHOT") == Descriptor::SHOT);
  EXPECT_TRUE((int)Descriptor::PFH == 0 && (int)Descriptor::SC3D == 5);
  EXPECT_TRUE(enums::from_string<Keypoint>("HARRIS") == Keypoint::HARRIS);
  EXPECT_TRUE(enums::from_string<EstimationMethod>("SAC_IA") == EstimationMethod::SAC_IA);
  bool thrown = false;
  try {
    enums::from_string<Keypoint>("sift");  // case-sensitive
  } catch (const std::runtime_error&) {
    thrown = true;
  }
  EXPECT_TRUE(thrown);
  // dependent defaults are frozen at resolution 0.1 (map_merging.h:29-39)
  const char* argv[] = {"tool", "a.pcd", "--resolution", "0.2", "--descriptor_type", "FPFH", "--matching_k", "-3", "--bogus", "1",
                        "--refine_transform", "0", "b.pcd"};
  MapMergingParams p = MapMergingParams::fromCommandLine(13, const_cast<char**>(argv));
  EXPECT_TRUE(p.resolution == 0.2 && p.descriptor_radius == 0.1 * 8.0 && p.normal_radius == 0.1 * 6.0);
  EXPECT_TRUE(p.inlier_threshold == 0.1 * 5.0 && p.max_correspondence_distance == 0.1 * 5.0 * 2.0);
  EXPECT_TRUE(p.descriptor_type == Descriptor::FPFH && p.matching_k == 5 && !p.refine_transform);
  std::ostringstream os;
  os << MapMergingParams();
  EXPECT_TRUE(os.str().find("descriptor_type: PFH\n") != std::string::npos);
  EXPECT_TRUE(os.str().find("keypoint_type: SIFT\n") != std::string::npos);
  EXPECT_TRUE(os.str().find("output_resolution: 0.05\n") != std::string::npos);
}

int main()
{
  RUN(estimateMapsTransforms_empty);
  RUN(estimateMapsTransforms_one);
  RUN(composeMaps_empty);
  RUN(composeMaps_wrongSizes);
  RUN(composeMaps_one);
  std::printf("%d passed\n", passed);
  const int before = failures;
  api_surface();
  std::printf("api surface: %s\n", failures == before ? "ok" : "FAILED");
  return failures ? 1 : 0;
}
