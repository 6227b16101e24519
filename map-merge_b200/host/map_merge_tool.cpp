// map_merge_tool — the reference's CLI front-end (map_merge_3d/src/map_merge_tool.cpp:8-55) on the shim:
//   map_merge_tool a.pcd b.pcd [...] [--param value ...]
// loads the PCDs, estimates the transforms, prints them and writes the composed map to output.pcd.
#include <iostream>

#include <map_merge_3d/map_merging.h>

#include "pcd_io.h"

using namespace map_merge_3d;

int main(int argc, char** argv)
{
  // pcl::console::parse_file_extension_argument(argc, argv, ".pcd")
  std::vector<int> pcd_file_indices;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.size() > 4 && a.compare(a.size() - 4, 4, ".pcd") == 0) pcd_file_indices.push_back(i);
  }
  const std::string output_name = "output.pcd";
  if (pcd_file_indices.size() < 2) {
    std::cerr << "Need at least 2 input files!\n";
    return -1;
  }
  MapMergingParams params;
  try {
    params = MapMergingParams::fromCommandLine(argc, argv);
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return -1;
  }
  std::cout << "params: " << std::endl << params << std::endl;

  std::vector<PointCloudConstPtr> clouds;
  for (int idx : pcd_file_indices) {
    PointCloudPtr cloud(new PointCloud);
    int rc = -1;
    try {
      rc = mm3d_io::loadPCDFile(argv[idx], *cloud);
    } catch (const std::exception& e) {  // corrupt file: same exit as an unreadable one (map_merge_tool.cpp:27-31)
      std::cerr << e.what() << "\n";
    }
    if (rc < 0) {
      std::cerr << "Error loading pointcloud file " << argv[idx] << ". Aborting.\n";
      return -1;
    }
    clouds.push_back(cloud);
  }
  std::cout << "> Estimating transforms.\n";
  std::vector<Eigen::Matrix4f> transforms;
  try {
    transforms = estimateMapsTransforms(clouds, params);
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return -2;
  }
  std::cout << "> Estimated transforms:\n";
  for (const auto& transform : transforms) std::cout << transform << std::endl;
  std::cout << "> Compositing clouds and writing to output.pcd\n";
  // clouds and transforms are passed as they are (map_merge_tool.cpp:49-50): when trailing maps have no keypoints the
  // transform list is shorter and composeMaps throws, as in the reference (which lets the exception terminate the tool)
  PointCloudPtr result;
  try {
    result = composeMaps(clouds, transforms, params.output_resolution);
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return -3;
  }
  if (!result) return -3;
  return mm3d_io::savePCDFileBinary(output_name, *result);
}
