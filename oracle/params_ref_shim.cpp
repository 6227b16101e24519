// Prints, as JSON, what the REFERENCE's own headers say: the default-constructed MapMergingParams
// (include/map_merge_3d/map_merging.h:28-44) and the three string-convertible enums (enum.h ENUM_CLASS, features.h:20-49,
// matching.h:103).  Compiled by `make -C oracle ref` against /root/reference/map_merge_3d/include with the two stand-in
// headers of hdr_stub/ (typedefs.h, ros/ros.h); the output is committed as tests/golden/params_ref.json.
// Test infrastructure only.
#include <cstdio>
#include <sstream>

#include <map_merge_3d/map_merging.h>

using namespace map_merge_3d;

template <typename E>
static void dump_enum(std::ostringstream& o, const char* name, int n)
{
  o << "\"" << name << "\": {\"names\": [";
  for (int i = 0; i < n; ++i) {
    std::ostringstream s;
    s << static_cast<E>(i);  // operator<< of ENUM_CLASS
    o << (i ? ", " : "") << "\"" << s.str() << "\"";
  }
  o << "], \"round_trip\": [";
  for (int i = 0; i < n; ++i) o << (i ? ", " : "") << (int)enums::from_string<E>(enums::to_string(static_cast<E>(i)));
  o << "], \"invalid\": \"";
  try {
    enums::from_string<E>("no_such_value");
    o << "(no exception)";
  } catch (const std::runtime_error& e) {
    o << e.what();
  }
  o << "\"}";
}

int main()
{
  const MapMergingParams p;
  std::ostringstream o;
  o.precision(17);
  o << "{\"defaults\": {"
    << "\"resolution\": " << p.resolution << ", \"descriptor_radius\": " << p.descriptor_radius
    << ", \"outliers_min_neighbours\": " << p.outliers_min_neighbours << ", \"normal_radius\": " << p.normal_radius
    << ", \"keypoint_type\": " << (int)p.keypoint_type << ", \"keypoint_threshold\": " << p.keypoint_threshold
    << ", \"descriptor_type\": " << (int)p.descriptor_type << ", \"estimation_method\": " << (int)p.estimation_method
    << ", \"refine_transform\": " << (p.refine_transform ? 1 : 0) << ", \"inlier_threshold\": " << p.inlier_threshold
    << ", \"max_correspondence_distance\": " << p.max_correspondence_distance << ", \"max_iterations\": " << p.max_iterations
    << ", \"matching_k\": " << p.matching_k << ", \"transform_epsilon\": " << p.transform_epsilon
    << ", \"confidence_threshold\": " << p.confidence_threshold << ", \"output_resolution\": " << p.output_resolution << "}, ";
  dump_enum<Descriptor>(o, "Descriptor", 6);
  o << ", ";
  dump_enum<Keypoint>(o, "Keypoint", 2);
  o << ", ";
  dump_enum<EstimationMethod>(o, "EstimationMethod", 2);
  o << "}";
  std::puts(o.str().c_str());
  return 0;
}
