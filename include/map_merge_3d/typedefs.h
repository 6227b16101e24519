// typedefs.h — the point-cloud typedefs of map_merge_3d/include/map_merge_3d/typedefs.h:15-33.
// With PCL installed the real types are used; otherwise layout-compatible minimal stand-ins
// (same member names, 32-byte PointXYZRGB / Normal, column-major Matrix4f) are provided so the
// shim, its tests and map_merge_tool build on a machine without PCL / Eigen / ROS.
#ifndef MM3D_SHIM_TYPEDEFS_H_
#define MM3D_SHIM_TYPEDEFS_H_

#if defined(__has_include)
#if __has_include(<pcl/point_cloud.h>) && __has_include(<Eigen/Core>) && !defined(MM3D_NO_PCL)
#define MM3D_HAVE_PCL 1
#endif
#endif

#ifdef MM3D_HAVE_PCL
#include <pcl/PCLPointCloud2.h>
#include <pcl/correspondence.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

namespace Eigen
{
struct Matrix4f {
  float m[16];  // column-major, like Eigen
  static Matrix4f Zero()
  {
    Matrix4f r;
    for (float& v : r.m) v = 0.0f;
    return r;
  }
  static Matrix4f Identity()
  {
    Matrix4f r = Zero();
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
    return r;
  }
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  float* data() { return m; }
  const float* data() const { return m; }
  void setZero() { *this = Zero(); }
  bool isZero(float prec = 1e-5f) const
  {
    for (float v : m)
      if (!(std::fabs(v) <= prec)) return false;
    return true;
  }
  bool operator==(const Matrix4f& o) const
  {
    for (int i = 0; i < 16; ++i)
      if (m[i] != o.m[i]) return false;
    return true;
  }
};
inline std::ostream& operator<<(std::ostream& s, const Matrix4f& t)
{
  for (int r = 0; r < 4; ++r) {
    for (int c = 0; c < 4; ++c) s << (c ? " " : "") << t(r, c);
    if (r < 3) s << "\n";
  }
  return s;
}
}  // namespace Eigen

namespace pcl
{
struct alignas(16) PointXYZRGB {
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;
  union {
    struct {
      uint8_t b, g, r, a;
    };
    float rgb;
    uint32_t rgba;
  };
  float pad_[3];
  PointXYZRGB() : rgba(0xff000000u), pad_{0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZRGB) == 32, "pcl::PointXYZRGB is 32 bytes");

struct alignas(16) Normal {
  union {
    float normal[3];
    struct {
      float normal_x, normal_y, normal_z;
    };
  };
  float data_n3 = 0.f;
  float curvature = 0.f;
  float pad_[3] = {0.f, 0.f, 0.f};
  Normal() : normal_x(0.f), normal_y(0.f), normal_z(0.f) {}
};
static_assert(sizeof(Normal) == 32, "pcl::Normal is 32 bytes");

template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  void push_back(const PointT& p)
  {
    points.push_back(p);
    width = (uint32_t)points.size();
  }
  void resize(size_t n)
  {
    points.resize(n);
    width = (uint32_t)n;
  }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

struct PCLPointField {
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 7;  // FLOAT32
  uint32_t count = 1;
};
struct PCLPointCloud2 {
  typedef std::shared_ptr<PCLPointCloud2> Ptr;
  typedef std::shared_ptr<const PCLPointCloud2> ConstPtr;
  uint32_t height = 1, width = 0;
  std::vector<PCLPointField> fields;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};

struct Correspondence {
  int index_query = 0;
  int index_match = -1;
  float distance = 0.f;
  Correspondence() {}
  Correspondence(int q, int m, float d) : index_query(q), index_match(m), distance(d) {}
};
typedef std::vector<Correspondence> Correspondences;
typedef std::shared_ptr<Correspondences> CorrespondencesPtr;
}  // namespace pcl
#endif  // MM3D_HAVE_PCL

namespace map_merge_3d
{
typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> PointCloud;
typedef pcl::PointCloud<PointT>::Ptr PointCloudPtr;
typedef pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;

typedef pcl::Normal NormalT;
typedef pcl::PointCloud<NormalT> SurfaceNormals;
typedef pcl::PointCloud<NormalT>::Ptr SurfaceNormalsPtr;
typedef pcl::PointCloud<NormalT>::ConstPtr SurfaceNormalsConstPtr;

typedef pcl::PCLPointCloud2 LocalDescriptors;
typedef pcl::PCLPointCloud2::Ptr LocalDescriptorsPtr;
typedef pcl::PCLPointCloud2::ConstPtr LocalDescriptorsConstPtr;

using pcl::Correspondences;
using pcl::CorrespondencesPtr;
}  // namespace map_merge_3d

#endif  // MM3D_SHIM_TYPEDEFS_H_
