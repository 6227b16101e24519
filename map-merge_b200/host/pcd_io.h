// pcd_io.h — minimal PCD reader/writer for pcl::PointXYZRGB clouds, standing in for
// pcl::io::loadPCDFile / savePCDFileBinary as used by map_merge_3d/src/map_merge_tool.cpp:24-33,52.
// Reads ascii, binary and binary_compressed (LZF) files with any field set that contains x y z and
// optionally rgb / rgba; writes the layout PCL writes for PointXYZRGB (FIELDS x y z rgb, DATA binary).
#ifndef MM3D_PCD_IO_H_
#define MM3D_PCD_IO_H_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <map_merge_3d/typedefs.h>

namespace mm3d_io
{
inline bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len)
{
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + out_len;
  while (ip < in_end) {
    unsigned ctrl = *ip++;
    if (ctrl < (1u << 5)) {  // literal run
      ++ctrl;
      if (ctrl > (size_t)(out_end - op) || ctrl > (size_t)(in_end - ip)) return false;
      std::memcpy(op, ip, ctrl);
      op += ctrl;
      ip += ctrl;
    } else {  // back reference
      unsigned len = ctrl >> 5;
      if (ip >= in_end) return false;
      if (len == 7) {
        len += *ip++;
        if (ip >= in_end) return false;
      }
      const size_t back = ((size_t)(ctrl & 0x1f) << 8) + 1 + *ip++;  // distance of the reference, checked before a pointer is formed
      if (back > (size_t)(op - out) || (size_t)len + 2 > (size_t)(out_end - op)) return false;
      const uint8_t* ref = op - back;
      len += 2;
      while (len--) *op++ = *ref++;
    }
  }
  return op == out_end;
}

struct Field {
  std::string name;
  int size = 4;
  char type = 'F';
  int count = 1;
  int offset = 0;
};

// returns 0 on success, < 0 on failure (like pcl::io::loadPCDFile)
inline int loadPCDFile(const std::string& path, map_merge_3d::PointCloud& cloud)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::vector<Field> fields;
  size_t width = 0, height = 1, points = 0;
  std::string data_mode, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS" || key == "COLUMNS") {
      std::string n;
      while (ss >> n) {
        Field fd;
        fd.name = n;
        fields.push_back(fd);
      }
    } else if (key == "SIZE") {
      for (Field& fd : fields) ss >> fd.size;
    } else if (key == "TYPE") {
      for (Field& fd : fields) ss >> fd.type;
    } else if (key == "COUNT") {
      for (Field& fd : fields) ss >> fd.count;
    } else if (key == "WIDTH") {
      ss >> width;
    } else if (key == "HEIGHT") {
      ss >> height;
    } else if (key == "POINTS") {
      ss >> points;
    } else if (key == "DATA") {
      ss >> data_mode;
      break;
    }
  }
  if (fields.empty() || data_mode.empty()) return -1;
  if (points == 0) points = width * height;
  int step = 0;
  for (Field& fd : fields) {
    fd.offset = step;
    step += fd.size * fd.count;
  }
  int ix = -1, iy = -1, iz = -1, ic = -1;
  for (size_t i = 0; i < fields.size(); ++i) {
    if (fields[i].name == "x") ix = (int)i;
    if (fields[i].name == "y") iy = (int)i;
    if (fields[i].name == "z") iz = (int)i;
    if (fields[i].name == "rgb" || fields[i].name == "rgba") ic = (int)i;
  }
  if (ix < 0 || iy < 0 || iz < 0) return -1;
  // the payload cannot be larger than what is left of the file: refuse headers that claim more (corrupt or truncated files)
  const std::streampos data_pos = f.tellg();
  f.seekg(0, std::ios::end);
  const std::streampos end_pos = f.tellg();
  f.seekg(data_pos);
  if (data_pos < 0 || end_pos < data_pos || step <= 0) return -1;
  const size_t remaining = (size_t)(end_pos - data_pos);
  if (data_mode == "ascii" && points > remaining) return -1;  // at least one byte per point
  if (data_mode == "binary" && points > remaining / (size_t)step) return -1;
  if (data_mode == "binary_compressed" && points > (size_t)0xffffffffu / (size_t)step) return -1;
  cloud.points.assign(points, map_merge_3d::PointT());
  auto read_float = [](const uint8_t* p, const Field& fd) -> float {
    if (fd.type == 'F' && fd.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
    if (fd.type == 'F' && fd.size == 8) { double v; std::memcpy(&v, p, 8); return (float)v; }
    if (fd.size == 4) { int32_t v; std::memcpy(&v, p, 4); return (float)v; }
    if (fd.size == 2) { int16_t v; std::memcpy(&v, p, 2); return (float)v; }
    return (float)*p;
  };
  if (data_mode == "ascii") {
    for (size_t i = 0; i < points; ++i) {
      if (!std::getline(f, line)) return -1;
      std::istringstream ss(line);
      map_merge_3d::PointT& p = cloud.points[i];
      for (size_t k = 0; k < fields.size(); ++k)
        for (int c = 0; c < fields[k].count; ++c) {
          std::string tok;
          if (!(ss >> tok)) return -1;
          if (c) continue;
          if ((int)k == ix) p.x = std::stof(tok);
          else if ((int)k == iy) p.y = std::stof(tok);
          else if ((int)k == iz) p.z = std::stof(tok);
          else if ((int)k == ic) {
            if (fields[k].type == 'F') { float v = std::stof(tok); std::memcpy(&p.rgba, &v, 4); }
            else p.rgba = (uint32_t)std::stoul(tok);
          }
        }
    }
  } else {
    std::vector<uint8_t> raw((size_t)step * points);
    if (data_mode == "binary") {
      f.read((char*)raw.data(), (std::streamsize)raw.size());
      if ((size_t)f.gcount() != raw.size()) return -1;
    } else if (data_mode == "binary_compressed") {
      uint32_t comp = 0, uncomp = 0;
      f.read((char*)&comp, 4);
      if (f.gcount() != 4) return -1;
      f.read((char*)&uncomp, 4);
      if (f.gcount() != 4) return -1;
      if (uncomp != raw.size() || remaining < 8 || comp > remaining - 8) return -1;
      std::vector<uint8_t> cbuf(comp), soa(uncomp);
      f.read((char*)cbuf.data(), comp);
      if ((size_t)f.gcount() != (size_t)comp) return -1;
      if (!lzf_decompress(cbuf.data(), comp, soa.data(), uncomp)) return -1;
      // compressed files store field after field (SoA); rebuild the AoS rows
      size_t off = 0;
      for (const Field& fd : fields) {
        const size_t fs = (size_t)fd.size * fd.count;
        for (size_t i = 0; i < points; ++i) std::memcpy(&raw[i * step + fd.offset], &soa[off + i * fs], fs);
        off += fs * points;
      }
    } else {
      return -1;
    }
    for (size_t i = 0; i < points; ++i) {
      const uint8_t* row = &raw[i * step];
      map_merge_3d::PointT& p = cloud.points[i];
      p.x = read_float(row + fields[ix].offset, fields[ix]);
      p.y = read_float(row + fields[iy].offset, fields[iy]);
      p.z = read_float(row + fields[iz].offset, fields[iz]);
      if (ic >= 0) std::memcpy(&p.rgba, row + fields[ic].offset, 4);
    }
  }
  cloud.width = (uint32_t)points;
  cloud.height = 1;
  cloud.is_dense = true;
  for (const auto& p : cloud.points)
    if (p.x != p.x || p.y != p.y || p.z != p.z) cloud.is_dense = false;
  return 0;
}

inline int savePCDFileBinary(const std::string& path, const map_merge_3d::PointCloud& cloud)
{
  std::ofstream f(path, std::ios::binary);
  if (!f) return -1;
  const size_t n = cloud.points.size();
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  for (const auto& p : cloud.points) {
    float row[4] = {p.x, p.y, p.z, 0.f};
    std::memcpy(&row[3], &p.rgba, 4);
    f.write((const char*)row, 16);
  }
  return f ? 0 : -1;
}
}  // namespace mm3d_io

#endif  // MM3D_PCD_IO_H_
