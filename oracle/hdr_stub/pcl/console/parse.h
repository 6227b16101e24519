// Stand-in for <pcl/console/parse.h>: pcl::console::parse_argument as PCL 1.8 behaves for the four value types
// map_merging.cpp uses [PCL-recall common/src/parse.cpp]: the FIRST occurrence of the flag counts, the next argv entry is
// its value, numbers go through atof / atoi, bool is "atoi == 1".  Returns the index of the flag or -1.
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>
namespace pcl
{
namespace console
{
inline int find_argument(int argc, char** argv, const char* name)
{
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], name) == 0) return i;
  return -1;
}
inline int parse_argument(int argc, char** argv, const char* name, std::string& val)
{
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) val = argv[i];
  return i - 1;
}
inline int parse_argument(int argc, char** argv, const char* name, double& val)
{
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) val = std::atof(argv[i]);
  return i - 1;
}
inline int parse_argument(int argc, char** argv, const char* name, int& val)
{
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) val = std::atoi(argv[i]);
  return i - 1;
}
inline int parse_argument(int argc, char** argv, const char* name, bool& val)
{
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) val = std::atoi(argv[i]) == 1;
  return i - 1;
}
}  // namespace console
}  // namespace pcl
