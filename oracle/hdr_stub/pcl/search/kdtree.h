// Stand-in: pcl::search::KdTree lives in <pcl/stub_types.h>.
#pragma once
#include <pcl/stub_types.h>
