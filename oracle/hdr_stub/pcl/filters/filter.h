// Stand-in: see <pcl/stub_features.h>.
#pragma once
#include <pcl/stub_features.h>
