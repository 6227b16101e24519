// Stand-in: the estimator templates are declared in <pcl/stub_types.h>.
#pragma once
#include <pcl/stub_types.h>
