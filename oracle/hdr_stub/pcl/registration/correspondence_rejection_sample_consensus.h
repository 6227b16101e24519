// Stand-in: see <pcl/registration/stub_registration.h>.
#pragma once
#include <pcl/registration/stub_registration.h>
