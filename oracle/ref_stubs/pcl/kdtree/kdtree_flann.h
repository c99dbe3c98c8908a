// ORACLE — TEST INFRASTRUCTURE ONLY: empty stand-in (nothing of this PCL header is used on the compiled path).
#pragma once
#include "pcl/point_cloud.h"
#include "pcl/point_types.h"
