// segment.h -- forwarding header: the type lives in b200_value_types.h (kept so that code written against the
// reference's include paths compiles unchanged).
#pragma once
#include "mav_trajectory_generation/b200_value_types.h"
