// extremum.h -- reference include path kept: Extremum lives in b200_value_types.h.
#ifndef MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
#define MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
#include "mav_trajectory_generation/b200_value_types.h"
#endif  // MAV_TRAJECTORY_GENERATION_EXTREMUM_H_
