// motion_defines.h -- names of the position derivatives (mirror of the reference's
// include/mav_trajectory_generation/motion_defines.h:25-47; same constants, same namespace).
#ifndef MAV_TRAJECTORY_GENERATION_MOTION_DEFINES_H_
#define MAV_TRAJECTORY_GENERATION_MOTION_DEFINES_H_

#include <string>

namespace mav_trajectory_generation {
namespace derivative_order {
static constexpr int INVALID = -1;
static constexpr int POSITION = 0, VELOCITY = 1, ACCELERATION = 2, JERK = 3, SNAP = 4;
static constexpr int ORIENTATION = 0, ANGULAR_VELOCITY = 1, ANGULAR_ACCELERATION = 2;
}  // namespace derivative_order

std::string positionDerivativeToString(int derivative);
int positionDerivativeToInt(const std::string& string);
std::string orintationDerivativeToString(int derivative);  // (sic) reference spelling
int orientationDerivativeToInt(const std::string& string);
}  // namespace mav_trajectory_generation
#endif
