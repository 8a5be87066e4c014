#include "mav_trajectory_generation/motion_defines.h"

namespace mav_trajectory_generation {
namespace {
const char* const kPositionNames[] = {"position", "velocity", "acceleration", "jerk", "snap"};
const char* const kOrientationNames[] = {"orientation", "angular_velocity", "angular_acceleration"};
}  // namespace

std::string positionDerivativeToString(int derivative) {
  return (derivative >= 0 && derivative <= 4) ? kPositionNames[derivative] : "invalid";
}
int positionDerivativeToInt(const std::string& string) {
  for (int i = 0; i <= 4; ++i)
    if (string == kPositionNames[i]) return i;
  return derivative_order::INVALID;
}
std::string orintationDerivativeToString(int derivative) {
  return (derivative >= 0 && derivative <= 2) ? kOrientationNames[derivative] : "invalid";
}
int orientationDerivativeToInt(const std::string& string) {
  for (int i = 0; i <= 2; ++i)
    if (string == kOrientationNames[i]) return i;
  return derivative_order::INVALID;
}
}  // namespace mav_trajectory_generation
