// io.h -- YAML (de)serialisation of solved segments in the reference's on-disk schema
// (reference src/io.cpp:27-31, 126-219):
//
//   segments:
//     - N: 10
//       D: 3
//       time: 3970847830  # [ns]            <- Segment::getTimeNSec(), uint64 nanoseconds
//       coefficients:
//         - [c0, c1, ..., c9]               <- one flow sequence per dimension, increasing powers
//         - [...]
//         - [...]
//
// Written without yaml-cpp (not available here): the writer emits exactly this layout with 17
// significant digits; the reader accepts this subset of YAML (block map / block sequence / flow
// sequences of numbers, comments), which is what the reference's emitter produces.
// SURVEY.md section 8f-4.  Host-side only; nothing here touches the GPU.
#ifndef MAV_TRAJECTORY_GENERATION_IO_H_
#define MAV_TRAJECTORY_GENERATION_IO_H_

#include <string>

#include "mav_trajectory_generation/b200_value_types.h"

namespace mav_trajectory_generation {

bool segmentsToFile(const std::string& filename, const Segment::Vector& segments);
bool segmentsFromFile(const std::string& filename, Segment::Vector* segments);
std::string segmentsToYamlString(const Segment::Vector& segments);
bool segmentsFromYamlString(const std::string& yaml, Segment::Vector* segments);

inline bool trajectoryToFile(const std::string& filename, const Trajectory& trajectory) {
  Segment::Vector segments;
  trajectory.getSegments(&segments);
  return segmentsToFile(filename, segments);
}
inline bool trajectoryFromFile(const std::string& filename, Trajectory* trajectory) {
  Segment::Vector segments;
  if (!segmentsFromFile(filename, &segments) || segments.empty()) return false;
  CHECK_NOTNULL(trajectory)->setSegments(segments);
  return true;
}

// Matlab-readable dump of the sampled flat states (reference src/io.cpp:221-279): the whole trajectory sampled every
// 0.01 s, one row per sample, 27 columns
//   [t_ns, x y z, vx vy vz, ax ay az, jx jy jz, sx sy sz, qw qx qy qz, wx wy wz, w'x w'y w'z, tm]
// where column tm of row j (j < number of segments) holds the accumulated time at the end of segment j.  Text layout as
// the reference's `fs << Eigen::MatrixXd`: 6 significant digits, single-space separated, columns right-aligned to the
// widest entry.
bool sampledTrajectoryStatesToFile(const std::string& filename, const Trajectory& trajectory);

}  // namespace mav_trajectory_generation
#endif
