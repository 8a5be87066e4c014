// trajectory.h -- container of segments (mirror of the container half of the reference's
// include/mav_trajectory_generation/trajectory.h:31-149: what getTrajectory() needs plus
// evaluation; analytic extrema / time scaling are downstream of the hot path and not provided).
#ifndef MAV_TRAJECTORY_GENERATION_TRAJECTORY_H_
#define MAV_TRAJECTORY_GENERATION_TRAJECTORY_H_

#include <vector>

#include "mav_trajectory_generation/segment.h"
#include "mav_trajectory_generation/vertex.h"

namespace mav_trajectory_generation {

class Trajectory {
 public:
  Trajectory() : D_(0), N_(0), max_time_(0.0) {}

  bool operator==(const Trajectory& rhs) const;
  bool operator!=(const Trajectory& rhs) const { return !(*this == rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  int K() const { return static_cast<int>(segments_.size()); }
  bool empty() const { return segments_.empty(); }
  void clear() {
    segments_.clear();
    D_ = N_ = 0;
    max_time_ = 0.0;
  }
  void setSegments(const Segment::Vector& segments) {
    CHECK(!segments.empty());
    D_ = segments.front().D();
    N_ = segments.front().N();
    max_time_ = 0.0;
    segments_.clear();
    addSegments(segments);
  }
  void addSegments(const Segment::Vector& segments) {
    for (const Segment& segment : segments) {
      CHECK_EQ(segment.D(), D_);
      CHECK_EQ(segment.N(), N_);
      max_time_ += segment.getTime();
    }
    segments_.insert(segments_.end(), segments.begin(), segments.end());
  }
  void getSegments(Segment::Vector* segments) const { *CHECK_NOTNULL(segments) = segments_; }
  const Segment::Vector& segments() const { return segments_; }
  double getMinTime() const { return 0.0; }
  double getMaxTime() const { return max_time_; }
  std::vector<double> getSegmentTimes() const;

  // Value of one derivative at time t (clamped into the last segment like the reference,
  // trajectory.cpp:48-79).
  Eigen::VectorXd evaluate(double t, int derivative_order = derivative_order::POSITION) const;
  // Samples [t_start, t_end] every dt.
  void evaluateRange(double t_start, double t_end, double dt, int derivative_order,
                     std::vector<Eigen::VectorXd>* result, std::vector<double>* sampling_times = nullptr) const;

 private:
  int D_;
  int N_;
  double max_time_;
  Segment::Vector segments_;
};
}  // namespace mav_trajectory_generation
#endif
