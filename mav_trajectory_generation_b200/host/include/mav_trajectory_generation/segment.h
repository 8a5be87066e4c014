// segment.h -- D polynomials sharing one duration (mirror of the data half of the reference's
// include/mav_trajectory_generation/segment.h:43-128; the extrema search is out of scope).
#ifndef MAV_TRAJECTORY_GENERATION_SEGMENT_H_
#define MAV_TRAJECTORY_GENERATION_SEGMENT_H_

#include <cstdint>
#include <ostream>
#include <vector>

#include "mav_trajectory_generation/motion_defines.h"
#include "mav_trajectory_generation/polynomial.h"

namespace mav_trajectory_generation {

constexpr double kNumNSecPerSec = 1.0e9;
constexpr double kNumSecPerNsec = 1.0e-9;

class Segment {
 public:
  typedef std::vector<Segment> Vector;

  Segment(int N, int D) : time_(0.0), N_(N), D_(D) { polynomials_.resize(D_, Polynomial(N_)); }
  Segment(const Segment& segment) = default;
  Segment& operator=(const Segment& segment) = default;

  bool operator==(const Segment& rhs) const;
  bool operator!=(const Segment& rhs) const { return !(*this == rhs); }

  int D() const { return D_; }
  int N() const { return N_; }
  double getTime() const { return time_; }
  uint64_t getTimeNSec() const { return static_cast<uint64_t>(kNumNSecPerSec * time_); }
  void setTime(double time_sec) { time_ = time_sec; }
  void setTimeNSec(uint64_t time_ns) { time_ = time_ns * kNumSecPerNsec; }

  Polynomial& operator[](size_t idx);
  const Polynomial& operator[](size_t idx) const;
  const Polynomial::Vector& getPolynomialsRef() const { return polynomials_; }

  Eigen::VectorXd evaluate(double t, int derivative_order = derivative_order::POSITION) const;

  bool getSegmentWithSingleDimension(int dimension, Segment* new_segment) const;
  bool getSegmentWithAppendedDimension(const Segment& segment_to_append, Segment* new_segment) const;
  bool offsetSegment(const Eigen::VectorXd& A_r_B);

 protected:
  Polynomial::Vector polynomials_;
  double time_;

 private:
  int N_;
  int D_;
};

void printSegment(std::ostream& stream, const Segment& s, int derivative);
std::ostream& operator<<(std::ostream& stream, const Segment& s);
std::ostream& operator<<(std::ostream& stream, const std::vector<Segment>& segments);
}  // namespace mav_trajectory_generation
#endif
