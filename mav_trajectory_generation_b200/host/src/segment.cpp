// segment.cpp / trajectory -- container members (mirror of the data half of the reference's
// src/segment.cpp:25-81,187-245 and src/trajectory.cpp:26-141).
#include "mav_trajectory_generation/segment.h"

#include <cmath>

#include "mav_trajectory_generation/trajectory.h"

namespace mav_trajectory_generation {

bool Segment::operator==(const Segment& rhs) const {
  if (D_ != rhs.D_ || N_ != rhs.N_) return false;
  if (std::abs(time_ - rhs.time_) > kNumSecPerNsec) return false;  // times compared at ns resolution
  for (int i = 0; i < D_; ++i)
    if (polynomials_[i] != rhs.polynomials_[i]) return false;
  return true;
}

Polynomial& Segment::operator[](size_t idx) {
  CHECK_LT(idx, static_cast<size_t>(D_));
  return polynomials_[idx];
}

const Polynomial& Segment::operator[](size_t idx) const {
  CHECK_LT(idx, static_cast<size_t>(D_));
  return polynomials_[idx];
}

Eigen::VectorXd Segment::evaluate(double t, int derivative) const {
  Eigen::VectorXd result(D_);
  result.setZero();
  for (int d = 0; d < D_; ++d) result[d] = polynomials_[d].evaluate(t, derivative);
  return result;
}

bool Segment::getSegmentWithSingleDimension(int dimension, Segment* new_segment) const {
  if (dimension < 0 || dimension >= D_) return false;
  *new_segment = Segment(N_, 1);
  (*new_segment)[0] = polynomials_[dimension];
  new_segment->setTime(time_);
  return true;
}

bool Segment::getSegmentWithAppendedDimension(const Segment& segment_to_append, Segment* new_segment) const {
  if (N_ == 0 || D_ == 0) {
    *new_segment = segment_to_append;
    return true;
  }
  if (segment_to_append.N() == 0 || segment_to_append.D() == 0) {
    *new_segment = *this;
    return true;
  }
  if (std::abs(time_ - segment_to_append.getTime()) > kNumSecPerNsec) return false;
  const int new_N = N_ > segment_to_append.N() ? N_ : segment_to_append.N();
  const int new_D = D_ + segment_to_append.D();
  *new_segment = Segment(new_N, new_D);
  bool ok = true;
  for (int i = 0; i < new_D; ++i) {
    const Polynomial& src = i < D_ ? polynomials_[i] : segment_to_append[i - D_];
    Polynomial widened(new_N);
    ok = src.getPolynomialWithAppendedCoefficients(new_N, &widened) && ok;
    (*new_segment)[i] = widened;
  }
  new_segment->setTime(time_);
  return ok;
}

bool Segment::offsetSegment(const Eigen::VectorXd& A_r_B) {
  if (static_cast<int>(A_r_B.size()) < D_) return false;
  for (int d = 0; d < D_; ++d) polynomials_[d].offsetPolynomial(A_r_B[d]);
  return true;
}

void printSegment(std::ostream& stream, const Segment& s, int derivative) {
  CHECK(derivative >= 0 && derivative < s.N());
  stream << "t: " << s.getTime() << std::endl;
  stream << " coefficients for " << positionDerivativeToString(derivative) << ": " << std::endl;
  for (int i = 0; i < s.D(); ++i) {
    const Eigen::VectorXd c = s[i].getCoefficients(derivative);
    stream << "dim " << i << ": " << std::endl << "[";
    for (int j = 0; j < static_cast<int>(c.size()); ++j) stream << (j ? ", " : "") << c[j];
    stream << "]" << std::endl;
  }
}

std::ostream& operator<<(std::ostream& stream, const Segment& s) {
  printSegment(stream, s, derivative_order::POSITION);
  return stream;
}

std::ostream& operator<<(std::ostream& stream, const std::vector<Segment>& segments) {
  for (const Segment& s : segments) stream << s << std::endl;
  return stream;
}

// ---- Trajectory -----------------------------------------------------------------------------
bool Trajectory::operator==(const Trajectory& rhs) const {
  if (segments_.size() != rhs.segments_.size()) return false;
  for (size_t i = 0; i < segments_.size(); ++i)
    if (segments_[i] != rhs.segments_[i]) return false;
  return true;
}

std::vector<double> Trajectory::getSegmentTimes() const {
  std::vector<double> times;
  for (const Segment& s : segments_) times.push_back(s.getTime());
  return times;
}

Eigen::VectorXd Trajectory::evaluate(double t, int derivative_order) const {
  CHECK(!segments_.empty());
  // Same conventions as the reference (src/trajectory.cpp:48-79): a time that falls on a vertex belongs to
  // the segment on its right; t == total time evaluates the end of the last segment; t beyond the end is an
  // error and yields zeros.
  double start = 0.0;
  size_t i = 0;
  for (; i < segments_.size(); ++i) {
    if (start + segments_[i].getTime() > t) break;
    start += segments_[i].getTime();
  }
  if (i == segments_.size()) {
    if (t > start) {
      LOG(ERROR) << "Time out of range of the trajectory!";
      return Eigen::VectorXd::Zero(D_);
    }
    i = segments_.size() - 1;
    start -= segments_[i].getTime();
  }
  return segments_[i].evaluate(t - start, derivative_order);
}

void Trajectory::evaluateRange(double t_start, double t_end, double dt, int derivative_order,
                               std::vector<Eigen::VectorXd>* result, std::vector<double>* sampling_times) const {
  CHECK_NOTNULL(result)->clear();
  if (sampling_times) sampling_times->clear();
  CHECK_GT(dt, 0.0);
  const size_t n = static_cast<size_t>(std::floor((t_end - t_start) / dt + 1e-9)) + 1;
  for (size_t k = 0; k < n; ++k) {
    const double t = t_start + static_cast<double>(k) * dt;
    result->push_back(evaluate(t, derivative_order));
    if (sampling_times) sampling_times->push_back(t);
  }
}

}  // namespace mav_trajectory_generation
