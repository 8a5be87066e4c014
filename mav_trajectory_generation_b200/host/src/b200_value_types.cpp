// b200_value_types.cpp -- members of the value types declared in b200_value_types.h (Polynomial, Vertex and
// its fixture / time-allocation helpers, Segment, Trajectory, derivative names).
#include "mav_trajectory_generation/b200_value_types.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <random>

// ===== motion_defines ============================================================
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

// ===== polynomial ============================================================
// polynomial.cpp -- data half of the reference's Polynomial (src/polynomial.cpp:145-216 and the
// inline members of polynomial.h).  Written against the shim/Eigen common subset (element access).


namespace mav_trajectory_generation {

// B(d, j) = j (j-1) ... (j-d+1): each row is the previous one times the falling factor.
Eigen::MatrixXd computeBaseCoefficients(int N) {
  Eigen::MatrixXd b(N, N);
  b.setZero();
  for (int j = 0; j < N; ++j) b(0, j) = 1.0;
  for (int d = 1; d < N; ++d)
    for (int j = d; j < N; ++j) b(d, j) = b(d - 1, j) * static_cast<double>(j - d + 1);
  return b;
}

Eigen::MatrixXd Polynomial::base_coefficients_ = computeBaseCoefficients(Polynomial::kMaxConvolutionSize);

Eigen::VectorXd Polynomial::getCoefficients(int derivative) const {
  CHECK_LE(derivative, N_);
  if (derivative == 0) return coefficients_;
  Eigen::VectorXd result(N_);
  result.setZero();
  for (int j = derivative; j < N_; ++j) result[j - derivative] = base_coefficients_(derivative, j) * coefficients_[j];
  return result;
}

double Polynomial::evaluate(double t, int derivative) const {
  if (derivative >= N_) return 0.0;
  double acc = base_coefficients_(derivative, N_ - 1) * coefficients_[N_ - 1];
  for (int j = N_ - 2; j >= derivative; --j) acc = acc * t + base_coefficients_(derivative, j) * coefficients_[j];
  return acc;
}

void Polynomial::evaluate(double t, Eigen::VectorXd* result) const {
  CHECK_LE(static_cast<int>(result->size()), N_);
  for (int d = 0; d < static_cast<int>(result->size()); ++d) (*result)[d] = evaluate(t, d);
}

bool Polynomial::getPolynomialWithAppendedCoefficients(int new_N, Polynomial* new_polynomial) const {
  if (new_N == N_) {
    *new_polynomial = *this;
    return true;
  }
  if (new_N < N_) {
    LOG(WARNING) << "You shan't decrease the number of coefficients.";
    *new_polynomial = *this;
    return false;
  }
  Eigen::VectorXd coeffs(new_N);
  coeffs.setZero();
  for (int i = 0; i < N_; ++i) coeffs[i] = coefficients_[i];
  *new_polynomial = Polynomial(coeffs);
  return true;
}

void Polynomial::baseCoeffsWithTime(int N, int derivative, double t, Eigen::VectorXd* coeffs) {
  CHECK_LT(derivative, N);
  CHECK_GE(derivative, 0);
  coeffs->resize(N, 1);
  coeffs->setZero();
  (*coeffs)[derivative] = base_coefficients_(derivative, derivative);
  if (std::abs(t) < std::numeric_limits<double>::epsilon()) return;
  double t_power = t;
  for (int j = derivative + 1; j < N; ++j) {
    (*coeffs)[j] = base_coefficients_(derivative, j) * t_power;
    t_power *= t;
  }
}

Eigen::VectorXd Polynomial::convolve(const Eigen::VectorXd& data, const Eigen::VectorXd& kernel) {
  const int nd = static_cast<int>(data.size()), nk = static_cast<int>(kernel.size());
  Eigen::VectorXd out(getConvolutionLength(nd, nk));
  out.setZero();
  for (int i = 0; i < nd; ++i)
    for (int k = 0; k < nk; ++k) out[i + k] += data[i] * kernel[k];
  return out;
}

void Polynomial::scalePolynomialInTime(double scaling_factor) {
  double scale = 1.0;
  for (int n = 0; n < N_; ++n) {
    coefficients_[n] *= scale;
    scale *= scaling_factor;
  }
}

void Polynomial::offsetPolynomial(const double offset) {
  if (N_ > 0) coefficients_[0] += offset;
}

}  // namespace mav_trajectory_generation

// ===== vertex ============================================================
// vertex.cpp -- Vertex members, time-allocation heuristics and the random fixture generator
// (mirror of the reference's src/vertex.cpp; createRandomVertices is draw-for-draw identical
// under libstdc++: std::mt19937 + one uniform_real_distribution per dimension + 0.2 m rejection,
// reference vertex.cpp:37-72).


namespace mav_trajectory_generation {

void Vertex::addConstraint(int derivative_order, const Eigen::VectorXd& constraint) {
  CHECK_EQ(static_cast<long>(constraint.rows()), static_cast<long>(D_));
  constraints_[derivative_order] = constraint;
}

bool Vertex::removeConstraint(int type) { return constraints_.erase(type) > 0; }

void Vertex::makeStartOrEnd(const Eigen::VectorXd& constraint, int up_to_derivative) {
  addConstraint(derivative_order::POSITION, constraint);
  for (int i = 1; i <= up_to_derivative; ++i) constraints_[i] = ConstraintValue::Zero(D_);
}

bool Vertex::hasConstraint(int derivative_order) const { return constraints_.count(derivative_order) > 0; }

bool Vertex::getConstraint(int derivative_order, Eigen::VectorXd* value) const {
  CHECK_NOTNULL(value);
  const auto it = constraints_.find(derivative_order);
  if (it == constraints_.end()) return false;
  *value = it->second;
  return true;
}

bool Vertex::isEqualTol(const Vertex& rhs, double tol) const {
  if (constraints_.size() != rhs.constraints_.size()) return false;
  for (const auto& kv : constraints_) {
    const auto other = rhs.constraints_.find(kv.first);
    if (other == rhs.constraints_.end()) return false;
    if (!((kv.second - other->second).isZero(tol))) return false;
  }
  return true;
}

bool Vertex::getSubdimension(const std::vector<size_t>& subdimensions, int max_derivative_order,
                             Vertex* subvertex) const {
  CHECK_NOTNULL(subvertex);
  *subvertex = Vertex(subdimensions.size());
  for (size_t s : subdimensions)
    if (s >= static_cast<size_t>(D_)) return false;
  for (const auto& kv : constraints_) {
    if (kv.first > max_derivative_order) continue;
    ConstraintValue sub(static_cast<int>(subdimensions.size()));
    for (size_t i = 0; i < subdimensions.size(); ++i) sub[i] = kv.second[subdimensions[i]];
    subvertex->addConstraint(kv.first, sub);
  }
  return true;
}

std::ostream& operator<<(std::ostream& stream, const Vertex& v) {
  stream << "constraints: " << std::endl;
  for (auto it = v.cBegin(); it != v.cEnd(); ++it) {
    stream << "  type: " << positionDerivativeToString(it->first) << "  value: [";
    for (int d = 0; d < static_cast<int>(it->second.size()); ++d) stream << (d ? ", " : "") << it->second[d];
    stream << "]" << std::endl;
  }
  return stream;
}

std::ostream& operator<<(std::ostream& stream, const std::vector<Vertex>& vertices) {
  for (const Vertex& v : vertices) stream << v << std::endl;
  return stream;
}

namespace {
double waypointDistance(const Vertex& a, const Vertex& b) {
  Eigen::VectorXd start, end;
  a.getConstraint(derivative_order::POSITION, &start);
  b.getConstraint(derivative_order::POSITION, &end);
  return (end - start).norm();
}
}  // namespace

std::vector<double> estimateSegmentTimes(const Vertex::Vector& vertices, double v_max, double a_max) {
  return estimateSegmentTimesNfabian(vertices, v_max, a_max);
}

std::vector<double> estimateSegmentTimesVelocityRamp(const Vertex::Vector& vertices, double v_max, double a_max,
                                                     double time_factor) {
  CHECK_GE(vertices.size(), 2u);
  (void)time_factor;  // unused by the reference as well (vertex.cpp:233-253)
  constexpr double kMinSegmentTime = 0.1;
  std::vector<double> times;
  times.reserve(vertices.size() - 1);
  for (size_t i = 0; i + 1 < vertices.size(); ++i) {
    Eigen::VectorXd start, end;
    vertices[i].getConstraint(derivative_order::POSITION, &start);
    vertices[i + 1].getConstraint(derivative_order::POSITION, &end);
    times.push_back(std::max(kMinSegmentTime, computeTimeVelocityRamp(start, end, v_max, a_max)));
  }
  return times;
}

// t = 2 d / v * (1 + c * v / a * exp(-2 d / v))   (reference vertex.cpp:255-272)
std::vector<double> estimateSegmentTimesNfabian(const Vertex::Vector& vertices, double v_max, double a_max,
                                                double magic_fabian_constant) {
  CHECK_GE(vertices.size(), 2u);
  std::vector<double> times;
  times.reserve(vertices.size() - 1);
  for (size_t i = 0; i + 1 < vertices.size(); ++i) {
    const double distance = waypointDistance(vertices[i], vertices[i + 1]);
    const double t =
        distance / v_max * 2 * (1.0 + magic_fabian_constant * v_max / a_max * exp(-distance / v_max * 2));
    times.push_back(t);
  }
  return times;
}

double computeTimeVelocityRamp(const Eigen::VectorXd& start, const Eigen::VectorXd& goal, double v_max,
                               double a_max) {
  const double distance = (start - goal).norm();
  const double acc_time = v_max / a_max;
  const double acc_distance = 0.5 * v_max * acc_time;
  if (distance < 2.0 * acc_distance) return 2.0 * std::sqrt(distance / a_max);
  return 2.0 * acc_time + (distance - 2.0 * acc_distance) / v_max;
}

Vertex::Vector createRandomVertices(int maximum_derivative, size_t n_segments, const Eigen::VectorXd& pos_min,
                                    const Eigen::VectorXd& pos_max, size_t seed) {
  CHECK_GE(static_cast<int>(n_segments), 1);
  CHECK_EQ(pos_min.size(), pos_max.size());
  CHECK_GE((pos_max - pos_min).norm(), 0.2);
  CHECK_GT(maximum_derivative, 0);
  const size_t dimension = static_cast<size_t>(pos_min.size());
  std::mt19937 generator(seed);
  std::vector<std::uniform_real_distribution<double> > box(dimension);
  for (size_t d = 0; d < dimension; ++d) box[d] = std::uniform_real_distribution<double>(pos_min[d], pos_max[d]);
  auto draw = [&](Eigen::VectorXd* p) {
    for (size_t d = 0; d < dimension; ++d) (*p)[d] = box[d](generator);
  };
  const double min_distance = 0.2;
  Vertex::Vector vertices;
  vertices.reserve(n_segments + 1);
  Eigen::VectorXd last(static_cast<int>(dimension));
  draw(&last);
  vertices.push_back(Vertex(dimension));
  vertices.front().makeStartOrEnd(last, maximum_derivative);
  for (size_t i = 1; i <= n_segments; ++i) {
    Eigen::VectorXd pos(static_cast<int>(dimension));
    do {
      draw(&pos);
    } while (!((pos - last).norm() > min_distance));
    Vertex v(dimension);
    v.addConstraint(derivative_order::POSITION, pos);
    vertices.push_back(v);
    last = pos;
  }
  vertices.back().makeStartOrEnd(last, maximum_derivative);
  return vertices;
}

Vertex::Vector createSquareVertices(int maximum_derivative, const Eigen::Vector3d& center, double side_length,
                                    int rounds) {
  const double s = side_length / 2.0;
  const double dx[4] = {-s, -s, s, s}, dy[4] = {-s, s, s, -s};
  std::vector<Vertex> corner;
  for (int c = 0; c < 4; ++c) {
    Eigen::VectorXd p(3);
    p[0] = center[0] + dx[c];
    p[1] = center[1] + dy[c];
    p[2] = center[2];
    Vertex v(3);
    v.addConstraint(derivative_order::POSITION, p);
    corner.push_back(v);
  }
  Eigen::VectorXd first(3);
  corner[0].getConstraint(derivative_order::POSITION, &first);
  Vertex::Vector vertices;
  vertices.reserve(4 * rounds + 1);
  vertices.push_back(corner[0]);
  vertices.front().makeStartOrEnd(first, maximum_derivative);
  for (int i = 0; i < rounds; ++i)
    for (int c = 1; c <= 4; ++c) vertices.push_back(corner[c % 4]);
  vertices.back().makeStartOrEnd(first, maximum_derivative);
  return vertices;
}

Vertex::Vector createRandomVertices1D(int maximum_derivative, size_t n_segments, double pos_min, double pos_max,
                                      size_t seed) {
  return createRandomVertices(maximum_derivative, n_segments, Eigen::VectorXd::Constant(1, pos_min),
                              Eigen::VectorXd::Constant(1, pos_max), seed);
}

}  // namespace mav_trajectory_generation

// ===== segment ============================================================
// segment.cpp / trajectory -- container members (mirror of the data half of the reference's
// src/segment.cpp:25-81,187-245 and src/trajectory.cpp:26-141).



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
  // Common order and duration (reference src/segment.cpp:201-262): the shorter-lived operand is re-parametrised to the
  // longer duration (p(t) -> p(t * T_short / T_long)), the lower order is zero-padded.  (The reference applies the time
  // scaling only when both orders are equal and silently drops it otherwise; here it is applied in both cases.)
  const int new_N = std::max(N_, segment_to_append.N());
  const int new_D = D_ + segment_to_append.D();
  const double new_time = std::max(time_, segment_to_append.getTime());
  Segment mine = *this, other = segment_to_append;
  if (new_time > 0.0) {
    if (time_ < new_time)
      for (int d = 0; d < D_; ++d) mine[d].scalePolynomialInTime(time_ / new_time);
    else if (segment_to_append.getTime() < new_time)
      for (int d = 0; d < other.D(); ++d) other[d].scalePolynomialInTime(segment_to_append.getTime() / new_time);
  }
  *new_segment = Segment(new_N, new_D);
  bool ok = true;
  for (int i = 0; i < new_D; ++i) {
    const Polynomial& src = i < D_ ? mine[i] : other[i - D_];
    Polynomial widened(new_N);
    ok = src.getPolynomialWithAppendedCoefficients(new_N, &widened) && ok;
    (*new_segment)[i] = widened;
  }
  new_segment->setTime(new_time);
  return ok;
}

bool Segment::offsetSegment(const Eigen::VectorXd& A_r_B) {
  // only the translational part moves (the first three dimensions at most), as in the reference (src/segment.cpp:264-276)
  const int n = std::min(D_, 3);
  if (static_cast<int>(A_r_B.size()) < n) {
    LOG(WARNING) << "Offset vector size smaller than segment dimension.";
    return false;
  }
  for (int d = 0; d < n; ++d) polynomials_[d].offsetPolynomial(A_r_B[d]);
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

Trajectory Trajectory::getTrajectoryWithSingleDimension(int dimension) const {
  CHECK_LT(dimension, D_);
  Segment::Vector picked;
  picked.reserve(segments_.size());
  for (const Segment& s : segments_) {
    Segment one(N_, 1);
    CHECK(s.getSegmentWithSingleDimension(dimension, &one));
    picked.push_back(one);
  }
  Trajectory out;
  if (!picked.empty()) out.setSegments(picked);
  return out;
}

bool Trajectory::getTrajectoryWithAppendedDimension(const Trajectory& trajectory_to_append,
                                                    Trajectory* new_trajectory) const {
  CHECK_NOTNULL(new_trajectory);
  if (N_ == 0 || D_ == 0) {  // nothing here yet: the result is the other trajectory
    *new_trajectory = trajectory_to_append;
    return true;
  }
  if (trajectory_to_append.N() == 0 || trajectory_to_append.D() == 0) {
    *new_trajectory = *this;
    return true;
  }
  CHECK_EQ(K(), trajectory_to_append.K());
  Segment::Vector joined;
  joined.reserve(segments_.size());
  for (size_t k = 0; k < segments_.size(); ++k) {
    Segment both(0, 0);
    if (!segments_[k].getSegmentWithAppendedDimension(trajectory_to_append.segments()[k], &both)) return false;
    joined.push_back(both);
  }
  new_trajectory->setSegments(joined);
  return true;
}

bool Trajectory::addTrajectories(const std::vector<Trajectory>& trajectories, Trajectory* merged) const {
  CHECK_NOTNULL(merged);
  *merged = *this;
  for (const Trajectory& t : trajectories) {
    if (t.D() != D_ || t.N() != N_) {
      LOG(WARNING) << "addTrajectories: shape (D, N) = (" << t.D() << ", " << t.N() << ") does not match (" << D_
                   << ", " << N_ << ")";
      return false;
    }
    merged->addSegments(t.segments());
  }
  return true;
}

bool Trajectory::offsetTrajectory(const Eigen::VectorXd& A_r_B) {
  if (A_r_B.size() < std::min(D_, 3)) {
    LOG(WARNING) << "Offset vector size smaller than trajectory dimension.";
    return false;
  }
  for (Segment& s : segments_)
    if (!s.offsetSegment(A_r_B)) return false;
  return true;
}

Vertex Trajectory::getVertexAtTime(double t, int max_derivative_order) const {
  Vertex v(D_);
  for (int derivative = 0; derivative <= max_derivative_order; ++derivative)
    v.addConstraint(derivative, evaluate(t, derivative));
  return v;
}

Vertex Trajectory::getStartVertex(int max_derivative_order) const { return getVertexAtTime(0.0, max_derivative_order); }

Vertex Trajectory::getGoalVertex(int max_derivative_order) const {
  return getVertexAtTime(max_time_, max_derivative_order);
}

bool Trajectory::getVertices(int max_derivative_order, Vertex::Vector* vertices) const {
  CHECK_NOTNULL(vertices);
  vertices->assign(segments_.size() + 1, Vertex(D_));
  vertices->front() = getStartVertex(max_derivative_order);
  double t = 0.0;  // accumulated exactly like the reference: boundary k sits at the running sum of the segment times
  for (size_t i = 0; i < segments_.size(); ++i) {
    t += segments_[i].getTime();
    (*vertices)[i + 1] = getVertexAtTime(t, max_derivative_order);
  }
  return true;
}

bool Trajectory::getVertices(int max_derivative_order_pos, int max_derivative_order_yaw, Vertex::Vector* pos_vertices,
                             Vertex::Vector* yaw_vertices) const {
  CHECK_NOTNULL(pos_vertices);
  CHECK_NOTNULL(yaw_vertices);
  const std::vector<size_t> pos_dims = {0, 1, 2};
  const std::vector<size_t> yaw_dims = {3};
  const int order = std::max(max_derivative_order_pos, max_derivative_order_yaw);
  pos_vertices->assign(segments_.size() + 1, Vertex(3));
  yaw_vertices->assign(segments_.size() + 1, Vertex(1));
  double t = 0.0;
  for (size_t i = 0; i <= segments_.size(); ++i) {
    if (i > 0) t += segments_[i - 1].getTime();
    const Vertex full = getVertexAtTime(t, order);
    if (!full.getSubdimension(pos_dims, max_derivative_order_pos, &(*pos_vertices)[i])) return false;
    if (!full.getSubdimension(yaw_dims, max_derivative_order_yaw, &(*yaw_vertices)[i])) return false;
  }
  return true;
}

void Trajectory::evaluateRange(double t_start, double t_end, double dt, int derivative_order,
                               std::vector<Eigen::VectorXd>* result, std::vector<double>* sampling_times) const {
  // The reference's walk (src/trajectory.cpp:81-141), kept step for step so that sample counts and sample
  // times are identical for drop-in callers: the sample clock `accumulated_time` starts at the START of the
  // segment that contains t_start, advances by dt per sample, is what the loop compares with t_end (so t_end
  // itself is excluded) and what sampling_times reports; a sample whose local time exceeds the segment time
  // moves to the next segment; the walk ends after the last segment.  mtg_evaluate_range_batch_f64 is the
  // batched device version of exactly this loop.
  CHECK_NOTNULL(result)->clear();
  if (sampling_times) sampling_times->clear();
  double accumulated_time = 0.0;
  size_t i = 0;
  for (i = 0; i < segments_.size(); ++i) {
    accumulated_time += segments_[i].getTime();
    if (accumulated_time > t_start) break;
  }
  if (t_start > accumulated_time) {
    LOG(ERROR) << "Start time out of range of the trajectory!";
    return;
  }
  if (i >= segments_.size()) return;  // t_start == end of the trajectory (the reference reads past the vector here)
  accumulated_time -= segments_[i].getTime();
  double time_in_segment = t_start - accumulated_time;
  while (accumulated_time < t_end) {
    if (time_in_segment > segments_[i].getTime()) {
      time_in_segment = time_in_segment - segments_[i].getTime();
      i++;
      if (i >= segments_.size()) break;
      continue;
    }
    result->push_back(segments_[i].evaluate(time_in_segment, derivative_order));
    if (sampling_times) sampling_times->push_back(accumulated_time);
    time_in_segment += dt;
    accumulated_time += dt;
  }
}

}  // namespace mav_trajectory_generation
