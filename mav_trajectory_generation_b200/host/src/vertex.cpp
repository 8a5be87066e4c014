// vertex.cpp -- Vertex members, time-allocation heuristics and the random fixture generator
// (mirror of the reference's src/vertex.cpp; createRandomVertices is draw-for-draw identical
// under libstdc++: std::mt19937 + one uniform_real_distribution per dimension + 0.2 m rejection,
// reference vertex.cpp:37-72).
#include "mav_trajectory_generation/vertex.h"

#include <algorithm>
#include <cmath>
#include <random>

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
