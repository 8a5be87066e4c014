// vertex.h -- support point of a path with per-derivative constraints (mirror of the
// reference's include/mav_trajectory_generation/vertex.h:42-177 and src/vertex.cpp).
#ifndef MAV_TRAJECTORY_GENERATION_VERTEX_H_
#define MAV_TRAJECTORY_GENERATION_VERTEX_H_

#include <map>
#include <ostream>
#include <utility>
#include <vector>

#include "mav_trajectory_generation/eigen_shim.h"
#include "mav_trajectory_generation/glog_shim.h"
#include "mav_trajectory_generation/motion_defines.h"
#include "mav_trajectory_generation/polynomial.h"

namespace mav_trajectory_generation {

class Vertex {
 public:
  typedef std::vector<Vertex> Vector;
  typedef Eigen::VectorXd ConstraintValue;
  typedef std::pair<int, ConstraintValue> Constraint;
  typedef std::map<int, ConstraintValue> Constraints;

  explicit Vertex(size_t dimension) : D_(static_cast<int>(dimension)) {}
  int D() const { return D_; }

  // Same value in every dimension.
  void addConstraint(int derivative_order, double value) {
    constraints_[derivative_order] = ConstraintValue::Constant(D_, value);
  }
  void addConstraint(int type, const Eigen::VectorXd& constraint);
  bool removeConstraint(int type);
  // Position = constraint, derivatives 1..up_to_derivative = 0.
  void makeStartOrEnd(const Eigen::VectorXd& constraint, int up_to_derivative);
  void makeStartOrEnd(double value, int up_to_derivative) {
    makeStartOrEnd(Eigen::VectorXd::Constant(D_, value), up_to_derivative);
  }
  bool hasConstraint(int derivative_order) const;
  bool getConstraint(int derivative_order, Eigen::VectorXd* constraint) const;
  Constraints::const_iterator cBegin() const { return constraints_.begin(); }
  Constraints::const_iterator cEnd() const { return constraints_.end(); }
  size_t getNumberOfConstraints() const { return constraints_.size(); }
  bool isEqualTol(const Vertex& rhs, double tol) const;
  bool getSubdimension(const std::vector<size_t>& subdimensions, int max_derivative_order, Vertex* subvertex) const;

 private:
  int D_;
  Constraints constraints_;
};

std::ostream& operator<<(std::ostream& stream, const Vertex& v);
std::ostream& operator<<(std::ostream& stream, const std::vector<Vertex>& vertices);

std::vector<double> estimateSegmentTimes(const Vertex::Vector& vertices, double v_max, double a_max);
std::vector<double> estimateSegmentTimesVelocityRamp(const Vertex::Vector& vertices, double v_max, double a_max,
                                                     double time_factor = 1.0);
std::vector<double> estimateSegmentTimesNfabian(const Vertex::Vector& vertices, double v_max, double a_max,
                                                double magic_fabian_constant = 6.5);
double computeTimeVelocityRamp(const Eigen::VectorXd& start, const Eigen::VectorXd& goal, double v_max,
                               double a_max);
inline int getHighestDerivativeFromN(int N) { return N / 2 - 1; }

Vertex::Vector createRandomVertices(int maximum_derivative, size_t n_segments,
                                    const Eigen::VectorXd& minimum_position,
                                    const Eigen::VectorXd& maximum_position, size_t seed = 0);
Vertex::Vector createSquareVertices(int maximum_derivative, const Eigen::Vector3d& center, double side_length,
                                    int rounds);
Vertex::Vector createRandomVertices1D(int maximum_derivative, size_t n_segments, double minimum_position,
                                      double maximum_position, size_t seed = 0);
}  // namespace mav_trajectory_generation
#endif
