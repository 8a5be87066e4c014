// b200_value_types.h -- the value types of the reference's public API in ONE header:
// derivative_order (reference motion_defines.h), Polynomial (polynomial.h, data half), Vertex (vertex.h),
// Segment (segment.h, data half), Trajectory (trajectory.h, container half).  The reference's per-class
// header names (vertex.h, segment.h, ...) are kept as forwarding headers so existing #include lines work.
#ifndef MAV_TRAJECTORY_GENERATION_B200_VALUE_TYPES_H_
#define MAV_TRAJECTORY_GENERATION_B200_VALUE_TYPES_H_

#include <cstdint>
#include <map>
#include <ostream>
#include <string>
#include <utility>
#include <vector>

#include "mav_trajectory_generation/eigen_shim.h"
#include "mav_trajectory_generation/glog_shim.h"

// ===== motion_defines ============================================================
// motion_defines.h -- names of the position derivatives (mirror of the reference's
// include/mav_trajectory_generation/motion_defines.h:25-47; same constants, same namespace).


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

// ===== polynomial ============================================================
// polynomial.h -- value type for one polynomial (mirror of the data half of the reference's
// include/mav_trajectory_generation/polynomial.h:37-251: coefficients in INCREASING powers,
// evaluation, derivative coefficients, the base-coefficient table).  The root-finding half of
// the reference class (Jenkins-Traub extrema, polynomial.h:151-186) is outside the hot path
//  and is not provided.



namespace mav_trajectory_generation {

// Container holding the properties of an extremum: time relative to the segment start, value, segment index
// (mirror of the reference's include/mav_trajectory_generation/extremum.h:28-45).
struct Extremum {
  Extremum() : time(0.0), value(0.0), segment_idx(0) {}
  Extremum(double _time, double _value, int _segment_idx) : time(_time), value(_value), segment_idx(_segment_idx) {}
  bool operator<(const Extremum& rhs) const { return value < rhs.value; }
  bool operator>(const Extremum& rhs) const { return value > rhs.value; }
  double time;
  double value;
  int segment_idx;
};
inline std::ostream& operator<<(std::ostream& stream, const Extremum& e) {
  stream << "time: " << e.time << ", value: " << e.value << ", segment idx: " << e.segment_idx << std::endl;
  return stream;
}

class Polynomial {
 public:
  typedef std::vector<Polynomial> Vector;

  static constexpr int kMaxN = 12;                          // reference polynomial.h:44
  static constexpr int kMaxConvolutionSize = 2 * kMaxN - 2;  // :47
  // base_coefficients_(d, j) = j! / (j - d)!  (reference polynomial.h:50, polynomial.cpp:145-160)
  static Eigen::MatrixXd base_coefficients_;

  explicit Polynomial(int N) : N_(N), coefficients_(Eigen::VectorXd::Zero(N)) {}
  Polynomial(int N, const Eigen::VectorXd& coeffs) : N_(N), coefficients_(coeffs) {
    CHECK_EQ(N_, static_cast<int>(coeffs.size())) << "Number of coefficients has to match.";
  }
  explicit Polynomial(const Eigen::VectorXd& coeffs) : N_(static_cast<int>(coeffs.size())), coefficients_(coeffs) {}

  int N() const { return N_; }
  bool operator==(const Polynomial& rhs) const { return coefficients_ == rhs.coefficients_; }
  bool operator!=(const Polynomial& rhs) const { return !(*this == rhs); }
  Polynomial operator+(const Polynomial& rhs) const { return Polynomial(coefficients_ + rhs.coefficients_); }
  Polynomial& operator+=(const Polynomial& rhs) {
    coefficients_ += rhs.coefficients_;
    return *this;
  }
  Polynomial operator*(const Polynomial& rhs) const { return Polynomial(convolve(coefficients_, rhs.coefficients_)); }
  Polynomial operator*(const double& rhs) const { return Polynomial(coefficients_ * rhs); }

  void setCoefficients(const Eigen::VectorXd& coeffs) {
    CHECK_EQ(N_, static_cast<int>(coeffs.size())) << "Number of coefficients has to match.";
    coefficients_ = coeffs;
  }
  // Coefficients of the given derivative (same length N, trailing zeros).
  Eigen::VectorXd getCoefficients(int derivative = 0) const;
  // Fills derivatives 0 .. result->size()-1 at time t.
  void evaluate(double t, Eigen::VectorXd* result) const;
  // One derivative at time t.
  double evaluate(double t, int derivative) const;

  // ---- extrema (reference polynomial.h:151-190, src/polynomial.cpp:27-135).  getRoots returns ALL complex roots of
  // the given derivative; the reference uses Jenkins-Traub (src/rpoly/rpoly_ak1.cpp), here an Aberth-Ehrlich
  // simultaneous iteration with real-axis Newton polishing (roots that are real come back with imag() == 0).
  bool getRoots(int derivative, Eigen::VectorXcd* roots) const;
  static bool selectMinMaxCandidatesFromRoots(double t_start, double t_end,
                                              const Eigen::VectorXcd& roots_derivative_of_derivative,
                                              std::vector<double>* candidates);
  bool computeMinMaxCandidates(double t_start, double t_end, int derivative, std::vector<double>* candidates) const;
  bool selectMinMaxFromRoots(double t_start, double t_end, int derivative,
                             const Eigen::VectorXcd& roots_derivative_of_derivative, std::pair<double, double>* minimum,
                             std::pair<double, double>* maximum) const;
  bool computeMinMax(double t_start, double t_end, int derivative, std::pair<double, double>* minimum,
                     std::pair<double, double>* maximum) const;
  bool selectMinMaxFromCandidates(const std::vector<double>& candidates, int derivative,
                                  std::pair<double, double>* minimum, std::pair<double, double>* maximum) const;

  bool getPolynomialWithAppendedCoefficients(int new_N, Polynomial* new_polynomial) const;
  // Row of the mapping matrix: d-th derivative basis evaluated at t (reference polynomial.h:201-219).
  static void baseCoeffsWithTime(int N, int derivative, double t, Eigen::VectorXd* coeffs);
  static Eigen::VectorXd baseCoeffsWithTime(int N, int derivative, double t) {
    Eigen::VectorXd c(N);
    baseCoeffsWithTime(N, derivative, t, &c);
    return c;
  }
  static Eigen::VectorXd convolve(const Eigen::VectorXd& data, const Eigen::VectorXd& kernel);
  static inline int getConvolutionLength(int data_size, int kernel_size) { return data_size + kernel_size - 1; }
  void scalePolynomialInTime(double scaling_factor);
  void offsetPolynomial(const double offset);

 private:
  int N_;
  Eigen::VectorXd coefficients_;
};

Eigen::MatrixXd computeBaseCoefficients(int N);

}  // namespace mav_trajectory_generation

// ===== vertex ============================================================
// vertex.h -- support point of a path with per-derivative constraints (mirror of the
// reference's include/mav_trajectory_generation/vertex.h:42-177 and src/vertex.cpp).



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

// ===== segment ============================================================
// segment.h -- D polynomials sharing one duration (mirror of the data half of the reference's
// include/mav_trajectory_generation/segment.h:43-128; the extrema search is out of scope).



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

  // ---- extrema of the magnitude over a set of dimensions (reference segment.h:84-110, src/segment.cpp:83-180)
  bool computeMinMaxMagnitudeCandidateTimes(int derivative, double t_start, double t_end,
                                            const std::vector<int>& dimensions,
                                            std::vector<double>* candidate_times) const;
  bool computeMinMaxMagnitudeCandidates(int derivative, double t_start, double t_end,
                                        const std::vector<int>& dimensions, std::vector<Extremum>* candidates) const;
  bool selectMinMaxMagnitudeFromCandidates(int derivative, double t_start, double t_end,
                                           const std::vector<int>& dimensions, const std::vector<Extremum>& candidates,
                                           Extremum* minimum, Extremum* maximum) const;

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

// ===== trajectory ============================================================
// trajectory.h -- container of segments (mirror of the container half of the reference's
// include/mav_trajectory_generation/trajectory.h:31-149: what getTrajectory() needs plus
// evaluation; analytic extrema / time scaling are downstream of the hot path and not provided).



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
  // ---- extrema / time scaling (reference trajectory.h:126-141, src/trajectory.cpp:191-227, 346-429)
  bool computeMinMaxMagnitude(int derivative, const std::vector<int>& dimensions, Extremum* minimum,
                              Extremum* maximum) const;
  bool computeMaxVelocityAndAcceleration(double* v_max, double* a_max) const;
  bool scaleSegmentTimes(double scaling);
  bool scaleSegmentTimesToMeetConstraints(double v_max, double a_max);

  // The reference's sequential sampling walk (src/trajectory.cpp:81-141).
  void evaluateRange(double t_start, double t_end, double dt, int derivative_order,
                     std::vector<Eigen::VectorXd>* result, std::vector<double>* sampling_times = nullptr) const;

  // ---- re-shaping and vertex extraction (reference trajectory.h:86-124, src/trajectory.cpp:143-189, 238-343)
  // One dimension of every segment as a 1-D trajectory.
  Trajectory getTrajectoryWithSingleDimension(int dimension) const;
  // This trajectory's dimensions followed by the other's (same number of segments, same segment times; the
  // polynomial orders may differ, the shorter one is zero-padded by Segment::getSegmentWithAppendedDimension).
  // An empty operand yields the other one.
  bool getTrajectoryWithAppendedDimension(const Trajectory& trajectory_to_append, Trajectory* new_trajectory) const;
  // This trajectory followed in time by the given ones (same D and N); false when a shape differs.
  bool addTrajectories(const std::vector<Trajectory>& trajectories, Trajectory* merged) const;
  // Adds A_r_B to the position polynomials' constant terms of every segment (at most the first 3 dimensions).
  bool offsetTrajectory(const Eigen::VectorXd& A_r_B);
  // Vertex carrying derivatives 0..max_derivative_order of the trajectory at time t / at its start / at its end.
  Vertex getVertexAtTime(double t, int max_derivative_order) const;
  Vertex getStartVertex(int max_derivative_order) const;
  Vertex getGoalVertex(int max_derivative_order) const;
  // The K+1 vertices at the segment boundaries; the 4-D overload splits position (dimensions 0..2) and yaw (3).
  bool getVertices(int max_derivative_order, Vertex::Vector* vertices) const;
  bool getVertices(int max_derivative_order_pos, int max_derivative_order_yaw, Vertex::Vector* pos_vertices,
                   Vertex::Vector* yaw_vertices) const;

 private:
  int D_;
  int N_;
  double max_time_;
  Segment::Vector segments_;
};
}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_B200_VALUE_TYPES_H_
