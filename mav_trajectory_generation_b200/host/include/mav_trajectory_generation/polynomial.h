// polynomial.h -- value type for one polynomial (mirror of the data half of the reference's
// include/mav_trajectory_generation/polynomial.h:37-251: coefficients in INCREASING powers,
// evaluation, derivative coefficients, the base-coefficient table).  The root-finding half of
// the reference class (Jenkins-Traub extrema, polynomial.h:151-186) is outside the hot path
//  and is not provided.
#ifndef MAV_TRAJECTORY_GENERATION_POLYNOMIAL_H_
#define MAV_TRAJECTORY_GENERATION_POLYNOMIAL_H_

#include <vector>

#include "mav_trajectory_generation/eigen_shim.h"
#include "mav_trajectory_generation/glog_shim.h"

namespace mav_trajectory_generation {

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
#endif
