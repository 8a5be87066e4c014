// polynomial.cpp -- data half of the reference's Polynomial (src/polynomial.cpp:145-216 and the
// inline members of polynomial.h).  Written against the shim/Eigen common subset (element access).
#include "mav_trajectory_generation/polynomial.h"

#include <cmath>
#include <limits>

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
