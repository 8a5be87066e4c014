// b200_extrema.cpp -- analytic extrema of polynomials / segments / trajectories and the time scaling built on
// them: the value-type half of the reference that the constraint-aware callers need
// (Polynomial::getRoots ... computeMinMax, reference src/polynomial.cpp:27-135; Segment::computeMinMaxMagnitude*,
// src/segment.cpp:83-180; Trajectory::computeMinMaxMagnitude / computeMaxVelocityAndAcceleration /
// scaleSegmentTimes / scaleSegmentTimesToMeetConstraints, src/trajectory.cpp:191-227, 346-429).
//
// Root finding: the reference calls Jenkins-Traub (src/rpoly/rpoly_ak1.cpp, 948 lines of translated Fortran).
// This file does NOT translate it.  All roots are found at once with the Aberth-Ehrlich iteration (cubically
// convergent simultaneous Newton with root repulsion, started on a circle of the Cauchy bound radius), then every
// root that is real to working precision is polished by Newton on the real axis and returned with an exactly
// zero imaginary part -- which is what selectMinMaxCandidatesFromRoots keys on (|imag| <= epsilon,
// src/polynomial.cpp:47-50).  Degrees here are <= 2*kMaxN - 3 = 21.
#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <numeric>

#include "mav_trajectory_generation/b200_value_types.h"

namespace mav_trajectory_generation {

namespace {

typedef std::complex<double> cplx;

// p(z) and p'(z) by Horner; c[0..n] in INCREASING powers.
inline void horner(const std::vector<double>& c, cplx z, cplx* p, cplx* dp) {
  const int n = static_cast<int>(c.size()) - 1;
  cplx a = c[n], b = 0.0;
  for (int i = n - 1; i >= 0; --i) {
    b = b * z + a;
    a = a * z + c[i];
  }
  *p = a;
  *dp = b;
}
inline void horner_real(const std::vector<double>& c, double x, double* p, double* dp) {
  const int n = static_cast<int>(c.size()) - 1;
  double a = c[n], b = 0.0;
  for (int i = n - 1; i >= 0; --i) {
    b = b * x + a;
    a = a * x + c[i];
  }
  *p = a;
  *dp = b;
}

// All roots of c[0] + c[1] z + ... + c[n] z^n, c[n] != 0, n >= 1.
std::vector<cplx> aberth_roots(std::vector<double> c) {
  // strip exact zero roots (c[0] == 0) so that the iteration works on a polynomial with p(0) != 0
  std::vector<cplx> roots;
  while (c.size() > 1 && c.front() == 0.0) {
    roots.emplace_back(0.0, 0.0);
    c.erase(c.begin());
  }
  const int n = static_cast<int>(c.size()) - 1;
  if (n < 1) return roots;
  if (n == 1) {
    roots.emplace_back(-c[0] / c[1], 0.0);
    return roots;
  }
  // scale the variable so that roots are O(1): z = s w with s = |c0/cn|^(1/n)
  const double s = std::pow(std::abs(c[0] / c[n]), 1.0 / n);
  std::vector<double> q(c.size());
  double sp = 1.0;
  for (int i = 0; i <= n; ++i) {
    q[i] = c[i] * sp;
    sp *= s;
  }
  const double lead = q[n];
  for (double& v : q) v /= lead;
  // Cauchy bound on |w|
  double bound = 0.0;
  for (int i = 0; i < n; ++i) bound = std::max(bound, std::abs(q[i]));
  bound += 1.0;
  const double r0 = std::min(bound, 2.0);
  std::vector<cplx> w(n);
  const double kPi = 3.14159265358979323846;
  for (int k = 0; k < n; ++k) w[k] = std::polar(r0 * (0.6 + 0.4 * (k + 1.0) / n), 2.0 * kPi * k / n + 0.4);
  for (int it = 0; it < 200; ++it) {
    double move = 0.0;
    for (int k = 0; k < n; ++k) {
      cplx p, dp;
      horner(q, w[k], &p, &dp);
      if (p == cplx(0.0, 0.0)) continue;
      cplx newton = (dp == cplx(0.0, 0.0)) ? cplx(1e-3, 1e-3) : p / dp;
      cplx rep = 0.0;
      for (int j = 0; j < n; ++j)
        if (j != k) {
          const cplx d = w[k] - w[j];
          rep += (d == cplx(0.0, 0.0)) ? cplx(1e6, 0.0) : 1.0 / d;
        }
      const cplx denom = 1.0 - newton * rep;
      const cplx step = (denom == cplx(0.0, 0.0)) ? newton : newton / denom;
      w[k] -= step;
      move = std::max(move, std::abs(step) / std::max(1.0, std::abs(w[k])));
    }
    if (move < 1e-15) break;
  }
  for (int k = 0; k < n; ++k) {
    cplx z = w[k] * s;
    // real to working precision?  polish on the real axis and return an exactly real root
    if (std::abs(z.imag()) <= 1e-7 * std::max(1.0, std::abs(z.real()))) {
      double x = z.real();
      bool ok = false;
      for (int it = 0; it < 8; ++it) {
        double p, dp;
        horner_real(c, x, &p, &dp);
        if (dp == 0.0) break;
        const double dx = p / dp;
        x -= dx;
        if (std::abs(dx) <= 4.0 * std::numeric_limits<double>::epsilon() * std::max(1.0, std::abs(x))) {
          ok = true;
          break;
        }
      }
      // accept only if Newton stayed at the same root (a close complex pair must not collapse onto the axis)
      if (ok && std::abs(x - z.real()) <= 1e-5 * std::max(1.0, std::abs(z.real()))) z = cplx(x, 0.0);
    }
    roots.push_back(z);
  }
  return roots;
}

}  // namespace

// reference src/polynomial.cpp:27-29 + src/rpoly/rpoly_ak1.cpp:70-130 (trailing zeros removed, degree < 1 -> no roots)
bool Polynomial::getRoots(int derivative, Eigen::VectorXcd* roots) const {
  CHECK_NOTNULL(roots);
  const Eigen::VectorXd coeffs = getCoefficients(derivative);
  int last = -1;
  for (int i = static_cast<int>(coeffs.size()) - 1; i >= 0; --i)
    if (std::abs(coeffs[i]) >= std::numeric_limits<double>::min()) {
      last = i;
      break;
    }
  if (last < 1) {  // all zero or constant: no roots
    roots->resize(0, 1);
    return true;
  }
  std::vector<double> c(last + 1);
  for (int i = 0; i <= last; ++i) c[i] = coeffs[i];
  const std::vector<cplx> r = aberth_roots(c);
  roots->resize(static_cast<Eigen::Index>(r.size()), 1);
  for (size_t i = 0; i < r.size(); ++i) (*roots)[static_cast<Eigen::Index>(i)] = r[i];
  for (const cplx& z : r)
    if (!std::isfinite(z.real()) || !std::isfinite(z.imag())) return false;
  return true;
}

// reference src/polynomial.cpp:31-62
bool Polynomial::selectMinMaxCandidatesFromRoots(double t_start, double t_end,
                                                 const Eigen::VectorXcd& roots_derivative_of_derivative,
                                                 std::vector<double>* candidates) {
  CHECK_NOTNULL(candidates);
  if (t_start > t_end) {
    LOG(WARNING) << "t_start is greater than t_end.";
    return false;
  }
  candidates->clear();
  candidates->reserve(static_cast<size_t>(roots_derivative_of_derivative.size()) + 2);
  candidates->push_back(t_start);  // the interval ends are always candidates
  candidates->push_back(t_end);
  for (Eigen::Index i = 0; i < roots_derivative_of_derivative.size(); ++i) {
    const cplx z = roots_derivative_of_derivative[i];
    if (std::abs(z.imag()) > std::numeric_limits<double>::epsilon()) continue;  // only real critical points
    if (z.real() < t_start || z.real() > t_end) continue;
    candidates->push_back(z.real());
  }
  return true;
}

// reference src/polynomial.cpp:64-83
bool Polynomial::computeMinMaxCandidates(double t_start, double t_end, int derivative,
                                         std::vector<double>* candidates) const {
  CHECK_NOTNULL(candidates)->clear();
  if (N_ - derivative - 1 < 0) {
    LOG(WARNING) << "N - derivative - 1 has to be at least 0.";
    return false;
  }
  Eigen::VectorXcd roots;
  if (!getRoots(derivative + 1, &roots)) VLOG(1) << "Couldn't find roots, polynomial may be constant.";
  return selectMinMaxCandidatesFromRoots(t_start, t_end, roots, candidates);
}

// reference src/polynomial.cpp:85-99
bool Polynomial::selectMinMaxFromRoots(double t_start, double t_end, int derivative,
                                       const Eigen::VectorXcd& roots_derivative_of_derivative,
                                       std::pair<double, double>* minimum, std::pair<double, double>* maximum) const {
  std::vector<double> candidates;
  if (!selectMinMaxCandidatesFromRoots(t_start, t_end, roots_derivative_of_derivative, &candidates)) return false;
  return selectMinMaxFromCandidates(candidates, derivative, minimum, maximum);
}

// reference src/polynomial.cpp:101-113
bool Polynomial::computeMinMax(double t_start, double t_end, int derivative, std::pair<double, double>* minimum,
                               std::pair<double, double>* maximum) const {
  std::vector<double> candidates;
  if (!computeMinMaxCandidates(t_start, t_end, derivative, &candidates)) return false;
  return selectMinMaxFromCandidates(candidates, derivative, minimum, maximum);
}

// reference src/polynomial.cpp:115-141
bool Polynomial::selectMinMaxFromCandidates(const std::vector<double>& candidates, int derivative,
                                            std::pair<double, double>* minimum,
                                            std::pair<double, double>* maximum) const {
  CHECK_NOTNULL(minimum);
  CHECK_NOTNULL(maximum);
  if (candidates.empty()) {
    LOG(WARNING) << "Cannot find extrema from an empty candidates vector.";
    return false;
  }
  *minimum = std::make_pair(candidates[0], std::numeric_limits<double>::max());
  *maximum = std::make_pair(candidates[0], std::numeric_limits<double>::lowest());
  for (const double t : candidates) {
    const double value = evaluate(t, derivative);
    if (value < minimum->second) *minimum = std::make_pair(t, value);
    if (value > maximum->second) *maximum = std::make_pair(t, value);
  }
  return true;
}

// reference src/segment.cpp:83-134.  For several dimensions the critical points of |x^(k)(t)|^2 are the roots of
// sum_dim x^(k) x^(k+1) (a polynomial convolution); for one dimension simply the roots of x^(k+1).
bool Segment::computeMinMaxMagnitudeCandidateTimes(int derivative, double t_start, double t_end,
                                                   const std::vector<int>& dimensions,
                                                   std::vector<double>* candidate_times) const {
  CHECK_NOTNULL(candidate_times)->clear();
  if (dimensions.empty()) {
    LOG(WARNING) << "No dimensions specified.";
    return false;
  }
  if (dimensions.size() == 1) {
    const int dim = dimensions[0];
    if (dim < 0 || dim >= D_) {
      LOG(WARNING) << "Specified dimension " << dim << " is out of bounds [0.." << D_ - 1 << "].";
      return false;
    }
    return polynomials_[dim].computeMinMaxCandidates(t_start, t_end, derivative, candidate_times);
  }
  const int n_d = N_ - derivative, n_dd = n_d - 1;
  if (n_dd < 1) {
    LOG(WARNING) << "N - derivative - 1 has to be greater than 0.";
    return false;
  }
  Eigen::VectorXd summed(Polynomial::getConvolutionLength(n_d, n_dd));
  for (const int dim : dimensions) {
    if (dim < 0 || dim >= D_) {
      LOG(WARNING) << "Specified dimension " << dim << " is out of bounds [0.." << D_ - 1 << "].";
      return false;
    }
    // increasing coefficients: the derivative's non-zero part is the head
    const Eigen::VectorXd dk = polynomials_[dim].getCoefficients(derivative);
    const Eigen::VectorXd dk1 = polynomials_[dim].getCoefficients(derivative + 1);
    Eigen::VectorXd a(n_d), b(n_dd);
    for (int i = 0; i < n_d; ++i) a[i] = dk[i];
    for (int i = 0; i < n_dd; ++i) b[i] = dk1[i];
    const Eigen::VectorXd conv = Polynomial::convolve(a, b);
    for (Eigen::Index i = 0; i < summed.size(); ++i) summed[i] += conv[i];
  }
  // the convolved polynomial already IS the derivative of the squared magnitude: "derivative -1"
  return Polynomial(summed).computeMinMaxCandidates(t_start, t_end, -1, candidate_times);
}

// reference src/segment.cpp:136-158
bool Segment::computeMinMaxMagnitudeCandidates(int derivative, double t_start, double t_end,
                                               const std::vector<int>& dimensions,
                                               std::vector<Extremum>* candidates) const {
  CHECK_NOTNULL(candidates);
  std::vector<double> times;
  computeMinMaxMagnitudeCandidateTimes(derivative, t_start, t_end, dimensions, &times);
  candidates->resize(times.size());
  for (size_t i = 0; i < times.size(); ++i) {
    double sq = 0.0;
    for (const int dim : dimensions) {
      const double v = polynomials_[dim].evaluate(times[i], derivative);
      sq += v * v;
    }
    (*candidates)[i] = Extremum(times[i], std::sqrt(sq), 0);
  }
  return true;
}

// reference src/segment.cpp:160-184
bool Segment::selectMinMaxMagnitudeFromCandidates(int /*derivative*/, double t_start, double t_end,
                                                  const std::vector<int>& /*dimensions*/,
                                                  const std::vector<Extremum>& candidates, Extremum* minimum,
                                                  Extremum* maximum) const {
  CHECK_NOTNULL(minimum);
  CHECK_NOTNULL(maximum);
  if (t_start > t_end) {
    LOG(WARNING) << "t_start is greater than t_end.";
    return false;
  }
  minimum->value = std::numeric_limits<double>::max();
  maximum->value = std::numeric_limits<double>::lowest();
  for (const Extremum& c : candidates) {
    if (c.time < t_start || c.time > t_end) continue;
    *maximum = std::max(*maximum, c);
    *minimum = std::min(*minimum, c);
  }
  return true;
}

// reference src/trajectory.cpp:191-227
bool Trajectory::computeMinMaxMagnitude(int derivative, const std::vector<int>& dimensions, Extremum* minimum,
                                        Extremum* maximum) const {
  CHECK_NOTNULL(minimum)->value = std::numeric_limits<double>::max();
  CHECK_NOTNULL(maximum)->value = std::numeric_limits<double>::lowest();
  for (size_t i = 0; i < segments_.size(); ++i) {
    const Segment& seg = segments_[i];
    std::vector<Extremum> candidates;
    if (!seg.computeMinMaxMagnitudeCandidates(derivative, 0.0, seg.getTime(), dimensions, &candidates)) return false;
    Extremum lo, hi;
    if (!seg.selectMinMaxMagnitudeFromCandidates(derivative, 0.0, seg.getTime(), dimensions, candidates, &lo, &hi))
      return false;
    if (lo < *minimum) {
      *minimum = lo;
      minimum->segment_idx = static_cast<int>(i);
    }
    if (hi > *maximum) {
      *maximum = hi;
      maximum->segment_idx = static_cast<int>(i);
    }
  }
  return true;
}

// reference src/trajectory.cpp:346-365
bool Trajectory::computeMaxVelocityAndAcceleration(double* v_max, double* a_max) const {
  std::vector<int> dimensions(D_);
  std::iota(dimensions.begin(), dimensions.end(), 0);
  Extremum v_lo, v_hi, a_lo, a_hi;
  bool ok = computeMinMaxMagnitude(derivative_order::VELOCITY, dimensions, &v_lo, &v_hi);
  ok &= computeMinMaxMagnitude(derivative_order::ACCELERATION, dimensions, &a_lo, &a_hi);
  *CHECK_NOTNULL(v_max) = v_hi.value;
  *CHECK_NOTNULL(a_max) = a_hi.value;
  return ok;
}

// reference src/trajectory.cpp:367-384: stretching time by `scaling` scales coefficient j by scaling^-j
bool Trajectory::scaleSegmentTimes(double scaling) {
  if (scaling < 1.0e-6) return false;
  const double inverse = 1.0 / scaling;
  double total = 0.0;
  for (Segment& seg : segments_) {
    const double t_new = seg.getTime() * scaling;
    for (int d = 0; d < seg.D(); ++d) seg[d].scalePolynomialInTime(inverse);
    seg.setTime(t_new);
    total += t_new;
  }
  max_time_ = total;
  return true;
}

// reference src/trajectory.cpp:386-429 (Liu et al., RA-L 2017): stretch uniformly until v_max / a_max hold
bool Trajectory::scaleSegmentTimesToMeetConstraints(double v_max, double a_max) {
  constexpr int kMaxIterations = 20;
  constexpr double kTolerance = 1e-3;
  bool within_range = false;
  for (int it = 0; it < kMaxIterations; ++it) {
    double v_actual = 0.0, a_actual = 0.0;
    computeMaxVelocityAndAcceleration(&v_actual, &a_actual);
    const double v_violation = v_actual / v_max, a_violation = a_actual / a_max;
    within_range = v_violation <= 1.0 + kTolerance && a_violation <= 1.0 + kTolerance;
    if (within_range) break;
    const double stretch = std::max(1.0, std::max(v_violation, std::sqrt(a_violation)));
    const double inverse = 1.0 / stretch;
    double total = 0.0;
    for (Segment& seg : segments_) {
      const double t_new = seg.getTime() * stretch;
      for (int d = 0; d < seg.D(); ++d) seg[d].scalePolynomialInTime(inverse);
      seg.setTime(t_new);
      total += t_new;
    }
    max_time_ = total;
  }
  return within_range;
}

}  // namespace mav_trajectory_generation
