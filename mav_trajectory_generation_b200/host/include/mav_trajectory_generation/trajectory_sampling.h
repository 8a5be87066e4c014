// trajectory_sampling.h -- single-object sampling of a Trajectory / Segment into flat states:
// sampleTrajectoryAtTime, sampleTrajectoryInRange, sampleTrajectoryStartDuration, sampleWholeTrajectory,
// sampleSegmentAtTime, sampleFlatStateAtTime (reference include/mav_trajectory_generation/trajectory_sampling.h,
// src/trajectory_sampling.cpp:27-196).  The batched device form of the same sample set is
// mtg_evaluate_range_batch_f64 / BatchPolynomialOptimization::evaluateRange with derivatives {0..4}.
//
// mav_msgs is not part of this tree, so mav_msgs::EigenTrajectoryPoint is mirrored as a plain struct with the same
// field names and meaning (mav_msgs/eigen_mav_msgs.h): world-frame position ... snap, orientation quaternion,
// world-frame angular velocity / acceleration, time from start in nanoseconds, degrees of freedom.  A maintainer
// with mav_msgs on the include path deletes the mirror namespace below and includes <mav_msgs/eigen_mav_msgs.h>.
// Same rules as the reference: at least 3 dimensions; a 4th dimension is yaw (orientation about the world z axis,
// yaw rate / acceleration in the z components); 6 dimensions carry a rotation vector in the last three, turned into
// the orientation quaternion and -- through the left Jacobian of SO(3), w = J(phi) phi', w' = J phi'' + J' phi' --
// into world-frame angular velocity / acceleration (the published formulas of mav_msgs' omegaFromRotationVector /
// omegaDotFromRotationVector; pinned here by finite differences of the rotation itself, tests/cpp).
#ifndef MAV_TRAJECTORY_GENERATION_TRAJECTORY_SAMPLING_H_
#define MAV_TRAJECTORY_GENERATION_TRAJECTORY_SAMPLING_H_

#include <cmath>
#include <cstdint>
#include <vector>

#include "mav_trajectory_generation/trajectory.h"

namespace mav_msgs {

enum MavActuation { DOF4 = 4, DOF6 = 6 };

struct Quaternion {  // Eigen::Quaterniond's coefficients
  double w = 1.0, x = 0.0, y = 0.0, z = 0.0;
};

struct EigenTrajectoryPoint {
  typedef std::vector<EigenTrajectoryPoint> Vector;
  int64_t time_from_start_ns = 0;
  Eigen::Vector3d position_W, velocity_W, acceleration_W, jerk_W, snap_W;
  Quaternion orientation_W_B;
  Eigen::Vector3d angular_velocity_W, angular_acceleration_W;
  MavActuation degrees_of_freedom = DOF4;

  void setFromYaw(double yaw) {
    orientation_W_B.w = std::cos(0.5 * yaw);
    orientation_W_B.x = orientation_W_B.y = 0.0;
    orientation_W_B.z = std::sin(0.5 * yaw);
  }
  void setFromYawRate(double yaw_rate) {
    angular_velocity_W[0] = angular_velocity_W[1] = 0.0;
    angular_velocity_W[2] = yaw_rate;
  }
  void setFromYawAcc(double yaw_acc) {
    angular_acceleration_W[0] = angular_acceleration_W[1] = 0.0;
    angular_acceleration_W[2] = yaw_acc;
  }
  double getYaw() const {
    return std::atan2(2.0 * (orientation_W_B.w * orientation_W_B.z + orientation_W_B.x * orientation_W_B.y),
                      1.0 - 2.0 * (orientation_W_B.y * orientation_W_B.y + orientation_W_B.z * orientation_W_B.z));
  }
  double getYawRate() const { return angular_velocity_W[2]; }
  double getYawAcc() const { return angular_acceleration_W[2]; }
};
typedef std::vector<EigenTrajectoryPoint> EigenTrajectoryPointVector;

}  // namespace mav_msgs

namespace mav_trajectory_generation {

namespace sampling_detail {
const double kNumNanosecondsPerSecond = 1.e9;

struct V3 {
  double v[3];
};
inline V3 cross(const V3& a, const V3& b) {
  return {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}
inline double dot(const V3& a, const V3& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
inline V3 axpy(double s, const V3& a, const V3& b) { return {{s * a.v[0] + b.v[0], s * a.v[1] + b.v[1], s * a.v[2] + b.v[2]}}; }

// coefficients a = (1 - cos t)/t^2, b = (t - sin t)/t^3 of J = I + a [phi]x + b [phi]x^2 and their derivatives in t
inline void jacobianCoefficients(double t, double* a, double* b, double* da, double* db) {
  if (t < 1e-4) {  // series: the closed forms cancel catastrophically near 0
    const double t2 = t * t;
    *a = 0.5 - t2 / 24.0;
    *b = 1.0 / 6.0 - t2 / 120.0;
    *da = -t / 12.0;
    *db = -t / 60.0;
    return;
  }
  const double s = std::sin(t), c = std::cos(t);
  *a = (1.0 - c) / (t * t);
  *b = (t - s) / (t * t * t);
  *da = (t * s - 2.0 * (1.0 - c)) / (t * t * t);
  *db = ((1.0 - c) * t - 3.0 * (t - s)) / (t * t * t * t);
}

// orientation, world-frame angular velocity and acceleration of the rotation vector phi(t) with derivatives dphi, ddphi
inline void rotationVectorKinematics(const V3& phi, const V3& dphi, const V3& ddphi, mav_msgs::Quaternion* q, V3* omega,
                                     V3* omega_dot) {
  const double t = std::sqrt(dot(phi, phi));
  if (t < 1e-12) {
    *q = mav_msgs::Quaternion();
  } else {
    const double s = std::sin(0.5 * t) / t;
    q->w = std::cos(0.5 * t);
    q->x = s * phi.v[0];
    q->y = s * phi.v[1];
    q->z = s * phi.v[2];
    if (q->w < 0.0) {  // the matrix -> quaternion conversion the reference goes through returns w >= 0
      q->w = -q->w;
      q->x = -q->x;
      q->y = -q->y;
      q->z = -q->z;
    }
  }
  double a, b, da, db;
  jacobianCoefficients(t, &a, &b, &da, &db);
  auto applyJ = [&](const V3& x) {  // J x = x + a phi x x + b phi x (phi x x)
    const V3 px = cross(phi, x);
    return axpy(b, cross(phi, px), axpy(a, px, x));
  };
  *omega = applyJ(dphi);
  // J' dphi with t' = phi . dphi / t:  a' phi x dphi + b' phi x (phi x dphi) + b (dphi x (phi x dphi))   (dphi x dphi = 0)
  const double tdot = t > 1e-12 ? dot(phi, dphi) / t : 0.0;
  const V3 pd = cross(phi, dphi);
  V3 jd = {{0.0, 0.0, 0.0}};
  jd = axpy(da * tdot, pd, jd);
  jd = axpy(db * tdot, cross(phi, pd), jd);
  jd = axpy(b, cross(dphi, pd), jd);
  const V3 jdd = applyJ(ddphi);
  *omega_dot = axpy(1.0, jdd, jd);
}

inline V3 tail3(const Eigen::VectorXd& v) { return {{v[v.size() - 3], v[v.size() - 2], v[v.size() - 1]}}; }
inline void head3(const Eigen::VectorXd& v, Eigen::Vector3d* out) {
  for (int i = 0; i < 3; ++i) (*out)[i] = v[i];
}
inline void store(const V3& a, Eigen::Vector3d* out) {
  for (int i = 0; i < 3; ++i) (*out)[i] = a.v[i];
}

// attitude part shared by the single-time and the range samplers (reference trajectory_sampling.cpp:87-105, 170-188)
inline void fillAttitude(int D, const Eigen::VectorXd& position, const Eigen::VectorXd& velocity,
                         const Eigen::VectorXd& acceleration, mav_msgs::EigenTrajectoryPoint* state) {
  state->degrees_of_freedom = mav_msgs::DOF4;
  if (D == 4) {
    state->setFromYaw(position[3]);
    state->setFromYawRate(velocity[3]);
    state->setFromYawAcc(acceleration[3]);
  } else if (D == 6) {
    V3 omega, omega_dot;
    rotationVectorKinematics(tail3(position), tail3(velocity), tail3(acceleration), &state->orientation_W_B, &omega,
                             &omega_dot);
    store(omega, &state->angular_velocity_W);
    store(omega_dot, &state->angular_acceleration_W);
    state->degrees_of_freedom = mav_msgs::DOF6;
  }
}
}  // namespace sampling_detail

// reference trajectory_sampling.cpp:151-196
template <class T>
bool sampleFlatStateAtTime(const T& type, double sample_time, mav_msgs::EigenTrajectoryPoint* state) {
  CHECK_NOTNULL(state);
  if (type.D() < 3) {
    LOG(ERROR) << "Dimension has to be 3, 4, or 6 but is " << type.D();
    return false;
  }
  const Eigen::VectorXd position = type.evaluate(sample_time, derivative_order::POSITION);
  const Eigen::VectorXd velocity = type.evaluate(sample_time, derivative_order::VELOCITY);
  const Eigen::VectorXd acceleration = type.evaluate(sample_time, derivative_order::ACCELERATION);
  sampling_detail::head3(position, &state->position_W);
  sampling_detail::head3(velocity, &state->velocity_W);
  sampling_detail::head3(acceleration, &state->acceleration_W);
  sampling_detail::head3(type.evaluate(sample_time, derivative_order::JERK), &state->jerk_W);
  sampling_detail::head3(type.evaluate(sample_time, derivative_order::SNAP), &state->snap_W);
  sampling_detail::fillAttitude(type.D(), position, velocity, acceleration, state);
  state->time_from_start_ns = static_cast<int64_t>(sample_time * sampling_detail::kNumNanosecondsPerSecond);
  return true;
}

// reference trajectory_sampling.cpp:27-43
inline bool sampleTrajectoryAtTime(const Trajectory& trajectory, double sample_time,
                                   mav_msgs::EigenTrajectoryPoint* state) {
  CHECK_NOTNULL(state);
  if (sample_time < trajectory.getMinTime() || sample_time > trajectory.getMaxTime()) {
    LOG(ERROR) << "Sample time should be within [" << trajectory.getMinTime() << " " << trajectory.getMaxTime()
               << "] but is " << sample_time;
    return false;
  }
  if (trajectory.D() < 3) {
    LOG(ERROR) << "Dimension has to be at least 3, but is " << trajectory.D();
    return false;
  }
  return sampleFlatStateAtTime<Trajectory>(trajectory, sample_time, state);
}

// reference trajectory_sampling.cpp:45-110: five evaluateRange walks (position .. snap), one state per sample,
// time stamps min_time + i * sampling_interval
inline bool sampleTrajectoryInRange(const Trajectory& trajectory, double min_time, double max_time,
                                    double sampling_interval, mav_msgs::EigenTrajectoryPointVector* states) {
  CHECK_NOTNULL(states);
  if (min_time < trajectory.getMinTime() || max_time > trajectory.getMaxTime()) {
    LOG(ERROR) << "Sample time should be within [" << trajectory.getMinTime() << " " << trajectory.getMaxTime()
               << "] but is [" << min_time << " " << max_time << "]";
    return false;
  }
  if (trajectory.D() < 3) {
    LOG(ERROR) << "Dimension has to be at least 3, but is " << trajectory.D();
    return false;
  }
  std::vector<Eigen::VectorXd> sampled[5];
  for (int derivative = 0; derivative < 5; ++derivative)
    trajectory.evaluateRange(min_time, max_time, sampling_interval, derivative, &sampled[derivative]);
  const size_t n_samples = sampled[0].size();
  states->assign(n_samples, mav_msgs::EigenTrajectoryPoint());
  for (size_t i = 0; i < n_samples; ++i) {
    mav_msgs::EigenTrajectoryPoint& state = (*states)[i];
    sampling_detail::head3(sampled[0][i], &state.position_W);
    sampling_detail::head3(sampled[1][i], &state.velocity_W);
    sampling_detail::head3(sampled[2][i], &state.acceleration_W);
    sampling_detail::head3(sampled[3][i], &state.jerk_W);
    sampling_detail::head3(sampled[4][i], &state.snap_W);
    state.time_from_start_ns =
        static_cast<int64_t>((min_time + sampling_interval * i) * sampling_detail::kNumNanosecondsPerSecond);
    sampling_detail::fillAttitude(trajectory.D(), sampled[0][i], sampled[1][i], sampled[2][i], &state);
  }
  return true;
}

inline bool sampleTrajectoryStartDuration(const Trajectory& trajectory, double start_time, double duration,
                                          double sampling_interval, mav_msgs::EigenTrajectoryPointVector* states) {
  return sampleTrajectoryInRange(trajectory, start_time, start_time + duration, sampling_interval, states);
}

inline bool sampleWholeTrajectory(const Trajectory& trajectory, double sampling_interval,
                                  mav_msgs::EigenTrajectoryPoint::Vector* states) {
  return sampleTrajectoryInRange(trajectory, trajectory.getMinTime(), trajectory.getMaxTime(), sampling_interval, states);
}

// reference trajectory_sampling.cpp:137-149
inline bool sampleSegmentAtTime(const Segment& segment, double sample_time, mav_msgs::EigenTrajectoryPoint* state) {
  CHECK_NOTNULL(state);
  if (sample_time < 0.0 || sample_time > segment.getTime()) {
    LOG(ERROR) << "Sample time should be within [" << 0.0 << " " << segment.getTime() << "] but is " << sample_time;
    return false;
  }
  return sampleFlatStateAtTime<Segment>(segment, sample_time, state);
}

}  // namespace mav_trajectory_generation

#endif  // MAV_TRAJECTORY_GENERATION_TRAJECTORY_SAMPLING_H_
