// ros_conversions.h -- Trajectory <-> polynomial trajectory message mapping, the wire format
// existing consumers of the reference speak (reference mav_trajectory_generation_ros/src/ros_conversions.cpp:25-178).
//
// ROS / mav_planning_msgs are not part of this tree, so the messages are mirrored as plain structs with the SAME
// field names and meaning as mav_planning_msgs/PolynomialSegment(.msg), PolynomialSegment4D, PolynomialTrajectory
// and PolynomialTrajectory4D: per segment the number of coefficients, the segment time (ros::Duration there; here
// the same quantity in integer nanoseconds, exactly Segment::getTimeNSec(), reference segment.h:58-63) and one
// coefficient array per axis in INCREASING powers -- x, y, z, then yaw (4-D) or rx, ry, rz (6-D rotation vector).
// A maintainer with ROS available assigns these fields to the generated message classes one to one
// (INTEGRATION.md section E).  The dimension rules are the reference's: 3, 4 or 6 dimensions (3 or 4 for the 4D
// message), anything else fails and clears the message.
#ifndef MAV_TRAJECTORY_GENERATION_ROS_ROS_CONVERSIONS_H_
#define MAV_TRAJECTORY_GENERATION_ROS_ROS_CONVERSIONS_H_

#include <cstdint>
#include <vector>

#include "mav_trajectory_generation/trajectory.h"

namespace mav_planning_msgs {

struct PolynomialSegment {   // mav_planning_msgs/PolynomialSegment.msg
  int32_t num_coeffs = 0;
  uint64_t segment_time_ns = 0;  // ros::Duration segment_time
  std::vector<double> x, y, z, rx, ry, rz, yaw;
};
struct PolynomialSegment4D {  // mav_planning_msgs/PolynomialSegment4D.msg
  int32_t num_coeffs = 0;
  uint64_t segment_time_ns = 0;
  std::vector<double> x, y, z, yaw;
};
struct PolynomialTrajectory {
  std::vector<PolynomialSegment> segments;
};
struct PolynomialTrajectory4D {
  std::vector<PolynomialSegment4D> segments;
};

}  // namespace mav_planning_msgs

namespace mav_trajectory_generation {

namespace ros_detail {
inline std::vector<double> toStd(const Eigen::VectorXd& v) {
  std::vector<double> out(static_cast<size_t>(v.size()));
  for (Eigen::Index i = 0; i < v.size(); ++i) out[static_cast<size_t>(i)] = v[i];
  return out;
}
inline Eigen::VectorXd toEigen(const std::vector<double>& v) {
  Eigen::VectorXd out(static_cast<Eigen::Index>(v.size()));
  for (size_t i = 0; i < v.size(); ++i) out[static_cast<Eigen::Index>(i)] = v[i];
  return out;
}
}  // namespace ros_detail

// reference ros_conversions.cpp:25-68
inline bool trajectoryToPolynomialTrajectoryMsg(const Trajectory& trajectory, mav_planning_msgs::PolynomialTrajectory* msg) {
  CHECK_NOTNULL(msg)->segments.clear();
  Segment::Vector segments;
  trajectory.getSegments(&segments);
  msg->segments.reserve(segments.size());
  for (const Segment& segment : segments) {
    if (segment.D() != 3 && segment.D() != 4 && segment.D() != 6) {
      LOG(ERROR) << "Dimension of position segment has to be 3, 4 or 6, but is " << segment.D();
      msg->segments.clear();
      return false;
    }
    mav_planning_msgs::PolynomialSegment m;
    m.x = ros_detail::toStd(segment[0].getCoefficients());
    m.y = ros_detail::toStd(segment[1].getCoefficients());
    m.z = ros_detail::toStd(segment[2].getCoefficients());
    if (segment.D() == 4) {
      m.yaw = ros_detail::toStd(segment[3].getCoefficients());
    } else if (segment.D() == 6) {
      m.rx = ros_detail::toStd(segment[3].getCoefficients());
      m.ry = ros_detail::toStd(segment[4].getCoefficients());
      m.rz = ros_detail::toStd(segment[5].getCoefficients());
    }
    m.num_coeffs = segment.N();
    m.segment_time_ns = segment.getTimeNSec();
    msg->segments.push_back(m);
  }
  return true;
}

// reference ros_conversions.cpp:70-106
inline bool polynomialTrajectoryMsgToTrajectory(const mav_planning_msgs::PolynomialTrajectory& msg, Trajectory* trajectory) {
  Segment::Vector segment_vector;
  for (const mav_planning_msgs::PolynomialSegment& m : msg.segments) {
    int D = 3;
    if (!m.yaw.empty()) D = 4;
    if (!m.rx.empty() && !m.ry.empty() && !m.rz.empty()) D = 6;
    Segment segment(static_cast<int>(m.x.size()), D);
    segment[0].setCoefficients(ros_detail::toEigen(m.x));
    segment[1].setCoefficients(ros_detail::toEigen(m.y));
    segment[2].setCoefficients(ros_detail::toEigen(m.z));
    if (D == 4) {
      segment[3].setCoefficients(ros_detail::toEigen(m.yaw));
    } else if (D == 6) {
      segment[3].setCoefficients(ros_detail::toEigen(m.rx));
      segment[4].setCoefficients(ros_detail::toEigen(m.ry));
      segment[5].setCoefficients(ros_detail::toEigen(m.rz));
    }
    segment.setTimeNSec(m.segment_time_ns);
    segment_vector.push_back(segment);
  }
  CHECK_NOTNULL(trajectory)->setSegments(segment_vector);
  return true;
}

// reference ros_conversions.cpp:108-150
inline bool trajectoryToPolynomialTrajectoryMsg(const Trajectory& trajectory, mav_planning_msgs::PolynomialTrajectory4D* msg) {
  CHECK_NOTNULL(msg)->segments.clear();
  Segment::Vector segments;
  trajectory.getSegments(&segments);
  msg->segments.reserve(segments.size());
  for (const Segment& segment : segments) {
    if (segment.D() != 3 && segment.D() != 4) {
      LOG(ERROR) << "Dimension of position segment has to be 3 or 4, but is " << segment.D();
      msg->segments.clear();
      return false;
    }
    mav_planning_msgs::PolynomialSegment4D m;
    m.x = ros_detail::toStd(segment[0].getCoefficients());
    m.y = ros_detail::toStd(segment[1].getCoefficients());
    m.z = ros_detail::toStd(segment[2].getCoefficients());
    if (segment.D() == 4) m.yaw = ros_detail::toStd(segment[3].getCoefficients());
    m.num_coeffs = segment.N();
    m.segment_time_ns = segment.getTimeNSec();
    msg->segments.push_back(m);
  }
  return true;
}

// reference ros_conversions.cpp:152-178
inline bool polynomialTrajectoryMsgToTrajectory(const mav_planning_msgs::PolynomialTrajectory4D& msg, Trajectory* trajectory) {
  Segment::Vector segment_vector;
  for (const mav_planning_msgs::PolynomialSegment4D& m : msg.segments) {
    const int D = m.yaw.empty() ? 3 : 4;
    Segment segment(static_cast<int>(m.x.size()), D);
    segment[0].setCoefficients(ros_detail::toEigen(m.x));
    segment[1].setCoefficients(ros_detail::toEigen(m.y));
    segment[2].setCoefficients(ros_detail::toEigen(m.z));
    if (D == 4) segment[3].setCoefficients(ros_detail::toEigen(m.yaw));
    segment.setTimeNSec(m.segment_time_ns);
    segment_vector.push_back(segment);
  }
  CHECK_NOTNULL(trajectory)->setSegments(segment_vector);
  return true;
}

}  // namespace mav_trajectory_generation
#endif  // MAV_TRAJECTORY_GENERATION_ROS_ROS_CONVERSIONS_H_
