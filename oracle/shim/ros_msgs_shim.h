// oracle/shim/ros_msgs_shim.h — TEST INFRASTRUCTURE ONLY.  Plain-struct stand-ins for the ROS message types that
// mono-slam/src/RosVSLAMRansac.{hpp,cpp} names (geometry_msgs, nav_msgs, visualization_msgs, ros::Time), written from the
// public message definitions, so that RosVSLAM::getPointsFeatures can be compiled from the reference's own source into
// oracle/_ref.  Only the fields the reference touches exist; nothing is published anywhere.
#ifndef EKF_SHIM_ROS_MSGS_H_
#define EKF_SHIM_ROS_MSGS_H_
#include <string>
#include <vector>
#ifdef float   // message fields keep their real types whatever the scalar re-typing of the build
#undef float
#define EKF_SHIM_MSGS_RESTORE_FLOAT 1
#endif
namespace ros {
struct Time {
  double sec = 0;
  static Time now() { return Time(); }
  double toSec() const { return sec; }
};
struct Duration {
  double sec = 0;
  Duration() {}
  Duration(double s) : sec(s) {}
};
}  // namespace ros
namespace std_msgs {
struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}  // namespace std_msgs
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseArray { std_msgs::Header header; std::vector<Pose> poses; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36] = {0}; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
}  // namespace geometry_msgs
namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; };
}  // namespace nav_msgs
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8,
         TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns, text, mesh_resource;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
  std::vector<std_msgs::ColorRGBA> colors;
  ros::Duration lifetime;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs
#ifdef EKF_SHIM_MSGS_RESTORE_FLOAT
#undef EKF_SHIM_MSGS_RESTORE_FLOAT
#include "shim_retype.h"
#endif
#endif
