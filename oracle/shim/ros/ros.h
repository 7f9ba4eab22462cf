// oracle/shim/ros/ros.h — TEST INFRASTRUCTURE ONLY.  The reference's filter uses ROS only for its
// logging macros (vslamRansac.cpp:4,518,784,837,1063).
#ifndef EKF_SHIM_ROS_H_
#define EKF_SHIM_ROS_H_
#include "../shim_prelude.h"
extern "C" void ekf_shim_log(int level, const char* fmt, ...);
#define ROS_ERROR(...) ekf_shim_log(2, __VA_ARGS__)
#define ROS_INFO(...) ekf_shim_log(1, __VA_ARGS__)
#define ROS_DEBUG(...) ekf_shim_log(0, __VA_ARGS__)
#include "../ros_msgs_shim.h"   // ros::Time and the message structs RosVSLAMRansac.hpp names
#include "../shim_retype.h"
#endif
