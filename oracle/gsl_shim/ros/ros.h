/* oracle shim (test infrastructure): Robot.cpp includes <ros/ros.h> but uses nothing from it. */
#ifndef EKF_ORACLE_ROS_SHIM_H
#define EKF_ORACLE_ROS_SHIM_H
#endif
