#ifndef CVSHIM_ROS_PACKAGE_H
#define CVSHIM_ROS_PACKAGE_H
#include "ros.h"
namespace ros {
namespace package {
inline std::string getPath(const std::string&) { return ros::shim::package_path(); }
} // namespace package
} // namespace ros
#endif
