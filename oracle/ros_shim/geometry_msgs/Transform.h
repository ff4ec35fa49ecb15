/* oracle shim (test infrastructure): geometry_msgs/Transform as main.cpp:150-169 uses it (translation.{x,y,z}, rotation.{x,y,z}) */
#ifndef EKF_ORACLE_TRANSFORM_SHIM_H
#define EKF_ORACLE_TRANSFORM_SHIM_H
#include "geometry_msgs/Vector3.h"
namespace geometry_msgs {
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(0) {} };
struct Transform { Vector3 translation; Quaternion rotation; };
}
#endif
