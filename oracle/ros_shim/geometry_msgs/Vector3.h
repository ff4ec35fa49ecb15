/* oracle shim (test infrastructure): geometry_msgs/Vector3 as main.cpp:79-89, 150-152 uses it */
#ifndef EKF_ORACLE_VECTOR3_SHIM_H
#define EKF_ORACLE_VECTOR3_SHIM_H
namespace geometry_msgs {
struct Vector3 { double x, y, z; Vector3() : x(0), y(0), z(0) {} };
}
#endif
