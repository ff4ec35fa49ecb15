/* oracle shim (test infrastructure): the two fields of std_msgs::Float32MultiArray that
 * Robot.h:59 / Robot.cpp:31,873-878 touch (data) and main.cpp:41 reads (layout.dim[].size). */
#ifndef EKF_ORACLE_F32MA_SHIM_H
#define EKF_ORACLE_F32MA_SHIM_H
#include <vector>
#include <string>
#include <random>
namespace std_msgs {
struct MultiArrayDimension { std::string label; unsigned int size; unsigned int stride; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned int data_offset; };
struct Float32MultiArray { MultiArrayLayout layout; std::vector<float> data; };
}
#endif
