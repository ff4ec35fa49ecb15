/* oracle shim (test infrastructure): included by main.cpp:21, not used */
#ifndef EKF_ORACLE_FLOAT32_SHIM_H
#define EKF_ORACLE_FLOAT32_SHIM_H
namespace std_msgs { struct Float32 { float data; }; }
#endif
