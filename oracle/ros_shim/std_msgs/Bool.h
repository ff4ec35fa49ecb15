/* oracle shim (test infrastructure): included by main.cpp:20, not used */
#ifndef EKF_ORACLE_BOOL_SHIM_H
#define EKF_ORACLE_BOOL_SHIM_H
namespace std_msgs { struct Bool { bool data; }; }
#endif
