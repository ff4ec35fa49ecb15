/* force-included (-include) ahead of slam_ros/main.cpp in the node builds (test infrastructure):
 *   - main() becomes slam_node_main(), called by oracle/node_harness.cpp;
 *   - the node's one-second start-up pause (main.cpp:125, "the nodes require time to connect") is not slept. */
#ifndef EKF_ORACLE_NODE_PRE_H
#define EKF_ORACLE_NODE_PRE_H
#include <unistd.h>
#define usleep(x) ((void)(x))
#define main slam_node_main
int slam_node_main(int argc, char* argv[]);
#endif
