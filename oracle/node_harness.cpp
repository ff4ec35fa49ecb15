/* oracle/node_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the reference NODE itself: slam_ros/main.cpp, compiled unmodified where it lies under /root/reference over the
 * in-process roscpp restatement of oracle/ros_shim/ (its main() renamed by a force-included header).  Two builds share this
 * file (oracle/Makefile):
 *
 *   _ref/libslamnode_ref.so      main.cpp + the reference's Robot.cpp (Q1 patched as everywhere), lineFitting.cpp, ...
 *   _ref/libslamnode_dropin.so   main.cpp + dropin/Robot_cuda.cpp over libekfcuda.so (the drop-in), same other sources
 *
 * A test supplies the messages tick by tick through a callback (it sees what the node published so far, so it can close
 * the loop the way the simulator does) and reads back every `robotPosition` and `lines` message the node published
 * (main.cpp:150-174).
 */
#include <cstdlib>
#include <iostream>
#include <new>
#include <vector>

#include "ros/ros.h"

/* Robot's constructor leaves y[], savedLineCount and most of P_t0 unset; main.cpp:98 does `new Robot(0, 0, 0)` and
 * works because a fresh process gets that 330 KB block as zero pages from the kernel.  Inside a long-lived test process the
 * allocator hands out recycled memory instead, so the node libraries (linked -Bsymbolic: only they see this) allocate
 * zero-filled storage -- the behaviour the node has when launched the way it is meant to be. */
void* operator new(std::size_t n) {
  void* p = std::calloc(1, n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](std::size_t n) {
  void* p = std::calloc(1, n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

int slam_node_main(int argc, char* argv[]);

extern "C" {

/* fills real_pose[3] (topic realRoboPose) and *points / *n_floats (topic mappingPoints: (r, angle) float pairs, 0 floats =
 * no scan this tick); last_pose6 = the latest robotPosition message (translation xyz, rotation xyz) or NULL.
 * Returns 0 when the script is over. */
typedef int (*node_tick_fn)(int tick, const double* last_pose6, int n_published, double* real_pose, const float** points, int* n_floats);

int slam_node_run(node_tick_fn fn, int max_ticks) {
  std::cout.setstate(std::ios_base::failbit);        /* the node prints ~100 lines per scan */
  ros::Bus& b = ros::bus();
  b.next = [fn, max_ticks](int tick, geometry_msgs::Vector3* real_pose, std_msgs::Float32MultiArray* points, bool* has_points,
                           geometry_msgs::Vector3*, bool* has_encoder) -> bool {
    if (tick >= max_ticks) return false;
    ros::Bus& bb = ros::bus();
    double last[6] = {0, 0, 0, 0, 0, 0};
    const bool have = !bb.robot_position.empty();
    if (have) {
      const geometry_msgs::Transform& t = bb.robot_position.back();
      last[0] = t.translation.x; last[1] = t.translation.y; last[2] = t.translation.z;
      last[3] = t.rotation.x; last[4] = t.rotation.y; last[5] = t.rotation.z;
    }
    double rp[3] = {0, 0, 0};
    const float* pts = 0;
    int n = 0;
    if (!fn(tick, have ? last : 0, (int)bb.robot_position.size(), rp, &pts, &n)) return false;
    real_pose->x = rp[0]; real_pose->y = rp[1]; real_pose->z = rp[2];
    *has_encoder = false;
    *has_points = n > 0 && pts;
    if (*has_points) {
      points->layout.dim.resize(1);
      points->layout.dim[0].size = (unsigned)n;          /* main.cpp:41, 46 */
      points->layout.dim[0].stride = 1;
      points->data.assign(pts, pts + n);
    }
    return true;
  };
  const int rc = slam_node_main(0, 0);
  std::cout.clear();
  return rc;
}

int slam_node_published(void) { return (int)ros::bus().robot_position.size(); }

int slam_node_get_pose(int i, double out6[6]) {
  ros::Bus& b = ros::bus();
  if (i < 0 || i >= (int)b.robot_position.size()) return -1;
  const geometry_msgs::Transform& t = b.robot_position[i];
  out6[0] = t.translation.x; out6[1] = t.translation.y; out6[2] = t.translation.z;
  out6[3] = t.rotation.x; out6[4] = t.rotation.y; out6[5] = t.rotation.z;
  return 0;
}

/* the i-th `lines` message: returns its length in floats (4 per appended line), copies up to cap floats */
int slam_node_get_lines(int i, float* out, int cap) {
  ros::Bus& b = ros::bus();
  if (i < 0 || i >= (int)b.lines.size()) return -1;
  const std::vector<float>& v = b.lines[i];
  for (int k = 0; k < (int)v.size() && k < cap; ++k) out[k] = v[k];
  return (int)v.size();
}

}  /* extern "C" */
