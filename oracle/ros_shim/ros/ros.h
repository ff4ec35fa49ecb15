/* oracle/ros_shim/ros/ros.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A restatement of the sliver of the roscpp API that the reference node, slam_ros/main.cpp, uses (ros::init, NodeHandle::
 * advertise / subscribe, Publisher::publish, Subscriber::shutdown, Rate, ok, spinOnce, shutdown), as an in-process bus with
 * no network and no threads, so that the reference's OWN main.cpp can be compiled, unmodified, where it lies and driven by a
 * test (oracle/node_harness.cpp):
 *
 *   - ros::spinOnce() asks the harness for the next tick's messages and delivers them to the subscribed callbacks, in
 *     the order the callbacks were subscribed;
 *   - ros::ok() turns false when the harness has no further tick;
 *   - Publisher::publish records what the node publishes; the harness hands it to the test.
 *
 * ROS itself (Kinetic, by the reference's package.xml) is not in this image; nothing here is copied from it.  The message
 * structs carry only the fields main.cpp touches (main.cpp:37-89, 150-173).
 */
#ifndef EKF_ORACLE_ROS_FULL_SHIM_H
#define EKF_ORACLE_ROS_FULL_SHIM_H
#include <functional>
#include <string>
#include <vector>

#include "geometry_msgs/Transform.h"
#include "geometry_msgs/Vector3.h"
#include "std_msgs/Float32MultiArray.h"

namespace ros {

struct Bus {
  struct Sub { std::string topic; std::function<void(const void*)> deliver; };
  std::vector<Sub> subs;
  /* the harness: fills the tick's messages, returns false when the script is over */
  std::function<bool(int tick, geometry_msgs::Vector3* real_pose, std_msgs::Float32MultiArray* points, bool* has_points,
                     geometry_msgs::Vector3* encoder, bool* has_encoder)> next;
  int tick;
  bool running;
  std::vector<geometry_msgs::Transform> robot_position;       /* topic "robotPosition" */
  std::vector<std::vector<float> > lines;                     /* topic "lines": one entry per publish */
  Bus() : tick(0), running(true) {}
};
inline Bus& bus() { static Bus b; return b; }

inline void init(int&, char**, const std::string&) {}
inline bool ok() { return bus().running; }
inline void shutdown() { bus().running = false; }

inline void deliver(const std::string& topic, const void* msg) {
  Bus& b = bus();
  for (size_t i = 0; i < b.subs.size(); ++i) if (b.subs[i].topic == topic) b.subs[i].deliver(msg);
}
inline void spinOnce() {
  Bus& b = bus();
  if (!b.running) return;
  geometry_msgs::Vector3 real_pose, encoder;
  std_msgs::Float32MultiArray points;
  bool has_points = false, has_encoder = false;
  if (!b.next || !b.next(b.tick, &real_pose, &points, &has_points, &encoder, &has_encoder)) { b.running = false; return; }
  b.tick += 1;
  if (has_encoder) deliver("encoderPosition", &encoder);
  deliver("realRoboPose", &real_pose);
  if (has_points) deliver("mappingPoints", &points);
}

struct Rate { explicit Rate(double) {} void sleep() {} };
struct Subscriber { void shutdown() {} };
struct Publisher {
  std::string topic;
  void publish(const geometry_msgs::Transform& m) const { bus().robot_position.push_back(m); }
  void publish(const std_msgs::Float32MultiArray& m) const { bus().lines.push_back(m.data); }
};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string& topic, int) { Publisher p; p.topic = topic; return p; }
  template <class M> Subscriber subscribe(const std::string& topic, int, void (*cb)(M)) {
    Bus::Sub s;
    s.topic = topic;
    s.deliver = [cb](const void* m) { cb(*static_cast<const M*>(m)); };      /* main.cpp's callbacks take the message by value */
    bus().subs.push_back(s);
    return Subscriber();
  }
};

}  // namespace ros
#endif
