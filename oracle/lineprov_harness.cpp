/* oracle/lineprov_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's own end-point transform: lineprovider/main.cpp is compiled unmodified (oracle/Makefile, target
 * _ref/libslamlineprov.so, over oracle/ros_shim; its main() renamed and never called) and its free function Transform()
 * (lineprovider/main.cpp:60-84) is driven through the globals it works on: `lines` (only lineInterval is read), `theta`,
 * `pose` in, `lineIntervals.data` out (four floats per line: the `lines_1` topic).  The parity test of
 * ekf_lx_world_segments compares against this.
 */
#include <vector>

#include "std_msgs/Float32MultiArray.h"
#include "LineXtraction.h"          /* lineprovider's: line, polar_point, Vec2 */

extern std::vector<line> lines;                       /* lineprovider/main.cpp:16 */
extern std_msgs::Float32MultiArray lineIntervals;     /* :17 */
extern double theta;                                  /* :18 */
extern Vec2 pose;                                     /* :19 */
void Transform();                                     /* :60 */

extern "C" {

/* iv: n x (alfa0, r0, alfa1, r1) -- the two polar end points LineExtraction leaves in line::lineInterval;
 * out: 4 n floats (x0, y0, x1, y1 in the world frame).  Returns the number of floats Transform() produced. */
int lp_transform(int n, const double* iv, const double pose3[3], float* out) {
  lines.clear();
  lines.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    polar_point p0, p1;                               /* fields set one by one: the (alfa, r) constructor scales alfa */
    p0.alfa = iv[4 * i]; p0.r = iv[4 * i + 1]; p1.alfa = iv[4 * i + 2]; p1.r = iv[4 * i + 3];
    lines[i].lineInterval.clear();
    lines[i].lineInterval.push_back(p0); lines[i].lineInterval.push_back(p1);
  }
  pose.x = pose3[0]; pose.y = pose3[1]; theta = pose3[2];
  lineIntervals.data.clear();
  Transform();
  const int k = (int)lineIntervals.data.size();
  for (int i = 0; i < k; ++i) out[i] = lineIntervals.data[i];
  lineIntervals.data.clear();
  lines.clear();
  return k;
}

}  /* extern "C" */
