/* Robot_cuda.cpp -- the translation unit a maintainer of slam_ros puts in place of slam_ros/Robot.cpp.
 *
 * It is compiled against the reference's OWN, UNMODIFIED headers (slam_ros/Robot.h:21-77, simplifyPath.h:62-79,
 * lineFitting.h) and defines the members of `class Robot` the node uses (slam_ros/main.cpp:98, 144-174):
 *
 *     Robot::Robot, ~Robot            Robot.cpp:20-39      state created in HBM (ekf_create)
 *     Robot::localize                 Robot.cpp:126-943    one ekf_scan: predict, associate, update, augment, reset
 *     Robot::getEllipse               Robot.cpp:73-124     ekf_get_ellipse
 *     Robot::normalizeRadian          Robot.cpp:62-71      unchanged arithmetic
 *     Robot::measure                  Robot.cpp:944-947    (an empty stub in the reference, never called by the node)
 *
 * Robot.h cannot grow a member, so the ekf_ctx lives in a side table keyed by the object's address.
 *
 * ONE behavioural difference: the covariance stays resident in HBM, so the public host array Robot::P_t0 (and the
 * private y[]) are not rewritten by every localize.  They are brought up to date ON DEMAND by Robot::measure() --
 * the reference declares and defines it as an empty function and nothing calls it, which makes it the one spot of the
 * unmodified class where "bring the host mirrors up to date" fits.  main.cpp reads neither array (its only
 * covariance consumer is getEllipse), so the node itself never needs the call.
 *
 * Built and driven by the same harness as the literal reference: oracle/Makefile target _ref/libslamdropin.so,
 * tests/test_gpu_dropin_ref_headers.py.
 */
#include <cstring>
#include <map>

#include "Robot.h"          /* the reference's */
#include "ekf_robot.hpp"    /* this repository: ekfcuda::Robot over the C ABI of include/ekf.h */

namespace {
typedef ekfcuda::Robot Impl;
std::map<const Robot*, Impl*>& impls() { static std::map<const Robot*, Impl*> table; return table; }
Impl* impl_of(const Robot* r) {
  std::map<const Robot*, Impl*>::iterator it = impls().find(r);
  return it == impls().end() ? 0 : it->second;
}
}  // namespace

Robot::Robot(double x, double y0, double theta) {                       /* Robot.cpp:20-35 */
  this->xPos = x; this->yPos = y0; this->thetaPos = theta;
  this->savedLineCount = 0; this->matchesNum = 0;                       /* the reference leaves these unset */
  std::memset(this->y, 0, sizeof this->y);
  std::memset(this->P_t0, 0, sizeof this->P_t0);
  this->P_t0[0] = 0.05; this->P_t0[SLAMSIZE + 1] = 0.05;                /* Robot.cpp:28-30 */
  impls()[this] = new Impl(x, y0, theta, LINESIZE);
  lineIntervals.data.reserve(80);
}

Robot::~Robot() {
  delete impl_of(this);
  impls().erase(this);
}

void Robot::normalizeRadian(double& rad) {                              /* Robot.cpp:62-71 */
  if (rad > M_PI) rad = rad - (2.0 * M_PI + floor(rad / (2.0 * M_PI)) * 2.0 * M_PI);
  else if (rad < -M_PI) rad = rad + (2.0 * M_PI + floor(std::abs(rad) / (2.0 * M_PI)) * 2.0 * M_PI);
}

bool Robot::getEllipse(float axii[], float& angle) { return impl_of(this)->getEllipse(axii, angle); }   /* Robot.cpp:73-124 */

void Robot::localize(const std::vector<line>& lines, float* rot, const double* encoder) {                /* Robot.cpp:126-943 */
  Impl* ekf = impl_of(this);
  ekf->xPos = xPos; ekf->yPos = yPos; ekf->thetaPos = thetaPos;         /* the node may have overwritten the public pose */
  ekf->localize(lines, rot, encoder);
  xPos = ekf->xPos; yPos = ekf->yPos; thetaPos = ekf->thetaPos;         /* read by main.cpp:150-152 */
  savedLineCount = ekf->savedLineCount;
  lineIntervals.data.insert(lineIntervals.data.end(), ekf->lineIntervals.data.begin(), ekf->lineIntervals.data.end());
  ekf->lineIntervals.data.clear();                                      /* main.cpp:172-174 publishes, then clears Robot::lineIntervals */
}

void Robot::measure() {                                                 /* host mirrors on demand (see the header comment) */
  Impl* ekf = impl_of(this);
  ekf->syncCovariance();
  std::memcpy(this->y, ekf->y.data(), sizeof this->y);
  std::memcpy(this->P_t0, ekf->P_t0.data(), sizeof this->P_t0);
  savedLineCount = ekf->savedLineCount;
}
