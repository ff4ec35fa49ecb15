/* ref_lines_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * C-ABI entry around the reference's own line extraction, `LineExtraction(vector<polar_point>&)`
 * (slam_ros/lineFitting.cpp:640-702), preceded and followed by exactly what the node's callback does around it
 * (slam_ros/main.cpp:37-71 with SIMULATIONOFF: alfa = angle - pi, r kept if > 0.05, variance 0.01; afterwards
 * alfa += pi wrapped into (-pi, pi]).  The reference sources are compiled where they lie by oracle/Makefile into
 * oracle/_ref/libslamlines.so; nothing is copied here.
 *
 * The reference reads indeterminate memory on this path: residual_error accumulates into
 * an uninitialised `sum_dist` (lineFitting.cpp:109,122) and Covariancia multiplies by a gsl_matrix_alloc'ed C_x of
 * which only the diagonal is written (:382, :416-420).  To get ONE deterministic instance of the reference this
 * library is built with -ftrivial-auto-var-init=zero (every automatic variable starts at zero) and with the
 * shim's gsl_matrix_alloc zero-filling (-DGSL_SHIM_ZERO_ALLOC): both are behaviours the reference may legally
 * exhibit, and they are the ones its authors evidently intended.
 *
 * WriteCov (lineFitting.cpp:53-59) prints rejected covariances with printf and cout; both are silenced for the call.
 * LineExtraction also writes dist.txt / p_data.txt / p_raw_data.txt into the working directory on every call;
 * the harness runs it inside a scratch directory.
 */
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include <cmath>
#include <unistd.h>
#include <fcntl.h>
#include <cstdio>
#include "LineXtraction.h"

extern "C" {

/* data: n_pairs x (r, angle) as float32, the `mappingPoints` payload (main.cpp:46-56).
 * out:  per line 10 doubles: alfa, r, C_AR[0..3], interval0 (alfa, r), interval1 (alfa, r).
 * Returns the number of lines (<= max_lines are written), or -1. */
int ref_extract_lines(int n_pairs, const float* data, int max_lines, double* out) {
  static char scratch[256] = {0};
  char cwd[1024];
  if (!getcwd(cwd, sizeof cwd)) return -1;
  if (!scratch[0]) {
    std::snprintf(scratch, sizeof scratch, "/tmp/slamref_lines_XXXXXX");
    if (!mkdtemp(scratch)) return -1;
  }
  if (chdir(scratch) != 0) return -1;
  std::ios_base::iostate old = std::cout.rdstate();
  std::cout.setstate(std::ios_base::failbit);                /* lineFitting prints from WriteCov: cout ... */
  std::fflush(stdout);                                       /* ... and printf: park fd 1 on /dev/null for the call */
  const int saved_fd = dup(1), null_fd = open("/dev/null", O_WRONLY);
  if (saved_fd >= 0 && null_fd >= 0) dup2(null_fd, 1);
  std::vector<polar_point> points;
  polar_point temp;
  for (int i = 0; i < 2 * n_pairs; i += 2) {                 /* main.cpp:46-62 */
    if (data[i] > 0.05) {
      points.push_back(temp);
      points[points.size() - 1].alfa = data[i + 1] - M_PI;
      points[points.size() - 1].r = data[i];
      points[points.size() - 1].variance = 0.01;
    }
  }
  std::vector<line> lines;
  if (points.size() >= 2) lines = LineExtraction(points);    /* main.cpp:65 */
  for (auto& lin : lines) {                                  /* main.cpp:66-69 */
    lin.alfa += M_PI;
    lin.alfa = lin.alfa > M_PI ? lin.alfa - 2.0 * M_PI : lin.alfa;
  }
  std::fflush(stdout);
  if (saved_fd >= 0) { dup2(saved_fd, 1); close(saved_fd); }
  if (null_fd >= 0) close(null_fd);
  std::cout.clear(old);
  if (chdir(cwd) != 0) return -1;
  const int n = (int)lines.size();
  for (int i = 0; i < n && i < max_lines; ++i) {
    double* o = out + 10 * i;
    o[0] = lines[i].alfa; o[1] = lines[i].r;
    for (int k = 0; k < 4; ++k) o[2 + k] = lines[i].C_AR ? lines[i].C_AR->data[k] : 0.0;
    for (int k = 0; k < 2; ++k) {
      const bool have = lines[i].lineInterval.size() > (size_t)k;
      o[6 + 2 * k] = have ? lines[i].lineInterval[k].alfa : 0.0;
      o[7 + 2 * k] = have ? lines[i].lineInterval[k].r : 0.0;
    }
  }
  return n;
}

}  /* extern "C" */
