/* Drives include/ekf_robot.hpp (the C++ drop-in `Robot`) exactly like slam_ros/main.cpp:139-147 drives the
 * reference's Robot::localize, on a scan sequence read from a flat binary file, and dumps the results.
 * Built and run by tests/test_gpu_cpp_dropin.py. */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ekf_robot.hpp"

/* same members as the reference's `line` / `polar_point` / gsl_matrix (simplifyPath.h:48-79) */
struct Mat { double* data; };
struct PolarPoint { double alfa, r; };
struct Line { double alfa, r; Mat* C_AR; std::vector<PolarPoint> lineInterval; };

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  if (!fi) return 3;
  int steps = 0, maxl = 0;
  if (std::fread(&steps, sizeof(int), 1, fi) != 1 || std::fread(&maxl, sizeof(int), 1, fi) != 1) return 4;
  std::vector<int> count(steps);
  std::vector<double> u(3 * (size_t)steps), z(2 * (size_t)steps * maxl), R(4 * (size_t)steps * maxl);
  if (std::fread(count.data(), sizeof(int), steps, fi) != (size_t)steps) return 4;
  if (std::fread(u.data(), sizeof(double), u.size(), fi) != u.size()) return 4;
  if (std::fread(z.data(), sizeof(double), z.size(), fi) != z.size()) return 4;
  if (std::fread(R.data(), sizeof(double), R.size(), fi) != R.size()) return 4;
  std::fclose(fi);
  ekfcuda::Robot rover(0, 0, 0);
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fo) return 5;
  std::vector<double> Rstore(4 * (size_t)maxl);
  std::vector<Mat> mats(maxl);
  for (int s = 0; s < steps; ++s) {
    const int m = count[s];
    std::vector<Line> lines(m);
    for (int i = 0; i < m; ++i) {
      for (int t = 0; t < 4; ++t) Rstore[4 * i + t] = R[4 * ((size_t)s * maxl + i) + t];
      mats[i].data = &Rstore[4 * i];
      lines[i].alfa = z[2 * ((size_t)s * maxl + i)];
      lines[i].r = z[2 * ((size_t)s * maxl + i) + 1];
      lines[i].C_AR = &mats[i];
      lines[i].lineInterval.push_back(PolarPoint{lines[i].alfa - 0.1, lines[i].r + 0.2});
      lines[i].lineInterval.push_back(PolarPoint{lines[i].alfa + 0.1, lines[i].r + 0.3});
    }
    /* the node feeds the external pose; here the one that makes (pose - encoder) the intended odometry (Q5) */
    const double enc[3] = {rover.xPos - u[3 * s], rover.yPos, rover.thetaPos - u[3 * s + 2]};
    rover.localize(lines, (float*)0, enc);
    const double rec[4] = {rover.xPos, rover.yPos, rover.thetaPos, (double)rover.savedLineCount};
    std::fwrite(rec, sizeof(double), 4, fo);
  }
  float ax[2], ang = 0.f;
  const bool ok = rover.getEllipse(ax, ang);
  const double ell[4] = {ok ? 1.0 : 0.0, ax[0], ax[1], ang};
  std::fwrite(ell, sizeof(double), 4, fo);
  rover.syncCovariance();
  const double nn = (double)rover.y.size();
  std::fwrite(&nn, sizeof(double), 1, fo);
  std::fwrite(rover.y.data(), sizeof(double), rover.y.size(), fo);
  std::fwrite(rover.P_t0.data(), sizeof(double), rover.P_t0.size(), fo);
  const double nli = (double)rover.lineIntervals.data.size();
  std::fwrite(&nli, sizeof(double), 1, fo);
  std::fclose(fo);
  return 0;
}
