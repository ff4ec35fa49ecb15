/* The node's callback + loop in C++ over libekfcuda (slam_ros/main.cpp:37-71 and :139-174): payloads from a flat
 * binary file -> ekfcuda::LineExtractor -> ekfcuda::Robot::localize.  Dumps the lines of every scan and the pose.
 * Built and run by tests/test_gpu_cpp_dropin.py. */
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <vector>
#include "ekf_robot.hpp"

struct Mat { double data[4]; };
struct PolarPoint { double alfa, r; };
struct Line { double alfa, r; Mat* C_AR; std::vector<PolarPoint> lineInterval; };

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* fi = std::fopen(argv[1], "rb");
  if (!fi) return 3;
  int steps = 0, beams = 0;
  if (std::fread(&steps, sizeof(int), 1, fi) != 1 || std::fread(&beams, sizeof(int), 1, fi) != 1) return 4;
  std::vector<float> scans(2 * (size_t)steps * beams);
  std::vector<double> u(3 * (size_t)steps);
  if (std::fread(scans.data(), sizeof(float), scans.size(), fi) != scans.size()) return 4;
  if (std::fread(u.data(), sizeof(double), u.size(), fi) != u.size()) return 4;
  std::fclose(fi);
  ekfcuda::LineExtractor extractor;
  ekfcuda::Robot rover(0, 0, 0);
  FILE* fo = std::fopen(argv[2], "wb");
  if (!fo) return 5;
  for (int s = 0; s < steps; ++s) {
    std::deque<Mat> pool;                                    /* stands in for gsl_matrix_alloc(2, 2) */
    std::vector<Line> lines;
    const int n = extractor.extract(&scans[2 * (size_t)s * beams], 2 * beams, lines, [&pool]() { pool.push_back(Mat()); return &pool.back(); });
    const double nn = (double)n;
    std::fwrite(&nn, sizeof(double), 1, fo);
    for (int i = 0; i < n; ++i) {
      const double rec[10] = {lines[i].alfa, lines[i].r, lines[i].C_AR->data[0], lines[i].C_AR->data[1], lines[i].C_AR->data[2],
                              lines[i].C_AR->data[3], lines[i].lineInterval[0].alfa, lines[i].lineInterval[0].r,
                              lines[i].lineInterval[1].alfa, lines[i].lineInterval[1].r};
      std::fwrite(rec, sizeof(double), 10, fo);
    }
    if (lines.size() > 9) lines.resize(9);                   /* the reference cannot append more than 9 lines per scan (Q4) */
    const double enc[3] = {rover.xPos - u[3 * s], rover.yPos, rover.thetaPos - u[3 * s + 2]};
    rover.localize(lines, (float*)0, enc);
    const double pose[4] = {rover.xPos, rover.yPos, rover.thetaPos, (double)rover.savedLineCount};
    std::fwrite(pose, sizeof(double), 4, fo);
    /* the consumer after the filter (lineprovider/main.cpp:60-84): every extracted line's end points in the world frame */
    std::vector<float> seg;
    const int ns = extractor.worldSegments(rover.xPos, rover.yPos, rover.thetaPos, seg);
    if (ns != n || seg.size() != 4 * (size_t)n) return 6;
    for (size_t i = 0; i < seg.size(); ++i) { const double v = seg[i]; std::fwrite(&v, sizeof(double), 1, fo); }
  }
  std::fclose(fo);
  return 0;
}
