"""CPU tests of the checkers themselves (no GPU): the structured oracle is pinned, bitwise, to the
reference's own Robot.cpp compiled over the GSL shim, and to the committed golden vectors that the
reference produced (tests/golden/make_golden.py)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle.oracle import LiteralReference, StructuredOracle, have_literal
from slam_ros_b200 import scenario as sc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "room_literal.npz")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_vectors_pin_the_structured_oracle():
    g = np.load(GOLD)
    so = StructuredOracle(100)
    steps = g["u"].shape[0]
    for s in range(steps):
        m = int(g["count"][s])
        st, j = so.localize(g["z"][s, :m], g["R"][s, :m], g["encoder"][s])
        assert so.lines == int(g["L"][s]), "line count at step %d" % s
        # libm (sin/cos) may differ in the last ulp between hosts: tight tolerance instead of bit equality
        assert np.abs(so.pose - g["pose"][s]).max() < 1e-11, "pose at step %d" % s
        if ("P_%d" % s) in g.files:
            P = so.P_full(); y = so.y_full()
            assert np.abs(P - g["P_%d" % s]).max() <= 1e-12 * np.abs(P).max(), "P at step %d" % s
            assert np.abs(y - g["y_%d" % s]).max() <= 1e-11, "y at step %d" % s
        assert abs(np.trace(so.P_view()) - g["trace"][s]) <= 1e-10 * abs(g["trace"][s])
    assert so.stats()["resets"] > 0 and so.stats()["matches"] > 100
    ok, ax, ang = so.get_ellipse()
    assert ok and np.allclose(ax, g["ellipse"][:2], rtol=1e-5)


@pytest.mark.skipif(not have_literal(), reason="oracle/_ref/libslamref.so not built (needs /root/reference)")
def test_structured_oracle_is_bitwise_the_literal_reference():
    room = sc.room_scenario(steps=1000, seed=7, range_sigma=5e-5)
    lit = LiteralReference(); so = StructuredOracle(lit.capacity, lit.gate, lit.encoder_noise)
    assert lit.capacity == 100 and lit.n == 203
    for s in range(1000):
        m = room["count"][s]
        y, P, L, pose = lit.state()
        enc = sc.encoder_for(pose, room["u"][s])
        lit.localize(room["z"][s, :m], room["R"][s, :m], enc)
        so.localize(room["z"][s, :m], room["R"][s, :m], enc)
        if s % 25 == 0 or s > 990:
            y, P, L, pose = lit.state()
            assert L == so.lines and np.array_equal(pose, so.pose), "step %d" % s
            assert np.array_equal(y, so.y_full()) and np.array_equal(P, so.P_full()), "step %d" % s
    st = so.stats()
    assert st["resets"] >= 1 and st["matches"] > 500
    print("literal == structured bitwise over 1000 steps; min gate margin %.3e over %d gates; %d shim range errors"
          % (st["min_margin"], st["gates"], lit.range_errors()))


@pytest.mark.skipif(not have_literal(), reason="oracle/_ref/libslamref.so not built")
def test_structured_oracle_bitwise_on_noisy_mismatching_scans():
    """Honest noise (the node's own 1 cm range sigma): few matches, duplicates, many resets."""
    room = sc.room_scenario(steps=250, seed=3, range_sigma=1e-2)
    lit = LiteralReference(); so = StructuredOracle(100)
    for s in range(250):
        m = room["count"][s]
        y, P, L, pose = lit.state()
        enc = sc.encoder_for(pose, room["u"][s])
        lit.localize(room["z"][s, :m], room["R"][s, :m], enc)
        so.localize(room["z"][s, :m], room["R"][s, :m], enc)
    y, P, L, pose = lit.state()
    assert L == so.lines and np.array_equal(y, so.y_full()) and np.array_equal(P, so.P_full())


@pytest.mark.skipif(not have_literal(), reason="oracle/_ref/libslamref.so not built")
def test_empty_scans_and_ellipse_against_literal():
    lit = LiteralReference(); so = StructuredOracle(100)
    u = np.array([0.05, 0.0, 0.03])
    for s in range(5):                      # no lines at all: Robot.cpp:702-716
        y, P, L, pose = lit.state()
        enc = sc.encoder_for(pose, u)
        lit.localize(np.zeros((0, 2)), np.zeros((0, 4)), enc)
        so.localize(np.zeros((0, 2)), np.zeros((0, 4)), enc)
    y, P, L, pose = lit.state()
    assert L == so.lines == 0 and np.array_equal(P, so.P_full()) and np.array_equal(pose, so.pose)
    okl, axl, angl = lit.get_ellipse()
    oko, axo, ango = so.get_ellipse()
    assert okl and oko and np.allclose(axl, axo, rtol=1e-6)


def test_oracle_scan_equals_localize_given_same_odometry():
    scn = sc.map_scenario(30, 20, m=5, seed=5)
    a = StructuredOracle(64); b = StructuredOracle(64)
    a.scan(np.zeros(3), scn["seed_z"], scn["seed_R"]); b.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
    for s in range(20):
        u = scn["u"][s]
        # encoder chosen so that (pose - encoder) reproduces u exactly in floating point
        st, ja = a.scan(u, scn["z"][s], scn["R"][s])
        pose = b.pose
        enc = np.array([pose[0] - u[0], pose[1], pose[2] - u[2]])
        st, jb = b.localize(scn["z"][s], scn["R"][s], enc)
        assert np.array_equal(ja, jb)
    assert np.abs(a.P_full() - b.P_full()).max() < 1e-9


def test_oracle_openmp_is_bitwise_thread_independent():
    scn = sc.map_scenario(120, 10, m=6, seed=2)
    a = StructuredOracle(160, threads=1); b = StructuredOracle(160, threads=4)
    for o in (a, b):
        o.scan(np.zeros(3), scn["seed_z"], scn["seed_R"])
        for s in range(10):
            o.scan(scn["u"][s], scn["z"][s], scn["R"][s])
    assert np.array_equal(a.P_full(), b.P_full()) and np.array_equal(a.y_full(), b.y_full())


KAT_SRC = r'''
#include "gsl/gsl_matrix.h"
#include "gsl/gsl_blas.h"
#include "gsl/gsl_linalg.h"
#include <cstdio>
/* generic EKF quiz of slam_ros/Robot.h:84-218 (inputs there; expected outputs derived in SURVEY.md section 4) */
int main() {
  double P[4] = {1, 0, 0, 1}, H[4] = {2, 1, 1, 2}, R[4] = {0.5, 0, 0, 0.5}, x[2] = {1, 2}, z[2] = {1.1, 1.9}, h[2] = {1, 2};
  gsl_matrix_view Pv = gsl_matrix_view_array(P, 2, 2), Hv = gsl_matrix_view_array(H, 2, 2), Rv = gsl_matrix_view_array(R, 2, 2);
  double HP[4], S[4], Sc[4], Si[4], PHt[4], K[4], KS[4], KSK[4];
  gsl_matrix_view HPv = gsl_matrix_view_array(HP, 2, 2), Sv = gsl_matrix_view_array(S, 2, 2), Scv = gsl_matrix_view_array(Sc, 2, 2),
                  Siv = gsl_matrix_view_array(Si, 2, 2), PHtv = gsl_matrix_view_array(PHt, 2, 2), Kv = gsl_matrix_view_array(K, 2, 2),
                  KSv = gsl_matrix_view_array(KS, 2, 2), KSKv = gsl_matrix_view_array(KSK, 2, 2);
  gsl_blas_dgemm(CblasNoTrans, CblasNoTrans, 1.0, &Hv.matrix, &Pv.matrix, 0.0, &HPv.matrix);
  gsl_blas_dgemm(CblasNoTrans, CblasTrans, 1.0, &HPv.matrix, &Hv.matrix, 0.0, &Sv.matrix);
  gsl_matrix_add(&Sv.matrix, &Rv.matrix);
  gsl_matrix_memcpy(&Scv.matrix, &Sv.matrix);
  gsl_permutation* p = gsl_permutation_alloc(2); int sg;
  gsl_linalg_LU_decomp(&Scv.matrix, p, &sg);
  gsl_linalg_LU_invert(&Scv.matrix, p, &Siv.matrix);
  gsl_blas_dgemm(CblasNoTrans, CblasTrans, 1.0, &Pv.matrix, &Hv.matrix, 0.0, &PHtv.matrix);
  gsl_blas_dgemm(CblasNoTrans, CblasNoTrans, 1.0, &PHtv.matrix, &Siv.matrix, 0.0, &Kv.matrix);
  gsl_blas_dgemm(CblasNoTrans, CblasNoTrans, 1.0, &Kv.matrix, &Sv.matrix, 0.0, &KSv.matrix);
  gsl_blas_dgemm(CblasNoTrans, CblasTrans, 1.0, &KSv.matrix, &Kv.matrix, 0.0, &KSKv.matrix);
  gsl_matrix_sub(&Pv.matrix, &KSKv.matrix);
  double v[2] = {z[0] - h[0], z[1] - h[1]};
  double xp[2] = {x[0] + K[0] * v[0] + K[1] * v[1], x[1] + K[2] * v[0] + K[3] * v[1]};
  std::printf("%.17g %.17g %.17g %.17g\n", S[0], S[1], S[2], S[3]);
  std::printf("%.17g %.17g %.17g %.17g\n", K[0], K[1], K[2], K[3]);
  std::printf("%.17g %.17g\n", xp[0], xp[1]);
  std::printf("%.17g %.17g %.17g %.17g\n", P[0], P[1], P[2], P[3]);
  { /* P_prior = Fx P Fx' + Fu Q Fu' with Fx = Fu = [2 1;1 2], P = I, Q = 0.5 I */
    double F[4] = {2, 1, 1, 2}, P0[4] = {1, 0, 0, 1}, Q[4] = {0.5, 0, 0, 0.5}, T[4], A[4], B[4];
    gsl_matrix_view Fv = gsl_matrix_view_array(F, 2, 2), P0v = gsl_matrix_view_array(P0, 2, 2), Qv = gsl_matrix_view_array(Q, 2, 2),
                    Tv = gsl_matrix_view_array(T, 2, 2), Av = gsl_matrix_view_array(A, 2, 2), Bv = gsl_matrix_view_array(B, 2, 2);
    gsl_blas_dgemm(CblasNoTrans, CblasNoTrans, 1.0, &Fv.matrix, &P0v.matrix, 0.0, &Tv.matrix);
    gsl_blas_dgemm(CblasNoTrans, CblasTrans, 1.0, &Tv.matrix, &Fv.matrix, 0.0, &Av.matrix);
    gsl_blas_dgemm(CblasNoTrans, CblasNoTrans, 1.0, &Fv.matrix, &Qv.matrix, 0.0, &Tv.matrix);
    gsl_blas_dgemm(CblasNoTrans, CblasTrans, 1.0, &Tv.matrix, &Fv.matrix, 0.0, &Bv.matrix);
    gsl_matrix_add(&Av.matrix, &Bv.matrix);
    std::printf("%.17g %.17g %.17g %.17g\n", A[0], A[1], A[2], A[3]);
  }
  /* shape mismatch must return GSL_EBADLEN (19) and a range error must read 0 */
  gsl_matrix_view bad = gsl_matrix_view_array(HP, 1, 4);
  std::printf("%d %g\n", gsl_matrix_add(&Sv.matrix, &bad.matrix), gsl_matrix_get(&Sv.matrix, 5, 0));
  return 0;
}
'''


def test_gsl_shim_known_answers():
    """The restated GSL subset reproduces the generic-EKF quiz of Robot.h:84-218 (values: SURVEY.md 4)."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "kat.cpp"); exe = os.path.join(d, "kat")
        open(src, "w").write(KAT_SRC)
        subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++11", "-O2",
                               "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle", "gsl_shim"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    S = np.array(out[0].split(), dtype=float); K = np.array(out[1].split(), dtype=float)
    xp = np.array(out[2].split(), dtype=float); Pp = np.array(out[3].split(), dtype=float)
    assert np.allclose(S, [5.5, 4, 4, 5.5], rtol=0, atol=1e-15)
    assert np.allclose(K, [0.4912280701754385, -0.17543859649122812, -0.17543859649122817, 0.4912280701754386], atol=1e-15)
    assert np.allclose(xp, [1.0666666666666667, 1.9333333333333333], atol=1e-14)
    assert np.allclose(Pp, [0.19298245614035103, -0.14035087719298228, -0.14035087719298228, 0.19298245614035092], atol=1e-15)
    assert np.array_equal(np.array(out[4].split(), dtype=float), [7.5, 6, 6, 7.5])
    assert out[5].split() == ["19", "0"]


def test_motion_model_quiz():
    """f(x, u) of Robot.cpp:148 against the quiz value in SURVEY.md section 4."""
    so = StructuredOracle(4)
    so.set_pose([1.0, 2.0, np.pi / 4])
    x = so.predict([0.1, 0.0, np.pi / 4])
    # the quiz model adds u(2) once; the reference code uses theta + u[2]/2 for the heading of the step
    # (Robot.cpp:148), so u[2] = pi/4 reproduces the quiz's x and y; theta advances by the full u[2]
    assert np.allclose(x[:2], [1.038268343236509, 2.0923879532511287], atol=1e-15)
    assert abs(x[2] - np.pi / 2) < 1e-15


def test_structured_oracle_is_bitwise_the_reference_at_linesize_1000():
    """BASELINE configs[1] size: the reference's own Robot.cpp compiled with LINESIZE = 1000 / SLAMSIZE = 2003
    (oracle/_ref/libslamref1k.so: the two macros rewritten into a generated header, nothing else touched) against
    the structured oracle at capacity 1000 -- state and the full 2003 x 2003 covariance bit for bit after a
    300-line seeding scan and an update scan.  (~12 s per reference call: its prediction is a dense n^3 dgemm.)"""
    from oracle.oracle import LiteralReference, StructuredOracle, have_literal_1k
    if not have_literal_1k():
        pytest.skip("literal reference at LINESIZE=1000 not built (no /root/reference)")
    lit = LiteralReference(big=True)
    assert lit.capacity == 1000 and lit.n == 2003
    so = StructuredOracle(1000)
    scn = sc.map_scenario(300, 1, m=8, seed=3)
    zero = np.zeros(3)
    lit.localize(scn["seed_z"], scn["seed_R"], sc.encoder_for(zero, zero))
    so.localize(scn["seed_z"], scn["seed_R"], sc.encoder_for(zero, zero))
    y, P, L, pose = lit.state()
    assert L == so.lines == 300
    lit.localize(scn["z"][0], scn["R"][0], sc.encoder_for(pose, scn["u"][0]))
    st, j = so.localize(scn["z"][0], scn["R"][0], sc.encoder_for(so.pose, scn["u"][0]))
    y, P, L, pose = lit.state()
    assert (j >= 0).sum() >= 6
    assert L == so.lines and np.array_equal(y, so.y_full()) and np.array_equal(P, so.P_full())
    assert np.array_equal(pose, so.pose)


def test_structured_oracle_against_golden_vectors_of_the_1k_reference():
    """tests/golden/literal_1k.npz: produced by the reference's own Robot.cpp at LINESIZE = 1000 (make_golden_1k.py).
    Travels to machines without /root/reference: the structured oracle must reproduce it bit for bit."""
    import os
    from oracle.oracle import StructuredOracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "literal_1k.npz"))
    so = StructuredOracle(1000)
    zero = np.zeros(3)
    so.localize(g["seed_z"], g["seed_R"], sc.encoder_for(zero, zero))
    for s in range(g["u"].shape[0]):
        so.localize(g["z"][s], g["R"][s], g["encoder"][s])
        assert np.array_equal(so.pose, g["pose"][s])
    L = int(g["L"][0]); nl = 3 + 2 * L
    assert so.lines == L
    P = so.P_full()
    assert np.array_equal(so.y_full()[:nl], g["y"])
    assert np.array_equal(np.diag(P)[:nl], g["diag"]) and np.array_equal(P[:3, :nl], g["top"])
    for (r, c), blk in zip(g["corners"], g["blocks"]):
        assert np.array_equal(P[r:r + blk.shape[0], c:c + blk.shape[1]], blk)


def test_line_interval_end_points_value_for_value_against_the_reference():
    """Robot::lineIntervals (Robot.cpp:868-879, the node's `lines` topic).  The literal reference is
    driven with line::lineInterval set; what it leaves in lineIntervals.data must equal, float for float, the host-side
    formula of the drop-ins (slam_ros_b200/robot.py interval_end_point == include/ekf_robot.hpp push_endpoint) evaluated
    at the reference's own post-update pose, for exactly the lines it appended, in order."""
    from oracle.oracle import LiteralReference, have_literal
    from slam_ros_b200.robot import interval_end_point
    if not have_literal():
        pytest.skip("oracle/_ref/libslamref.so not present")
    steps = 200
    room = sc.room_scenario(steps=steps, seed=7, range_sigma=5e-5)
    rng = np.random.default_rng(3)
    lit = LiteralReference()
    so = StructuredOracle(100)
    total = 0
    for s in range(steps):
        m = room["count"][s]
        z, R = room["z"][s, :m], room["R"][s, :m]
        iv = np.column_stack([z[:, 0] - 0.3 + 0.05 * rng.standard_normal(m), z[:, 1] * 1.07,
                              z[:, 0] + 0.4 + 0.05 * rng.standard_normal(m), z[:, 1] * 1.11])
        _, _, L0, pose0 = lit.state(want_cov=False)
        enc = sc.encoder_for(pose0, room["u"][s])
        got = lit.localize_intervals(z, R, enc, iv)
        st, j = so.localize(z, R, enc)                  # the structured oracle says which lines were appended
        _, _, L1, pose1 = lit.state(want_cov=False)
        want = []
        for i in range(m):
            if j[i] < 0:
                want.extend(interval_end_point(iv[i, 0], iv[i, 1], *pose1))
                want.extend(interval_end_point(iv[i, 2], iv[i, 3], *pose1))
        want = np.array(want, dtype=np.float32)
        assert got.shape == want.shape, "step %d: %d vs %d floats" % (s, got.size, want.size)
        assert np.array_equal(got, want), "step %d: %s vs %s" % (s, got[:8], want[:8])
        total += got.size
    assert total >= 4 * 20
