"""Replay of a committed full-size oracle fingerprint (tests/golden/oracle_*.npz, made by make_golden_fullsize.py)
on a libekfcuda filter -- TEST INFRASTRUCTURE (used by tests/ and by bench.py's parity extras; never by the product).

No oracle code runs here: the CPU oracle ran when the fixture was generated; this module only feeds the same seeded
scan sequence to the GPU filter through the host-buffer C-ABI call (ekf_scan) and compares
  * the association of every line of every step, bit for bit,
  * the pose after every step,
  * at the fixture's checkpoints: trace / sum of squares of the covariance, sum / sum of squares of y, and the P blocks,
  * y at the last step,
against what the oracle produced.  For a row-sharded filter the ranks' partial read-outs are summed with `reduce_sum`
(each element is owned by exactly one rank, so the sum is a gather).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_path(name):
    return os.path.join(HERE, "oracle_%s.npz" % name)


def load_golden(name):
    p = golden_path(name)
    if not os.path.exists(p):
        return None
    d = np.load(p)
    return {k: d[k] for k in d.files}


def replay(f, g, scn, reduce_sum=None, steps=None):
    """f: EkfFilter already seeded with scn's landmarks; g: load_golden(...); scn: scenario.map_scenario(N, >= steps,
    m, seed) of the SAME N / m / seed.  Returns a dict of error measures (all relative ones against the bar's scale:
    the largest diagonal entry for P, the largest |y| for y)."""
    red = reduce_sum if reduce_sum is not None else (lambda a: a)
    S = int(g["steps"]) if steps is None else min(int(steps), int(g["steps"]))
    bs = int(g["bs"])
    pos = g["pos"]
    ck = {int(s): i for i, s in enumerate(g["ck_step"])}
    scale = float(g["diag_max"])
    out = {"steps": S, "assoc_exact": True, "first_assoc_mismatch": None, "pose_max_abs_err": 0.0, "P_rel_err": 0.0,
           "trace_rel_err": 0.0, "sumsq_rel_err": 0.0, "y_stats_rel_err": 0.0, "y_rel_err": None, "checkpoints": 0,
           "status_nonzero": 0}
    for s in range(S):
        rc, j, pose = f.scan(scn["u"][s], scn["z"][s], scn["R"][s])
        if rc != 0:
            out["status_nonzero"] += 1
        if not np.array_equal(j, g["j_out"][s]):
            if out["assoc_exact"]:
                out["first_assoc_mismatch"] = [int(s), [int(v) for v in j], [int(v) for v in g["j_out"][s]]]
            out["assoc_exact"] = False
        out["pose_max_abs_err"] = max(out["pose_max_abs_err"], float(np.abs(pose - g["pose"][s]).max()))
        if s in ck:
            i = ck[s]
            tr_o, sm_o, sq_o, ys_o, yq_o = (float(v) for v in g["ck_stats"][i])
            tr, sm, sq = (float(v) for v in red(np.array(f.cov_stats())))
            out["trace_rel_err"] = max(out["trace_rel_err"], abs(tr - tr_o) / abs(tr_o))
            out["sumsq_rel_err"] = max(out["sumsq_rel_err"], abs(sq - sq_o) / abs(sq_o))
            y = f.download_y()
            out["y_stats_rel_err"] = max(out["y_stats_rel_err"], abs(float((y * y).sum()) - yq_o) / abs(yq_o))
            for b, (r0, c0) in enumerate(pos):
                blk = red(f.download_block(int(r0), int(c0), bs, bs))
                out["P_rel_err"] = max(out["P_rel_err"], float(np.abs(blk - g["ck_blocks"][i][b]).max()) / scale)
            if "lines" in g and f.lines != int(g["lines"][s]):
                out["lines_mismatch"] = out.get("lines_mismatch", 0) + 1
            out["checkpoints"] += 1
    if S == int(g["steps"]):
        y = f.download_y()
        out["y_rel_err"] = float(np.abs(y - g["y_last"]).max() / np.abs(g["y_last"]).max())
    out["oracle_min_gate_margin"] = float(g["min_margin"])
    return out


def passed(r, tol=1e-9):
    """north_star's bar: association bit-exact, state and P within 1e-9 relative."""
    ok = r["assoc_exact"] and r["status_nonzero"] == 0 and not r.get("lines_mismatch") and r["P_rel_err"] < tol and r["trace_rel_err"] < tol and \
        r["sumsq_rel_err"] < 10 * tol and r["y_stats_rel_err"] < tol and r["pose_max_abs_err"] < 1e-9
    if r["y_rel_err"] is not None:
        ok = ok and r["y_rel_err"] < tol
    return bool(ok)
