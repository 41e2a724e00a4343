"""CPU check of the kernel's factorisation update (thin Q + W, Householder drop; tests/qp_kernel_model.py) against the
oracle's full-QR Goldfarb-Idnani solver on the QPs of real closed-loop swarm steps, contact phase included."""
import numpy as np
import pytest

import oracle_lib as O
import qp_kernel_model as KM


def _agent_rows(a, n, nr, d, pred):
    rows = []
    for j in range(n):
        if j == a:
            continue
        for m in range(5):
            a3 = nr[a, j, m].astype(np.float64)
            rhs = d[a, j, m] + pred[j, m].astype(np.float64) @ a3
            rows.append((m, a3, rhs))
    return rows


@pytest.mark.parametrize("n,steps,stride", [(20, 100, 5), (48, 70, 7)])
def test_kernel_update_matches_oracle(n, steps, stride):
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(n, r0=2.0 if n > 20 else 3.0)
    sw = O.Swarm(n, scn.world_min, scn.world_max)
    sw.set_state(scn.start); sw.set_goals(scn.goal); sw.set_capture(True)
    T = O.Tables()
    lb = np.repeat(np.asarray(scn.world_min, float), 30); ub = np.repeat(np.asarray(scn.world_max, float), 30)
    checked = drops = n_fail = 0
    worst = 0.0
    for step in range(steps):
        pos, vel, acc = sw.state()
        sw.step()
        q = sw.qp()
        if step % stride == 0 or step > steps - 25:
            nr, d, _ = sw.capture()
            pred = sw.pred()
            x_o = sw.traj().astype(np.float64)
            # the agents with the most active rows (contact) and a spread
            pick = sorted(set(np.argsort(-q["n_active"])[:4].tolist()) | {0, n // 2})
            for a in pick:
                rows = _agent_rows(a, n, nr, d, pred)
                ts = O.terminal_segments(pos[a], scn.goal[a])
                r = KM.solve(T, np.stack([pos[a], vel[a], acc[a]]).astype(np.float64), scn.goal[a].astype(np.float64), ts,
                             lb, ub, [1, 1, 1], [2, 2, 2], rows)
                assert r["status"] == q["status"][a], (step, a, r["status"], q["status"][a])
                checked += 1; drops += r["drops"]; n_fail += r["status"] != 0
                if r["status"] == 0:
                    xm = r["x"].reshape(3, 5, 6).transpose(1, 2, 0)
                    diff = np.abs(xm - x_o[a]).max()
                    tol = 2e-5 if q["maxviol"][a] > 1e-9 else 2e-6      # oracle trajectories are float32
                    assert diff <= tol, (step, a, diff, q["maxviol"][a])
                    worst = max(worst, diff)
        sw.advance()
    assert checked > 50 and drops > 0
