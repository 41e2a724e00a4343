"""Independent numpy/scipy statement of the LSC trajectory QP (test infrastructure only).

This is the *second*, independent solve path SURVEY.md §8(c) asks for: it assembles the QP of
reference `src/traj_optimizer.cpp:261-539` (populatebyrow) as dense matrices and solves it by
null-space elimination + Lawson-Hanson least-distance programming (scipy.optimize.nnls).
It is used by the CPU tests to validate the C++ oracle's assembly and its active-set solver,
and by `tools/make_golden.py` to parse the reference's `log/QPmodel.lp` dump.

Nothing here is imported by the product path.
"""
from __future__ import annotations

import math
import re

import numpy as np
from scipy.optimize import nnls

M, NCP, DIM = 5, 6, 3          # segments, control points per segment (n+1), axes
NV = DIM * M * NCP             # 90 decision variables, index k*30 + m*6 + i (traj_optimizer.cpp:277)


def bernstein_basis(n: int = 5) -> np.ndarray:
    """include/polynomial.hpp:415-428 (buildBernsteinBasis)."""
    B = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(i, n + 1):
            B[i, j] = math.comb(n, i) * math.comb(n - i, n - j) * (-1.0) ** (j - i)
    return B


def q_base(dt: float, n: int = 5, phi: int = 3) -> np.ndarray:
    """src/traj_optimizer.cpp:169-184 (buildQBase) with phi_n = 1."""
    def cd(a, k):
        if a < k:
            return 0
        c = 1
        for t in range(k):
            c *= a - t
        return c
    B = bernstein_basis(n)
    Z = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        for j in range(n + 1):
            if i + j - 2 * phi + 1 > 0:
                Z[i, j] = cd(i, phi) * cd(j, phi) / (i + j - 2 * phi + 1)
    return B @ Z @ B.T * dt ** (-2 * phi + 1)


def aeq_axis(dt: float, n: int = 5, phi: int = 3, m_seg: int = M) -> np.ndarray:
    """Per-axis equality matrix: 15 rows of Aeq_base (src/traj_optimizer.cpp:186-236) followed by the
    2 LSC stop rows (src/traj_optimizer.cpp:529-536). Shape (17, 30)."""
    A0 = np.array([[1, 0, 0, 0, 0, 0], [-1, 1, 0, 0, 0, 0], [1, -2, 1, 0, 0, 0]], float)
    AT = np.array([[0, 0, 0, 0, 0, 1], [0, 0, 0, 0, -1, 1], [0, 0, 0, 1, -2, 1]], float)
    rows = []
    nn = 1
    for j in range(phi):
        r = np.zeros(m_seg * (n + 1))
        r[: n + 1] = dt ** (-j) * nn * A0[j]
        rows.append(r)
        nn *= n - j
    for m in range(1, m_seg):
        nn = 1
        for j in range(phi):
            r = np.zeros(m_seg * (n + 1))
            r[(n + 1) * (m - 1): (n + 1) * m] = dt ** (-j) * nn * AT[j]
            r[(n + 1) * m: (n + 1) * (m + 1)] = -(dt ** (-j)) * nn * A0[j]
            rows.append(r)
            nn *= n - j
    for i in range(1, phi):
        r = np.zeros(m_seg * (n + 1))
        r[(m_seg - 1) * (n + 1) + n] = 1.0
        r[(m_seg - 1) * (n + 1) + n - i] = -1.0
        rows.append(r)
    return np.array(rows)


def terminal_segments(pos, goal, v_nom, dt: float) -> int:
    """src/traj_optimizer.cpp:541-548 (float32 point arithmetic, double norm)."""
    d = (np.asarray(goal, np.float32) - np.asarray(pos, np.float32)).astype(np.float32)
    nsq = np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])
    ideal = math.sqrt(float(nsq)) / v_nom
    return max(int((M * dt - ideal + 1e-9) / dt), 1)


def assemble(state, goal, ts, lsc_rows, boxes, world_min, world_max, max_vel, max_acc,
             dt=0.2, w=0.01, wT=1.0):
    """Dense QP:  min x'Px + q'x + c0   s.t.  Aeq x = beq,  Ain x >= bin,  lb <= x <= ub.

    state: (3,3) rows pos/vel/acc (float32 values); goal (3,); ts terminal segments;
    lsc_rows: iterable of (m, i, n3, rhs) meaning n3 . c_{m,i} >= rhs  (rhs = d + n.o);
    boxes: None or (M, 6) array [min xyz, max xyz]  (SFC rows, src/traj_optimizer.cpp:409-434).
    """
    Qb = q_base(dt)
    P = np.zeros((NV, NV))
    q = np.zeros(NV)
    c0 = 0.0
    for k in range(DIM):
        for m in range(M):
            o = k * 30 + m * 6
            P[o:o + 6, o:o + 6] += w * Qb
        for m in range(M - ts, M):
            j = k * 30 + m * 6 + 5
            P[j, j] += wT
            q[j] += -2.0 * wT * float(goal[k])
            c0 += wT * float(goal[k]) ** 2
    A17 = aeq_axis(dt)
    Aeq = np.zeros((51, NV))
    beq = np.zeros(51)
    for k in range(DIM):
        Aeq[17 * k:17 * k + 17, 30 * k:30 * k + 30] = A17
        beq[17 * k:17 * k + 3] = [float(state[0][k]), float(state[1][k]), float(state[2][k])]
    rows, rhs = [], []
    if boxes is not None:
        for m in range(M):
            for k in range(DIM):
                for j in range(NCP):
                    if m == 0 and j < 3:
                        continue
                    r = np.zeros(NV); r[k * 30 + m * 6 + j] = 1.0
                    rows.append(r); rhs.append(float(boxes[m][k]))
                    rows.append(-r); rhs.append(-float(boxes[m][k + 3]))
    for (m, i, n3, b) in lsc_rows:
        r = np.zeros(NV)
        for k in range(DIM):
            r[k * 30 + m * 6 + i] = float(n3[k])
        rows.append(r); rhs.append(float(b))
    for k in range(DIM):
        for m in range(M):
            for i in range(5):
                if m == 0 and i < 2:
                    continue
                r = np.zeros(NV)
                r[k * 30 + m * 6 + i + 1] = 5.0 / dt
                r[k * 30 + m * 6 + i] = -5.0 / dt
                rows.append(-r); rhs.append(-float(max_vel[k]))
                rows.append(r); rhs.append(-float(max_vel[k]))
            for i in range(4):
                if m == 0 and i == 0:
                    continue
                r = np.zeros(NV)
                c = 20.0 / dt ** 2
                r[k * 30 + m * 6 + i + 2] = c
                r[k * 30 + m * 6 + i + 1] = -2 * c
                r[k * 30 + m * 6 + i] = c
                rows.append(-r); rhs.append(-float(max_acc[k]))
                rows.append(r); rhs.append(-float(max_acc[k]))
    lb = np.full(NV, -np.inf); ub = np.full(NV, np.inf)
    for k in range(DIM):
        for m in range(M):
            for i in range(NCP):
                if m == 0 and i < 3:
                    continue
                lb[k * 30 + m * 6 + i] = float(world_min[k])
                ub[k * 30 + m * 6 + i] = float(world_max[k])
    Ain = np.array(rows) if rows else np.zeros((0, NV))
    bin_ = np.array(rhs) if rhs else np.zeros(0)
    return dict(P=P, q=q, c0=c0, Aeq=Aeq, beq=beq, Ain=Ain, bin=bin_, lb=lb, ub=ub)


def solve_ldp(qp, tol=1e-9):
    """Null-space elimination + least-distance programming via NNLS.

    Returns (x, objective, status) with status 'ok' or 'infeasible'.
    """
    P, q, c0 = qp["P"], qp["q"], qp["c0"]
    Aeq, beq = qp["Aeq"], qp["beq"]
    G = [qp["Ain"]]; h = [qp["bin"]]
    I = np.eye(P.shape[0])
    fin = np.isfinite(qp["lb"]); G.append(I[fin]); h.append(qp["lb"][fin])
    fin = np.isfinite(qp["ub"]); G.append(-I[fin]); h.append(-qp["ub"][fin])
    G = np.vstack(G); h = np.concatenate(h)
    # x = xp + Z y
    U, s, Vt = np.linalg.svd(Aeq)
    r = int((s > 1e-9 * s[0]).sum())
    Z = Vt[r:].T
    xp = np.linalg.lstsq(Aeq, beq, rcond=None)[0]
    H = Z.T @ P @ Z
    H = 0.5 * (H + H.T)
    f = Z.T @ (2 * P @ xp + q)          # cost = y'Hy + f'y + const
    L = np.linalg.cholesky(H)
    # u = L' y ; cost = |u|^2 + (L^-1 f)'u = |u - u0|^2 + c ; u0 = -L^-1 f / 2
    u0 = -0.5 * np.linalg.solve(L, f)
    Gt = G @ Z @ np.linalg.inv(L).T      # rows in u-space
    ht = h - G @ xp
    # v = u - u0: min |v|^2 s.t. Gt v >= ht - Gt u0
    hv = ht - Gt @ u0
    scale = np.maximum(np.linalg.norm(Gt, axis=1), 1e-300)
    Gs = Gt / scale[:, None]; hs = hv / scale
    E = np.vstack([Gs.T, hs[None, :]])
    fvec = np.zeros(E.shape[0]); fvec[-1] = 1.0
    lam, rnorm = nnls(E, fvec, maxiter=50 * E.shape[1])
    res = E @ lam - fvec
    if np.linalg.norm(res) < tol:
        return None, None, "infeasible"
    v = -res[:-1] / res[-1]
    u = v + u0
    y = np.linalg.solve(L.T, u)
    x = xp + Z @ y
    obj = float(x @ P @ x + q @ x + c0)
    return x, obj, "ok"


def with_slack(qp, lsc_rows, row_slack, slack_w, n_sfc_rows=0):
    """Extend a dense QP (assemble / oracle dense) by the reference's slack variables (src/traj_optimizer.cpp:317-326,
    383-390,455-457): one eps in (-inf, 0] per (obstacle, segment) entry r of `lsc_rows` (tuples whose first element is
    m) with row_slack[r] != 0, cost slack_w * (M - m) / M * eps^2, subtracted from the left-hand side of the entry's LSC
    rows (6 per entry, 3 for m == 0; they follow the n_sfc_rows SFC rows in Ain). Entries outside the set get no variable:
    theirs only appears in the objective and stays 0. Returns the extended problem; the eps are the trailing variables."""
    idx = [r for r in range(len(lsc_rows)) if row_slack[r]]
    ns = len(idx)
    n0 = qp["P"].shape[0]
    n = n0 + ns
    P = np.zeros((n, n)); P[:n0, :n0] = qp["P"]
    for c, r in enumerate(idx):
        P[n0 + c, n0 + c] = slack_w * (M - int(lsc_rows[r][0])) / M
    q = np.concatenate([qp["q"], np.zeros(ns)])
    Aeq = np.hstack([qp["Aeq"], np.zeros((qp["Aeq"].shape[0], ns))])
    Ain = np.hstack([qp["Ain"], np.zeros((qp["Ain"].shape[0], ns))])
    at = n_sfc_rows
    col = {r: c for c, r in enumerate(idx)}
    for r in range(len(lsc_rows)):
        cnt = 3 if int(lsc_rows[r][0]) == 0 else 6
        if r in col:
            Ain[at:at + cnt, n0 + col[r]] = -1.0
        at += cnt
    lb = np.concatenate([qp["lb"], np.full(ns, -np.inf)])
    ub = np.concatenate([qp["ub"], np.zeros(ns)])
    return dict(P=P, q=q, c0=qp["c0"], Aeq=Aeq, beq=qp["beq"], Ain=Ain, bin=qp["bin"], lb=lb, ub=ub), idx


def kkt_violation(qp, x):
    """max primal violation of all constraint families (equalities, rows, bounds)."""
    v_eq = float(np.max(np.abs(qp["Aeq"] @ x - qp["beq"])))
    v_in = float(np.max(np.maximum(qp["bin"] - qp["Ain"] @ x, 0.0))) if len(qp["bin"]) else 0.0
    v_b = float(max(np.max(np.maximum(qp["lb"] - x, 0)), np.max(np.maximum(x - qp["ub"], 0))))
    return v_eq, max(v_in, v_b)


# --------------------------------------------------------------------------------------
# CPLEX LP-format reader (enough for log/QPmodel.lp)
# --------------------------------------------------------------------------------------
def var_index(name: str) -> int:
    k = "xyz".index(name[0])
    _, m, i = name.split("_")
    return k * 30 + int(m) * 6 + int(i)


_TERM = re.compile(r"([+-])?\s*([0-9.eE+-]+)?\s*([xyz]_\d_\d)(\s*\^2|\s*\*\s*([xyz]_\d_\d))?")


def _lin_terms(expr: str):
    out = []
    for mt in re.finditer(r"([+-])?\s*(\d[0-9.eE+-]*)?\s*([xyz]_\d_\d)", expr):
        sign = -1.0 if mt.group(1) == "-" else 1.0
        coef = float(mt.group(2)) if mt.group(2) else 1.0
        out.append((var_index(mt.group(3)), sign * coef))
    return out


def parse_lp(text: str):
    """Parse a CPLEX LP dump into dense arrays (P,q,c0 with obj = x'Px + q'x + c0)."""
    text = re.sub(r"\\.*", "", text)
    obj_s = text.index("Minimize"); st_s = text.index("Subject To")
    bd_s = text.index("Bounds"); end_s = text.rindex("End")
    obj = text[obj_s + len("Minimize"):st_s]
    obj = obj.split(":", 1)[1]
    lin_part, rest = obj.split("[", 1)
    quad_part, tail = rest.split("]", 1)
    P = np.zeros((NV, NV)); q = np.zeros(NV)
    for j, c in _lin_terms(lin_part):
        q[j] += c
    for mt in re.finditer(r"([+-])?\s*(\d[0-9.eE+-]*)?\s*([xyz]_\d_\d)\s*(\^2|\*\s*([xyz]_\d_\d))", quad_part):
        sign = -1.0 if mt.group(1) == "-" else 1.0
        coef = sign * (float(mt.group(2)) if mt.group(2) else 1.0) / 2.0   # "[ ... ] / 2"
        a = var_index(mt.group(3))
        if mt.group(4).startswith("^"):
            P[a, a] += coef
        else:
            b = var_index(mt.group(5))
            P[a, b] += coef / 2; P[b, a] += coef / 2
    mt = re.search(r"/\s*2\s*([+-])\s*([0-9.eE+-]+)", tail)
    c0 = (1 if mt.group(1) == "+" else -1) * float(mt.group(2)) if mt else 0.0
    cons = text[st_s + len("Subject To"):bd_s]
    eqA, eqb, inA, inb, names = [], [], [], [], []
    for chunk in re.split(r"\n\s*(?=c\d+:)", cons):
        chunk = chunk.strip()
        if not chunk:
            continue
        name, body = chunk.split(":", 1)
        body = " ".join(body.split())
        mt = re.search(r"(>=|<=|=)\s*([+-]?[0-9.eE+-]+)\s*$", body)
        op, rhs = mt.group(1), float(mt.group(2))
        row = np.zeros(NV)
        for j, c in _lin_terms(body[:mt.start()]):
            row[j] += c
        if op == "=":
            eqA.append(row); eqb.append(rhs)
        elif op == ">=":
            inA.append(row); inb.append(rhs); names.append(name.strip())
        else:
            inA.append(-row); inb.append(-rhs); names.append(name.strip())
    lb = np.zeros(NV); ub = np.full(NV, np.inf)      # LP default bounds 0..inf
    for line in text[bd_s + len("Bounds"):end_s].splitlines():
        line = line.strip()
        if not line:
            continue
        mt = re.match(r"([+-]?[0-9.eE+-]+)\s*<=\s*([xyz]_\d_\d)\s*<=\s*([+-]?[0-9.eE+-]+)", line)
        if mt:
            j = var_index(mt.group(2)); lb[j] = float(mt.group(1)); ub[j] = float(mt.group(3)); continue
        mt = re.match(r"([xyz]_\d_\d)\s+[Ff]ree", line)
        if mt:
            j = var_index(mt.group(1)); lb[j] = -np.inf; ub[j] = np.inf; continue
        raise ValueError("unparsed bound line: " + line)
    return dict(P=P, q=q, c0=c0, Aeq=np.array(eqA), beq=np.array(eqb), Ain=np.array(inA),
                bin=np.array(inb), lb=lb, ub=ub, in_names=names)
