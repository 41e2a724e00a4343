"""Scalar numpy model of the active-set iteration the CUDA kernel runs (lsc_planner_b200/csrc/qp_core.cuh).

Test infrastructure only. The kernel keeps a THIN orthonormal basis Q (39 x q) of the active whitened normals N and a
dense coefficient matrix W with N W = Q; no triangular factor. Adding a row appends z/|z| to Q and the column
(-W d / |z|, 1/|z|) to W; dropping active row l reflects the unit vector y ~ W[l, :] onto the last coordinate with one
Householder matrix H (Q <- Q H, W <- W H), deletes the last column of both and moves the last row of W into row l.
This file states exactly that update so that its numerics can be checked against the oracle's full-QR Goldfarb-Idnani
solver (oracle/qp.hpp) on the CPU, on the QPs of real swarm steps, before and independently of the GPU run.
"""
from __future__ import annotations

import numpy as np

FEAS_TOL = 1e-6
ZERO_TOL = 1e-13
NR = 39


def build_rows(lb, ub, vmax, amax, lsc_rows, dt=0.2):
    """All inequality rows a.x >= b in the kernel's canonical id order (device_common.cuh): bounds [0,180),
    dynamic limits [180,450), LSC 450 + 6 s + i. Returns (A [R][90], b [R], ids [R])."""
    A, b, ids = [], [], []
    for v in range(90):
        k, m, i = v // 30, (v % 30) // 6, v % 6
        if m == 0 and i < 3:
            continue
        r = np.zeros(90); r[v] = 1.0
        A.append(r); b.append(lb[v]); ids.append(2 * v)
        A.append(-r); b.append(-ub[v]); ids.append(2 * v + 1)
    vc, ac = 5.0 / dt, 20.0 / dt ** 2
    for k in range(3):
        for m in range(5):
            for j in range(9):
                base = k * 30 + m * 6
                r = np.zeros(90)
                if j < 5:
                    if m == 0 and j < 2:
                        continue
                    r[base + j + 1] = vc; r[base + j] = -vc; lim = vmax[k]
                else:
                    i = j - 5
                    if m == 0 and i == 0:
                        continue
                    r[base + i + 2] = ac; r[base + i + 1] = -2 * ac; r[base + i] = ac; lim = amax[k]
                rid = 180 + ((k * 5 + m) * 9 + j) * 2
                A.append(-r); b.append(-lim); ids.append(rid)          # side 0:  expr <= lim
                A.append(r); b.append(-lim); ids.append(rid + 1)       # side 1: -expr <= lim
    for s, (m, a3, rhs6) in enumerate(lsc_rows):
        for i in range(6):
            if m == 0 and i < 3:
                continue
            r = np.zeros(90)
            for k in range(3):
                r[k * 30 + m * 6 + i] = a3[k]
            A.append(r); b.append(rhs6[i]); ids.append(450 + 6 * s + i)
    return np.array(A), np.array(b), np.array(ids)


def solve(T, state, goal, ts, lb, ub, vmax, amax, lsc_rows, max_iter=2000):
    """state [3][3] rows pos / vel / acc. Returns dict(x, status, iters, n_active, drops)."""
    G = T.G[ts - 1]
    Gk = np.zeros((90, NR))
    for k in range(3):
        Gk[k * 30:(k + 1) * 30, k * 13:(k + 1) * 13] = G
    x = np.zeros(90)
    st = np.asarray(state, float).reshape(3, 3)
    for k in range(3):
        x[k * 30:(k + 1) * 30] = T.Xs[ts - 1] @ st[:, k] + T.xg[ts - 1] * goal[k]
    A, b, ids = build_rows(lb, ub, vmax, amax, lsc_rows, T.dt)
    An = A @ Gk
    nlen = np.linalg.norm(An, axis=1)
    Q = np.zeros((NR, NR)); W = np.zeros((NR, NR))
    act, lam = [], []
    q = 0; iters = 0; drops = 0; status = 0

    def drop(l):
        nonlocal q, Q, W, drops
        drops += 1
        j = q - 1
        y = W[l, :q].copy()
        ny = np.sqrt(y @ y)
        sg = 1.0 if y[j] >= 0 else -1.0
        v = y.copy(); v[j] += sg * ny
        beta = 1.0 / (ny * (ny + abs(y[j])))
        s = Q[:, :q] @ v
        Q[:, :j] -= beta * np.outer(s, v[:j])
        t = W[:q, :q] @ v
        W[:q, :j] -= beta * np.outer(t, v[:j])
        if l != j:
            W[l, :j] = W[j, :j]
            act[l] = act[j]; lam[l] = lam[j]
        act.pop(); lam.pop()
        q = j

    while True:
        slack = A @ x - b
        viol = slack < -FEAS_TOL
        if not viol.any():
            break
        mu = np.where(viol, slack / np.maximum(nlen, 1e-300), np.inf)
        p = int(np.argmin(mu))                      # first minimum = smallest id (rows are in id order)
        if ids[p] in act:
            status = 2; break
        if not nlen[p] > 0:
            status = 1; break
        nv = An[p] / nlen[p]
        sl = slack[p] / nlen[p]
        lam_p = 0.0
        fail = False
        while True:
            iters += 1
            if iters > max_iter:
                status = 2; fail = True; break
            z = nv.copy(); d = np.zeros(q); zz = 1.0
            if q:
                for _ in range(2):
                    c = Q[:, :q].T @ z
                    d += c
                    z = z - Q[:, :q] @ c
                    zz_new = z @ z
                    again = zz_new < 0.25 * zz
                    zz = zz_new
                    if not again:
                        break
            rr = W[:q, :q] @ d
            t1, l = np.inf, -1
            for k in range(q):
                if rr[k] > ZERO_TOL:
                    t = lam[k] / rr[k]
                    if t < t1:
                        t1, l = t, k
            primal = zz > ZERO_TOL
            t2 = max(-sl / zz, 0.0) if primal else np.inf
            t = min(t1, t2)
            if not t < np.inf:
                status = 1; fail = True; break
            for k in range(q):
                lam[k] -= t * rr[k]
            lam_p += t
            if not primal:
                drop(l); continue
            x = x + Gk @ (t * z)
            sl += t * zz
            if t2 <= t1:
                izn = 1.0 / np.sqrt(zz)
                Q[:, q] = z * izn
                W[:q, q] = -rr * izn
                W[q, :q] = 0.0
                W[q, q] = izn
                act.append(int(ids[p])); lam.append(lam_p)
                q += 1
                break
            drop(l)
        if fail:
            break
    return dict(x=x, status=status, iters=iters, n_active=q, drops=drops)
