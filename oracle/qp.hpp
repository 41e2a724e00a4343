// ORACLE — test infrastructure only (see oracle/README.md).
// CPU restatement of the Bernstein trajectory QP of the reference and an FP64 dual active-set
// solver standing in for CPLEX (third-party, IBM ILOG CPLEX 20.1, not in the reference tree;
// call sites src/traj_optimizer.cpp:35-60,76,80,96). Citations relative to /root/reference:
//   src/traj_optimizer.cpp:169-184  buildQBase          -> q_base()
//   src/traj_optimizer.cpp:186-236  buildAeqBase        -> aeq_axis() rows 0..14
//   src/traj_optimizer.cpp:529-536  LSC stop rows       -> aeq_axis() rows 15..16
//   src/traj_optimizer.cpp:239-259  buildDeq            -> equality rhs = (pos,vel,acc)
//   src/traj_optimizer.cpp:261-539  populatebyrow       -> assemble_dense() (row order of the
//                                                          reference: eq, SFC, LSC, dyn, stop)
//   src/traj_optimizer.cpp:541-548  getTerminalSegments -> terminal_segments()
//   include/polynomial.hpp:224-234,415-428  coef_derivative, buildBernsteinBasis
// PARITY UNPINNED by the reference (it records no solutions): the solver is pinned by (i) the
// coefficient-exact comparison of assemble_dense() with the reference's log/QPmodel.lp dump and
// its INFEASIBLE verdict, (ii) an independent null-space + NNLS solve in tests/qp_pyref.py, and
// (iii) KKT residuals. The QP is strictly convex on the equality manifold, so the minimiser is
// unique and the comparison is on coefficients as well as objective/residuals.
#pragma once
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "geom.hpp"

namespace orc {

constexpr int QM = 5, QN = 5, QNCP = 6, QPHI = 3, QDIM = 3;
constexpr int QAX = QM * QNCP;       // 30 variables per axis
constexpr int QNV = QDIM * QAX;      // 90
constexpr int QFREE = 13;            // degrees of freedom per axis
constexpr int QRED = QDIM * QFREE;   // 39

enum QpStatus { QP_OK = 0, QP_INFEASIBLE = 1, QP_MAXITER = 2 };

// Primal feasibility tolerance: CPLEX's default EpRHS = 1e-6 (the reference changes no tolerance,
// src/traj_optimizer.cpp:42-54). As in a dual simplex, an inequality row enters the working set only when it is
// violated by MORE than the tolerance; rows that do enter are then satisfied exactly. Without it the closed loop
// breaks down: trajectories are handed on as float32 (traj_t), so two agents in contact see each other's hulls
// 1e-7 closer than r_i + r_j and an exact solver reports INFEASIBLE where CPLEX returns "optimal".
constexpr double QP_FEAS_TOL = 1e-6;

inline int n_choose_k(int n, int k) {
    if (k > n) return 0;
    if (k * 2 > n) k = n - k;
    if (k == 0) return 1;
    int r = n;
    for (int i = 2; i <= k; i++) { r *= (n - i + 1); r /= i; }
    return r;
}
inline int coef_derivative(int n, int phi) {
    if (n < phi) return 0;
    int c = 1;
    for (int i = 0; i < phi; i++) c *= n - i;
    return c;
}

inline void q_base(double dt, double Q[6][6]) {
    double B[6][6] = {}, Z[6][6] = {}, T[6][6] = {};
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++)
            B[i][j] = n_choose_k(QN, i) * n_choose_k(QN - i, QN - j) * ((j - i) % 2 ? -1.0 : 1.0);
    const int k = QPHI;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++)
            if (i + j - 2 * k + 1 > 0)
                Z[i][j] = (double)coef_derivative(i, k) * coef_derivative(j, k) / (i + j - 2 * k + 1);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int l = 0; l < 6; l++) s += B[i][l] * Z[l][j]; T[i][j] = s; }
    const double sc = std::pow(dt, -2 * k + 1);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int l = 0; l < 6; l++) s += T[i][l] * B[j][l]; Q[i][j] = s * sc; }
}

// 17 x 30: initial-state rows (3), continuity rows (12), stop rows (2)
inline void aeq_axis(double dt, double A[17][QAX]) {
    static const double A0[3][6] = {{1, 0, 0, 0, 0, 0}, {-1, 1, 0, 0, 0, 0}, {1, -2, 1, 0, 0, 0}};
    static const double AT[3][6] = {{0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, -1, 1}, {0, 0, 0, 1, -2, 1}};
    std::memset(A, 0, sizeof(double) * 17 * QAX);
    int nn = 1;
    for (int j = 0; j < QPHI; j++) {
        for (int c = 0; c < 6; c++) A[j][c] = std::pow(dt, -j) * nn * A0[j][c];
        nn *= (QN - j);
    }
    for (int m = 1; m < QM; m++) {
        nn = 1;
        for (int j = 0; j < QPHI; j++) {
            int r = QPHI * m + j;
            for (int c = 0; c < 6; c++) {
                A[r][6 * (m - 1) + c] = std::pow(dt, -j) * nn * AT[j][c];
                A[r][6 * m + c] = -std::pow(dt, -j) * nn * A0[j][c];
            }
            nn *= (QN - j);
        }
    }
    for (int i = 1; i < QPHI; i++) {
        A[14 + i][(QM - 1) * 6 + QN] = 1.0;
        A[14 + i][(QM - 1) * 6 + QN - i] = -1.0;
    }
}

inline int terminal_segments(F3 pos, F3 goal, double v_nom, double dt) {
    double ideal = normf(goal - pos) / v_nom;
    return std::max((int)((QM * dt - ideal + 1e-9) / dt), 1);
}

// ---------------------------------------------------------------------------------------------
// Constant tables of the reduced (null-space, whitened) problem
// ---------------------------------------------------------------------------------------------
struct QpTables {
    double dt, w, wT;
    double Qb[6][6];
    double A17[17][QAX];
    double Xp[QAX][3];            // particular solution map: x = Xp s + Z y
    double Z[QAX][QFREE];
    // per terminal-segment count ts = 1..5 (index ts-1)
    double G[5][QAX][QFREE];      // x = x0 + G v, objective = J(x0) + |v|^2 (per axis block)
    double Xs[5][QAX][3];         // unconstrained minimiser x0 = Xs s + xg * goal
    double xg[5][QAX];
    double gnorm[5][QAX];         // |G row|
};

inline void gauss_solve(std::vector<double>& A, int n, std::vector<double>& Bm, int nrhs) {
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
        if (std::fabs(A[piv * n + c]) < 1e-300) throw std::runtime_error("singular basis");
        if (piv != c) {
            for (int k = 0; k < n; k++) std::swap(A[c * n + k], A[piv * n + k]);
            for (int k = 0; k < nrhs; k++) std::swap(Bm[c * nrhs + k], Bm[piv * nrhs + k]);
        }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = A[r * n + c] / A[c * n + c];
            if (f == 0) continue;
            for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
            for (int k = 0; k < nrhs; k++) Bm[r * nrhs + k] -= f * Bm[c * nrhs + k];
        }
    }
    for (int r = 0; r < n; r++) for (int k = 0; k < nrhs; k++) Bm[r * nrhs + k] /= A[r * n + r];
}

inline void build_tables(double dt, double w, double wT, QpTables& T) {
    T.dt = dt; T.w = w; T.wT = wT;
    q_base(dt, T.Qb);
    aeq_axis(dt, T.A17);
    // free variables: control points 3..5 of segments 0..3 and the last point of segment 4
    int free_idx[QFREE], nfree = 0, basic_idx[17], nbasic = 0;
    for (int m = 0; m < QM; m++)
        for (int i = 0; i < 6; i++) {
            bool is_free = (m < QM - 1) ? (i >= 3) : (i == 5);
            if (is_free) free_idx[nfree++] = 6 * m + i; else basic_idx[nbasic++] = 6 * m + i;
        }
    std::vector<double> AB(17 * 17), RH(17 * (3 + QFREE), 0.0);
    for (int r = 0; r < 17; r++) {
        for (int c = 0; c < 17; c++) AB[r * 17 + c] = T.A17[r][basic_idx[c]];
        if (r < 3) RH[r * 16 + r] = 1.0;
        for (int c = 0; c < QFREE; c++) RH[r * 16 + 3 + c] = -T.A17[r][free_idx[c]];
    }
    gauss_solve(AB, 17, RH, 16);
    std::memset(T.Xp, 0, sizeof T.Xp); std::memset(T.Z, 0, sizeof T.Z);
    for (int b = 0; b < 17; b++) {
        for (int c = 0; c < 3; c++) T.Xp[basic_idx[b]][c] = RH[b * 16 + c];
        for (int c = 0; c < QFREE; c++) T.Z[basic_idx[b]][c] = RH[b * 16 + 3 + c];
    }
    for (int c = 0; c < QFREE; c++) T.Z[free_idx[c]][c] = 1.0;
    for (int ts = 1; ts <= QM; ts++) {
        double P[QAX][QAX] = {};
        for (int m = 0; m < QM; m++)
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) P[6 * m + i][6 * m + j] += w * T.Qb[i][j];
        for (int m = QM - ts; m < QM; m++) P[6 * m + 5][6 * m + 5] += wT;
        double PZ[QAX][QFREE] = {}, H[QFREE][QFREE] = {}, L[QFREE][QFREE] = {};
        for (int i = 0; i < QAX; i++) for (int c = 0; c < QFREE; c++) { double s = 0; for (int j = 0; j < QAX; j++) s += P[i][j] * T.Z[j][c]; PZ[i][c] = s; }
        for (int a = 0; a < QFREE; a++) for (int c = 0; c < QFREE; c++) { double s = 0; for (int i = 0; i < QAX; i++) s += T.Z[i][a] * PZ[i][c]; H[a][c] = s; }
        for (int a = 0; a < QFREE; a++) for (int c = 0; c < a; c++) { double s = 0.5 * (H[a][c] + H[c][a]); H[a][c] = H[c][a] = s; }
        for (int j = 0; j < QFREE; j++) {
            double s = H[j][j];
            for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
            if (s <= 0) throw std::runtime_error("reduced Hessian not PD");
            L[j][j] = std::sqrt(s);
            for (int i = j + 1; i < QFREE; i++) {
                double t = H[i][j];
                for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
                L[i][j] = t / L[j][j];
            }
        }
        // G^T = L^{-1} Z^T  (forward substitution per variable row)
        for (int i = 0; i < QAX; i++) {
            double y[QFREE];
            for (int a = 0; a < QFREE; a++) {
                double t = T.Z[i][a];
                for (int k = 0; k < a; k++) t -= L[a][k] * y[k];
                y[a] = t / L[a][a];
            }
            double nn = 0;
            for (int a = 0; a < QFREE; a++) { T.G[ts - 1][i][a] = y[a]; nn += y[a] * y[a]; }
            T.gnorm[ts - 1][i] = std::sqrt(nn);
        }
        // Xs = Xp - G G^T P Xp ; xg = wT G G^T e_T
        double PX[QAX][3] = {}, eT[QAX] = {};
        for (int i = 0; i < QAX; i++) for (int c = 0; c < 3; c++) { double s = 0; for (int j = 0; j < QAX; j++) s += P[i][j] * T.Xp[j][c]; PX[i][c] = s; }
        for (int m = QM - ts; m < QM; m++) eT[6 * m + 5] = 1.0;
        double GtPX[QFREE][3] = {}, GteT[QFREE] = {};
        for (int a = 0; a < QFREE; a++) {
            for (int c = 0; c < 3; c++) { double s = 0; for (int i = 0; i < QAX; i++) s += T.G[ts - 1][i][a] * PX[i][c]; GtPX[a][c] = s; }
            double s = 0; for (int i = 0; i < QAX; i++) s += T.G[ts - 1][i][a] * eT[i]; GteT[a] = s;
        }
        for (int i = 0; i < QAX; i++) {
            for (int c = 0; c < 3; c++) { double s = 0; for (int a = 0; a < QFREE; a++) s += T.G[ts - 1][i][a] * GtPX[a][c]; T.Xs[ts - 1][i][c] = T.Xp[i][c] - s; }
            double s = 0; for (int a = 0; a < QFREE; a++) s += T.G[ts - 1][i][a] * GteT[a];
            T.xg[ts - 1][i] = wT * s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Problem / result containers
// ---------------------------------------------------------------------------------------------
struct LscRows {        // one entry per (obstacle, segment): a . c_{m,i} >= rhs[i]
    int m;
    double a[3];        // widened float normal (traj_optimizer.cpp:446-452)
    double rhs[6];      // d_i + a . o_{m,i}
    int slack = 0;      // 1: the obstacle is in obs_slack_indices -> rows read a . c - eps_{oi,m} >= rhs (traj_optimizer.cpp:455-457)
};

struct QpProblem {
    double s[3][3];         // [pos|vel|acc][axis]
    double goal[3];
    int ts;
    double lb[QNV], ub[QNV];   // -inf/+inf where free (m == 0, i < 3)
    double vmax[3], amax[3];
    const LscRows* rows; int n_rows;
    double slack_w = 1.0;   // opt/slack_collision_weight (src/param.cpp:75): cost w ((M - m) / M) eps^2 (traj_optimizer.cpp:383-390)
};

struct QpResult {
    double x[QNV];
    double cost;
    int status, iters, n_active;
    double kkt_stationarity;     // | v - sum lambda_k n_k |_inf in the whitened space
    double max_violation;        // max over all rows of the (unnormalised) violation
    int act_ids[QRED]; int n_act_ids = 0;   // canonical ids of the rows active at the solution (qp_solve)
    int warm_accepted = 0;       // qp_solve with a guess: rows of the guess the start point kept (0: cold start)
    std::vector<double> eps;     // qp_solve_slack: slack variable of every (obstacle, segment) row entry, <= 0
    double slack_cost = 0;       // ... their share of `cost`
};

inline double objective(const QpTables& T, const QpProblem& p, const double* x) {
    double J = 0;
    for (int k = 0; k < 3; k++) {
        for (int m = 0; m < QM; m++) {
            const double* c = x + k * QAX + 6 * m;
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++)
                if (T.Qb[i][j] != 0) J += T.w * T.Qb[i][j] * c[i] * c[j];
        }
        for (int m = QM - p.ts; m < QM; m++) {
            double e = x[k * QAX + 6 * m + 5] - p.goal[k];
            J += T.wT * e * e;
        }
    }
    return J;
}

// canonical constraint ids (shared convention with the CUDA engine):
//   [0,180)    bounds   ((k*5+m)*6+i)*2 + side         side 0: x >= lb, side 1: x <= ub
//   [180,450)  dynamic  180 + ((k*5+m)*9+j)*2 + side   j<5 velocity i=j, j>=5 acceleration i=j-5
//                                                      side 0: +expr <= lim, side 1: -expr <= lim
//   [450,..)   LSC      450 + row*6 + i
struct RowView { int nnz; int idx[3]; double a[3]; double b; };

inline bool row_get(const QpTables& T, const QpProblem& p, int id, RowView& r) {
    const double dt = T.dt;
    if (id < 180) {
        int side = id & 1, v = id >> 1; int i = v % 6, m = (v / 6) % 5, k = v / 30;
        if (m == 0 && i < 3) return false;
        int j = k * QAX + 6 * m + i;
        r.nnz = 1; r.idx[0] = j;
        if (side == 0) { if (!(p.lb[j] > -INFINITY)) return false; r.a[0] = 1; r.b = p.lb[j]; }
        else { if (!(p.ub[j] < INFINITY)) return false; r.a[0] = -1; r.b = -p.ub[j]; }
        return true;
    }
    if (id < 450) {
        int e = id - 180; int side = e & 1, v = e >> 1; int j = v % 9, m = (v / 9) % 5, k = v / 45;
        double sg = side == 0 ? -1.0 : 1.0;
        if (j < 5) {
            if (m == 0 && j < 2) return false;
            double c = std::pow(dt, -1) * QN;
            r.nnz = 2; r.idx[0] = k * QAX + 6 * m + j + 1; r.idx[1] = k * QAX + 6 * m + j;
            r.a[0] = sg * c; r.a[1] = -sg * c; r.b = -p.vmax[k];
        } else {
            int i = j - 5;
            if (m == 0 && i == 0) return false;
            double c = std::pow(dt, -2) * QN * (QN - 1);
            r.nnz = 3; r.idx[0] = k * QAX + 6 * m + i + 2; r.idx[1] = k * QAX + 6 * m + i + 1; r.idx[2] = k * QAX + 6 * m + i;
            r.a[0] = sg * c; r.a[1] = -2 * sg * c; r.a[2] = sg * c; r.b = -p.amax[k];
        }
        return true;
    }
    int e = id - 450; int row = e / 6, i = e % 6;
    if (row >= p.n_rows) return false;
    const LscRows& L = p.rows[row];
    if (L.m == 0 && i < QPHI) return false;
    r.nnz = 3;
    for (int k = 0; k < 3; k++) { r.idx[k] = k * QAX + 6 * L.m + i; r.a[k] = L.a[k]; }
    r.b = L.rhs[i];
    return true;
}

// Test hook (debugging the CUDA kernel's two-tier pricing on the CPU): when finite, LSC rows are priced in two tiers
// exactly like k_qp_solve — a working set of (obstacle, segment) pairs whose whitened slack at x0 is below this
// threshold, extended by a full sweep whenever nothing in it is violated. Default: every row priced every iteration.
inline double g_tier_threshold = INFINITY;
inline int g_trace_agent = -1;            // debugging: >= 0 prints every pivot of the solve (stderr)
// warm-start experiment knob: how often candidates with a negative multiplier are removed and the rest re-tried before the
// solve starts cold (environment ORC_WARM_ATTEMPTS; the kernel's rule is 2)
inline int g_warm_attempts = std::getenv("ORC_WARM_ATTEMPTS") ? std::atoi(std::getenv("ORC_WARM_ATTEMPTS")) : 2;

// Goldfarb-Idnani dual active set on  min |v|^2  s.t.  n_j . v >= -slack0_j   (x = x0 + G v)
// Optional warm start (guess / n_guess: canonical row ids, e.g. the previous step's active set shifted by one segment):
// the rows of the guess are factorised at once, v = argmin |v|^2 s.t. n_k . v = b_k on them, and the pair (v, guess) is
// accepted as the starting point iff every multiplier is >= 0 (then it is an S-pair of Goldfarb-Idnani, from which the
// iteration below converges to the same unique minimiser); rows with a negative multiplier are removed once and the
// rest re-tried, otherwise the solve starts cold. Restates the kernel's warm start (csrc/qp_core.cuh) for CPU experiments.
inline void qp_solve(const QpTables& T, const QpProblem& p, QpResult& out, int max_iter = 2000, const int* guess = nullptr,
                     int n_guess = 0) {
    const int n = QRED;
    const int ts = p.ts;
    double x[QNV], v[QRED] = {};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < QAX; i++)
            x[k * QAX + i] = T.Xs[ts - 1][i][0] * p.s[0][k] + T.Xs[ts - 1][i][1] * p.s[1][k] +
                             T.Xs[ts - 1][i][2] * p.s[2][k] + T.xg[ts - 1][i] * p.goal[k];
    std::vector<double> J(n * n, 0.0), R(n * n, 0.0);
    for (int i = 0; i < n; i++) J[i * n + i] = 1.0;
    int q = 0, act[QRED]; double lam[QRED];
    std::vector<double> Nact(n * n, 0.0);   // active normals (columns), for the KKT report only
    const int n_ids = 450 + 6 * p.n_rows;
    std::vector<char> is_active(n_ids, 0);
    const double tol = 1e-10;
    int iters = 0;
    out.status = QP_OK;

    auto row_normal = [&](const RowView& r, double* nv) -> double {     // nv = G^T a ; returns |nv|
        std::fill(nv, nv + n, 0.0);
        for (int t = 0; t < r.nnz; t++) {
            int k = r.idx[t] / QAX, i = r.idx[t] % QAX;
            for (int c = 0; c < QFREE; c++) nv[k * QFREE + c] += r.a[t] * T.G[ts - 1][i][c];
        }
        double s = 0; for (int c = 0; c < n; c++) s += nv[c] * nv[c];
        return std::sqrt(s);
    };
    auto row_slack = [&](const RowView& r) { double s = -r.b; for (int t = 0; t < r.nnz; t++) s += r.a[t] * x[r.idx[t]]; return s; };
    auto drop = [&](int l) {
        is_active[act[l]] = 0;
        for (int j = l; j < q - 1; j++) {
            act[j] = act[j + 1]; lam[j] = lam[j + 1];
            for (int r = 0; r < n; r++) { R[r * n + j] = R[r * n + j + 1]; Nact[r * n + j] = Nact[r * n + j + 1]; }
        }
        for (int r = 0; r < n; r++) R[r * n + q - 1] = 0.0;
        q--;
        for (int j = l; j < q; j++) {              // zero the sub-diagonal R[j+1][j]
            double a = R[j * n + j], b = R[(j + 1) * n + j];
            if (b == 0.0) continue;
            double h = std::hypot(a, b), c = a / h, s = b / h;
            for (int k = j; k < q; k++) {
                double t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
                R[j * n + k] = c * t1 + s * t2; R[(j + 1) * n + k] = -s * t1 + c * t2;
            }
            for (int r = 0; r < n; r++) {
                double t1 = J[r * n + j], t2 = J[r * n + j + 1];
                J[r * n + j] = c * t1 + s * t2; J[r * n + j + 1] = -s * t1 + c * t2;
            }
        }
    };

    double nv[QRED], d[QRED], z[QRED], rr[QRED];
    const bool tiered = g_tier_threshold < INFINITY;
    std::vector<char> in_work(tiered ? p.n_rows : 0, 0);
    int pbest = -1; double mu_best = -tol; RowView rb; double nb = 0;
    auto price = [&](int id) -> bool {          // returns true when the row is violated beyond the tolerance
        if (is_active[id]) return false;
        RowView r;
        if (!row_get(T, p, id, r)) return false;
        double nn;
        if (r.nnz == 1) nn = T.gnorm[ts - 1][r.idx[0] % QAX];
        else if (id >= 450) nn = std::sqrt(r.a[0] * r.a[0] + r.a[1] * r.a[1] + r.a[2] * r.a[2]) * T.gnorm[ts - 1][r.idx[0] % QAX];
        else nn = row_normal(r, nv);
        double sl = row_slack(r);
        if (tiered && id >= 450 && pbest == -2) {       // initial working-set selection at x0
            if (!(sl / std::max(nn, 1e-300) >= g_tier_threshold)) in_work[(id - 450) / 6] = 1;
            return false;
        }
        if (!(sl < -QP_FEAS_TOL)) return false;
        double mu = sl / std::max(nn, 1e-300);
        if (mu < mu_best) { mu_best = mu; pbest = id; rb = r; nb = nn; }
        return true;
    };
    if (tiered) { pbest = -2; for (int id = 450; id < n_ids; id++) price(id); }
    out.warm_accepted = 0;
    if (guess && n_guess > 0) {
        std::vector<int> cand(guess, guess + n_guess);
        for (int attempt = 0; attempt < g_warm_attempts && !cand.empty(); attempt++) {
            // factorise the candidate rows: J, R as after adding them one by one (no steps)
            std::fill(J.begin(), J.end(), 0.0); std::fill(R.begin(), R.end(), 0.0);
            for (int i = 0; i < n; i++) J[i * n + i] = 1.0;
            int qq = 0; int ids[QRED]; double bb[QRED];
            for (int id : cand) {
                RowView r;
                if (qq >= n || id < 0 || id >= n_ids || !row_get(T, p, id, r)) continue;
                bool dup = false; for (int k = 0; k < qq; k++) dup |= ids[k] == id;
                if (dup) continue;
                const double nrm = row_normal(r, nv);
                if (!(nrm > 0)) continue;
                for (int c = 0; c < n; c++) nv[c] /= nrm;
                for (int c = 0; c < n; c++) { double s2 = 0; for (int rr2 = 0; rr2 < n; rr2++) s2 += J[rr2 * n + c] * nv[rr2]; d[c] = s2; }
                double zz = 0; for (int c = qq; c < n; c++) zz += d[c] * d[c];
                if (!(zz > 1e-10)) continue;                     // (numerically) dependent on the rows taken so far
                for (int j = n - 1; j > qq; j--) {
                    double a = d[j - 1], b = d[j];
                    if (b == 0.0) continue;
                    double h = std::hypot(a, b), c = a / h, s2 = b / h;
                    d[j - 1] = h; d[j] = 0;
                    for (int rr2 = 0; rr2 < n; rr2++) {
                        double u1 = J[rr2 * n + j - 1], u2 = J[rr2 * n + j];
                        J[rr2 * n + j - 1] = c * u1 + s2 * u2; J[rr2 * n + j] = -s2 * u1 + c * u2;
                    }
                }
                for (int k = 0; k <= qq; k++) R[k * n + qq] = d[k];
                for (int rr2 = 0; rr2 < n; rr2++) Nact[rr2 * n + qq] = nv[rr2];
                ids[qq] = id; bb[qq] = -row_slack(r) / nrm;      // n . v >= b  with v measured from x0
                qq++;
            }
            // R^T R lambda = b;  v = J1 (R lambda)
            double y[QRED], lm[QRED];
            for (int k = 0; k < qq; k++) { double s2 = bb[k]; for (int c = 0; c < k; c++) s2 -= R[c * n + k] * y[c]; y[k] = s2 / R[k * n + k]; }
            for (int k = qq - 1; k >= 0; k--) { double s2 = y[k]; for (int c = k + 1; c < qq; c++) s2 -= R[k * n + c] * lm[c]; lm[k] = s2 / R[k * n + k]; }
            bool ok = true;
            for (int k = 0; k < qq; k++) ok &= lm[k] >= -1e-12;
            if (ok) {
                q = qq;
                for (int k = 0; k < q; k++) { act[k] = ids[k]; lam[k] = std::max(lm[k], 0.0); is_active[ids[k]] = 1; }
                for (int rr2 = 0; rr2 < n; rr2++) { double s2 = 0; for (int k = 0; k < q; k++) s2 += J[rr2 * n + k] * y[k]; v[rr2] = s2; }
                for (int k = 0; k < 3; k++)
                    for (int i = 0; i < QAX; i++) {
                        double s2 = 0;
                        for (int c = 0; c < QFREE; c++) s2 += T.G[ts - 1][i][c] * v[k * QFREE + c];
                        x[k * QAX + i] += s2;
                    }
                out.warm_accepted = q;
                break;
            }
            std::vector<int> keep;
            for (int k = 0; k < qq; k++) if (lm[k] >= -1e-12) keep.push_back(ids[k]);
            cand.swap(keep);
            if (attempt == g_warm_attempts - 1 || cand.empty()) {                  // cold start
                std::fill(J.begin(), J.end(), 0.0); std::fill(R.begin(), R.end(), 0.0);
                for (int i = 0; i < n; i++) J[i * n + i] = 1.0;
            }
        }
        if (!out.warm_accepted) { std::fill(J.begin(), J.end(), 0.0); std::fill(R.begin(), R.end(), 0.0); for (int i = 0; i < n; i++) J[i * n + i] = 1.0; }
    }
    while (true) {
        // pricing: most violated row in whitened distance
        pbest = -1; mu_best = -tol;
        if (!tiered) {
            for (int id = 0; id < n_ids; id++) price(id);
        } else {
            for (int id = 0; id < 450; id++) price(id);
            for (int id = 450; id < n_ids; id++) if (in_work[(id - 450) / 6]) price(id);
            if (pbest < 0)
                for (int id = 450; id < n_ids; id++) if (price(id)) in_work[(id - 450) / 6] = 1;
        }
        if (pbest < 0) break;
        if (g_trace_agent >= 0) std::fprintf(stderr, "  pick id %d (%s) mu %.4g q %d\n", pbest, pbest < 180 ? "bound" : pbest < 450 ? "dyn" : "lsc", mu_best, q);
        double nrm = row_normal(rb, nv);
        (void)nb;
        if (!(nrm > 0)) { out.status = QP_INFEASIBLE; break; }
        for (int c = 0; c < n; c++) nv[c] /= nrm;
        double lam_p = 0;
        bool fail = false;
        while (true) {
            if (++iters > max_iter) { out.status = QP_MAXITER; fail = true; break; }
            for (int c = 0; c < n; c++) { double s = 0; for (int r = 0; r < n; r++) s += J[r * n + c] * nv[r]; d[c] = s; }
            double zz = 0;
            for (int c = q; c < n; c++) zz += d[c] * d[c];
            for (int r = 0; r < n; r++) { double s = 0; for (int c = q; c < n; c++) s += J[r * n + c] * d[c]; z[r] = s; }
            for (int k = q - 1; k >= 0; k--) {
                double s = d[k];
                for (int c = k + 1; c < q; c++) s -= R[k * n + c] * rr[c];
                rr[k] = s / R[k * n + k];
            }
            double t1 = INFINITY; int l = -1;
            for (int k = 0; k < q; k++)
                if (rr[k] > 1e-13) { double t = lam[k] / rr[k]; if (t < t1) { t1 = t; l = k; } }
            const bool primal = zz > 1e-13;
            double slack = row_slack(rb) / nrm;
            double t2 = primal ? -slack / zz : INFINITY;
            if (t2 < 0) t2 = 0;
            double t = std::min(t1, t2);
            if (!(t < INFINITY)) { out.status = QP_INFEASIBLE; fail = true; break; }
            for (int k = 0; k < q; k++) lam[k] -= t * rr[k];
            lam_p += t;
            if (g_trace_agent >= 0 && !(t2 <= t1)) std::fprintf(stderr, "    drop id %d (t1 %.3g t2 %.3g)\n", act[l], t1, t2);
            if (!primal) { drop(l); continue; }
            for (int c = 0; c < n; c++) v[c] += t * z[c];
            for (int k = 0; k < 3; k++)
                for (int i = 0; i < QAX; i++) {
                    double s = 0;
                    for (int c = 0; c < QFREE; c++) s += T.G[ts - 1][i][c] * z[k * QFREE + c];
                    x[k * QAX + i] += t * s;
                }
            if (t2 <= t1) {
                // add constraint: Givens from the bottom so that J^T n = (d_0..d_{q-1}, +-|d2|, 0..)
                for (int j = n - 1; j > q; j--) {
                    double a = d[j - 1], b = d[j];
                    if (b == 0.0) continue;
                    double h = std::hypot(a, b), c = a / h, s = b / h;
                    d[j - 1] = h; d[j] = 0;
                    for (int r = 0; r < n; r++) {
                        double u1 = J[r * n + j - 1], u2 = J[r * n + j];
                        J[r * n + j - 1] = c * u1 + s * u2; J[r * n + j] = -s * u1 + c * u2;
                    }
                }
                for (int k = 0; k <= q; k++) R[k * n + q] = d[k];
                for (int r = 0; r < n; r++) Nact[r * n + q] = nv[r];
                act[q] = pbest; lam[q] = lam_p; is_active[pbest] = 1; q++;
                break;
            }
            drop(l);
        }
        if (fail) break;
    }
    std::memcpy(out.x, x, sizeof x);
    out.cost = objective(T, p, x);
    out.iters = iters; out.n_active = q;
    out.n_act_ids = q; for (int k = 0; k < q; k++) out.act_ids[k] = act[k];
    double st = 0;
    for (int r = 0; r < n; r++) { double s = v[r]; for (int k = 0; k < q; k++) s -= lam[k] * Nact[r * n + k]; st = std::max(st, std::fabs(s)); }
    out.kkt_stationarity = st;
    double mv = 0;
    for (int id = 0; id < n_ids; id++) { RowView r; if (!row_get(T, p, id, r)) continue; mv = std::max(mv, -row_slack(r)); }
    out.max_violation = mv;
}

// ---------------------------------------------------------------------------------------------
// The QP with slack variables (src/traj_optimizer.cpp:317-326,383-390,455-457): once obs_slack_indices is not empty the
// reference adds one variable eps_{oi,m} in (-inf, 0] per (obstacle, segment) with cost w_s ((M - m) / M) eps^2, and the
// LSC rows of the obstacles IN the set read  a . (c_{m,i} - o_{m,i}) - d_i - eps_{oi,m} >= 0.  Variables of obstacles
// outside the set appear in the objective only and stay 0.
//
// Same dual active set in the extended whitened space u = (v, e), e_p = sqrt(c_p) eps_p, objective |v|^2 + |e|^2:
//   slack row   (G^T a, -1/sqrt(c_p) at coordinate p) . u >= -slack0        |N|^2 = |G^T a|^2 + 1/c_p
//   bound row   (0, -1/sqrt(c_p) at p) . u >= 0                              (eps_p <= 0)
// A coordinate p is created when a row of its pair first enters the working set (until then e_p = 0 and no active
// normal touches it: the factorisation is extended by a unit column). Row ids: LSC rows as in qp_solve; the bound of
// row entry r is 450 + 6 n_rows + r. Without slack rows the iteration is qp_solve's, operation for operation.
// ---------------------------------------------------------------------------------------------
inline void qp_solve_slack(const QpTables& T, const QpProblem& p, QpResult& out, int max_iter = 2000) {
    const int ts = p.ts;
    int n = QRED;                          // current dimension
    int cap = QRED + 32;                   // stride of J, R, Nact
    double x[QNV];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < QAX; i++)
            x[k * QAX + i] = T.Xs[ts - 1][i][0] * p.s[0][k] + T.Xs[ts - 1][i][1] * p.s[1][k] +
                             T.Xs[ts - 1][i][2] * p.s[2][k] + T.xg[ts - 1][i] * p.goal[k];
    std::vector<double> J((size_t)cap * cap, 0.0), R((size_t)cap * cap, 0.0), Nact((size_t)cap * cap, 0.0), v(cap, 0.0);
    for (int i = 0; i < n; i++) J[(size_t)i * cap + i] = 1.0;
    int q = 0;
    std::vector<int> act(cap); std::vector<double> lam(cap);
    const int n_ids = 450 + 6 * p.n_rows + p.n_rows;
    const int id_bound0 = 450 + 6 * p.n_rows;
    std::vector<char> is_active(n_ids, 0);
    std::vector<int> coord(p.n_rows, -1);          // row entry -> slack coordinate (index into u, >= QRED)
    std::vector<double> isc(p.n_rows, 0.0);        // 1 / sqrt(c_p)
    for (int r = 0; r < p.n_rows; r++)
        if (p.rows[r].slack) isc[r] = 1.0 / std::sqrt(p.slack_w * ((double)(QM - p.rows[r].m) / QM));
    const double tol = 1e-10;
    int iters = 0;
    out.status = QP_OK;

    auto grow = [&]() {
        const int ncap = cap * 2;
        auto re = [&](std::vector<double>& A) {
            std::vector<double> B((size_t)ncap * ncap, 0.0);
            for (int r = 0; r < cap; r++) for (int c = 0; c < cap; c++) B[(size_t)r * ncap + c] = A[(size_t)r * cap + c];
            A.swap(B);
        };
        re(J); re(R); re(Nact);
        v.resize(ncap, 0.0); act.resize(ncap); lam.resize(ncap);
        cap = ncap;
    };
    auto eps_of = [&](int r) { return coord[r] >= 0 ? v[coord[r]] * isc[r] : 0.0; };
    // slack and extended norm of a row id; false when the row does not exist
    struct Priced { double slack, norm; };
    auto eval = [&](int id, Priced& pr) -> bool {
        if (id >= id_bound0) {
            const int r = id - id_bound0;
            if (coord[r] < 0) return false;
            pr.slack = -eps_of(r); pr.norm = isc[r];
            return true;
        }
        RowView rv;
        if (!row_get(T, p, id, rv)) return false;
        double s = -rv.b;
        for (int t = 0; t < rv.nnz; t++) s += rv.a[t] * x[rv.idx[t]];
        double nn;
        if (rv.nnz == 1) nn = T.gnorm[ts - 1][rv.idx[0] % QAX];
        else if (id >= 450) nn = std::sqrt(rv.a[0] * rv.a[0] + rv.a[1] * rv.a[1] + rv.a[2] * rv.a[2]) * T.gnorm[ts - 1][rv.idx[0] % QAX];
        else {
            double nv[QRED] = {};
            for (int t = 0; t < rv.nnz; t++) {
                int k = rv.idx[t] / QAX, i = rv.idx[t] % QAX;
                for (int c = 0; c < QFREE; c++) nv[k * QFREE + c] += rv.a[t] * T.G[ts - 1][i][c];
            }
            double s2 = 0; for (int c = 0; c < QRED; c++) s2 += nv[c] * nv[c];
            nn = std::sqrt(s2);
        }
        if (id >= 450) {
            const int r = (id - 450) / 6;
            if (p.rows[r].slack) { s -= eps_of(r); nn = std::sqrt(nn * nn + isc[r] * isc[r]); }
        }
        pr.slack = s; pr.norm = nn;
        return true;
    };
    // unnormalised extended normal of the entering row (creates the slack coordinate when needed); returns its length
    std::vector<double> nv, d, z, rr;
    auto ext_normal = [&](int id) -> double {
        int r_slack = -1;
        if (id >= id_bound0) r_slack = id - id_bound0;
        else if (id >= 450 && p.rows[(id - 450) / 6].slack) r_slack = (id - 450) / 6;
        if (r_slack >= 0 && coord[r_slack] < 0) {
            if (n == cap) grow();
            coord[r_slack] = n;
            J[(size_t)n * cap + n] = 1.0;
            n++;
        }
        nv.assign(cap, 0.0);
        if (id < id_bound0) {
            RowView rv; row_get(T, p, id, rv);
            for (int t = 0; t < rv.nnz; t++) {
                int k = rv.idx[t] / QAX, i = rv.idx[t] % QAX;
                for (int c = 0; c < QFREE; c++) nv[k * QFREE + c] += rv.a[t] * T.G[ts - 1][i][c];
            }
        }
        if (r_slack >= 0) nv[coord[r_slack]] = -isc[r_slack];
        double s = 0; for (int c = 0; c < n; c++) s += nv[c] * nv[c];
        return std::sqrt(s);
    };
    auto drop = [&](int l) {
        is_active[act[l]] = 0;
        for (int j = l; j < q - 1; j++) {
            act[j] = act[j + 1]; lam[j] = lam[j + 1];
            for (int r = 0; r < n; r++) { R[(size_t)r * cap + j] = R[(size_t)r * cap + j + 1]; Nact[(size_t)r * cap + j] = Nact[(size_t)r * cap + j + 1]; }
        }
        for (int r = 0; r < n; r++) R[(size_t)r * cap + q - 1] = 0.0;
        q--;
        for (int j = l; j < q; j++) {
            double a = R[(size_t)j * cap + j], b = R[(size_t)(j + 1) * cap + j];
            if (b == 0.0) continue;
            double h = std::hypot(a, b), c = a / h, s = b / h;
            for (int k = j; k < q; k++) {
                double t1 = R[(size_t)j * cap + k], t2 = R[(size_t)(j + 1) * cap + k];
                R[(size_t)j * cap + k] = c * t1 + s * t2; R[(size_t)(j + 1) * cap + k] = -s * t1 + c * t2;
            }
            for (int r = 0; r < n; r++) {
                double t1 = J[(size_t)r * cap + j], t2 = J[(size_t)r * cap + j + 1];
                J[(size_t)r * cap + j] = c * t1 + s * t2; J[(size_t)r * cap + j + 1] = -s * t1 + c * t2;
            }
        }
    };

    while (true) {
        // pricing: most violated row by distance in the extended whitened space
        int pbest = -1; double mu_best = -tol;
        for (int id = 0; id < n_ids; id++) {
            if (is_active[id]) continue;
            Priced pr;
            if (!eval(id, pr)) continue;
            if (!(pr.slack < -QP_FEAS_TOL)) continue;
            const double mu = pr.slack / std::max(pr.norm, 1e-300);
            if (mu < mu_best) { mu_best = mu; pbest = id; }
        }
        if (pbest < 0) break;
        const double nrm = ext_normal(pbest);
        if (!(nrm > 0)) { out.status = QP_INFEASIBLE; break; }
        d.assign(cap, 0.0); z.assign(cap, 0.0); rr.assign(cap, 0.0);
        for (int c = 0; c < n; c++) nv[c] /= nrm;
        double lam_p = 0;
        bool fail = false;
        while (true) {
            if (++iters > max_iter) { out.status = QP_MAXITER; fail = true; break; }
            for (int c = 0; c < n; c++) { double s = 0; for (int r = 0; r < n; r++) s += J[(size_t)r * cap + c] * nv[r]; d[c] = s; }
            double zz = 0;
            for (int c = q; c < n; c++) zz += d[c] * d[c];
            for (int r = 0; r < n; r++) { double s = 0; for (int c = q; c < n; c++) s += J[(size_t)r * cap + c] * d[c]; z[r] = s; }
            for (int k = q - 1; k >= 0; k--) {
                double s = d[k];
                for (int c = k + 1; c < q; c++) s -= R[(size_t)k * cap + c] * rr[c];
                rr[k] = s / R[(size_t)k * cap + k];
            }
            double t1 = INFINITY; int l = -1;
            for (int k = 0; k < q; k++)
                if (rr[k] > 1e-13) { double t = lam[k] / rr[k]; if (t < t1) { t1 = t; l = k; } }
            const bool primal = zz > 1e-13;
            Priced pr{0.0, 1.0}; eval(pbest, pr);
            const double slack = pr.slack / nrm;
            double t2 = primal ? -slack / zz : INFINITY;
            if (t2 < 0) t2 = 0;
            double t = std::min(t1, t2);
            if (!(t < INFINITY)) { out.status = QP_INFEASIBLE; fail = true; break; }
            for (int k = 0; k < q; k++) lam[k] -= t * rr[k];
            lam_p += t;
            if (!primal) { drop(l); continue; }
            for (int c = 0; c < n; c++) v[c] += t * z[c];
            for (int k = 0; k < 3; k++)
                for (int i = 0; i < QAX; i++) {
                    double s = 0;
                    for (int c = 0; c < QFREE; c++) s += T.G[ts - 1][i][c] * z[k * QFREE + c];
                    x[k * QAX + i] += t * s;
                }
            if (t2 <= t1) {
                for (int j = n - 1; j > q; j--) {
                    double a = d[j - 1], b = d[j];
                    if (b == 0.0) continue;
                    double h = std::hypot(a, b), c = a / h, s = b / h;
                    d[j - 1] = h; d[j] = 0;
                    for (int r = 0; r < n; r++) {
                        double u1 = J[(size_t)r * cap + j - 1], u2 = J[(size_t)r * cap + j];
                        J[(size_t)r * cap + j - 1] = c * u1 + s * u2; J[(size_t)r * cap + j] = -s * u1 + c * u2;
                    }
                }
                for (int k = 0; k <= q; k++) R[(size_t)k * cap + q] = d[k];
                for (int r = 0; r < n; r++) Nact[(size_t)r * cap + q] = nv[r];
                act[q] = pbest; lam[q] = lam_p; is_active[pbest] = 1; q++;
                break;
            }
            drop(l);
        }
        if (fail) break;
    }
    std::memcpy(out.x, x, sizeof x);
    out.eps.assign(p.n_rows, 0.0);
    out.slack_cost = 0;
    for (int r = 0; r < p.n_rows; r++)
        if (coord[r] >= 0) { out.eps[r] = eps_of(r); out.slack_cost += v[coord[r]] * v[coord[r]]; }
    out.cost = objective(T, p, x) + out.slack_cost;
    out.iters = iters; out.n_active = q;
    double st = 0;
    for (int r = 0; r < n; r++) { double s = v[r]; for (int k = 0; k < q; k++) s -= lam[k] * Nact[(size_t)r * cap + k]; st = std::max(st, std::fabs(s)); }
    out.kkt_stationarity = st;
    double mv = 0;
    for (int id = 0; id < n_ids; id++) { Priced pr{0.0, 1.0}; if (!eval(id, pr)) continue; mv = std::max(mv, -pr.slack); }
    out.max_violation = mv;
}

// ---------------------------------------------------------------------------------------------
// Dense assembly in the reference's row order (for the golden comparison with log/QPmodel.lp).
//   P (90x90), qlin (90), c0; Aeq (51x90: 45 Aeq_base rows axis-major, then 6 stop rows), beq;
//   inequality rows "a.x >= b": SFC (if boxes), LSC, dynamic (as <= rows negated).
// ---------------------------------------------------------------------------------------------
struct DenseQp {
    std::vector<double> P, qlin, Aeq, beq, Ain, bin, lb, ub;
    double c0 = 0; int n_in = 0;
};

inline void assemble_dense(const QpTables& T, const QpProblem& p, const float* boxes /*[M][6] or null*/, DenseQp& D) {
    D.P.assign(QNV * QNV, 0); D.qlin.assign(QNV, 0); D.c0 = 0;
    for (int k = 0; k < 3; k++) {
        for (int m = 0; m < QM; m++)
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++)
                D.P[(k * QAX + 6 * m + i) * QNV + k * QAX + 6 * m + j] += T.w * T.Qb[i][j];
        for (int m = QM - p.ts; m < QM; m++) {
            int j = k * QAX + 6 * m + 5;
            D.P[j * QNV + j] += T.wT; D.qlin[j] += -2 * T.wT * p.goal[k]; D.c0 += T.wT * p.goal[k] * p.goal[k];
        }
    }
    D.Aeq.assign(51 * QNV, 0); D.beq.assign(51, 0);
    for (int k = 0; k < 3; k++) {
        for (int r = 0; r < 15; r++) {
            for (int c = 0; c < QAX; c++) D.Aeq[(15 * k + r) * QNV + k * QAX + c] = T.A17[r][c];
            if (r < 3) D.beq[15 * k + r] = p.s[r][k];
        }
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < QAX; c++) D.Aeq[(45 + 2 * k + r) * QNV + k * QAX + c] = T.A17[15 + r][c];
    }
    D.Ain.clear(); D.bin.clear(); D.n_in = 0;
    auto push = [&](const int* idx, const double* a, int nnz, double b) {
        size_t o = D.Ain.size(); D.Ain.resize(o + QNV, 0.0);
        for (int t = 0; t < nnz; t++) D.Ain[o + idx[t]] += a[t];
        D.bin.push_back(b); D.n_in++;
    };
    if (boxes) {
        for (int m = 0; m < QM; m++)
            for (int k = 0; k < 3; k++)
                for (int side = 0; side < 2; side++)
                    for (int j = 0; j < 6; j++) {
                        if (m == 0 && j < QPHI) continue;
                        int idx = k * QAX + 6 * m + j; double a = side == 0 ? 1.0 : -1.0;
                        double b = side == 0 ? (double)boxes[m * 6 + k] : -(double)boxes[m * 6 + 3 + k];
                        push(&idx, &a, 1, b);
                    }
    }
    for (int r = 0; r < p.n_rows; r++)
        for (int i = 0; i < 6; i++) {
            RowView rv; if (!row_get(T, p, 450 + 6 * r + i, rv)) continue;
            push(rv.idx, rv.a, rv.nnz, rv.b);
        }
    for (int k = 0; k < 3; k++) for (int m = 0; m < QM; m++) for (int j = 0; j < 9; j++) for (int side = 0; side < 2; side++) {
        RowView rv; if (!row_get(T, p, 180 + ((k * 5 + m) * 9 + j) * 2 + side, rv)) continue;
        push(rv.idx, rv.a, rv.nnz, rv.b);
    }
    D.lb.assign(p.lb, p.lb + QNV); D.ub.assign(p.ub, p.ub + QNV);
}

}  // namespace orc
