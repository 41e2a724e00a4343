// Host-side FP64 precompute of everything in the trajectory QP that does not depend on the agent:
// built ONCE per engine (the reference builds Q_base / Aeq_base once per TrajOptimizer,
// src/traj_optimizer.cpp:4-25,169-236) and uploaded to the device as read-only tables.
//
// The QP (src/traj_optimizer.cpp:261-539) separates per axis except for the LSC
// rows: 30 variables, 17 equalities (3 initial state, 12 continuity, 2 terminal stop) -> 13 degrees
// of freedom per axis, identical for all axes and agents. For every terminal-segment count
// ts = 1..5 we compute a *whitened* null-space basis G_ts (30 x 13) with  G^T P_ts G = I  and
// A G = 0, so that with x = x0 + (G (+) G (+) G) v the objective is J(x0) + |v|^2 and the QP becomes
// the least-distance problem   min |v|^2  s.t.  (G^T a_j) . v >= b_j - a_j . x0   over the
// inequality rows. x0 (the equality-constrained minimiser) is linear in the initial state and the
// goal: x0 = Xs s + xg g.
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace lscgpu {

constexpr int kM = 5, kN = 5, kNcp = 6, kPhi = 3, kDim = 3;
constexpr int kAx = kM * kNcp;     // 30 variables per axis
constexpr int kNv = kDim * kAx;    // 90
constexpr int kFree = 13;          // degrees of freedom per axis
constexpr int kRed = kDim * kFree; // 39
constexpr int kEq = 17;            // equalities per axis

// Tables as uploaded (plain arrays, one struct = one cudaMemcpy).
struct QpTablesDev {
    double Qw[6][6];               // control_input_weight * Q_base
    double wT;
    double dt;
    double vel_coef, acc_coef;     // n/dt, n(n-1)/dt^2   (src/traj_optimizer.cpp:469-525)
    double G[5][kAx][kFree];       // whitened basis per ts (index ts-1)
    double Xs[5][kAx][3];          // x0 = Xs (pos,vel,acc) + xg * goal
    double xg[5][kAx];
    double gnorm[5][kAx];          // |G row|   (whitened length of a unit bound normal)
    double dyn_norm[5][5][9];      // [ts-1][m][j] whitened length of the velocity (j<5) / acceleration (j>=5) row
};

namespace detail {

inline long binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    long r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
inline long falling(int n, int k) {          // n!/(n-k)!, 0 when n < k   (include/polynomial.hpp:224-234)
    if (n < k) return 0;
    long r = 1;
    for (int i = 0; i < k; i++) r *= (n - i);
    return r;
}

// column-major-free tiny dense matrix
struct Mat {
    int r, c;
    std::vector<double> a;
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};
inline Mat mul(const Mat& A, const Mat& B) {
    Mat C(A.r, B.c);
    for (int i = 0; i < A.r; i++)
        for (int k = 0; k < A.c; k++) {
            double f = A(i, k);
            if (f == 0) continue;
            for (int j = 0; j < B.c; j++) C(i, j) += f * B(k, j);
        }
    return C;
}
inline Mat transpose(const Mat& A) {
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; i++) for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
    return T;
}

// Householder QR of a tall matrix X (rows x cols, rows >= cols): returns the full orthogonal Q (rows x rows)
// and overwrites X with R in its upper triangle.
inline Mat householder_qr(Mat& X) {
    const int m = X.r, n = X.c;
    Mat Q(m, m);
    for (int i = 0; i < m; i++) Q(i, i) = 1.0;
    std::vector<double> u(m);
    for (int k = 0; k < n; k++) {
        double nrm = 0;
        for (int i = k; i < m; i++) nrm += X(i, k) * X(i, k);
        nrm = std::sqrt(nrm);
        if (nrm == 0) throw std::runtime_error("equality rows are rank deficient");
        double alpha = X(k, k) > 0 ? -nrm : nrm;
        for (int i = 0; i < m; i++) u[i] = 0;
        for (int i = k; i < m; i++) u[i] = X(i, k);
        u[k] -= alpha;
        double uu = 0;
        for (int i = k; i < m; i++) uu += u[i] * u[i];
        if (uu == 0) continue;
        for (int j = k; j < n; j++) {
            double s = 0;
            for (int i = k; i < m; i++) s += u[i] * X(i, j);
            s *= 2.0 / uu;
            for (int i = k; i < m; i++) X(i, j) -= s * u[i];
        }
        for (int r = 0; r < m; r++) {          // Q <- Q H
            double s = 0;
            for (int i = k; i < m; i++) s += Q(r, i) * u[i];
            s *= 2.0 / uu;
            for (int i = k; i < m; i++) Q(r, i) -= s * u[i];
        }
    }
    return Q;
}

}  // namespace detail

// Q_base = dt^(1-2 phi) * B Z B^T  (src/traj_optimizer.cpp:169-184; Bernstein basis include/polynomial.hpp:415-428)
inline void jerk_cost_block(double dt, double Q[6][6]) {
    using namespace detail;
    Mat B(6, 6), Z(6, 6);
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) B(i, j) = (double)(binom(kN, i) * binom(kN - i, kN - j)) * (((j - i) & 1) ? -1.0 : 1.0);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            int p = i + j - 2 * kPhi + 1;
            if (p > 0) Z(i, j) = (double)falling(i, kPhi) * (double)falling(j, kPhi) / p;
        }
    Mat Qm = mul(mul(B, Z), transpose(B));
    const double scale = std::pow(dt, 1 - 2 * kPhi);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Q[i][j] = Qm(i, j) * scale;
}

// Per-axis equality matrix (17 x 30): rows 0-2 initial pos/vel/acc, rows 3-14 C0/C1/C2 continuity between
// consecutive segments (src/traj_optimizer.cpp:186-236), rows 15-16 terminal stop c_{4,5} = c_{4,4} = c_{4,3}
// (src/traj_optimizer.cpp:529-536).
inline detail::Mat equality_rows(double dt) {
    using namespace detail;
    Mat A(kEq, kAx);
    // derivative stencils at the start (forward differences) and at the end (backward differences)
    const double st[3][3] = {{1, 0, 0}, {-1, 1, 0}, {1, -2, 1}};
    for (int j = 0; j < kPhi; j++) {
        const double scale = std::pow(dt, -j) * (double)falling(kN, j);
        for (int t = 0; t <= j; t++) A(j, t) = scale * st[j][t];
        for (int m = 1; m < kM; m++) {
            const int row = kPhi * m + j;
            // end of segment m-1: the stencil of order j ends at control point 5
            for (int t = 0; t <= j; t++) A(row, 6 * (m - 1) + 5 - j + t) = scale * st[j][t];
            for (int t = 0; t <= j; t++) A(row, 6 * m + t) = -scale * st[j][t];
        }
    }
    for (int i = 1; i < kPhi; i++) {
        A(14 + i, 6 * (kM - 1) + kN) = 1.0;
        A(14 + i, 6 * (kM - 1) + kN - i) = -1.0;
    }
    return A;
}

inline void build_qp_tables(double dt, double w, double wT, QpTablesDev& T) {
    using namespace detail;
    double Qb[6][6];
    jerk_cost_block(dt, Qb);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) T.Qw[i][j] = w * Qb[i][j];
    T.wT = wT; T.dt = dt;
    T.vel_coef = std::pow(dt, -1) * kN;
    T.acc_coef = std::pow(dt, -2) * kN * (kN - 1);

    Mat A = equality_rows(dt);
    Mat At = transpose(A);                 // 30 x 17
    Mat Q = householder_qr(At);            // At = Q [R; 0]
    // orthonormal null-space basis Z = Q[:, 17:], min-norm particular map Xp = Q1 R^-T E3
    Mat Z(kAx, kFree);
    for (int i = 0; i < kAx; i++) for (int c = 0; c < kFree; c++) Z(i, c) = Q(i, kEq + c);
    Mat Xp(kAx, 3);
    for (int col = 0; col < 3; col++) {
        // solve R^T y = e_col (forward substitution; R upper triangular in At)
        double y[kEq];
        for (int i = 0; i < kEq; i++) {
            double s = (i == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= At(k, i) * y[k];
            y[i] = s / At(i, i);
        }
        for (int r = 0; r < kAx; r++) {
            double s = 0;
            for (int k = 0; k < kEq; k++) s += Q(r, k) * y[k];
            Xp(r, col) = s;
        }
    }
    for (int ts = 1; ts <= kM; ts++) {
        Mat P(kAx, kAx);
        for (int m = 0; m < kM; m++)
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) P(6 * m + i, 6 * m + j) += T.Qw[i][j];
        std::vector<double> eT(kAx, 0.0);
        for (int m = kM - ts; m < kM; m++) { P(6 * m + 5, 6 * m + 5) += wT; eT[6 * m + 5] = 1.0; }
        Mat PZ = mul(P, Z);
        Mat H = mul(transpose(Z), PZ);     // 13 x 13 reduced Hessian (of J = x^T P x, no 1/2)
        for (int a = 0; a < kFree; a++) for (int b = 0; b < a; b++) { double s = 0.5 * (H(a, b) + H(b, a)); H(a, b) = H(b, a) = s; }
        Mat L(kFree, kFree);
        for (int j = 0; j < kFree; j++) {
            double s = H(j, j);
            for (int k = 0; k < j; k++) s -= L(j, k) * L(j, k);
            if (!(s > 0)) throw std::runtime_error("reduced Hessian is not positive definite");
            L(j, j) = std::sqrt(s);
            for (int i = j + 1; i < kFree; i++) {
                double t = H(i, j);
                for (int k = 0; k < j; k++) t -= L(i, k) * L(j, k);
                L(i, j) = t / L(j, j);
            }
        }
        // G = Z L^-T  (row i of G solves L g = Z_i^T)
        Mat G(kAx, kFree);
        for (int i = 0; i < kAx; i++) {
            double nn = 0;
            for (int a = 0; a < kFree; a++) {
                double t = Z(i, a);
                for (int k = 0; k < a; k++) t -= L(a, k) * G(i, k);
                G(i, a) = t / L(a, a);
                nn += G(i, a) * G(i, a);
            }
            T.gnorm[ts - 1][i] = std::sqrt(nn);
            for (int a = 0; a < kFree; a++) T.G[ts - 1][i][a] = G(i, a);
        }
        // x0 = Xp s - G G^T (P Xp s - wT eT g)
        Mat PX = mul(P, Xp);
        Mat GtPX = mul(transpose(G), PX);          // 13 x 3
        Mat corr = mul(G, GtPX);                   // 30 x 3
        for (int i = 0; i < kAx; i++) for (int c = 0; c < 3; c++) T.Xs[ts - 1][i][c] = Xp(i, c) - corr(i, c);
        for (int i = 0; i < kAx; i++) {
            double s = 0;
            for (int a = 0; a < kFree; a++) {
                double ge = 0;
                for (int r = 0; r < kAx; r++) ge += G(r, a) * eT[r];
                s += G(i, a) * ge;
            }
            T.xg[ts - 1][i] = wT * s;
        }
        // whitened lengths of the dynamic-limit rows
        for (int m = 0; m < kM; m++)
            for (int j = 0; j < 9; j++) {
                double nv[kFree];
                for (int a = 0; a < kFree; a++) {
                    if (j < 5) nv[a] = T.vel_coef * (G(6 * m + j + 1, a) - G(6 * m + j, a));
                    else { int i = j - 5; nv[a] = T.acc_coef * (G(6 * m + i + 2, a) - 2 * G(6 * m + i + 1, a) + G(6 * m + i, a)); }
                }
                double nn = 0;
                for (int a = 0; a < kFree; a++) nn += nv[a] * nv[a];
                T.dyn_norm[ts - 1][m][j] = std::sqrt(nn);
            }
    }
}

}  // namespace lscgpu
