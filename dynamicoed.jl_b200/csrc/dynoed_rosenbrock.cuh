// dynoed_rosenbrock.cuh — fixed-grid Rosenbrock integration of the sensitivity-augmented system
// for the stiff / DAE configuration (BASELINE.json config 4; the reference runs Rodas5 on the torn
// system, /root/reference/test/references/Rober.jl:23-31, /root/reference/src/problem.jl:82).
//
// State per design lane: y = [x; vec(G); P] with P_i = int (h_x[i,:]G)^T (h_x[i,:]G) dt the
// UNWEIGHTED per-observation quadratures of one merged interval; F += w_i P_i at the interval end
// (F is linear in w because nothing depends on F, /root/reference/src/augment.jl:223-226), and
// the per-interval P_i are exactly what d objective / d w needs.
//
// One step (Hairer/Wanner IV.7, implementation form without Jacobian-vector products):
//     (1/(h gamma) I - J_y(y_n)) U_i = g(y_n + sum_{j<i} a_ij U_j) + sum_{j<i} (c_ij / h) U_j
//     y_{n+1} = y_n + sum_i m_i U_i
// J_y is block lower triangular with the SAME diagonal block f_x for x and for every column of G
// and a zero diagonal block for P: one nx x nx LU per step (partial
// pivoting, registers) serves all 1 + np right-hand sides; the off-diagonal products
// d(J G_j + f_p,j)/dx . U_x and dq/dy . U are emitted symbolically (Model::ros_ling / ros_linp).
// The reference factors the dense n_aug x n_aug matrix instead.
//
// Gradient: d/dw and d/dtau need no extra integration; d/d(unknown initial conditions) and
// d/d(control values) are carried as a forward-mode dual part through the whole scheme (the way
// the reference's ForwardDiff does, /root/reference/src/problem.jl:154), one integration pass per
// unknown initial condition / control value (controls: Rosenbrock only — the explicit tableaus have
// the adjoint sweep of oed_eval_kernel; with DYNOED_GRAD_NO_CONTROLS their entries are NaN).
#pragma once
#include <type_traits>

#include "dynoed_kernels.cuh"

namespace dynoed {

// ----------------------------------------------------------------------------------------------
// forward-mode dual number with K directions
// ----------------------------------------------------------------------------------------------
template <int K>
struct Dual {
    double v;
    double d[K];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < K; ++i) d[i] = 0.0;
    }
};
#define DYNOED_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < K; ++i)
template <int K> __device__ __forceinline__ Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v + b.v; DYNOED_DUAL_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v - b.v; DYNOED_DUAL_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator-(const Dual<K>& a) { Dual<K> r; r.v = -a.v; DYNOED_DUAL_LOOP r.d[i] = -a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v * b.v; DYNOED_DUAL_LOOP r.d[i] = fma(a.d[i], b.v, a.v * b.d[i]); return r; }
template <int K> __device__ __forceinline__ Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; const double ib = 1.0 / b.v; r.v = a.v * ib; DYNOED_DUAL_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator+(const Dual<K>& a, double b) { Dual<K> r = a; r.v += b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator+(double b, const Dual<K>& a) { Dual<K> r = a; r.v += b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator-(const Dual<K>& a, double b) { Dual<K> r = a; r.v -= b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator-(double b, const Dual<K>& a) { Dual<K> r; r.v = b - a.v; DYNOED_DUAL_LOOP r.d[i] = -a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator*(const Dual<K>& a, double b) { Dual<K> r; r.v = a.v * b; DYNOED_DUAL_LOOP r.d[i] = a.d[i] * b; return r; }
template <int K> __device__ __forceinline__ Dual<K> operator*(double b, const Dual<K>& a) { return a * b; }
template <int K> __device__ __forceinline__ Dual<K> operator/(const Dual<K>& a, double b) { return a * (1.0 / b); }
template <int K> __device__ __forceinline__ Dual<K> operator/(double a, const Dual<K>& b) { Dual<K> r; r.v = a / b.v; const double s = -r.v / b.v; DYNOED_DUAL_LOOP r.d[i] = s * b.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> sqrt(const Dual<K>& a) { Dual<K> r; r.v = ::sqrt(a.v); const double s = 0.5 / r.v; DYNOED_DUAL_LOOP r.d[i] = s * a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> exp(const Dual<K>& a) { Dual<K> r; r.v = ::exp(a.v); DYNOED_DUAL_LOOP r.d[i] = r.v * a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> log(const Dual<K>& a) { Dual<K> r; r.v = ::log(a.v); const double s = 1.0 / a.v; DYNOED_DUAL_LOOP r.d[i] = s * a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> sin(const Dual<K>& a) { Dual<K> r; r.v = ::sin(a.v); const double s = ::cos(a.v); DYNOED_DUAL_LOOP r.d[i] = s * a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> cos(const Dual<K>& a) { Dual<K> r; r.v = ::cos(a.v); const double s = -::sin(a.v); DYNOED_DUAL_LOOP r.d[i] = s * a.d[i]; return r; }
template <int K> __device__ __forceinline__ Dual<K> pow(const Dual<K>& a, double e) { Dual<K> r; r.v = ::pow(a.v, e); const double s = e * ::pow(a.v, e - 1.0); DYNOED_DUAL_LOOP r.d[i] = s * a.d[i]; return r; }
#undef DYNOED_DUAL_LOOP
__device__ __forceinline__ double value_of(double x) { return x; }
template <int K> __device__ __forceinline__ double value_of(const Dual<K>& x) { return x.v; }
__device__ __forceinline__ double dual_part(double, int) { return 0.0; }
template <int K> __device__ __forceinline__ double dual_part(const Dual<K>& x, int k) { return x.d[k]; }

// ----------------------------------------------------------------------------------------------
// Rosenbrock tableaus in implementation form (gamma, a_ij, c_ij, m_i)
// ----------------------------------------------------------------------------------------------
// ROS2 (Verwer, Spee, Blom, Hundsdorfer 1999): 2 stages, order 2, L-stable, gamma = 1 + 1/sqrt(2):
//   (I - gamma h J) k1 = f(y),  (I - gamma h J) k2 = f(y + h k1) - 2 k1,  y+ = y + 1.5 h k1 + 0.5 h k2
// with U_i = gamma h k_i.
struct TabROS2 {
    static constexpr int S = 2;
    static constexpr double gamma = 1.7071067811865475244;
    __host__ __device__ static constexpr double a(int i, int j) { return (i == 1 && j == 0) ? 1.0 / gamma : 0.0; }
    __host__ __device__ static constexpr double c(int i, int j) { return (i == 1 && j == 0) ? -2.0 / gamma : 0.0; }
    __host__ __device__ static constexpr double m(int i) { return i == 0 ? 1.5 / gamma : 0.5 / gamma; }
    __host__ __device__ static constexpr bool same_stage_state(int) { return false; }
};
// ROS3P (Lang, Verwer 2001): 3 stages, order 3, A-stable (R(inf) = 0.73), two function evaluations
// (stages 2 and 3 share their stage state).
struct TabROS3P {
    static constexpr int S = 3;
    static constexpr double gamma = 7.886751345948129e-01;
    __host__ __device__ static constexpr double a(int i, int j) {
        return (i == 1 && j == 0) ? 1.267949192431123 : (i == 2 && j == 0) ? 1.267949192431123 : 0.0;
    }
    __host__ __device__ static constexpr double c(int i, int j) {
        return (i == 1 && j == 0) ? -1.607695154586736
             : (i == 2 && j == 0) ? -3.464101615137755
             : (i == 2 && j == 1) ? -1.732050807568877 : 0.0;
    }
    __host__ __device__ static constexpr double m(int i) {
        return i == 0 ? 2.0 : i == 1 ? 5.773502691896258e-01 : 4.226497308103742e-01;
    }
    __host__ __device__ static constexpr bool same_stage_state(int i) { return i == 2; }
};

// RODAS3 (Sandu, Verwer, Blom, Spee, Carmichael, Potra 1997): 4 stages, order 3, stiffly accurate
// and L-stable (R(inf) = 0), three function evaluations (stages 1 and 2 share their stage state).
struct TabRODAS3 {
    static constexpr int S = 4;
    static constexpr double gamma = 0.5;
    __host__ __device__ static constexpr double a(int i, int j) {
        return (i == 2 && j == 0) ? 2.0 : (i == 3 && j == 0) ? 2.0 : (i == 3 && j == 2) ? 1.0 : 0.0;
    }
    __host__ __device__ static constexpr double c(int i, int j) {
        return (i == 1 && j == 0) ? 4.0
             : (i == 2 && j == 0) ? 1.0 : (i == 2 && j == 1) ? -1.0
             : (i == 3 && j == 0) ? 1.0 : (i == 3 && j == 1) ? -1.0 : (i == 3 && j == 2) ? -8.0 / 3.0 : 0.0;
    }
    __host__ __device__ static constexpr double m(int i) { return i == 0 ? 2.0 : i == 1 ? 0.0 : 1.0; }
    __host__ __device__ static constexpr bool same_stage_state(int i) { return i == 1; }
};

// ----------------------------------------------------------------------------------------------
// LU with partial pivoting of a small matrix held in registers.
//
// The COMMON case needs no row exchange at all: W = 1/(h gamma) I - f_x is column diagonally dominant
// for small steps and for kinetics of any step (f_x has non-positive column sums), and partial pivoting
// never exchanges rows of such a matrix.  So the factorisation runs SPECULATIVELY without exchanges and
// only records, per lane, whether partial pivoting would have picked another row at some elimination
// step (the column maxima it needs are by-products of the elimination).  If no lane of the warp needed
// one — one vote per factorisation — the result IS the pivoted factorisation and every solve is plain
// straight-line substitution.  Otherwise the warp rebuilds the matrix and takes the pivoted path:
// row exchanges as SELECTS over statically indexed rows (dynamic register indexing would move the whole
// matrix to local memory), whole rows as in LAPACK's getrf, so that a solve applies the exchanges to its
// right-hand side first (laswp) and then runs the same substitution code.  The vote only decides which
// (equivalent) code runs: results do not depend on which lanes share a warp.
//
// (Round 2's first version voted at every elimination step and guarded the exchanges of every solve
// step: ptxas paid for those ~50 warp-uniform branches per Rosenbrock step with phi copies of the matrix
// at every join — 470 IMAD.MOV + 160 FSEL of 1,780 executed instructions per chem6 ROS3P step.)
// ----------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void select_swap(bool sw, T& x, T& y) {
    const T tx = x, ty = y;
    x = sw ? ty : tx;
    y = sw ? tx : ty;
}

template <class T, int N>
struct SmallLU {
    T a[N][N];         // L (unit diagonal, multipliers below) and U; the diagonal keeps 1 / pivot
    int piv[N];        // pivoted path: row exchanged with row k at elimination step k
    bool pivoted;      // warp-uniform: a[][] factorises the row-exchanged matrix

    // elimination step k on rows k+1.. (shared by both paths)
    template <int k>
    __device__ __forceinline__ void eliminate() {
        const T inv = 1.0 / a[k][k];
        a[k][k] = inv;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const T l = a[i][k] * inv;
            a[i][k] = l;
#pragma unroll
            for (int j = k + 1; j < N; ++j) a[i][j] = a[i][j] - l * a[k][j];
        }
    }
    // no exchanges; returns whether partial pivoting would have exchanged rows at some step (then a[][] is
    // not to be used: the caller rebuilds the matrix and calls factor_pivoted)
    template <int k>
    __device__ __forceinline__ bool plain_step() {
        bool need = false;
        const double d = fabs(value_of(a[k][k]));
#pragma unroll
        for (int i = k + 1; i < N; ++i) need = need || (fabs(value_of(a[i][k])) > d);
        eliminate<k>();
        if constexpr (k + 1 < N) need = plain_step<k + 1>() || need;
        return need;
    }
    __device__ __forceinline__ bool factor_plain() {
        pivoted = false;
        return plain_step<0>();
    }
    template <int k>
    __device__ __forceinline__ void pivoted_step() {
        int p = k;
        double best = fabs(value_of(a[k][k]));
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double v = fabs(value_of(a[i][k]));
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const bool sw = (p == i);
#pragma unroll
            for (int j = 0; j < N; ++j) select_swap(sw, a[k][j], a[i][j]);   // whole rows (getrf)
        }
        eliminate<k>();
        if constexpr (k + 1 < N) pivoted_step<k + 1>();
    }
    __device__ __forceinline__ void factor_pivoted() {
        pivoted = true;
        pivoted_step<0>();
    }
    __device__ __forceinline__ void solve(T (&b)[N]) const {
        if (pivoted) {   // warp-uniform, rare
#pragma unroll
            for (int k = 0; k + 1 < N; ++k)
#pragma unroll
                for (int i = k + 1; i < N; ++i) select_swap(piv[k] == i, b[k], b[i]);
        }
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = b[i] - a[i][k] * b[k];
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            T s = b[k];
#pragma unroll
            for (int j = k + 1; j < N; ++j) s = s - a[k][j] * b[j];
            b[k] = s * a[k][k];
        }
    }
};

// W = c0 I - A  ->  lu: speculative plain factorisation, pivoted redo when some lane of the warp needs it.
// `jac(A)` fills A = f_x row-major; it is evaluated AGAIN on the rare path instead of keeping NX^2 more
// doubles alive across the elimination.
template <int NX, class Jac>
__device__ __forceinline__ void factor_shifted(SmallLU<double, NX>& lu, double c0, Jac jac) {
    {
        double A[NX * NX];
        jac(A);
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) lu.a[i][j] = (i == j ? c0 : 0.0) - A[i * NX + j];
    }
    const bool need = lu.factor_plain();
    if (__any_sync(__activemask(), need)) {
        double A[NX * NX];
        jac(A);
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) lu.a[i][j] = (i == j ? c0 : 0.0) - A[i * NX + j];
        lu.factor_pivoted();
    }
}

// ----------------------------------------------------------------------------------------------
// The linear systems of one step:  W U = b  with  W = 1/(h gamma) I - f_x(x_n).
// Plain doubles: one pivoted LU.  Forward-mode duals (gradient directions): the factorisation is
// still the VALUE factorisation W0; a direction's part follows from differentiating the system,
//     W0 U1 = b1 - W1 U0 = b1 + (d f_x . direction) U0 ,
// with the second-derivative product emitted sparsely (Model::ros_hvp).  A dual-number LU would
// triple the factorisation work and double its registers (36 more doubles for nx = 6).
// ----------------------------------------------------------------------------------------------
template <class M, class T>
struct StepSolver;

template <class M>
struct StepSolver<M, double> {
    SmallLU<double, M::NX> lu;
    __device__ __forceinline__ void init(double t, double c0, const double* z, const double* u, const double* th,
                                         const double* c) {
        constexpr int NX = M::NX;
        factor_shifted<NX>(lu, c0, [&](double (&A)[NX * NX]) { M::template ros_jac<double>(t, z, u, th, c, A); });
    }
    __device__ __forceinline__ void solve(double (&b)[M::NX]) const { lu.solve(b); }
};

template <class M, int K>
struct StepSolver<M, Dual<K>> {
    static constexpr int NX = M::NX, NCU = M::NCTRL > 0 ? M::NCTRL : 1;
    SmallLU<double, NX> lu;
    double xv[NX], xd[K][NX], uv[NCU], ud[K][NCU], cv[M::NCONST];
    double t;
    const double* th;
    __device__ __forceinline__ void init(double t_, double c0, const Dual<K>* z, const Dual<K>* u, const double* th_,
                                         const Dual<K>* c) {
        t = t_; th = th_;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            xv[i] = z[i].v;
#pragma unroll
            for (int d = 0; d < K; ++d) xd[d][i] = z[i].d[d];
        }
#pragma unroll
        for (int k = 0; k < M::NCTRL; ++k) {
            uv[k] = u[k].v;
#pragma unroll
            for (int d = 0; d < K; ++d) ud[d][k] = u[k].d[d];
        }
#pragma unroll
        for (int k = 0; k < M::NCONST; ++k) cv[k] = c[k].v;
        factor_shifted<NX>(lu, c0, [&](double (&A)[NX * NX]) { M::template ros_jac<double>(t, xv, uv, th, cv, A); });
    }
    __device__ __forceinline__ void solve(Dual<K> (&b)[NX]) const {
        double b0[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) b0[i] = b[i].v;
        lu.solve(b0);
#pragma unroll
        for (int d = 0; d < K; ++d) {
            double hv[NX], b1[NX];
            M::ros_hvp(t, xv, xd[d], uv, ud[d], th, cv, b0, hv);
#pragma unroll
            for (int i = 0; i < NX; ++i) b1[i] = b[i].d[d] + hv[i];
            lu.solve(b1);
#pragma unroll
            for (int i = 0; i < NX; ++i) b[i].d[d] = b1[i];
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) b[i].v = b0[i];
    }
};

// ----------------------------------------------------------------------------------------------
// One Rosenbrock step on y = [z (x, G); Pq (unweighted quadratures)]
// ----------------------------------------------------------------------------------------------
template <class M, class RT, class T>
__device__ __forceinline__ void rosenbrock_step(double t, double h, double ih, T (&z)[M::NZ], T (&Pq)[M::NOBS * M::NF],
                                                const T* u, const double* th, const T* c) {
    constexpr int S = RT::S, NX = M::NX, NP = M::NP, NZ = M::NZ, NQ = M::NOBS * M::NF;
    // ih = 1 / h comes from the host's step table: two FP64 divisions per step less (for a 2-state model like
    // Robertson they were ~15 % of the step's instructions)
    const double c0 = ih * (1.0 / RT::gamma);
    StepSolver<M, T> lu;
    lu.init(t, c0, z, u, th, c);
    T Uz[S][NZ], Uq[S][NQ];
    T gz[NZ], gq[NQ];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (!RT::same_stage_state(s)) {
            T Z[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) Z[a] = z[a];
#pragma unroll
            for (int j = 0; j < s; ++j) {
                if (RT::a(s, j) != 0.0) {
#pragma unroll
                    for (int a = 0; a < NZ; ++a) Z[a] = Z[a] + RT::a(s, j) * Uz[j][a];
                }
            }
            M::template ros_f<T>(t, Z, u, th, c, gz, gq);
        }
        T rz[NZ], rq[NQ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) rz[a] = gz[a];
#pragma unroll
        for (int a = 0; a < NQ; ++a) rq[a] = gq[a];
#pragma unroll
        for (int j = 0; j < s; ++j) {
            if (RT::c(s, j) != 0.0) {
                const double cj = RT::c(s, j) * ih;
#pragma unroll
                for (int a = 0; a < NZ; ++a) rz[a] = rz[a] + cj * Uz[j][a];
#pragma unroll
                for (int a = 0; a < NQ; ++a) rq[a] = rq[a] + cj * Uq[j][a];
            }
        }
        // x block
        T bx[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) bx[i] = rz[i];
        lu.solve(bx);
#pragma unroll
        for (int i = 0; i < NX; ++i) Uz[s][i] = bx[i];
        // G columns: (c0 I - f_x) U_Gj = r_Gj + B_j U_x
        T BU[NX * NP];
        M::template ros_ling<T>(t, z, u, th, c, Uz[s], BU);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T bg[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) bg[i] = rz[NX + j * NX + i] + BU[j * NX + i];
            lu.solve(bg);
#pragma unroll
            for (int i = 0; i < NX; ++i) Uz[s][NX + j * NX + i] = bg[i];
        }
        // quadrature rows: c0 U_P = r_P + dq/dy . U
        T dPU[NQ];
        M::template ros_linp<T>(t, z, u, th, c, Uz[s], dPU);
        const double ic0 = h * RT::gamma;
#pragma unroll
        for (int a = 0; a < NQ; ++a) Uq[s][a] = (rq[a] + dPU[a]) * ic0;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (RT::m(s) != 0.0) {
#pragma unroll
            for (int a = 0; a < NZ; ++a) z[a] = z[a] + RT::m(s) * Uz[s][a];
#pragma unroll
            for (int a = 0; a < NQ; ++a) Pq[a] = Pq[a] + RT::m(s) * Uq[s][a];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Steppers for the forward-only kernel below: Rosenbrock (stiff configuration) or an explicit
// Runge-Kutta tableau on the same state y = [z; Pq] (the "criterion + d/dw + d/dtau" mode of the
// explicit integrators: no backward sweep, controls are not differentiated).
// ----------------------------------------------------------------------------------------------
template <class RT>
struct RosenbrockStepper {
    static constexpr int STAGES = RT::S;
    template <class M, class T>
    __device__ __forceinline__ static void step(double t, double h, double ih, T (&z)[M::NZ], T (&Pq)[M::NOBS * M::NF],
                                                const T* u, const double* th, const T* c) {
        rosenbrock_step<M, RT, T>(t, h, ih, z, Pq, u, th, c);
    }
};

template <class Tab>
struct ExplicitStepper {
    static constexpr int STAGES = Tab::S;
    template <class M, class T>
    __device__ __forceinline__ static void step(double t, double h, double /*ih*/, T (&z)[M::NZ], T (&Pq)[M::NOBS * M::NF],
                                                const T* u, const double* th, const T* c) {
        constexpr int S = Tab::S, NZ = M::NZ, NQ = M::NOBS * M::NF;
        T k[S][NZ], zacc[NZ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) zacc[a] = z[a];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            T Z[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) Z[a] = z[a];
#pragma unroll
            for (int j = 0; j < s; ++j) {
                if (Tab::a(s, j) != 0.0) {
                    const double ha = h * Tab::a(s, j);
#pragma unroll
                    for (int a = 0; a < NZ; ++a) Z[a] = Z[a] + ha * k[j][a];
                }
            }
            T dq[NQ];
            M::template ros_f<T>(t + Tab::c(s) * h, Z, u, th, c, k[s], dq);
            const double hb = h * Tab::b(s);
#pragma unroll
            for (int a = 0; a < NZ; ++a) zacc[a] = zacc[a] + hb * k[s][a];
#pragma unroll
            for (int a = 0; a < NQ; ++a) Pq[a] = Pq[a] + hb * dq[a];
        }
#pragma unroll
        for (int a = 0; a < NZ; ++a) z[a] = zacc[a];
    }
};

// ----------------------------------------------------------------------------------------------
// The forward-only evaluation kernel (Rosenbrock, or explicit RK without control gradients): same
// tile staging (TMA), epilogue and arg-min as oed_eval_kernel; no backward sweep (see the header
// comment for the gradient).  Control entries of the gradient are NaN ("not computed").
// P.traj is used as the per-interval quadrature scratch [grid][N][NOBS*NF][kMaxTile].
// ----------------------------------------------------------------------------------------------
template <class M, class Stepper, bool GRAD, bool CTRLGRAD, int NDIR>
__global__ void __launch_bounds__(kMaxTile, 1) ros_eval_kernel(const EvalParams P) {
    constexpr int NX = M::NX, NP = M::NP, NZ = M::NZ, NF = M::NF, NOBS = M::NOBS, NCTRL = M::NCTRL;
    constexpr int NCU = NCTRL > 0 ? NCTRL : 1, NQ = NOBS * NF;
    // gradient directions (unknown initial conditions, and with CTRLGRAD every control VALUE of the design
    // vector) are carried as forward-mode dual parts, NDIR directions per integration pass.  The linear
    // systems of every direction are solved with the value factorisation (StepSolver): a pass costs about
    // 1 + NDIR units of a plain integration.
    constexpr bool CG = GRAD && CTRLGRAD && M::NCTRL > 0;
    constexpr bool DUAL = GRAD && (M::NIC > 0 || CG);
    using T = typename std::conditional<DUAL, Dual<NDIR>, double>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* tile_buf = reinterpret_cast<double*>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int TILE = P.tile;
    const int n_var = P.n_var;
    const long long ntiles = (P.n + TILE - 1) / TILE;
    const int N = P.n_intervals;
    const double* th = P.theta;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see oed_eval_kernel / launch_device
    if (tid == 0 && P.use_tma) mbar_init(bar, 1);
    scratch_acquire(P);
    uint32_t phase = 0;
    double best_val = __longlong_as_double(0x7ff0000000000000LL);
    long long best_idx = -1;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long g0 = tile * TILE;
        const int count = (int)((P.n - g0 < TILE) ? (P.n - g0) : TILE);
        const double* src = P.designs + g0 * n_var;
        const uint32_t bytes = (uint32_t)count * n_var * 8u;
        const bool tma_ok = P.use_tma && (bytes % 16u == 0);
        if (tma_ok) {
            if (tid == 0) {
                mbar_expect_tx(bar, bytes);
                tma_load_1d(tile_buf, src, bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            for (int i = tid; i < count * n_var; i += blockDim.x) tile_buf[i] = src[i];
            __syncthreads();
        }

        if (tid < count) {
            double* row = tile_buf + (size_t)tid * n_var;
            const long long g = g0 + tid;
            // per-lane scratch: [N][NQ] interval quadratures, then one slot per control value
            double* pq = P.traj + ((size_t)blockIdx.x * (N * NQ + P.n_ctrl_values)) * kMaxTile + tid;
            double* cg = pq + (size_t)N * NQ * kMaxTile;
            double Ls[NF], L[NF], icg[M::NIC > 0 ? M::NIC : 1];
            const int ndir = DUAL ? M::NIC + (CG ? P.n_ctrl_values : 0) : 1;
            // direction -> (control k, entry idx) for dir >= NIC: controls are laid out block after block
            for (int dir0 = 0; dir0 < (ndir > 0 ? ndir : 1); dir0 += NDIR) {
                int dk[NDIR], didx[NDIR];
#pragma unroll
                for (int q = 0; q < NDIR; ++q) {
                    dk[q] = -1; didx[q] = -1;
                    if (CG && dir0 + q >= M::NIC && dir0 + q < ndir) {
                        int rest = dir0 + q - M::NIC;
#pragma unroll
                        for (int k = 0; k < NCTRL; ++k) {
                            if (dk[q] < 0) {
                                if (rest < P.ctrl_len[k]) { dk[q] = k; didx[q] = rest; }
                                else rest -= P.ctrl_len[k];
                            }
                        }
                    }
                }
                T z[NZ], F[NF];
#pragma unroll
                for (int a = 0; a < NX; ++a) z[a] = T(P.x0[a]);
#pragma unroll
                for (int a = 0; a < NX * NP; ++a) z[NX + a] = T(P.G0[a]);
                if constexpr (M::NIC > 0) {
#pragma unroll
                    for (int a = 0; a < M::NIC; ++a) {
                        T v(row[P.ic_off + a]);
                        if constexpr (GRAD) {
#pragma unroll
                            for (int q = 0; q < NDIR; ++q) v.d[q] = (a == dir0 + q) ? 1.0 : 0.0;
                        }
                        z[M::ic_state(a)] = v;
                    }
                }
#pragma unroll
                for (int a = 0; a < NF; ++a) F[a] = T(0.0);
                double* xg = (!GRAD && P.xgrid && dir0 == 0) ? P.xgrid + (size_t)g * (N + 1) * (NZ + NF) : nullptr;
                if (xg) {
#pragma unroll
                    for (int a = 0; a < NZ; ++a) xg[a] = value_of(z[a]);
#pragma unroll
                    for (int a = 0; a < NF; ++a) xg[NZ + a] = 0.0;
                }
                for (int j = 0; j < N; ++j) {
                    T u[NCU], c[M::NCONST];
                    double w[NOBS];
#pragma unroll
                    for (int k = 0; k < NCTRL; ++k) {
                        const int idx = __ldg(P.indicators + k * N + j);
                        u[k] = T(row[P.ctrl_off[k] + idx]);
                        if constexpr (CG) {
#pragma unroll
                            for (int q = 0; q < NDIR; ++q) u[k].d[q] = (k == dk[q] && idx == didx[q]) ? 1.0 : 0.0;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NOBS; ++i) w[i] = row[P.w_off[i] + __ldg(P.indicators + (NCTRL + i) * N + j)];
                    M::template consts_t<T>(u, th, c);
                    T Pq[NQ];
#pragma unroll
                    for (int a = 0; a < NQ; ++a) Pq[a] = T(0.0);
                    const int s0 = __ldg(P.first_step + j), s1 = __ldg(P.first_step + j + 1);
                    for (int s = s0; s < s1; ++s)
                        Stepper::template step<M, T>(__ldg(P.step_t + s), __ldg(P.step_h + s), __ldg(P.step_ih + s), z, Pq, u, th, c);
#pragma unroll
                    for (int i = 0; i < NOBS; ++i)
#pragma unroll
                        for (int k = 0; k < NF; ++k) F[k] = F[k] + w[i] * Pq[i * NF + k];
                    if (GRAD && dir0 == 0) {
#pragma unroll
                        for (int a = 0; a < NQ; ++a) pq[((size_t)j * NQ + a) * kMaxTile] = value_of(Pq[a]);
                    }
                    if (xg) {
                        double* o = xg + (size_t)(j + 1) * (NZ + NF);
#pragma unroll
                        for (int a = 0; a < NZ; ++a) o[a] = value_of(z[a]);
#pragma unroll
                        for (int a = 0; a < NF; ++a) o[NZ + a] = value_of(F[a]);
                    }
                }

                if (dir0 == 0) {
                    // ---- criterion epilogue ----
                    const double tau = row[P.tau_off];
                    double Fv[NF], value;
#pragma unroll
                    for (int a = 0; a < NF; ++a) Fv[a] = value_of(F[a]);
                    criterion_eval<NP>(Fv, tau, P.criterion, value, L);
                    const double obj = value + tau;
                    P.crit[g] = obj;
                    if (P.fim) {
#pragma unroll
                        for (int a = 0; a < NF; ++a) P.fim[g * NF + a] = Fv[a];
                    }
                    if (obj < best_val) { best_val = obj; best_idx = P.index_base + g; }
#pragma unroll
                    for (int jj = 0; jj < NP; ++jj)
#pragma unroll
                        for (int ii = 0; ii <= jj; ++ii) Ls[tri(ii, jj)] = (ii == jj ? 1.0 : 2.0) * L[tri(ii, jj)];
                }
                if constexpr (DUAL) {   // <Lambda, dF/d(direction)> with the symmetric pairing
#pragma unroll
                    for (int q = 0; q < NDIR; ++q) {
                        double sd = 0.0;
#pragma unroll
                        for (int k = 0; k < NF; ++k) sd = fma(Ls[k], dual_part(F[k], q), sd);
#pragma unroll
                        for (int a = 0; a < M::NIC; ++a)
                            if (a == dir0 + q) icg[a] = sd;          // static indexing keeps icg in registers
                        if (CG && dir0 + q >= M::NIC && dir0 + q < ndir) cg[(size_t)(dir0 + q - M::NIC) * kMaxTile] = sd;
                    }
                }
            }

            // ---- gradient: <Lambda, .> with the symmetric pairing (off-diagonals count twice) ----
            if (GRAD) {
                double wb[NOBS];
                int wk[NOBS];
#pragma unroll
                for (int i = 0; i < NOBS; ++i) { wb[i] = 0.0; wk[i] = -1; }
                for (int j = 0; j < N; ++j) {
#pragma unroll
                    for (int i = 0; i < NOBS; ++i) {
                        const int idx = __ldg(P.indicators + (NCTRL + i) * N + j);
                        if (idx != wk[i]) {
                            if (wk[i] >= 0) row[P.w_off[i] + wk[i]] = wb[i];
                            wb[i] = 0.0; wk[i] = idx;
                        }
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < NF; ++k) s = fma(Ls[k], pq[((size_t)j * NQ + i * NF + k) * kMaxTile], s);
                        wb[i] += s;
                    }
                }
#pragma unroll
                for (int i = 0; i < NOBS; ++i)
                    if (wk[i] >= 0) row[P.w_off[i] + wk[i]] = wb[i];
#pragma unroll
                for (int a = 0; a < M::NIC; ++a) row[P.ic_off + a] = icg[a];
                // controls: one dual pass per value (CTRLGRAD), otherwise "not computed" (NaN)
                {
                    int d = 0;
#pragma unroll
                    for (int k = 0; k < NCTRL; ++k)
                        for (int idx = 0; idx < P.ctrl_len[k]; ++idx, ++d)
                            row[P.ctrl_off[k] + idx] = CG ? cg[(size_t)d * kMaxTile]
                                                          : __longlong_as_double(0x7ff8000000000000LL);
                }
                double trL = 1.0;
#pragma unroll
                for (int i = 0; i < NP; ++i) trL += L[tri(i, i)];
                row[P.tau_off] = trL;
            }
        }

        if (GRAD) {
            double* dst = P.grad + g0 * n_var;
            if (tma_ok) {
                fence_proxy_async();
                __syncthreads();
                if (tid == 0) {
                    tma_store_1d(dst, tile_buf, bytes);
                    tma_store_wait_read();
                }
                __syncthreads();
            } else {
                __syncthreads();
                for (int i = tid; i < count * n_var; i += blockDim.x) dst[i] = tile_buf[i];
                __syncthreads();
            }
        } else {
            __syncthreads();
        }
    }
    argmin_tail(P, best_val, best_idx);
}

// ----------------------------------------------------------------------------------------------
// Rosenbrock gradient kernel for UNKNOWN INITIAL CONDITIONS (BASELINE.json config 4: stiff kinetics,
// ~6 states, 12 sensitivities, Fisher criteria).  d objective / d ic_a needs dF/d ic_a, i.e. the
// second-order sensitivities H_(j,a) = dG_j/d ic_a.  Differentiating a Rosenbrock scheme that uses
// the exact Jacobian gives the same scheme on the system extended by H and dP = dP/d ic (emit.py,
// "second-order pieces"), whose Jacobian is still block lower triangular with diagonal block f_x:
// every H column is one more right-hand side for the step's ONE value factorisation.  Compared
// with forward-mode dual numbers through the whole step this (a) does not carry d x/d ic_a twice
// (it IS a G column), (b) uses H_(j,a) = H_(a,j), (c) evaluates only the non-zero second derivatives.
//
// Registers cannot hold the extended state and its S stage increments (chem6: 54 + 54 S doubles), so
// the step is ordered by COLUMN and only what later columns need is kept, in shared memory:
//   phase 1  (x, G, P) stage by stage as in ros_eval_kernel; the increments U_s of (x, G) go to
//            shared memory  [S][NZ][tile]
//   phase 2  one H column at a time through all S stages: stage states of (x, G) are rebuilt from y_n
//            (registers) and the stored increments; the column's own increments live in registers for
//            the duration of the column; its share of the dP rows is accumulated (the stage recursion of
//            the quadrature rows is linear, so each column applies it to its own contribution)
//   phase 3  y_n <- y_n + sum_s m_s U_s
// Shared memory per lane: (S + NH/NP... ) see Ic2Layout; each lane touches only its own slots, so the
// phases need no barrier.
// ----------------------------------------------------------------------------------------------
// Lane stride of the shared-memory slots = lanes per CTA of this kernel.  Four-stage tableaus (RODAS3) need
// S NZ + NH NX slots per lane (chem6: 90 doubles = 92 KB at 128 lanes, one CTA per SM); with 64-lane CTAs
// three or four fit.  The host picks the matching tile (dynoed_create).
template <class RT>
__host__ __device__ constexpr int ic_stride() { return RT::S >= 4 ? 64 : kMaxTile; }

template <class M, class RT>
struct Ic2Layout {
    static constexpr int S = RT::S;
    static constexpr int U = 0;                          // [S][NZ] increments of (x, G)
    static constexpr int HN = U + S * M::NZ;             // [NH][NX] second-order sensitivities at y_n
    static constexpr int SLOTS = HN + M::NH * M::NX;     // doubles per lane
};

// Shared-memory reads of the stage storage are VOLATILE: otherwise the compiler forwards the values it has
// just stored from registers, keeps every increment live and spills them to local memory — the opposite
// of what the shared-memory staging is for.
template <class RT, int NZ>
struct StageZ {                       // Z_s[k] = y_n[k] + sum_j a_sj U_j[k]
    const double (&z)[NZ];
    const volatile double* U;
    int s;
    __device__ __forceinline__ double operator[](int k) const {
        double v = z[k];
#pragma unroll
        for (int j = 0; j < RT::S; ++j)
            if (j < s && RT::a(s, j) != 0.0) v = fma(RT::a(s, j), U[(size_t)(j * NZ + k) * ic_stride<RT>()], v);
        return v;
    }
};
template <class RT, int NZ>
struct StageU {
    const volatile double* U;
    int s;
    __device__ __forceinline__ double operator[](int k) const { return U[(size_t)(s * NZ + k) * ic_stride<RT>()]; }
};

template <class M, class RT, int Q>
struct Ic2Column {
    __device__ __forceinline__ static void run(double t, double ih, double ic0, const StepSolver<M, double>& lu,
                                               const double (&z)[M::NZ], double (&dP)[M::NOBS * M::NF * M::NIC],
                                               double* __restrict__ sm, const double* u, const double* th) {
        constexpr int S = RT::S, NX = M::NX, NZ = M::NZ, NQ2 = M::NOBS * M::NF * M::NIC;
        constexpr unsigned long long rows = M::ic2_qrows(Q);
        using Lay = Ic2Layout<M, RT>;
        const volatile double* Us = sm + (size_t)Lay::U * ic_stride<RT>();
        volatile double* Hs = sm + (size_t)(Lay::HN + Q * NX) * ic_stride<RT>();
        double Hn[NX], UH[S][NX], Ud[S][NQ2];
#pragma unroll
        for (int i = 0; i < NX; ++i) Hn[i] = Hs[(size_t)i * ic_stride<RT>()];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const StageZ<RT, NZ> zs{z, Us, s};
            const StageU<RT, NZ> us{Us, s};
            double ZH[NX], r[NX], b[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double v = Hn[i];
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (RT::a(s, j) != 0.0) v = fma(RT::a(s, j), UH[j][i], v);
                ZH[i] = v;
            }
            M::template ic2_rhs<Q>(t, zs, ZH, u, th, r);
            M::template ic2_lin<Q>(t, z, Hn, us, u, th, b);
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double v = r[i] + b[i];
#pragma unroll
                for (int j = 0; j < s; ++j)
                    if (RT::c(s, j) != 0.0) v = fma(RT::c(s, j) * ih, UH[j][i], v);
                r[i] = v;
            }
            lu.solve(r);
#pragma unroll
            for (int i = 0; i < NX; ++i) UH[s][i] = r[i];
            if constexpr (rows != 0ull) {
                double acc[NQ2];
#pragma unroll
                for (int a = 0; a < NQ2; ++a) acc[a] = 0.0;
                M::template ic2_qcol<Q>(t, zs, ZH, z, Hn, us, UH[s], u, th, acc);
#pragma unroll
                for (int a = 0; a < NQ2; ++a) {
                    if (rows >> a & 1ull) {
                        double v = acc[a];
#pragma unroll
                        for (int j = 0; j < s; ++j)
                            if (RT::c(s, j) != 0.0) v = fma(RT::c(s, j) * ih, Ud[j][a], v);
                        Ud[s][a] = v * ic0;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = Hn[i];
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (RT::m(s) != 0.0) v = fma(RT::m(s), UH[s][i], v);
            Hs[(size_t)i * ic_stride<RT>()] = v;
        }
#pragma unroll
        for (int a = 0; a < NQ2; ++a) {
            if (rows >> a & 1ull) {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (RT::m(s) != 0.0) dP[a] = fma(RT::m(s), Ud[s][a], dP[a]);
            }
        }
        if constexpr (Q + 1 < M::NH) Ic2Column<M, RT, Q + 1>::run(t, ih, ic0, lu, z, dP, sm, u, th);
    }
};

template <class M, class RT>
__device__ __forceinline__ void rosenbrock_ic_step(double t, double h, double ih, double (&z)[M::NZ], double (&Pq)[M::NOBS * M::NF],
                                                   double (&dP)[M::NOBS * M::NF * M::NIC], double* __restrict__ sm,
                                                   const double* u, const double* th, const double* c) {
    constexpr int S = RT::S, NX = M::NX, NP = M::NP, NZ = M::NZ, NQ = M::NOBS * M::NF, NQ2 = NQ * M::NIC;
    using Lay = Ic2Layout<M, RT>;
    const double c0 = ih * (1.0 / RT::gamma);
    const double ic0 = h * RT::gamma;
    volatile double* Us = sm + (size_t)Lay::U * ic_stride<RT>();
    StepSolver<M, double> lu;
    lu.init(t, c0, z, u, th, c);
    // ---- phase 1: (x, G, P) ----
    double Uq[S][NQ];
    double gz[NZ], gq[NQ];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (!RT::same_stage_state(s)) {
            const StageZ<RT, NZ> zs{z, Us, s};
            double Z[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) Z[a] = zs[a];
            M::template ros_f<double>(t, Z, u, th, c, gz, gq);
        }
        double rz[NZ], rq[NQ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) rz[a] = gz[a];
#pragma unroll
        for (int a = 0; a < NQ; ++a) rq[a] = gq[a];
#pragma unroll
        for (int j = 0; j < s; ++j) {
            if (RT::c(s, j) != 0.0) {
                const double cj = RT::c(s, j) * ih;
#pragma unroll
                for (int a = 0; a < NZ; ++a) rz[a] = fma(cj, Us[(size_t)(j * NZ + a) * ic_stride<RT>()], rz[a]);
#pragma unroll
                for (int a = 0; a < NQ; ++a) rq[a] = fma(cj, Uq[j][a], rq[a]);
            }
        }
        double Uc[NZ];
        {
            double bx[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) bx[i] = rz[i];
            lu.solve(bx);
#pragma unroll
            for (int i = 0; i < NX; ++i) Uc[i] = bx[i];
        }
        double BU[NX * NP];
        M::template ros_ling<double>(t, z, u, th, c, Uc, BU);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            double bg[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) bg[i] = rz[NX + j * NX + i] + BU[j * NX + i];
            lu.solve(bg);
#pragma unroll
            for (int i = 0; i < NX; ++i) Uc[NX + j * NX + i] = bg[i];
        }
        double dPU[NQ];
        M::template ros_linp<double>(t, z, u, th, c, Uc, dPU);
#pragma unroll
        for (int a = 0; a < NQ; ++a) Uq[s][a] = (rq[a] + dPU[a]) * ic0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) Us[(size_t)(s * NZ + a) * ic_stride<RT>()] = Uc[a];
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (RT::m(s) != 0.0) {
#pragma unroll
            for (int a = 0; a < NQ; ++a) Pq[a] = fma(RT::m(s), Uq[s][a], Pq[a]);
        }
    }
    // ---- phase 2: dP rows that do not involve H, then the H columns ----
    if constexpr (M::IC2_QBASE_ROWS != 0ull) {
        double Ub[S][NQ2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const StageZ<RT, NZ> zs{z, Us, s};
            const StageU<RT, NZ> us{Us, s};
            double rb[NQ2];
            M::ic2_qbase(t, zs, z, us, u, th, rb);
#pragma unroll
            for (int a = 0; a < NQ2; ++a) {
                if (M::IC2_QBASE_ROWS >> a & 1ull) {
                    double v = rb[a];
#pragma unroll
                    for (int j = 0; j < s; ++j)
                        if (RT::c(s, j) != 0.0) v = fma(RT::c(s, j) * ih, Ub[j][a], v);
                    Ub[s][a] = v * ic0;
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NQ2; ++a) {
            if (M::IC2_QBASE_ROWS >> a & 1ull) {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (RT::m(s) != 0.0) dP[a] = fma(RT::m(s), Ub[s][a], dP[a]);
            }
        }
    }
    Ic2Column<M, RT, 0>::run(t, ih, ic0, lu, z, dP, sm, u, th);
    // ---- phase 3 ----
#pragma unroll
    for (int a = 0; a < NZ; ++a) {
        double v = z[a];
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (RT::m(s) != 0.0) v = fma(RT::m(s), Us[(size_t)(s * NZ + a) * ic_stride<RT>()], v);
        z[a] = v;
    }
}

template <class M, class RT>
__global__ void __launch_bounds__(kMaxTile, 1) ros_ic_kernel(const EvalParams P) {
    constexpr int NX = M::NX, NP = M::NP, NZ = M::NZ, NF = M::NF, NOBS = M::NOBS, NCTRL = M::NCTRL, NIC = M::NIC;
    constexpr int NCU = NCTRL > 0 ? NCTRL : 1, NQ = NOBS * NF, NQ2 = NQ * NIC;
    using Lay = Ic2Layout<M, RT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* tile_buf = reinterpret_cast<double*>(smem_raw + 128);
    const int tid = threadIdx.x;
    const int TILE = P.tile;
    const int n_var = P.n_var;
    const long long ntiles = (P.n + TILE - 1) / TILE;
    const int N = P.n_intervals;
    const double* th = P.theta;
    // this lane's column of the stage / second-order storage, behind the design tile
    double* sm = tile_buf + (((size_t)TILE * n_var + 15) & ~(size_t)15) + tid;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see oed_eval_kernel / launch_device
    if (tid == 0 && P.use_tma) mbar_init(bar, 1);
    scratch_acquire(P);
    uint32_t phase = 0;
    double best_val = __longlong_as_double(0x7ff0000000000000LL);
    long long best_idx = -1;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long g0 = tile * TILE;
        const int count = (int)((P.n - g0 < TILE) ? (P.n - g0) : TILE);
        const double* src = P.designs + g0 * n_var;
        const uint32_t bytes = (uint32_t)count * n_var * 8u;
        const bool tma_ok = P.use_tma && (bytes % 16u == 0);
        if (tma_ok) {
            if (tid == 0) {
                mbar_expect_tx(bar, bytes);
                tma_load_1d(tile_buf, src, bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            for (int i = tid; i < count * n_var; i += blockDim.x) tile_buf[i] = src[i];
            __syncthreads();
        }

        if (tid < count) {
            double* row = tile_buf + (size_t)tid * n_var;
            const long long g = g0 + tid;
            double* pq = P.traj + ((size_t)blockIdx.x * (N * NQ + P.n_ctrl_values)) * kMaxTile + tid;
            double z[NZ], F[NF], Fd[NIC][NF];
#pragma unroll
            for (int a = 0; a < NX; ++a) z[a] = P.x0[a];
#pragma unroll
            for (int a = 0; a < NX * NP; ++a) z[NX + a] = P.G0[a];
#pragma unroll
            for (int a = 0; a < NIC; ++a) z[M::ic_state(a)] = row[P.ic_off + a];
#pragma unroll
            for (int a = 0; a < NF; ++a) F[a] = 0.0;
#pragma unroll
            for (int a = 0; a < NIC; ++a)
#pragma unroll
                for (int k = 0; k < NF; ++k) Fd[a][k] = 0.0;
#pragma unroll
            for (int a = 0; a < M::NH * NX; ++a) sm[(size_t)(Lay::HN + a) * ic_stride<RT>()] = 0.0;   // G0 is constant
            double* const xg = nullptr;      // gradient kernel: the trajectory comes from the objective-only kernels
            if (xg) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) xg[a] = z[a];
#pragma unroll
                for (int a = 0; a < NF; ++a) xg[NZ + a] = 0.0;
            }
            for (int j = 0; j < N; ++j) {
                double u[NCU], c[M::NCONST], w[NOBS];
#pragma unroll
                for (int k = 0; k < NCTRL; ++k) u[k] = row[P.ctrl_off[k] + __ldg(P.indicators + k * N + j)];
#pragma unroll
                for (int i = 0; i < NOBS; ++i) w[i] = row[P.w_off[i] + __ldg(P.indicators + (NCTRL + i) * N + j)];
                M::consts(u, th, c);
                double Pq[NQ], dP[NQ2];
#pragma unroll
                for (int a = 0; a < NQ; ++a) Pq[a] = 0.0;
#pragma unroll
                for (int a = 0; a < NQ2; ++a) dP[a] = 0.0;
                const int s0 = __ldg(P.first_step + j), s1 = __ldg(P.first_step + j + 1);
                for (int s = s0; s < s1; ++s)
                    rosenbrock_ic_step<M, RT>(__ldg(P.step_t + s), __ldg(P.step_h + s), __ldg(P.step_ih + s), z, Pq, dP, sm, u, th, c);
#pragma unroll
                for (int i = 0; i < NOBS; ++i)
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        F[k] = fma(w[i], Pq[i * NF + k], F[k]);
#pragma unroll
                        for (int a = 0; a < NIC; ++a) Fd[a][k] = fma(w[i], dP[a * NQ + i * NF + k], Fd[a][k]);
                    }
#pragma unroll
                for (int a = 0; a < NQ; ++a) pq[((size_t)j * NQ + a) * kMaxTile] = Pq[a];
                if (xg) {
                    double* o = xg + (size_t)(j + 1) * (NZ + NF);
#pragma unroll
                    for (int a = 0; a < NZ; ++a) o[a] = z[a];
#pragma unroll
                    for (int a = 0; a < NF; ++a) o[NZ + a] = F[a];
                }
            }

            // ---- criterion epilogue ----
            const double tau = row[P.tau_off];
            double value, L[NF], Ls[NF];
            criterion_eval<NP>(F, tau, P.criterion, value, L);
            const double obj = value + tau;
            P.crit[g] = obj;
            if (P.fim) {
#pragma unroll
                for (int a = 0; a < NF; ++a) P.fim[g * NF + a] = F[a];
            }
            if (obj < best_val) { best_val = obj; best_idx = P.index_base + g; }
#pragma unroll
            for (int jj = 0; jj < NP; ++jj)
#pragma unroll
                for (int ii = 0; ii <= jj; ++ii) Ls[tri(ii, jj)] = (ii == jj ? 1.0 : 2.0) * L[tri(ii, jj)];

            // ---- gradient: <Lambda, .> with the symmetric pairing (off-diagonals count twice) ----
            double wb[NOBS];
            int wk[NOBS];
#pragma unroll
            for (int i = 0; i < NOBS; ++i) { wb[i] = 0.0; wk[i] = -1; }
            for (int j = 0; j < N; ++j) {
#pragma unroll
                for (int i = 0; i < NOBS; ++i) {
                    const int idx = __ldg(P.indicators + (NCTRL + i) * N + j);
                    if (idx != wk[i]) {
                        if (wk[i] >= 0) row[P.w_off[i] + wk[i]] = wb[i];
                        wb[i] = 0.0; wk[i] = idx;
                    }
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NF; ++k) s = fma(Ls[k], pq[((size_t)j * NQ + i * NF + k) * kMaxTile], s);
                    wb[i] += s;
                }
            }
#pragma unroll
            for (int i = 0; i < NOBS; ++i)
                if (wk[i] >= 0) row[P.w_off[i] + wk[i]] = wb[i];
#pragma unroll
            for (int a = 0; a < NIC; ++a) {
                double sd = 0.0;
#pragma unroll
                for (int k = 0; k < NF; ++k) sd = fma(Ls[k], Fd[a][k], sd);
                row[P.ic_off + a] = sd;
            }
#pragma unroll
            for (int k = 0; k < NCTRL; ++k)       // controls: "not computed" (the CTRLGRAD kernels do them)
                for (int idx = 0; idx < P.ctrl_len[k]; ++idx)
                    row[P.ctrl_off[k] + idx] = __longlong_as_double(0x7ff8000000000000LL);
            double trL = 1.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) trL += L[tri(i, i)];
            row[P.tau_off] = trL;
        }

        double* dst = P.grad + g0 * n_var;
        if (tma_ok) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tma_store_1d(dst, tile_buf, bytes);
                tma_store_wait_read();
            }
            __syncthreads();
        } else {
            __syncthreads();
            for (int i = tid; i < count * n_var; i += blockDim.x) dst[i] = tile_buf[i];
            __syncthreads();
        }
    }
    argmin_tail(P, best_val, best_idx);
}

// ----------------------------------------------------------------------------------------------
// Direction-parallel gradient kernel for SMALL batches — the optimiser callback (one design per
// call, /root/reference/src/problem.jl:115-120 and its ForwardDiff gradient :154) and the
// 2 n_var + 1 designs of a finite-difference Hessian.  A batch that cannot fill the GPU is bound by
// the sequential depth of one design (forward sweep + adjoint sweep), not by arithmetic, so the
// work of ONE design is spread over 1 + n_dir lanes: lane 0 integrates the plain system and keeps
// the per-interval quadratures (objective, F, d/dw, d/dtau); lane 1 + k carries a forward-mode dual
// part for direction k (an unknown initial condition or one control value) through the same sweep
// and writes that single gradient entry.  No backward sweep, no checkpoints; every lane of a warp
// runs the same instruction stream (the base lane is a dual lane with a zero seed).  Designs are
// read and gradients written straight from/to global memory (a few KB per call).
// P.traj is used as [n][N][NOBS*NF] quadrature scratch of the base lanes.
// ----------------------------------------------------------------------------------------------
template <class M, class Stepper>
__global__ void __launch_bounds__(32, 1) dir_eval_kernel(const EvalParams P) {
    constexpr int NX = M::NX, NP = M::NP, NZ = M::NZ, NF = M::NF, NOBS = M::NOBS, NCTRL = M::NCTRL;
    constexpr int NCU = NCTRL > 0 ? NCTRL : 1, NQ = NOBS * NF;
    using T = Dual<1>;
    const int n_var = P.n_var, N = P.n_intervals;
    const int nl = 1 + M::NIC + P.n_ctrl_values;          // lanes per design
    const long long lane = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long g = lane / nl;
    const int dir = (int)(lane - g * nl) - 1;             // -1: base lane
    const double* th = P.theta;
    double best_val = __longlong_as_double(0x7ff0000000000000LL);
    long long best_idx = -1;
    scratch_acquire(P);

    // P.tile > 0: the design rows of this CTA's lanes are staged in shared memory first (one
    // coalesced read — the rows may live in mapped host memory, see eval_host's small-call path)
    extern __shared__ __align__(16) double dir_rows[];
    const long long d_first = (blockIdx.x * (long long)blockDim.x) / nl;
    // ... and so are the interval tables (indicator matrix, first step of every interval): a lone
    // warp has nothing to hide a global load behind
    const int* ind = P.indicators;
    const int* fst = P.first_step;
    if (P.tile > 0) {
        long long d_last = (blockIdx.x * (long long)blockDim.x + blockDim.x - 1) / nl;
        if (d_last > P.n - 1) d_last = P.n - 1;
        const long long cnt = (d_last - d_first + 1) * n_var;
        const double* src = P.designs + d_first * n_var;
        for (long long i = threadIdx.x; i < cnt; i += blockDim.x) dir_rows[i] = src[i];
        int* s_tab = reinterpret_cast<int*>(dir_rows + (size_t)P.tile * (n_var + N * NQ));
        const int n_ind = (NCTRL + NOBS) * N;
        for (int i = threadIdx.x; i < n_ind; i += blockDim.x) s_tab[i] = P.indicators[i];
        for (int i = threadIdx.x; i <= N; i += blockDim.x) s_tab[n_ind + i] = P.first_step[i];
        ind = s_tab; fst = s_tab + n_ind;
        __syncthreads();
    }

    if (g < P.n) {
        const double* row = P.tile > 0 ? dir_rows + (g - d_first) * n_var : P.designs + g * n_var;
        double* grow = P.grad + g * n_var;
        int dk = -1, didx = -1;                           // direction -> (control k, entry idx)
        if (dir >= M::NIC) {
            int rest = dir - M::NIC;
#pragma unroll
            for (int k = 0; k < NCTRL; ++k) {
                if (dk < 0) {
                    if (rest < P.ctrl_len[k]) { dk = k; didx = rest; }
                    else rest -= P.ctrl_len[k];
                }
            }
        }
        T z[NZ], F[NF];
#pragma unroll
        for (int a = 0; a < NX; ++a) z[a] = T(P.x0[a]);
#pragma unroll
        for (int a = 0; a < NX * NP; ++a) z[NX + a] = T(P.G0[a]);
        if constexpr (M::NIC > 0) {
#pragma unroll
            for (int a = 0; a < M::NIC; ++a) {
                T v(row[P.ic_off + a]);
                v.d[0] = (a == dir) ? 1.0 : 0.0;
                z[M::ic_state(a)] = v;
            }
        }
#pragma unroll
        for (int a = 0; a < NF; ++a) F[a] = T(0.0);
        // per-interval quadratures of the base lane: shared memory next to the rows, else global scratch
        double* pq = P.tile > 0 ? dir_rows + (size_t)P.tile * n_var + (size_t)(g - d_first) * N * NQ
                                : P.traj + (size_t)g * N * NQ;
        int s_cur = fst[0];
        double t_next = __ldg(P.step_t + s_cur), h_next = __ldg(P.step_h + s_cur), ih_next = __ldg(P.step_ih + s_cur);
        for (int j = 0; j < N; ++j) {
            T u[NCU], c[M::NCONST];
            double w[NOBS];
#pragma unroll
            for (int k = 0; k < NCTRL; ++k) {
                const int idx = ind[k * N + j];
                u[k] = T(row[P.ctrl_off[k] + idx]);
                u[k].d[0] = (k == dk && idx == didx) ? 1.0 : 0.0;
            }
#pragma unroll
            for (int i = 0; i < NOBS; ++i) w[i] = row[P.w_off[i] + ind[(NCTRL + i) * N + j]];
            M::template consts_t<T>(u, th, c);
            T Pq[NQ];
#pragma unroll
            for (int a = 0; a < NQ; ++a) Pq[a] = T(0.0);
            // one warp per SM sub-partition has nothing to hide a load behind: (t, h) of the NEXT step
            // are fetched before the current step's arithmetic (the step tables are padded by one entry)
            const int s1 = fst[j + 1];
            for (; s_cur < s1; ++s_cur) {
                const double t_now = t_next, h_now = h_next, ih_now = ih_next;
                t_next = __ldg(P.step_t + s_cur + 1); h_next = __ldg(P.step_h + s_cur + 1);
                ih_next = __ldg(P.step_ih + s_cur + 1);
                Stepper::template step<M, T>(t_now, h_now, ih_now, z, Pq, u, th, c);
            }
#pragma unroll
            for (int i = 0; i < NOBS; ++i)
#pragma unroll
                for (int k = 0; k < NF; ++k) F[k] = F[k] + w[i] * Pq[i * NF + k];
            if (dir < 0) {
#pragma unroll
                for (int a = 0; a < NQ; ++a) pq[(size_t)j * NQ + a] = Pq[a].v;
            }
        }
        // every lane evaluates the criterion on its own (identical) F: no exchange between lanes
        const double tau = row[P.tau_off];
        double Fv[NF], L[NF], Ls[NF], value;
#pragma unroll
        for (int a = 0; a < NF; ++a) Fv[a] = F[a].v;
        criterion_eval<NP>(Fv, tau, P.criterion, value, L);
#pragma unroll
        for (int jj = 0; jj < NP; ++jj)
#pragma unroll
            for (int ii = 0; ii <= jj; ++ii) Ls[tri(ii, jj)] = (ii == jj ? 1.0 : 2.0) * L[tri(ii, jj)];
        if (dir >= 0) {
            double sd = 0.0;
#pragma unroll
            for (int k = 0; k < NF; ++k) sd = fma(Ls[k], F[k].d[0], sd);
            if (dir < M::NIC) grow[P.ic_off + dir] = sd;
            else grow[P.ctrl_off[dk] + didx] = sd;
        } else {
            const double obj = value + tau;
            P.crit[g] = obj;
            if (P.fim) {
#pragma unroll
                for (int a = 0; a < NF; ++a) P.fim[g * NF + a] = Fv[a];
            }
            if (obj < best_val) { best_val = obj; best_idx = P.index_base + g; }   // NaN never wins
            double wb[NOBS];
            int wk[NOBS];
#pragma unroll
            for (int i = 0; i < NOBS; ++i) { wb[i] = 0.0; wk[i] = -1; }
            for (int j = 0; j < N; ++j) {
#pragma unroll
                for (int i = 0; i < NOBS; ++i) {
                    const int idx = ind[(NCTRL + i) * N + j];
                    if (idx != wk[i]) {
                        if (wk[i] >= 0) grow[P.w_off[i] + wk[i]] = wb[i];
                        wb[i] = 0.0; wk[i] = idx;
                    }
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < NF; ++k) sacc = fma(Ls[k], pq[(size_t)j * NQ + i * NF + k], sacc);
                    wb[i] += sacc;
                }
            }
#pragma unroll
            for (int i = 0; i < NOBS; ++i)
                if (wk[i] >= 0) grow[P.w_off[i] + wk[i]] = wb[i];
            double trL = 1.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) trL += L[tri(i, i)];
            grow[P.tau_off] = trL;
        }
    }
    argmin_tail(P, best_val, best_idx);
}

// ----------------------------------------------------------------------------------------------
// Time-parallel gradient kernel for the SMALLEST batches — one CTA per design.
//
// dir_eval_kernel above spreads the gradient DIRECTIONS of one design over lanes; every lane still
// walks all n_steps steps with a dual number, so one call costs n_steps x (a dual-number step).
// Here the sequential part is cut down to what is sequential by nature:
//   phase 1 (one lane)        the plain forward sweep x_0 .. x_n of the STATES only (doubles); nothing
//                             else is nonlinear in time;
//   phase 1b                  the sensitivities G: given x_k a step maps G affinely, so task (k, q) gets
//                             column q of that map from one dual-number step, and the chain of small affine
//                             maps is evaluated in chunks of steps (compose per chunk, chain the chunks,
//                             replay) — z_k = (x_k, G_k) at every step start ends up in shared memory;
//   phase 2 (all warps)       the steps are now INDEPENDENT: task (k, m) pushes one dual number through
//                             step k alone, seeded in state component m (or in the control value active
//                             on step k), which yields column m of the step Jacobian d z_{k+1} / d (z_k, u),
//                             the quadrature increment of step k and its derivative;
//   phase 2b                  interval quadratures, F (lanes over intervals + butterfly), the criterion and
//                             Lambda = d criterion / d F; then every task contracts its quadrature
//                             derivative with Lambda and the measurement weights (s_k);
//   phase 3                   the discrete adjoint recursion lambda_k = s_k + J_k^T lambda_{k+1} over these
//                             small matrices, itself cut into <= 16 chunks of steps (one lane alone needs
//                             ~200 cycles per step):  (1) every chunk composes its steps into one affine map
//                             lambda_lo = b_c + A_c lambda_hi — one lane per column of [A_c | b_c], no
//                             communication;  (2) one warp chains the chunk maps (<= 16 small mat-vecs) and
//                             gets lambda at every chunk boundary, lambda_0 = d/d(initial conditions);
//                             (3) models with controls: one lane per chunk replays its steps from the known
//                             boundary value and adds lambda_{k+1} . d z_{k+1}/d u + s_k,u into its own
//                             partial of d/d(control value); partials are summed in a fixed order.
//                             Warp 1 writes d/dw and d/dtau meanwhile.
// The adjoint of the step map is exact for the discrete scheme (same derivative as the dual sweep of
// dir_eval_kernel / the reference's ForwardDiff gradient, /root/reference/src/problem.jl:154), for
// explicit and Rosenbrock steppers alike — the step is differentiated as a black box.
// Shared memory: row | step tables | z_k | step and interval quadratures | J columns | quadrature
// derivatives | s | Lambda | int tables.
// ----------------------------------------------------------------------------------------------
#ifndef DYNOED_TP_BLOCK
#define DYNOED_TP_BLOCK 256
#endif
constexpr int kTpBlock = DYNOED_TP_BLOCK;
// MEASUREMENT-ONLY build switch DYNOED_TP_TIMING: thread 0 overwrites the first gradient entries of its
// design with the SM-cycle counts of the phases (scripts/tp_phase_probe.py); results are WRONG.
#ifdef DYNOED_TP_TIMING
#define DYNOED_TP_MARK(i) { if (threadIdx.x == 0) tp_clk[i] = clock64(); __syncwarp(); }
// BAR.SYNC.DEFER_BLOCKING lets the clock read that follows it issue before the warp blocks: a barrier
// whose result is consumed does block
#define DYNOED_TP_SYNC() { tp_sink += __syncthreads_count(1); }
#else
#define DYNOED_TP_MARK(i)
#define DYNOED_TP_SYNC() __syncthreads();
#endif
constexpr int kTpMaxChunks = 16;

// SPILL: the two large per-step tables (Jacobian columns, quadrature derivatives) live in the launch's global
// scratch (L2) instead of shared memory — models / grids whose tables exceed one SM's shared memory
// (chem6 on a 130-step grid: 18 x 18 columns per step; LV with 16 substeps per interval).
template <class M>
__host__ __device__ inline size_t tp_spill_doubles(int n_steps) {
    return (size_t)n_steps * (M::NZ + M::NCTRL) * (M::NZ + M::NOBS * M::NF);
}
template <class M>
__host__ __device__ inline size_t tp_smem_bytes(int n_var, int N, int n_steps, int n_ctrl_values, bool spill) {
    constexpr int NZ = M::NZ, NQ = M::NOBS * M::NF, ND = M::NZ + M::NCTRL;
    size_t dbl = (size_t)n_var + 3 * (size_t)n_steps + (size_t)(n_steps + 1) * NZ + (size_t)n_steps * NQ +
                 (size_t)N * NQ + (spill ? 0 : tp_spill_doubles<M>(n_steps)) + (size_t)n_steps * ND + M::NF + 2 +
                 (size_t)kTpMaxChunks * (NZ + 1) * NZ + (size_t)(kTpMaxChunks + 1) * NZ +
                 (size_t)kTpMaxChunks * (n_ctrl_values > 0 ? n_ctrl_values : 1);
    size_t ints = (size_t)(M::NCTRL + M::NOBS) * N + (N + 1) + n_steps;
    return dbl * sizeof(double) + ints * sizeof(int) + 16;
}

template <class M, class Stepper, bool SPILL>
__global__ void __launch_bounds__(kTpBlock, 1) tp_eval_kernel(const EvalParams P) {
    constexpr int NX = M::NX, NP = M::NP, NZ = M::NZ, NF = M::NF, NOBS = M::NOBS, NCTRL = M::NCTRL;
    constexpr int NCU = NCTRL > 0 ? NCTRL : 1, NQ = NOBS * NF, ND = NZ + NCTRL;
    using T = Dual<1>;
    const int n_var = P.n_var, N = P.n_intervals, n_steps = P.n_steps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long g = blockIdx.x;
    const double* th = P.theta;
#ifdef DYNOED_TP_TIMING
    long long tp_clk[7];
    int tp_sink = 0;
    const unsigned long long tp_ns0 = global_ns();
#endif
    DYNOED_TP_MARK(0)
    __shared__ int tp_prog;            // steps whose start state x_k is in shared memory (phase 1 -> phase 1b)
    if (threadIdx.x == 0) tp_prog = 0;
    if constexpr (SPILL) scratch_acquire(P);      // the spilled tables belong to the scratch set

    extern __shared__ __align__(16) double tp_sm[];
    double* row = tp_sm;
    double* s_t = row + n_var;
    double* s_h = s_t + n_steps;
    double* s_ih = s_h + n_steps;
    double* zs = s_ih + n_steps;                         // [n_steps + 1][NZ] state at every step start
    double* pqs = zs + (size_t)(n_steps + 1) * NZ;       // [n_steps][NQ]     quadrature increment of every step
    double* pq = pqs + (size_t)n_steps * NQ;             // [N][NQ]           interval quadratures
    double* big = SPILL ? P.traj + (size_t)blockIdx.x * tp_spill_doubles<M>(n_steps) : pq + (size_t)N * NQ;
    double* jc = big;                                    // [n_steps][ND][NZ] step Jacobian columns
    double* dpq = jc + (size_t)n_steps * ND * NZ;        // [n_steps][ND][NQ] derivative of the step's quadrature increment
    double* sc = SPILL ? pq + (size_t)N * NQ : dpq + (size_t)n_steps * ND * NQ;   // [n_steps][ND] ... contracted with Lambda and the weights
    double* ls = sc + (size_t)n_steps * ND;              // [NF] (+ trace term)
    double* comp = ls + NF + 2;                          // [chunks][NZ + 1][NZ] columns of [A_c | b_c]
    double* lamb = comp + (size_t)kTpMaxChunks * (NZ + 1) * NZ;          // [chunks + 1][NZ] lambda at chunk boundaries
    double* part = lamb + (size_t)(kTpMaxChunks + 1) * NZ;               // [chunks][n_ctrl_values] partial d/d(control value)
    const int ncv = P.n_ctrl_values > 0 ? P.n_ctrl_values : 1;
    int* ind = reinterpret_cast<int*>(part + (size_t)kTpMaxChunks * ncv);   // [(NCTRL + NOBS)][N]
    int* fst = ind + (NCTRL + NOBS) * N;                 // [N + 1]
    int* siv = fst + (N + 1);                            // [n_steps] interval of every step

    double* grow = P.grad + g * n_var;
    {
        const double* src = P.designs + g * n_var;
        for (int i = tid; i < n_var; i += blockDim.x) row[i] = src[i];
        for (int i = tid; i < n_steps; i += blockDim.x) {
            s_t[i] = P.step_t[i]; s_h[i] = P.step_h[i]; s_ih[i] = P.step_ih[i];
        }
        for (int i = tid; i < (NCTRL + NOBS) * N; i += blockDim.x) ind[i] = P.indicators[i];
        for (int i = tid; i <= N; i += blockDim.x) fst[i] = P.first_step[i];
        for (int j = warp; j < N; j += (int)(blockDim.x >> 5)) {
            const int a = P.first_step[j], b = P.first_step[j + 1];
            for (int s = a + lane; s < b; s += 32) siv[s] = j;
        }
    }
    DYNOED_TP_SYNC()
    DYNOED_TP_MARK(1)

    // ---- phase 1: plain forward sweep of the STATES x, one lane.  The stepper runs on the whole z, but only
    //      x is stored: the sensitivities and quadratures feed nothing that is read, so their arithmetic is
    //      dead code here (the compiler removes it) ----
    constexpr int NG = NX * NP;                                  // sensitivities
    if (tid == 0) {
        double z[NZ];
#pragma unroll
        for (int a = 0; a < NX; ++a) z[a] = P.x0[a];
#pragma unroll
        for (int a = 0; a < NG; ++a) z[NX + a] = P.G0[a];
        if constexpr (M::NIC > 0) {
#pragma unroll
            for (int a = 0; a < M::NIC; ++a) z[M::ic_state(a)] = row[P.ic_off + a];
        }
        int s_cur = fst[0];
        for (int j = 0; j < N; ++j) {
            double u[NCU], c[M::NCONST];
#pragma unroll
            for (int k = 0; k < NCTRL; ++k) u[k] = row[P.ctrl_off[k] + ind[k * N + j]];
            M::template consts_t<double>(u, th, c);
            const int s1 = fst[j + 1];
            for (; s_cur < s1; ++s_cur) {
#pragma unroll
                for (int a = 0; a < NX; ++a) zs[(size_t)s_cur * NZ + a] = z[a];
                double Pq[NQ];
#pragma unroll
                for (int a = 0; a < NQ; ++a) Pq[a] = 0.0;
                Stepper::template step<M, double>(s_t[s_cur], s_h[s_cur], s_ih[s_cur], z, Pq, u, th, c);
            }
            __threadfence_block();
            *(volatile int*)&tp_prog = s1;      // the tasks of phase 1b start on these steps while the sweep goes on
        }
    }
    DYNOED_TP_MARK(2)

    // ---- phase 1b: the sensitivities.  Given x_k, one step maps G affinely: G_{k+1} = psi_k + Phi_k G_k
    //      (G' = f_x G + f_p, /root/reference/src/augment.jl:231; true for the explicit and the
    //      Rosenbrock steps alike).  Task (k, q) takes the step from (x_k, G = 0) with a dual seed in
    //      sensitivity q: the dual part is column q of Phi_k (exact: the map is affine), the value part is
    //      psi_k.  The chain of these small affine maps is then evaluated in chunks like phase 3, forward. ----
    constexpr int NCG = NG + 1;                                   // columns of [Phi | psi]
    constexpr int LGG = NCG <= 8 ? 8 : NCG <= 16 ? 16 : 32;
    constexpr int CPWG = 32 / LGG;
    static_assert(NCG <= 32, "one lane per column of a chunk map");
    double* phi = jc;                                            // [n_steps][NG + 1][NG]  (jc is written in phase 2)
    double* compg = comp;                                        // [chunks][NG + 1][NG]
    double* gb = lamb;                                           // [chunks + 1][NG] G at the chunk boundaries
    // The tasks run WHILE lane 0 is still sweeping (they only need x_k): every warp except warp 0 and the one
    // that shares its scheduler (warp 4) takes tasks in step order and waits for the sweep's progress counter.
    const int nwork = ((int)(blockDim.x >> 5) - (blockDim.x > 128 ? 2 : 1)) * 32;
    const int wtid = warp == 0 || (warp == 4 && blockDim.x > 128) ? -1 : (warp - 1 - (warp > 4 ? 1 : 0)) * 32 + lane;
    for (int task = wtid; wtid >= 0 && task < n_steps * NCG; task += nwork) {
        const int k = task / NCG, q = task - k * NCG;
        {
            const int klast = min((task - lane + 31) / NCG, n_steps - 1);     // warp-uniform
            while (*(volatile int*)&tp_prog <= klast) __nanosleep(64);
            __threadfence_block();
        }
        const int j = siv[k];
        T z[NZ], u[NCU], c[M::NCONST], Pq[NQ];
#pragma unroll
        for (int a = 0; a < NX; ++a) z[a] = T(zs[(size_t)k * NZ + a]);
#pragma unroll
        for (int a = 0; a < NG; ++a) { z[NX + a].v = 0.0; z[NX + a].d[0] = (a == q) ? 1.0 : 0.0; }
#pragma unroll
        for (int kk = 0; kk < NCTRL; ++kk) u[kk] = T(row[P.ctrl_off[kk] + ind[kk * N + j]]);
        M::template consts_t<T>(u, th, c);
#pragma unroll
        for (int a = 0; a < NQ; ++a) Pq[a] = T(0.0);
        Stepper::template step<M, T>(s_t[k], s_h[k], s_ih[k], z, Pq, u, th, c);
#pragma unroll
        for (int a = 0; a < NG; ++a) phi[(size_t)task * NG + a] = (q < NG) ? z[NX + a].d[0] : z[NX + a].v;
    }
    DYNOED_TP_SYNC()
    const int nchunk_g = min(min(kTpMaxChunks, CPWG * (int)(blockDim.x >> 5)), n_steps);
    {   // compose the steps of every chunk, forward in time
        const int c = warp * CPWG + lane / LGG, q = lane % LGG;
        if (c < nchunk_g && q < NCG) {
            double v[NG];
#pragma unroll
            for (int a = 0; a < NG; ++a) v[a] = (a == q) ? 1.0 : 0.0;
            const int lo = (int)((long long)c * n_steps / nchunk_g), hi = (int)((long long)(c + 1) * n_steps / nchunk_g);
            const double sw = (q == NG) ? 1.0 : 0.0;
            for (int k = lo; k < hi; ++k) {
                const double* ph = phi + (size_t)k * NCG * NG;
                double nv[NG];
#pragma unroll
                for (int a = 0; a < NG; ++a) nv[a] = sw * ph[NG * NG + a];
#pragma unroll
                for (int r = 0; r < NG; ++r)
#pragma unroll
                    for (int a = 0; a < NG; ++a) nv[a] = fma(ph[r * NG + a], v[r], nv[a]);
#pragma unroll
                for (int a = 0; a < NG; ++a) v[a] = nv[a];
            }
#pragma unroll
            for (int a = 0; a < NG; ++a) compg[((size_t)c * NCG + q) * NG + a] = v[a];
        }
    }
    DYNOED_TP_SYNC()
    if (warp == 0) {   // chain the chunk maps from G_0: gb[c] = G at the lower end of chunk c
        double g_a = lane < NG ? P.G0[lane] : 0.0;
        for (int c = 0; c < nchunk_g; ++c) {
            if (lane < NG) gb[(size_t)c * NG + lane] = g_a;
            __syncwarp();
            if (lane < NG) {
                double acc = compg[((size_t)c * NCG + NG) * NG + lane];
#pragma unroll
                for (int r = 0; r < NG; ++r) acc = fma(compg[((size_t)c * NCG + r) * NG + lane], gb[(size_t)c * NG + r], acc);
                g_a = acc;
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane < nchunk_g) {   // replay every chunk from its boundary value: G_k at every step start
            const int c = lane;
            double G[NG];
#pragma unroll
            for (int a = 0; a < NG; ++a) G[a] = gb[(size_t)c * NG + a];
            const int lo = (int)((long long)c * n_steps / nchunk_g), hi = (int)((long long)(c + 1) * n_steps / nchunk_g);
            for (int k = lo; k < hi; ++k) {
                const double* ph = phi + (size_t)k * NCG * NG;
                double nv[NG];
#pragma unroll
                for (int a = 0; a < NG; ++a) { zs[(size_t)k * NZ + NX + a] = G[a]; nv[a] = ph[NG * NG + a]; }
#pragma unroll
                for (int r = 0; r < NG; ++r)
#pragma unroll
                    for (int a = 0; a < NG; ++a) nv[a] = fma(ph[r * NG + a], G[r], nv[a]);
#pragma unroll
                for (int a = 0; a < NG; ++a) G[a] = nv[a];
            }
        }
    }
    DYNOED_TP_SYNC()
    DYNOED_TP_MARK(3)

    // ---- phase 2: every step on its own, one dual direction per task ----
    // (objective-only call: one unseeded task per step, for its quadrature increment alone)
    const bool want_grad = P.grad != nullptr;
    const int nd_run = want_grad ? ND : 1;
    {
        for (int task = tid; task < n_steps * nd_run; task += blockDim.x) {
            const int k = task / nd_run, m = want_grad ? task - k * nd_run : -1;
            const int j = siv[k];
            T z[NZ], u[NCU], c[M::NCONST], Pq[NQ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                z[a].v = zs[(size_t)k * NZ + a];
                z[a].d[0] = (a == m) ? 1.0 : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < NCTRL; ++kk) {
                u[kk].v = row[P.ctrl_off[kk] + ind[kk * N + j]];
                u[kk].d[0] = (NZ + kk == m) ? 1.0 : 0.0;
            }
            M::template consts_t<T>(u, th, c);
#pragma unroll
            for (int a = 0; a < NQ; ++a) Pq[a] = T(0.0);
            Stepper::template step<M, T>(s_t[k], s_h[k], s_ih[k], z, Pq, u, th, c);
            if (want_grad) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) jc[(size_t)task * NZ + a] = z[a].d[0];
#pragma unroll
                for (int a = 0; a < NQ; ++a) dpq[(size_t)task * NQ + a] = Pq[a].d[0];
            }
            if (m <= 0) {
#pragma unroll
                for (int a = 0; a < NQ; ++a) pqs[(size_t)k * NQ + a] = Pq[a].v;
            }
        }
    }
    DYNOED_TP_SYNC()
    DYNOED_TP_MARK(4)

    // ---- phase 2b: interval quadratures (steps summed in order), F, criterion, Lambda ----
    double best_val = __longlong_as_double(0x7ff0000000000000LL);
    long long best_idx = -1;
    for (int i = tid; i < N * NQ; i += blockDim.x) {
        const int j = i / NQ, a = i - j * NQ;
        double acc = 0.0;
        for (int k = fst[j]; k < fst[j + 1]; ++k) acc += pqs[(size_t)k * NQ + a];
        pq[i] = acc;
    }
    for (int i = tid; i < kTpMaxChunks * ncv; i += blockDim.x) part[i] = 0.0;
    DYNOED_TP_SYNC()
    if (warp == 0) {
        double F[NF];
#pragma unroll
        for (int a = 0; a < NF; ++a) F[a] = 0.0;
        for (int j = lane; j < N; j += 32) {
#pragma unroll
            for (int i = 0; i < NOBS; ++i) {
                const double w = row[P.w_off[i] + ind[(NCTRL + i) * N + j]];
#pragma unroll
                for (int k = 0; k < NF; ++k) F[k] = fma(w, pq[(size_t)j * NQ + i * NF + k], F[k]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int a = 0; a < NF; ++a) F[a] += __shfl_xor_sync(0xffffffffu, F[a], o);
        }
        if (lane == 0) {
            const double tau = row[P.tau_off];
            double L[NF], value;
            criterion_eval<NP>(F, tau, P.criterion, value, L);
#pragma unroll
            for (int jj = 0; jj < NP; ++jj)
#pragma unroll
                for (int ii = 0; ii <= jj; ++ii) ls[tri(ii, jj)] = (ii == jj ? 1.0 : 2.0) * L[tri(ii, jj)];
            double trL = 1.0;
#pragma unroll
            for (int i = 0; i < NP; ++i) trL += L[tri(i, i)];
            ls[NF] = trL;
            const double obj = value + tau;
            P.crit[g] = obj;
            if (P.fim) {
#pragma unroll
                for (int a = 0; a < NF; ++a) P.fim[g * NF + a] = F[a];
            }
            if (obj < best_val) { best_val = obj; best_idx = P.index_base + g; }   // NaN never wins
        }
    }
    DYNOED_TP_SYNC()
    for (int task = tid; want_grad && task < n_steps * ND; task += blockDim.x) {
        const int j = siv[task / ND];
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NOBS; ++i) {
            const double w = row[P.w_off[i] + ind[(NCTRL + i) * N + j]];
            double si = 0.0;
#pragma unroll
            for (int f = 0; f < NF; ++f) si = fma(ls[f], dpq[(size_t)task * NQ + i * NF + f], si);
            s = fma(w, si, s);
        }
        sc[task] = s;
    }
    DYNOED_TP_SYNC()
    DYNOED_TP_MARK(5)

    // ---- phase 3: adjoint recursion over the step Jacobians, cut into chunks of steps ----
    constexpr int NC = NZ + 1;                                   // columns of [A_c | b_c]
    constexpr int LG = NC <= 8 ? 8 : NC <= 16 ? 16 : 32;         // lanes per chunk
    constexpr int CPW = 32 / LG;                                 // chunks per warp
    static_assert(NC <= 32, "one lane per column of a chunk map");
    const int nchunk = want_grad ? min(min(kTpMaxChunks, CPW * (int)(blockDim.x >> 5)), n_steps) : 0;   // 0: nothing to do
    {   // (1) compose the steps of every chunk: column q of A_c is the chain applied to e_q, b_c starts from 0
        const int c = warp * CPW + lane / LG, q = lane % LG;
        if (c < nchunk && q < NC) {
            double v[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) v[a] = (a == q) ? 1.0 : 0.0;
            const int lo = (int)((long long)c * n_steps / nchunk), hi = (int)((long long)(c + 1) * n_steps / nchunk);
            const double sw = (q == NZ) ? 1.0 : 0.0;
            for (int k = hi - 1; k >= lo; --k) {
                const double* col = jc + (size_t)k * ND * NZ;
                const double* sk = sc + (size_t)k * ND;
                double nv[NZ];
#pragma unroll
                for (int m = 0; m < NZ; ++m) nv[m] = sw * sk[m];
#pragma unroll
                for (int r = 0; r < NZ; ++r)
#pragma unroll
                    for (int m = 0; m < NZ; ++m) nv[m] = fma(v[r], col[m * NZ + r], nv[m]);
#pragma unroll
                for (int m = 0; m < NZ; ++m) v[m] = nv[m];
            }
#pragma unroll
            for (int a = 0; a < NZ; ++a) comp[((size_t)c * NC + q) * NZ + a] = v[a];
        }
    }
    DYNOED_TP_SYNC()
    if (warp == 0 && want_grad) {   // (2) chain the chunk maps from the end: lamb[c + 1] = lambda at the upper end of chunk c
        double lam_m = 0.0;
        for (int c = nchunk - 1; c >= 0; --c) {
            if (lane < NZ) lamb[(size_t)(c + 1) * NZ + lane] = lam_m;
            __syncwarp();
            if (lane < NZ) {
                double acc = comp[((size_t)c * NC + NZ) * NZ + lane];
#pragma unroll
                for (int r = 0; r < NZ; ++r) acc = fma(comp[((size_t)c * NC + r) * NZ + lane], lamb[(size_t)(c + 1) * NZ + r], acc);
                lam_m = acc;
            }
            __syncwarp();
        }
        if (lane < NZ) lamb[lane] = lam_m;
        __syncwarp();
        if constexpr (M::NIC > 0) {
            if (lane < M::NIC) grow[P.ic_off + lane] = lamb[M::ic_state(lane)];
        }
    } else if (warp == 1 && want_grad) {   // d/dw from the interval quadratures, d/dtau
        if (lane < NOBS) {
            const int i = lane;
            double wb = 0.0;
            int wk = -1;
            for (int j = 0; j < N; ++j) {
                const int idx = ind[(NCTRL + i) * N + j];
                if (idx != wk) {
                    if (wk >= 0) grow[P.w_off[i] + wk] = wb;
                    wb = 0.0; wk = idx;
                }
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < NF; ++k) sacc = fma(ls[k], pq[(size_t)j * NQ + i * NF + k], sacc);
                wb += sacc;
            }
            if (wk >= 0) grow[P.w_off[i] + wk] = wb;
        } else if (lane == NOBS) {
            grow[P.tau_off] = ls[NF];
        }
    }
    if constexpr (NCTRL > 0) {
        DYNOED_TP_SYNC()
        if (warp == 0 && lane < nchunk) {   // (3) replay every chunk from its boundary value: control columns
            const int c = lane;
            double lam[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) lam[a] = lamb[(size_t)(c + 1) * NZ + a];
            const int lo = (int)((long long)c * n_steps / nchunk), hi = (int)((long long)(c + 1) * n_steps / nchunk);
            for (int k = hi - 1; k >= lo; --k) {
                const double* col = jc + (size_t)k * ND * NZ;
                const double* sk = sc + (size_t)k * ND;
                const int j = siv[k];
                int cbase = 0;
#pragma unroll
                for (int kc = 0; kc < NCTRL; ++kc) {
                    double mu = sk[NZ + kc];
#pragma unroll
                    for (int r = 0; r < NZ; ++r) mu = fma(lam[r], col[(NZ + kc) * NZ + r], mu);
                    part[(size_t)c * ncv + cbase + ind[kc * N + j]] += mu;
                    cbase += P.ctrl_len[kc];
                }
                double nl[NZ];
#pragma unroll
                for (int m = 0; m < NZ; ++m) nl[m] = sk[m];
#pragma unroll
                for (int r = 0; r < NZ; ++r)
#pragma unroll
                    for (int m = 0; m < NZ; ++m) nl[m] = fma(lam[r], col[m * NZ + r], nl[m]);
#pragma unroll
                for (int m = 0; m < NZ; ++m) lam[m] = nl[m];
            }
        }
        DYNOED_TP_SYNC()
        for (int i = tid; want_grad && i < P.n_ctrl_values; i += blockDim.x) {   // later chunks first, as the recursion runs
            double acc = 0.0;
            for (int c = nchunk - 1; c >= 0; --c) acc += part[(size_t)c * ncv + i];
            int rest = i, off = -1;
#pragma unroll
            for (int kc = 0; kc < NCTRL; ++kc) {
                if (off < 0) {
                    if (rest < P.ctrl_len[kc]) off = P.ctrl_off[kc] + rest;
                    else rest -= P.ctrl_len[kc];
                }
            }
            if (off >= 0) grow[off] = acc;
        }
    }
    DYNOED_TP_MARK(6)
    DYNOED_TP_SYNC()
#ifdef DYNOED_TP_TIMING
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) grow[i] = (double)(tp_clk[i + 1] - tp_clk[i]);
        grow[6] = (double)(global_ns() - tp_ns0);      // wall time of the same span, ns
        if (tp_sink == -12345) grow[7] = 0.0;
    }
#endif

    // ---- arg-min record of this CTA (= this design); the last CTA reduces the batch.  Only these record
    //      slots belong to the scratch set (unless tables are spilled), so its guard is taken here, after the work ----
    if constexpr (!SPILL) scratch_acquire(P);
    __shared__ int tp_last;
    if (tid == 0) {
        P.part_val[P.part_base + blockIdx.x] = best_val;
        P.part_idx[P.part_base + blockIdx.x] = best_idx;
        __threadfence();
        const unsigned int t = atomicAdd(P.ticket, 1u);
        tp_last = (t == P.ticket_target - 1u);
    }
    __syncthreads();
    if (tp_last && warp == 0) {
        __threadfence();
        double bv = __longlong_as_double(0x7ff0000000000000LL);
        long long bi = -1;
        for (int i = lane; i < P.n_parts; i += 32) {
            const double v = __ldcg(P.part_val + i);
            const long long k = __ldcg(P.part_idx + i);
            if (k >= 0 && (v < bv || (v == bv && (bi < 0 || k < bi)))) { bv = v; bi = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (ov < bv || (ov == bv && (bi < 0 || oi < bi)))) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            *P.best_val = bv;
            *P.best_idx = bi;
            if (P.ring_rec) { P.ring_rec[0] = __double_as_longlong(bv); P.ring_rec[1] = bi; }
            *P.ticket = 0u;
            if (P.guard) { __threadfence(); atomicAdd(P.guard, 1ULL); }   // this call is done with the scratch set
        }
    }
}

}  // namespace dynoed
