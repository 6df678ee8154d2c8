// CPU ORACLE (test infrastructure, not product code) — C++ restatement of the reference's hot path
// for the models named in BASELINE.json, fast enough for 10^4..10^5-design parity samples and for
// the `cpu_baseline` leg of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
// legs may load this library; the product path never does.
//
// What is restated (from the equations, not translated from the Julia sources):
//   * the augmented system  x' = f,  0 = f_xdot G' + f_x G + f_p,  F' = sum_i w_i (h_x[i,:]G)^T(h_x[i,:]G)
//     (/root/reference/src/augment.jl:113-126, :182-188, :223-231) built generically from per-model
//     hand-written f, f_x, f_p, h_x (model sources cited at each model below)
//   * the sequential solve with a restart at every merged-grid point and piecewise-constant
//     controls / weights chosen through the indicator matrix
//     (/root/reference/src/discretize.jl:247-265, :305-344)
//   * the objective c(Symmetric(F + tau I)) + tau (/root/reference/src/problem.jl:115-120,
//     /root/reference/src/fisher.jl:12-107)
//   * the gradient the way the reference computes it: forward-mode dual numbers pushed through the
//     whole integrator in chunks of 12 directions (ForwardDiff's default chunk size for n > 12;
//     AutoForwardDiff at /root/reference/src/problem.jl:154)
// Integrator: FIXED-step classic RK4 or the Tsit5 tableau (parity is only meaningful on a fixed
// tableau/grid), or a fixed-grid Rosenbrock method (ROS2 / ROS3P) for the stiff
// configuration, applied the way the reference applies Rodas5: to the FLAT augmented system
// [x; vec(G); F] with its DENSE Jacobian (obtained here by nested dual numbers, one directional
// derivative per column, so no hand-written second derivatives) and a dense pivoted LU per step.  PARITY PIN: cross-checked against oracle/oed_oracle.py (which
// derives its Jacobians symbolically) in tests/test_oracle.py; integration results against the
// reference itself are unpinned beyond the reference's own 1e-2 test tolerances (no Julia here).
//
// Build: make -C oracle   (g++ -O3 -ffp-contract=off -pthread; std::thread over designs)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <thread>

namespace {

// ---------------------------------------------------------------------------------------------
// forward-mode dual number
// ---------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0.0) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0.0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + b.d[i] * a.v; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { Dual<N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& b, double a) { return a * b; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { return a + b; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { return -a + b; }
// ---------------------------------------------------------------------------------------------
// one-direction dual number over an arbitrary base type (double or Dual<N>): nesting it around
// the gradient duals gives Jacobian-vector products of the right-hand side for the Rosenbrock
// methods without any hand-written second derivative
// ---------------------------------------------------------------------------------------------
template <class B>
struct Du {
    B v, d;
    Du() : v(0.0), d(0.0) {}
    Du(double x) : v(x), d(0.0) {}
    Du(const B& a, const B& b) : v(a), d(b) {}
};
template <class B> inline Du<B> operator+(const Du<B>& a, const Du<B>& b) { return Du<B>(a.v + b.v, a.d + b.d); }
template <class B> inline Du<B> operator-(const Du<B>& a, const Du<B>& b) { return Du<B>(a.v - b.v, a.d - b.d); }
template <class B> inline Du<B> operator-(const Du<B>& a) { return Du<B>(-a.v, -a.d); }
template <class B> inline Du<B> operator*(const Du<B>& a, const Du<B>& b) { return Du<B>(a.v * b.v, a.d * b.v + a.v * b.d); }
template <class B> inline Du<B> operator*(double a, const Du<B>& b) { return Du<B>(a * b.v, a * b.d); }
template <class B> inline Du<B> operator*(const Du<B>& b, double a) { return a * b; }
template <class B> inline Du<B> operator+(const Du<B>& a, double b) { return Du<B>(a.v + b, a.d); }
template <class B> inline Du<B> operator+(double b, const Du<B>& a) { return a + b; }
template <class B> inline Du<B> operator-(const Du<B>& a, double b) { return Du<B>(a.v - b, a.d); }
template <class B> inline Du<B> operator-(double b, const Du<B>& a) { return -a + b; }
inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Dual<N>& x) { return x.v; }

// ---------------------------------------------------------------------------------------------
// models: hand-written f, f_x (differential | algebraic columns), f_p, torn states, h_x
// ---------------------------------------------------------------------------------------------
// x' = p x, y = x            /root/reference/test/references/1D.jl:14-24
struct OneD {
    static constexpr int NX = 1, NXA = 0, NPO = 1, NIC = 0, NOBS = 1, NCTRL = 0, NTH = 1;
    static constexpr bool SQUARE_OBS = false;
    template <class T> static void eval(double, const T* x, const T*, const double* th, const T*, T* f, T (*Jd)[NX], T (*)[1], T (*fp)[1], T (*Hd)[NX], T (*)[1]) {
        f[0] = th[0] * x[0]; Jd[0][0] = T(th[0]); fp[0][0] = x[0]; Hd[0][0] = T(1.0);
    }
    template <class T> static void torn(double, const T*, const double*, T*, T (*)[NX], T (*)[1]) {}
    static constexpr int ic_state(int) { return 0; }
};
// x' = p x, y = x^2          /root/reference/README.md:31-40
struct Readme {
    static constexpr int NX = 1, NXA = 0, NPO = 1, NIC = 0, NOBS = 1, NCTRL = 0, NTH = 1;
    template <class T> static void eval(double, const T* x, const T*, const double* th, const T*, T* f, T (*Jd)[NX], T (*)[1], T (*fp)[1], T (*Hd)[NX], T (*)[1]) {
        f[0] = th[0] * x[0]; Jd[0][0] = T(th[0]); fp[0][0] = x[0]; Hd[0][0] = 2.0 * x[0];
    }
    template <class T> static void torn(double, const T*, const double*, T*, T (*)[NX], T (*)[1]) {}
    static constexpr int ic_state(int) { return 0; }
};
// Lotka-Volterra with fishing control   /root/reference/test/references/LotkaVolterra.jl:10-36
// theta = (p1, p2, p3, p4, c1, c2); tunables (c1, c2); control u; obs (x, y)
struct LVFishing {
    static constexpr int NX = 2, NXA = 0, NPO = 2, NIC = 0, NOBS = 2, NCTRL = 1, NTH = 6;
    template <class T> static void eval(double, const T* x, const T*, const double* th, const T* u, T* f, T (*Jd)[NX], T (*)[1], T (*fp)[2], T (*Hd)[NX], T (*)[1]) {
        const double p1 = th[0], p2 = th[1], p3 = th[2], p4 = th[3], c1 = th[4], c2 = th[5];
        const T X = x[0], Y = x[1], U = u[0];
        f[0] = p1 * X - c1 * (X * Y) - p3 * (X * U);
        f[1] = -p2 * Y + c2 * (X * Y) - p4 * (Y * U);
        Jd[0][0] = p1 - c1 * Y - p3 * U;  Jd[0][1] = -c1 * X;
        Jd[1][0] = c2 * Y;                Jd[1][1] = -p2 + c2 * X - p4 * U;
        fp[0][0] = -(X * Y); fp[0][1] = T(0.0);
        fp[1][0] = T(0.0);   fp[1][1] = X * Y;
        Hd[0][0] = T(1.0); Hd[0][1] = T(0.0); Hd[1][0] = T(0.0); Hd[1][1] = T(1.0);
    }
    template <class T> static void torn(double, const T*, const double*, T*, T (*)[NX], T (*)[2]) {}
    static constexpr int ic_state(int) { return 0; }
};
// LV with unknown initial condition x0   /root/reference/test/references/LotkaVolterra.jl:305-330
// theta = (p1, p2, p3, c1); tunables (c1, x0); obs y + x
struct LVIC {
    static constexpr int NX = 2, NXA = 0, NPO = 1, NIC = 1, NOBS = 1, NCTRL = 0, NTH = 4;
    template <class T> static void eval(double, const T* x, const T*, const double* th, const T*, T* f, T (*Jd)[NX], T (*)[1], T (*fp)[1], T (*Hd)[NX], T (*)[1]) {
        const double p1 = th[0], p2 = th[1], p3 = th[2], c1 = th[3];
        const T X = x[0], Y = x[1];
        f[0] = p1 * X - c1 * (X * Y);
        f[1] = -p2 * Y + p3 * (X * Y);
        Jd[0][0] = p1 - c1 * Y; Jd[0][1] = -c1 * X;
        Jd[1][0] = p3 * Y;      Jd[1][1] = -p2 + p3 * X;
        fp[0][0] = -(X * Y); fp[1][0] = T(0.0);
        Hd[0][0] = T(1.0); Hd[0][1] = T(1.0);
    }
    template <class T> static void torn(double, const T*, const double*, T*, T (*)[NX], T (*)[1]) {}
    static constexpr int ic_state(int) { return 0; }
};
// Robertson DAE, y3 torn from 0 = y1 + y2 + y3 - 1   /root/reference/test/references/Rober.jl:9-21
// theta = (k1, k2, k3); tunable k1; obs y1
struct Robertson {
    static constexpr int NX = 2, NXA = 1, NPO = 1, NIC = 0, NOBS = 1, NCTRL = 0, NTH = 3;
    template <class T> static void eval(double, const T* x, const T* xa, const double* th, const T*, T* f, T (*Jd)[NX], T (*Ja)[1], T (*fp)[1], T (*Hd)[NX], T (*Ha)[1]) {
        const double k1 = th[0], k2 = th[1], k3 = th[2];
        const T y1 = x[0], y2 = x[1], y3 = xa[0];
        f[0] = -k1 * y1 + k3 * (y2 * y3);
        f[1] = k1 * y1 - k3 * (y2 * y3) - k2 * (y2 * y2);
        Jd[0][0] = T(-k1); Jd[0][1] = k3 * y3;                       Ja[0][0] = k3 * y2;
        Jd[1][0] = T(k1);  Jd[1][1] = -k3 * y3 - (2.0 * k2) * y2;    Ja[1][0] = -k3 * y2;
        fp[0][0] = -y1; fp[1][0] = y1;
        Hd[0][0] = T(1.0); Hd[0][1] = T(0.0); Ha[0][0] = T(0.0);
    }
    // y3 = 1 - y1 - y2 ; d y3/d(y1,y2) = (-1,-1) ; d y3/d k1 = 0
    template <class T> static void torn(double, const T* x, const double*, T* xa, T (*phix)[NX], T (*phip)[1]) {
        xa[0] = 1.0 - x[0] - x[1]; phix[0][0] = T(-1.0); phix[0][1] = T(-1.0); phip[0][0] = T(0.0);
    }
    static constexpr int ic_state(int) { return 0; }
};

// Stiff 6-species kinetics with two unknown initial conditions (builder-defined instance of
// BASELINE.json config 4; dynamicoed.jl_b200/models.py::chem6).  States (A, D, B, C, E, Fs);
// theta = (k1..k6); tunables = (A(0), D(0)); obs (E, C)
struct Chem6 {
    static constexpr int NX = 6, NXA = 0, NPO = 0, NIC = 2, NOBS = 2, NCTRL = 0, NTH = 6;
    template <class T> static void eval(double, const T* x, const T*, const double* th, const T*, T* f, T (*Jd)[NX], T (*)[1], T (*)[1], T (*Hd)[NX], T (*)[1]) {
        const double k1 = th[0], k2 = th[1], k3 = th[2], k4 = th[3], k5 = th[4], k6 = th[5];
        const T A = x[0], D = x[1], B = x[2], C = x[3], E = x[4], Fs = x[5];
        f[0] = -k1 * A + k3 * (B * C);
        f[1] = -k4 * (C * D) + k6 * Fs;
        f[2] = k1 * A - k3 * (B * C) - k2 * (B * B);
        f[3] = k2 * (B * B) - k4 * (C * D);
        f[4] = k4 * (C * D) - k5 * E;
        f[5] = k5 * E - k6 * Fs;
        for (int i = 0; i < NX; ++i) for (int j = 0; j < NX; ++j) Jd[i][j] = T(0.0);
        Jd[0][0] = T(-k1); Jd[0][2] = k3 * C; Jd[0][3] = k3 * B;
        Jd[1][1] = -k4 * C; Jd[1][3] = -k4 * D; Jd[1][5] = T(k6);
        Jd[2][0] = T(k1); Jd[2][2] = -k3 * C - (2.0 * k2) * B; Jd[2][3] = -k3 * B;
        Jd[3][1] = -k4 * C; Jd[3][2] = (2.0 * k2) * B; Jd[3][3] = -k4 * D;
        Jd[4][1] = k4 * C; Jd[4][3] = k4 * D; Jd[4][4] = T(-k5);
        Jd[5][4] = T(k5); Jd[5][5] = T(-k6);
        for (int o = 0; o < NOBS; ++o) for (int j = 0; j < NX; ++j) Hd[o][j] = T(0.0);
        Hd[0][4] = T(1.0); Hd[1][3] = T(1.0);
    }
    template <class T> static void torn(double, const T*, const double*, T*, T (*)[NX], T (*)[1]) {}
    static constexpr int ic_state(int i) { return i; }
};

struct Grid {
    int N, n_var, method, crit;
    const double* tgrid;
    const int* first_step; const double* step_t; const double* step_h;
    const int* ind;         // [(NCTRL+NOBS)][N] 0-based
    const int* ctrl_off; const int* w_off; int ic_off, tau_off;
    const double* theta; const double* x0; const double* G0;
};

// ---------------------------------------------------------------------------------------------
// augmented RHS on [x; vec(G) col-major; F_triu col-major]
// ---------------------------------------------------------------------------------------------
template <class M, class T>
inline void rhs(double t, const T* z, const T* u, const T* w, const double* th, T* dz) {
    constexpr int NX = M::NX, NXA = M::NXA, NPO = M::NPO, NP = M::NPO + M::NIC, NOBS = M::NOBS;
    constexpr int NXA1 = NXA > 0 ? NXA : 1, NPO1 = NPO > 0 ? NPO : 1;
    T xa[NXA1], phix[NXA1][NX], phip[NXA1][NPO1];
    M::template torn<T>(t, z, th, xa, phix, phip);
    T f[NX], Jd[NX][NX], Ja[NX][NXA1], fp[NX][NPO1], Hd[NOBS][NX], Ha[NOBS][NXA1];
    M::template eval<T>(t, z, xa, th, u, f, Jd, Ja, fp, Hd, Ha);
    const T* G = z + NX;   // G[j*NX + i]
    T Ga[NXA1][NP];
    for (int r = 0; r < NXA; ++r)
        for (int j = 0; j < NP; ++j) {
            T s(0.0);
            for (int k = 0; k < NX; ++k) s = s + phix[r][k] * G[j * NX + k];
            if (j < NPO) s = s + phip[r][j];
            Ga[r][j] = s;
        }
    for (int i = 0; i < NX; ++i) dz[i] = f[i];
    for (int j = 0; j < NP; ++j)
        for (int i = 0; i < NX; ++i) {
            T s(0.0);
            for (int k = 0; k < NX; ++k) s = s + Jd[i][k] * G[j * NX + k];
            for (int k = 0; k < NXA; ++k) s = s + Ja[i][k] * Ga[k][j];
            if (j < NPO) s = s + fp[i][j];
            dz[NX + j * NX + i] = s;
        }
    T Q[NOBS][NP];
    for (int o = 0; o < NOBS; ++o)
        for (int j = 0; j < NP; ++j) {
            T s(0.0);
            for (int k = 0; k < NX; ++k) s = s + Hd[o][k] * G[j * NX + k];
            for (int k = 0; k < NXA; ++k) s = s + Ha[o][k] * Ga[k][j];
            Q[o][j] = s;
        }
    int k = NX + NX * NP;
    for (int j = 0; j < NP; ++j)
        for (int i = 0; i <= j; ++i) {
            T s(0.0);
            for (int o = 0; o < NOBS; ++o) s = s + w[o] * (Q[o][i] * Q[o][j]);
            dz[k++] = s;
        }
}

static const double RK4_A[4][4] = {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.5, 0, 0}, {0, 0, 1.0, 0}};
static const double RK4_B[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
static const double RK4_C[4] = {0.0, 0.5, 0.5, 1.0};
// Tsitouras 5(4), fixed step (see oracle/oed_oracle.py::tableau)
static const double TS_A[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0}};
static const double TS_B[6] = {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774};
static const double TS_C[6] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0};

// ---------------------------------------------------------------------------------------------
// Rosenbrock methods in implementation form (Hairer/Wanner IV.7):
//   (1/(h gamma) I - J) U_i = g(y + sum_{j<i} a_ij U_j) + sum_{j<i} (c_ij/h) U_j,  y+ = y + sum m_i U_i
// method 2: ROS2 (Verwer et al. 1999), method 3: ROS3P (Lang/Verwer 2001), method 4: RODAS3 (Sandu et al. 1997)
// ---------------------------------------------------------------------------------------------
struct RosTab { int S; double gamma; double a[4][4]; double c[4][4]; double m[4]; };
static const double ROS2_G = 1.7071067811865475244;
static const RosTab ROS2_TAB = {2, ROS2_G, {{0}, {1.0 / ROS2_G}, {0}, {0}},
                                {{0}, {-2.0 / ROS2_G}, {0}, {0}}, {1.5 / ROS2_G, 0.5 / ROS2_G, 0, 0}};
static const RosTab ROS3P_TAB = {3, 7.886751345948129e-01,
                                 {{0}, {1.267949192431123}, {1.267949192431123}, {0}},
                                 {{0}, {-1.607695154586736}, {-3.464101615137755, -1.732050807568877}, {0}},
                                 {2.0, 5.773502691896258e-01, 4.226497308103742e-01, 0}};
// RODAS3 (Sandu et al. 1997): 4 stages, order 3, stiffly accurate, L-stable
static const RosTab RODAS3_TAB = {4, 0.5, {{0}, {0}, {2.0}, {2.0, 0, 1.0}},
                                  {{0}, {4.0}, {1.0, -1.0}, {1.0, -1.0, -8.0 / 3.0}}, {2.0, 0.0, 1.0, 1.0}};

inline double scalar_value(double x) { return x; }
template <int N> inline double scalar_value(const Dual<N>& x) { return x.v; }
inline double div_t(double a, double b) { return a / b; }
template <int N> inline Dual<N> div_t(const Dual<N>& a, const Dual<N>& b) { return a / b; }

template <class M, class T>
void rosenbrock_step(const RosTab& tab, double t, double h, T* y, const T* u, const T* w, const double* th) {
    constexpr int NX = M::NX, NP = M::NPO + M::NIC;
    constexpr int NA = NX + NX * NP + NP * (NP + 1) / 2;
    constexpr int NOBS = M::NOBS, NC1 = M::NCTRL > 0 ? M::NCTRL : 1;
    // dense Jacobian of the flat augmented system, column by column: J e_c = d/d eps g(y + eps e_c)
    T W[NA][NA];
    {
        Du<T> yd[NA], ud[NC1], wd[NOBS], gd[NA];
        for (int k = 0; k < M::NCTRL; ++k) ud[k] = Du<T>(u[k], T(0.0));
        for (int o = 0; o < NOBS; ++o) wd[o] = Du<T>(w[o], T(0.0));
        for (int c = 0; c < NA; ++c) {
            for (int a = 0; a < NA; ++a) yd[a] = Du<T>(y[a], T(a == c ? 1.0 : 0.0));
            rhs<M, Du<T>>(t, yd, ud, wd, th, gd);
            for (int a = 0; a < NA; ++a) W[a][c] = -gd[a].d;
        }
        const double c0 = 1.0 / (h * tab.gamma);
        for (int a = 0; a < NA; ++a) W[a][a] = W[a][a] + c0;
    }
    // LU with partial pivoting (pivot chosen on the value part)
    int perm[NA];
    for (int k = 0; k < NA; ++k) {
        int p = k;
        double best = std::fabs(scalar_value(W[k][k]));
        for (int i = k + 1; i < NA; ++i) {
            const double v = std::fabs(scalar_value(W[i][k]));
            if (v > best) { best = v; p = i; }
        }
        perm[k] = p;
        if (p != k) for (int j = k; j < NA; ++j) std::swap(W[k][j], W[p][j]);   // active part only: the
        // multipliers of earlier columns stay put, matching the interleaved swap/eliminate solve below
        for (int i = k + 1; i < NA; ++i) {
            W[i][k] = div_t(W[i][k], W[k][k]);
            for (int j = k + 1; j < NA; ++j) W[i][j] = W[i][j] - W[i][k] * W[k][j];
        }
    }
    T U[4][NA], Y[NA], g[NA], r[NA];
    for (int s = 0; s < tab.S; ++s) {
        for (int a = 0; a < NA; ++a) Y[a] = y[a];
        for (int j = 0; j < s; ++j)
            if (tab.a[s][j] != 0.0)
                for (int a = 0; a < NA; ++a) Y[a] = Y[a] + tab.a[s][j] * U[j][a];
        rhs<M, T>(t, Y, u, w, th, g);
        for (int a = 0; a < NA; ++a) r[a] = g[a];
        for (int j = 0; j < s; ++j)
            if (tab.c[s][j] != 0.0)
                for (int a = 0; a < NA; ++a) r[a] = r[a] + (tab.c[s][j] / h) * U[j][a];
        for (int k = 0; k < NA; ++k) {
            if (perm[k] != k) std::swap(r[k], r[perm[k]]);
            for (int i = k + 1; i < NA; ++i) r[i] = r[i] - W[i][k] * r[k];
        }
        for (int k = NA - 1; k >= 0; --k) {
            T sum = r[k];
            for (int j = k + 1; j < NA; ++j) sum = sum - W[k][j] * r[j];
            r[k] = div_t(sum, W[k][k]);
        }
        for (int a = 0; a < NA; ++a) U[s][a] = r[a];
    }
    for (int s = 0; s < tab.S; ++s)
        for (int a = 0; a < NA; ++a) y[a] = y[a] + tab.m[s] * U[s][a];
}

template <class M, class T>
void integrate(const Grid& g, const T* design, T* z) {
    constexpr int NX = M::NX, NP = M::NPO + M::NIC, NOBS = M::NOBS, NCTRL = M::NCTRL;
    constexpr int NA = NX + NX * NP + NP * (NP + 1) / 2;
    constexpr int NC1 = NCTRL > 0 ? NCTRL : 1;
    for (int i = 0; i < NX; ++i) z[i] = T(g.x0[i]);
    for (int i = 0; i < NX * NP; ++i) z[NX + i] = T(g.G0[i]);
    for (int i = NX + NX * NP; i < NA; ++i) z[i] = T(0.0);
    for (int a = 0; a < M::NIC; ++a) z[M::ic_state(a)] = design[g.ic_off + a];   // discretize.jl:260-263
    const int S = g.method == 0 ? 4 : 6;
    T k[6][NA], zs[NA];
    for (int j = 0; j < g.N; ++j) {
        T u[NC1], w[NOBS];
        for (int c = 0; c < NCTRL; ++c) u[c] = design[g.ctrl_off[c] + g.ind[c * g.N + j]];
        for (int o = 0; o < NOBS; ++o) w[o] = design[g.w_off[o] + g.ind[(NCTRL + o) * g.N + j]];
        for (int s = g.first_step[j]; s < g.first_step[j + 1]; ++s) {
            const double t = g.step_t[s], h = g.step_h[s];
            if (g.method >= 2) {
                rosenbrock_step<M, T>(g.method == 2 ? ROS2_TAB : g.method == 3 ? ROS3P_TAB : RODAS3_TAB, t, h, z, u, w, g.theta);
                continue;
            }
            for (int st = 0; st < S; ++st) {
                for (int a = 0; a < NA; ++a) zs[a] = z[a];
                for (int q = 0; q < st; ++q) {
                    const double aq = g.method == 0 ? RK4_A[st][q] : TS_A[st][q];
                    if (aq != 0.0)
                        for (int a = 0; a < NA; ++a) zs[a] = zs[a] + (h * aq) * k[q][a];
                }
                rhs<M, T>(t + (g.method == 0 ? RK4_C[st] : TS_C[st]) * h, zs, u, w, g.theta, k[st]);
            }
            for (int st = 0; st < S; ++st) {
                const double b = g.method == 0 ? RK4_B[st] : TS_B[st];
                for (int a = 0; a < NA; ++a) z[a] = z[a] + (h * b) * k[st][a];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// criteria (fisher.jl:26-107) and d c / d M, symmetric n x n, n <= 8, cyclic Jacobi for eigenpairs
// ---------------------------------------------------------------------------------------------
constexpr int MAXN = 8;
inline int tri(int i, int j) { return i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j; }

void jacobi(int n, double A[MAXN][MAXN], double V[MAXN][MAXN]) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dia = 0;
        for (int p = 0; p < n; ++p) { dia += A[p][p] * A[p][p]; for (int q = p + 1; q < n; ++q) off += A[p][q] * A[p][q]; }
        if (off == 0.0 || off <= 1e-36 * dia) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                if (A[p][q] == 0.0) continue;
                const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
                for (int k = 0; k < n; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
                for (int k = 0; k < n; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
            }
    }
}

// value of c(F + tau I) and packed gradient matrix L
double criterion(int crit, int n, const double* Fp, double tau, double* L) {
    double A[MAXN][MAXN], V[MAXN][MAXN], M[MAXN][MAXN];
    for (int j = 0; j < n; ++j) for (int i = 0; i <= j; ++i) { const double v = Fp[tri(i, j)] + (i == j ? tau : 0.0); A[i][j] = A[j][i] = v; M[i][j] = M[j][i] = v; }
    const int nF = n * (n + 1) / 2;
    if (crit == 0) {
        double tr = 0; for (int i = 0; i < n; ++i) tr += M[i][i];
        for (int k = 0; k < nF; ++k) L[k] = 0.0;
        for (int i = 0; i < n; ++i) L[tri(i, i)] = -1.0;
        return -tr;
    }
    jacobi(n, A, V);
    double lam[MAXN], det = 1.0, sinv = 0.0; int kmin = 0, kinv = 0;
    for (int i = 0; i < n; ++i) { lam[i] = A[i][i]; det *= lam[i]; sinv += 1.0 / lam[i]; if (lam[i] < lam[kmin]) kmin = i; }
    for (int i = 0; i < n; ++i) if (1.0 / lam[i] > 1.0 / lam[kinv]) kinv = i;
    double gk[MAXN], val;
    switch (crit) {
        case 1: val = -det; for (int i = 0; i < n; ++i) gk[i] = -det / lam[i]; break;
        case 2: val = -lam[kmin]; for (int i = 0; i < n; ++i) gk[i] = i == kmin ? -1.0 : 0.0; break;
        case 3: val = sinv; for (int i = 0; i < n; ++i) gk[i] = -1.0 / (lam[i] * lam[i]); break;
        case 4: val = 1.0 / det; for (int i = 0; i < n; ++i) gk[i] = -1.0 / (det * lam[i]); break;
        case 6: val = -std::log(det); for (int i = 0; i < n; ++i) gk[i] = -1.0 / lam[i]; break;   // log form of D (extension)
        default: val = 1.0 / lam[kinv]; for (int i = 0; i < n; ++i) gk[i] = i == kinv ? -1.0 / (lam[i] * lam[i]) : 0.0; break;
    }
    for (int j = 0; j < n; ++j) for (int i = 0; i <= j; ++i) { double s = 0; for (int k = 0; k < n; ++k) s += V[i][k] * gk[k] * V[j][k]; L[tri(i, j)] = s; }
    return val;
}

constexpr int CHUNK = 12;   // ForwardDiff.pickchunksize(n) for n > 12 (DEFAULT_CHUNK_THRESHOLD)

template <class M>
void eval_one(const Grid& g, const double* design, double* crit_out, double* grad_out, double* fim_out) {
    constexpr int NX = M::NX, NP = M::NPO + M::NIC, NF = NP * (NP + 1) / 2;
    constexpr int NA = NX + NX * NP + NF;
    const double tau = design[g.tau_off];
    double L[NF > 0 ? NF : 1];
    if (!grad_out) {
        double z[NA];
        integrate<M, double>(g, design, z);
        *crit_out = criterion(g.crit, NP, z + NX + NX * NP, tau, L) + tau;
        if (fim_out) for (int k = 0; k < NF; ++k) fim_out[k] = z[NX + NX * NP + k];
        return;
    }
    std::vector<Dual<CHUNK>> des(g.n_var);
    Dual<CHUNK> z[NA];
    bool have = false;
    for (int c0 = 0; c0 < g.n_var; c0 += CHUNK) {
        for (int v = 0; v < g.n_var; ++v) {
            des[v] = Dual<CHUNK>(design[v]);
            if (v >= c0 && v < c0 + CHUNK) des[v].d[v - c0] = 1.0;
        }
        integrate<M, Dual<CHUNK>>(g, des.data(), z);
        if (!have) {
            double Fv[NF];
            for (int k = 0; k < NF; ++k) Fv[k] = z[NX + NX * NP + k].v;
            *crit_out = criterion(g.crit, NP, Fv, tau, L) + tau;
            if (fim_out) for (int k = 0; k < NF; ++k) fim_out[k] = Fv[k];
            have = true;
        }
        for (int v = c0; v < std::min(c0 + CHUNK, g.n_var); ++v) {
            double s = 0.0;   // <L, dF/dv> over the full symmetric matrix
            for (int j = 0; j < NP; ++j)
                for (int i = 0; i <= j; ++i) s += (i == j ? 1.0 : 2.0) * L[tri(i, j)] * z[NX + NX * NP + tri(i, j)].d[v - c0];
            grad_out[v] = s;
        }
    }
    double trL = 0.0;
    for (int i = 0; i < NP; ++i) trL += L[tri(i, i)];
    grad_out[g.tau_off] += trL + 1.0;
}

template <class M>
void eval_batch(const Grid& g, const double* designs, int64_t n, double* crit, double* grad, double* fim, int threads) {
    constexpr int NP = M::NPO + M::NIC, NF = NP * (NP + 1) / 2;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n));
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k)
            eval_one<M>(g, designs + k * g.n_var, crit + k, grad ? grad + k * g.n_var : nullptr,
                        fim ? fim + k * NF : nullptr);
    };
    if (nt == 1) { work(0, n); return; }
    std::vector<std::thread> pool;   // static contiguous partition of the designs over threads
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& th : pool) th.join();
}

template <class M>
void traj_one(const Grid& g, const double* design, double* X) {
    // X: [N+1][n_aug] at the merged grid points; re-integrates prefix by prefix (test sizes only)
    constexpr int NX = M::NX, NP = M::NPO + M::NIC, NA = NX + NX * NP + NP * (NP + 1) / 2;
    for (int j = 0; j <= g.N; ++j) {
        Grid gj = g; gj.N = j;
        // indicators are indexed with stride g.N: rebuild a compact copy
        std::vector<int> ind((M::NCTRL + M::NOBS) * std::max(j, 1));
        for (int v = 0; v < M::NCTRL + M::NOBS; ++v) for (int q = 0; q < j; ++q) ind[v * j + q] = g.ind[v * g.N + q];
        gj.ind = ind.data();
        double z[NA];
        integrate<M, double>(gj, design, z);
        for (int a = 0; a < NA; ++a) X[j * NA + a] = z[a];
    }
}

}  // namespace

extern "C" {

struct oracle_desc {
    int32_t model;       // 0 one_d, 1 readme, 2 lv_fishing, 3 lv_ic, 4 robertson, 5 chem6
    int32_t crit, method, N, n_var, n_steps;
    const double* tgrid; const int32_t* first_step; const double* step_t; const double* step_h;
    const int32_t* ind; const int32_t* ctrl_off; const int32_t* w_off; int32_t ic_off, tau_off;
    const double* theta; const double* x0; const double* G0;
};

// Model constants written down HERE from the reference's own test files, for callers that pass
// theta / x0 / G0 = NULL: an oracle fed from the product's models.py would share a wrong constant with it.
//   lv_fishing: /root/reference/test/references/LotkaVolterra.jl:11 (x = 0.5, y = 0.7), :13-17
//               (p = [1.0; 1.0; 0.4; 0.6], c = [1.0; 1.0]); no unknown initial conditions => G0 = 0
//   one_d     : /root/reference/test/references/1D.jl:15-16 (x = 1.0, p = -2.0)
static bool own_constants(int model, const double** th, const double** x0, const double** G0) {
    static const double lv_th[6] = {1.0, 1.0, 0.4, 0.6, 1.0, 1.0}, lv_x0[2] = {0.5, 0.7}, zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static const double od_th[1] = {-2.0}, od_x0[1] = {1.0};
    if (model == 2) { *th = lv_th; *x0 = lv_x0; *G0 = zeros; return true; }
    if (model == 0) { *th = od_th; *x0 = od_x0; *G0 = zeros; return true; }
    return false;
}

static Grid to_grid(const oracle_desc* d) {
    Grid g;
    g.N = d->N; g.n_var = d->n_var; g.method = d->method; g.crit = d->crit; g.tgrid = d->tgrid;
    g.first_step = d->first_step; g.step_t = d->step_t; g.step_h = d->step_h; g.ind = d->ind;
    g.ctrl_off = d->ctrl_off; g.w_off = d->w_off; g.ic_off = d->ic_off; g.tau_off = d->tau_off;
    g.theta = d->theta; g.x0 = d->x0; g.G0 = d->G0;
    if (!d->theta && !d->x0 && !d->G0) own_constants(d->model, &g.theta, &g.x0, &g.G0);
    return g;
}

int oracle_eval_batch(const oracle_desc* d, const double* designs, int64_t n, double* crit, double* grad, double* fim, int threads) {
    const Grid g = to_grid(d);
    switch (d->model) {
        case 0: eval_batch<OneD>(g, designs, n, crit, grad, fim, threads); break;
        case 1: eval_batch<Readme>(g, designs, n, crit, grad, fim, threads); break;
        case 2: eval_batch<LVFishing>(g, designs, n, crit, grad, fim, threads); break;
        case 3: eval_batch<LVIC>(g, designs, n, crit, grad, fim, threads); break;
        case 4: eval_batch<Robertson>(g, designs, n, crit, grad, fim, threads); break;
        case 5: eval_batch<Chem6>(g, designs, n, crit, grad, fim, threads); break;
        default: return -1;
    }
    return 0;
}

int oracle_trajectory(const oracle_desc* d, const double* design, double* X) {
    const Grid g = to_grid(d);
    switch (d->model) {
        case 0: traj_one<OneD>(g, design, X); break;
        case 1: traj_one<Readme>(g, design, X); break;
        case 2: traj_one<LVFishing>(g, design, X); break;
        case 3: traj_one<LVIC>(g, design, X); break;
        case 4: traj_one<Robertson>(g, design, X); break;
        case 5: traj_one<Chem6>(g, design, X); break;
        default: return -1;
    }
    return 0;
}

double oracle_criterion(int crit, int n, const double* F_packed, double tau, double* L_packed) {
    double L[MAXN * (MAXN + 1) / 2];
    const double v = criterion(crit, n, F_packed, tau, L);
    if (L_packed) for (int k = 0; k < n * (n + 1) / 2; ++k) L_packed[k] = L[k];
    return v;
}

int oracle_max_threads(void) {
    return (int)std::max(1u, std::thread::hardware_concurrency());
}

}  // extern "C"
