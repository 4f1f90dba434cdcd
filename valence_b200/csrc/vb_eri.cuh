// Obara-Saika vertical recurrence for contracted [e0|f0] auxiliary integrals,
// Boys function, and the cartesian index helpers shared by host and device.
//
// This replaces the arithmetic the reference delegates to SIMINT at
// /root/reference/src/valence.F90:3398 (simint_compute_eri) -- re-derived, not
// translated: the reference's HRR + "digestion" (valence.F90:3372-3412) is
// folded on the host into the orbital-pair density (vb_setup.cpp, hrr_fold),
// so the device only ever forms [e0|f0] with e in [la, la+lb], f in [lc, lc+ld].
#pragma once
#include <cmath>
#include <type_traits>

#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD inline
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#endif

namespace vb {

constexpr int LMAX_SHELL = 2;             // s, p, d (nang <= 2 in every reference input)
constexpr int EMAX = 2 * LMAX_SHELL;      // la + lb
constexpr int MTOP = 2 * EMAX;            // highest Boys order of an ERI
constexpr double PI = 3.14159265358979323846264338327950288;

VB_HD constexpr int ncart(int L) { return (L + 1) * (L + 2) / 2; }
VB_HD constexpr int coff(int L) { return L * (L + 1) * (L + 2) / 6; }   // comps with l < L
VB_HD constexpr int ncum(int L) { return coff(L + 1); }                 // comps with l <= L
// CCA order inside one l: for i=0..L, j=0..i: (L-i, i-j, j)   (valence.F90:2365-2378)
VB_HD constexpr int cidx(int lx, int ly, int lz)
{
    int L = lx + ly + lz, i = L - lx;
    return coff(L) + i * (i + 1) / 2 + lz;
}
VB_HD constexpr int c_L(int c) { int L = 0; while (coff(L + 1) <= c) ++L; return L; }
VB_HD constexpr int c_i(int c) { int r = c - coff(c_L(c)), i = 0; while ((i + 1) * (i + 2) / 2 <= r) ++i; return i; }
VB_HD constexpr int c_lx(int c) { return c_L(c) - c_i(c); }
VB_HD constexpr int c_lz(int c) { int i = c_i(c); return c - coff(c_L(c)) - i * (i + 1) / 2; }
VB_HD constexpr int c_ly(int c) { return c_i(c) - c_lz(c); }
VB_HD constexpr int c_l(int c, int d) { return d == 0 ? c_lx(c) : (d == 1 ? c_ly(c) : c_lz(c)); }
// direction used to build component c from lower ones: first non-zero of x,y,z
VB_HD constexpr int c_dir(int c) { return c_lx(c) > 0 ? 0 : (c_ly(c) > 0 ? 1 : 2); }
VB_HD constexpr int c_dec(int c, int d)   // index of c - 1_d  (requires l_d(c) >= 1)
{
    return cidx(c_lx(c) - (d == 0), c_ly(c) - (d == 1), c_lz(c) - (d == 2));
}

// pair types: (la >= lb) encoded as la*(la+1)/2 + lb : ss=0 ps=1 pp=2 ds=3 dp=4 dd=5
constexpr int NPTYPE = 6;
VB_HD constexpr int ptype(int la, int lb) { return la * (la + 1) / 2 + lb; }
VB_HD constexpr int pt_la(int t) { return t < 1 ? 0 : (t < 3 ? 1 : 2); }
VB_HD constexpr int pt_lb(int t) { return t - pt_la(t) * (pt_la(t) + 1) / 2; }
VB_HD constexpr int pt_E(int t) { return pt_la(t) + pt_lb(t); }                       // highest e
VB_HD constexpr int pt_ne(int t) { return ncum(pt_E(t)) - coff(pt_la(t)); }           // # [e0| comps kept

// ---------------------------------------------------------------------------
// Boys function.  Table: row k (T_k = k/16, k = 0..BOYS_ROWS-1) holds
// F_m(T_k), m = 0..BOYS_COLS-1.  F_mtop(T) by an 8-term Taylor series about the
// nearest grid point (|dT| <= 1/32 -> truncation < 3e-17 relative), lower
// orders by downward recursion; beyond the table F_0 = sqrt(pi/T)/2 and
// upward recursion (stable for T >> m).
// ---------------------------------------------------------------------------
constexpr int BOYS_COLS = 16;                 // MTOP + 8
constexpr double BOYS_STEP = 1.0 / 16.0;
constexpr double BOYS_TMAX = 64.0;              // beyond: exp(-T) < 2e-28, pure asymptotic series
constexpr int BOYS_ROWS = 64 * 16 + 1;

inline void boys_reference(int mmax, double T, double* F)   // host: exact series / erf
{
    if (T < 35.0) {
        double eT = std::exp(-T), term = 1.0 / (2.0 * mmax + 1.0), sum = term;
        for (int k = 1; k < 500; ++k) {
            term *= 2.0 * T / (2.0 * mmax + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-19 * sum) break;
        }
        F[mmax] = eT * sum;
        for (int m = mmax; m > 0; --m) F[m - 1] = (2.0 * T * F[m] + eT) / (2.0 * m - 1.0);
    } else {
        double eT = std::exp(-T), st = std::sqrt(T);
        F[0] = 0.5 * std::sqrt(PI) / st * std::erf(st);
        for (int m = 0; m < mmax; ++m) F[m + 1] = ((2.0 * m + 1.0) * F[m] - eT) / (2.0 * T);
    }
}

inline void boys_make_table(double* tab)   // BOYS_ROWS * BOYS_COLS doubles
{
    for (int k = 0; k < BOYS_ROWS; ++k) boys_reference(BOYS_COLS - 1, k * BOYS_STEP, tab + (size_t)k * BOYS_COLS);
}

VB_HD double boys_taylor(const double* __restrict__ r, double d)   // sum_j r[j] d^j / j!, j < 8
{
    double f = r[7] * (1.0 / 5040.0);
    f = f * d + r[6] * (1.0 / 720.0);
    f = f * d + r[5] * (1.0 / 120.0);
    f = f * d + r[4] * (1.0 / 24.0);
    f = f * d + r[3] * (1.0 / 6.0);
    f = f * d + r[2] * 0.5;
    f = f * d + r[1];
    return f * d + r[0];
}

template <int M>
VB_HD void boys(const double* __restrict__ tab, double T, double* F)   // F[0..M]
{
    if (T < BOYS_TMAX) {
        int k = (int)(T * (1.0 / BOYS_STEP) + 0.5);
        double d = k * BOYS_STEP - T;                 // F_m(T) = sum_j F_{m+j}(T_k) d^j / j!
        const double* r = tab + (size_t)k * BOYS_COLS;
        if constexpr (M <= 2) {                       // low orders: one Taylor series each, no exp
#pragma unroll
            for (int m = 0; m <= M; ++m) F[m] = boys_taylor(r + m, d);
        } else {
            F[M] = boys_taylor(r + M, d);
            double eT = exp(-T), t2 = 2.0 * T;
#pragma unroll
            for (int m = M; m > 0; --m) F[m - 1] = (t2 * F[m] + eT) * (1.0 / (2.0 * m - 1.0));
        }
    } else {
        double r = rsqrt(T), rt = r * r;
        F[0] = 0.88622692545275801365 * r;            // sqrt(pi)/2 / sqrt(T)
#pragma unroll
        for (int m = 0; m < M; ++m) F[m + 1] = (m + 0.5) * F[m] * rt;
    }
}

// Compact table for the s/p kernel (orders <= 4), small enough to live in shared memory:
// step 1/8, T < 40, 9-term Taylor series (|dT| <= 1/16 -> remainder < 3e-18), F_0 .. F_12 per row.
constexpr int BOYS_S_COLS = 13;
constexpr double BOYS_S_STEP = 0.125, BOYS_S_TMAX = 40.0;
constexpr int BOYS_S_ROWS = 321;
constexpr int BOYS_S_SIZE = (BOYS_S_ROWS * BOYS_S_COLS + 1) & ~1;

inline void boys_make_table_small(double* tab)
{
    for (int k = 0; k < BOYS_S_ROWS; ++k) boys_reference(BOYS_S_COLS - 1, k * BOYS_S_STEP, tab + (size_t)k * BOYS_S_COLS);
}

VB_HD double boys_taylor9(const double* __restrict__ r, double d)   // sum_j r[j] d^j / j!, j < 9
{
    double f = r[8] * (1.0 / 40320.0);
    f = f * d + r[7] * (1.0 / 5040.0);
    f = f * d + r[6] * (1.0 / 720.0);
    f = f * d + r[5] * (1.0 / 120.0);
    f = f * d + r[4] * (1.0 / 24.0);
    f = f * d + r[3] * (1.0 / 6.0);
    f = f * d + r[2] * 0.5;
    f = f * d + r[1];
    return f * d + r[0];
}

template <int M>
VB_HD void boys_s(const double* __restrict__ tab, double T, double* F)   // F[0..M], M <= 4
{
    static_assert(M + 9 <= BOYS_S_COLS, "compact Boys table too narrow");
    if (T < BOYS_S_TMAX) {
        const int k = (int)(T * (1.0 / BOYS_S_STEP) + 0.5);
        const double d = k * BOYS_S_STEP - T;
        const double* r = tab + k * BOYS_S_COLS;
        // Horner form of sum_j r[j] d^j / j! with the factorials folded into the variable: the scaled
        // steps d/j are shared by every order, so each order costs 8 fused multiply-adds
        const double d2 = d * 0.5, d3 = d * (1.0 / 3.0), d4 = d * 0.25, d5 = d * 0.2, d6 = d * (1.0 / 6.0), d7 = d * (1.0 / 7.0),
                     d8 = d * 0.125;
        auto series = [&](const double* c) {
            double f = c[8];
            f = f * d8 + c[7];
            f = f * d7 + c[6];
            f = f * d6 + c[5];
            f = f * d5 + c[4];
            f = f * d4 + c[3];
            f = f * d3 + c[2];
            f = f * d2 + c[1];
            return f * d + c[0];
        };
        if constexpr (M <= 2) {
#pragma unroll
            for (int m = 0; m <= M; ++m) F[m] = series(r + m);
        } else {
            F[M] = series(r + M);
            const double eT = exp(-T), t2 = 2.0 * T;
#pragma unroll
            for (int m = M; m > 0; --m) F[m - 1] = (t2 * F[m] + eT) * (1.0 / (2.0 * m - 1.0));
        }
    } else {
        const double r = rsqrt(T), rt = r * r;
        F[0] = 0.88622692545275801365 * r;            // sqrt(pi)/2 / sqrt(T); exp(-T) < 5e-18 is dropped
#pragma unroll
        for (int m = 0; m < M; ++m) F[m + 1] = (m + 0.5) * F[m] * rt;
    }
}

VB_HD void boys_rt(int M, const double* __restrict__ tab, double T, double* F)   // runtime order
{
    if (T < BOYS_TMAX) {
        int k = (int)(T * (1.0 / BOYS_STEP) + 0.5);
        double d = k * BOYS_STEP - T;
        F[M] = boys_taylor(tab + (size_t)k * BOYS_COLS + M, d);
        double eT = exp(-T), t2 = 2.0 * T;
        for (int m = M; m > 0; --m) F[m - 1] = (t2 * F[m] + eT) / (2.0 * m - 1.0);
    } else {
        double r = 1.0 / sqrt(T), rt = r * r;
        F[0] = 0.88622692545275801365 * r;
        for (int m = 0; m < M; ++m) F[m + 1] = (m + 0.5) * F[m] * rt;
    }
}

// ---------------------------------------------------------------------------
// primitive shell-pair record (64 bytes, staged through shared memory)
// ---------------------------------------------------------------------------
struct alignas(16) PrimPair {   // 96 bytes (multiple of 16: moved with cp.async.bulk)
    double Px, Py, Pz;     // Gaussian product centre
    double p;              // a + b
    double ip;             // 1 / p
    double Kp;             // c_a c_b exp(-ab/p |AB|^2) * sqrt(2) pi^(5/4) / p
    double PAx, PAy, PAz;  // P - A, A = centre carrying the angular momentum (la >= lb)
    double w;              // magnitude bound used to skip negligible primitive quartets
    double wseg;           // largest w of this primitive pair's shell pair (non-increasing within a pair type)
    int eoff, pad;         // e-offset of the shell pair inside its pair group
};

struct QuartetGeom {       // everything the VRR needs for one primitive quartet
    double PA[3], QC[3], WP[3], WQ[3];
    double h2p, h2q, h2pq, rp, rq;   // 1/2p, 1/2q, 1/2(p+q), rho/p, rho/q
};

// compile-time loop
template <int B, int E, class F>
VB_HD void sfor(F&& f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        sfor<B + 1, E>(f);
    }
}

// ---------------------------------------------------------------------------
// Fully unrolled VRR for EA, EC <= 2 (the s/p classes of 6-31G).  Adds
// pref * [e0|f0]^(0) for e in [coff(LA), ncum(EA)), f in [coff(LC), ncum(EC))
// to acc[(e - coff(LA)) * NF + (f - coff(LC))].
//   [e+1_d 0|00]^m = PA_d [e0|00]^m + WP_d [e0|00]^(m+1)
//                    + e_d/2p ([e-1_d 0|00]^m - rho/p [e-1_d 0|00]^(m+1))
//   [e0|f+1_d 0]^m = QC_d [e0|f0]^m + WQ_d [e0|f0]^(m+1)
//                    + f_d/2q ([e0|f-1_d 0]^m - rho/q [e0|f-1_d 0]^(m+1))
//                    + e_d/2(p+q) [e-1_d 0|f0]^(m+1)
// ---------------------------------------------------------------------------
template <int LA, int EA, int LC, int EC>
VB_HD void vrr_unrolled(const QuartetGeom& g, const double* __restrict__ Fm /* pref*F_m, m=0..EA+EC */,
                        double* __restrict__ acc)
{
    constexpr int NE = ncum(EA), NFc = ncum(EC), MT = EA + EC;
    constexpr int NF = ncum(EC) - coff(LC);
    double T[MT + 1][NE][NFc];
    sfor<0, MT + 1>([&](auto M) { T[M][0][0] = Fm[M]; });
    sfor<1, NE>([&](auto EI) {
        constexpr int e = EI, d = c_dir(e), e1 = c_dec(e, d), n1 = c_l(e1, d), Le = c_L(e);
        sfor<0, MT + 1 - Le>([&](auto M) {
            constexpr int m = M;
            double v = g.PA[d] * T[m][e1][0] + g.WP[d] * T[m + 1][e1][0];
            if constexpr (n1 > 0) {
                constexpr int e2 = c_dec(e1, d);
                v += (n1 * g.h2p) * (T[m][e2][0] - g.rp * T[m + 1][e2][0]);
            }
            T[m][e][0] = v;
        });
    });
    sfor<1, NFc>([&](auto FI) {
        constexpr int f = FI, d = c_dir(f), f1 = c_dec(f, d), n1 = c_l(f1, d), Lf = c_L(f);
        sfor<0, NE>([&](auto EI) {
            constexpr int e = EI, Le = c_L(e), ne = c_l(e, d);
            sfor<0, MT + 1 - Le - Lf>([&](auto M) {
                constexpr int m = M;
                double v = g.QC[d] * T[m][e][f1] + g.WQ[d] * T[m + 1][e][f1];
                if constexpr (n1 > 0) {
                    constexpr int f2 = c_dec(f1, d);
                    v += (n1 * g.h2q) * (T[m][e][f2] - g.rq * T[m + 1][e][f2]);
                }
                if constexpr (ne > 0) {
                    constexpr int em = c_dec(e, d);
                    v += (ne * g.h2pq) * T[m + 1][em][f1];
                }
                T[m][e][f] = v;
            });
        });
    });
    sfor<coff(LA), NE>([&](auto EI) {
        sfor<coff(LC), NFc>([&](auto FI) {
            constexpr int e = EI, f = FI;
            acc[(e - coff(LA)) * NF + (f - coff(LC))] += T[0][e][f];
        });
    });
}

// ---------------------------------------------------------------------------
// Loop-based VRR for any EA, EC <= EMAX (d shells).  `T` is caller-provided
// scratch of (MTOP+1) * ncum(EMAX) * ncum(EMAX) doubles.
// ---------------------------------------------------------------------------
constexpr int GEN_NE = ncum(EMAX);
constexpr int GEN_SCRATCH = (MTOP + 1) * GEN_NE * GEN_NE;

VB_HD void vrr_generic(int LA, int EA, int LC, int EC, const QuartetGeom& g, const double* __restrict__ Fm,
                       double* __restrict__ T, double* __restrict__ acc)
{
    const int NE = ncum(EA), NFc = ncum(EC), MT = EA + EC;
    const int NF = ncum(EC) - coff(LC);
#define TT(m, e, f) T[((m) * GEN_NE + (e)) * GEN_NE + (f)]
    for (int m = 0; m <= MT; ++m) TT(m, 0, 0) = Fm[m];
    for (int e = 1; e < NE; ++e) {
        int d = c_dir(e), e1 = c_dec(e, d), n1 = c_l(e1, d), Le = c_L(e);
        int e2 = n1 > 0 ? c_dec(e1, d) : 0;
        for (int m = 0; m <= MT - Le; ++m) {
            double v = g.PA[d] * TT(m, e1, 0) + g.WP[d] * TT(m + 1, e1, 0);
            if (n1 > 0) v += (n1 * g.h2p) * (TT(m, e2, 0) - g.rp * TT(m + 1, e2, 0));
            TT(m, e, 0) = v;
        }
    }
    for (int f = 1; f < NFc; ++f) {
        int d = c_dir(f), f1 = c_dec(f, d), n1 = c_l(f1, d), Lf = c_L(f);
        int f2 = n1 > 0 ? c_dec(f1, d) : 0;
        for (int e = 0; e < NE; ++e) {
            int Le = c_L(e), ne = c_l(e, d);
            int em = ne > 0 ? c_dec(e, d) : 0;
            for (int m = 0; m <= MT - Le - Lf; ++m) {
                double v = g.QC[d] * TT(m, e, f1) + g.WQ[d] * TT(m + 1, e, f1);
                if (n1 > 0) v += (n1 * g.h2q) * (TT(m, e, f2) - g.rq * TT(m + 1, e, f2));
                if (ne > 0) v += (ne * g.h2pq) * TT(m + 1, em, f1);
                TT(m, e, f) = v;
            }
        }
    }
    for (int e = coff(LA); e < NE; ++e)
        for (int f = coff(LC); f < NFc; ++f) acc[(e - coff(LA)) * NF + (f - coff(LC))] += TT(0, e, f);
#undef TT
}

// geometry of one primitive quartet; returns T = rho |PQ|^2 and the prefactor K_a K_b / (p q sqrt(p+q)).
// Division-free: one rsqrt.
VB_HD void quartet_geom(const PrimPair& a, const PrimPair& b, QuartetGeom& g, double& T, double& pref)
{
    const double p = a.p, q = b.p;
    const double r = rsqrt(p + q), ipq = r * r;
    const double wq = q * ipq, wp = p * ipq;             // rho/p, rho/q
    const double dx = a.Px - b.Px, dy = a.Py - b.Py, dz = a.Pz - b.Pz;
    T = p * wq * (dx * dx + dy * dy + dz * dz);
    pref = a.Kp * b.Kp * r;
    g.PA[0] = a.PAx; g.PA[1] = a.PAy; g.PA[2] = a.PAz;
    g.QC[0] = b.PAx; g.QC[1] = b.PAy; g.QC[2] = b.PAz;
    g.WP[0] = -wq * dx; g.WP[1] = -wq * dy; g.WP[2] = -wq * dz;   // W - P
    g.WQ[0] = wp * dx; g.WQ[1] = wp * dy; g.WQ[2] = wp * dz;      // W - Q
    g.h2p = 0.5 * a.ip; g.h2q = 0.5 * b.ip; g.h2pq = 0.5 * ipq; g.rp = wq; g.rq = wp;
}

}  // namespace vb
