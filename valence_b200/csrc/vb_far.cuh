// Far-field form of the light integral classes (ss|ss) (ps|ss) (ss|ps) (ps|ps), shared by host and device.
//
// For T = rho |PQ|^2 >= 40 the Boys functions are their asymptotic values to 4e-19 relative, and the primitive
// integral is the Coulomb interaction of two point multipoles (the Hermite expansion of the Gaussian products):
//     [e0|f0] = chat_a chat_k  sum_t sum_tau E^e_t E^f_tau (-1)^|tau| D_{t+tau}(R),   R = P_a - Q_k,
//     chat = Kp sqrt(sqrt(pi)/2) / sqrt(p),  D_0 = 1/R = u,  D_x = -R_x u^3,  D_xy = 3 R_x R_y u^5 - delta_xy u^3,
//     E^s_0 = 1;  E^x_0 = PA_x, E^x_x = h = 1/2p            (e on the centre A that carries the angular momentum):
//     (ss|ss) = cc' u
//     (ps|ss)_x = cc' [PA_x u - h R_x u^3]
//     (ss|ps)_x = cc' [QC_x u + h' R_x u^3]
//     (ps|ps)_xy = cc' [PA_x Y_y - h R_x Z_y + delta_xy h h' u^3],   Y = QC u + h' u^3 R,  Z = QC u^3 + 3 h' u^5 R
// No exponent survives except through chat and h; no table, no recurrence: 14 FP64 instructions per (ss|ss) quartet
// including the far-field test, against the 63 operations of the general path.
// (replaces SIMINT's simint_compute_eri for these classes in the regime where most of a large cluster's quartets live,
// /root/reference/src/valence.F90:3398)
#pragma once
#include <cstring>

#include "vb_eri.cuh"

namespace vb {

constexpr double FAR_T = 40.0;                   // = BOYS_S_TMAX: the asymptotic regime of the class kernels

struct alignas(16) FarPrim {                     // 80 bytes, one per primitive pair (index-aligned with pps / pps_flat)
    double Px, Py;                               // |
    double Pz, c;                                // | centre of the Gaussian product, chat
    double ip40, w;                              // FAR_T / p;  magnitude bound (PrimPair::w)
    double PAx, PAy;                             // P - A
    double PAz, ch;                              // chat h, h = 1/2p
};

VB_HD unsigned far_hi32(double x)
{
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(x);
#else
    unsigned long long b; std::memcpy(&b, &x, 8); return (unsigned)(b >> 32);
#endif
}
// Conservative integer form of  wa wk >= tau  for positive normal doubles: the high word of a double is a piecewise
// linear log2 (offset 1023, scale 2^20, between 0 and 0.0861 below log2), so
//     wa wk >= tau   =>   hi(wa) + hi(wk) >= hi(tau) + hi(1.0) - 180700.
// The test keeps everything the exact comparison keeps (and a few quartets just below the cut); it runs on the integer pipe.
VB_HD unsigned far_lthr(double tau) { return far_hi32(tau) + 0x3FF00000u - 180700u; }

#ifdef __CUDA_ARCH__
// 1/sqrt(x), x normal and positive: hardware seed (2^-22) and one third-order correction, no special cases
__device__ __forceinline__ double far_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(x, -(y0 * y0), 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}
#else
inline double far_rsqrt(double x) { return 1.0 / std::sqrt(x); }
#endif

VB_HD FarPrim far_make_prim(const PrimPair& a)
{
    FarPrim f;
    const double c = a.Kp * 0.94139626377671481 * sqrt(a.ip);      // Kp sqrt(sqrt(pi)/2) / sqrt(p)
    f.Px = a.Px; f.Py = a.Py; f.Pz = a.Pz; f.c = c;
    f.ip40 = FAR_T * a.ip * (1.0 + 1e-12); f.w = a.w;
    f.PAx = a.PAx; f.PAy = a.PAy; f.PAz = a.PAz; f.ch = c * (0.5 * a.ip);
    return f;
}

// running sums of one lane over the bra primitives of a shell-pair segment, for one ket primitive
template <int TB, int TK>
struct FarSums {
    static constexpr int N = (TB == 0 && TK == 0) ? 1 : (TB == 1 && TK == 0) ? 3 : (TB == 0 && TK == 1) ? 4 : 9;
    double v[N];
    VB_HD void clear()
    {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = 0.0;
    }
};

// ket-side constants of the lane's ket primitive
struct FarKet {
    double Qx, Qy, Qz, ip40;
    double c, QCx, QCy, QCz, ch;                 // ch = chat' h'
};
VB_HD FarKet far_ket(const FarPrim& k)
{
    FarKet r;
    r.Qx = k.Px; r.Qy = k.Py; r.Qz = k.Pz; r.ip40 = k.ip40; r.c = k.c; r.QCx = k.PAx; r.QCy = k.PAy; r.QCz = k.PAz; r.ch = k.ch;
    return r;
}

// One primitive quartet: returns true when it is in the far field (T >= FAR_T); S accumulates its far-field value when
// `take` (in range, above the magnitude cut) and far, nothing otherwise.  Everything that depends on the ket primitive
// only is applied once per segment in far_finish.
template <int TB, int TK>
VB_HD bool far_quartet(const FarPrim& a, const FarKet& k, bool take, FarSums<TB, TK>& S)
{
    const double Rx = a.Px - k.Qx, Ry = a.Py - k.Qy, Rz = a.Pz - k.Qz;
    const double r2 = fma(Rx, Rx, fma(Ry, Ry, Rz * Rz));
    const bool far = r2 >= a.ip40 + k.ip40;
    double u = far_rsqrt(r2);
    u = (take && far) ? u : 0.0;
    if constexpr (TB == 0 && TK == 0) {
        S.v[0] = fma(a.c, u, S.v[0]);
    } else if constexpr (TB == 1 && TK == 0) {
        // sum_a c [PA u - h R u^3]
        const double g0 = a.c * u, g1 = a.ch * (u * u * u);
        S.v[0] = fma(a.PAx, g0, fma(-Rx, g1, S.v[0]));
        S.v[1] = fma(a.PAy, g0, fma(-Ry, g1, S.v[1]));
        S.v[2] = fma(a.PAz, g0, fma(-Rz, g1, S.v[2]));
    } else if constexpr (TB == 0 && TK == 1) {
        // s0 = sum_a c u,  v = sum_a c u^3 R;   block = c' (QC s0 + h' v)
        const double g0 = a.c * u, g3 = g0 * (u * u);
        S.v[0] += g0;
        S.v[1] = fma(Rx, g3, S.v[1]); S.v[2] = fma(Ry, g3, S.v[2]); S.v[3] = fma(Rz, g3, S.v[3]);
    } else {
        // per quartet, ket constants folded in: c c' [PA_x Y_y - h R_x Z_y + delta_xy h h' u^3]
        const double u2 = u * u, u3 = u2 * u, u5 = u3 * u2;
        const double cu = k.c * u, cu3 = k.c * u3, chu3 = k.ch * u3, chu5 = 3.0 * k.ch * u5;
        const double Y[3] = {fma(k.QCx, cu, chu3 * Rx), fma(k.QCy, cu, chu3 * Ry), fma(k.QCz, cu, chu3 * Rz)};      // c' Y
        const double Z[3] = {fma(k.QCx, cu3, chu5 * Rx), fma(k.QCy, cu3, chu5 * Ry), fma(k.QCz, cu3, chu5 * Rz)};  // c' Z
        const double PA[3] = {a.c * a.PAx, a.c * a.PAy, a.c * a.PAz}, hR[3] = {-a.ch * Rx, -a.ch * Ry, -a.ch * Rz};
        const double dg = a.ch * chu3;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y) S.v[x * 3 + y] = fma(PA[x], Y[y], fma(hR[x], Z[y], S.v[x * 3 + y]));
        S.v[0] += dg; S.v[4] += dg; S.v[8] += dg;
    }
    return far;
}

// acc[e * NF + f] += the segment's far-field block for this ket primitive
template <int TB, int TK>
VB_HD void far_finish(const FarSums<TB, TK>& S, const FarKet& k, double* __restrict__ acc)
{
    if constexpr (TB == 0 && TK == 0) {
        acc[0] = fma(k.c, S.v[0], acc[0]);
    } else if constexpr (TB == 1 && TK == 0) {
        acc[0] = fma(k.c, S.v[0], acc[0]); acc[1] = fma(k.c, S.v[1], acc[1]); acc[2] = fma(k.c, S.v[2], acc[2]);
    } else if constexpr (TB == 0 && TK == 1) {
        const double cs = k.c * S.v[0];
        acc[0] += fma(k.QCx, cs, k.ch * S.v[1]); acc[1] += fma(k.QCy, cs, k.ch * S.v[2]); acc[2] += fma(k.QCz, cs, k.ch * S.v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] += S.v[i];
    }
}

// executed FP64 operations per primitive quartet of the far-field form (FMA = 2; the reciprocal square root is one
// special-function seed + 2 multiplications + 3 FMAs = 9): distance 8, far test 2, rsqrt 9, class sums
VB_HD constexpr double far_flops(int tb, int tk)
{
    return tb + tk == 0 ? 8 + 2 + 9 + 2 : (tb == 1 && tk == 0) ? 8 + 2 + 9 + 5 + 12 : (tb == 0 && tk == 1) ? 8 + 2 + 9 + 3 + 7 : 8 + 2 + 9 + 7 + 24 + 6 + 36 + 4;
}

}  // namespace vb
