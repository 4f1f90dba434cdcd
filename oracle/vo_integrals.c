/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * AO integrals over contracted cartesian Gaussian shells, McMurchie-Davidson
 * scheme with an exact Boys function.  This restates what the reference gets
 * from the third-party SIMINT library (un-vendored, un-pinned:
 * /root/reference/install-simint.sh:4, `git clone .../simint-generator` HEAD,
 * `create.py -l 3 -p 3`, SIMINT_VECTOR=scalar), at the reference's call sites:
 *   simint_compute_overlap    /root/reference/src/valence.F90:2988
 *   simint_compute_ke         /root/reference/src/valence.F90:3132
 *   simint_compute_potential  /root/reference/src/valence.F90:3151-3153
 *   simint_compute_eri        /root/reference/src/valence.F90:3398
 * Conventions: plain cartesian x^lx y^ly z^lz times
 * sum_g c_g exp(-a_g r^2), c_g as handed over by the caller (VALENCE passes
 * its own normalised con_coeff, valence_simint_module.F90:45-48, and applies
 * the per-component factor angn itself); component order is CCA
 * (valence.F90:2365-2378); ERI block layout is ((i*nj+j)*nk+k)*nl+l
 * (valence.F90:3401-3412); the potential already carries -Z (valence.F90:3160).
 *
 * The algorithm is deliberately different from the product's CUDA kernels
 * (Obara-Saika/HGP), so agreement between the two is independent evidence.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vo_internal.h"

#define LMAX 3               /* per shell */
#define L2 (2 * LMAX + 3)    /* 1D E table extent (kinetic needs j+2) */
#define L4 (4 * LMAX + 1)

static const double PI = 3.14159265358979323846264338327950288;

/* ---- Boys function F_m(T), m = 0..mmax ---------------------------------- */
void vo_boys(int mmax, double T, double *F)
{
    if (T < 35.0) {
        /* F_m(T) = exp(-T) sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1)), all terms
         * positive; then downward recursion F_{m-1} = (2T F_m + e^-T)/(2m-1) */
        double eT = exp(-T);
        double term = 1.0 / (2.0 * mmax + 1.0), sum = term;
        for (int k = 1; k < 400; ++k) {
            term *= 2.0 * T / (2.0 * mmax + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-18 * sum) break;
        }
        F[mmax] = eT * sum;
        for (int m = mmax; m > 0; --m) F[m - 1] = (2.0 * T * F[m] + eT) / (2.0 * m - 1.0);
    } else {
        /* erf form for F_0, upward recursion (stable for T >> m) */
        double eT = exp(-T), st = sqrt(T);
        F[0] = 0.5 * sqrt(PI) / st * erf(st);
        for (int m = 0; m < mmax; ++m) F[m + 1] = ((2.0 * m + 1.0) * F[m] - eT) / (2.0 * T);
    }
}

/* ---- cartesian components, CCA order ------------------------------------ */
int vo_ncart(int l) { return (l + 1) * (l + 2) / 2; }

void vo_cart(int l, int idx, int *lx, int *ly, int *lz)
{
    int n = 0;
    for (int i = 0; i <= l; ++i)
        for (int j = 0; j <= i; ++j) {
            if (n == idx) { *lx = l - i; *ly = i - j; *lz = j; return; }
            ++n;
        }
    *lx = *ly = *lz = 0;
}

/* ---- 1D Hermite expansion coefficients E[i][j][t] (without exp factor) --- */
static void hermite_E(int imax, int jmax, double p, double PA, double PB, double E[L2][L2][2 * L2])
{
    double h = 0.5 / p;
    memset(E, 0, sizeof(double) * L2 * L2 * 2 * L2);
    E[0][0][0] = 1.0;
    for (int i = 0; i <= imax; ++i) {
        if (i > 0)
            for (int t = 0; t <= i; ++t) {
                double v = PA * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
                if (t > 0) v += h * E[i - 1][0][t - 1];
                E[i][0][t] = v;
            }
        for (int j = 1; j <= jmax; ++j)
            for (int t = 0; t <= i + j; ++t) {
                double v = PB * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
                if (t > 0) v += h * E[i][j - 1][t - 1];
                E[i][j][t] = v;
            }
    }
}

/* ---- Hermite Coulomb integrals R_{tuv} = R^(0)_{tuv}, t+u+v <= L ---------- */
static void hermite_R(int L, double alpha, const double PQ[3], double R[L4][L4][L4])
{
    static __thread double Rn[L4 + 1][L4][L4][L4];
    double F[L4 + 1];
    double T = alpha * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
    vo_boys(L, T, F);
    double f = 1.0;
    for (int n = 0; n <= L; ++n) { Rn[n][0][0][0] = f * F[n]; f *= -2.0 * alpha; }
    for (int N = 1; N <= L; ++N)              /* total order t+u+v = N */
        for (int n = 0; n <= L - N; ++n)
            for (int t = 0; t <= N; ++t)
                for (int u = 0; u <= N - t; ++u) {
                    int v = N - t - u;
                    double val;
                    if (t > 0) {
                        val = PQ[0] * Rn[n + 1][t - 1][u][v];
                        if (t > 1) val += (t - 1) * Rn[n + 1][t - 2][u][v];
                    } else if (u > 0) {
                        val = PQ[1] * Rn[n + 1][t][u - 1][v];
                        if (u > 1) val += (u - 1) * Rn[n + 1][t][u - 2][v];
                    } else {
                        val = PQ[2] * Rn[n + 1][t][u][v - 1];
                        if (v > 1) val += (v - 1) * Rn[n + 1][t][u][v - 2];
                    }
                    Rn[n][t][u][v] = val;
                }
    for (int t = 0; t <= L; ++t)
        for (int u = 0; u <= L - t; ++u)
            for (int v = 0; v <= L - t - u; ++v) R[t][u][v] = Rn[0][t][u][v];
}

/* ---- one-electron blocks ------------------------------------------------- */
void vo_overlap_block(const vo_shell *A, const vo_shell *B, double *out)
{
    int na = vo_ncart(A->l), nb = vo_ncart(B->l);
    double AB2 = 0.0;
    for (int d = 0; d < 3; ++d) AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
    for (int n = 0; n < na * nb; ++n) out[n] = 0.0;
    static __thread double E[3][L2][L2][2 * L2];
    for (int ia = 0; ia < A->nprim; ++ia)
        for (int ib = 0; ib < B->nprim; ++ib) {
            double a = A->exps[ia], b = B->exps[ib], p = a + b, mu = a * b / p;
            double pref = A->coef[ia] * B->coef[ib] * exp(-mu * AB2) * pow(PI / p, 1.5);
            for (int d = 0; d < 3; ++d) {
                double P = (a * A->r[d] + b * B->r[d]) / p;
                hermite_E(A->l, B->l, p, P - A->r[d], P - B->r[d], E[d]);
            }
            for (int i = 0; i < na; ++i) {
                int ax, ay, az; vo_cart(A->l, i, &ax, &ay, &az);
                for (int j = 0; j < nb; ++j) {
                    int bx, by, bz; vo_cart(B->l, j, &bx, &by, &bz);
                    out[i * nb + j] += pref * E[0][ax][bx][0] * E[1][ay][by][0] * E[2][az][bz][0];
                }
            }
        }
}

void vo_kinetic_block(const vo_shell *A, const vo_shell *B, double *out)
{
    int na = vo_ncart(A->l), nb = vo_ncart(B->l);
    double AB2 = 0.0;
    for (int d = 0; d < 3; ++d) AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
    for (int n = 0; n < na * nb; ++n) out[n] = 0.0;
    static __thread double E[3][L2][L2][2 * L2];
    for (int ia = 0; ia < A->nprim; ++ia)
        for (int ib = 0; ib < B->nprim; ++ib) {
            double a = A->exps[ia], b = B->exps[ib], p = a + b, mu = a * b / p;
            double pref = A->coef[ia] * B->coef[ib] * exp(-mu * AB2) * pow(PI / p, 1.5);
            for (int d = 0; d < 3; ++d) {
                double P = (a * A->r[d] + b * B->r[d]) / p;
                hermite_E(A->l, B->l + 2, p, P - A->r[d], P - B->r[d], E[d]);
            }
            for (int i = 0; i < na; ++i) {
                int al[3]; vo_cart(A->l, i, &al[0], &al[1], &al[2]);
                for (int j = 0; j < nb; ++j) {
                    int bl[3]; vo_cart(B->l, j, &bl[0], &bl[1], &bl[2]);
                    double S[3], T[3];
                    for (int d = 0; d < 3; ++d) {
                        int ii = al[d], jj = bl[d];
                        S[d] = E[d][ii][jj][0];
                        /* -1/2 d^2/dx^2 on the ket 1D Gaussian */
                        T[d] = -2.0 * b * b * E[d][ii][jj + 2][0] + b * (2.0 * jj + 1.0) * E[d][ii][jj][0];
                        if (jj >= 2) T[d] -= 0.5 * jj * (jj - 1) * E[d][ii][jj - 2][0];
                    }
                    out[i * nb + j] += pref * (T[0] * S[1] * S[2] + S[0] * T[1] * S[2] + S[0] * S[1] * T[2]);
                }
            }
        }
}

/* potential of one point charge Z at C; result carries -Z */
void vo_potential_block(const vo_shell *A, const vo_shell *B, double Z, const double C[3], double *out)
{
    int na = vo_ncart(A->l), nb = vo_ncart(B->l), L = A->l + B->l;
    double AB2 = 0.0;
    for (int d = 0; d < 3; ++d) AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
    for (int n = 0; n < na * nb; ++n) out[n] = 0.0;
    static __thread double E[3][L2][L2][2 * L2];
    static __thread double R[L4][L4][L4];
    for (int ia = 0; ia < A->nprim; ++ia)
        for (int ib = 0; ib < B->nprim; ++ib) {
            double a = A->exps[ia], b = B->exps[ib], p = a + b, mu = a * b / p;
            double pref = -Z * A->coef[ia] * B->coef[ib] * exp(-mu * AB2) * 2.0 * PI / p;
            double PC[3];
            for (int d = 0; d < 3; ++d) {
                double P = (a * A->r[d] + b * B->r[d]) / p;
                hermite_E(A->l, B->l, p, P - A->r[d], P - B->r[d], E[d]);
                PC[d] = P - C[d];
            }
            hermite_R(L, p, PC, R);
            for (int i = 0; i < na; ++i) {
                int ax, ay, az; vo_cart(A->l, i, &ax, &ay, &az);
                for (int j = 0; j < nb; ++j) {
                    int bx, by, bz; vo_cart(B->l, j, &bx, &by, &bz);
                    double s = 0.0;
                    for (int t = 0; t <= ax + bx; ++t)
                        for (int u = 0; u <= ay + by; ++u)
                            for (int v = 0; v <= az + bz; ++v)
                                s += E[0][ax][bx][t] * E[1][ay][by][u] * E[2][az][bz][v] * R[t][u][v];
                    out[i * nb + j] += pref * s;
                }
            }
        }
}

/* ---- two-electron block (ab|cd), no primitive screening (tolerance 0.0 at
 * valence.F90:3398) ------------------------------------------------------- */
void vo_eri_block(const vo_shell *A, const vo_shell *B, const vo_shell *C, const vo_shell *D, double *out)
{
    int na = vo_ncart(A->l), nb = vo_ncart(B->l), nc = vo_ncart(C->l), nd = vo_ncart(D->l);
    int Lab = A->l + B->l, Lcd = C->l + D->l, L = Lab + Lcd;
    int nab = na * nb, ncd = nc * nd;
    double AB2 = 0.0, CD2 = 0.0;
    for (int d = 0; d < 3; ++d) {
        AB2 += (A->r[d] - B->r[d]) * (A->r[d] - B->r[d]);
        CD2 += (C->r[d] - D->r[d]) * (C->r[d] - D->r[d]);
    }
    for (int n = 0; n < nab * ncd; ++n) out[n] = 0.0;
    static __thread double Eab[3][L2][L2][2 * L2], Ecd[3][L2][L2][2 * L2];
    static __thread double R[L4][L4][L4];
    /* bra Hermite coefficients per component, per primitive pair */
    int nhab = (Lab + 1) * (Lab + 1) * (Lab + 1);
    double *hab = (double *)malloc(sizeof(double) * (size_t)nab * nhab);
    int lxa[36], lya[36], lza[36], lxb[36], lyb[36], lzb[36];
    for (int i = 0; i < na; ++i) vo_cart(A->l, i, &lxa[i], &lya[i], &lza[i]);
    for (int j = 0; j < nb; ++j) vo_cart(B->l, j, &lxb[j], &lyb[j], &lzb[j]);
    int lxc[36], lyc[36], lzc[36], lxd[36], lyd[36], lzd[36];
    for (int k = 0; k < nc; ++k) vo_cart(C->l, k, &lxc[k], &lyc[k], &lzc[k]);
    for (int l = 0; l < nd; ++l) vo_cart(D->l, l, &lxd[l], &lyd[l], &lzd[l]);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)nhab);

    for (int ia = 0; ia < A->nprim; ++ia)
        for (int ib = 0; ib < B->nprim; ++ib) {
            double a = A->exps[ia], b = B->exps[ib], p = a + b;
            double Kab = A->coef[ia] * B->coef[ib] * exp(-a * b / p * AB2);
            double P[3];
            for (int d = 0; d < 3; ++d) {
                P[d] = (a * A->r[d] + b * B->r[d]) / p;
                hermite_E(A->l, B->l, p, P[d] - A->r[d], P[d] - B->r[d], Eab[d]);
            }
            for (int i = 0; i < na; ++i)
                for (int j = 0; j < nb; ++j) {
                    double *h = hab + (size_t)(i * nb + j) * nhab;
                    for (int t = 0; t <= Lab; ++t)
                        for (int u = 0; u <= Lab; ++u)
                            for (int v = 0; v <= Lab; ++v)
                                h[(t * (Lab + 1) + u) * (Lab + 1) + v] =
                                    (t <= lxa[i] + lxb[j] && u <= lya[i] + lyb[j] && v <= lza[i] + lzb[j])
                                        ? Eab[0][lxa[i]][lxb[j]][t] * Eab[1][lya[i]][lyb[j]][u] * Eab[2][lza[i]][lzb[j]][v]
                                        : 0.0;
                }
            for (int ic = 0; ic < C->nprim; ++ic)
                for (int id = 0; id < D->nprim; ++id) {
                    double c = C->exps[ic], dd = D->exps[id], q = c + dd;
                    double Kcd = C->coef[ic] * D->coef[id] * exp(-c * dd / q * CD2);
                    double PQ[3];
                    for (int d = 0; d < 3; ++d) {
                        double Q = (c * C->r[d] + dd * D->r[d]) / q;
                        hermite_E(C->l, D->l, q, Q - C->r[d], Q - D->r[d], Ecd[d]);
                        PQ[d] = P[d] - Q;
                    }
                    double alpha = p * q / (p + q);
                    double pref = 2.0 * pow(PI, 2.5) / (p * q * sqrt(p + q)) * Kab * Kcd;
                    hermite_R(L, alpha, PQ, R);
                    for (int k = 0; k < nc; ++k)
                        for (int l = 0; l < nd; ++l) {
                            int mx = lxc[k] + lxd[l], my = lyc[k] + lyd[l], mz = lzc[k] + lzd[l];
                            /* tmp[tuv] = sum_{tau nu phi} (-1)^(tau+nu+phi) Ecd R_{t+tau,u+nu,v+phi} */
                            for (int t = 0; t <= Lab; ++t)
                                for (int u = 0; u <= Lab - t; ++u)
                                    for (int v = 0; v <= Lab - t - u; ++v) {
                                        double s = 0.0;
                                        for (int tau = 0; tau <= mx; ++tau)
                                            for (int nu = 0; nu <= my; ++nu)
                                                for (int phi = 0; phi <= mz; ++phi) {
                                                    double e = Ecd[0][lxc[k]][lxd[l]][tau] * Ecd[1][lyc[k]][lyd[l]][nu] *
                                                               Ecd[2][lzc[k]][lzd[l]][phi];
                                                    if ((tau + nu + phi) & 1) e = -e;
                                                    s += e * R[t + tau][u + nu][v + phi];
                                                }
                                        tmp[(t * (Lab + 1) + u) * (Lab + 1) + v] = s;
                                    }
                            for (int ij = 0; ij < nab; ++ij) {
                                const double *h = hab + (size_t)ij * nhab;
                                double s = 0.0;
                                for (int t = 0; t <= Lab; ++t)
                                    for (int u = 0; u <= Lab - t; ++u)
                                        for (int v = 0; v <= Lab - t - u; ++v) {
                                            int x = (t * (Lab + 1) + u) * (Lab + 1) + v;
                                            s += h[x] * tmp[x];
                                        }
                                out[(size_t)ij * ncd + k * nd + l] += pref * s;
                            }
                        }
                }
        }
    free(tmp);
    free(hab);
}
