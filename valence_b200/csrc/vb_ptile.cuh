// The fused hot kernel for s/p basis sets (every class up to (pp|pp)), primitive-level formulation.
//
// For one tile (bra pair group P, ket pair group Q) the orbital-level integrals are
//     G[q][p] = sum_{k,f} sum_{b,e}  Dq[e_k + f][q] * [b,e | k,f] * Dp[e_b + e][p]
// with b / k running over the *primitive* pairs of P / Q (flat lists per pair type, sorted by
// magnitude) and Dp / Dq the HRR-folded pair densities of the primitive's shell pair.  Written like
// this the contraction of primitives into shells, and both density transformations, are two dense
// matrix products around the matrix of primitive integrals -- and they run on the FP64 tensor cores:
//
//   lane (g = lane/4, t = lane%4) of a warp owns ket primitive g of an octet and bra shell pair t of a
//   quad (shell pairs sorted by contraction length, so the four lanes run nearly equal trip counts):
//   it sums the Obara-Saika values of its ket primitive against the primitives of its bra shell pair
//   and feeds each component straight into DMMA m8n8k4 as the A fragment A[g][t];
//   B[t][n] = Dp[e_sp + e][n] comes from the staged bra densities, and X_f[k][p] (8 ket primitives x
//   32 orbital pairs) accumulates in the C fragments while the warp walks all bra shell pairs.
//   No shuffles, no segmented reductions; pruning by magnitude is per lane.  (FP64 DMMA shares the
//   FP64 pipe with DFMA on B200 -- one m8n8k4 costs 8 DFMA issue slots -- hence the contraction over
//   the bra primitives happens in the lane *before* the tensor-core product.)
//   After the bra loop a second DMMA product folds the 8 ket primitives with Dq into G.
//
// (replaces int2e, valence.F90:3184-3438, and the 2e loop of vsvb_energy, :1153-1433; the screening
// and bookkeeping of the contraction phase follow the reference line by line, see below)
#pragma once
#include "vb_tile.cuh"

namespace vb {

// The classes are split over two launches so that each gets the register budget it needs:
//   PART_HEAVY : every class with a pp pair on either side (up to 81 components per quartet), 8 warps x 255 registers,
//                leaves its share of G in a global buffer;
//   PART_LIGHT : (ss|ss) (ps|ss) (ss|ps) (ps|ps) -- 77 % of the flops, 90 % of the primitive quartets of a water
//                cluster -- 16 warps x 128 registers, starts from the heavy share and runs the contraction;
//   PART_ALL   : everything in one launch (Schwarz pass).
enum { PART_ALL = 0, PART_HEAVY = 1, PART_LIGHT = 2 };
#ifndef VB_PT_LIGHT_THREADS
#define VB_PT_LIGHT_THREADS 512
#endif
#ifndef VB_PT_ALL_THREADS
#define VB_PT_ALL_THREADS 384
#endif
__host__ __device__ constexpr int pt_threads(int part)
{
    return part == PART_LIGHT ? VB_PT_LIGHT_THREADS : (part == PART_ALL ? VB_PT_ALL_THREADS : 256);
}
constexpr int PT_MAX_WARPS = (VB_PT_LIGHT_THREADS > VB_PT_ALL_THREADS ? VB_PT_LIGHT_THREADS : VB_PT_ALL_THREADS) / 32;
__host__ __device__ constexpr bool pt_class_in_part(int part, int tb, int tk)
{
    return part == PART_ALL || ((tb < 2 && tk < 2) == (part == PART_LIGHT));
}
#ifndef VB_EXP_SKIP
#define VB_EXP_SKIP 0     // experiment builds only: 1 no contraction, 2 no second transform / G reduction, 4 no first transform, 8 no integrals
#endif
constexpr int PT_SLD = 33;                      // row stride of the per-warp X scratch (8 x 32 doubles)
constexpr int PT_SCRATCH = 8 * PT_SLD + 1;      // doubles per warp (odd row stride, even total keeps the next region aligned)

// one primitive quartet per lane: acc += [e0|f0] for e in the kept bra range, f in the kept ket range
template <int TB, int TK>
__device__ __forceinline__ void quartet_values(const double* __restrict__ boys_tab, const PrimPair& a, const PrimPair& b,
                                               double (&acc)[pt_ne(TB) * pt_ne(TK)])
{
    constexpr int LA = pt_la(TB), EA = pt_E(TB), LC = pt_la(TK), EC = pt_E(TK), M = EA + EC;
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    if constexpr (M == 0) {
        // (ss|ss): the far field needs one reciprocal square root.  With s = p q |PQ|^2 and T = s / (p + q):
        // T >= Tmax  <=>  s >= Tmax (p + q), and  pref F_0(T) = Ka Kb (p+q)^-1/2 * (sqrt(pi)/2) T^-1/2 = Ka Kb (sqrt(pi)/2) s^-1/2
        const double u = a.p + b.p;
        const double dx = a.Px - b.Px, dy = a.Py - b.Py, dz = a.Pz - b.Pz;
        const double s = a.p * b.p * (dx * dx + dy * dy + dz * dz);
        if (s >= BOYS_S_TMAX * u) {
            acc[0] += a.Kp * b.Kp * 0.88622692545275801365 * rsqrt(s);
            return;
        }
    }
    QuartetGeom geo;
    double T, pref;
    quartet_geom(a, b, geo, T, pref);
    double F[M + 1];
    boys_s<M>(boys_tab, T, F);
#pragma unroll
    for (int m = 0; m <= M; ++m) F[m] *= pref;
    if constexpr (M == 0) acc[0] += F[0];
    else vrr_unrolled<LA, EA, LC, EC>(geo, F, acc);
}

// the quartet values go straight into the tensor cores: X_f[k][p] += sum_b A[k][b] Dp[e_b + e][p]
template <int TB, int TK>
__device__ __forceinline__ void feed_dmma(const double (&acc)[pt_ne(TB) * pt_ne(TK)], int eoff, const double* __restrict__ Dp_s,
                                          int npP, int g, double (&X)[pt_ne(TK)][4][2])
{
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    const double* drow = Dp_s + eoff * npP + g;        // B[t][n = 8j + g] = Dp[e_sp + e][8j + g]
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double bf = drow[e * npP + 8 * j];
#pragma unroll
            for (int f = 0; f < NF; ++f) dmma_884(X[f][j][0], X[f][j][1], acc[e * NF + f], bf);
        }
}

// One warp task: a ket octet of pair type TK against every bra primitive of P.
template <int PART, int TK>
__device__ __forceinline__ void ptask(const TileArgs& A, const PGDesc& P, const PGDesc& Q, int oct, const PrimPair* __restrict__ bpps,
                                      const SPRec* __restrict__ spss,
                                      const double* __restrict__ Dp_s, const double* __restrict__ Dq_g, const double* __restrict__ boys_tab,
                                      double* __restrict__ scratch, double* __restrict__ Gw, double* __restrict__ Gglob, bool priv, int lane,
                                      unsigned long long* __restrict__ s_pq)
{
    constexpr int NF = pt_ne(TK);
    const int g = lane >> 2, t = lane & 3;
    const int nk = Q.pp_beg[TK + 1] - Q.pp_beg[TK];
    const PrimPair* __restrict__ kl = A.pps_flat + Q.pp_beg[TK];
    const int k0 = 8 * oct;
    const bool kact = k0 + g < nk;
    PrimPair b = kl[k0 + (kact ? g : 0)];
    if (!kact) b.Kp = 0.0;
    const double wk = kl[k0].w;                        // the octet's largest magnitude (lists are sorted)
    double X[NF][4][2];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) { X[f][j][0] = 0.0; X[f][j][1] = 0.0; }
    bool any = false;
    sfor<0, 3>([&](auto TBc) {
        constexpr int TB = TBc;
        constexpr int NE = pt_ne(TB);
        if constexpr (!pt_class_in_part(PART, TB, TK)) return;
        const int nsp = P.sp_beg[TB + 1] - P.sp_beg[TB];
        const SPRec* __restrict__ sl = spss + (P.sp_beg[TB] - P.sp_beg[0]);
        if (!(nsp > 0 && P.kwmax[TB] * wk >= A.tau)) return;
        unsigned nq = 0;                                // primitive quartets evaluated by this lane
        for (int q0 = 0; q0 < nsp; q0 += 4) {
            const bool sact = q0 + t < nsp;
            const SPRec sp = sl[q0 + (sact ? t : 0)];
            const int cnt = sact ? sp.pp_cnt : 0;
            const PrimPair* __restrict__ bl = bpps + (sp.pp_beg - P.pp_beg[0]);
            double acc[NE * NF];
#pragma unroll
            for (int i = 0; i < NE * NF; ++i) acc[i] = 0.0;
            int ip = 0;
            // (processing two primitives per trip for more instruction-level parallelism was measured 20-35 % slower:
            //  the kernel is bound by instruction supply -- 9 unrolled class bodies, 12 warps in different ones -- not
            //  by dependency latency)
            for (;; ++ip) {
                // primitives of a shell pair are sorted by magnitude: a lane that stops stays stopped
                PrimPair a = bl[ip < cnt ? ip : 0];
                const bool act = ip < cnt && a.w * wk >= A.tau;
                if (!__any_sync(0xffffffffu, act)) break;
                if (!act) a.Kp = 0.0;
                if ((VB_EXP_SKIP & 8) && A.mode == 1) acc[0] += a.Kp * b.Kp;
                else quartet_values<TB, TK>(boys_tab, a, b, acc);
                nq += (act && kact) ? 1u : 0u;
            }
            if ((VB_EXP_SKIP & 4) && A.mode == 1) {
                double sacc = 0.0;
#pragma unroll
                for (int i = 0; i < NE * NF; ++i) sacc += acc[i];
                X[0][0][0] += sacc;
            } else if (ip > 0) feed_dmma<TB, TK>(acc, sp.eoff, Dp_s, P.np, g, X);
        }
        for (int o = 16; o > 0; o >>= 1) nq += __shfl_xor_sync(0xffffffffu, nq, o);
        if (nq) {
            any = true;
            if (lane == 0) atomicAdd(&s_pq[TB * NPTYPE + TK], (unsigned long long)nq);
        }
    });
    if (!any) return;
    if ((VB_EXP_SKIP & 2) && A.mode == 1) {
        double sx = 0.0;
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) sx += X[f][j][0] + X[f][j][1];
        if (sx == 1.2345e301) scratch[lane] = sx;
        return;
    }
    // G[q][p] += sum_{k,f} Dq[e_k + f][q] X_f[k][p]:  A'[q][k] from the staged ket densities, B'[k][p] = X_f
    // re-laid out through the warp's scratch (C fragment -> B fragment).
    double C[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) { C[m][j][0] = 0.0; C[m][j][1] = 0.0; }
    const int eo_own = b.eoff;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            scratch[g * PT_SLD + 8 * j + 2 * t] = X[f][j][0];
            scratch[g * PT_SLD + 8 * j + 2 * t + 1] = X[f][j][1];
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int kk = 4 * s + t;                                   // ket primitive this lane supplies
            const int eo = __shfl_sync(0xffffffffu, eo_own, 4 * kk);    // lane 4*kk evaluates primitive kk
            double bfr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bfr[j] = scratch[kk * PT_SLD + 8 * j + g];
            const double* arow = Dq_g + (eo + f) * Q.np;                // read once per task: straight from L2
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int q = 8 * m + g;
                const double af = q < Q.np ? arow[q] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(C[m][j][0], C[m][j][1], af, bfr[j]);
            }
        }
    }
    // Schwarz pass: warp-private partials (fixed order inside a warp; the warps are summed in fixed order
    // later); energy pass: FP64 reductions into the tile's G in the CTA's global slice (RED.ADD.F64 at L2)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int q = 8 * m + g;
        if (q >= Q.np) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int p = 8 * j + 2 * t + i;
                if (p < P.np) {
                    if (priv) Gw[q * P.np + p] += C[m][j][i];
                    else atomicAdd(&Gglob[q * P.np + p], C[m][j][i]);
                }
            }
    }
}

// Contraction of one tile's orbital-level integrals G[q][p] with the cofactor densities, with the reference's
// screening and task bookkeeping line by line (valence.F90:1153-1433).  Threads of the CTA stride over the
// tile's entries; returns this thread's share of the tile energy and adds to its counters.
struct TileIdx { int np, pair_beg; };     // what the contraction needs of a pair group
// SYM = A.sym as a compile-time constant: the image loops unroll and the image table stays in registers.
// sigma * W'(a,b,c,d) for P'_alpha = Q + x y^T / sigma (the 1/sigma^2 terms cancel exactly, so this stays finite for a
// singular substituted block):  sigma W0 + x_a y_b T_cd + x_c y_d T_ab - x_a y_d Q_cb - x_c y_b Q_ad,  T = Q + R
__device__ __forceinline__ double w_rank1(const TileArgs& A, int nso, int a, int b, int c, int d)
{
    const double* __restrict__ Q = A.Pa;
    const double* __restrict__ R = A.Pb;
    const double qab = Q[a * nso + b], qcd = Q[c * nso + d], qad = Q[a * nso + d], qcb = Q[c * nso + b];
    const double rab = R[a * nso + b], rcd = R[c * nso + d], rad = R[a * nso + d], rcb = R[c * nso + b];
    const double w0 = __dmul_rn(qab + rab, qcd + rcd) - __dmul_rn(qad, qcb) - __dmul_rn(rad, rcb);
    const double xa = A.r1x[a], xc = A.r1x[c], yb = A.r1y[b], yd = A.r1y[d];
    return A.r1sigma * w0 + xa * yb * (qcd + rcd) + xc * yd * (qab + rab) - xa * yd * qcb - xc * yb * qad;
}

// FMODE (first_order_opt, tiles free of the subject entry): besides the energy with the base densities (W0), the
// Fock-like matrix F[s][t] with  sum_entries val * (bilinear part of sigma W') = x^T F y  is accumulated, so that every
// (ib,jb) element of ham needs no pass over these tiles at all.
template <int THREADS, bool SYM, bool FMODE, class GFetch>
__device__ __forceinline__ double contract_tile(const TileArgs& A, const TileIdx P, const TileIdx Q, GFetch&& Gval,
                                                int tid, unsigned long long (&cnt)[CNT_N])
{
    constexpr int NIM = SYM ? 8 : 2;
    const int nso = A.nso;
    const bool diag_tile = P.pair_beg == Q.pair_beg;
    double epart = 0.0;
    for (int idx = tid; idx < P.np * Q.np; idx += THREADS) {
        const int p = idx / Q.np, q = idx % Q.np;
        if (diag_tile && q > p) continue;
        const int s = A.pg_pairs[2 * (P.pair_beg + p)], t = A.pg_pairs[2 * (P.pair_beg + p) + 1];
        const int u = A.pg_pairs[2 * (Q.pair_beg + q)], v = A.pg_pairs[2 * (Q.pair_beg + q) + 1];
        if (A.mode == 0) {
            if (diag_tile && p == q) {
                const double G = Gval(p, q);
                A.diag[(size_t)s * nso + t] = G;
                if (SYM) A.diag[(size_t)t * nso + s] = G;
            }
            continue;
        }
        // reference screen on the Schwarz product (valence.F90:1189-1190); screened entries
        // contribute nothing and are counted nowhere
        const bool ssig = A.sch[s * nso + t] * A.sch[u * nso + v] > A.itol;
        if (!ssig) continue;
        const double G = Gval(p, q);
        // images of (s,t,u,v) under the integral's permutational symmetry
        int im[NIM][4];
        {
            const int base4[2][4] = {{s, t, u, v}, {u, v, s, t}};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int a = base4[k][0], b = base4[k][1], c = base4[k][2], d = base4[k][3];
                constexpr int W = SYM ? 4 : 1;
                im[W * k][0] = a; im[W * k][1] = b; im[W * k][2] = c; im[W * k][3] = d;
                if constexpr (SYM) {
                    im[4 * k + 1][0] = b; im[4 * k + 1][1] = a; im[4 * k + 1][2] = c; im[4 * k + 1][3] = d;
                    im[4 * k + 2][0] = a; im[4 * k + 2][1] = b; im[4 * k + 2][2] = d; im[4 * k + 2][3] = c;
                    im[4 * k + 3][0] = b; im[4 * k + 3][1] = a; im[4 * k + 3][2] = d; im[4 * k + 3][3] = c;
                }
            }
        }
        double wsum = 0.0;
        cnt[CNT_ENTRIES]++;
        const bool e_st = s == t, e_uv = u == v, e_pq = s == u && t == v, e_px = s == v && t == u;
#pragma unroll
        for (int k = 0; k < NIM; ++k) {
            const int a = im[k][0], b = im[k][1], c = im[k][2], d = im[k][3];
            // an image repeats an earlier one only through s == t, u == v, (s,t) == (u,v) or (s,t) == (v,u)
            bool dup;
            if constexpr (SYM) {
                dup = (k == 1 && e_st) || (k == 2 && e_uv) || (k == 3 && (e_st || e_uv)) || (k == 4 && (e_pq || e_px)) ||
                      (k == 5 && (e_uv || e_pq || e_px)) || (k == 6 && (e_st || e_pq || e_px)) || (k == 7 && (e_st || e_uv || e_pq || e_px));
            } else {
                dup = k == 1 && e_pq;
            }
            if (dup) continue;
            // task bookkeeping exactly as the reference visits it (valence.F90:1167-1190)
            const bool shortcut = (a == c && b == d) && a != A.subject && b != A.subject;   // :1213 (nonsub)
            const double val = shortcut ? A.sch[a * nso + b] * A.sch[a * nso + b] : G;
            const bool vsig = fabs(val) > A.itol;
            // as the direct integral of task io=a, ko=b, jo=c, lo=d
            bool vd = a >= c && b >= d && !((a == c && a < A.nnd) || (b == d && b < A.nnd));
            if (vd && SYM) vd = tri_index(a, c) >= tri_index(b, d);
            // as the exchanged integral of task io=a, lo=b, jo=c, ko=d
            bool vx = a >= c && d >= b && !((a == c && a < A.nnd) || (b == d && b < A.nnd));
            if (vx && SYM) vx = tri_index(a, c) >= tri_index(d, b);
            if (vd) { cnt[CNT_SCHWARZ_EREP]++; cnt[CNT_VALUE_EREP] += vsig; }
            if (vx) { cnt[CNT_SCHWARZ_EXCH]++; cnt[CNT_VALUE_EXCH] += vsig; }
            if (!shortcut) {
                const int calls = (vd ? 1 : 0) + ((vx && b != d) ? 1 : 0);
                cnt[CNT_INT2E] += calls;
                cnt[CNT_SHELLQ] += (unsigned long long)calls * A.nsh_bra[a] * A.nsh_ket[b] * A.nsh_bra[c] * A.nsh_ket[d];
            }
            if (shortcut && vd) cnt[CNT_SHORTCUT]++;
            if (A.debug) printf("ENTRY (%d %d|%d %d) val %.12f W %.12f vsig %d\n", a, b, c, d, val, A.ndp ? w_general(A.cof, A.ndp, A.cof_stride, nso, a, b, c, d) : w_term(A.Pa, A.Pb, nso, a, b, c, d), (int)vsig);
            if (vsig) {
                if constexpr (FMODE) {
                    const double* __restrict__ Q = A.Pa;
                    const double* __restrict__ R = A.Pb;
                    const double qab = Q[a * nso + b], qcd = Q[c * nso + d], qad = Q[a * nso + d], qcb = Q[c * nso + b];
                    const double rab = R[a * nso + b], rcd = R[c * nso + d], rad = R[a * nso + d], rcb = R[c * nso + b];
                    wsum += val * (__dmul_rn(qab + rab, qcd + rcd) - __dmul_rn(qad, qcb) - __dmul_rn(rad, rcb));
                    const double h = 0.5 * val;
                    atomicAdd(&A.Fmat[a * nso + b], h * (qcd + rcd));
                    atomicAdd(&A.Fmat[c * nso + d], h * (qab + rab));
                    atomicAdd(&A.Fmat[a * nso + d], -h * qcb);
                    atomicAdd(&A.Fmat[c * nso + b], -h * qad);
                } else {
                    wsum += val * (A.r1 ? w_rank1(A, nso, a, b, c, d)
                                        : (A.ndp ? w_general(A.cof, A.ndp, A.cof_stride, nso, a, b, c, d) : w_term(A.Pa, A.Pb, nso, a, b, c, d)));
                }
            }
        }
        epart += 0.5 * wsum;
    }
    return epart;
}

constexpr int PT_MAXQ = 8;     // ket pair groups processed together against one staged bra pair group

// Persistent kernel.  One work item = a bra pair group P with up to PT_MAXQ ket pair groups (tiles
// (P,Q_i), consecutive in the tile list): P's densities and primitive pairs are staged once (TMA bulk),
// the ket octets of all Q_i form one task pool (heaviest pair types first), and the contraction phase
// runs over all of the item's tiles.
template <int PART>
__global__ void __launch_bounds__(pt_threads(PART), 1) k_ptile(const TileArgs A)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int PT_THREADS = pt_threads(PART);
    constexpr int nw = PT_THREADS / 32;
    double* Dp_s = smem;                                        // [P.ne][P.np]
    double* Gs = Dp_s + A.dq_cap;                               // G[q][p] per tile (energy pass) / per warp (Schwarz pass), 8 x g_cap
    double* scr = Gs + PT_MAXQ * A.g_cap;                       // per-warp X scratch
    double* boys_sm = scr + PT_MAX_WARPS * PT_SCRATCH;                    // compact Boys table (when it fits)
    SPRec* sps_s = reinterpret_cast<SPRec*>(boys_sm + A.boys_cap);           // bra shell pairs of P
    PrimPair* bpp_s = reinterpret_cast<PrimPair*>(sps_s + A.sp_cap);         // bra primitive pairs of P (when they fit)
    constexpr int nwp = nw < PT_MAXQ ? nw : PT_MAXQ;            // warps that take tasks in the Schwarz pass (the G region holds their partials)
    for (int i = threadIdx.x; i < A.boys_cap; i += PT_THREADS) boys_sm[i] = A.boys_small[i];
    const double* boys_tab = A.boys_cap ? boys_sm : A.boys_small;
    __shared__ unsigned long long s_bar;
    unsigned phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int s_item, s_unit;
    __shared__ int s_cum[3 * PT_MAXQ + 1];
    __shared__ double s_red[nw][PT_MAXQ];
    __shared__ unsigned long long s_cnt[CNT_N];
    __shared__ unsigned long long s_pq[NPTYPE * NPTYPE];
    __shared__ PGDesc s_P, s_Q[PT_MAXQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < CNT_N) s_cnt[tid] = 0ull;
    if (tid < NPTYPE * NPTYPE) s_pq[tid] = 0ull;
    const bool priv = A.mode == 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(A.counter, 1u);   // work stealing inside this rank's shard
        __syncthreads();
        const long long it = (long long)A.tile_first + (long long)s_item * A.tile_stride;
        if (it >= A.nitems) break;
        const int4 item = A.items[it];                          // first tile, # tiles, slot of the first tile in the G hand-over buffer
        const int tl0 = item.x, ntl = item.y;
        double* gbuf = A.gbuf + ((size_t)item.z - A.gslot_base) * A.g_cap;
        {
            // descriptors: P once, Q_i per tile (word-wise copy by the first warps)
            constexpr int W = sizeof(PGDesc) / 4;
            const int2 t0 = A.tiles[tl0];
            if (warp == 0)
                for (int i = lane; i < W; i += 32) reinterpret_cast<int*>(&s_P)[i] = reinterpret_cast<const int*>(A.pgs + t0.x)[i];
            for (int qi = warp; qi < ntl; qi += nw) {
                const int y = A.tiles[tl0 + qi].y;
                for (int i = lane; i < W; i += 32) reinterpret_cast<int*>(&s_Q[qi])[i] = reinterpret_cast<const int*>(A.pgs + y)[i];
            }
        }
        __syncthreads();
        const PGDesc& P = s_P;
        const int nbpp = P.pp_beg[NPTYPE] - P.pp_beg[0];
        if (tid == 0) {
            // the buffers were last read through the generic proxy (previous item): order before async writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bp = (unsigned)((((size_t)P.ne * P.np + 1) & ~(size_t)1) * sizeof(double));
            const unsigned bb = A.pp_cap ? (unsigned)(nbpp * sizeof(PrimPair)) : 0u;
            const unsigned bs = (unsigned)((P.sp_beg[NPTYPE] - P.sp_beg[0]) * sizeof(SPRec));
            mbar_expect_tx(&s_bar, bp + bb + bs);
            tma_bulk_g2s(Dp_s, A.dmat + P.d_off, bp, &s_bar);
            tma_bulk_g2s(sps_s, A.sps + P.sp_beg[0], bs, &s_bar);
            if (A.pp_cap) tma_bulk_g2s(bpp_s, A.pps + P.pp_beg[0], bb, &s_bar);
            // task pool: (pair type, tile, octet), heaviest pair type first so the tail of the item is made of light tasks
            int n = 0;
            const bool bra_pp = P.sp_beg[3] > P.sp_beg[2];
            for (int r = 0; r < 3; ++r)
                for (int qi = 0; qi < PT_MAXQ; ++qi) {
                    s_cum[r * PT_MAXQ + qi] = n;
                    const int tk = 2 - r;
                    bool wanted = qi < ntl;
                    if (PART == PART_LIGHT && tk == 2) wanted = false;
                    if (PART == PART_HEAVY && tk < 2 && !bra_pp) wanted = false;
                    if (wanted) {
                        const int nk = s_Q[qi].pp_beg[tk + 1] - s_Q[qi].pp_beg[tk];
                        if (nk > 0) n += (nk + 7) >> 3;
                    }
                }
            s_cum[3 * PT_MAXQ] = n;
            s_unit = 0;
        }
        const PrimPair* bpps = A.pp_cap ? bpp_s : A.pps + P.pp_beg[0];
        // energy pass: the tiles' G live in this CTA's slice of a global (L2-resident) buffer and are accumulated with
        // fire-and-forget FP64 reductions (native at L2; shared memory has no FP64 atomic add, only a CAS loop)
        double* Gg = A.gred + (size_t)blockIdx.x * PT_MAXQ * A.g_cap;
        if (priv) {
            for (int i = tid; i < nwp * A.g_cap; i += PT_THREADS) Gs[i] = 0.0;
        } else {
            for (int i = tid; i < ntl * A.g_cap; i += PT_THREADS) __stcg(&Gg[i], PART == PART_LIGHT ? gbuf[i] : 0.0);
            __threadfence();
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
        __syncthreads();

        // ---- primitive integrals + both density transformations (tensor cores) ----------------
        //   energy pass : tasks are handed out through a shared counter (dynamic), G through shared atomics;
        //   Schwarz pass: one tile per item, tasks dealt round-robin (static) into warp-private partials, so
        //                 the table is bitwise reproducible and identical on every rank (the tile list is
        //                 derived from it).
        const int nunits = s_cum[3 * PT_MAXQ];
        double* scratch = scr + warp * PT_SCRATCH;
        int ustat = warp - nwp;
        for (;;) {
            int u;
            if (!priv) {
                u = 0;
                if (lane == 0) u = atomicAdd(&s_unit, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
            } else {
                ustat += nwp;
                u = ustat;
                if (warp >= nwp) break;
            }
            if (u >= nunits) break;
            int c = 0;
            while (s_cum[c + 1] <= u) ++c;
            const int tk = 2 - c / PT_MAXQ, qi = c % PT_MAXQ, oct = u - s_cum[c];
            const PGDesc& Q = s_Q[qi];
            double* Gw = Gs + warp * A.g_cap;                 // Schwarz pass: warp-private partial
            double* Gglob = A.gred + ((size_t)blockIdx.x * PT_MAXQ + qi) * A.g_cap;   // energy pass: the tile's G in the CTA's global slice
            const double* Dq_g = A.dmat + Q.d_off;
            switch (tk) {
                case 0: ptask<PART, 0>(A, P, Q, oct, bpps, sps_s, Dp_s, Dq_g, boys_tab, scratch, Gw, Gglob, priv, lane, s_pq); break;
                case 1: ptask<PART, 1>(A, P, Q, oct, bpps, sps_s, Dp_s, Dq_g, boys_tab, scratch, Gw, Gglob, priv, lane, s_pq); break;
                default: ptask<PART, 2>(A, P, Q, oct, bpps, sps_s, Dp_s, Dq_g, boys_tab, scratch, Gw, Gglob, priv, lane, s_pq); break;
            }
        }
        if (!priv) __threadfence();
        __syncthreads();
        if (!priv)
            for (int i = tid; i < ntl * A.g_cap; i += PT_THREADS) Gs[i] = __ldcg(&Gg[i]);
        if (PART == PART_HEAVY || A.mode == 2) {
            // hand-over of the heavy classes' share / mode 2: the tiles' G go to the integral cache of first_order_opt
            for (int i = tid; i < ntl * A.g_cap; i += PT_THREADS) gbuf[i] = Gs[i];
            continue;
        }
        if (!priv) __syncthreads();
        if (priv) {
            // fixed-order sum of the warp partials -> Gs[0 .. gsz)
            const int gsz = P.np * s_Q[0].np;
            for (int i = tid; i < gsz; i += PT_THREADS) {
                double v = Gs[i];
#pragma unroll
                for (int w = 1; w < nwp; ++w) v += Gs[w * A.g_cap + i];
                Gs[i] = v;
            }
            __syncthreads();
        }

        // ---- contraction with the cofactor densities --------------------------------------
        unsigned long long cnt[CNT_N];
#pragma unroll
        for (int i = 0; i < CNT_N; ++i) cnt[i] = 0ull;
        for (int qi = 0; qi < (((VB_EXP_SKIP & 1) && A.mode == 1) ? 0 : ntl); ++qi) {
            const double* G_s = Gs + qi * A.g_cap;
            const int npP = P.np;
            const TileIdx Pi{P.np, P.pair_beg}, Qi{s_Q[qi].np, s_Q[qi].pair_beg};
            auto gv = [&](int p, int q) { return G_s[q * npP + p]; };
            double epart = A.sym ? contract_tile<PT_THREADS, true, false>(A, Pi, Qi, gv, tid, cnt) : contract_tile<PT_THREADS, false, false>(A, Pi, Qi, gv, tid, cnt);
            if (A.mode == 1) {
                for (int o = 16; o > 0; o >>= 1) epart += __shfl_down_sync(0xffffffffu, epart, o);
                if (lane == 0) s_red[warp][qi] = epart;
            }
        }
        if (A.mode == 1) {
            // deterministic block reduction of the tile energies; counters: one shared atomic per warp
#pragma unroll
            for (int i = 0; i < CNT_N; ++i) {
                unsigned long long c = cnt[i];
                for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
                if (lane == 0 && c) atomicAdd(&s_cnt[i], c);
            }
            __syncthreads();
            if (tid < ntl) {
                double e = 0.0;
                for (int w = 0; w < nw; ++w) e += s_red[w][tid];
                A.tileE[tl0 + tid] = e * A.c0;
            }
        }
    }
    // one flush per CTA (same-address global atomics per tile would serialise in L2)
    __syncthreads();
    if (tid < CNT_N && s_cnt[tid]) atomicAdd(&A.counters[tid], s_cnt[tid]);
    if (tid < NPTYPE * NPTYPE && s_pq[tid]) atomicAdd(&A.pq_counters[tid], s_pq[tid]);
}

// Contraction-only pass over cached integrals (first_order_opt, valence.F90:527-764: the (ib,jb) loop changes one
// orbital of the bra and of the ket, so every tile that does not touch that entry keeps its integrals -- the GPU
// counterpart of the reference's eribuf cache, valence.F90:1227-1273).  One variant tile per WARP, no barriers and no
// shared memory: the kernel is a chain of dependent gathers (tile record -> pairs -> Schwarz values -> G, densities),
// so it lives on the number of independent warps in flight.
//   vts[2i]   = (first pair of P, first pair of Q, np(P) | np(Q) << 8 | swap << 16, -)
//   vts[2i+1] = (cache slot of the canonical tile (low, high 32 bits), -, -)
//   swap: the canonical tile is stored as (canon(Q), canon(P));  perm[pair] = index, inside the canonical pair group,
//   of the pair (t,s) [identity for canonical pair groups].  Tile energies go to A.tileE[tile_base + i]; only entries
//   that pass the reference's Schwarz screen fetch their integral from the cache.
constexpr int CT_THREADS = 256;
__global__ void __launch_bounds__(CT_THREADS, 2) k_contract(const TileArgs A, const int4* __restrict__ vts, long long nvt,
                                                            const int* __restrict__ perm, const double* __restrict__ gcache,
                                                            long long tile_base)
{
    __shared__ unsigned long long s_cnt[CNT_N];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < CNT_N) s_cnt[tid] = 0ull;
    __syncthreads();
    unsigned long long cnt[CNT_N];
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) cnt[i] = 0ull;
    const long long nwarps = (long long)gridDim.x * (CT_THREADS / 32);
    for (long long it = (long long)blockIdx.x * (CT_THREADS / 32) + (tid >> 5); it < nvt; it += nwarps) {
        const int4 v0 = vts[2 * it], v1 = vts[2 * it + 1];
        const TileIdx P{v0.z & 0xff, v0.x}, Q{(v0.z >> 8) & 0xff, v0.y};
        const bool swp = (v0.z >> 16) & 1;
        const long long slot = (long long)(unsigned)v1.x | ((long long)v1.y << 32);
        const double* __restrict__ src = gcache + (size_t)slot * A.g_cap;
        auto gv = [&](int p, int q) {
            const int pc = perm[P.pair_beg + p], qc = perm[Q.pair_beg + q];
            return swp ? src[pc * Q.np + qc] : src[qc * P.np + pc];
        };
        double epart;
        if (A.Fmat) epart = contract_tile<32, false, true>(A, P, Q, gv, lane, cnt);      // first_order_opt: W0 energy + F accumulators
        else epart = A.sym ? contract_tile<32, true, false>(A, P, Q, gv, lane, cnt) : contract_tile<32, false, false>(A, P, Q, gv, lane, cnt);
        for (int o = 16; o > 0; o >>= 1) epart += __shfl_down_sync(0xffffffffu, epart, o);
        if (lane == 0) A.tileE[tile_base + it] = epart * A.c0;
    }
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) {
        unsigned long long c = cnt[i];
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if (lane == 0 && c) atomicAdd(&s_cnt[i], c);
    }
    __syncthreads();
    if (tid < CNT_N && s_cnt[tid]) atomicAdd(&A.counters[tid], s_cnt[tid]);
}

}  // namespace vb
