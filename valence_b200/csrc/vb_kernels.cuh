// CUDA kernels of the VSVB energy engine (sm_100a).
//
//   k_ao_1e        AO overlap and core-Hamiltonian matrices (replaces the SIMINT
//                  overlap/ke/potential calls of ovint/int1e, valence.F90:2988,3132,3151)
//   k_orb_1e       orbital-level <bra_s|ket_t>, <bra_s|h|ket_t> (ovint/int1e, :2891-3176)
//   k_gather_block / k_gj_inverse_grid / k_entry_density
//                  spin-block overlap matrices, their inverses and determinants
//                  (replaces density/det/givdr, :1535-2144, by the inverse form,
//                  SURVEY.md appendix B)
//   k_tile         fused shell-quartet ERI generation + transformation to the
//                  orbital-pair basis + contraction with the cofactor densities
//                  (replaces int2e :3184-3438 and the 2e loop of vsvb_energy :1153-1433)
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "vb_eri.cuh"
#include "vb_setup.h"

namespace vb {

enum { CNT_SCHWARZ_EREP = 0, CNT_SCHWARZ_EXCH, CNT_VALUE_EREP, CNT_VALUE_EXCH, CNT_INT2E, CNT_SHELLQ, CNT_SHORTCUT,
       CNT_ENTRIES, CNT_N };

struct DevShell {          // global shell on the device
    double x, y, z;
    int l, nprim, prim_off, ao_off;
};

struct TileArgs {
    const PGDesc* pgs;
    const int* pg_pairs;
    const SPRec* sps;
    const PrimPair* pps;             // grouped by shell pair (bra side)
    const PrimPair* pps_flat;        // k_ptile: per pair type sorted by magnitude across shell pairs (ket side)
    double tau;                      // primitive-quartet magnitude cut (0 = none)
    unsigned long long* pq_counters; // primitive quartets evaluated, per class tb*NPTYPE+tk
    const double* dmat;
    const double* boys;
    const int2* tiles;
    int ntiles;
    const int4* items;               // k_ptile work items: (first tile, # tiles, G slot, -) runs of tiles sharing the bra pair group
    int nitems;
    double* gred;                    // k_ptile energy pass: per-CTA G accumulation slices, grid x 8 x g_cap doubles
    double* gbuf;                    // hand-over of the heavy classes' share of G, g_cap doubles per tile slot
    long long gslot_base;            // slot of the first tile of the current chunk
    int tile_first, tile_stride;     // static block-cyclic shard of this rank; work stealing inside
    unsigned int* counter;
    int mode;                        // 0 = diagonal (Schwarz) pass, 1 = energy, 2 = export G
    int nso, nnd, sym, subject;
    int dq_cap;                      // doubles reserved for the staged ket density
    int hs_cap, pp_cap, sp_cap;      // shared-memory capacities: H tile (doubles), primitive pairs per side, shell pairs
    int hs_ld;                       // row stride of the H tile (k_tile)
    int g_cap;                       // k_ptile: doubles per warp-private G[q][p] partial
    const double* boys_small;        // compact Boys table (vb_eri.cuh, boys_s)
    int boys_cap;                    // BOYS_S_SIZE when the table is staged in shared memory, else 0
    double* diag;                    // mode 0: (s,t) -> (st|st)
    const double* sch;               // mode 1: Schwarz table as the reference indexes it, nso*nso
    double itol;
    const double* Pa;                // entry-level alpha / beta densities, nso*nso
    const double* Pb;
    double c0;                       // det(alpha block) * det(beta block) (fast path), 1 on the general path
    const double* cof;               // general path: packed cofactor data per determinant pair (vb_cofactor.h)
    int ndp;                         // # determinant pairs on the general path, 0 = fast path (Pa/Pb/c0)
    long long cof_stride;
    const int* nsh_bra;              // # shells passing the weight screen, per entry (bra / ket orbital)
    const int* nsh_ket;
    double* tileE;                   // mode 1: per-tile energy partial (deterministic reduction later)
    unsigned long long* counters;    // CNT_N
    // k_tile (d shells): the first half transformation of one tile split over hsplit CTAs.  hphase 1: CTA (tile, slice) takes the
    // units u = slice (mod hsplit) and leaves its share of the half-transformed tile in hpart[(tile * hsplit + slice) * hs_cap];
    // hphase 2: one CTA per tile sums the shares in slice order (deterministic) and contracts.  hsplit <= 1: one CTA does both.
    // A slice is (us, ks) with hsplit = US * hksplit: units u = us (mod US), ket batches b = ks (mod hksplit) of every unit.
    int hsplit, hphase, hksplit;
    double* hpart;
    double* gfull;                   // mode 2: dense G over ordered pairs
    const int* pair_index;           // mode 2: (s*nso+t) -> dense pair index, or -1
    int npairs_total;
    int debug;                       // print every contracted entry (tiny inputs only)
    // first_order_opt in rank-one form (vb_engine.cu: first_order_cached): the substituted alpha (or beta) block differs from
    // the block N without the subject entry by one bordered row / column, so sigma * P' = sigma * Q + x y^T (Q in Pa, the
    // other spin in Pb) and sigma * W' is sigma * W0 plus terms bilinear in (x, y)
    int r1;                          // != 0: W' of the entries is taken in rank-one form
    const double* r1x;               // x[s], bra side (from chi_ib), nso
    const double* r1y;               // y[t], ket side (from chi_jb), nso
    double r1sigma;                  // Schur complement <chi_ib|chi_jb> - r^T N^-1 c
    double* Fmat;                    // FMODE pass: F[s][t] accumulators (nso x nso), see contract_tile
    double* gen_scratch;             // generic (d-shell) path scratch, GEN_SCRATCH doubles per thread
};

// ------------------------------------------------------------------------------------------------
// one-electron AO matrices
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dev_binom(int n, int k)
{
    double r = 1.0;
    for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
    return r;
}

// Contracted nuclear-attraction auxiliaries sum_{prim pairs, nuclei} [e|C], e <= L on centre A, for one
// shell pair: lanes stride over the nuclei, the primitive pairs are walked by the whole warp.
template <int L>
__device__ __forceinline__ void nuc_aux(const DevShell& A, const DevShell& B, double AB2, const double* __restrict__ exps,
                                        const double* __restrict__ coefs, const double* __restrict__ nuc, int natom,
                                        const double* __restrict__ boys_tab, int lane, double kcut, double* __restrict__ out /* ncum(L) */)
{
    constexpr int NE = ncum(L);
    double acc[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) acc[e] = 0.0;
    for (int ia = 0; ia < A.nprim; ++ia)
        for (int ib = 0; ib < B.nprim; ++ib) {
            const double a = exps[A.prim_off + ia], b = exps[B.prim_off + ib], p = a + b, ip = 1.0 / p, h2p = 0.5 * ip;
            const double K = coefs[A.prim_off + ia] * coefs[B.prim_off + ib] * exp(-a * b * ip * AB2);
            if (!(fabs(K) > kcut)) continue;
            const double P[3] = {(a * A.x + b * B.x) * ip, (a * A.y + b * B.y) * ip, (a * A.z + b * B.z) * ip};
            const double PA[3] = {P[0] - A.x, P[1] - A.y, P[2] - A.z};
            const double pk = -K * 2.0 * PI * ip;
            for (int c = lane; c < natom; c += 32) {
                const double Z = nuc[4 * c + 3];
                if (!(fabs(Z) > 1.0e-12)) continue;     // valence.F90:3149
                const double PC[3] = {P[0] - nuc[4 * c], P[1] - nuc[4 * c + 1], P[2] - nuc[4 * c + 2]};
                const double U = p * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]);
                double F[L + 1];
                boys<L>(boys_tab, U, F);
                double R[L + 1][NE];
                const double pv = Z * pk;
#pragma unroll
                for (int m = 0; m <= L; ++m) R[m][0] = pv * F[m];
                sfor<1, NE>([&](auto EI) {
                    constexpr int e = EI, d = c_dir(e), e1 = c_dec(e, d), n1 = c_l(e1, d), Le = c_L(e);
                    sfor<0, L + 1 - Le>([&](auto M) {
                        constexpr int m = M;
                        double v = PA[d] * R[m][e1] - PC[d] * R[m + 1][e1];
                        if constexpr (n1 > 0) {
                            constexpr int e2 = c_dec(e1, d);
                            v += n1 * h2p * (R[m][e2] - R[m + 1][e2]);
                        }
                        R[m][e] = v;
                    });
                });
#pragma unroll
                for (int e = 0; e < NE; ++e) acc[e] += R[0][e];
            }
        }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        double v = acc[e];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        out[e] = v;
    }
}

// One warp per AO shell pair (i >= j).  Primitive pairs whose Gaussian-product prefactor is below
// kcut (1e-30: far below anything that reaches 1e-10 Eh) are skipped.
__global__ void __launch_bounds__(128) k_ao_1e(const DevShell* __restrict__ sh, int nshell, const double* __restrict__ exps,
                        const double* __restrict__ coefs, const double* __restrict__ nuc /* x,y,z,Z per atom */, int natom,
                        const double* __restrict__ boys_tab, int nao, double kcut, double* __restrict__ S, double* __restrict__ H,
                        int first = 0, int stride = 1 /* shell pairs first, first + stride, ... (one process per GPU) */)
{
    const int lane = threadIdx.x & 31;
    const long long idx = first + (long long)stride * (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const long long npair = (long long)nshell * (nshell + 1) / 2;
    if (idx >= npair) return;
    int i = (int)floor(sqrt(2.0 * (double)idx + 0.25) - 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= idx) ++i;
    while ((long long)i * (i + 1) / 2 > idx) --i;
    int j = (int)(idx - (long long)i * (i + 1) / 2);
    DevShell A = sh[i], B = sh[j];
    if (A.l < B.l) { DevShell t = A; A = B; B = t; }   // A carries la >= lb
    const int la = A.l, lb = B.l, L = la + lb, na = ncart(la), nb = ncart(lb);
    const double AB[3] = {A.x - B.x, A.y - B.y, A.z - B.z};
    const double AB2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    // nuclear attraction auxiliaries, all lanes
    double Raux[ncum(EMAX)];
    switch (L) {
        case 0: nuc_aux<0>(A, B, AB2, exps, coefs, nuc, natom, boys_tab, lane, kcut, Raux); break;
        case 1: nuc_aux<1>(A, B, AB2, exps, coefs, nuc, natom, boys_tab, lane, kcut, Raux); break;
        case 2: nuc_aux<2>(A, B, AB2, exps, coefs, nuc, natom, boys_tab, lane, kcut, Raux); break;
        case 3: nuc_aux<3>(A, B, AB2, exps, coefs, nuc, natom, boys_tab, lane, kcut, Raux); break;
        default: nuc_aux<4>(A, B, AB2, exps, coefs, nuc, natom, boys_tab, lane, kcut, Raux); break;
    }
    if (lane != 0) return;
    double Sb[36], Hb[36];
    for (int n = 0; n < na * nb; ++n) { Sb[n] = 0.0; Hb[n] = 0.0; }
    for (int ia = 0; ia < A.nprim; ++ia)
        for (int ib = 0; ib < B.nprim; ++ib) {
            double a = exps[A.prim_off + ia], b = exps[B.prim_off + ib], p = a + b, ip = 1.0 / p, h2p = 0.5 * ip;
            double K = coefs[A.prim_off + ia] * coefs[B.prim_off + ib] * exp(-a * b * ip * AB2);
            if (!(fabs(K) > kcut)) continue;
            double P[3] = {(a * A.x + b * B.x) * ip, (a * A.y + b * B.y) * ip, (a * A.z + b * B.z) * ip};
            double PA[3] = {P[0] - A.x, P[1] - A.y, P[2] - A.z};
            double PB[3] = {P[0] - B.x, P[1] - B.y, P[2] - B.z};
            // 1D overlap tables s[d][i][j], i <= la, j <= lb + 2
            double s[3][3][5];
            for (int d = 0; d < 3; ++d) {
                s[d][0][0] = 1.0;
                for (int ii = 1; ii <= la; ++ii) s[d][ii][0] = PA[d] * s[d][ii - 1][0] + (ii > 1 ? (ii - 1) * h2p * s[d][ii - 2][0] : 0.0);
                for (int jj = 1; jj <= lb + 2; ++jj)
                    for (int ii = 0; ii <= la; ++ii) {
                        double v = PB[d] * s[d][ii][jj - 1];
                        if (ii > 0) v += ii * h2p * s[d][ii - 1][jj - 1];
                        if (jj > 1) v += (jj - 1) * h2p * s[d][ii][jj - 2];
                        s[d][ii][jj] = v;
                    }
            }
            double pref = K * pow(PI * ip, 1.5);
            for (int ca = 0; ca < na; ++ca) {
                int c1 = coff(la) + ca, al[3] = {c_lx(c1), c_ly(c1), c_lz(c1)};
                for (int cb = 0; cb < nb; ++cb) {
                    int c2 = coff(lb) + cb, bl[3] = {c_lx(c2), c_ly(c2), c_lz(c2)};
                    double sx[3], tx[3];
                    for (int d = 0; d < 3; ++d) {
                        int ii = al[d], jj = bl[d];
                        sx[d] = s[d][ii][jj];
                        tx[d] = -2.0 * b * b * s[d][ii][jj + 2] + b * (2.0 * jj + 1.0) * s[d][ii][jj];
                        if (jj >= 2) tx[d] -= 0.5 * jj * (jj - 1) * s[d][ii][jj - 2];
                    }
                    Sb[ca * nb + cb] += pref * sx[0] * sx[1] * sx[2];
                    Hb[ca * nb + cb] += pref * (tx[0] * sx[1] * sx[2] + sx[0] * tx[1] * sx[2] + sx[0] * sx[1] * tx[2]);
                }
            }
        }
    // horizontal shift of the contracted [e|V|0] to (a|V|b) by binomials in A - B
    for (int ca = 0; ca < na; ++ca) {
        int c1 = coff(la) + ca, ax = c_lx(c1), ay = c_ly(c1), az = c_lz(c1);
        for (int cb = 0; cb < nb; ++cb) {
            int c2 = coff(lb) + cb, bx = c_lx(c2), by = c_ly(c2), bz = c_lz(c2);
            double v = 0.0;
            for (int kx = 0; kx <= bx; ++kx)
                for (int ky = 0; ky <= by; ++ky)
                    for (int kz = 0; kz <= bz; ++kz)
                        v += dev_binom(bx, kx) * dev_binom(by, ky) * dev_binom(bz, kz) * pow(AB[0], (double)(bx - kx)) *
                             pow(AB[1], (double)(by - ky)) * pow(AB[2], (double)(bz - kz)) * Raux[cidx(ax + kx, ay + ky, az + kz)];
            Hb[ca * nb + cb] += v;
        }
    }
    for (int ca = 0; ca < na; ++ca)
        for (int cb = 0; cb < nb; ++cb) {
            size_t r = (size_t)A.ao_off + ca, c = (size_t)B.ao_off + cb;
            S[r * nao + c] = Sb[ca * nb + cb]; S[c * nao + r] = Sb[ca * nb + cb];
            H[r * nao + c] = Hb[ca * nb + cb]; H[c * nao + r] = Hb[ca * nb + cb];
        }
}

// orbital-level one-electron integrals: out_s[k], out_h[k] for the pair list (x_k, y_k) of orbital ids.
// Orbitals are CSR lists of (AO, weight * angn).
__global__ void k_orb_1e(const int* __restrict__ optr, const int* __restrict__ oao, const double* __restrict__ oc,
                         const int2* __restrict__ pairs, int npairs, int nao, const double* __restrict__ S,
                         const double* __restrict__ H, double* __restrict__ out_s, double* __restrict__ out_h)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    int x = pairs[k].x, y = pairs[k].y;
    double ss = 0.0, hh = 0.0;
    for (int i = optr[x]; i < optr[x + 1]; ++i) {
        const double ci = oc[i];
        const size_t row = (size_t)oao[i] * nao;
        double ts = 0.0, th = 0.0;
        for (int j = optr[y]; j < optr[y + 1]; ++j) {
            ts += oc[j] * S[row + oao[j]];
            if (H) th += oc[j] * H[row + oao[j]];
        }
        ss += ci * ts;
        hh += ci * th;
    }
    out_s[k] = ss;
    if (out_h) out_h[k] = hh;
}

// M[r][c] = Se[entry_of_slot(bra_list[r])][entry_of_slot(ket_list[c])]   (row-major n x n)
__global__ void k_gather_block(const double* __restrict__ Se, int nso, const int* __restrict__ bra_entry,
                               const int* __restrict__ ket_entry, int n, double* __restrict__ M)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    int r = idx / n, c = idx % n;
    M[idx] = Se[(size_t)bra_entry[r] * nso + ket_entry[c]];
}

// Multi-CTA in-place Gauss-Jordan inverse for large blocks (cooperative launch, grid-wide barriers).
// Implicit partial pivoting: column k takes its pivot from the not-yet-used row with the largest
// |A[r][k]| (rows are never moved; prow/colf hold the scaled pivot row and the eliminated column so
// that no CTA reads data another CTA is overwriting).  With p(k) the pivot row of column k, the
// storage ends up holding S[p(i)][c] = (A^-1)[i][p(c)]; Ainv receives the un-permuted inverse.
// ws: prow[n] colf[n]; iws: prow_of_col[n], used[n], pivot_row (1).  out[0] = det, out[1] = min|piv|/max|piv|.
__global__ void __launch_bounds__(1024) k_gj_inverse_grid(double* __restrict__ A, int n, double* __restrict__ Ainv,
                                                          double* __restrict__ ws, int* __restrict__ iws, double* __restrict__ out)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    double* prow = ws;
    double* colf = ws + n;
    int* pcol = iws;            // pcol[k] = pivot row of column k
    int* used = iws + n;
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double s_det, s_min, s_max;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (blockIdx.x == 0) {
        for (int r = tid; r < n; r += nt) used[r] = 0;
        if (tid == 0) { s_det = 1.0; s_min = 1e300; s_max = 0.0; }
    }
    __threadfence();
    grid.sync();
    for (int k = 0; k < n; ++k) {
        if (blockIdx.x == 0) {
            double best = -1.0; int bi = n;
            for (int r = tid; r < n; r += nt) {
                if (used[r]) continue;
                double v = fabs(A[(size_t)r * n + k]);
                if (v > best || (v == best && r < bi)) { best = v; bi = r; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_down_sync(0xffffffffu, best, o); int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
            __syncthreads();
            if (tid < 32) {
                best = tid < (nt >> 5) ? s_val[tid] : -1.0; bi = tid < (nt >> 5) ? s_idx[tid] : n;
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_down_sync(0xffffffffu, best, o); int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (tid == 0) { s_idx[0] = bi; }
            }
            __syncthreads();
            const int pr = s_idx[0];
            const double pv = A[(size_t)pr * n + k];
            const double ipv = pv != 0.0 ? 1.0 / pv : 0.0;
            for (int c = tid; c < n; c += nt) {
                prow[c] = (c == k) ? ipv : A[(size_t)pr * n + c] * ipv;
                colf[c] = (c == pr) ? 0.0 : A[(size_t)c * n + k];
            }
            if (tid == 0) {
                pcol[k] = pr; used[pr] = 1; iws[2 * n] = pr;
                s_det *= pv;
                s_min = fmin(s_min, fabs(pv)); s_max = fmax(s_max, fabs(pv));
            }
            __syncthreads();
        }
        __threadfence();
        grid.sync();
        const int pr = iws[2 * n];
        for (int r = blockIdx.x; r < n; r += gridDim.x) {
            double* row = A + (size_t)r * n;
            if (r == pr) {
                for (int c = tid; c < n; c += nt) row[c] = prow[c];
            } else {
                const double f = colf[r];
                if (f == 0.0) continue;
                for (int c = tid; c < n; c += nt) row[c] = (c == k) ? -f * prow[k] : row[c] - f * prow[c];
            }
        }
        __threadfence();
        grid.sync();
    }
    // un-permute: Ainv[i][p(c)] = S[p(i)][c]
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const double* src = A + (size_t)pcol[i] * n;
        for (int c = tid; c < n; c += nt) Ainv[(size_t)i * n + pcol[c]] = src[c];
    }
    if (blockIdx.x == 0 && tid == 0) {
        // sign of the permutation k -> p(k) by cycle counting (used[] is reused as the visited flag)
        int sign = 1;
        for (int r = 0; r < n; ++r) used[r] = 0;
        for (int r = 0; r < n; ++r) {
            if (used[r]) continue;
            int len = 0, x = r;
            while (!used[x]) { used[x] = 1; x = pcol[x]; ++len; }
            if ((len & 1) == 0) sign = -sign;
        }
        out[0] = sign * s_det;
        out[1] = s_max > 0.0 ? s_min / s_max : 0.0;
    }
}

// One step of iterative refinement of the Gauss-Jordan inverse: R = I - M X with every dot product accumulated
// in double-double (error-free products through FMA, two-sum accumulation), then X' = X + X R.  The inverse
// enters E = numerator / wfnorm with weights of order |E| ~ 1e4 Eh, so its n * eps ~ 1e-13 relative error would
// show at the 1e-10 Eh level for clusters of a few hundred orbitals; after this step X is accurate to a few ulp.
constexpr int RF_T = 16;
__global__ void __launch_bounds__(RF_T * RF_T) k_inv_residual_dd(const double* __restrict__ M, const double* __restrict__ X, int n,
                                                                 double* __restrict__ R)
{
    __shared__ double sm[RF_T][RF_T + 1], sx[RF_T][RF_T + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.y * RF_T + ty, j = blockIdx.x * RF_T + tx;
    double hi = 0.0, lo = 0.0;
    for (int k0 = 0; k0 < n; k0 += RF_T) {
        sm[ty][tx] = (i < n && k0 + tx < n) ? M[(size_t)i * n + k0 + tx] : 0.0;
        sx[ty][tx] = (k0 + ty < n && j < n) ? X[(size_t)(k0 + ty) * n + j] : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RF_T; ++k) {
            const double a = sm[ty][k], b = sx[k][tx];
            const double p = __dmul_rn(a, b), e = __fma_rn(a, b, -p);
            const double t = __dadd_rn(hi, p), bp = __dsub_rn(t, hi);
            const double err = __dadd_rn(__dsub_rn(hi, __dsub_rn(t, bp)), __dsub_rn(p, bp));
            hi = t;
            lo = __dadd_rn(lo, __dadd_rn(err, e));
        }
        __syncthreads();
    }
    if (i < n && j < n) R[(size_t)i * n + j] = __dsub_rn(__dsub_rn(i == j ? 1.0 : 0.0, hi), lo);
}
__global__ void __launch_bounds__(RF_T * RF_T) k_inv_update(const double* __restrict__ X, const double* __restrict__ R, int n, double* __restrict__ Xn)
{
    __shared__ double sa[RF_T][RF_T + 1], sb[RF_T][RF_T + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.y * RF_T + ty, j = blockIdx.x * RF_T + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < n; k0 += RF_T) {
        sa[ty][tx] = (i < n && k0 + tx < n) ? X[(size_t)i * n + k0 + tx] : 0.0;
        sb[ty][tx] = (k0 + ty < n && j < n) ? R[(size_t)(k0 + ty) * n + j] : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RF_T; ++k) acc += sa[ty][k] * sb[k][tx];
        __syncthreads();
    }
    if (i < n && j < n) Xn[(size_t)i * n + j] = X[(size_t)i * n + j] + acc;
}

// entry-level density of one spin block: P[s][t] = Minv[pos_ket(t)][pos_bra(s)], 0 when an entry has no slot of this spin
__global__ void k_entry_density(const double* __restrict__ Minv, int n, const int* __restrict__ pos_bra,
                                const int* __restrict__ pos_ket, int nso, double* __restrict__ P)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nso * nso) return;
    int s = idx / nso, t = idx % nso;
    int r = pos_bra[s], c = pos_ket[t];
    P[idx] = (r >= 0 && c >= 0) ? Minv[(size_t)c * n + r] : 0.0;
}

// deterministic two-level sum of n doubles (fixed tree; same result on every run and for every grid)
// double-double accumulation (error-free product through FMA, two-sum): the one-electron numerator of a few hundred
// molecules is a sum of ~1e6 terms of magnitude 1e2 that adds up to ~1e5 Hartree -- a plain double sum alone costs 1e-10 Eh
struct dd { double hi, lo; };
__device__ __forceinline__ void dd_add(dd& s, double p, double e)   // s += p + e
{
    const double t = __dadd_rn(s.hi, p), bp = __dsub_rn(t, s.hi);
    const double err = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, bp)), __dsub_rn(p, bp));
    s.hi = t;
    s.lo = __dadd_rn(s.lo, __dadd_rn(err, e));
}
// out_lo == nullptr: out[0] = the sum rounded to double.  Otherwise the sum leaves as an unevaluated pair: out[0] = its
// multiple of `grid` (a power of two shared by every rank, chosen so that the per-rank values add up EXACTLY in a double
// all-reduce), out_lo[0] = the rest -- the energy is assembled from the pair in extended precision on the host.
__global__ void k_sum(const double* __restrict__ x, long long n, double* __restrict__ out, double* __restrict__ out_lo = nullptr, double grid = 0.0)
{
    __shared__ dd sh[1024];
    dd acc = {0.0, 0.0};            // double-double per thread over a fixed stride pattern: deterministic and exact to ~1e-30
    for (long long i = threadIdx.x; i < n; i += blockDim.x) dd_add(acc, x[i], 0.0);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) dd_add(sh[threadIdx.x], sh[threadIdx.x + o].hi, sh[threadIdx.x + o].lo);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (!out_lo) out[0] = sh[0].hi + sh[0].lo;
        else if (grid > 0.0) {
            const double top = rint(sh[0].hi / grid) * grid;     // exact: a power-of-two grid
            out[0] = top; out_lo[0] = (sh[0].hi - top) + sh[0].lo;
        } else { out[0] = sh[0].hi; out_lo[0] = sh[0].lo; }
    }
}

// E1 = sum_st h[s][t] (Pa+Pb)[s][t],  N1 = sum_st S[s][t] (Pa+Pb)[s][t]   (valence.F90:1072-1106)
__global__ void k_one_electron_energy(const double* __restrict__ Se, const double* __restrict__ He,
                                      const double* __restrict__ Pa, const double* __restrict__ Pb, int n2,
                                      double* __restrict__ out /* [4]: E1, N1, then their low-order parts */)
{
    __shared__ dd sh[2][1024];
    dd e = {0.0, 0.0}, w = {0.0, 0.0};
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const double p = Pa[i] + Pb[i];
        const double pe = __dmul_rn(He[i], p), pw = __dmul_rn(Se[i], p);
        dd_add(e, pe, __fma_rn(He[i], p, -pe));
        dd_add(w, pw, __fma_rn(Se[i], p, -pw));
    }
    sh[0][threadIdx.x] = e; sh[1][threadIdx.x] = w;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            dd_add(sh[0][threadIdx.x], sh[0][threadIdx.x + o].hi, sh[0][threadIdx.x + o].lo);
            dd_add(sh[1][threadIdx.x], sh[1][threadIdx.x + o].hi, sh[1][threadIdx.x + o].lo);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double e = sh[0][0].hi + sh[0][0].lo, w = sh[1][0].hi + sh[1][0].lo;
        out[0] = e; out[1] = w;
        out[2] = (sh[0][0].hi - e) + sh[0][0].lo; out[3] = (sh[1][0].hi - w) + sh[1][0].lo;
    }
}

}  // namespace vb
