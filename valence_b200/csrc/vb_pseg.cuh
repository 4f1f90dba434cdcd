// Far-field form of the four light integral classes (ss|ss) (ps|ss) (ss|ps) (ps|ps): 90 % of the primitive quartets of a
// water cluster, 96 % of them in the asymptotic regime of the Boys function for (H2O)_256.  Same launch structure and the
// same lane mapping as k_pclass (persistent CTAs, work items (P, <= 8 Q), TMA-staged bra tables; lane = ket primitive g of
// an octet x bra shell-pair segment t of a quad, in-lane contraction over the segment's primitives, two DMMA density
// transforms, G reduced into the L2-resident buffer), but the inner loop is the closed far-field form of vb_far.cuh on
// compact 80-byte primitive records: no Boys table, no exponent-dependent geometry, no recurrence, the magnitude cut on the
// integer pipe -- 14 FP64 instructions per (ss|ss) quartet.  Quartets that fail the far-field test are skipped there and
// flagged; a cold fix-up loop evaluates exactly those through the general Obara-Saika path afterwards.
// (replaces int2e, valence.F90:3184-3438, for these classes)
#pragma once
#include "vb_far.cuh"
#include "vb_pclass.cuh"

namespace vb {

// compact far-field records of every primitive pair, index-aligned with pps (bra role) and pps_flat (ket role)
__global__ void k_far_tables(const PrimPair* __restrict__ pps, const PrimPair* __restrict__ pps_flat, long long n, FarPrim* __restrict__ fcp,
                             FarPrim* __restrict__ fcpf)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fcp[i] = far_make_prim(pps[i]);
    fcpf[i] = far_make_prim(pps_flat[i]);
}

struct SegCfg {             // shared-memory capacities of one launch (host: max over the pair groups) and the far-field tables
    int d_cap;              // doubles for the staged bra densities of pair type TB (+2 slack for the alignment shift)
    int sp_cap, pp_cap;     // bra segments / primitive pairs of type TB
    const FarPrim* fcp;     // aligned with TileArgs::pps
    const FarPrim* fcpf;    // aligned with TileArgs::pps_flat
    unsigned lthr;          // far_lthr(tau)
    int skip;               // experiments only (VB_SEG_SKIP): 1 skips the general fix-up, 2 the far-field form, 4 the first transform
};

#ifndef VB_SEG_MB_SS
#define VB_SEG_MB_SS 2
#endif
#ifndef VB_SEG_MB_SP
#define VB_SEG_MB_SP 2
#endif
#ifndef VB_SEG_MB_PSPS
#define VB_SEG_MB_PSPS 2
#endif
__host__ __device__ constexpr int ps_minblocks(int tb, int tk) { return tb + tk == 0 ? VB_SEG_MB_SS : (tb + tk == 1 ? VB_SEG_MB_SP : VB_SEG_MB_PSPS); }
constexpr int PS_THREADS = 256;
constexpr int PQ_SEGFAR = 27;      // pq_counters[PQ_SEGFAR + tb*2 + tk]: quartets of light class (tb|tk) evaluated in the far-field form
constexpr int FAR_PAD = 64;        // records behind the staged primitive table that the far-field loop may touch (>= longest segment)

// cold path: the quartets of one segment step that are NOT in the far field (bit ia of `mask`: bra primitive ia of the lane's
// segment against the lane's ket primitive), through the Boys table and the recurrence
template <int TB, int TK>
__device__ __noinline__ void near_fixup(const PrimPair* __restrict__ bl, const PrimPair* __restrict__ kp, unsigned long long mask, int trips,
                                        const double* __restrict__ boys_tab, double (&acc)[pt_ne(TB) * pt_ne(TK)])
{
    const PrimPair b = *kp;
    for (int ia = 0; ia < trips; ++ia) {
        const bool act = (mask >> ia) & 1ull;
        if (!__any_sync(0xffffffffu, act)) continue;
        PrimPair a = bl[act ? ia : 0];
        if (!act) a.Kp = 0.0;
        quartet_values<TB, TK>(boys_tab, a, b, acc);
    }
}

// One warp task: the octet oct of ket primitives of pair type TK of Q against every bra segment of type TB of P.
template <int TB, int TK>
__device__ __forceinline__ void pfar_task(const TileArgs& A, const SegCfg& C, int nsp, int npP, int pp_base, int e_beg, const PGDesc& Q, int oct,
                                          const FarPrim* __restrict__ bfp, const SPRec* __restrict__ spss, const PrimPair* __restrict__ bpp,
                                          const double* __restrict__ Dp_s /* first row = e_beg */, const double* __restrict__ Dq_g,
                                          const double* __restrict__ boys_tab, double* __restrict__ scratch,
                                          double* __restrict__ Gglob, int lane, unsigned long long* __restrict__ s_pq)
{
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    const int g = lane >> 2, t = lane & 3;
    const int nk = Q.pp_beg[TK + 1] - Q.pp_beg[TK];
    const int k0 = 8 * oct;
    const bool kact = k0 + g < nk;
    const int kidx = Q.pp_beg[TK] + (kact ? k0 + g : k0);
    const FarPrim kf = C.fcpf[kidx];
    const FarKet k = far_ket(kf);
    const double wl = kact ? kf.w : 0.0;                            // this lane's ket magnitude
    const double wk = C.fcpf[Q.pp_beg[TK] + k0].w;                  // the octet's largest (the list is sorted)
    const unsigned lwk = C.lthr - far_hi32(kact ? kf.w : 1e-300);   // a bra primitive passes when hi(w) >= lwk (as a signed difference)
    const int eo_own = A.pps_flat[kidx].eoff;
    double X[NF][4][2];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) { X[f][j][0] = 0.0; X[f][j][1] = 0.0; }
    unsigned nq = 0, nnear = 0;      // quartets evaluated / of those through the general path
    for (int q0 = 0; q0 < nsp; q0 += 4) {
        const bool sact = q0 + t < nsp;
        const SPRec sp = spss[q0 + (sact ? t : 0)];
        if (!__any_sync(0xffffffffu, sact && sp.wmax * wk >= A.tau)) continue;
        const int cnt = (sact && sp.wmax * wl >= A.tau) ? sp.pp_cnt : 0;
        const int trips = __reduce_max_sync(0xffffffffu, cnt);
        if (trips == 0) continue;
        const FarPrim* __restrict__ bl = bfp + (sp.pp_beg - pp_base);
        FarSums<TB, TK> S;
        S.clear();
        unsigned long long nearmask = 0ull;
        unsigned ntake = 0;
        if (!(C.skip & 2)) {
#pragma unroll 2
            for (int ia = 0; ia < trips; ++ia) {
                const FarPrim& a = bl[ia];                          // up to FAR_PAD records past the segment: finite, discarded
                const bool take = ia < cnt && (int)(far_hi32(a.w) - lwk) >= 0;
                const bool far = far_quartet<TB, TK>(a, k, take, S);
                if (take && !far) nearmask |= 1ull << ia;
                ntake += take ? 1u : 0u;
            }
        }
        if (!__any_sync(0xffffffffu, ntake != 0u)) continue;
        nq += ntake;
        double acc[NE * NF];
#pragma unroll
        for (int i = 0; i < NE * NF; ++i) acc[i] = 0.0;
        far_finish<TB, TK>(S, k, acc);
        if (__any_sync(0xffffffffu, nearmask != 0ull)) {
            nnear += __popcll(nearmask);
            if (!(C.skip & 1)) near_fixup<TB, TK>(bpp + (sp.pp_beg - pp_base), A.pps_flat + kidx, nearmask, trips, boys_tab, acc);
        }
        if (!(C.skip & 4)) {
            // X_f[k][p] += sum_b A[k][b] Dp[e_b + e][p]: lane (g, t) supplies A[g][t], B[t][8j + g] comes from the staged densities
            const double* drow = Dp_s + (sp.eoff - e_beg) * npP + g;
#pragma unroll
            for (int e = 0; e < NE; ++e)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double bf = drow[e * npP + 8 * j];
#pragma unroll
                    for (int f = 0; f < NF; ++f) dmma_884(X[f][j][0], X[f][j][1], acc[e * NF + f], bf);
                }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { nq += __shfl_xor_sync(0xffffffffu, nq, o); nnear += __shfl_xor_sync(0xffffffffu, nnear, o); }
    if (!nq) return;
    if (lane == 0) { atomicAdd(&s_pq[0], (unsigned long long)nq); atomicAdd(&s_pq[1], (unsigned long long)(nq - nnear)); }
    // G[q][p] += sum_{k,f} Dq[e_k + f][q] X_f[k][p]:  A'[q][k] from the ket densities, B'[k][p] = X_f re-laid out through the
    // warp's scratch (C fragment -> B fragment)
    double Cf[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) { Cf[m][j][0] = 0.0; Cf[m][j][1] = 0.0; }
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
            const bool kk_act = k0 + kk < nk;
            double bfr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bfr[j] = scratch[kk * PT_SLD + 8 * j + g];
            const double* arow = Dq_g + (eo + f) * Q.np;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int q = 8 * m + g;
                const double af = (kk_act && q < Q.np) ? arow[q] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(Cf[m][j][0], Cf[m][j][1], af, bfr[j]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int q = 8 * m + g;
        if (q >= Q.np) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int p = 8 * j + 2 * t + i;
                if (p < npP) atomicAdd(&Gglob[q * npP + p], Cf[m][j][i]);
            }
    }
}

// Persistent kernel of one light class in far-field form.  Work item = bra pair group P with up to PT_MAXQ ket pair groups;
// G of tile j of the item lives at gbuf[(item.z + j - gslot_base) * g_cap].
template <int TB, int TK>
__global__ void __launch_bounds__(PS_THREADS, ps_minblocks(TB, TK)) k_pseg(const TileArgs A, const SegCfg C)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int nw = PS_THREADS / 32, NE = pt_ne(TB);
    double* Dp_s = smem;                                                   // bra densities of type TB
    double* scr = Dp_s + C.d_cap;                                          // per-warp X scratch
    double* boys_sm = scr + nw * PT_SCRATCH;                               // compact Boys table (general path)
    SPRec* sps_s = reinterpret_cast<SPRec*>(boys_sm + BOYS_S_SIZE);        // bra segments of type TB
    FarPrim* bfp_s = reinterpret_cast<FarPrim*>(sps_s + C.sp_cap);         // their primitives (far-field form, + FAR_PAD)
    PrimPair* bpp_s = reinterpret_cast<PrimPair*>(bfp_s + C.pp_cap + FAR_PAD);   // their primitives (general form)
    for (int i = threadIdx.x; i < BOYS_S_SIZE; i += PS_THREADS) boys_sm[i] = A.boys_small[i];
    // the far-field loops read up to FAR_PAD records past a lane's own segment and discard them by selection: keep the whole
    // table finite (only FarPrim records are ever written over the zeros)
    for (int i = threadIdx.x; i < (C.pp_cap + FAR_PAD) * (int)(sizeof(FarPrim) / sizeof(double)); i += PS_THREADS) reinterpret_cast<double*>(bfp_s)[i] = 0.0;
    __shared__ unsigned long long s_bar;
    unsigned phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int s_item, s_unit;
    __shared__ int s_cum[PT_MAXQ + 1];
    __shared__ unsigned long long s_pq[3];
    __shared__ PGDesc s_P, s_Q[PT_MAXQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 3) s_pq[tid] = 0ull;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(A.counter, 1u);   // work stealing inside this rank's shard
        __syncthreads();
        const long long it = (long long)s_item;
        if (it >= A.nitems) break;
        const int4 item = A.items[it];
        const int tl0 = item.x, ntl = item.y;
        {
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
        const int nsp = P.sp_beg[TB + 1] - P.sp_beg[TB], nbpp = P.pp_beg[TB + 1] - P.pp_beg[TB];
        if (nsp == 0 || nbpp == 0) continue;
        const int e_beg = P.e_beg[TB], nrows = P.e_beg[TB + 1] - e_beg;   // rows of pair type TB inside P's density block
        const long long o = P.d_off + (long long)e_beg * P.np;  // TMA wants 16-byte alignment: copy from the even element below
        const int shift = (int)(o & 1);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bd = (unsigned)(((shift + nrows * P.np + 1) & ~1) * sizeof(double));
            const unsigned bb = (unsigned)(nbpp * sizeof(FarPrim));
            const unsigned bs = (unsigned)(nsp * sizeof(SPRec));
            const unsigned bg = (unsigned)(nbpp * sizeof(PrimPair));
            mbar_expect_tx(&s_bar, bd + bb + bs + bg);
            tma_bulk_g2s(Dp_s, A.dmat + (o - shift), bd, &s_bar);
            tma_bulk_g2s(sps_s, A.sps + P.sp_beg[TB], bs, &s_bar);
            tma_bulk_g2s(bfp_s, C.fcp + P.pp_beg[TB], bb, &s_bar);
            tma_bulk_g2s(bpp_s, A.pps + P.pp_beg[TB], bg, &s_bar);
            int n = 0;
            for (int qi = 0; qi < PT_MAXQ; ++qi) {
                s_cum[qi] = n;
                if (qi < ntl) {
                    const int nkq = s_Q[qi].pp_beg[TK + 1] - s_Q[qi].pp_beg[TK];
                    if (nkq > 0 && P.kwmax[TB] * s_Q[qi].kwmax[TK] >= A.tau) n += (nkq + 7) / 8;
                }
            }
            s_cum[PT_MAXQ] = n;
            s_unit = 0;
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
        __syncthreads();
        const int nunits = s_cum[PT_MAXQ];
        double* scratch = scr + warp * PT_SCRATCH;
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_unit, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= nunits) break;
            int qi = 0;
            while (s_cum[qi + 1] <= u) ++qi;
            const int oct = u - s_cum[qi];
            const PGDesc& Q = s_Q[qi];
            double* Gglob = A.gbuf + ((size_t)item.z + qi - A.gslot_base) * A.g_cap;
            pfar_task<TB, TK>(A, C, nsp, P.np, P.pp_beg[TB], e_beg, Q, oct, bfp_s, sps_s, bpp_s, Dp_s + shift, A.dmat + Q.d_off, boys_sm, scratch, Gglob, lane, s_pq);
        }
    }
    __syncthreads();
    if (tid == 0 && s_pq[0]) {
        atomicAdd(&A.pq_counters[TB * NPTYPE + TK], s_pq[0]); atomicAdd(&A.pq_counters[PQ_FAR + TB * 3 + TK], s_pq[1]);
        atomicAdd(&A.pq_counters[PQ_SEGFAR + TB * 2 + TK], s_pq[1]);
    }
}

}  // namespace vb
