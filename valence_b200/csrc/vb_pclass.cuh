// Class-split form of the fused hot path (s/p basis sets): ONE launch per integral class (TB|TK), TB/TK in
// {ss, ps, pp}, all of them accumulating the orbital-level integrals G[q][p] of every tile in an HBM/L2-resident
// buffer with FP64 reductions, followed by one contraction launch (k_contract_items) that screens and contracts
// the tiles with the cofactor densities exactly as the reference does (contract_tile, vb_ptile.cuh).
//
// Why (measured on B200, profiles/r2_*): the all-in-one kernel k_ptile<PART_ALL> is 250 KB of unrolled code in
// which 12 warps per SM sit in nine different class bodies (instruction-fetch stalls as frequent as issues), and
// its register allocation is dictated by the (pp|pp) body (168 registers + 5.5 KB of spills per thread) although
// 90 % of the primitive quartets belong to the four light classes.  Here every class is its own small kernel with
// its own register budget and occupancy, only the tables of ITS pair types are staged (TMA bulk), and the inner
// loop can afford instruction-level parallelism.
//
// The arithmetic is that of vb_ptile.cuh (same formulation: lane = ket primitive x bra shell pair, in-lane
// contraction over the bra primitives, two DMMA transforms); pruning is per lane (own ket weight).
// (replaces int2e, valence.F90:3184-3438, and the 2e loop of vsvb_energy, :1153-1433)
#pragma once
#include "vb_ptile.cuh"

namespace vb {

constexpr int PQ_FAR = 18;   // pq_counters[PQ_FAR + tb*3 + tk]: quartets of class (tb|tk) evaluated with the asymptotic Boys values
constexpr int PQ_DMMA = 31;  // pq_counters[PQ_DMMA]: FP64 tensor-core instructions (m8n8k4, 512 flops each) of the two density transforms

struct ClassCfg {           // shared-memory capacities of one class launch (host: max over the pair groups)
    int d_cap;              // doubles for the staged bra densities of pair type TB (+2 slack for the alignment shift)
    int sp_cap, pp_cap;     // bra shell pairs / primitive pairs of type TB
};

// kets per lane (KB): every lane carries KB ket primitives through the bra loop -- KB independent dependency chains
// per lane (the loop is latency-bound: 8 resident warps per scheduler at most), one bra-primitive load, one loop
// head and one density-fragment load for KB quartets, and one reduction of G per KB octets.
#ifndef VB_KB_SS
#define VB_KB_SS 2
#endif
#ifndef VB_KB_PSSS
#define VB_KB_PSSS 2
#endif
#ifndef VB_KB_SSPS
#define VB_KB_SSPS 1     // three ket components: two kets per lane do not fit 128 registers (measured slower)
#endif
#ifndef VB_KB_PSPS
#define VB_KB_PSPS 1
#endif
#ifndef VB_KB_PP
#define VB_KB_PP 1
#endif
__host__ __device__ constexpr int pc_kb(int tb, int tk)
{
    return (tb == 2 || tk == 2) ? VB_KB_PP : (tb + tk == 0 ? VB_KB_SS : (tb + tk == 1 ? (tb == 1 ? VB_KB_PSSS : VB_KB_SSPS) : VB_KB_PSPS));
}

// 1/sqrt(x) for normal positive x without the special-case branch of the library routine (that branch and the
// far/near branches of the Boys function keep the scheduler from interleaving the KB dependency chains):
// hardware seed (2^-22) + one third-order correction y0 (1 + e/2 + 3 e^2/8), e = 1 - x y0^2.
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(x, -(y0 * y0), 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}
// quartet_geom (vb_eri.cuh) with the branch-free reciprocal square root
__device__ __forceinline__ void quartet_geom_fast(const PrimPair& a, const PrimPair& b, QuartetGeom& g, double& T, double& pref)
{
    const double p = a.p, q = b.p;
    const double r = rsqrt_fast(p + q), ipq = r * r;
    const double wq = q * ipq, wp = p * ipq;             // rho/p, rho/q
    const double dx = a.Px - b.Px, dy = a.Py - b.Py, dz = a.Pz - b.Pz;
    T = p * wq * (dx * dx + dy * dy + dz * dz);
    pref = a.Kp * b.Kp * r;
    g.PA[0] = a.PAx; g.PA[1] = a.PAy; g.PA[2] = a.PAz;
    g.QC[0] = b.PAx; g.QC[1] = b.PAy; g.QC[2] = b.PAz;
    g.WP[0] = -wq * dx; g.WP[1] = -wq * dy; g.WP[2] = -wq * dz;
    g.WQ[0] = wp * dx; g.WQ[1] = wp * dy; g.WQ[2] = wp * dz;
    g.h2p = 0.5 * a.ip; g.h2q = 0.5 * b.ip; g.h2pq = 0.5 * ipq; g.rp = wq; g.rq = wp;
}
// Boys values in the asymptotic regime T >= BOYS_S_TMAX (exp(-T) < 5e-18 dropped, as boys_s does), branch-free
template <int M>
__device__ __forceinline__ void boys_far(double T, double* F)
{
    const double r = rsqrt_fast(T), rt = r * r;
    F[0] = 0.88622692545275801365 * r;
#pragma unroll
    for (int m = 0; m < M; ++m) F[m + 1] = (m + 0.5) * F[m] * rt;
}

// the quartet values of KB kets go straight into the tensor cores: X[kb]_f[k][p] += sum_b A[k][b] Dp[e_b + e][p]
template <int TB, int TK, int KB>
__device__ __forceinline__ void feed_dmma_kb(const double (&acc)[KB][pt_ne(TB) * pt_ne(TK)], int eoff, const double* __restrict__ Dp_s,
                                             int npP, int g, double (&X)[KB][pt_ne(TK)][4][2])
{
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    const double* drow = Dp_s + eoff * npP + g;        // B[t][n = 8j + g] = Dp[e_sp + e][8j + g]
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double bf = drow[e * npP + 8 * j];
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int f = 0; f < NF; ++f) dmma_884(X[kb][f][j][0], X[kb][f][j][1], acc[kb][e * NF + f], bf);
        }
}

// One warp task: noct groups of KB consecutive ket octets of pair type TK (starting with group oct) against every bra
// primitive of type TB of P; the groups accumulate into one G fragment that is reduced into the tile's G once.
template <int TB, int TK, int KB, int KO>
__device__ __forceinline__ void pclass_task(const TileArgs& A, int nsp, int npP, int pp_base, int e_beg, const PGDesc& Q, int oct0,
                                            const PrimPair* __restrict__ bpps, const SPRec* __restrict__ spss,
                                            const double* __restrict__ Dp_s /* first row = e_beg */, const double* __restrict__ Dq_g,
                                            const double* __restrict__ boys_tab, double* __restrict__ scratch,
                                            double* __restrict__ Gglob, int lane, unsigned long long* __restrict__ s_pq)
{
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    const int g = lane >> 2, t = lane & 3;
    const int nk = Q.pp_beg[TK + 1] - Q.pp_beg[TK];
    const PrimPair* __restrict__ kl = A.pps_flat + Q.pp_beg[TK];
    double C[4][4][2];
    if constexpr (KO > 1) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) { C[m][j][0] = 0.0; C[m][j][1] = 0.0; }
    }
    unsigned long long nq_task = 0ull, nfar_task = 0ull, ndmma_task = 0ull;
    auto octet_group = [&](const int oct) -> bool {     // false: nothing left further down the (sorted) ket list
    const int k0 = 8 * KB * oct;
    if (k0 >= nk) return false;
    PrimPair b[KB];
    double wl[KB];                                     // this lane's own ket magnitudes (non-increasing in kb: the list is sorted)
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const bool kact = k0 + 8 * kb + g < nk;
        b[kb] = kl[kact ? k0 + 8 * kb + g : k0];
        if (!kact) { b[kb].Kp = 0.0; b[kb].w = 0.0; }
        wl[kb] = b[kb].w;
    }
    const double wk = kl[k0].w;                        // the task's largest magnitude
    double X[KB][NF][4][2];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) { X[kb][f][j][0] = 0.0; X[kb][f][j][1] = 0.0; }
    unsigned nq = 0, nfar = 0;                          // primitive quartets evaluated by this lane; those in the asymptotic regime
    for (int q0 = 0; q0 < nsp; q0 += 4) {
        // shell pairs are sorted by contraction length, not by weight: test the quad's own bound
        const bool sact = q0 + t < nsp;
        const SPRec sp = spss[q0 + (sact ? t : 0)];
        if (!__any_sync(0xffffffffu, sact && sp.wmax * wk >= A.tau)) {
            continue;
        }
        const int cnt = (sact && sp.wmax * wl[0] >= A.tau) ? sp.pp_cnt : 0;
        const PrimPair* __restrict__ bl = bpps + (sp.pp_beg - pp_base);
        double acc[KB][NE * NF];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int i = 0; i < NE * NF; ++i) acc[kb][i] = 0.0;
        int ip = 0;
        for (;; ++ip) {
            // primitives of a shell pair are sorted by magnitude: a lane that stops stays stopped
            const PrimPair a = bl[ip < cnt ? ip : 0];
            const bool act = ip < cnt && a.w * wl[0] >= A.tau;
            if (!__any_sync(0xffffffffu, act)) break;
            // All lanes and all KB kets in the asymptotic regime (true for most trips of a large cluster): straight-line
            // code without a table look-up, so the KB chains interleave; otherwise the general routine per ket.
            constexpr int M = pt_E(TB) + pt_E(TK);
            bool actk[KB];
            PrimPair ak[KB];
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                actk[kb] = kb == 0 ? act : (act && a.w * wl[kb] >= A.tau);
                ak[kb] = a;
                if (!actk[kb]) ak[kb].Kp = 0.0;
                nq += actk[kb] ? 1u : 0u;
            }
            if constexpr (M == 0) {
                double sv[KB];
                bool far = true;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    const double u = a.p + b[kb].p;
                    const double dx = a.Px - b[kb].Px, dy = a.Py - b[kb].Py, dz = a.Pz - b[kb].Pz;
                    const double sq = a.p * b[kb].p * (dx * dx + dy * dy + dz * dz);
                    const bool fk = sq >= BOYS_S_TMAX * u;
                    far = far && (!actk[kb] || fk);
                    nfar += (actk[kb] && fk) ? 1u : 0u;
                    sv[kb] = actk[kb] ? sq : 1.0;
                }
                if (__all_sync(0xffffffffu, far)) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) acc[kb][0] += ak[kb].Kp * b[kb].Kp * 0.88622692545275801365 * rsqrt_fast(sv[kb]);
                } else {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) quartet_values<TB, TK>(boys_tab, ak[kb], b[kb], acc[kb]);
                }
            } else {
                constexpr int LA = pt_la(TB), EA = pt_E(TB), LC = pt_la(TK), EC = pt_E(TK);
                QuartetGeom geo[KB];
                double T[KB], pref[KB];
                bool far = true;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    quartet_geom_fast(ak[kb], b[kb], geo[kb], T[kb], pref[kb]);
                    const bool fk = T[kb] >= BOYS_S_TMAX;
                    far = far && (!actk[kb] || fk);
                    nfar += (actk[kb] && fk) ? 1u : 0u;
                }
                if (__all_sync(0xffffffffu, far)) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        double F[M + 1];
                        boys_far<M>(actk[kb] ? T[kb] : 64.0, F);
#pragma unroll
                        for (int m = 0; m <= M; ++m) F[m] *= pref[kb];
                        vrr_unrolled<LA, EA, LC, EC>(geo[kb], F, acc[kb]);
                    }
                } else {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        double F[M + 1];
                        boys_s<M>(boys_tab, T[kb], F);
#pragma unroll
                        for (int m = 0; m <= M; ++m) F[m] *= pref[kb];
                        vrr_unrolled<LA, EA, LC, EC>(geo[kb], F, acc[kb]);
                    }
                }
            }
        }
        if (ip > 0) { feed_dmma_kb<TB, TK, KB>(acc, sp.eoff - e_beg, Dp_s, npP, g, X); ndmma_task += NE * NF * 4 * KB; }
    }
    for (int o = 16; o > 0; o >>= 1) { nq += __shfl_xor_sync(0xffffffffu, nq, o); nfar += __shfl_xor_sync(0xffffffffu, nfar, o); }
    if (!nq) return false;                             // the ket list is sorted by magnitude: nothing further down either
    nq_task += nq; nfar_task += nfar; ndmma_task += NF * KB * 32;
    if constexpr (KO == 1) {                           // single-octet tasks: C only lives from here
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) { C[m][j][0] = 0.0; C[m][j][1] = 0.0; }
    }
    // G[q][p] += sum_{k,f} Dq[e_k + f][q] X_f[k][p]:  A'[q][k] from the ket densities, B'[k][p] = X_f re-laid out
    // through the warp's scratch (C fragment -> B fragment); all octets of the task accumulate into one C
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const int eo_own = b[kb].eoff;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                scratch[g * PT_SLD + 8 * j + 2 * t] = X[kb][f][j][0];
                scratch[g * PT_SLD + 8 * j + 2 * t + 1] = X[kb][f][j][1];
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
    }
    return true;
    };   // octet groups of the task
    if constexpr (KO == 1) octet_group(oct0);
    else {
#pragma unroll 1
        for (int oct = oct0; oct < oct0 + KO; ++oct)
            if (!octet_group(oct)) break;
    }
    if (!nq_task) return;
    if (lane == 0) { atomicAdd(&s_pq[0], nq_task); atomicAdd(&s_pq[1], nfar_task); atomicAdd(&s_pq[2], ndmma_task); }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int q = 8 * m + g;
        if (q >= Q.np) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int p = 8 * j + 2 * t + i;
                if (p < npP) atomicAdd(&Gglob[q * npP + p], C[m][j][i]);
            }
    }
}

#ifndef VB_MB_SS
#define VB_MB_SS 2
#endif
#ifndef VB_MB_SP
#define VB_MB_SP 2
#endif
#ifndef VB_MB_PSPS
#define VB_MB_PSPS 2
#endif
// octet groups per task (KO): classes with a pp bra pair have ONE quad of bra shell pairs, i.e. a few trips per octet --
// their tasks take many octets so that the reduction of G into L2 (20 atomics per lane) is paid once per task
#ifndef VB_KO_BPP
#define VB_KO_BPP 64
#endif
__host__ __device__ constexpr int pc_ko(int tb, int tk) { return tb == 2 ? VB_KO_BPP : 1; }
__host__ __device__ constexpr int pc_threads(int tb, int tk) { return 256; }
__host__ __device__ constexpr int pc_minblocks(int tb, int tk)
{
    return (tb == 2 || tk == 2) ? 1 : (tb + tk == 0 ? VB_MB_SS : (tb + tk == 1 ? VB_MB_SP : VB_MB_PSPS));
}

// Persistent kernel of one class.  Work item = bra pair group P with up to PT_MAXQ ket pair groups (consecutive tiles);
// G of tile j of the item lives at gbuf[(item.z + j - gslot_base) * g_cap].
template <int TB, int TK>
__global__ void __launch_bounds__(pc_threads(TB, TK), pc_minblocks(TB, TK)) k_pclass(const TileArgs A, const ClassCfg C)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int THREADS = pc_threads(TB, TK), nw = THREADS / 32, NE = pt_ne(TB), KB = pc_kb(TB, TK), KO = pc_ko(TB, TK);
    double* Dp_s = smem;                                                   // bra densities of type TB
    double* scr = Dp_s + C.d_cap;                                          // per-warp X scratch
    double* boys_sm = scr + nw * PT_SCRATCH;                               // compact Boys table
    SPRec* sps_s = reinterpret_cast<SPRec*>(boys_sm + BOYS_S_SIZE);        // bra shell pairs of type TB
    PrimPair* bpp_s = reinterpret_cast<PrimPair*>(sps_s + C.sp_cap);       // bra primitive pairs of type TB
    for (int i = threadIdx.x; i < BOYS_S_SIZE; i += THREADS) boys_sm[i] = A.boys_small[i];
    __shared__ unsigned long long s_bar;
    unsigned phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int s_item, s_unit;
    __shared__ int s_cum[PT_MAXQ + 1];
    __shared__ unsigned long long s_pq[3];        // primitive quartets evaluated / of those in the asymptotic regime / DMMA instructions
    __shared__ PGDesc s_P, s_Q[PT_MAXQ];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 3) s_pq[tid] = 0ull;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(A.counter, 1u);   // work stealing inside this rank's shard
        __syncthreads();
        const long long it = (long long)s_item;
        if (it >= A.nitems) break;
        const int4 item = A.items[it];                          // first tile, # tiles, slot of the first tile in gbuf
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
        const int e_beg = P.e_beg[TB];                          // first e-row of pair type TB inside P's density block
        const long long o = P.d_off + (long long)e_beg * P.np;  // TMA wants 16-byte alignment: copy from the even element below
        const int shift = (int)(o & 1);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bd = (unsigned)(((shift + (P.e_beg[TB + 1] - e_beg) * P.np + 1) & ~1) * sizeof(double));
            const unsigned bb = (unsigned)(nbpp * sizeof(PrimPair));
            const unsigned bs = (unsigned)(nsp * sizeof(SPRec));
            mbar_expect_tx(&s_bar, bd + bb + bs);
            tma_bulk_g2s(Dp_s, A.dmat + (o - shift), bd, &s_bar);
            tma_bulk_g2s(sps_s, A.sps + P.sp_beg[TB], bs, &s_bar);
            tma_bulk_g2s(bpp_s, A.pps + P.pp_beg[TB], bb, &s_bar);
            int n = 0;
            for (int qi = 0; qi < PT_MAXQ; ++qi) {
                s_cum[qi] = n;
                if (qi < ntl) {
                    const int nk = s_Q[qi].pp_beg[TK + 1] - s_Q[qi].pp_beg[TK];
                    if (nk > 0 && P.kwmax[TB] * s_Q[qi].kwmax[TK] >= A.tau) n += (nk + 8 * KB * KO - 1) / (8 * KB * KO);
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
            const int oct = (u - s_cum[qi]) * KO;
            const PGDesc& Q = s_Q[qi];
            double* Gglob = A.gbuf + ((size_t)item.z + qi - A.gslot_base) * A.g_cap;
            pclass_task<TB, TK, KB, KO>(A, nsp, P.np, P.pp_beg[TB], e_beg, Q, oct, bpp_s, sps_s, Dp_s + shift, A.dmat + Q.d_off, boys_sm, scratch, Gglob, lane, s_pq);
        }
    }
    __syncthreads();
    if (tid == 0 && s_pq[0]) {
        atomicAdd(&A.pq_counters[TB * NPTYPE + TK], s_pq[0]); atomicAdd(&A.pq_counters[PQ_FAR + TB * 3 + TK], s_pq[1]);
        atomicAdd(&A.pq_counters[PQ_DMMA], s_pq[2]);
    }
}

// Contraction of the finished tiles of a chunk of work items with the cofactor densities: one tile per warp,
// reference screening / bookkeeping in contract_tile (valence.F90:1153-1433).
constexpr int CI_THREADS = 256;
__global__ void __launch_bounds__(CI_THREADS, 2) k_contract_items(const TileArgs A, long long ntile_slots)
{
    __shared__ unsigned long long s_cnt[CNT_N];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < CNT_N) s_cnt[tid] = 0ull;
    __syncthreads();
    unsigned long long cnt[CNT_N];
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) cnt[i] = 0ull;
    const long long nwarps = (long long)gridDim.x * (CI_THREADS / 32);
    for (long long w = (long long)blockIdx.x * (CI_THREADS / 32) + (tid >> 5); w < (long long)A.nitems * PT_MAXQ; w += nwarps) {
        const int4 item = A.items[w / PT_MAXQ];
        const int j = (int)(w % PT_MAXQ);
        if (j >= item.y) continue;
        const int2 tq = A.tiles[item.x + j];
        const TileIdx P{A.pgs[tq.x].np, A.pgs[tq.x].pair_beg}, Q{A.pgs[tq.y].np, A.pgs[tq.y].pair_beg};
        const double* __restrict__ src = A.gbuf + ((size_t)item.z + j - A.gslot_base) * A.g_cap;
        const int npP = P.np;
        auto gv = [&](int p, int q) { return __ldcs(&src[q * npP + p]); };
        double epart = A.sym ? contract_tile<32, true, false>(A, P, Q, gv, lane, cnt) : contract_tile<32, false, false>(A, P, Q, gv, lane, cnt);
        for (int o = 16; o > 0; o >>= 1) epart += __shfl_down_sync(0xffffffffu, epart, o);
        if (lane == 0) A.tileE[item.x + j] = epart * A.c0;
    }
    (void)ntile_slots;
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
