// Shell-pair formulation of the fused kernel, kept for basis sets with d shells (k_tile<true>; the s/p
// kernel is vb_ptile.cuh): for one tile (bra pair group P, ket pair group Q)
//   1. every contracted shell quartet of the AO block (P-support | Q-support) is generated once
//      (Obara-Saika VRR per angular-momentum class; warp = one bra shell pair, lanes = ket
//      primitive-pair chunks of one class),
//   2. half-transformed on the fly with the staged ket pair densities (lanes = ket orbital pairs,
//      integrals broadcast by warp shuffles), then with the bra pair densities,
//   3. the resulting orbital-level integrals (st|uv) are screened exactly as the reference
//      screens them (valence.F90:1189-1190, 1213-1216, 1286-1287) and contracted with the
//      cofactor densities; nothing but one double per tile reaches HBM.
#pragma once
#include "vb_cofactor.h"
#include "vb_kernels.cuh"

namespace vb {

// ---- TMA bulk copy (cp.async.bulk, global -> shared) completed through an mbarrier ---------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ int tri_index(int a, int c) { return a * (a + 1) / 2 + c; }   // a >= c, 0-based

// W(s,t,u,v) = P[s,t] P[u,v] - sum_sigma P^s[s,v] P^s[u,t]   (times c0 outside)
__device__ __forceinline__ double w_term(const double* __restrict__ Pa, const double* __restrict__ Pb, int nso, int s, int t, int u, int v)
{
    if (Pa == Pb) {     // closed shell: one density for both spins (the host passes the same pointer) -- half the gathers, same arithmetic
        const double ast = Pa[s * nso + t], auv = Pa[u * nso + v], asv = Pa[s * nso + v], aut = Pa[u * nso + t];
        const double x = __dmul_rn(asv, aut);
        return __dmul_rn(ast + ast, auv + auv) - x - x;
    }
    double ast = Pa[s * nso + t], bst = Pb[s * nso + t], auv = Pa[u * nso + v], buv = Pb[u * nso + v];
    double asv = Pa[s * nso + v], bsv = Pb[s * nso + v], aut = Pa[u * nso + t], but = Pb[u * nso + t];
    return __dmul_rn(ast + bst, auv + buv) - __dmul_rn(asv, aut) - __dmul_rn(bsv, but);
}

#ifndef VB_TILE_THREADS
#define VB_TILE_THREADS 256
#endif
#ifndef VB_MINBLOCKS
#define VB_MINBLOCKS 1
#endif
constexpr int TILE_THREADS = VB_TILE_THREADS;
constexpr int HMAX_UNR = 9;                      // pt_ne(pp)
constexpr int HMAX_GEN = 31;                     // pt_ne(dd)
constexpr int GEN_PER_THREAD = GEN_SCRATCH + HMAX_GEN * HMAX_GEN;

// General cofactor weight (vb_cofactor.h): sum over determinant pairs of
//   C2a C0b + C1a(s,t) C1b(u,v) + C1b(s,t) C1a(u,v) + C0a C2b
struct CofBlock {
    double dN, pZ, pw0, pw1;
    int nz;
    const double *G, *uz, *vz;
};
__device__ __forceinline__ double cof_c1(const CofBlock& B, int nso, int s, int t)
{
    double v = B.pZ * B.G[s * nso + t];
    if (B.nz > 0) v += B.pw0 * B.uz[s] * B.vz[t];
    if (B.nz > 1) v += B.pw1 * B.uz[nso + s] * B.vz[nso + t];
    return B.dN * v;
}
__device__ __forceinline__ double cof_c2(const CofBlock& B, int nso, int s, int t, int u, int v)
{
    const double gst = B.G[s * nso + t], guv = B.G[u * nso + v], gsv = B.G[s * nso + v], gut = B.G[u * nso + t];
    // products rounded separately (no FMA contraction): identical operands must cancel exactly
    double r = B.pZ * (__dmul_rn(gst, guv) - __dmul_rn(gsv, gut));
    for (int z = 0; z < B.nz; ++z) {
        const double* uz = B.uz + z * nso;
        const double* vz = B.vz + z * nso;
        r += (z == 0 ? B.pw0 : B.pw1) * (vz[v] * uz[u] * gst - vz[v] * uz[s] * gut - vz[t] * uz[u] * gsv + vz[t] * uz[s] * guv);
    }
    if (B.nz == 2)
        r += (B.vz[t] * B.vz[nso + v] - B.vz[nso + t] * B.vz[v]) * (B.uz[s] * B.uz[nso + u] - B.uz[nso + s] * B.uz[u]);
    return B.dN * r;
}
__device__ __noinline__ double w_general(const double* __restrict__ cof, int ndp, long long stride, int nso, int s, int t, int u, int v)
{
    double sum = 0.0;
    for (int d = 0; d < ndp; ++d) {
        const double* D = cof + (long long)d * stride;
        CofBlock a, b;
        a.dN = D[1]; a.pZ = D[2]; a.pw0 = D[3]; a.pw1 = D[4]; a.nz = (int)D[5];
        b.dN = D[6]; b.pZ = D[7]; b.pw0 = D[8]; b.pw1 = D[9]; b.nz = (int)D[10];
        a.G = D + COF_HEADER; b.G = a.G + nso * nso;
        a.uz = b.G + nso * nso; a.vz = a.uz + 2 * nso; b.uz = a.vz + 2 * nso; b.vz = b.uz + 2 * nso;
        if (a.dN == 0.0 && b.dN == 0.0) continue;
        const double c0a = a.dN * a.pZ, c0b = b.dN * b.pZ;
        double w = 0.0;
        if (c0b != 0.0) w += cof_c2(a, nso, s, t, u, v) * c0b;
        if (c0a != 0.0) w += cof_c2(b, nso, s, t, u, v) * c0a;
        w += cof_c1(a, nso, s, t) * cof_c1(b, nso, u, v) + cof_c1(b, nso, s, t) * cof_c1(a, nso, u, v);
        sum += D[0] * w;
    }
    return sum;
}

// One warp-batch of class (TB|TK): the warp owns bra shell pair `sp`; lane = one ket primitive
// pair of type TK (flat list, shell pairs contiguous).  Every lane runs the same trip count
// (the bra contraction), so there is no divergence inside the batch.  The contracted partial
// blocks of lanes that belong to the same ket shell pair are summed by a segmented warp
// reduction, then the segment heads are half-transformed with the staged ket densities.
template <int TB, int TK>
__device__ __forceinline__ void batch_unrolled(const TileArgs& A, const SPRec& sp, const PrimPair* __restrict__ bpp,
                                               const PrimPair* __restrict__ kpp, int npQ, int base, int nket, int lane,
                                               const double* __restrict__ Dq, double* __restrict__ H, unsigned long long& npq)
{
    constexpr int LA = pt_la(TB), EA = pt_E(TB), LC = pt_la(TK), EC = pt_E(TK), M = EA + EC;
    constexpr int NE = pt_ne(TB), NF = pt_ne(TK);
    double acc[NE * NF];
#pragma unroll
    for (int i = 0; i < NE * NF; ++i) acc[i] = 0.0;
    const bool active = base + lane < nket;
    PrimPair b = kpp[base + (active ? lane : 0)];
    int seg = active ? b.eoff : -1 - lane;
    if (!active) { b.Kp = 0.0; b.w = 0.0; }
    // upper bound of the ket weights of this batch (shell pairs are sorted by weight) -> how many
    // (sorted) bra primitives can still matter
    const double wq = kpp[base].wseg;
    int nbra = 0;
    for (int ip = 0; ip < sp.pp_cnt; ++ip) {
        const PrimPair a = bpp[ip];
        if (!(a.w * wq >= A.tau)) break;           // bra primitives are sorted by weight
        ++nbra;
        QuartetGeom g;
        double T, pref;
        quartet_geom(a, b, g, T, pref);
        double F[M + 1];
        boys<M>(A.boys, T, F);
#pragma unroll
        for (int m = 0; m <= M; ++m) F[m] *= pref;
        if constexpr (M == 0) acc[0] += F[0];
        else vrr_unrolled<LA, EA, LC, EC>(g, F, acc);
    }
    if (nbra == 0) return;
    if (lane == 0) npq += (unsigned long long)nbra * min(32, nket - base);
    (void)A;
    // segmented reduction over lanes of the same ket shell pair (segments are contiguous)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int so = __shfl_down_sync(0xffffffffu, seg, o);
        const bool take = (lane + o < 32) && (so == seg);
        if (!__any_sync(0xffffffffu, take)) break;     // no run is longer than the stride: done
#pragma unroll
        for (int i = 0; i < NE * NF; ++i) {
            const double v = __shfl_down_sync(0xffffffffu, acc[i], o);
            if (take) acc[i] += v;
        }
    }
    const int sprev = __shfl_up_sync(0xffffffffu, seg, 1);
    unsigned heads = __ballot_sync(0xffffffffu, active && (lane == 0 || sprev != seg));
    while (heads) {
        const int src = __ffs(heads) - 1;
        heads &= heads - 1;
        const int eo = __shfl_sync(0xffffffffu, seg, src);
        const double* row = Dq + eo * npQ + (lane < npQ ? lane : 0);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const double dq = row[f * npQ];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const double v = __shfl_sync(0xffffffffu, acc[e * NF + f], src);
                H[e] += v * dq;
            }
        }
    }
}

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8):  lane = 4*g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// generic path (any class with a d shell): runtime loops, scratch in global memory
__device__ __noinline__ void batch_generic(const TileArgs& A, int tb, int tk, const SPRec& sp, const PrimPair* __restrict__ bpp,
                                           const PrimPair* __restrict__ kpp, int npQ, int base, int nket,
                                           int lane, const double* __restrict__ Dq, double* __restrict__ H,
                                           double* __restrict__ scratch, unsigned long long& npq)
{
    const int LA = pt_la(tb), EA = pt_E(tb), LC = pt_la(tk), EC = pt_E(tk), M = EA + EC;
    const int NE = pt_ne(tb), NF = pt_ne(tk);
    double* T = scratch;                       // GEN_SCRATCH
    double* acc = scratch + GEN_SCRATCH;       // up to 31*31
    for (int i = 0; i < NE * NF; ++i) acc[i] = 0.0;
    const bool active = base + lane < nket;
    PrimPair b = kpp[base + (active ? lane : 0)];
    const int eoff = b.eoff;
    if (!active) { b.Kp = 0.0; b.w = 0.0; }
    double wq = b.w;
    for (int o = 16; o > 0; o >>= 1) wq = fmax(wq, __shfl_xor_sync(0xffffffffu, wq, o));
    int nbra = 0;
    for (; nbra < sp.pp_cnt; ++nbra)
        if (!(bpp[nbra].w * wq >= A.tau)) break;
    if (nbra == 0) return;
    if (lane == 0) npq += (unsigned long long)nbra * min(32, nket - base);
    if (active)
        for (int ip = 0; ip < nbra; ++ip) {
            const PrimPair a = bpp[ip];
            QuartetGeom g;
            double Tt, pref, F[MTOP + 1];
            quartet_geom(a, b, g, Tt, pref);
            boys_rt(M, A.boys, Tt, F);
            for (int m = 0; m <= M; ++m) F[m] *= pref;
            vrr_generic(LA, EA, LC, EC, g, F, T, acc);
        }
    const int cnt = min(32, nket - base);
    for (int src = 0; src < cnt; ++src) {
        const int eo = __shfl_sync(0xffffffffu, eoff, src);
        for (int f = 0; f < NF; ++f) {
            const double dq = (lane < npQ) ? Dq[(eo + f) * npQ + lane] : 0.0;
            for (int e = 0; e < NE; ++e) {
                const double v = __shfl_sync(0xffffffffu, acc[e * NF + f], src);
                H[e] += v * dq;
            }
        }
    }
}


// add a warp's half-transformed rows of one bra shell pair into the CTA's H tile (lane = ket pair q)
template <int NEH>
__device__ __forceinline__ void add_rows(double* __restrict__ Hs, int hs_ld, int eoff, const double* __restrict__ H, int lane)
{
#pragma unroll
    for (int e = 0; e < NEH; ++e) atomicAdd(&Hs[(eoff + e) * hs_ld + lane], H[e]);
}

template <bool GEN>
__global__ void __launch_bounds__(TILE_THREADS, GEN ? 1 : VB_MINBLOCKS) k_tile(const TileArgs A)
{
    extern __shared__ __align__(16) double smem[];
    double* Dq = smem;                                          // [Q.ne][Q.np]
    double* Hs = Dq + A.dq_cap;                                 // half-transformed tile [P.ne][hs_ld]
    SPRec* spss = reinterpret_cast<SPRec*>(Hs + A.hs_cap);      // bra shell pairs of P
    PrimPair* kpp_s = reinterpret_cast<PrimPair*>(spss + A.sp_cap);   // ket primitive pairs of Q (when they fit)
    PrimPair* bpp_s = kpp_s + A.pp_cap;                         // bra primitive pairs of P
    __shared__ unsigned long long s_bar;
    unsigned phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int s_tile, s_unit;
    __shared__ int s_cum[NPTYPE * NPTYPE + 1];
    __shared__ double s_red[TILE_THREADS / 32];
    __shared__ unsigned long long s_cnt[CNT_N];
    __shared__ unsigned long long s_pq[NPTYPE * NPTYPE];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = TILE_THREADS / 32;
    constexpr int HM = GEN ? HMAX_GEN : HMAX_UNR;
    double* scratch = nullptr;
    if constexpr (GEN) scratch = A.gen_scratch + ((size_t)blockIdx.x * TILE_THREADS + tid) * GEN_PER_THREAD;

    if (tid < CNT_N) s_cnt[tid] = 0ull;
    if (tid < NPTYPE * NPTYPE) s_pq[tid] = 0ull;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = (int)atomicAdd(A.counter, 1u);   // work stealing inside this rank's shard
        __syncthreads();
        const int hs = A.hsplit > 1 ? A.hsplit : 1;
        const long long wloc = A.hphase == 1 ? (long long)s_tile / hs : (long long)s_tile;     // this rank's tile number
        const int slice = A.hphase == 1 ? (int)((long long)s_tile % hs) : 0;
        const long long tl = (long long)A.tile_first + wloc * A.tile_stride;
        if (tl >= A.ntiles) break;
        const int2 tq = A.tiles[tl];
        const PGDesc P = A.pgs[tq.x];
        const PGDesc Q = A.pgs[tq.y];
        if (A.hphase == 2) {
            // the half-transformed tile = sum of the slices' shares, in slice order
            const double* __restrict__ hp = A.hpart + (size_t)wloc * hs * A.hs_cap;
            for (int i = tid; i < P.ne * A.hs_ld; i += TILE_THREADS) {
                double v = 0.0;
                for (int k = 0; k < hs; ++k) v += hp[(size_t)k * A.hs_cap + i];
                Hs[i] = v;
            }
            __syncthreads();
        } else {
        // stage the tile's tables with TMA bulk copies (one elected thread issues, all wait on the mbarrier)
        const int nkpp = Q.pp_beg[NPTYPE] - Q.pp_beg[0], nbpp = P.pp_beg[NPTYPE] - P.pp_beg[0];
        const int nspP = P.sp_beg[NPTYPE] - P.sp_beg[0];
        if (tid == 0) {
            // the buffers were last read through the generic proxy (previous tile): order before async writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bq = (unsigned)((((size_t)Q.ne * Q.np + 1) & ~(size_t)1) * sizeof(double));
            const unsigned bk = A.pp_cap ? (unsigned)(nkpp * sizeof(PrimPair)) : 0u, bb = A.pp_cap ? (unsigned)(nbpp * sizeof(PrimPair)) : 0u;
            const unsigned bs = (unsigned)(nspP * sizeof(SPRec));
            mbar_expect_tx(&s_bar, bq + bk + bb + bs);
            tma_bulk_g2s(Dq, A.dmat + Q.d_off, bq, &s_bar);
            tma_bulk_g2s(spss, A.sps + P.sp_beg[0], bs, &s_bar);
            if (A.pp_cap) {
                tma_bulk_g2s(kpp_s, A.pps + Q.pp_beg[0], bk, &s_bar);
                tma_bulk_g2s(bpp_s, A.pps + P.pp_beg[0], bb, &s_bar);
            }
        }
        // large orbital basis sets: the primitive tables stay in global memory (L1/L2)
        const PrimPair* kpps = A.pp_cap ? kpp_s : A.pps + Q.pp_beg[0];
        const PrimPair* bpps = A.pp_cap ? bpp_s : A.pps + P.pp_beg[0];
        for (int i = tid; i < P.ne * A.hs_ld; i += TILE_THREADS) Hs[i] = 0.0;
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
        __syncthreads();

        // First half transformation.  All warps walk the classes (tb|tk) in the same order, so at any
        // time the CTA executes one class body (instruction-cache friendly); inside a class the bra
        // shell pairs of type tb are dealt round-robin to the warps in cost order.  The assignment is
        // static, hence every tile result is bitwise reproducible (all ranks derive the same Schwarz
        // table), and each row of Hs is only ever touched by one warp.
        // Units (class, bra shell pair) in class order; most expensive shell pairs first inside a class.
        //   Schwarz pass: dealt round-robin (static), so the table is bitwise reproducible and
        //                 identical on every rank (the tile list is derived from it);
        //   energy pass : handed out through a shared counter (dynamic), which keeps all warps busy
        //                 until the tile is done; rows of Hs are accumulated with shared atomics.
        constexpr int NT = GEN ? NPTYPE : 3;
        if (tid == 0) {
            int n = 0;
            for (int tb = 0; tb < NT; ++tb)
                for (int tk = 0; tk < NT; ++tk) {
                    const int nspb = P.sp_beg[tb + 1] - P.sp_beg[tb], nket = Q.pp_beg[tk + 1] - Q.pp_beg[tk];
                    s_cum[tb * NT + tk] = n;
                    if (nspb > 0 && nket > 0 && P.kwmax[tb] * Q.kwmax[tk] >= A.tau) n += nspb;
                }
            s_cum[NT * NT] = n;
            s_unit = 0;
        }
        __syncthreads();
        const int nunits = s_cum[NT * NT];
        const bool dynamic = A.mode != 0 && A.hphase != 1;
        const int KS = (A.hphase == 1 && A.hksplit > 1) ? A.hksplit : 1, US = hs / KS;      // slice = (us, ks)
        const int us = slice % US, ks = slice / US;
        int ustat = us + warp * US;
        for (;;) {
            int u;
            if (dynamic) {
                u = 0;
                if (lane == 0) u = atomicAdd(&s_unit, 1);
                u = __shfl_sync(0xffffffffu, u, 0);
            } else {
                u = ustat;
                ustat += nw * US;
            }
            if (u >= nunits) break;
            int c = 0;
            while (s_cum[c + 1] <= u) ++c;
            const int tb = c / NT, tk = c - tb * NT;
            const int nket = Q.pp_beg[tk + 1] - Q.pp_beg[tk];
            const int cls = tb * NPTYPE + tk;
            const SPRec sp = spss[P.sp_beg[tb] - P.sp_beg[0] + (u - s_cum[c])];
            if (!(sp.wmax * Q.kwmax[tk] >= A.tau)) continue;
            const PrimPair* bpp = bpps + (sp.pp_beg - P.pp_beg[0]);
            const PrimPair* kpp = kpps + (Q.pp_beg[tk] - Q.pp_beg[0]);
            double H[HM];
#pragma unroll
            for (int e = 0; e < HM; ++e) H[e] = 0.0;
            unsigned long long npq = 0ull;
            for (int base = 32 * ks; base < nket; base += 32 * KS) {
                // ket shell pairs are sorted by weight: once a batch is negligible, so are the rest
                if (!(sp.wmax * kpp[base].wseg >= A.tau)) break;
                switch (cls) {
#define VB_CASE(TB, TK) case TB * NPTYPE + TK: batch_unrolled<TB, TK>(A, sp, bpp, kpp, Q.np, base, nket, lane, Dq, H, npq); break;
                    VB_CASE(0, 0) VB_CASE(0, 1) VB_CASE(0, 2)
                    VB_CASE(1, 0) VB_CASE(1, 1) VB_CASE(1, 2)
                    VB_CASE(2, 0) VB_CASE(2, 1) VB_CASE(2, 2)
#undef VB_CASE
                    default:
                        if constexpr (GEN) batch_generic(A, tb, tk, sp, bpp, kpp, Q.np, base, nket, lane, Dq, H, scratch, npq);
                        break;
                }
            }
            npq = __shfl_sync(0xffffffffu, npq, 0);
            if (npq == 0ull) continue;
            if (lane == 0) atomicAdd(&s_pq[cls], npq);
            if (lane < Q.np) {
                switch (pt_ne(tb)) {
                    case 1: add_rows<1>(Hs, A.hs_ld, sp.eoff, H, lane); break;
                    case 3: add_rows<3>(Hs, A.hs_ld, sp.eoff, H, lane); break;
                    case 9: add_rows<9>(Hs, A.hs_ld, sp.eoff, H, lane); break;
                    default:
                        if constexpr (GEN)
                            for (int e = 0; e < pt_ne(tb); ++e) atomicAdd(&Hs[(sp.eoff + e) * A.hs_ld + lane], H[e]);
                        break;
                }
            }
        }
        __syncthreads();
        if (A.hphase == 1) {
            double* __restrict__ hp = A.hpart + ((size_t)wloc * hs + slice) * A.hs_cap;
            for (int i = tid; i < P.ne * A.hs_ld; i += TILE_THREADS) hp[i] = Hs[i];
            continue;
        }
        }   // hphase != 2

        // ---- contraction with the cofactor densities --------------------------------------
        double epart = 0.0;
        unsigned long long cnt[CNT_N];
#pragma unroll
        for (int i = 0; i < CNT_N; ++i) cnt[i] = 0ull;
        const bool diag_tile = tq.x == tq.y;
        const int nso = A.nso;
        for (int idx = tid; idx < P.np * Q.np; idx += TILE_THREADS) {
            const int p = idx / Q.np, q = idx % Q.np;
            if (diag_tile && q > p) continue;
            const int s = A.pg_pairs[2 * (P.pair_beg + p)], t = A.pg_pairs[2 * (P.pair_beg + p) + 1];
            const int u = A.pg_pairs[2 * (Q.pair_beg + q)], v = A.pg_pairs[2 * (Q.pair_beg + q) + 1];
            if (A.mode == 0 && !(diag_tile && p == q)) continue;
            bool ssig = true;
            if (A.mode == 1) {
                // reference screen on the Schwarz product (valence.F90:1189-1190); screened entries
                // contribute nothing and are counted nowhere, so their integral is not even formed
                ssig = A.sch[s * nso + t] * A.sch[u * nso + v] > A.itol;
                if (!ssig) continue;
            }
            // second half transformation for this entry: (st|uv) = sum_e Dp[e][p] Hs[e][q]
            double G = 0.0;
            {
                const double* Dp = A.dmat + P.d_off + p;
                for (int e = 0; e < P.ne; ++e) G += Dp[(size_t)e * P.np] * Hs[e * A.hs_ld + q];
            }
            if (A.mode == 0) {
                if (diag_tile && p == q) {
                    A.diag[(size_t)s * nso + t] = G;
                    if (A.sym) A.diag[(size_t)t * nso + s] = G;
                }
                continue;
            }
            if (A.mode == 2) {
                const int i1 = A.pair_index[s * nso + t], i2 = A.pair_index[u * nso + v];
                A.gfull[(size_t)i1 * A.npairs_total + i2] = G;
                A.gfull[(size_t)i2 * A.npairs_total + i1] = G;
                if (A.sym) {
                    const int j1 = A.pair_index[t * nso + s], j2 = A.pair_index[v * nso + u];
                    A.gfull[(size_t)j1 * A.npairs_total + i2] = G; A.gfull[(size_t)i2 * A.npairs_total + j1] = G;
                    A.gfull[(size_t)i1 * A.npairs_total + j2] = G; A.gfull[(size_t)j2 * A.npairs_total + i1] = G;
                    A.gfull[(size_t)j1 * A.npairs_total + j2] = G; A.gfull[(size_t)j2 * A.npairs_total + j1] = G;
                }
                continue;
            }
            // images of (s,t,u,v) under the integral's permutational symmetry
            int im[8][4];
            int nim = 0;
            {
                const int base4[2][4] = {{s, t, u, v}, {u, v, s, t}};
                for (int k = 0; k < 2; ++k) {
                    const int a = base4[k][0], b = base4[k][1], c = base4[k][2], d = base4[k][3];
                    im[nim][0] = a; im[nim][1] = b; im[nim][2] = c; im[nim][3] = d; ++nim;
                    if (A.sym) {
                        im[nim][0] = b; im[nim][1] = a; im[nim][2] = c; im[nim][3] = d; ++nim;
                        im[nim][0] = a; im[nim][1] = b; im[nim][2] = d; im[nim][3] = c; ++nim;
                        im[nim][0] = b; im[nim][1] = a; im[nim][2] = d; im[nim][3] = c; ++nim;
                    }
                }
            }
            double wsum = 0.0;
            int stab = 0;
            cnt[CNT_ENTRIES]++;
            for (int k = 0; k < nim; ++k) {
                const int a = im[k][0], b = im[k][1], c = im[k][2], d = im[k][3];
                if (a == s && b == t && c == u && d == v) ++stab;
                bool dup = false;
                for (int k2 = 0; k2 < k; ++k2)
                    dup = dup || (im[k2][0] == a && im[k2][1] == b && im[k2][2] == c && im[k2][3] == d);
                if (dup) continue;
                // task bookkeeping exactly as the reference visits it (valence.F90:1167-1190)
                const bool shortcut = (a == c && b == d) && a != A.subject && b != A.subject;   // :1213 (nonsub)
                const double val = shortcut ? A.sch[a * nso + b] * A.sch[a * nso + b] : G;
                const bool vsig = ssig && fabs(val) > A.itol;
                // as the direct integral of task io=a, ko=b, jo=c, lo=d
                bool vd = a >= c && b >= d && !((a == c && a < A.nnd) || (b == d && b < A.nnd));
                if (vd && A.sym) vd = tri_index(a, c) >= tri_index(b, d);
                // as the exchanged integral of task io=a, lo=b, jo=c, ko=d
                bool vx = a >= c && d >= b && !((a == c && a < A.nnd) || (b == d && b < A.nnd));
                if (vx && A.sym) vx = tri_index(a, c) >= tri_index(d, b);
                if (vd) { cnt[CNT_SCHWARZ_EREP] += ssig; cnt[CNT_VALUE_EREP] += vsig; }
                if (vx) { cnt[CNT_SCHWARZ_EXCH] += ssig; cnt[CNT_VALUE_EXCH] += vsig; }
                if (ssig && !shortcut) {
                    const int calls = (vd ? 1 : 0) + ((vx && b != d) ? 1 : 0);
                    cnt[CNT_INT2E] += calls;
                    cnt[CNT_SHELLQ] += (unsigned long long)calls * A.nsh_bra[a] * A.nsh_ket[b] * A.nsh_bra[c] * A.nsh_ket[d];
                }
                if (ssig && shortcut && vd) cnt[CNT_SHORTCUT]++;
                if (A.debug) printf("ENTRY (%d %d|%d %d) val %.12f W %.12f vsig %d\n", a, b, c, d, val, A.ndp ? w_general(A.cof, A.ndp, A.cof_stride, nso, a, b, c, d) : w_term(A.Pa, A.Pb, nso, a, b, c, d), (int)vsig);
                if (vsig) wsum += val * (A.ndp ? w_general(A.cof, A.ndp, A.cof_stride, nso, a, b, c, d) : w_term(A.Pa, A.Pb, nso, a, b, c, d));
            }
            (void)stab;
            epart += 0.5 * wsum;
        }
        if (A.mode == 1) {
            // deterministic block reduction of the tile's energy
            for (int o = 16; o > 0; o >>= 1) epart += __shfl_down_sync(0xffffffffu, epart, o);
            if (lane == 0) s_red[warp] = epart;
            for (int i = 0; i < CNT_N; ++i)
                if (cnt[i]) atomicAdd(&s_cnt[i], cnt[i]);
            __syncthreads();
            if (tid == 0) {
                double e = 0.0;
                for (int w = 0; w < nw; ++w) e += s_red[w];
                A.tileE[tl] = e * A.c0;
            }
        }
    }
    // one flush per CTA (same-address global atomics per tile would serialise in L2)
    __syncthreads();
    if (tid < CNT_N && s_cnt[tid]) atomicAdd(&A.counters[tid], s_cnt[tid]);
    if (tid < NPTYPE * NPTYPE && s_pq[tid]) atomicAdd(&A.pq_counters[tid], s_pq[tid]);
}

}  // namespace vb
