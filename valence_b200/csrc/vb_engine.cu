// GPU VSVB energy engine: host orchestration.  See vb_kernels.cuh / vb_tile.cuh for the kernels
// and DESIGN.md for the data layout.  No CPU fallback exists: every integral, inverse and
// contraction below runs in a CUDA kernel, and construction throws when no device is usable.
#include "vb_engine.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <atomic>

#include "vb_cofactor.h"
#include "vb_nccl.h"
#include "vb_kernels.cuh"
#include "vb_tilelist.h"
#include "vb_ptile.cuh"
#include "vb_pseg.cuh"
#include "vb_tile.cuh"

namespace vb {

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #call); \
    } while (0)

namespace {

// dynamic shared memory k_ptile may ask for: 227 KB per CTA minus its static part
constexpr size_t PT_SMEM_MAX = 225 * 1024;
static_assert(sizeof(TilePair) == sizeof(int2) && sizeof(WorkItem) == sizeof(int4) && alignof(WorkItem) == alignof(int4), "tile list records must match int2 / int4");
static_assert(PT_MAXQ == TILES_PER_ITEM_MAX, "work items hold at most PT_MAXQ tiles");

long long g_h2d_bytes = 0, g_d2h_bytes = 0;   // host<->device traffic of the current call

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0, n = 0;
    ~DBuf() { if (p) cudaFree(p); }
    void alloc(size_t m)
    {
        if (m > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            size_t want = m + m / 8 + 16;
            CK(cudaMalloc((void**)&p, want * sizeof(T)));
            cap = want;
        }
        n = m;
    }
    void upload(const std::vector<T>& v, cudaStream_t st)
    {
        alloc(v.size());
        if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        g_h2d_bytes += (long long)(v.size() * sizeof(T));
    }
    void upload_at(size_t off, const std::vector<T>& v, cudaStream_t st)   // into an allocation made with alloc()
    {
        if (off + v.size() > cap) throw std::runtime_error("valence_b200: device table overflow");
        if (!v.empty()) CK(cudaMemcpyAsync(p + off, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        g_h2d_bytes += (long long)(v.size() * sizeof(T));
    }
    void upload_at(size_t off, const T* src, size_t cnt, cudaStream_t st)   // straight from the caller's storage (no temporary)
    {
        if (off + cnt > cap) throw std::runtime_error("valence_b200: device table overflow");
        if (cnt) CK(cudaMemcpyAsync(p + off, src, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
        g_h2d_bytes += (long long)(cnt * sizeof(T));
    }
    void zero(cudaStream_t st) { if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
    void download(std::vector<T>& v, cudaStream_t st)
    {
        v.resize(n);
        if (n) CK(cudaMemcpyAsync(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_d2h_bytes += (long long)(n * sizeof(T));
    }
};

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// algorithmic FP64 operation count of one primitive quartet of class (tb | tk); FMA = 2
// far != 0: the quartet is in the asymptotic regime T >= 40 (one reciprocal square root instead of the table look-up and
// Taylor series; (ss|ss) also skips the general geometry): what the class kernels execute for it
double flops_prim_quartet(int tb, int tk, int far = 0)
{
    const int LA = pt_la(tb), EA = pt_E(tb), LC = pt_la(tk), EC = pt_E(tk), M = EA + EC;
    if (far && M == 0) return 27.0;                   // |PQ|^2, p q, test, rsqrt (9), K_a K_b sqrt(pi)/2, accumulate
    double f = 45.0;                                  // geometry, T, prefactor (1 rsqrt, 1 div)
    if (far) f += 10.0 + 3.0 * M;                     // F_0 = sqrt(pi/T)/2, upward recursion
    else f += 16.0 + (M > 0 ? 25.0 + 3.0 * M : 0.0);  // Boys: 8-term Taylor (+ exp and downward recursion)
    f += M + 1;                                       // prefactor scaling
    const int NE = ncum(EA), NF = ncum(EC);
    for (int e = 1; e < NE; ++e) {
        int d = c_dir(e), e1 = c_dec(e, d), n1 = c_l(e1, d);
        f += (M + 1 - c_L(e)) * (3.0 + (n1 > 0 ? 5.0 : 0.0));
    }
    for (int ff = 1; ff < NF; ++ff) {
        int d = c_dir(ff), f1 = c_dec(ff, d), n1 = c_l(f1, d);
        for (int e = 0; e < NE; ++e) {
            int cnt = M + 1 - c_L(e) - c_L(ff);
            if (cnt <= 0) continue;
            f += cnt * (3.0 + (n1 > 0 ? 5.0 : 0.0) + (c_l(e, d) > 0 ? 3.0 : 0.0));
        }
    }
    f += (double)(NE - coff(LA)) * (NF - coff(LC));   // accumulation into the contracted block
    return f;
}


// ---- class-split tile pass (vb_pclass.cuh): one launch per integral class into gbuf, then (optionally) the contraction ----
struct ClassPlan {
    bool present[3][3];
    ClassCfg cfg[3];
    size_t smem[3][3];
    bool ok = true;
    // segment form of the four light classes (vb_pseg.cuh): far-field tables of this evaluation, or null -> k_pclass for every class
    const FarPrim* fcp = nullptr;      // aligned with pps
    const FarPrim* fcpf = nullptr;     // aligned with pps_flat
    unsigned lthr = 0;
    size_t smem_seg[2] = {0, 0};
};

ClassPlan plan_classes(const std::vector<PGDesc>& pgs)
{
    ClassPlan pl;
    int mx_d[3] = {0, 0, 0}, mx_sp[3] = {0, 0, 0}, mx_pp[3] = {0, 0, 0};
    bool bra[3] = {false, false, false};
    for (const PGDesc& pg : pgs)
        for (int t = 0; t < 3; ++t) {
            const int nsp = pg.sp_beg[t + 1] - pg.sp_beg[t], npp = pg.pp_beg[t + 1] - pg.pp_beg[t];
            if (nsp > 0 && npp > 0) bra[t] = true;
            mx_d[t] = std::max(mx_d[t], (pg.e_beg[t + 1] - pg.e_beg[t]) * pg.np);
            mx_sp[t] = std::max(mx_sp[t], nsp);
            mx_pp[t] = std::max(mx_pp[t], npp);
        }
    for (int tb = 0; tb < 3; ++tb) {
        pl.cfg[tb].d_cap = ((mx_d[tb] + 3) & ~1) + 2;
        pl.cfg[tb].sp_cap = std::max(1, mx_sp[tb]);
        pl.cfg[tb].pp_cap = std::max(1, mx_pp[tb]);
        for (int tk = 0; tk < 3; ++tk) {
            pl.present[tb][tk] = bra[tb] && bra[tk];
            const int nw = pc_threads(tb, tk) / 32;
            pl.smem[tb][tk] = ((size_t)pl.cfg[tb].d_cap + (size_t)nw * PT_SCRATCH + BOYS_S_SIZE) * sizeof(double) +
                              (size_t)pl.cfg[tb].sp_cap * sizeof(SPRec) + (size_t)pl.cfg[tb].pp_cap * sizeof(PrimPair);
            if (pl.present[tb][tk] && pl.smem[tb][tk] > PT_SMEM_MAX) pl.ok = false;
        }
        if (tb < 2)
            pl.smem_seg[tb] = ((size_t)pl.cfg[tb].d_cap + (size_t)(PS_THREADS / 32) * PT_SCRATCH + BOYS_S_SIZE) * sizeof(double) +
                              (size_t)pl.cfg[tb].sp_cap * sizeof(SPRec) + ((size_t)pl.cfg[tb].pp_cap + FAR_PAD) * sizeof(FarPrim) +
                              (size_t)pl.cfg[tb].pp_cap * sizeof(PrimPair);
    }
    return pl;
}

template <int TB, int TK>
void launch_seg(const TileArgs& A, const ClassPlan& pl, int nsm, int nitems, cudaStream_t st)
{
    const size_t smem = pl.smem_seg[TB];
    CK(cudaFuncSetAttribute(k_pseg<TB, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pseg<TB, TK>, PS_THREADS, smem));
    per_sm = std::max(1, per_sm);
    const int grid = std::max(1, std::min(nsm * per_sm, nitems));
    SegCfg C;
    C.d_cap = pl.cfg[TB].d_cap; C.sp_cap = pl.cfg[TB].sp_cap; C.pp_cap = pl.cfg[TB].pp_cap; C.fcp = pl.fcp; C.fcpf = pl.fcpf; C.lthr = pl.lthr;
    C.skip = 0;
    if (const char* e = std::getenv("VB_SEG_SKIP")) C.skip = std::atoi(e);
    k_pseg<TB, TK><<<grid, PS_THREADS, smem, st>>>(A, C);
    CK(cudaGetLastError());
}

template <int TB, int TK>
void launch_class(const TileArgs& A, const ClassPlan& pl, int nsm, int nitems, cudaStream_t st)
{
    const size_t smem = pl.smem[TB][TK];
    CK(cudaFuncSetAttribute(k_pclass<TB, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pclass<TB, TK>, pc_threads(TB, TK), smem));
    per_sm = std::max(1, per_sm);
    const int grid = std::max(1, std::min(nsm * per_sm, nitems));
    k_pclass<TB, TK><<<grid, pc_threads(TB, TK), smem, st>>>(A, pl.cfg[TB]);
    CK(cudaGetLastError());
}

// all classes of the items A.items[0 .. A.nitems) into A.gbuf (zeroed here); returns the number of launches
int run_class_pass(TileArgs A, const ClassPlan& pl, int nsm, long long ntile_slots, unsigned int* counter, cudaStream_t st, float* ms_class /* 9 or null */,
                   cudaEvent_t e0, cudaEvent_t e1)
{
    CK(cudaMemsetAsync(A.gbuf, 0, (size_t)ntile_slots * A.g_cap * sizeof(double), st));
    int n = 0;
    // heaviest classes first
    static const int order[9][2] = {{2, 2}, {2, 1}, {1, 2}, {2, 0}, {0, 2}, {1, 1}, {1, 0}, {0, 1}, {0, 0}};
    for (const auto& c : order) {
        const int tb = c[0], tk = c[1];
        if (!pl.present[tb][tk]) continue;
        CK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
        if (ms_class) CK(cudaEventRecord(e0, st));
        const bool seg = pl.fcp && tb < 2 && tk < 2 && pl.smem_seg[tb] <= PT_SMEM_MAX;
        if (seg) {
            switch (tb * 2 + tk) {
                case 0: launch_seg<0, 0>(A, pl, nsm, A.nitems, st); break;
                case 1: launch_seg<0, 1>(A, pl, nsm, A.nitems, st); break;
                case 2: launch_seg<1, 0>(A, pl, nsm, A.nitems, st); break;
                default: launch_seg<1, 1>(A, pl, nsm, A.nitems, st); break;
            }
        } else
        switch (tb * 3 + tk) {
            case 0: launch_class<0, 0>(A, pl, nsm, A.nitems, st); break;
            case 1: launch_class<0, 1>(A, pl, nsm, A.nitems, st); break;
            case 2: launch_class<0, 2>(A, pl, nsm, A.nitems, st); break;
            case 3: launch_class<1, 0>(A, pl, nsm, A.nitems, st); break;
            case 4: launch_class<1, 1>(A, pl, nsm, A.nitems, st); break;
            case 5: launch_class<1, 2>(A, pl, nsm, A.nitems, st); break;
            case 6: launch_class<2, 0>(A, pl, nsm, A.nitems, st); break;
            case 7: launch_class<2, 1>(A, pl, nsm, A.nitems, st); break;
            default: launch_class<2, 2>(A, pl, nsm, A.nitems, st); break;
        }
        if (ms_class) {
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms_class[tb * 3 + tk] += ms;
        }
        ++n;
    }
    return n;
}

}  // namespace

struct Engine::Impl {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    int nsm = 0;
    std::vector<std::vector<double>> coeff;          // current orbital weights (normalised in energy())
    std::vector<double> xyz_angs;
    // device data
    DBuf<FarPrim> fcp, fcpf;
    DBuf<double> hpart;                              // d-shell tile kernel: per-slice shares of the half-transformed tiles
    DBuf<double> boys, boys_small, gbuf, gred, exps, coefs, nuc, S, H, Se, He, Ma, Mb, Mai, Mbi, gj_ws, gj_res, Pa, Pb, gjout, diag, sch, tileE, accum, one_e, dmat, gen_scratch;
    DBuf<DevShell> shells;
    DBuf<int> optr, oao, piv, ea_bra, ea_ket, eb_bra, eb_ket, posa_bra, posa_ket, posb_bra, posb_ket, pg_pairs, nsh_bra, nsh_ket;
    DBuf<double> oc;
    DBuf<int2> opairs;
    DBuf<TilePair> tiles;       // same layout as int2 / int4 (vb_tilelist.h)
    DBuf<WorkItem> items;
    DBuf<PGDesc> pgs;
    DBuf<SPRec> sps;
    DBuf<PrimPair> pps, pps_flat;
    DBuf<unsigned int> counter;
    DBuf<unsigned long long> counters, pq_counters;
    // state carried from energy_partial to energy_finish
    Basis bas;
    std::vector<ExpOrb> orbs1e, orbs2e;
    std::vector<std::vector<double>> cnorm;          // normalised weights of the current call
    DBuf<double> cof;
    double dtol = 0, itol = 0;
    void prepare(const Input& in, int subject);
    void evaluate(const Input& in, const Wavefunction& wf, const std::vector<double>* sch_in, bool diag_only, int rank, int nranks,
                  EnergyResult* out, std::vector<double>* sch_out);
    bool first_order_cached(const Input& in, const Wavefunction& wf, int es, int iorb, const std::vector<double>& sch, int rank, int nranks,
                            std::vector<double>* ham, std::vector<double>* ovl, EnergyResult* acc);
    TileSetup ts_keep;                               // host tables of the last evaluation (storage kept)
    // Schwarz table of the last guess energy, orbital by orbital, and a digest of the state it belongs to (geometry, orbital
    // weights, screen): first_order_opt needs the same table for its unsubstituted lists (valence.F90:666-667) and takes it from
    // here instead of repeating the table build and the diagonal pass when nothing has changed in between
    std::vector<double> sch_cache;
    int sch_cache_n = 0;
    unsigned long long sch_cache_key = 0;
    unsigned long long state_key(const Input& in) const
    {
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
        mix(xyz_angs.data(), xyz_angs.size() * sizeof(double));
        for (const std::vector<double>& c : coeff) mix(c.data(), c.size() * sizeof(double));
        const double t[3] = {(double)in.ntol_i, (double)in.ntol_d, tau};
        mix(t, sizeof t);
        return h;
    }
    // table build shared by the ranks of a node (Engine::shard_tables): consumed by the next energy_partial
    bool shard_pending = false, prepared = false;
    int shard_rank = 0, shard_nranks = 1;
    std::string shard_prefix, shard_file;
    DBuf<double> gcache;                             // first_order_opt: orbital-level integrals of the subject-free tiles
    DBuf<int4> vtiles;
    DBuf<int> fo_perm;
    void cofactor_stage(const Input& in, const Wavefunction& wf, bool diag_only, EnergyResult* out, bool* fast_out, double* c0_out, int* ndp_out);
    int only_isc = -1, only_jsc = -1;                 // spin_opt: restrict the cofactors to one coupling pair
    // GPU inverse + determinant of Se[ent][ent] (Gauss-Jordan, then one double-double refinement step); res_dev[0] = det,
    // res_dev[1] = smallest / largest pivot
    void gpu_inverse(int nso, const std::vector<int>& ent, DBuf<int>& ent_dev, DBuf<double>& M, DBuf<double>& Minv, double* res_dev);
    void sc_cofactors_gpu(const Input& in, const Wavefunction& wf, EnergyResult* out, int* ndp_out);
    DBuf<int> sc_idx;                                 // spin-coupled determinant pairs on the GPU: entry lists / positions of one pair
    DBuf<double> Fmat, r1x, r1y, Qd, Rd, Ni;       // first_order_opt in rank-one form
    DBuf<int> en_dev, eo_dev, posn_dev, poso_dev;
    bool use_gather = false;                         // energy(): table shares are exchanged on the device (all-gather)
    bool fo_collective = false;                      // first_order through the engine's communicator: decisions are agreed on
    std::vector<double> coeff_sc;                    // current spin-coupling weights
    double enuc = 0, e1 = 0, wfnorm = 0;
    // the same two sums per unit determinant product, as unevaluated (high, low) pairs, and that product: energy_finish
    // assembles E = (E1 + E2 / c0) / (N1 / nelec) + Enuc in extended precision (a 256-molecule cluster has |E_elec| ~ 2e5 Eh,
    // one ulp of that is 3e-11 Eh: five roundings of the plain formula are the whole 1e-10 Eh budget)
    double e1u[2] = {0, 0}, wnu[2] = {0, 0}, c0_keep = 0;
    bool pb_same = false;       // closed shell: Pb holds the same numbers as Pa (the contraction then reads Pa only)
    bool split_sum = false;     // energy_partial: leave E2 as the (grid multiple, rest) pair in accum[0], accum[1 + CNT_N]
    int nelec_keep = 0;
    // primitive-quartet magnitude cut (VB_PRIM_TAU overrides).  Measured on (H2O)_64 / (H2O)_128: the energy is the same to
    // 13 digits for every cut between 1e-24 and 1e-18 (profiles/r2_tau_sweep.log); the screening counters do not depend on it.
    double tau = 1e-20;
    int launches = 0;
    double t_begin = 0;
    std::unique_ptr<Comm> comm;                      // one process per GPU: NCCL communicator of the job (null = alone)
    int crank() const { return comm ? comm->rank() : 0; }
    int csize() const { return comm ? comm->nranks() : 1; }
};

Engine::Engine(const Input& in, int device) : in_(in), impl_(new Impl)
{
    Impl& I = *impl_;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error("valence_b200: no CUDA device available; this engine has no CPU path");
    if (device < 0 || device >= ndev) throw std::runtime_error("valence_b200: CUDA device index out of range");
    I.device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    I.nsm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&I.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&I.ev0)); CK(cudaEventCreate(&I.ev1)); CK(cudaEventCreate(&I.ev2)); CK(cudaEventCreate(&I.ev3));
    std::vector<double> tab((size_t)BOYS_ROWS * BOYS_COLS);
    boys_make_table(tab.data());
    I.boys.upload(tab, I.st);
    std::vector<double> tabs(BOYS_S_SIZE, 0.0);
    boys_make_table_small(tabs.data());
    I.boys_small.upload(tabs, I.st);
    I.xyz_angs = in.coords;
    I.coeff_sc = in.coeff_sc;
    if (const char* t = std::getenv("VB_PRIM_TAU")) I.tau = std::atof(t);
    reset_orbitals();
    I.accum.alloc(2 + CNT_N);     // [E2 (high part), counters, E2 (low part)]
    CK(cudaStreamSynchronize(I.st));
}

Engine::~Engine()
{
    if (!impl_) return;
    Impl& I = *impl_;
    if (!I.shard_file.empty()) std::remove(I.shard_file.c_str());
    cudaSetDevice(I.device);
    if (I.st) cudaStreamSynchronize(I.st);
    for (cudaEvent_t ev : {I.ev0, I.ev1, I.ev2, I.ev3}) if (ev) cudaEventDestroy(ev);
    if (I.st) cudaStreamDestroy(I.st);
}

void Engine::set_coords_angstrom(const double* x) { impl_->xyz_angs.assign(x, x + 3 * in_.natom); }

void Engine::reset_orbitals()
{
    impl_->coeff.clear();
    for (const OrbitalDef& o : in_.orbitals) impl_->coeff.push_back(o.coeff);
}

double* Engine::accum_device() const { return impl_->accum.p; }
long long Engine::debug_tile_energies(double* out, long long cap) const
{
    long long n = (long long)impl_->tileE.n;
    if (out && cap >= n) {
        cudaMemcpy(out, impl_->tileE.p, n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    return n;
}
int Engine::accum_len() const { return 2 + CNT_N; }
void* Engine::stream() const { return (void*)impl_->st; }

namespace {

// CSR of (AO index, weight * angn) per orbital, for the one-electron kernels
void build_csr(const Basis& bas, const std::vector<ExpOrb>& orbs, std::vector<int>* ptr, std::vector<int>* ao, std::vector<double>* c)
{
    ptr->assign(1, 0); ao->clear(); c->clear();
    for (const ExpOrb& o : orbs) {
        for (const OrbShell& s : o.sh) {
            const GShell& g = bas.shells[s.gshell];
            for (int k = 0; k < ncart(g.l); ++k)
                if (s.c[k] != 0.0) { ao->push_back(g.ao_off + k); c->push_back(s.c[k] * bas.angn[coff(g.l) + k]); }
        }
        ptr->push_back((int)ao->size());
    }
}

}  // namespace

// Everything that depends only on geometry and orbital weights (valence.F90:71-145): basis,
// nuclear repulsion, AO one-electron matrices, normalised orbitals and their expansions.
// subject >= 0 appends one single-AO orbital per expansion term of orbital `subject`
// (the dummy orbitals of first_order_opt, valence.F90:619-664), ids norbs, norbs+1, ...
void Engine::Impl::prepare(const Input& in, int subject)
{
    CK(cudaSetDevice(device));
    const int norbs = in.norbs(), nval = norbs - in.ndf;
    std::vector<double> xyz(3 * in.natom);
    for (int i = 0; i < 3 * in.natom; ++i) xyz[i] = xyz_angs[i] * ANGS2BOHR;
    bas = build_basis(in, xyz);
    enuc = nuclear_repulsion(in, xyz);
    dtol = std::pow(10.0, -in.ntol_d);
    itol = std::pow(10.0, -in.ntol_i);
    const int nao = bas.nao, nshell = (int)bas.shells.size();
    {
        std::vector<DevShell> ds(nshell);
        for (int s = 0; s < nshell; ++s) {
            const GShell& g = bas.shells[s];
            ds[s] = {g.r[0], g.r[1], g.r[2], g.l, g.nprim, g.prim_off, g.ao_off};
        }
        std::vector<double> nuc(4 * in.natom);
        for (int a = 0; a < in.natom; ++a) {
            for (int d = 0; d < 3; ++d) nuc[4 * a + d] = xyz[3 * a + d];
            nuc[4 * a + 3] = in.types[in.atom_t[a] - 1].charge;
        }
        shells.upload(ds, st); exps.upload(bas.exps, st); coefs.upload(bas.coefs, st); this->nuc.upload(nuc, st);
    }
    S.alloc((size_t)nao * nao); H.alloc((size_t)nao * nao);
    {
        long long npair = (long long)nshell * (nshell + 1) / 2;
        const int wpb = 4;   // warps (= shell pairs) per block
        // with a communicator every rank takes the shell pairs k = rank (mod N) and the matrices are summed (NVLink)
        const int cr = crank(), cn = csize();
        if (cn > 1) { S.zero(st); H.zero(st); }
        const long long mine = (npair - cr + cn - 1) / cn;
        if (mine > 0) k_ao_1e<<<(unsigned)((mine + wpb - 1) / wpb), 32 * wpb, 0, st>>>(shells.p, nshell, exps.p, coefs.p, this->nuc.p, in.natom, boys.p, nao, 1e-30, S.p, H.p, cr, cn);
        CK(cudaGetLastError());
        launches++;
        if (cn > 1) { comm->allreduce_sum(S.p, (size_t)nao * nao, st); comm->allreduce_sum(H.p, (size_t)nao * nao, st); }
    }
    // normalise DBFs, then orbitals (valence.F90:144-145, normal :2157-2182); the raw weights stay untouched
    cnorm = coeff;
    auto self_overlaps = [&](int lo, int hi) {
        if (hi <= lo) return;
        std::vector<ExpOrb> ex;
        for (int o = lo; o < hi; ++o) ex.push_back(expand_orbital(in, bas, cnorm, o));
        std::vector<int> ptr, ao;
        std::vector<double> c;
        build_csr(bas, ex, &ptr, &ao, &c);
        std::vector<int2> prs;
        for (int o = 0; o < hi - lo; ++o) prs.push_back(make_int2(o, o));
        optr.upload(ptr, st); oao.upload(ao, st); oc.upload(c, st); opairs.upload(prs, st);
        one_e.alloc(prs.size());
        k_orb_1e<<<(unsigned)((prs.size() + 127) / 128), 128, 0, st>>>(optr.p, oao.p, oc.p, opairs.p, (int)prs.size(), nao, S.p, nullptr, one_e.p, nullptr);
        CK(cudaGetLastError());
        launches++;
        std::vector<double> sv;
        one_e.download(sv, st);
        for (int o = lo; o < hi; ++o) {
            double f = std::pow(sv[o - lo], -0.5);
            for (double& w : cnorm[o]) w = w * f;
        }
    };
    self_overlaps(nval, norbs);
    self_overlaps(0, nval);
    // orbital expansions: full (1e) and weight-screened (2e, valence.F90:3296-3348)
    orbs1e.assign(norbs, ExpOrb());
    for (int o = 0; o < norbs; ++o) orbs1e[o] = expand_orbital(in, bas, cnorm, o);
    if (subject >= 0) {
        const OrbitalDef& od = in.orbitals[subject];
        for (size_t ib = 0; ib < od.xp.size(); ++ib) {
            ExpOrb d;
            int gs = 0, cmp = 0;
            if (od.xp[ib] >= 1 && obs_position(in, bas, subject, od.xp[ib], &gs, &cmp)) {
                OrbShell os;
                os.gshell = gs;
                std::memset(os.c, 0, sizeof os.c);
                os.c[cmp] = 1.0;
                d.sh.push_back(os);
            }
            orbs1e.push_back(d);   // a DBF term leaves a null orbital here (valence.F90:616-617)
        }
    }
    orbs2e.assign(orbs1e.size(), ExpOrb());
    for (size_t o = 0; o < orbs1e.size(); ++o)
        for (const OrbShell& s : orbs1e[o].sh) {
            double sum = 0.0;
            for (int k = 0; k < ncart(bas.shells[s.gshell].l); ++k) sum = sum + s.c[k] * s.c[k];
            if (sum > dtol) orbs2e[o].sh.push_back(s);
        }
    {
        std::vector<int> ptr, ao;
        std::vector<double> c;
        build_csr(bas, orbs1e, &ptr, &ao, &c);
        optr.upload(ptr, st); oao.upload(ao, st); oc.upload(c, st);
    }
}

// Entry-level overlap / core-Hamiltonian matrices and the cofactor densities of one bra/ket list pair
// (wfndet, the 1e loop and the density set-up of vsvb_energy): leaves Se, He, Pa/Pb or cof on the device,
// e1 and wfnorm in the object.
void Engine::Impl::cofactor_stage(const Input& in, const Wavefunction& wf, bool diag_only, EnergyResult* out, bool* fast_out, double* c0_out, int* ndp_out)
{
    const int nso = wf.nso, nelec = in.nelec(), nao = bas.nao;
    // ---- entry-level overlap and core-Hamiltonian matrices (wfndet :1440-1480, 1e loop :1072) ----
    {
        std::vector<int2> prs((size_t)nso * nso);
        for (int s = 0; s < nso; ++s)
            for (int t = 0; t < nso; ++t) prs[(size_t)s * nso + t] = make_int2(wf.bra[wf.slot(s, 0)], wf.ket[wf.slot(t, 0)]);
        opairs.upload(prs, st);
        Se.alloc(prs.size()); He.alloc(prs.size());
        k_orb_1e<<<(unsigned)((prs.size() + 127) / 128), 128, 0, st>>>(optr.p, oao.p, oc.p, opairs.p, (int)prs.size(), nao, S.p, H.p, Se.p, He.p);
        CK(cudaGetLastError());
        launches++;
    }
    // ---- cofactor densities ------------------------------------------------------------------------
    const int na = in.nalpha(), nb = in.nbeta();
    // large single determinant (closed/open shell, also the substituted lists of first_order_opt): inverse form on the GPU
    int fast_min = 64;
    if (const char* e = std::getenv("VB_FAST_MIN_N")) fast_min = std::atoi(e);
    bool fast = in.npair == 0 && std::max(na, nb) > fast_min;
    double c0 = 1.0;
    int ndp = 0;
    if (!diag_only) {
        bool singular = false, same = false;
        if (fast) {
            // spin of an electron slot is its position in the list (set_up_unpaired_docc, valence.F90:2461-2480):
            // unpaired slots and the first slot of every DOCC pair are alpha, the second slots beta
            std::vector<int> entry_of_slot(wf.bra.size());
            for (int s = 0; s < nso; ++s)
                for (int k = 0; k < wf.nslots(s); ++k) entry_of_slot[wf.slot(s, k)] = s;
            std::vector<int> ea, eb, posa(nso, -1), posb(nso, -1);
            for (int i = 0; i < in.nunpd; ++i) { const int s = entry_of_slot[i]; posa[s] = (int)ea.size(); ea.push_back(s); }
            for (int d = 0; d < in.ndocc; ++d) {
                const int sa = entry_of_slot[in.nunpd + 2 * d], sb = entry_of_slot[in.nunpd + 2 * d + 1];
                posa[sa] = (int)ea.size(); ea.push_back(sa);
                posb[sb] = (int)eb.size(); eb.push_back(sb);
            }
            ea_bra.upload(ea, st); eb_bra.upload(eb, st); posa_bra.upload(posa, st); posb_bra.upload(posb, st);
            Ma.alloc((size_t)na * na + 1); Mb.alloc((size_t)nb * nb + 1);
            // one inverse when both spin blocks hold the same entries (closed shell)
            same = ea == eb;
            if (na) { k_gather_block<<<(na * na + 255) / 256, 256, 0, st>>>(Se.p, nso, ea_bra.p, ea_bra.p, na, Ma.p); launches++; }
            if (nb && !same) { k_gather_block<<<(nb * nb + 255) / 256, 256, 0, st>>>(Se.p, nso, eb_bra.p, eb_bra.p, nb, Mb.p); launches++; }
            CK(cudaGetLastError());
            gjout.alloc(4);
            auto invert = [&](DBuf<double>& M, DBuf<double>& Minv, int n, double* res) {
                Minv.alloc((size_t)n * n + 1);
                if (n == 0) { const double one[2] = {1.0, 1.0}; CK(cudaMemcpyAsync(res, one, sizeof one, cudaMemcpyHostToDevice, st)); return; }
                gj_ws.alloc(2 * (size_t)n); piv.alloc(2 * (size_t)n + 2);
                int grid = std::max(1, std::min(nsm, n / 4));
                double* Ap = M.p; double* Ip = Minv.p; double* wsp = gj_ws.p; int* iwp = piv.p; int nn = n;
                void* args[] = {&Ap, &nn, &Ip, &wsp, &iwp, &res};
                CK(cudaLaunchCooperativeKernel((void*)k_gj_inverse_grid, dim3(grid), dim3(1024), args, 0, st));
                launches++;
            };
            // one refinement step with a double-double residual (k_inv_residual_dd): M is gathered again (the
            // elimination destroyed it), R = I - M X, X <- X + X R; the buffers swap roles
            auto refine = [&](DBuf<double>& M, DBuf<double>& Minv, const DBuf<int>& ent, int n) {
                if (n == 0) return;
                if (const char* e = std::getenv("VB_INV_REFINE")) if (std::atoi(e) == 0) return;
                k_gather_block<<<(n * n + 255) / 256, 256, 0, st>>>(Se.p, nso, ent.p, ent.p, n, M.p);
                gj_res.alloc((size_t)n * n);
                const dim3 g((n + RF_T - 1) / RF_T, (n + RF_T - 1) / RF_T), b(RF_T, RF_T);
                k_inv_residual_dd<<<g, b, 0, st>>>(M.p, Minv.p, n, gj_res.p);
                k_inv_update<<<g, b, 0, st>>>(Minv.p, gj_res.p, n, M.p);
                CK(cudaGetLastError());
                launches += 3;
                std::swap(M.p, Minv.p); std::swap(M.cap, Minv.cap); std::swap(M.n, Minv.n);
            };
            invert(Ma, Mai, na, gjout.p);
            refine(Ma, Mai, ea_bra, na);
            if (same) CK(cudaMemcpyAsync(gjout.p + 2, gjout.p, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
            else { invert(Mb, Mbi, nb, gjout.p + 2); refine(Mb, Mbi, eb_bra, nb); }
            std::vector<double> g;
            gjout.download(g, st);
            out->min_pivot_ratio = std::min(g[1], g[3]);
            if (!(g[0] != 0.0) || !(g[2] != 0.0) || out->min_pivot_ratio < 1e-13) {
                // symmetric lists: linearly dependent orbitals.  Substituted lists (first_order_opt) can be singular
                // legitimately (a basis function orthogonal to the space it replaces): exact null-space treatment on
                // the host while that is affordable
                if (wf.sym || std::max(na, nb) > 320)
                    throw std::runtime_error("valence_b200: singular spin-block overlap matrix (linearly dependent orbitals)");
                singular = true;
            }
            c0 = g[0] * g[2];
        }
        if (fast && !singular) {
            Pa.alloc((size_t)nso * nso); Pb.alloc((size_t)nso * nso);
            k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(Mai.p, na, posa_bra.p, posa_bra.p, nso, Pa.p);
            k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(same ? Mai.p : Mbi.p, nb, posb_bra.p, posb_bra.p, nso, Pb.p);
            pb_same = same && na == nb;      // same entry lists, hence the same positions
            one_e.alloc(4);
            k_one_electron_energy<<<1, 1024, 0, st>>>(Se.p, He.p, Pa.p, Pb.p, nso * nso, one_e.p);
            CK(cudaGetLastError());
            launches += 3;
            std::vector<double> oe;
            one_e.download(oe, st);
            e1 = c0 * oe[0];
            wfnorm = c0 * oe[1] / (double)nelec;    // valence.F90:1106
            e1u[0] = oe[0]; e1u[1] = oe[2]; wnu[0] = oe[1]; wnu[1] = oe[3]; c0_keep = c0; nelec_keep = nelec;
        } else {
            fast = false;
            c0 = 1.0;
            c0_keep = 0.0;
            if (in.npair > 0 && std::max(na, nb) > fast_min) {
                // spin-coupled pairs in a large wavefunction: every determinant pair's blocks inverted on the GPU
                sc_cofactors_gpu(in, wf, out, &ndp);
                *fast_out = false; *c0_out = 1.0; *ndp_out = ndp;
                return;
            }
            // small blocks / several determinant pairs / possibly singular blocks: factorise on the host
            // (O(n^3) once per determinant pair, n <= 64), contract on the GPU
            std::vector<double> hSe, hHe;
            Se.download(hSe, st); He.download(hHe, st);
            CofactorSet cs;
            Input in2 = in;
            if (!coeff_sc.empty()) in2.coeff_sc = coeff_sc;
            build_cofactors(in2, wf, hSe, &cs, only_isc, only_jsc);
            one_electron_from_cofactors(cs, hSe, hHe, nelec, &e1, &wfnorm);
            out->min_pivot_ratio = cs.min_sigma_ratio;
            cof.upload(cs.data, st);
            ndp = cs.ndp;
        }
    }
    *fast_out = fast; *c0_out = c0; *ndp_out = ndp;
}


// Cofactor data of a spin-coupled wavefunction with large determinants, built on the GPU.  Same enumeration and the same packed
// layout as build_cofactors (vb_cofactor.cpp; density_sc / dbra / dket of the reference, valence.F90:1576-1588, 1648-1870): for every
// (bra coupling, ket coupling, bra spin assignment, ket spin assignment) the alpha and beta overlap blocks are gathered from the
// entry-level overlaps, inverted by the cooperative Gauss-Jordan kernel (+ one double-double refinement step) and scattered to
// entry level straight into the pair's slot of `cof`; the one-electron sums of the pair are reduced on the device.  Blocks must
// be non-singular (the host path treats null spaces exactly, for blocks up to 64).  Leaves e1, wfnorm in the object.
void Engine::Impl::sc_cofactors_gpu(const Input& in, const Wavefunction& wf, EnergyResult* out, int* ndp_out)
{
    const int nso = wf.nso, nelec = (int)wf.bra.size();
    std::vector<int> entry_of_slot(nelec);
    for (int s = 0; s < nso; ++s)
        for (int k = 0; k < wf.nslots(s); ++k) entry_of_slot[wf.slot(s, k)] = s;
    const int npair = in.npair, nunpd = in.nunpd, ndocc = in.ndocc;
    const int nsc = std::max(1, in.nspinc);
    const std::vector<double>& csc = coeff_sc.empty() ? in.coeff_sc : coeff_sc;
    const size_t stride = cof_stride(nso);
    const long long nmask = 1LL << npair;
    const long long ncoup = only_isc >= 0 ? 1 : (long long)nsc * nsc;
    const long long ndp_tot = ncoup * nmask * nmask;
    if (ndp_tot > 65536 || (double)ndp_tot * stride * 8.0 > 64e9) throw std::runtime_error("valence_b200: too many determinant pairs for the GPU cofactor path");
    cof.alloc((size_t)ndp_tot * stride);
    cof.zero(st);
    std::vector<int> a_fixed, b_fixed;     // set_up_unpaired_docc, valence.F90:2461-2480 (0-based slots)
    for (int i = 0; i < nunpd; ++i) a_fixed.push_back(2 * npair + i);
    for (int d = 0; d < ndocc; ++d) { a_fixed.push_back(2 * npair + nunpd + 2 * d); b_fixed.push_back(2 * npair + nunpd + 2 * d + 1); }
    const int na = npair + (int)a_fixed.size(), nb = npair + (int)b_fixed.size();
    Ma.alloc((size_t)na * na + 1); Mai.alloc((size_t)na * na + 1); Mb.alloc((size_t)nb * nb + 1); Mbi.alloc((size_t)nb * nb + 1);
    gj_res.alloc((size_t)std::max(na, nb) * std::max(na, nb));
    gj_ws.alloc(2 * (size_t)std::max(na, nb)); piv.alloc(2 * (size_t)std::max(na, nb) + 2);
    gjout.alloc(4); one_e.alloc(4);
    sc_idx.alloc(2 * (size_t)(na + nb) + 4 * (size_t)nso);
    long double e1s = 0.0L, wns = 0.0L;
    double minratio = 1.0;
    int d = 0;
    auto invert = [&](DBuf<double>& M, DBuf<double>& Minv, const int* rows, const int* cols, int n, double* res) {
        if (n == 0) { const double one[2] = {1.0, 1.0}; CK(cudaMemcpyAsync(res, one, sizeof one, cudaMemcpyHostToDevice, st)); return; }
        k_gather_block<<<(n * n + 255) / 256, 256, 0, st>>>(Se.p, nso, rows, cols, n, M.p);
        int grid = std::max(1, std::min(nsm, n / 4));
        double* Ap = M.p; double* Ip = Minv.p; double* wsp = gj_ws.p; int* iwp = piv.p; int nn = n;
        void* args[] = {&Ap, &nn, &Ip, &wsp, &iwp, &res};
        CK(cudaLaunchCooperativeKernel((void*)k_gj_inverse_grid, dim3(grid), dim3(1024), args, 0, st));
        k_gather_block<<<(n * n + 255) / 256, 256, 0, st>>>(Se.p, nso, rows, cols, n, M.p);
        const dim3 g((n + RF_T - 1) / RF_T, (n + RF_T - 1) / RF_T), b(RF_T, RF_T);
        k_inv_residual_dd<<<g, b, 0, st>>>(M.p, Minv.p, n, gj_res.p);
        k_inv_update<<<g, b, 0, st>>>(Minv.p, gj_res.p, n, M.p);       // refined inverse lands in M
        CK(cudaGetLastError());
        launches += 5;
    };
    for (int isc = 0; isc < nsc; ++isc)
        for (int jsc = 0; jsc < nsc; ++jsc) {
            if (only_isc >= 0 && (isc != only_isc || jsc != only_jsc)) continue;
            for (long long bm = 0; bm < nmask; ++bm)
                for (long long km = 0; km < nmask; ++km, ++d) {
                    std::vector<int> abra, bbra, aket, bket;
                    for (int k = 0; k < npair; ++k) {
                        int b1 = in.pair(isc, k, 0) - 1, b2 = in.pair(isc, k, 1) - 1;
                        int k1 = in.pair(jsc, k, 0) - 1, k2 = in.pair(jsc, k, 1) - 1;
                        if ((bm >> k) & 1) std::swap(b1, b2);
                        if ((km >> k) & 1) std::swap(k1, k2);
                        abra.push_back(b1); bbra.push_back(b2); aket.push_back(k1); bket.push_back(k2);
                    }
                    abra.insert(abra.end(), a_fixed.begin(), a_fixed.end()); aket.insert(aket.end(), a_fixed.begin(), a_fixed.end());
                    bbra.insert(bbra.end(), b_fixed.begin(), b_fixed.end()); bket.insert(bket.end(), b_fixed.begin(), b_fixed.end());
                    // one upload: [ea_bra na][ea_ket na][eb_bra nb][eb_ket nb][posa_bra nso][posa_ket nso][posb_bra nso][posb_ket nso]
                    std::vector<int> h(2 * (size_t)(na + nb) + 4 * (size_t)nso, -1);
                    int* ea_b = h.data(); int* ea_k = ea_b + na; int* eb_b = ea_k + na; int* eb_k = eb_b + nb;
                    int* pa_b = eb_k + nb; int* pa_k = pa_b + nso; int* pb_b = pa_k + nso; int* pb_k = pb_b + nso;
                    for (int r = 0; r < na; ++r) { ea_b[r] = entry_of_slot[abra[r]]; ea_k[r] = entry_of_slot[aket[r]]; pa_b[ea_b[r]] = r; pa_k[ea_k[r]] = r; }
                    for (int r = 0; r < nb; ++r) { eb_b[r] = entry_of_slot[bbra[r]]; eb_k[r] = entry_of_slot[bket[r]]; pb_b[eb_b[r]] = r; pb_k[eb_k[r]] = r; }
                    CK(cudaMemcpyAsync(sc_idx.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st));
                    g_h2d_bytes += (long long)(h.size() * sizeof(int));
                    const int* D0 = sc_idx.p;
                    invert(Ma, Mai, D0, D0 + na, na, gjout.p);
                    invert(Mb, Mbi, D0 + 2 * na, D0 + 2 * na + nb, nb, gjout.p + 2);
                    double* slot = cof.p + (size_t)d * stride;
                    double* Ga = slot + COF_HEADER;
                    double* Gb = Ga + (size_t)nso * nso;
                    const int* P0 = D0 + 2 * (na + nb);
                    if (na) k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(Ma.p, na, P0, P0 + nso, nso, Ga);
                    if (nb) k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(Mb.p, nb, P0 + 2 * nso, P0 + 3 * nso, nso, Gb);
                    k_one_electron_energy<<<1, 1024, 0, st>>>(Se.p, He.p, Ga, Gb, nso * nso, one_e.p);
                    CK(cudaGetLastError());
                    launches += 3;
                    double g4[4], oe[4];
                    CK(cudaMemcpyAsync(g4, gjout.p, sizeof g4, cudaMemcpyDeviceToHost, st));
                    CK(cudaMemcpyAsync(oe, one_e.p, sizeof oe, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));      // also keeps the host index vector alive until it is copied
                    g_d2h_bytes += 64;
                    const double ra = na ? g4[1] : 1.0, rb = nb ? g4[3] : 1.0;
                    minratio = std::min(minratio, std::min(ra, rb));
                    if (!(g4[0] != 0.0) || !(g4[2] != 0.0) || std::min(ra, rb) < 1e-13)
                        throw std::runtime_error("valence_b200: singular spin block in a spin-coupled wavefunction too large for the host factorisation");
                    const double w = only_isc >= 0 ? 1.0 : ((npair > 0 && in.nspinc > 0) ? csc[isc] * csc[jsc] : 1.0);
                    const double hdr[COF_HEADER] = {w, g4[0], 1.0, 0.0, 0.0, 0.0, g4[2], 1.0, 0.0, 0.0, 0.0, 0.0};
                    CK(cudaMemcpyAsync(slot, hdr, sizeof hdr, cudaMemcpyHostToDevice, st));
                    CK(cudaStreamSynchronize(st));
                    const long double dd0 = (long double)w * g4[0] * g4[2];
                    e1s += dd0 * ((long double)oe[0] + oe[2]);
                    wns += dd0 * ((long double)oe[1] + oe[3]);
                }
        }
    e1 = (double)e1s;
    wfnorm = (double)(wns / (long double)nelec);
    out->min_pivot_ratio = minratio;
    *ndp_out = d;
}

void Engine::Impl::gpu_inverse(int nso, const std::vector<int>& ent, DBuf<int>& ent_dev, DBuf<double>& M, DBuf<double>& Minv, double* res_dev)
{
    const int n = (int)ent.size();
    ent_dev.upload(ent, st);
    M.alloc((size_t)n * n + 1); Minv.alloc((size_t)n * n + 1);
    if (n == 0) { const double one[2] = {1.0, 1.0}; CK(cudaMemcpyAsync(res_dev, one, sizeof one, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); return; }
    k_gather_block<<<(n * n + 255) / 256, 256, 0, st>>>(Se.p, nso, ent_dev.p, ent_dev.p, n, M.p);
    gj_ws.alloc(2 * (size_t)n); piv.alloc(2 * (size_t)n + 2);
    int grid = std::max(1, std::min(nsm, n / 4));
    double* Ap = M.p; double* Ip = Minv.p; double* wsp = gj_ws.p; int* iwp = piv.p; int nn = n;
    void* args[] = {&Ap, &nn, &Ip, &wsp, &iwp, &res_dev};
    CK(cudaLaunchCooperativeKernel((void*)k_gj_inverse_grid, dim3(grid), dim3(1024), args, 0, st));
    k_gather_block<<<(n * n + 255) / 256, 256, 0, st>>>(Se.p, nso, ent_dev.p, ent_dev.p, n, M.p);
    gj_res.alloc((size_t)n * n);
    const dim3 g((n + RF_T - 1) / RF_T, (n + RF_T - 1) / RF_T), b(RF_T, RF_T);
    k_inv_residual_dd<<<g, b, 0, st>>>(M.p, Minv.p, n, gj_res.p);
    k_inv_update<<<g, b, 0, st>>>(Minv.p, gj_res.p, n, M.p);
    CK(cudaGetLastError());
    launches += 5;
    std::swap(M.p, Minv.p); std::swap(M.cap, Minv.cap); std::swap(M.n, Minv.n);
}

// One vsvb_energy evaluation (valence.F90:1010-1434) for the bra/ket lists in wf.
//   sch_in  : Schwarz table to screen with (first_order_opt reuses the unsubstituted one,
//             valence.F90:667 vs 674-704); null -> run the diagonal pass (schwarz_ints)
//   diag_only: stop after the Schwarz table (returned in sch_out)
void Engine::Impl::evaluate(const Input& in, const Wavefunction& wf, const std::vector<double>* sch_in, bool diag_only,
                            int rank, int nranks, EnergyResult* out, std::vector<double>* sch_out)
{
    const int nso = wf.nso;
    double t1 = now_ms(), t_cof = 0.0;
    bool fast = false;
    double c0 = 1.0;
    int ndp = 0;

    // ---- pair groups, shell-pair tables, folded densities ---------------------------------------------
    // (host cores) built WHILE the GPU forms the entry-level matrices, inverts the spin blocks and reduces the one-electron
    // sums (cofactor_stage below): the two do not depend on each other
    // magnitude cuts: the Schwarz (diagonal) pass must resolve (st|st) down to itol^2, the energy pass
    // needs the integrals to ~itol; never looser than the configured tau
    const double tau_diag = std::min(tau, 0.01 * itol * itol), tau_energy = std::min(tau, 0.01 * itol);
    int bas_lmax = 0;
    for (const GShell& gs : bas.shells) bas_lmax = std::max(bas_lmax, gs.l);
    const bool gen = bas_lmax >= 2;   // d shells: shell-pair kernel with the loop-based recurrence (k_tile<true>)
    TileSetup& ts = ts_keep;    // storage reused across calls (no zero-fill / page faults of ~0.6 GB per energy)
    TileOpts topts;
    if (shard_pending) {       // every rank published its share of the pair groups (shard_tables); merge them
        topts.shard_mode = 2; topts.shard_rank = shard_rank; topts.shard_nranks = shard_nranks; topts.shard_prefix = shard_prefix;
        shard_pending = false;
    }
    const bool gather = use_gather && comm && comm->nranks() > 1 && !gen && !shard_pending;
    if (gather) { topts.shard_mode = 3; topts.shard_rank = comm->rank(); topts.shard_nranks = comm->nranks(); }
    {
        std::exception_ptr bt_err;
        std::thread bt([&] { try { build_tiles(in, bas, wf, orbs2e, tau_diag, !gen, &ts, topts); } catch (...) { bt_err = std::current_exception(); } });
        try { cofactor_stage(in, wf, diag_only, out, &fast, &c0, &ndp); } catch (...) { bt.join(); throw; }
        t_cof = now_ms() - t1;
        bt.join();
        if (bt_err) std::rethrow_exception(bt_err);
    }
    const double t2 = t1 + t_cof;      // end of the device-side cofactor stage; what is left of the table build counts as host set-up
    if (gather) {
        // Every rank has built the pair groups it owns (blocks of 16, round-robin) with local offsets.  The shares meet on
        // the DEVICE: sizes by one small all-reduce, then each table is all-gathered over NVLink into a layout with one
        // equal-sized slot per rank (holes are pair groups with np = 0, which no tile ever references).  Host build and H2D
        // traffic per rank are 1/N of the single-GPU ones; the descriptors come back to the host for the tile list.
        const int N = comm->nranks(), r = comm->rank();
        constexpr int NX = 10;
        std::vector<double> x((size_t)N * NX, 0.0);
        const double mine[NX] = {(double)ts.pgs.size(), (double)(ts.pg_pairs.size() / 2), (double)ts.sps.size(), (double)ts.pps.size(), (double)ts.dmat.size(),
                                 (double)ts.max_ne, (double)ts.max_np, (double)ts.max_npp, (double)ts.max_nsp, (double)ts.max_ks};
        for (int k = 0; k < NX; ++k) x[(size_t)r * NX + k] = mine[k];
        one_e.alloc(x.size());
        CK(cudaMemcpyAsync(one_e.p, x.data(), x.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        comm->allreduce_sum(one_e.p, x.size(), st);
        CK(cudaMemcpyAsync(x.data(), one_e.p, x.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        size_t cap[5] = {0, 0, 0, 0, 0};
        for (int q = 0; q < N; ++q) {
            for (int k = 0; k < 5; ++k) cap[k] = std::max(cap[k], (size_t)x[(size_t)q * NX + k]);
            ts.max_ne = std::max(ts.max_ne, (int)x[(size_t)q * NX + 5]); ts.max_np = std::max(ts.max_np, (int)x[(size_t)q * NX + 6]);
            ts.max_npp = std::max(ts.max_npp, (int)x[(size_t)q * NX + 7]); ts.max_nsp = std::max(ts.max_nsp, (int)x[(size_t)q * NX + 8]);
            ts.max_ks = std::max(ts.max_ks, (int)x[(size_t)q * NX + 9]);
        }
        for (int k = 0; k < 5; ++k) cap[k] = (cap[k] + 15) & ~(size_t)15;
        const size_t cpg = cap[0], cpr = cap[1], csp = cap[2], cpp = cap[3], cd = cap[4];
        if ((cpp + 1) * (size_t)N > 2000000000ull || (csp + 1) * (size_t)N > 2000000000ull) throw std::runtime_error("valence_b200: tables too large for 32-bit offsets");
        for (PGDesc& pg : ts.pgs) {
            pg.pair_beg += (int)(r * cpr); pg.d_off += (long long)(r * cd);
            for (int t = 0; t <= NPTYPE; ++t) { pg.pp_beg[t] += (int)(r * cpp); pg.sp_beg[t] += (int)(r * csp); }
        }
        for (SPRec& sr : ts.sps) sr.pp_beg += (int)(r * cpp);
        pgs.alloc(N * cpg); pg_pairs.alloc(2 * N * cpr); sps.alloc(N * csp); pps.alloc(N * cpp); pps_flat.alloc(N * cpp); dmat.alloc(N * cd);
        pgs.zero(st);
        pgs.upload_at(r * cpg, ts.pgs, st); pg_pairs.upload_at(2 * r * cpr, ts.pg_pairs, st); sps.upload_at(r * csp, ts.sps, st);
        pps.upload_at(r * cpp, ts.pps, st); pps_flat.upload_at(r * cpp, ts.pps_flat, st); dmat.upload_at(r * cd, ts.dmat, st);
        comm->allgather_bytes(pgs.p, cpg * sizeof(PGDesc), st); comm->allgather_bytes(pg_pairs.p, 2 * cpr * sizeof(int), st);
        comm->allgather_bytes(sps.p, csp * sizeof(SPRec), st); comm->allgather_bytes(pps.p, cpp * sizeof(PrimPair), st);
        comm->allgather_bytes(pps_flat.p, cpp * sizeof(PrimPair), st); comm->allgather_bytes(dmat.p, cd * sizeof(double), st);
        pgs.download(ts.pgs, st);
        pg_pairs.download(ts.pg_pairs, st);
    }
    const double t_bt = now_ms();
    const int npg = (int)ts.pgs.size();
    if (ts.max_np > 32) throw std::runtime_error("valence_b200: pair group too large");
    const int dq_cap = std::max(1, ts.max_ne * ts.max_np);
    const int hs_ld = std::max(1, ts.max_np) | 1;
    const int dq_cap2 = (dq_cap + 1) & ~1, hs_cap = (ts.max_ne * hs_ld + 1) & ~1, sp_cap = std::max(1, ts.max_nsp);
    const int g_cap = (ts.max_np * ts.max_np + 1) & ~1;
    int pp_cap = std::max(1, ts.max_npp);
    size_t smem;
    int boys_cap = 0;
    double* A_gred = nullptr;
    if (gen) {
        smem = ((size_t)dq_cap2 + hs_cap) * sizeof(double) + (size_t)sp_cap * sizeof(SPRec);
        if (smem + 2 * (size_t)pp_cap * sizeof(PrimPair) <= 225 * 1024) smem += 2 * (size_t)pp_cap * sizeof(PrimPair);
        else pp_cap = 0;   // primitive tables stay in global memory
    } else {
        constexpr int nw = PT_MAX_WARPS;
        smem = ((size_t)dq_cap2 + (size_t)PT_MAXQ * g_cap + (size_t)nw * PT_SCRATCH) * sizeof(double) + (size_t)sp_cap * sizeof(SPRec);
        if (smem + (size_t)pp_cap * sizeof(PrimPair) <= PT_SMEM_MAX) smem += (size_t)pp_cap * sizeof(PrimPair);
        else pp_cap = 0;
        if (smem + BOYS_S_SIZE * sizeof(double) <= PT_SMEM_MAX) { boys_cap = BOYS_S_SIZE; smem += BOYS_S_SIZE * sizeof(double); }
    }
    if (smem > PT_SMEM_MAX) throw std::runtime_error("valence_b200: orbital basis set too large for one pair-group tile");
    std::vector<int> nshb(nso), nshk(nso);
    for (int s = 0; s < nso; ++s) {
        nshb[s] = (int)orbs2e[wf.bra[wf.slot(s, 0)]].sh.size();
        nshk[s] = (int)orbs2e[wf.ket[wf.slot(s, 0)]].sh.size();
    }
    if (!gather) {
        pgs.upload(ts.pgs, st); pg_pairs.upload(ts.pg_pairs, st); sps.upload(ts.sps, st);
        pps.upload(ts.pps, st); pps_flat.upload(ts.pps_flat, st); dmat.upload(ts.dmat, st);
    }
    nsh_bra.upload(nshb, st); nsh_ket.upload(nshk, st);
    counter.alloc(1); counters.alloc(CNT_N); pq_counters.alloc(NPTYPE * NPTYPE);
    if (!gen) { gred.alloc((size_t)nsm * PT_MAXQ * g_cap); A_gred = gred.p; }
    int grid_cap = nsm;
    if (gen) gen_scratch.alloc((size_t)grid_cap * TILE_THREADS * GEN_PER_THREAD);      // one CTA per SM (3.3 GB of recurrence scratch on 148 SMs)
    if (gen) CK(cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else {
        CK(cudaFuncSetAttribute(k_ptile<PART_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_ptile<PART_HEAVY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_ptile<PART_LIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    TileArgs A;
    std::memset(&A, 0, sizeof A);
    A.gred = A_gred;
    A.pgs = pgs.p; A.pg_pairs = pg_pairs.p; A.sps = sps.p; A.pps = pps.p; A.pps_flat = pps_flat.p; A.tau = tau_diag;
    A.pq_counters = pq_counters.p; A.dmat = dmat.p;
    A.boys = boys.p; A.counter = counter.p; A.nso = nso; A.nnd = wf.nnd; A.sym = wf.sym ? 1 : 0; A.subject = wf.subject;
    A.dq_cap = dq_cap2; A.hs_cap = hs_cap; A.hs_ld = hs_ld; A.g_cap = g_cap; A.boys_small = boys_small.p; A.boys_cap = boys_cap; A.pp_cap = pp_cap; A.sp_cap = sp_cap; A.itol = itol; A.Pa = Pa.p; A.Pb = (fast && pb_same) ? Pa.p : Pb.p; A.c0 = fast ? c0 : 1.0; A.nsh_bra = nsh_bra.p; A.nsh_ket = nsh_ket.p;
    A.cof = cof.p; A.ndp = ndp; A.cof_stride = (long long)cof_stride(nso);
    A.counters = counters.p; A.gen_scratch = gen_scratch.p;
    A.debug = std::getenv("VB_DEBUG_ENTRIES") ? 1 : 0;
    auto launch = [&](int nwork, int part) {
        int grid = std::max(1, std::min(grid_cap, nwork));
        if (gen) {
            // d-shell inputs are small (one transition-metal atom: tens of tiles, each needing every AO quartet of the atom): one CTA
            // per tile leaves most SMs idle and the pass takes as long as its heaviest tile.  The first half transformation of
            // every tile is therefore split over `hs` CTAs (static slices of the (class, bra shell pair) units, shares summed in
            // slice order by the contraction launch: bitwise reproducible, so every rank still derives the same Schwarz table).
            int hs = std::max(1, std::min(64, (4 * nsm + nwork - 1) / std::max(1, nwork)));
            if (const char* e = std::getenv("VB_GEN_SPLIT")) hs = std::max(1, std::min(64, std::atoi(e)));
            // a slice takes every US-th unit and every KS-th batch of 32 ket primitives of it (the heaviest unit -- a dd bra
            // shell pair against every ket primitive -- is otherwise one warp's work for the whole launch)
            const int KS = std::min(4, hs), US = std::max(1, hs / KS);
            hs = US * KS;
            if (hs > 1) {
                hpart.alloc((size_t)nwork * hs * hs_cap);
                A.hsplit = hs; A.hksplit = KS; A.hpart = hpart.p;
                A.hphase = 1;
                k_tile<true><<<std::max(1, std::min(grid_cap, nwork * hs)), TILE_THREADS, smem, st>>>(A);
                CK(cudaMemsetAsync(counter.p, 0, sizeof(unsigned int), st));
                A.hphase = 2;
                k_tile<true><<<grid, TILE_THREADS, smem, st>>>(A);
                A.hsplit = 1; A.hksplit = 1; A.hphase = 0;
                launches++;
            } else {
                k_tile<true><<<grid, TILE_THREADS, smem, st>>>(A);
            }
        }
        else if (part == PART_HEAVY) k_ptile<PART_HEAVY><<<grid, pt_threads(PART_HEAVY), smem, st>>>(A);
        else if (part == PART_LIGHT) k_ptile<PART_LIGHT><<<grid, pt_threads(PART_LIGHT), smem, st>>>(A);
        else k_ptile<PART_ALL><<<grid, pt_threads(PART_ALL), smem, st>>>(A);
        CK(cudaGetLastError());
        launches++;
    };
    double t3 = now_ms();
    const bool dbg_time = std::getenv("VB_DEBUG_TIME") != nullptr;
    if (dbg_time) { CK(cudaStreamSynchronize(st)); std::printf("[time] density %.1f build_tiles %.1f upload+setup %.1f ms\n", t2 - t1, t_bt - t2, now_ms() - t_bt); }
    // ---- diagonal pass: (st|st) for every pair -> Schwarz table (schwarz_ints, :1489-1523) ------------
    std::vector<double> sch;
    if (sch_in) {
        sch = *sch_in;
    } else {
        std::vector<double> dg;
        std::vector<TilePair> dt;
        std::vector<WorkItem> di;
        for (int i = 0; i < npg; ++i)
            if (ts.pgs[i].np > 0) { di.push_back(WorkItem{(int)dt.size(), 1, 0, 0}); dt.push_back(TilePair{i, i}); }
        const int ndiag = (int)dt.size();
        tiles.upload(dt, st); items.upload(di, st);
        diag.alloc((size_t)nso * nso);
        diag.zero(st); counter.zero(st); counters.zero(st); pq_counters.zero(st);
        A.tiles = reinterpret_cast<const int2*>(tiles.p); A.ntiles = ndiag; A.items = reinterpret_cast<const int4*>(items.p); A.nitems = ndiag; A.gbuf = nullptr; A.gslot_base = 0; A.tile_first = 0; A.tile_stride = 1; A.mode = 0; A.diag = diag.p;
        CK(cudaEventRecord(ev0, st));
        launch(ndiag, PART_ALL);
        CK(cudaEventRecord(ev1, st));
        out->diag_launches++;
        diag.download(dg, st);
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        out->t_diag += ms;
        // reference table: schwarz(indx(i,j)) = sqrt((bra_i ket_j|bra_i ket_j)), i >= j, looked up symmetrically
        sch.resize((size_t)nso * nso);
        for (int s = 0; s < nso; ++s)
            for (int t = 0; t < nso; ++t) sch[(size_t)s * nso + t] = std::sqrt(dg[(size_t)std::max(s, t) * nso + std::min(s, t)]);
    }
    if (sch_out) *sch_out = sch;
    out->n_entries = nso; out->n_groups = (long long)ts.groups.size(); out->n_pairgroups = npg;
    out->t_density += t2 - t1;
    if (diag_only) { out->t_host_setup += t3 - t2; return; }
    for (PGDesc& pg : ts.pgs) {
        double m = 0.0;
        for (int p = 0; p < pg.np; ++p) {
            double v = sch[(size_t)ts.pg_pairs[2 * (pg.pair_beg + p)] * nso + ts.pg_pairs[2 * (pg.pair_beg + p) + 1]];
            if (v > m) m = v;   // NaN never wins
        }
        pg.smax = m;
    }
    if (std::getenv("VB_DEBUG_PAIRS")) {
        // how many orbital pairs of a pair group can ever pass the Schwarz screen (sch_st * max sch > itol)?
        double smx = 0.0;
        for (const PGDesc& pg : ts.pgs) smx = std::max(smx, pg.smax);
        long long hist[33] = {0}, histw[33] = {0}, tot = 0, live = 0;
        for (const PGDesc& pg : ts.pgs) {
            if (pg.np <= 0) continue;
            int n = 0;
            for (int p = 0; p < pg.np; ++p) {
                double v = sch[(size_t)ts.pg_pairs[2 * (pg.pair_beg + p)] * nso + ts.pg_pairs[2 * (pg.pair_beg + p) + 1]];
                if (v * smx > itol) ++n;
            }
            hist[std::min(32, n)]++; tot += pg.np; live += n;
            (void)histw;
        }
        std::printf("[pairs] %lld pair groups, %lld pairs, %lld can pass the Schwarz screen (max sch %.3g); live pairs per group:", (long long)ts.pgs.size(), tot, live, smx);
        for (int n = 0; n <= 32; ++n) if (hist[n]) std::printf(" %d:%lld", n, hist[n]);
        std::printf("\n");
    }
    // ---- tile list: pair-group pairs that can hold a significant integral ----------------------------
    // Tile (a, b), b <= a, exists when smax_a * smax_b > itol.  Order: blocks of PB consecutive bra pair groups
    // against chunks of QC consecutive ket pair groups, so that the tables of one (block, chunk) -- a few tens of
    // MB -- stay in L2 while the CTAs work through it; inside a chunk the partners of a come by decreasing
    // Schwarz bound (early exit).  A run = the tiles of one (a, chunk).
    std::vector<int> all(npg);
    std::iota(all.begin(), all.end(), 0);
    std::vector<TilePair> tl;
    std::vector<std::pair<long long, int>> runs;      // (first tile, # tiles) of every non-empty (a, chunk)
    make_tile_list(ts.pgs, all, all, itol, &tl, &runs, rank, nranks);      // this rank's bra blocks only
    const long long ntiles = (long long)tl.size();
    if (std::getenv("VB_DEBUG_PAIRS")) {
        // per tile: orbital pairs of either side that can pass the Schwarz screen against the other side's largest value
        long long h1[5] = {0, 0, 0, 0, 0}, h2[17] = {0};
        auto live = [&](const PGDesc& a, double smax_other) {
            int n = 0;
            for (int p = 0; p < a.np; ++p)
                if (sch[(size_t)ts.pg_pairs[2 * (a.pair_beg + p)] * nso + ts.pg_pairs[2 * (a.pair_beg + p) + 1]] * smax_other > itol) ++n;
            return n;
        };
        for (const TilePair& t : tl) {
            const int nP = live(ts.pgs[t.x], ts.pgs[t.y].smax), nQ = live(ts.pgs[t.y], ts.pgs[t.x].smax);
            const int jb = (nP + 7) / 8, mb = (nQ + 7) / 8;
            h1[jb]++; h2[jb * mb]++;
        }
        std::printf("[pairs] %lld tiles; column blocks of the first transform (of 4): 0:%lld 1:%lld 2:%lld 3:%lld 4:%lld; blocks of the second (of 16):", ntiles, h1[0], h1[1], h1[2], h1[3], h1[4]);
        for (int k = 0; k <= 16; ++k) if (h2[k]) std::printf(" %d:%lld", k, h2[k]);
        std::printf("\n");
    }
    // work items of the s/p kernel: pieces of <= m tiles of a run (they share the bra pair group); the d-shell
    // kernel takes single tiles.  Only this rank's items are kept (block-cyclic over the ranks; z = slot of the
    // item's first tile in the G hand-over buffer).
    std::vector<WorkItem> itl;
    long long my_tiles = 0;
    if (!gen) make_items(runs, ntiles, nsm, 0, 1, &itl, &my_tiles);
    long long mine = (long long)itl.size();
    if (gen) mine = ntiles;
    double t4 = now_ms();
    if (dbg_time) std::printf("[time] diag pass + tile list %.1f ms (tiles %lld items %zu)\n", t4 - t3, ntiles, itl.size());
    // ---- energy pass ---------------------------------------------------------------------------------
    this->sch.upload(sch, st);
    tiles.upload(tl, st); items.upload(itl, st);
    tileE.alloc((size_t)std::max<long long>(ntiles, 1));
    tileE.zero(st); counter.zero(st); counters.zero(st); pq_counters.zero(st);
    A.tiles = reinterpret_cast<const int2*>(tiles.p); A.ntiles = (int)ntiles; A.items = reinterpret_cast<const int4*>(items.p); A.nitems = (int)itl.size(); A.gbuf = nullptr; A.gslot_base = 0; A.tile_first = 0; A.tile_stride = 1; A.mode = 1; A.tau = tau_energy;
    A.sch = this->sch.p; A.tileE = tileE.p;
    bool split = false;   // VB_SPLIT=1: separate heavy / light launches (measured slower: 5.97 s vs 5.46 s on (H2O)_256)
    if (const char* e = std::getenv("VB_SPLIT")) split = std::atoi(e) != 0;
    bool csplit = !gen;   // class-split pass (vb_pclass.cuh); VB_CLASS_SPLIT=0 selects the all-in-one kernel
    if (const char* e = std::getenv("VB_CLASS_SPLIT")) csplit = csplit && std::atoi(e) != 0;
    ClassPlan cplan;
    if (csplit) { cplan = plan_classes(ts.pgs); csplit = cplan.ok; }
    // VB_PSEG=1: closed far-field form for the light classes (vb_pseg.cuh).  Measured slower than k_pclass (DESIGN.md section 5:
    // both are bound by the two DMMA transforms and the latency of short inner loops, not by the integral arithmetic) -- kept as
    // a documented experiment, off by default
    bool use_seg = false;
    if (const char* e = std::getenv("VB_PSEG")) use_seg = csplit && std::atoi(e) != 0;
    // (Tried: orbital pairs sorted by Schwarz value on the device and the density transforms of a tile restricted to the pairs that
    // can pass the reference's Schwarz screen in that tile -- on average 2.6 of 4 column blocks in the first and 7 of 16 blocks in
    // the second transform over the tiles of (H2O)_256, VB_DEBUG_PAIRS.  No gain in the light classes -- the tiles that carry
    // the work are the ones that need every block -- and the extra predicates cost the pp classes registers: 4.40 s against
    // 4.03 s for the pass.  profiles/r2_pair_screen_experiment.log)
    CK(cudaEventRecord(ev2, st));
    if (use_seg && mine > 0) {
        // far-field tables of every segment / primitive pair of this evaluation (part of the timed tile pass)
        fcp.alloc(pps.n); fcpf.alloc(pps.n);
        const long long npp = (long long)pps.n;
        k_far_tables<<<(unsigned)((npp + 255) / 256), 256, 0, st>>>(pps.p, pps_flat.p, npp, fcp.p, fcpf.p);
        CK(cudaGetLastError());
        launches++;
        cplan.fcp = fcp.p; cplan.fcpf = fcpf.p; cplan.lthr = far_lthr(tau_energy);
    }
    if (gen) {
        if (mine > 0) launch((int)mine, PART_ALL);
    } else if (mine > 0 && csplit) {
        long long cap_mb = 24576;
        if (const char* e = std::getenv("VB_GBUF_MB")) cap_mb = std::max(1, std::atoi(e));
        const long long chunk_tiles = std::max<long long>(PT_MAXQ, cap_mb * 1024 * 1024 / ((long long)g_cap * 8));
        gbuf.alloc((size_t)std::min(my_tiles, chunk_tiles) * g_cap);
        A.gbuf = gbuf.p; A.tile_first = 0; A.tile_stride = 1;
        float msc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ms_con = 0.f;
        for (size_t i0 = 0; i0 < itl.size();) {
            size_t i1 = i0;
            long long nt = 0;
            while (i1 < itl.size() && nt + itl[i1].y <= chunk_tiles) { nt += itl[i1].y; ++i1; }
            A.items = reinterpret_cast<const int4*>(items.p + i0); A.nitems = (int)(i1 - i0); A.gslot_base = itl[i0].z;
            launches += run_class_pass(A, cplan, nsm, nt, counter.p, st, dbg_time ? msc : nullptr, ev0, ev1);
            if (dbg_time) CK(cudaEventRecord(ev0, st));
            k_contract_items<<<std::max(1, (int)std::min<long long>(((long long)A.nitems * PT_MAXQ + 7) / 8, (long long)nsm * 16)), CI_THREADS, 0, st>>>(A, nt);
            CK(cudaGetLastError());
            launches++;
            if (dbg_time) { CK(cudaEventRecord(ev1, st)); CK(cudaEventSynchronize(ev1)); float ms = 0.f; CK(cudaEventElapsedTime(&ms, ev0, ev1)); ms_con += ms; }
            out->tile_launches += 1;
            i0 = i1;
        }
        if (dbg_time) {
            std::printf("[time] class pass:");
            for (int c = 0; c < 9; ++c) std::printf(" (%d|%d) %.1f", c / 3, c % 3, msc[c]);
            std::printf(" ms; contraction %.1f ms\n", ms_con);
        }
    } else if (mine > 0 && !split) {
        A.items = reinterpret_cast<const int4*>(items.p); A.nitems = (int)itl.size(); A.tile_first = 0; A.tile_stride = 1;
        launch((int)mine, PART_ALL);
        out->tile_launches += 1;
    } else if (mine > 0) {
        // heavy classes first (their share of G goes through the hand-over buffer), then the light classes and the
        // contraction; chunked so that the buffer stays bounded
        long long cap_mb = 4096;
        if (const char* e = std::getenv("VB_GBUF_MB")) cap_mb = std::max(1, std::atoi(e));
        const long long chunk_tiles = std::max<long long>(PT_MAXQ, cap_mb * 1024 * 1024 / ((long long)g_cap * 8));
        gbuf.alloc((size_t)std::min(my_tiles, chunk_tiles) * g_cap);
        A.gbuf = gbuf.p; A.tile_first = 0; A.tile_stride = 1;
        for (size_t i0 = 0; i0 < itl.size();) {
            size_t i1 = i0;
            long long nt = 0;
            while (i1 < itl.size() && nt + itl[i1].y <= chunk_tiles) { nt += itl[i1].y; ++i1; }
            A.items = reinterpret_cast<const int4*>(items.p + i0); A.nitems = (int)(i1 - i0); A.gslot_base = itl[i0].z;
            counter.zero(st);
            if (dbg_time) CK(cudaEventRecord(ev0, st));
            launch((int)(i1 - i0), PART_HEAVY);
            if (dbg_time) CK(cudaEventRecord(ev1, st));
            counter.zero(st);
            launch((int)(i1 - i0), PART_LIGHT);
            if (dbg_time) {
                CK(cudaEventRecord(ev3, st));
                CK(cudaStreamSynchronize(st));
                float mh = 0.f, ml = 0.f;
                CK(cudaEventElapsedTime(&mh, ev0, ev1)); CK(cudaEventElapsedTime(&ml, ev1, ev3));
                std::printf("[time] split chunk: heavy %.1f ms, light %.1f ms\n", mh, ml);
            }
            out->tile_launches += 1;
            i0 = i1;
        }
    }
    CK(cudaEventRecord(ev3, st));
    if (gen) out->tile_launches += mine > 0 ? 1 : 0;
    {
        // grid of the high part: 2^-42 of the next power of two above |E1| (|E2| < |E1|), the same on every rank
        double grid = 0.0;
        if (e1 != 0.0 && std::isfinite(e1)) grid = std::ldexp(1.0, std::ilogb(e1) + 1 - 42);
        if (split_sum) k_sum<<<1, 1024, 0, st>>>(tileE.p, ntiles, accum.p, accum.p + 1 + CNT_N, grid);
        else {
            k_sum<<<1, 1024, 0, st>>>(tileE.p, ntiles, accum.p);
            CK(cudaMemsetAsync(accum.p + 1 + CNT_N, 0, sizeof(double), st));
        }
    }
    CK(cudaGetLastError());
    launches++;
    {
        std::vector<unsigned long long> c;
        counters.download(c, st);
        // executed primitive quartets per class -> algorithmic flop count of this rank's tile pass
        std::vector<unsigned long long> pq;
        pq_counters.download(pq, st);
        double fl = 0.0;
        long long npq = 0, nfarq = 0;
        for (int a = 0; a < NPTYPE; ++a)
            for (int b = 0; b < NPTYPE; ++b) {
                if (!gen && (a >= 3 || b >= 3)) continue;      // s/p runs: the upper slots hold the far-field counts of the class kernels
                const unsigned long long far = (!gen && a < 3 && b < 3) ? std::min(pq[PQ_FAR + a * 3 + b], pq[a * NPTYPE + b]) : 0ull;
                // light classes: the quartets that went through the far-field segment form (vb_far.cuh) at ITS operation count
                const unsigned long long seg = (!gen && a < 2 && b < 2) ? std::min(pq[PQ_SEGFAR + a * 2 + b], far) : 0ull;
                fl += (double)(pq[a * NPTYPE + b] - far) * flops_prim_quartet(a, b) + (double)(far - seg) * flops_prim_quartet(a, b, 1) +
                      (double)seg * far_flops(a, b);
                npq += (long long)pq[a * NPTYPE + b];
                nfarq += (long long)far;
            }
        out->flops_model += fl; out->n_prim_quartets += npq; out->n_ao_quartets += nfarq;   // n_ao_quartets carries the far-field count
        if (!gen) out->flops_transform += 512.0 * (double)pq[PQ_DMMA];
        if (std::getenv("VB_DEBUG_PQ"))
            for (int a = 0; a < NPTYPE; ++a)
                for (int b = 0; b < NPTYPE; ++b)
                    if (pq[a * NPTYPE + b])
                        std::printf("class (%d|%d): %llu primitive quartets, %llu asymptotic, %llu in far-field segment form\n", a, b, pq[a * NPTYPE + b],
                                    a < 3 && b < 3 ? pq[PQ_FAR + a * 3 + b] : 0ull, a < 2 && b < 2 ? pq[PQ_SEGFAR + a * 2 + b] : 0ull);
        std::vector<double> cd(CNT_N);
        for (int i = 0; i < CNT_N; ++i) cd[i] = (double)c[i];
        CK(cudaMemcpyAsync(accum.p + 1, cd.data(), CNT_N * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev2, ev3)); out->t_tiles += ms;
    out->t_host_setup += (t3 - t2) + (t4 - t3);
    out->n_tiles += ntiles; out->n_tiles_mine += mine;
    out->enucrep = enuc; out->e1 = e1; out->wfnorm = wfnorm;
}

namespace {

Wavefunction default_wavefunction(const Input& in)   // guess_energy, valence.F90:324-336
{
    Wavefunction wf;
    wf.nnd = in.nnd();
    wf.nso = wf.nnd + in.ndocc;
    wf.sym = true;
    wf.subject = -1;
    for (int i = 0; i < wf.nnd; ++i) { wf.bra.push_back(i); wf.ket.push_back(i); }
    for (int d = 0; d < in.ndocc; ++d) for (int k = 0; k < 2; ++k) { wf.bra.push_back(wf.nnd + d); wf.ket.push_back(wf.nnd + d); }
    return wf;
}

}  // namespace

void Engine::attach_comm(int rank, int nranks, const std::string& key)
{
    CK(cudaSetDevice(impl_->device));
    impl_->comm.reset(nranks > 1 ? new Comm(rank, nranks, key) : nullptr);
}
void Engine::attach_nccl(int rank, int nranks, void* nccl_comm)
{
    CK(cudaSetDevice(impl_->device));
    impl_->comm.reset(nranks > 1 ? new Comm(rank, nranks, nccl_comm) : nullptr);
}
int Engine::comm_rank() const { return impl_->crank(); }
int Engine::comm_size() const { return impl_->csize(); }

// guess_energy; collective when a communicator is attached
void Engine::energy(EnergyResult* out)
{
    Impl& I = *impl_;
    if (I.csize() == 1) { energy_partial(0, 1, out); energy_finish(out); return; }
    const int r = I.crank(), n = I.csize();
    bool shard = true;                         // host tables once per node (VB_SHARD_SETUP=0: every rank builds everything)
    if (const char* e = std::getenv("VB_SHARD_SETUP")) shard = std::atoi(e) != 0;
    bool gather = shard;                       // exchange the shares on the device (VB_GATHER_TABLES=0: host shared memory)
    if (const char* e = std::getenv("VB_GATHER_TABLES")) gather = gather && std::atoi(e) != 0;
    if (gather) {
        I.use_gather = true;
        try { energy_partial(r, n, out); } catch (...) { I.use_gather = false; throw; }
        I.use_gather = false;
        I.comm->allreduce_sum(I.accum.p, 2 + CNT_N, I.st);
        energy_finish(out);
        return;
    }
    if (shard) {
        std::string key = "job";
        if (const char* k = std::getenv("VB_SHARD_KEY")) key = k;
        else if (const char* p = std::getenv("MASTER_PORT")) key = std::string("port") + p;
        shard_tables(r, n, "/dev/shm/valence_b200_tables_" + key + "_");
        I.comm->barrier(I.st);
    }
    energy_partial(r, n, out);
    I.comm->allreduce_sum(I.accum.p, 2 + CNT_N, I.st);
    energy_finish(out);
}

void Engine::energy_partial(int rank, int nranks, EnergyResult* out)
{
    Impl& I = *impl_;
    CK(cudaSetDevice(I.device));
    *out = EnergyResult();
    if (!I.prepared) {      // otherwise shard_tables opened this evaluation
        I.launches = 0;
        I.t_begin = now_ms();
        g_h2d_bytes = 0; g_d2h_bytes = 0;
        I.prepare(in_, -1);
    }
    I.prepared = false;
    double t1 = now_ms();
    out->t_1e = t1 - I.t_begin;
    Wavefunction wf = default_wavefunction(in_);
    I.split_sum = true;
    I.sch_cache_key = 0;
    try { I.evaluate(in_, wf, nullptr, false, rank, nranks, out, &I.sch_cache); } catch (...) { I.split_sum = false; throw; }
    I.split_sum = false;
    I.sch_cache_n = wf.nso; I.sch_cache_key = I.state_key(in_);     // default lists: entry s is orbital s
}

// Table build shared by the ranks of one node.  One process per GPU means N ranks on one host build the same pair
// tables N times per energy (the largest host cost of an 8-GPU step).  Here every rank builds the pair groups
// i with (i / 16) mod N == rank and publishes them as the file prefix + rank (put the prefix on /dev/shm); after a
// barrier of the caller (api.py: Engine.energy_distributed) the next energy_partial maps all N files and merges them:
// the tables are bitwise those of an unsharded build (tests/host/test_setup_host.cpp).
void Engine::shard_tables(int rank, int nranks, const std::string& prefix)
{
    Impl& I = *impl_;
    CK(cudaSetDevice(I.device));
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("shard_tables: bad rank / nranks");
    I.launches = 0;
    I.t_begin = now_ms();
    g_h2d_bytes = 0; g_d2h_bytes = 0;
    I.prepare(in_, -1);
    I.prepared = true;
    int bas_lmax = 0;
    for (const GShell& gs : I.bas.shells) bas_lmax = std::max(bas_lmax, gs.l);
    const double tau_diag = std::min(I.tau, 0.01 * I.itol * I.itol);
    TileOpts o;
    o.shard_mode = 1; o.shard_rank = rank; o.shard_nranks = nranks; o.shard_prefix = prefix;
    TileSetup scratch;
    build_tiles(in_, I.bas, default_wavefunction(in_), I.orbs2e, tau_diag, bas_lmax < 2, &scratch, o);
    I.shard_pending = true; I.shard_rank = rank; I.shard_nranks = nranks; I.shard_prefix = prefix;
    I.shard_file = prefix + std::to_string(rank);
}

namespace {

struct PtCfg { int dq_cap2, g_cap, pp_cap, sp_cap, boys_cap, hs_cap, hs_ld; size_t smem; };

// shared-memory layout of k_ptile for the given table maxima (same rules as the energy pass)
PtCfg pt_cfg(int max_ne, int max_np, int max_nsp, int max_npp)
{
    PtCfg c;
    const int dq_cap = std::max(1, max_ne * max_np);
    c.hs_ld = std::max(1, max_np) | 1;
    c.dq_cap2 = (dq_cap + 1) & ~1;
    c.hs_cap = (max_ne * c.hs_ld + 1) & ~1;
    c.sp_cap = std::max(1, max_nsp);
    c.g_cap = (max_np * max_np + 1) & ~1;
    c.pp_cap = std::max(1, max_npp);
    c.boys_cap = 0;
    constexpr int nw = PT_MAX_WARPS;
    c.smem = ((size_t)c.dq_cap2 + (size_t)PT_MAXQ * c.g_cap + (size_t)nw * PT_SCRATCH) * sizeof(double) + (size_t)c.sp_cap * sizeof(SPRec);
    if (c.smem + (size_t)c.pp_cap * sizeof(PrimPair) <= PT_SMEM_MAX) c.smem += (size_t)c.pp_cap * sizeof(PrimPair);
    else c.pp_cap = 0;
    if (c.smem + BOYS_S_SIZE * sizeof(double) <= PT_SMEM_MAX) { c.boys_cap = BOYS_S_SIZE; c.smem += BOYS_S_SIZE * sizeof(double); }
    return c;
}

}  // namespace

// first_order_opt with an integral cache.  The (ib,jb) loop of first_order_opt (valence.F90:674-764) replaces ONE
// orbital of the bra and of the ket list; every orbital-level integral that does not involve that entry is the same in
// all norbas(norbas+1)/2 evaluations -- the reference keeps those in eribuf when store_eri is set
// (valence.F90:1227-1273).  Here:
//   * the subject entry forms a pair-group family of its own ("subject" pair groups, rebuilt per (ib,jb)); all other
//     ("free") pair groups and their tables are built and uploaded once;
//   * one integral pass (k_ptile, mode 2) leaves G of every canonical free tile in HBM (g_cap doubles per tile; sized
//     for the 180 GB of a B200: 38 GB for (H2O)_256);
//   * every (ib,jb) evaluation = cofactors of the substituted lists + integrals of the subject tiles (k_ptile) +
//     a contraction-only pass over the cached tiles (k_contract), in the reference's non-symmetric task order.
// Returns false when the cache does not apply (d shells, not enough memory); the caller then runs the plain loop.
bool Engine::Impl::first_order_cached(const Input& in, const Wavefunction& wf, int es, int iorb, const std::vector<double>& sch_tab,
                                      int rank, int nranks, std::vector<double>* ham, std::vector<double>* ovl, EnergyResult* acc)
{
    int bas_lmax = 0;
    for (const GShell& gs : bas.shells) bas_lmax = std::max(bas_lmax, gs.l);
    if (bas_lmax >= 2) return false;
    if (const char* e = std::getenv("VB_FO_CACHE")) if (std::atoi(e) == 0) return false;
    const int nso = wf.nso, norbs = in.norbs();
    const int norbas = (int)in.orbitals[iorb].xp.size();
    const double tau_diag = std::min(tau, 0.01 * itol * itol), tau_energy = std::min(tau, 0.01 * itol);
    const bool dbg_time = std::getenv("VB_DEBUG_TIME") != nullptr;
    double tm0 = now_ms();
    Wavefunction wb = wf;              // unsubstituted lists in the non-symmetric task order
    wb.sym = false;
    wb.subject = -1;
    Wavefunction w2 = wb;
    w2.subject = es;
    auto substitute = [&](int ib, int jb) {
        const int idf = in.orbitals[iorb].xp[ib], jdf = in.orbitals[iorb].xp[jb];
        w2.bra[es] = idf < 1 ? norbs + idf - 1 : norbs + ib;
        w2.ket[es] = jdf < 1 ? norbs + jdf - 1 : norbs + jb;
    };
    // ---- bound of the primitive weights over every list of the loop (consistent pruning of all tables) ----
    double wall = 0.0;
    {
        TileOpts mo;
        mo.isolate = es; mo.measure_only = true;
        TileSetup tmp;
        build_tiles(in, bas, wb, orbs2e, tau_diag, true, &tmp, mo);
        wall = tmp.wmax;
        mo.only_subject = true;
        for (int ib = 0; ib < norbas; ++ib) {
            substitute(ib, ib);
            build_tiles(in, bas, w2, orbs2e, tau_diag, true, &tmp, mo);
            wall = std::max(wall, tmp.wmax);
        }
    }
    const double tmA = now_ms();
    // ---- all tables of the unsubstituted lists; the free part stays resident ---------------------------------
    TileOpts fo;
    fo.isolate = es; fo.wcut = wall;
    TileSetup& tsF = ts_keep;        // the storage of the energy pass's host tables (already paged in; rebuilt by every evaluation anyway)
    build_tiles(in, bas, wb, orbs2e, tau_diag, true, &tsF, fo);
    const double tmB = now_ms();
    if (dbg_time) std::printf("[fo] weight bounds %.1f ms, table build %.1f ms\n", tmA - tm0, tmB - tmA);
    const int nfree = tsF.n_free_pg, npgF = (int)tsF.pgs.size();
    if (nfree == 0) return false;
    const size_t fPairs = tsF.n_free_pairs, fSps = tsF.n_free_sps, fPps = tsF.n_free_pps, fD = tsF.n_free_d;
    const size_t sPairs = tsF.pg_pairs.size() / 2 - fPairs, sSps = tsF.sps.size() - fSps, sPps = tsF.pps.size() - fPps, sD = tsF.dmat.size() - fD;
    if (tsF.max_np > 32) throw std::runtime_error("valence_b200: pair group too large");
    auto set_smax = [&](std::vector<PGDesc>& v, const std::vector<int>& pairs, size_t pair_shift, size_t from) {
        for (size_t i = from; i < v.size(); ++i) {
            PGDesc& pg = v[i];
            double m = 0.0;
            for (int p = 0; p < pg.np; ++p) {
                const size_t k = (size_t)pg.pair_beg - pair_shift + p;
                const double x = sch_tab[(size_t)pairs[2 * k] * nso + pairs[2 * k + 1]];
                if (x > m) m = x;
            }
            pg.smax = m;
        }
    };
    std::vector<PGDesc> hp(tsF.pgs.begin(), tsF.pgs.begin() + nfree);
    set_smax(hp, tsF.pg_pairs, 0, 0);
    // device tables with room for the per-evaluation subject part (3x the size of the unsubstituted one)
    const size_t room = 3;
    pgs.alloc((size_t)nfree + room * (npgF - nfree) + 64);
    pg_pairs.alloc(2 * (fPairs + room * sPairs + 64));
    sps.alloc(fSps + room * sSps + 256);
    pps.alloc(fPps + room * sPps + 1024);
    pps_flat.alloc(fPps + room * sPps + 1024);
    dmat.alloc(fD + room * sD + 4096);
    {
        pg_pairs.upload_at(0, tsF.pg_pairs.data(), 2 * fPairs, st);
        sps.upload_at(0, tsF.sps.data(), fSps, st);
        pps.upload_at(0, tsF.pps.data(), fPps, st);
        pps_flat.upload_at(0, tsF.pps_flat.data(), fPps, st);
        dmat.upload_at(0, tsF.dmat.data(), fD, st);
        pgs.upload_at(0, hp, st);
        CK(cudaStreamSynchronize(st));
    }
    // ---- canonical pair groups (g >= h) and the pair permutation of the flipped ones ----------------------------
    const long long ng = (long long)tsF.groups.size();
    std::vector<int> idx_of((size_t)(ng * ng), -1);
    for (int x = 0; x < nfree; ++x) idx_of[(size_t)(hp[x].g * ng + hp[x].h)] = x;
    std::vector<int> flip(nfree), canon(nfree), perm(fPairs);
    for (int x = 0; x < nfree; ++x) {
        const PGDesc& pg = hp[x];
        flip[x] = idx_of[(size_t)(pg.h * ng + pg.g)];
        canon[x] = (pg.g < pg.h && flip[x] >= 0) ? flip[x] : x;
        const PGDesc& cg = hp[canon[x]];
        for (int p = 0; p < pg.np; ++p) {
            int found = p;
            if (canon[x] != x) {
                const int s_ = tsF.pg_pairs[2 * (pg.pair_beg + p)], t_ = tsF.pg_pairs[2 * (pg.pair_beg + p) + 1];
                found = -1;
                for (int q = 0; q < cg.np && found < 0; ++q)
                    if (tsF.pg_pairs[2 * (cg.pair_beg + q)] == t_ && tsF.pg_pairs[2 * (cg.pair_beg + q) + 1] == s_) found = q;
                if (found < 0 || cg.np != pg.np) throw std::runtime_error("valence_b200: first_order cache: pair groups do not mirror");
            }
            perm[pg.pair_beg + p] = found;
        }
    }
    fo_perm.upload(perm, st);
    // ---- integral pass over the canonical free tiles -> cache ----------------------------------------------
    std::vector<int> cand;
    for (int x = 0; x < nfree; ++x) if (canon[x] == x) cand.push_back(x);
    std::vector<TilePair> tlc;
    std::vector<std::pair<long long, int>> runs;
    const double tmC = now_ms();
    make_tile_list(hp, cand, cand, itol, &tlc, &runs);
    if (dbg_time) std::printf("[fo] uploads + mirror map %.1f ms, tile list %.1f ms\n", tmC - tmB, now_ms() - tmC);
    std::vector<WorkItem> itc;
    long long my_tiles = 0;
    make_items(runs, (long long)tlc.size(), nsm, rank, nranks, &itc, &my_tiles);
    PtCfg cfg = pt_cfg(tsF.max_ne, tsF.max_np, tsF.max_nsp, tsF.max_npp);
    {
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        const double need = (double)my_tiles * cfg.g_cap * 8.0;
        double cap = 0.8 * ((double)fr + (double)gcache.cap * 8.0);
        if (const char* e = std::getenv("VB_FO_CACHE_MB")) cap = std::min(cap, std::atof(e) * 1048576.0);
        // The cached and the plain loop split the tiles over the ranks differently, so every rank must take the same one:
        // with the engine's communicator the verdicts are summed; a caller that does its own all-reduce (nranks > 1, no
        // communicator) gets an error instead of a silent mix.
        double misfit = need > cap ? 1.0 : 0.0;
        if (fo_collective) {
            one_e.alloc(1);
            CK(cudaMemcpyAsync(one_e.p, &misfit, sizeof(double), cudaMemcpyHostToDevice, st));
            comm->allreduce_sum(one_e.p, 1, st);
            CK(cudaMemcpyAsync(&misfit, one_e.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        if (misfit > 0.0) {
            // the plain loop regenerates every integral norbas(norbas+1)/2 times: a caller that cannot afford that
            // (bench.py on large clusters) asks for an error instead
            bool require = nranks > 1 && !fo_collective;
            if (const char* e = std::getenv("VB_FO_REQUIRE_CACHE")) require = require || std::atoi(e) != 0;
            if (require) throw std::runtime_error("valence_b200: the integral cache of first_order_opt does not fit in device memory");
            return false;
        }
    }
    if (cfg.smem > PT_SMEM_MAX) throw std::runtime_error("valence_b200: orbital basis set too large for one pair-group tile");
    gcache.alloc((size_t)std::max<long long>(1, my_tiles) * cfg.g_cap);
    counter.alloc(1); counters.alloc(CNT_N); pq_counters.alloc(NPTYPE * NPTYPE);
    gred.alloc((size_t)nsm * PT_MAXQ * cfg.g_cap);
    CK(cudaFuncSetAttribute(k_ptile<PART_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    TileArgs A;
    std::memset(&A, 0, sizeof A);
    auto fill_cfg = [&](const PtCfg& c) {
        A.dq_cap = c.dq_cap2; A.hs_cap = c.hs_cap; A.hs_ld = c.hs_ld; A.g_cap = c.g_cap; A.boys_cap = c.boys_cap; A.pp_cap = c.pp_cap; A.sp_cap = c.sp_cap;
    };
    A.gred = gred.p;
    A.pgs = pgs.p; A.pg_pairs = pg_pairs.p; A.sps = sps.p; A.pps = pps.p; A.pps_flat = pps_flat.p; A.dmat = dmat.p;
    A.pq_counters = pq_counters.p; A.boys = boys.p; A.boys_small = boys_small.p; A.counter = counter.p; A.counters = counters.p;
    A.nso = nso; A.nnd = wf.nnd; A.sym = 0; A.subject = es; A.itol = itol;
    A.tile_first = 0; A.tile_stride = 1;
    fill_cfg(cfg);
    auto add_pq = [&]() {
        std::vector<unsigned long long> pq;
        pq_counters.download(pq, st);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                const unsigned long long far = std::min(pq[PQ_FAR + a * 3 + b], pq[a * NPTYPE + b]);
                acc->flops_model += (double)(pq[a * NPTYPE + b] - far) * flops_prim_quartet(a, b) + (double)far * flops_prim_quartet(a, b, 1);
                acc->n_prim_quartets += (long long)pq[a * NPTYPE + b];
            }
    };
    double tm1 = now_ms();
    if (my_tiles > 0) {
        tiles.upload(tlc, st); items.upload(itc, st);
        counter.zero(st); counters.zero(st); pq_counters.zero(st);
        A.tiles = reinterpret_cast<const int2*>(tiles.p); A.ntiles = (int)tlc.size(); A.items = reinterpret_cast<const int4*>(items.p); A.nitems = (int)itc.size();
        A.gbuf = gcache.p; A.gslot_base = 0; A.mode = 2; A.tau = tau_energy;
        CK(cudaEventRecord(ev2, st));
        bool csplit = true;
        if (const char* e = std::getenv("VB_CLASS_SPLIT")) csplit = std::atoi(e) != 0;
        ClassPlan cplan;
        if (csplit) { cplan = plan_classes(std::vector<PGDesc>(hp.begin(), hp.begin() + nfree)); csplit = cplan.ok; }
        if (csplit) {
            launches += run_class_pass(A, cplan, nsm, my_tiles, counter.p, st, nullptr, ev0, ev1);
        } else {
            k_ptile<PART_ALL><<<std::max(1, std::min(nsm, (int)itc.size())), pt_threads(PART_ALL), cfg.smem, st>>>(A);
            CK(cudaGetLastError());
            launches++;
        }
        CK(cudaEventRecord(ev3, st));
        acc->tile_launches++;
        add_pq();
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev2, ev3));
        acc->t_tiles += ms;
    }
    acc->n_tiles += (long long)tlc.size();
    // ---- variant tiles: every free tile of the non-symmetric task order, mapped onto its canonical tile ---------
    std::vector<int4> vts;
    vts.reserve((size_t)my_tiles * 8);
    auto add_variant = [&](int x, int y, int slot, int swp) {       // two int4 per variant tile (k_contract)
        vts.push_back(make_int4(hp[x].pair_beg, hp[y].pair_beg, hp[x].np | (hp[y].np << 8) | (swp << 16), 0));
        vts.push_back(make_int4(slot, 0, 0, 0));
    };
    for (const WorkItem& it : itc)
        for (int j = 0; j < it.y; ++j) {
            const TilePair t = tlc[(size_t)it.x + j];
            const int slot = it.z + j;
            const int A0 = t.x, B0 = t.y;
            const int Af = (flip[A0] >= 0 && flip[A0] != A0 && canon[flip[A0]] == A0) ? flip[A0] : -1;
            const int Bf = (flip[B0] >= 0 && flip[B0] != B0 && canon[flip[B0]] == B0) ? flip[B0] : -1;
            if (A0 != B0) {
                const int xs[2] = {A0, Af}, ys[2] = {B0, Bf};
                for (int i = 0; i < 2; ++i)
                    for (int k = 0; k < 2; ++k) {
                        const int x = xs[i], y = ys[k];
                        if (x < 0 || y < 0) continue;
                        if (x >= y) add_variant(x, y, slot, 0);
                        else add_variant(y, x, slot, 1);
                    }
            } else {
                add_variant(A0, A0, slot, 0);
                if (Af >= 0) {
                    add_variant(Af, Af, slot, 0);
                    add_variant(std::max(A0, Af), std::min(A0, Af), slot, 0);
                }
            }
        }
    const long long nvt = (long long)vts.size() / 2;
    vtiles.upload(vts, st);
    { std::vector<int4>().swap(vts); }
    this->sch.upload(sch_tab, st);
    A.sch = this->sch.p;
    double tm2 = now_ms();
    if (dbg_time) { CK(cudaStreamSynchronize(st)); std::printf("[fo] tables %.1f ms, cache pass %.1f ms (%lld canonical tiles, %lld mine, %.2f GB), %lld variant tiles\n", tm1 - tm0, tm2 - tm1, (long long)tlc.size(), my_tiles, (double)my_tiles * cfg.g_cap * 8e-9, nvt); }
    // ---- rank-one form of the (ib,jb) loop (large single-determinant wavefunctions) ------------------------------
    // Substituting slot e of the bra by chi_ib and of the ket by chi_jb borders the block N (the spin block of the subject
    // without it) by one row and one column:  det M' = det N * sigma,  M'^-1 = Nbar^-1 + abar bbar^T / sigma  with
    // a = N^-1 c_jb, b = N^-T r_ib, sigma = <chi_ib|chi_jb> - r_ib^T a.  sigma * W' is then sigma * W0 plus terms bilinear in
    // (x, y) = (bbar, abar): for the tiles free of the subject entry  sum G W' sigma = sigma E0 + x^T F y  with ONE Fock-like
    // matrix F for all norbas(norbas+1)/2 elements (contract_tile<FMODE>); only the subject tiles are walked per element,
    // with W' in rank-one form.  Nothing is divided by sigma, so a singular substituted block (a basis function orthogonal
    // to the space it replaces, which Givens determinants handle silently) needs no special treatment.
    {
        int fast_min = 64;
        if (const char* e = std::getenv("VB_FAST_MIN_N")) fast_min = std::atoi(e);
        bool r1path = in.npair == 0 && std::max(in.nalpha(), in.nbeta()) > fast_min;
        if (const char* e = std::getenv("VB_FO_RANK1")) r1path = r1path && std::atoi(e) != 0;
        if (r1path) {
            const int nelec = in.nelec(), nao = bas.nao;
            // entry-level overlap / core-Hamiltonian of the unsubstituted lists
            {
                std::vector<int2> prs((size_t)nso * nso);
                for (int a = 0; a < nso; ++a)
                    for (int b = 0; b < nso; ++b) prs[(size_t)a * nso + b] = make_int2(wb.bra[wb.slot(a, 0)], wb.ket[wb.slot(b, 0)]);
                opairs.upload(prs, st);
                Se.alloc(prs.size()); He.alloc(prs.size());
                k_orb_1e<<<(unsigned)((prs.size() + 127) / 128), 128, 0, st>>>(optr.p, oao.p, oc.p, opairs.p, (int)prs.size(), nao, S.p, H.p, Se.p, He.p);
                CK(cudaGetLastError());
                launches++;
            }
            // spin lists by slot position (set_up_unpaired_docc, valence.F90:2461-2480)
            std::vector<int> entry_of_slot(wb.bra.size());
            for (int a = 0; a < nso; ++a)
                for (int k = 0; k < wb.nslots(a); ++k) entry_of_slot[wb.slot(a, k)] = a;
            std::vector<int> ea, eb;
            for (int i = 0; i < in.nunpd; ++i) ea.push_back(entry_of_slot[i]);
            for (int d = 0; d < in.ndocc; ++d) { ea.push_back(entry_of_slot[in.nunpd + 2 * d]); eb.push_back(entry_of_slot[in.nunpd + 2 * d + 1]); }
            const bool in_a = std::find(ea.begin(), ea.end(), es) != ea.end();
            const std::vector<int>& esub = in_a ? ea : eb;
            const std::vector<int>& eoth = in_a ? eb : ea;
            if (std::find(esub.begin(), esub.end(), es) == esub.end()) throw std::runtime_error("valence_b200: first_order: subject entry not in a spin block");
            std::vector<int> en, posn(nso, -1), poso(nso, -1);
            for (int a : esub) if (a != es) { posn[a] = (int)en.size(); en.push_back(a); }
            for (size_t k = 0; k < eoth.size(); ++k) poso[eoth[k]] = (int)k;
            const int nn = (int)en.size(), no = (int)eoth.size();
            gjout.alloc(4);
            gpu_inverse(nso, en, en_dev, Ma, Ni, gjout.p);
            gpu_inverse(nso, eoth, eo_dev, Mb, Mbi, gjout.p + 2);
            std::vector<double> gd;
            gjout.download(gd, st);
            if (!(gd[0] != 0.0) || !(gd[2] != 0.0) || std::min(gd[1], gd[3]) < 1e-13)
                throw std::runtime_error("valence_b200: singular spin-block overlap matrix (linearly dependent orbitals)");
            const double Dd = gd[0] * gd[2];
            posn_dev.upload(posn, st); poso_dev.upload(poso, st);
            Qd.alloc((size_t)nso * nso); Rd.alloc((size_t)nso * nso);
            k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(Ni.p, nn, posn_dev.p, posn_dev.p, nso, Qd.p);
            k_entry_density<<<(nso * nso + 255) / 256, 256, 0, st>>>(Mbi.p, no, poso_dev.p, poso_dev.p, nso, Rd.p);
            one_e.alloc(2);
            k_one_electron_energy<<<1, 1024, 0, st>>>(Se.p, He.p, Qd.p, Rd.p, nso * nso, one_e.p);
            CK(cudaGetLastError());
            launches += 3;
            std::vector<double> oe;
            one_e.download(oe, st);
            const double E1_0 = oe[0], S0 = oe[1];
            // ---- F pass over the free tiles: E2_0 and F ----
            Fmat.alloc((size_t)nso * nso);
            Fmat.zero(st);
            tileE.alloc((size_t)std::max<long long>(1, nvt));
            tileE.zero(st); counters.zero(st);
            A.tiles = nullptr; A.ntiles = 0; A.items = nullptr; A.nitems = 0; A.mode = 1; A.tau = tau_energy;
            A.Pa = Qd.p; A.Pb = Rd.p; A.c0 = 1.0; A.cof = nullptr; A.ndp = 0; A.r1 = 0; A.Fmat = Fmat.p; A.tileE = tileE.p;
            {
                std::vector<int> nshb(nso), nshk(nso);
                for (int a = 0; a < nso; ++a) { nshb[a] = (int)orbs2e[wb.bra[wb.slot(a, 0)]].sh.size(); nshk[a] = (int)orbs2e[wb.ket[wb.slot(a, 0)]].sh.size(); }
                nsh_bra.upload(nshb, st); nsh_ket.upload(nshk, st);
                A.nsh_bra = nsh_bra.p; A.nsh_ket = nsh_ket.p;
            }
            double E2_0 = 0.0;
            std::vector<unsigned long long> cfree(CNT_N, 0ull);
            CK(cudaEventRecord(ev2, st));
            if (nvt > 0) {
                const int grid = (int)std::min<long long>((nvt + CT_THREADS / 32 - 1) / (CT_THREADS / 32), (long long)nsm * 16);
                k_contract<<<grid, CT_THREADS, 0, st>>>(A, vtiles.p, nvt, fo_perm.p, gcache.p, 0);
                k_sum<<<1, 1024, 0, st>>>(tileE.p, nvt, accum.p);
                CK(cudaGetLastError());
                launches += 2;
                CK(cudaMemcpyAsync(&E2_0, accum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
                counters.download(cfree, st);
            }
            CK(cudaEventRecord(ev3, st));
            CK(cudaEventSynchronize(ev3));
            { float ms = 0.f; CK(cudaEventElapsedTime(&ms, ev2, ev3)); acc->t_tiles += ms; if (dbg_time) std::printf("[fo] F pass %.1f ms (%lld variant tiles)\n", ms, nvt); }
            A.Fmat = nullptr;
            std::vector<double> hF, hHe, hSe, hNi;
            Fmat.download(hF, st); He.download(hHe, st); Se.download(hSe, st); Ni.download(hNi, st);
            // ---- overlaps / core-Hamiltonian elements with the dummy orbitals ----
            auto dummy = [&](int k) { const int idf = in.orbitals[iorb].xp[k]; return idf < 1 ? norbs + idf - 1 : norbs + k; };
            std::vector<int2> prs;
            for (int ib = 0; ib < norbas; ++ib) for (int t = 0; t < nso; ++t) prs.push_back(make_int2(dummy(ib), wb.ket[wb.slot(t, 0)]));
            for (int jb = 0; jb < norbas; ++jb) for (int a = 0; a < nso; ++a) prs.push_back(make_int2(wb.bra[wb.slot(a, 0)], dummy(jb)));
            for (int ib = 0; ib < norbas; ++ib) for (int jb = 0; jb < norbas; ++jb) prs.push_back(make_int2(dummy(ib), dummy(jb)));
            opairs.upload(prs, st);
            DBuf<double> os, oh;
            os.alloc(prs.size()); oh.alloc(prs.size());
            k_orb_1e<<<(unsigned)((prs.size() + 127) / 128), 128, 0, st>>>(optr.p, oao.p, oc.p, opairs.p, (int)prs.size(), nao, S.p, H.p, os.p, oh.p);
            CK(cudaGetLastError());
            launches++;
            std::vector<double> vs, vh;
            os.download(vs, st); oh.download(vh, st);
            const double* Sx = vs.data(); const double* Sy = Sx + (size_t)norbas * nso; const double* Sd = Sy + (size_t)norbas * nso;
            const double* Hx = vh.data(); const double* Hy = Hx + (size_t)norbas * nso; const double* Hd = Hy + (size_t)norbas * nso;
            // x_ib (bra side), y_jb (ket side): entry-indexed, -1 at the subject
            std::vector<std::vector<double>> xs(norbas, std::vector<double>(nso, 0.0)), ys(norbas, std::vector<double>(nso, 0.0));
            for (int k = 0; k < norbas; ++k) {
                for (int r = 0; r < nn; ++r) {
                    double a = 0.0, b = 0.0;
                    for (int c = 0; c < nn; ++c) {
                        a += hNi[(size_t)r * nn + c] * Sy[(size_t)k * nso + en[c]];      // a[k'] = sum_r Ninv[k'][r] c[r]
                        b += Sx[(size_t)k * nso + en[c]] * hNi[(size_t)c * nn + r];      // b[r] = sum_c rrow[c] Ninv[c][r]
                    }
                    ys[k][en[r]] = a; xs[k][en[r]] = b;
                }
                xs[k][es] = -1.0; ys[k][es] = -1.0;
            }
            // F y_jb, He y_jb, Se y_jb over the entries other than the subject (rows s)
            auto bil = [&](const std::vector<double>& Mx, const std::vector<double>& x, const std::vector<double>& y) {
                double tsum = 0.0;
                for (int a = 0; a < nso; ++a) {
                    if (a == es || x[a] == 0.0) continue;
                    double rs = 0.0;
                    const double* row = Mx.data() + (size_t)a * nso;
                    for (int b = 0; b < nso; ++b) if (b != es) rs += row[b] * y[b];
                    tsum += x[a] * rs;
                }
                return tsum;
            };
            r1x.alloc(nso); r1y.alloc(nso);
            std::vector<int> avec2, bvec2;
            for (int ib = 0; ib < norbas; ++ib)
                for (int jb = 0; jb <= ib; ++jb) {
                    double tp0 = now_ms();
                    substitute(ib, jb);
                    const std::vector<double>& x = xs[ib];
                    const std::vector<double>& y = ys[jb];
                    double sigma = Sd[(size_t)ib * norbas + jb];
                    for (int c = 0; c < nn; ++c) sigma -= Sx[(size_t)ib * nso + en[c]] * y[en[c]];
                    // x^T He' y and x^T Se' y with row / column `es` of the substituted lists
                    double xHy = bil(hHe, x, y), xSy = bil(hSe, x, y);
                    for (int t = 0; t < nso; ++t) if (t != es) { xHy -= Hx[(size_t)ib * nso + t] * y[t]; xSy -= Sx[(size_t)ib * nso + t] * y[t]; }
                    for (int a = 0; a < nso; ++a) if (a != es) { xHy -= x[a] * Hy[(size_t)jb * nso + a]; xSy -= x[a] * Sy[(size_t)jb * nso + a]; }
                    xHy += Hd[(size_t)ib * norbas + jb]; xSy += Sd[(size_t)ib * norbas + jb];
                    const double xFy = bil(hF, x, y);
                    // ---- subject tiles with W' in rank-one form ----
                    TileOpts so;
                    so.isolate = es; so.only_subject = true; so.wcut = wall;
                    TileSetup tsS;
                    build_tiles(in, bas, w2, orbs2e, tau_diag, true, &tsS, so);
                    const int nS = (int)tsS.pgs.size();
                    if (tsS.max_np > 32) throw std::runtime_error("valence_b200: pair group too large");
                    hp.resize(nfree);
                    for (PGDesc pg : tsS.pgs) {
                        pg.pair_beg += (int)fPairs;
                        pg.d_off += (long long)fD;
                        for (int t = 0; t <= NPTYPE; ++t) { pg.pp_beg[t] += (int)fPps; pg.sp_beg[t] += (int)fSps; }
                        hp.push_back(pg);
                    }
                    for (SPRec& sr : tsS.sps) sr.pp_beg += (int)fPps;
                    set_smax(hp, tsS.pg_pairs, fPairs, nfree);
                    {
                        std::vector<PGDesc> tail(hp.begin() + nfree, hp.end());
                        pgs.upload_at(nfree, tail, st);
                        pg_pairs.upload_at(2 * fPairs, tsS.pg_pairs, st);
                        sps.upload_at(fSps, tsS.sps, st);
                        pps.upload_at(fPps, tsS.pps, st);
                        pps_flat.upload_at(fPps, tsS.pps_flat, st);
                        dmat.upload_at(fD, tsS.dmat, st);
                    }
                    std::vector<int> nshb(nso), nshk(nso);
                    for (int a = 0; a < nso; ++a) {
                        nshb[a] = (int)orbs2e[w2.bra[w2.slot(a, 0)]].sh.size();
                        nshk[a] = (int)orbs2e[w2.ket[w2.slot(a, 0)]].sh.size();
                    }
                    nsh_bra.upload(nshb, st); nsh_ket.upload(nshk, st);
                    avec2.clear(); bvec2.clear();
                    for (int q = nfree; q < nfree + nS; ++q) avec2.push_back(q);
                    for (int q = 0; q < nfree + nS; ++q) bvec2.push_back(q);
                    std::vector<TilePair> tls;
                    make_tile_list(hp, avec2, bvec2, itol, &tls, &runs);
                    std::vector<WorkItem> its;
                    long long mine = 0;
                    make_items(runs, (long long)tls.size(), nsm, rank, nranks, &its, &mine);
                    const PtCfg c2 = pt_cfg(std::max(tsF.max_ne, tsS.max_ne), std::max(tsF.max_np, tsS.max_np), std::max(tsF.max_nsp, tsS.max_nsp),
                                            std::max(tsF.max_npp, tsS.max_npp));
                    if (c2.smem > PT_SMEM_MAX) throw std::runtime_error("valence_b200: orbital basis set too large for one pair-group tile");
                    CK(cudaFuncSetAttribute(k_ptile<PART_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2.smem));
                    fill_cfg(c2);
                    const long long nts = (long long)tls.size();
                    tiles.upload(tls, st); items.upload(its, st);
                    tileE.alloc((size_t)std::max<long long>(1, nts));
                    tileE.zero(st); counter.zero(st); counters.zero(st); pq_counters.zero(st);
                    CK(cudaMemcpyAsync(r1x.p, x.data(), nso * sizeof(double), cudaMemcpyHostToDevice, st));
                    CK(cudaMemcpyAsync(r1y.p, y.data(), nso * sizeof(double), cudaMemcpyHostToDevice, st));
                    A.tiles = reinterpret_cast<const int2*>(tiles.p); A.ntiles = (int)nts; A.items = reinterpret_cast<const int4*>(items.p); A.nitems = (int)its.size();
                    A.gbuf = nullptr; A.gslot_base = 0; A.mode = 1; A.tau = tau_energy;
                    A.Pa = Qd.p; A.Pb = Rd.p; A.c0 = Dd; A.cof = nullptr; A.ndp = 0; A.r1 = 1; A.r1x = r1x.p; A.r1y = r1y.p; A.r1sigma = sigma;
                    A.nsh_bra = nsh_bra.p; A.nsh_ket = nsh_ket.p; A.tileE = tileE.p;
                    CK(cudaEventRecord(ev2, st));
                    if (mine > 0) {
                        k_ptile<PART_ALL><<<std::max(1, std::min(nsm, (int)its.size())), pt_threads(PART_ALL), c2.smem, st>>>(A);
                        CK(cudaGetLastError());
                        launches++;
                        acc->tile_launches++;
                    }
                    CK(cudaEventRecord(ev3, st));
                    k_sum<<<1, 1024, 0, st>>>(tileE.p, nts, accum.p);
                    CK(cudaGetLastError());
                    launches++;
                    double e2s = 0.0;
                    CK(cudaMemcpyAsync(&e2s, accum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
                    std::vector<unsigned long long> c;
                    counters.download(c, st);
                    add_pq();
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, ev2, ev3));
                    acc->t_tiles += ms;
                    acc->n_tiles += nts + nvt;
                    for (int i = 0; i < CNT_N; ++i) acc->counters[i] += (long long)(c[i] + cfree[i]);
                    const double num = Dd * (sigma * E2_0 + xFy) + e2s + (rank == 0 ? Dd * (sigma * E1_0 + xHy) : 0.0);
                    (*ham)[(size_t)jb * norbas + ib] += num;
                    (*ovl)[(size_t)jb * norbas + ib] += Dd * (sigma * S0 + xSy) / (double)nelec;
                    if (dbg_time) std::printf("[fo] ib %d jb %d: sigma %.3e, subject tiles %lld, kernel %.1f ms, host %.1f ms\n", ib + 1, jb + 1, sigma, nts, ms, now_ms() - tp0 - ms);
                }
            A.r1 = 0;
            return true;
        }
    }
    // ---- the (ib,jb) loop -------------------------------------------------------------------------------
    std::vector<int> avec, bvec;
    for (int ib = 0; ib < norbas; ++ib)
        for (int jb = 0; jb <= ib; ++jb) {
            double tp0 = now_ms();
            substitute(ib, jb);
            EnergyResult r;
            bool fast = false;
            double c0 = 1.0;
            int ndp = 0;
            cofactor_stage(in, w2, false, &r, &fast, &c0, &ndp);
            double tp1 = now_ms();
            TileOpts so;
            so.isolate = es; so.only_subject = true; so.wcut = wall;
            TileSetup tsS;
            build_tiles(in, bas, w2, orbs2e, tau_diag, true, &tsS, so);
            const int nS = (int)tsS.pgs.size();
            if (tsS.max_np > 32) throw std::runtime_error("valence_b200: pair group too large");
            hp.resize(nfree);
            for (PGDesc pg : tsS.pgs) {
                pg.pair_beg += (int)fPairs;
                pg.d_off += (long long)fD;
                for (int t = 0; t <= NPTYPE; ++t) { pg.pp_beg[t] += (int)fPps; pg.sp_beg[t] += (int)fSps; }
                hp.push_back(pg);
            }
            for (SPRec& sr : tsS.sps) sr.pp_beg += (int)fPps;
            set_smax(hp, tsS.pg_pairs, fPairs, nfree);
            {
                std::vector<PGDesc> tail(hp.begin() + nfree, hp.end());
                pgs.upload_at(nfree, tail, st);
                pg_pairs.upload_at(2 * fPairs, tsS.pg_pairs, st);
                sps.upload_at(fSps, tsS.sps, st);
                pps.upload_at(fPps, tsS.pps, st);
                pps_flat.upload_at(fPps, tsS.pps_flat, st);
                dmat.upload_at(fD, tsS.dmat, st);
            }
            std::vector<int> nshb(nso), nshk(nso);
            for (int s = 0; s < nso; ++s) {
                nshb[s] = (int)orbs2e[w2.bra[w2.slot(s, 0)]].sh.size();
                nshk[s] = (int)orbs2e[w2.ket[w2.slot(s, 0)]].sh.size();
            }
            nsh_bra.upload(nshb, st); nsh_ket.upload(nshk, st);
            // subject tiles: a subject pair group against everything
            avec.clear(); bvec.clear();
            for (int x = nfree; x < nfree + nS; ++x) avec.push_back(x);
            for (int x = 0; x < nfree + nS; ++x) bvec.push_back(x);
            std::vector<TilePair> tls;
            make_tile_list(hp, avec, bvec, itol, &tls, &runs);
            std::vector<WorkItem> its;
            long long mine = 0;
            make_items(runs, (long long)tls.size(), nsm, rank, nranks, &its, &mine);
            const PtCfg c2 = pt_cfg(std::max(tsF.max_ne, tsS.max_ne), std::max(tsF.max_np, tsS.max_np), std::max(tsF.max_nsp, tsS.max_nsp),
                                    std::max(tsF.max_npp, tsS.max_npp));
            if (c2.smem > PT_SMEM_MAX) throw std::runtime_error("valence_b200: orbital basis set too large for one pair-group tile");
            if (c2.g_cap != cfg.g_cap) throw std::runtime_error("valence_b200: first_order cache: tile size changed");
            CK(cudaFuncSetAttribute(k_ptile<PART_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c2.smem));
            fill_cfg(c2);
            const long long nts = (long long)tls.size();
            tiles.upload(tls, st); items.upload(its, st);
            tileE.alloc((size_t)std::max<long long>(1, nts + nvt));
            tileE.zero(st); counter.zero(st); counters.zero(st); pq_counters.zero(st);
            A.tiles = reinterpret_cast<const int2*>(tiles.p); A.ntiles = (int)nts; A.items = reinterpret_cast<const int4*>(items.p); A.nitems = (int)its.size();
            A.gbuf = nullptr; A.gslot_base = 0; A.mode = 1; A.tau = tau_energy;
            A.Pa = Pa.p; A.Pb = Pb.p; A.c0 = fast ? c0 : 1.0; A.cof = cof.p; A.ndp = ndp; A.cof_stride = (long long)cof_stride(nso);
            A.nsh_bra = nsh_bra.p; A.nsh_ket = nsh_ket.p; A.tileE = tileE.p;
            CK(cudaEventRecord(ev2, st));
            if (mine > 0) {
                k_ptile<PART_ALL><<<std::max(1, std::min(nsm, (int)its.size())), pt_threads(PART_ALL), c2.smem, st>>>(A);
                CK(cudaGetLastError());
                launches++;
                acc->tile_launches++;
            }
            if (dbg_time) CK(cudaEventRecord(ev1, st));
            if (nvt > 0) {
                const int grid = (int)std::min<long long>((nvt + CT_THREADS / 32 - 1) / (CT_THREADS / 32), (long long)nsm * 16);
                k_contract<<<grid, CT_THREADS, 0, st>>>(A, vtiles.p, nvt, fo_perm.p, gcache.p, nts);
                CK(cudaGetLastError());
                launches++;
            }
            CK(cudaEventRecord(ev3, st));
            k_sum<<<1, 1024, 0, st>>>(tileE.p, nts + nvt, accum.p);
            CK(cudaGetLastError());
            launches++;
            double e2 = 0.0;
            CK(cudaMemcpyAsync(&e2, accum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            std::vector<unsigned long long> c;
            counters.download(c, st);
            add_pq();
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev2, ev3));
            acc->t_tiles += ms;
            acc->n_tiles += nts + nvt;
            for (int i = 0; i < CNT_N; ++i) acc->counters[i] += (long long)c[i];
            (*ham)[(size_t)jb * norbas + ib] += (rank == 0 ? e1 : 0.0) + e2;
            (*ovl)[(size_t)jb * norbas + ib] += wfnorm;
            if (dbg_time) {
                float ms1 = 0.f;
                CK(cudaEventElapsedTime(&ms1, ev2, ev1));
                std::printf("[fo] ib %d jb %d: cofactors %.1f ms, subject tables+tiles %.1f ms (%lld tiles), subject kernel %.1f ms, contraction %.1f ms, entries %llu\n",
                            ib + 1, jb + 1, tp1 - tp0, now_ms() - tp1 - ms, nts, ms1, ms - ms1, c[CNT_ENTRIES]);
            }
        }
    return true;
}

// first_order_opt matrices (valence.F90:527-764) for 0-based orbital `iorb`:
//   ham(ib,jb) = <Psi[slot e <- chi_ib] | H_el | Psi[slot e <- chi_jb]>, ovl likewise (not divided,
//   no nuclear repulsion); column-major n x n in ham/ovl (n = # expansion terms of the orbital).
int Engine::first_order(int iorb, std::vector<double>* ham, std::vector<double>* ovl, EnergyResult* stats, int rank, int nranks)
{
    Impl& I = *impl_;
    CK(cudaSetDevice(I.device));
    const Input& in = in_;
    if (iorb < 0 || iorb >= in.norbs() - in.ndf) throw std::runtime_error("first_order: orbital index out of range");
    const bool internal = rank < 0;            // the engine's own communicator shards the tiles and sums ham
    if (internal) { rank = I.crank(); nranks = I.csize(); }
    I.fo_collective = internal && nranks > 1;
    EnergyResult acc;
    I.launches = 0;
    I.t_begin = now_ms();
    g_h2d_bytes = 0; g_d2h_bytes = 0;
    I.prepare(in, iorb);
    const int norbs = in.norbs(), nnd0 = in.nnd(), ndocc = in.ndocc;
    const int norbas = (int)in.orbitals[iorb].xp.size();
    // wave-function lists (valence.F90:557-605)
    Wavefunction wf;
    int eslot;   // 0-based electron slot that is substituted
    if (iorb < nnd0) {
        wf = default_wavefunction(in);
        eslot = iorb;
    } else {
        wf.nnd = nnd0 + 2;
        wf.nso = nnd0 + ndocc + 1;
        for (int i = 0; i < nnd0; ++i) { wf.bra.push_back(i); wf.ket.push_back(i); }
        for (int k = 0; k < 2; ++k) { wf.bra.push_back(iorb); wf.ket.push_back(iorb); }
        for (int d = nnd0; d < nnd0 + ndocc; ++d)
            if (d != iorb) for (int k = 0; k < 2; ++k) { wf.bra.push_back(d); wf.ket.push_back(d); }
        eslot = nnd0;
    }
    wf.sym = true;
    wf.subject = -1;
    // Schwarz table of the unsubstituted lists (valence.F90:666-667); the (ib,jb) loop reuses it
    std::vector<double> sch;
    bool sch_cached = I.sch_cache_key != 0 && I.sch_cache_n == nnd0 + ndocc && I.sch_cache_key == I.state_key(in) &&
                      I.sch_cache.size() == (size_t)I.sch_cache_n * I.sch_cache_n;
    if (const char* e = std::getenv("VB_FO_SCH_CACHE")) sch_cached = sch_cached && std::atoi(e) != 0;
    if (sch_cached) {
        // same geometry, weights and screens as the last guess energy: its Schwarz table, re-indexed by the entries of these lists
        const int n0 = I.sch_cache_n, n1 = wf.nso;
        sch.resize((size_t)n1 * n1);
        for (int a = 0; a < n1; ++a)
            for (int b = 0; b < n1; ++b) sch[(size_t)a * n1 + b] = I.sch_cache[(size_t)wf.bra[wf.slot(a, 0)] * n0 + wf.bra[wf.slot(b, 0)]];
    } else {
        I.evaluate(in, wf, nullptr, true, 0, 1, &acc, &sch);
    }
    ham->assign((size_t)norbas * norbas, 0.0);
    ovl->assign((size_t)norbas * norbas, 0.0);
    const int npass = (in.nunpd > 0 && iorb >= nnd0) ? 2 : 1;   // spin average, valence.F90:709-749
    for (int pass = 0; pass < npass; ++pass) {
        Wavefunction w2 = wf;
        const int es = eslot + pass;              // beta-spin position in the second pass
        w2.sym = false;
        w2.subject = es;                          // slot index == entry index for the non-DOCC part of the list
        // integrals that do not involve the substituted entry are generated once and kept in HBM
        if (I.first_order_cached(in, wf, es, iorb, sch, rank, nranks, ham, ovl, &acc)) continue;
        for (int ib = 0; ib < norbas; ++ib) {
            int idf = in.orbitals[iorb].xp[ib];
            w2.bra[es] = idf < 1 ? norbs + idf - 1 : norbs + ib;
            for (int jb = 0; jb <= ib; ++jb) {
                int jdf = in.orbitals[iorb].xp[jb];
                w2.ket[es] = jdf < 1 ? norbs + jdf - 1 : norbs + jb;
                EnergyResult r;
                I.evaluate(in, w2, &sch, false, rank, nranks, &r, nullptr);
                std::vector<double> a(1 + CNT_N);
                CK(cudaMemcpyAsync(a.data(), I.accum.p, a.size() * sizeof(double), cudaMemcpyDeviceToHost, I.st));
                CK(cudaStreamSynchronize(I.st));
                const double num = (rank == 0 ? r.e1 : 0.0) + a[0];
                if (std::getenv("VB_DEBUG_FO")) {
                    std::printf("---- end of evaluation ib %d jb %d\n", ib + 1, jb + 1);
                    Wavefunction w3 = w2;
                    std::swap(w3.bra[es], w3.ket[es]);
                    EnergyResult r3;
                    I.evaluate(in, w3, &sch, false, 0, 1, &r3, nullptr);
                    std::vector<double> a3(1);
                    CK(cudaMemcpyAsync(a3.data(), I.accum.p, sizeof(double), cudaMemcpyDeviceToHost, I.st));
                    CK(cudaStreamSynchronize(I.st));
                    std::printf("FO ib %d jb %d : e1 %.10f e2 %.10f N %.10f | swapped e1 %.10f e2 %.10f N %.10f\n", ib + 1, jb + 1, r.e1, a[0], r.wfnorm, r3.e1, a3[0], r3.wfnorm);
                }
                (*ham)[(size_t)jb * norbas + ib] += num;
                (*ovl)[(size_t)jb * norbas + ib] += r.wfnorm;
                acc.t_tiles += r.t_tiles; acc.flops_model += r.flops_model; acc.n_prim_quartets += r.n_prim_quartets;
                acc.n_tiles += r.n_tiles; acc.tile_launches += r.tile_launches;
                for (int i = 0; i < CNT_N; ++i) acc.counters[i] += (long long)(a[1 + i] + 0.5);
            }
        }
    }
    if (internal && nranks > 1) {              // xm_equalize(ham), valence.F90:763
        I.one_e.alloc(ham->size());
        CK(cudaMemcpyAsync(I.one_e.p, ham->data(), ham->size() * sizeof(double), cudaMemcpyHostToDevice, I.st));
        I.comm->allreduce_sum(I.one_e.p, ham->size(), I.st);
        CK(cudaMemcpyAsync(ham->data(), I.one_e.p, ham->size() * sizeof(double), cudaMemcpyDeviceToHost, I.st));
        CK(cudaStreamSynchronize(I.st));
    }
    for (int i = 0; i < norbas; ++i)
        for (int j = 0; j < i; ++j) {
            (*ham)[(size_t)i * norbas + j] = (*ham)[(size_t)j * norbas + i];
            (*ovl)[(size_t)i * norbas + j] = (*ovl)[(size_t)j * norbas + i];
        }
    acc.enucrep = I.enuc;
    acc.launches = I.launches;
    acc.t_total = now_ms() - I.t_begin;
    acc.h2d_bytes = g_h2d_bytes; acc.d2h_bytes = g_d2h_bytes;
    if (stats) *stats = acc;
    return norbas;
}

namespace {

// Generalised symmetric eigenproblem A x = lambda B x (column-major, lower triangles used): Cholesky
// reduction + cyclic Jacobi; eigenvalues ascending, eigenvectors B-orthonormal.  Stands in for
// EISPACK rsg (/root/reference/src/rsg.F:1, called at valence.F90:771,912); returns 7n+1 when B is
// not positive definite, like rsg.
int gen_eig(int n, const std::vector<double>& A, const std::vector<double>& B, std::vector<double>* w, std::vector<double>* Z)
{
    std::vector<double> L((size_t)n * n, 0.0), C((size_t)n * n, 0.0), T((size_t)n * n, 0.0), V((size_t)n * n, 0.0);
    auto a = [&](int i, int j) { return i >= j ? A[(size_t)j * n + i] : A[(size_t)i * n + j]; };
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = B[(size_t)j * n + i];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            if (i == j) { if (s <= 0.0) return 7 * n + 1; L[i * n + i] = std::sqrt(s); }
            else L[i * n + j] = s / L[j * n + j];
        }
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double s = a(i, j);
            for (int k = 0; k < i; ++k) s -= L[i * n + k] * T[k * n + j];
            T[i * n + j] = s / L[i * n + i];
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = T[i * n + j];
            for (int k = 0; k < j; ++k) s -= C[i * n + k] * L[j * n + k];
            C[i * n + j] = s / L[j * n + j];
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) { double s = 0.5 * (C[i * n + j] + C[j * n + i]); C[i * n + j] = s; C[j * n + i] = s; }
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < n; ++i) { dg += C[i * n + i] * C[i * n + i]; for (int j = 0; j < i; ++j) off += C[i * n + j] * C[i * n + j]; }
        if (off <= 1e-34 * (dg + 1e-300)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = C[p * n + q];
                if (apq == 0.0) continue;
                double th = (C[q * n + q] - C[p * n + p]) / (2.0 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; ++k) { double x = C[k * n + p], y = C[k * n + q]; C[k * n + p] = cs * x - sn * y; C[k * n + q] = sn * x + cs * y; }
                for (int k = 0; k < n; ++k) { double x = C[p * n + k], y = C[q * n + k]; C[p * n + k] = cs * x - sn * y; C[q * n + k] = sn * x + cs * y; }
                for (int k = 0; k < n; ++k) { double x = V[k * n + p], y = V[k * n + q]; V[k * n + p] = cs * x - sn * y; V[k * n + q] = sn * x + cs * y; }
            }
    }
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](int x, int y) { return C[x * n + x] < C[y * n + y]; });
    w->assign(n, 0.0);
    Z->assign((size_t)n * n, 0.0);
    for (int e = 0; e < n; ++e) {
        (*w)[e] = C[ord[e] * n + ord[e]];
        for (int i = n - 1; i >= 0; --i) {
            double s = V[i * n + ord[e]];
            for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * (*Z)[(size_t)e * n + k];
            (*Z)[(size_t)e * n + i] = s / L[i * n + i];
        }
    }
    return 0;
}

}  // namespace

const std::vector<std::vector<double>>& Engine::weights() const { return impl_->coeff; }
std::vector<std::vector<double>> Engine::normalized_weights()
{
    Impl& I = *impl_;
    CK(cudaSetDevice(I.device));
    Comm* keep = I.comm.release();             // a local evaluation: not every rank writes files
    try { I.prepare(in_, -1); } catch (...) { I.comm.reset(keep); throw; }
    I.comm.reset(keep);
    return I.cnorm;
}
const std::vector<double>& Engine::coupling_weights() const { return impl_->coeff_sc; }

void Engine::run(RunResult* out, bool print)
{
    Impl& I = *impl_;
    const Input& in = in_;
    const double tokcal = 627.509469;
    *out = RunResult();
    print = print && I.crank() == 0;           // xm_print: rank 0 only (xm_module.F90:356-458)
    EnergyResult r;
    energy(&r);
    out->enucrep = r.enucrep;
    out->guess_energy = r.energy;
    double energy_now = r.energy;
    if (print) {
        if (in.natom > 1) std::printf(" %-32s  %24.16f\n", "nuclear repulsion", r.enucrep);   // valence.F90:92
        std::printf(" %-32s  %24.16f\n", "guess energy", r.energy);                          // valence.F90:191
        std::fflush(stdout);
    }
    out->total_energy = energy_now;
    if (on_save && I.crank() == 0) on_save(energy_now, false);                               // valence.F90:192
    if (in.max_iter <= 0) return;
    if (in.ptbnmax > 0.0 && in.nxorb == 0)
        throw std::runtime_error("direct energy minimisation (demgs_opt) is not supported (under development in the reference, README.md:108)");
    if (print) {
        std::printf(" %-71s\n\n", "orbital optimization");
        std::printf(" %-71s\n\n", "(full) first-order method");
        std::printf(" %-71s\n", "cycle  orbital   relaxation(kCal)   (..per orb.)/tol");
    }
    const int nsc = in.nspinc;
    double etol = 0.0, eprev = 0.0, eprv_sc = 0.0, eprv_orb = 0.0, cumulx = 0.0;
    int num_iter = 0;
    for (int ntol = in.ntol_e_min; ntol <= in.ntol_e_max; ++ntol) {
        etol = std::pow(10.0, -ntol);
        bool finished = false;
        while (!finished) {
            eprv_sc = energy_now;
            bool orbconv = false;
            while (!orbconv && num_iter < in.max_iter) {
                eprv_orb = energy_now;
                for (int iset = 0; iset < in.nset; ++iset) {
                    bool setconv = false;
                    while (!setconv && num_iter < in.max_iter) {
                        const double eprv_set = energy_now;
                        ++num_iter;
                        for (int iorb = in.orbset[2 * iset]; iorb <= in.orbset[2 * iset + 1]; ++iorb) {
                            eprev = energy_now;
                            std::vector<double> ham, ovl, w, Z;
                            EnergyResult st;
                            const int n = first_order(iorb - 1, &ham, &ovl, &st);
                            if (gen_eig(n, ham, ovl, &w, &Z) != 0) throw std::runtime_error("solver failed");   // valence.F90:786
                            int rootsel = 0;
                            for (int k = 0; k < in.nxorb; ++k) if (in.xorb[k] == iorb) rootsel = in.root[k];
                            energy_now = w[rootsel] + st.enucrep;                                          // valence.F90:791-795
                            for (int i = 0; i < n; ++i) I.coeff[iorb - 1][i] = Z[(size_t)rootsel * n + i];  // valence.F90:797-803
                            const double relaxn = (energy_now - eprev) * tokcal;
                            cumulx += relaxn;
                            if (print) { std::printf("%5d  %4d      %14.6f      %12.4E\n", num_iter, iorb, cumulx, relaxn / etol); std::fflush(stdout); }
                            if (on_save && I.crank() == 0) on_save(energy_now, false);               // valence.F90:2823
                        }
                        setconv = std::fabs(energy_now - eprv_set) * tokcal < etol;
                    }
                }
                orbconv = std::fabs(energy_now - eprv_orb) * tokcal < etol;
                if (in.nset == 1) { orbconv = true; eprv_orb = energy_now; }
            }
            if (nsc > 1) {
                // spin_opt (valence.F90:850-936): Hamiltonian / overlap between spin couplings
                if (print) std::printf(" %-71s\n", "spin optimization");
                ++num_iter;
                eprev = energy_now;
                std::vector<double> ham((size_t)nsc * nsc, 0.0), ovl((size_t)nsc * nsc, 0.0), w, Z;
                Wavefunction wf;
                wf.nnd = in.nnd(); wf.nso = wf.nnd + in.ndocc; wf.sym = false; wf.subject = -1;
                for (int i = 0; i < wf.nnd; ++i) { wf.bra.push_back(i); wf.ket.push_back(i); }
                for (int d = 0; d < in.ndocc; ++d) for (int k = 0; k < 2; ++k) { wf.bra.push_back(wf.nnd + d); wf.ket.push_back(wf.nnd + d); }
                I.launches = 0;
                I.prepare(in, -1);
                double enuc = I.enuc;
                for (int isc = 0; isc < nsc; ++isc)
                    for (int jsc = 0; jsc <= isc; ++jsc) {
                        I.only_isc = isc; I.only_jsc = jsc;
                        EnergyResult e2;
                        I.evaluate(in, wf, nullptr, false, 0, 1, &e2, nullptr);
                        std::vector<double> acc(1);
                        CK(cudaMemcpyAsync(acc.data(), I.accum.p, sizeof(double), cudaMemcpyDeviceToHost, I.st));
                        CK(cudaStreamSynchronize(I.st));
                        ham[(size_t)jsc * nsc + isc] = e2.e1 + acc[0];
                        ovl[(size_t)jsc * nsc + isc] = e2.wfnorm;
                    }
                I.only_isc = -1; I.only_jsc = -1;
                if (gen_eig(nsc, ham, ovl, &w, &Z) != 0) throw std::runtime_error("solver failed");
                energy_now = w[0] + enuc;
                for (int i = 0; i < nsc; ++i) I.coeff_sc[i] = Z[i];
                const double relaxn = (energy_now - eprev) * tokcal;
                cumulx += relaxn;
                if (print) { std::printf("%5d  %4d      %14.6f      %12.4E\n", num_iter, 0, cumulx, relaxn / etol); std::fflush(stdout); }
                if (on_save && I.crank() == 0) on_save(energy_now, false);                           // valence.F90:2859
                finished = std::fabs(energy_now - eprv_sc) * tokcal < etol;
            } else {
                finished = true;
            }
            if (num_iter >= in.max_iter) finished = true;
        }
    }
    eprev = nsc > 1 ? eprv_sc : eprv_orb;
    out->iterations = num_iter;
    out->total_energy = energy_now;
    if (num_iter >= in.max_iter) {
        if (print) std::printf("\n %-71s\n\n", "reached maximum number of iterations");
    } else if (std::fabs(energy_now - eprev) * tokcal < etol) {
        out->converged = 1;
        if (on_save && I.crank() == 0) on_save(energy_now, true);                                    // valence.F90:2882
        if (print) {
            std::printf("\n %-71s\n\n", "calculation converged");
            std::printf(" %-32s  %24.16f\n", "total energy", energy_now);
        }
    }
    if (print) std::fflush(stdout);
}

// FP64 FMA peak of this GPU, measured: 8 independent DFMA chains per thread, full occupancy.
__global__ void k_dfma_peak(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double measure_fp64_peak_tflops(int device)
{
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(d, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return best;
}

void Engine::energy_finish(EnergyResult* out)
{
    Impl& I = *impl_;
    CK(cudaSetDevice(I.device));
    std::vector<double> acc(2 + CNT_N);
    CK(cudaMemcpyAsync(acc.data(), I.accum.p, acc.size() * sizeof(double), cudaMemcpyDeviceToHost, I.st));
    CK(cudaStreamSynchronize(I.st));
    out->e2 = acc[0] + acc[1 + CNT_N];
    for (int i = 0; i < CNT_N; ++i) out->counters[i] = (long long)(acc[1 + i] + 0.5);
    out->ref_shell_quartets = out->counters[CNT_SHELLQ];
    out->numerator = I.e1 + out->e2;
    out->energy = out->numerator / I.wfnorm + I.enuc;     // valence.F90:344
    if (I.c0_keep != 0.0 && std::isfinite(I.c0_keep)) {
        typedef long double ld;     // x87 extended (64-bit mantissa) on the x86 hosts this runs on; plain double elsewhere
        const ld e1 = (ld)I.e1u[0] + (ld)I.e1u[1], wn = ((ld)I.wnu[0] + (ld)I.wnu[1]) / (ld)I.nelec_keep;
        const ld e2 = ((ld)acc[0] + (ld)acc[1 + CNT_N]) / (ld)I.c0_keep;
        const ld E = (e1 + e2) / wn + (ld)I.enuc;
        if (std::getenv("VB_DEBUG_PARTS"))
            std::fprintf(stderr, "[parts] E1/c0 %.20Lg  E2/c0 %.20Lg  N1/nelec - 1 %.3Le  c0 %.17g  E %.20Lg  (plain formula %.17g)\n", e1, e2, wn - 1.0L, I.c0_keep, E, out->energy);
        out->energy = (double)E;
    }
    out->launches = I.launches;
    out->h2d_bytes = g_h2d_bytes + 8 * CNT_N; out->d2h_bytes = g_d2h_bytes + 8 * (1 + CNT_N);
    out->t_total = now_ms() - I.t_begin;
}

}  // namespace vb
