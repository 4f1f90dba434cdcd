// Host-side check of the product's Obara-Saika VRR + HRR folding against the
// oracle's McMurchie-Davidson ERIs (independent algorithms).  Built and run by
// tests/test_host_math.py; no GPU needed (vb_eri.cuh is __host__ __device__).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../valence_b200/csrc/vb_eri.cuh"
extern "C" {
#include "../../oracle/vo_internal.h"
}
using namespace vb;

struct Sh { int l; std::vector<double> ex, co; double r[3]; };
static double urand() { return rand() / (double)RAND_MAX; }
static Sh mk(int l, int n) {
    Sh s; s.l = l;
    for (int i = 0; i < n; ++i) { s.ex.push_back(0.15 * std::pow(4.3, i) * (0.7 + 0.6 * urand())); s.co.push_back(0.3 + urand()); }
    for (int d = 0; d < 3; ++d) s.r[d] = 2.5 * (urand() - 0.5);
    return s;
}
static std::vector<PrimPair> pairs(const Sh& A, const Sh& B) {
    std::vector<PrimPair> v; double AB2 = 0; for (int d = 0; d < 3; ++d) AB2 += (A.r[d]-B.r[d])*(A.r[d]-B.r[d]);
    for (size_t i = 0; i < A.ex.size(); ++i) for (size_t j = 0; j < B.ex.size(); ++j) {
        double a = A.ex[i], b = B.ex[j], p = a + b; PrimPair pp;
        pp.Px = (a*A.r[0]+b*B.r[0])/p; pp.Py = (a*A.r[1]+b*B.r[1])/p; pp.Pz = (a*A.r[2]+b*B.r[2])/p; pp.p = p;
        pp.ip = 1.0/p; pp.Kp = A.co[i]*B.co[j]*std::exp(-a*b/p*AB2)*std::sqrt(2.0)*std::pow(PI,1.25)/p; pp.w = 1.0; pp.wseg = 1.0; pp.eoff = 0; pp.pad = 0;
        pp.PAx = pp.Px-A.r[0]; pp.PAy = pp.Py-A.r[1]; pp.PAz = pp.Pz-A.r[2]; v.push_back(pp);
    }
    return v;
}
static double binom(int n, int k) { double r = 1; for (int i = 0; i < k; ++i) r = r*(n-i)/(i+1); return r; }

template <int LA, int EA, int LC, int EC>
static void run_unrolled(const std::vector<PrimPair>& P, const std::vector<PrimPair>& Q, const double* tab, double* acc) {
    for (auto& a : P) for (auto& b : Q) {
        QuartetGeom g; double T, pref; quartet_geom(a, b, g, T, pref);
        double F[EA+EC+1]; boys<EA+EC>(tab, T, F); for (int m = 0; m <= EA+EC; ++m) F[m] *= pref;
        vrr_unrolled<LA,EA,LC,EC>(g, F, acc);
    }
}
template <int LA, int EA, int LC, int EC>
static void run_unrolled_small(const std::vector<PrimPair>& P, const std::vector<PrimPair>& Q, const double* tabs, double* acc) {
    for (auto& a : P) for (auto& b : Q) {   // what k_ptile executes: compact Boys table (boys_s) + unrolled VRR
        QuartetGeom g; double T, pref; quartet_geom(a, b, g, T, pref);
        double F[EA+EC+1]; boys_s<EA+EC>(tabs, T, F); for (int m = 0; m <= EA+EC; ++m) F[m] *= pref;
        vrr_unrolled<LA,EA,LC,EC>(g, F, acc);
    }
}
static void run_generic(int LA, int EA, int LC, int EC, const std::vector<PrimPair>& P, const std::vector<PrimPair>& Q, const double* tab, double* acc) {
    std::vector<double> scratch(GEN_SCRATCH);
    for (auto& a : P) for (auto& b : Q) {
        QuartetGeom g; double T, pref; quartet_geom(a, b, g, T, pref);
        double F[MTOP+1]; boys_rt(EA+EC, tab, T, F); for (int m = 0; m <= EA+EC; ++m) F[m] *= pref;
        vrr_generic(LA,EA,LC,EC,g,F,scratch.data(),acc);
    }
}

int main() {
    std::vector<double> tab((size_t)BOYS_ROWS*BOYS_COLS); boys_make_table(tab.data());
    std::vector<double> tabs(BOYS_S_SIZE); boys_make_table_small(tabs.data());
    // Boys check
    double worst_b = 0;
    for (double T = 0; T < 90; T += 0.0173) { double F[MTOP+1], R[MTOP+1]; boys_rt(MTOP, tab.data(), T, F); boys_reference(MTOP, T, R);
        for (int m = 0; m <= MTOP; ++m) worst_b = std::fmax(worst_b, std::fabs(F[m]-R[m])/R[m]); }
    printf("boys max rel err %.3e\n", worst_b);
    srand(7);
    double worst = 0, worst_ug = 0;
    for (int trial = 0; trial < 400; ++trial) {
        int l[4]; for (int i = 0; i < 4; ++i) l[i] = rand() % 3;
        if (l[0] < l[1]) std::swap(l[0], l[1]);
        if (l[2] < l[3]) std::swap(l[2], l[3]);
        Sh A = mk(l[0], 1 + rand()%3), B = mk(l[1], 1 + rand()%3), C = mk(l[2], 1 + rand()%3), D = mk(l[3], 1 + rand()%2);
        if (trial % 5 == 0) for (int d = 0; d < 3; ++d) { B.r[d] = A.r[d]; }        // one-centre pair
        if (trial % 7 == 0) for (int d = 0; d < 3; ++d) { B.r[d] = A.r[d]; C.r[d] = A.r[d]; D.r[d] = A.r[d]; }
        if (trial % 11 == 0) for (int d = 0; d < 3; ++d) { C.r[d] += (trial % 22 == 0 ? 30.0 : 9.0); D.r[d] += (trial % 22 == 0 ? 30.0 : 9.0); }  // large T
        int EA = l[0]+l[1], EC = l[2]+l[3];
        int NE = ncum(EA)-coff(l[0]), NF = ncum(EC)-coff(l[2]);
        std::vector<double> acc(NE*NF, 0.0), acc2(NE*NF, 0.0);
        auto P = pairs(A,B), Q = pairs(C,D);
        run_generic(l[0],EA,l[2],EC,P,Q,tab.data(),acc.data());
        if (l[0] <= 1 && l[2] <= 1) {   // the s/p kernel's path (compact Boys table) vs loop-based generic
            std::vector<double> acc3(NE*NF, 0.0);
            int tb = ptype(l[0],l[1]), tk = ptype(l[2],l[3]);
#define CASES(TB,TK) if (tb==TB && tk==TK) run_unrolled_small<pt_la(TB),pt_E(TB),pt_la(TK),pt_E(TK)>(P,Q,tabs.data(),acc3.data());
            CASES(0,0) CASES(0,1) CASES(0,2) CASES(1,0) CASES(1,1) CASES(1,2) CASES(2,0) CASES(2,1) CASES(2,2)
            double mx = 0; for (int i = 0; i < NE*NF; ++i) mx = std::fmax(mx, std::fabs(acc[i]));
            for (int i = 0; i < NE*NF; ++i) worst_ug = std::fmax(worst_ug, std::fabs(acc[i]-acc3[i])/mx);
        }
        bool unr = l[0] <= 1 && l[2] <= 1;
        if (unr) {
            int tb = ptype(l[0],l[1]), tk = ptype(l[2],l[3]);
#define CASE(TB,TK) if (tb==TB && tk==TK) run_unrolled<pt_la(TB),pt_E(TB),pt_la(TK),pt_E(TK)>(P,Q,tab.data(),acc2.data());
            CASE(0,0) CASE(0,1) CASE(0,2) CASE(1,0) CASE(1,1) CASE(1,2) CASE(2,0) CASE(2,1) CASE(2,2)
            double mx = 0; for (int i = 0; i < NE*NF; ++i) mx = std::fmax(mx, std::fabs(acc[i]));
            for (int i = 0; i < NE*NF; ++i) worst_ug = std::fmax(worst_ug, std::fabs(acc[i]-acc2[i])/mx);
        }
        // explicit HRR on both sides, compare with oracle
        vo_shell a{A.l,(int)A.ex.size(),A.ex.data(),A.co.data(),{A.r[0],A.r[1],A.r[2]}}, b{B.l,(int)B.ex.size(),B.ex.data(),B.co.data(),{B.r[0],B.r[1],B.r[2]}},
                 c{C.l,(int)C.ex.size(),C.ex.data(),C.co.data(),{C.r[0],C.r[1],C.r[2]}}, d{D.l,(int)D.ex.size(),D.ex.data(),D.co.data(),{D.r[0],D.r[1],D.r[2]}};
        int na=ncart(A.l), nb=ncart(B.l), nc=ncart(C.l), nd=ncart(D.l);
        std::vector<double> ref((size_t)na*nb*nc*nd); vo_eri_block(&a,&b,&c,&d,ref.data());
        double AB[3]={A.r[0]-B.r[0],A.r[1]-B.r[1],A.r[2]-B.r[2]}, CD[3]={C.r[0]-D.r[0],C.r[1]-D.r[1],C.r[2]-D.r[2]};
        double scale = 0; for (double v : ref) scale = std::fmax(scale, std::fabs(v));
        for (int ia=0; ia<na; ++ia) for (int ib=0; ib<nb; ++ib) for (int ic=0; ic<nc; ++ic) for (int id=0; id<nd; ++id) {
            int ca=coff(A.l)+ia, cb=coff(B.l)+ib, cc=coff(C.l)+ic, cd=coff(D.l)+id;
            double val = 0;
            for (int kx=0;kx<=c_lx(cb);++kx) for (int ky=0;ky<=c_ly(cb);++ky) for (int kz=0;kz<=c_lz(cb);++kz) {
                double fb = binom(c_lx(cb),kx)*binom(c_ly(cb),ky)*binom(c_lz(cb),kz)*std::pow(AB[0],c_lx(cb)-kx)*std::pow(AB[1],c_ly(cb)-ky)*std::pow(AB[2],c_lz(cb)-kz);
                int e = cidx(c_lx(ca)+kx,c_ly(ca)+ky,c_lz(ca)+kz)-coff(A.l);
                for (int jx=0;jx<=c_lx(cd);++jx) for (int jy=0;jy<=c_ly(cd);++jy) for (int jz=0;jz<=c_lz(cd);++jz) {
                    double fd = binom(c_lx(cd),jx)*binom(c_ly(cd),jy)*binom(c_lz(cd),jz)*std::pow(CD[0],c_lx(cd)-jx)*std::pow(CD[1],c_ly(cd)-jy)*std::pow(CD[2],c_lz(cd)-jz);
                    int f = cidx(c_lx(cc)+jx,c_ly(cc)+jy,c_lz(cc)+jz)-coff(C.l);
                    val += fb*fd*acc[e*NF+f];
                }
            }
            double r = ref[((size_t)(ia*nb+ib)*nc+ic)*nd+id];
            worst = std::fmax(worst, std::fabs(val-r)/scale);
        }
    }
    printf("eri max err rel. to block max %.3e ; unrolled vs generic %.3e\n", worst, worst_ug);
    return (worst < 5e-13 && worst_ug < 1e-12 && worst_b < 1e-13) ? 0 : 1;
}
