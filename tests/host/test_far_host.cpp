// Host-side check of the far-field form (vb_far.cuh) against the general Obara-Saika path (vb_eri.cuh) on random
// contracted shell pairs at all separations, for the four light classes.  Built and run by tests/test_host_math.py; no GPU needed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include "../../valence_b200/csrc/vb_far.cuh"
using namespace vb;

struct Sh { int l; std::vector<double> ex, co; double r[3]; };
static double urand() { return rand() / (double)RAND_MAX; }
static Sh mk(int l, int n, const double* c, double spread) {
    Sh s; s.l = l;
    for (int i = 0; i < n; ++i) { s.ex.push_back(0.15 * std::pow(4.3, i) * (0.7 + 0.6 * urand())); s.co.push_back(0.3 + urand()); }
    for (int d = 0; d < 3; ++d) s.r[d] = c[d] + spread * (urand() - 0.5);
    return s;
}
static std::vector<PrimPair> pairs(const Sh& A, const Sh& B) {   // A carries the angular momentum
    std::vector<PrimPair> v; double AB2 = 0; for (int d = 0; d < 3; ++d) AB2 += (A.r[d]-B.r[d])*(A.r[d]-B.r[d]);
    for (size_t i = 0; i < A.ex.size(); ++i) for (size_t j = 0; j < B.ex.size(); ++j) {
        double a = A.ex[i], b = B.ex[j], p = a + b; PrimPair pp;
        pp.Px = (a*A.r[0]+b*B.r[0])/p; pp.Py = (a*A.r[1]+b*B.r[1])/p; pp.Pz = (a*A.r[2]+b*B.r[2])/p; pp.p = p;
        pp.ip = 1.0/p; pp.Kp = A.co[i]*B.co[j]*std::exp(-a*b/p*AB2)*std::sqrt(2.0)*std::pow(PI,1.25)/p; pp.w = 0.5 + urand(); pp.wseg = 1.0; pp.eoff = 0; pp.pad = 0;
        pp.PAx = b/p*(B.r[0]-A.r[0]); pp.PAy = b/p*(B.r[1]-A.r[1]); pp.PAz = b/p*(B.r[2]-A.r[2]); pp.Px = A.r[0]+pp.PAx; pp.Py = A.r[1]+pp.PAy; pp.Pz = A.r[2]+pp.PAz; v.push_back(pp);
    }
    return v;
}
template <int TB, int TK>
static double one_case(const double* tab, double sep, bool same_centre, long* nfar, long* nnear)
{
    constexpr int LA = pt_la(TB), EA = pt_E(TB), LC = pt_la(TK), EC = pt_E(TK), NE = pt_ne(TB), NF = pt_ne(TK);
    const double c0[3] = {0.3, -0.2, 0.1}, dir[3] = {urand() - 0.5, urand() - 0.5, urand() - 0.5};
    const double nrm = std::sqrt(dir[0]*dir[0] + dir[1]*dir[1] + dir[2]*dir[2]);
    const double c1[3] = {c0[0] + sep * dir[0] / nrm, c0[1] + sep * dir[1] / nrm, c0[2] + sep * dir[2] / nrm};
    Sh A = mk(LA, 1 + rand() % 3, c0, 3.0), B = mk(pt_lb(TB), 1 + rand() % 3, c0, 3.0), C = mk(LC, 1 + rand() % 3, c1, 3.0), D = mk(pt_lb(TK), 1 + rand() % 3, c1, 3.0);
    if (same_centre) for (int d = 0; d < 3; ++d) B.r[d] = A.r[d];
    std::vector<PrimPair> P = pairs(A, B), Q = pairs(C, D);
    double ref[NE * NF] = {0}, got[NE * NF] = {0};
    for (auto& b : Q) {
        const FarKet k = far_ket(far_make_prim(b));
        FarSums<TB, TK> S; S.clear();
        for (auto& a : P) {
            QuartetGeom g; double T, pref; quartet_geom(a, b, g, T, pref);
            FarSums<TB, TK> S0 = S;
            const bool far = far_quartet<TB, TK>(far_make_prim(a), k, true, S);
            if (far && T < FAR_T) { std::printf("far test accepted T = %g\n", T); std::exit(1); }
            if (!far && T > FAR_T * (1.0 + 1e-9)) { std::printf("far test rejected T = %g\n", T); std::exit(1); }
            if (!far) { for (int i = 0; i < FarSums<TB, TK>::N; ++i) if (S.v[i] != S0.v[i]) { std::printf("near quartet changed the sums\n"); std::exit(1); } ++*nnear; continue; }
            ++*nfar;
            // a quartet that is not taken leaves the sums alone
            FarSums<TB, TK> S1 = S; far_quartet<TB, TK>(far_make_prim(a), k, false, S1);
            for (int i = 0; i < FarSums<TB, TK>::N; ++i) if (S.v[i] != S1.v[i]) { std::printf("masked quartet changed the sums\n"); std::exit(1); }
            double F[EA+EC+1]; boys<EA+EC>(tab, T, F); for (int m = 0; m <= EA+EC; ++m) F[m] *= pref;
            if constexpr (EA + EC == 0) ref[0] += F[0]; else vrr_unrolled<LA,EA,LC,EC>(g, F, ref);
        }
        far_finish<TB, TK>(S, k, got);
    }
    double mx = 0, dev = 0;
    for (int i = 0; i < NE * NF; ++i) { mx = std::fmax(mx, std::fabs(ref[i])); dev = std::fmax(dev, std::fabs(ref[i] - got[i])); }
    return mx > 0 ? dev / mx : 0.0;
}
int main() {
    std::vector<double> tab((size_t)BOYS_ROWS*BOYS_COLS); boys_make_table(tab.data());
    srand(7);
    double worst[4] = {0, 0, 0, 0};
    long nfar[4] = {0, 0, 0, 0}, nnear[4] = {0, 0, 0, 0};
    for (int it = 0; it < 4000; ++it) {
        const double sep = 1.0 + 40.0 * urand();
        const bool sc = it % 5 == 0;
        worst[0] = std::fmax(worst[0], one_case<0, 0>(tab.data(), sep, sc, &nfar[0], &nnear[0]));
        worst[1] = std::fmax(worst[1], one_case<1, 0>(tab.data(), sep, sc, &nfar[1], &nnear[1]));
        worst[2] = std::fmax(worst[2], one_case<0, 1>(tab.data(), sep, sc, &nfar[2], &nnear[2]));
        worst[3] = std::fmax(worst[3], one_case<1, 1>(tab.data(), sep, sc, &nfar[3], &nnear[3]));
    }
    // the integer weight test never drops what the exact comparison keeps
    for (int it = 0; it < 200000; ++it) {
        const double wa = std::pow(10.0, -12.0 * urand()), wk = std::pow(10.0, -12.0 * urand()), tau = std::pow(10.0, -8.0 - 14.0 * urand());
        const bool exact = wa * wk >= tau, cons = (int)(far_hi32(wa) - (far_lthr(tau) - far_hi32(wk))) >= 0;
        if (exact && !cons) { std::printf("weight test dropped a kept quartet: %g %g %g\n", wa, wk, tau); return 1; }
        if (cons && wa * wk < 0.8 * tau) { std::printf("weight test far too loose: %g %g %g\n", wa, wk, tau); return 1; }
    }
    const char* nm[4] = {"(ss|ss)", "(ps|ss)", "(ss|ps)", "(ps|ps)"};
    int bad = 0;
    for (int c = 0; c < 4; ++c) {
        std::printf("far %s: %ld far quartets, %ld near, max relative deviation of the block %.2e\n", nm[c], nfar[c], nnear[c], worst[c]);
        if (!(worst[c] < 5e-15) || nfar[c] < 10000 || nnear[c] < 1000) bad = 1;
    }
    std::printf(bad ? "FAIL\n" : "PASS\n");
    return bad;
}
