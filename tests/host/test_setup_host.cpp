// Host-side check of the pair-group table build (vb_setup.cpp): the layout must not depend on the
// number of host threads (every rank of a multi-GPU run derives its tile list from it), the flat
// (magnitude-sorted) primitive lists must be permutations of the shell-pair-grouped ones, and the
// compact Boys table must reproduce the exact series.  Built and run by tests/test_host_math.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>
#include "../../valence_b200/csrc/vb_setup.h"
using namespace vb;

static TileSetup build(const Input& in, const char* threads, bool sym = true, const TileOpts& opts = TileOpts())
{
    setenv("VB_HOST_THREADS", threads, 1);
    std::vector<double> xyz(3 * in.natom);
    for (int i = 0; i < 3 * in.natom; ++i) xyz[i] = in.coords[i] * ANGS2BOHR;
    Basis bas = build_basis(in, xyz);
    std::vector<std::vector<double>> coeff;
    for (const OrbitalDef& o : in.orbitals) coeff.push_back(o.coeff);
    std::vector<ExpOrb> orbs;
    for (int o = 0; o < in.norbs(); ++o) orbs.push_back(expand_orbital(in, bas, coeff, o));
    Wavefunction wf;
    wf.nnd = in.nnd(); wf.nso = wf.nnd + in.ndocc;
    for (int i = 0; i < wf.nnd; ++i) { wf.bra.push_back(i); wf.ket.push_back(i); }
    for (int d = 0; d < in.ndocc; ++d) for (int k = 0; k < 2; ++k) { wf.bra.push_back(wf.nnd + d); wf.ket.push_back(wf.nnd + d); }
    wf.sym = sym;
    TileSetup ts;
    build_tiles(in, bas, wf, orbs, 1e-22, true, &ts, opts);
    return ts;
}

template <class T>
static bool same(const std::vector<T>& a, const std::vector<T>& b)
{
    return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0);
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    Input in = parse_input_file(argv[1]);
    TileSetup a = build(in, "1"), b = build(in, "5");
    int bad = 0;
    if (!same(a.pg_pairs, b.pg_pairs) || !same(a.sps, b.sps) || !same(a.pps, b.pps) || !same(a.pps_flat, b.pps_flat) ||
        !same(a.dmat, b.dmat) || !same(a.pgs, b.pgs)) { std::printf("layout depends on the thread count\n"); ++bad; }
    if (a.pps.size() != a.pps_flat.size()) { std::printf("flat list size\n"); ++bad; }
    for (const PGDesc& pg : a.pgs)
        for (int t = 0; t < NPTYPE; ++t) {
            std::vector<double> g, f;
            for (int k = pg.pp_beg[t]; k < pg.pp_beg[t + 1]; ++k) { g.push_back(a.pps[k].w); f.push_back(a.pps_flat[k].w); }
            if (!std::is_sorted(f.begin(), f.end(), [](double x, double y) { return x > y; })) { std::printf("flat list not sorted\n"); ++bad; }
            std::sort(g.begin(), g.end()); std::sort(f.begin(), f.end());
            if (g != f) { std::printf("flat list is not a permutation\n"); ++bad; }
        }
    // shell pairs of a type come by non-increasing contraction length (quads of similar trip count)
    for (const PGDesc& pg : a.pgs)
        for (int t = 0; t < NPTYPE; ++t)
            for (int k = pg.sp_beg[t] + 1; k < pg.sp_beg[t + 1]; ++k)
                if (a.sps[k].pp_cnt > a.sps[k - 1].pp_cnt) { std::printf("shell pairs not sorted by length\n"); ++bad; }
    // P = A + (P - A) holds exactly as stored (P - A is formed from the difference of the centres, vb_setup.cpp), and the folded
    // density rows of ss / ps shell pairs are outer products over the pair list, D[e][(s,t)] = u_e[s] v_e[t] (every 2 x 2 minor
    // vanishes): the fact the one-index-at-a-time transforms of DESIGN.md "Next" rest on.  pp rows (HRR-folded d components
    // mix both centres) are counted, not required.
    {
        long long rows = 0, rank1 = 0, pp_rows = 0, pp_rank1 = 0;
        for (const PGDesc& pg : a.pgs) {
            if (pg.np <= 0) continue;
            for (int k = pg.sp_beg[0]; k < pg.sp_beg[NPTYPE]; ++k) {
                const SPRec& sp = a.sps[k];
                for (int e = 0; e < pt_ne(sp.type); ++e) {
                    const double* row = a.dmat.data() + pg.d_off + (size_t)(sp.eoff + e) * pg.np;
                    double mx = 0.0;
                    for (int p = 0; p < pg.np; ++p) mx = std::fmax(mx, std::fabs(row[p]));
                    if (mx == 0.0) continue;
                    bool ok = true;
                    for (int p = 0; p < pg.np && ok; ++p)
                        for (int q = 0; q < p && ok; ++q) {
                            const int s1 = a.pg_pairs[2 * (pg.pair_beg + p)], t1 = a.pg_pairs[2 * (pg.pair_beg + p) + 1];
                            const int s2 = a.pg_pairs[2 * (pg.pair_beg + q)], t2 = a.pg_pairs[2 * (pg.pair_beg + q) + 1];
                            if (s1 == s2 || t1 == t2) continue;
                            // the two crossed pairs (s1,t2), (s2,t1), if the group holds them
                            int pc = -1, qc = -1;
                            for (int r = 0; r < pg.np; ++r) {
                                const int sr = a.pg_pairs[2 * (pg.pair_beg + r)], tr = a.pg_pairs[2 * (pg.pair_beg + r) + 1];
                                if (sr == s1 && tr == t2) pc = r;
                                if (sr == s2 && tr == t1) qc = r;
                            }
                            if (pc < 0 || qc < 0) continue;
                            if (std::fabs(row[p] * row[q] - row[pc] * row[qc]) > 1e-13 * mx * mx) ok = false;
                        }
                    if (sp.type <= 1) { ++rows; rank1 += ok; } else { ++pp_rows; pp_rank1 += ok; }
                }
            }
        }
        std::printf("folded density rows: ss/ps %lld, outer products %lld; pp %lld, outer products %lld\n", rows, rank1, pp_rows, pp_rank1);
        if (rows == 0 || rank1 != rows) { std::printf("ss / ps density rows are not outer products over the pair list\n"); ++bad; }
        for (const PrimPair& pp : a.pps) {
            // A = P - (P - A) must be one of the atom centres to rounding (1e-12 bohr is four orders above the rounding of a 40 bohr coordinate)
            bool hit = false;
            for (int at = 0; at < in.natom && !hit; ++at) {
                const double dx = pp.Px - pp.PAx - in.coords[3 * at] * ANGS2BOHR, dy = pp.Py - pp.PAy - in.coords[3 * at + 1] * ANGS2BOHR,
                             dz = pp.Pz - pp.PAz - in.coords[3 * at + 2] * ANGS2BOHR;
                hit = std::fabs(dx) < 1e-12 && std::fabs(dy) < 1e-12 && std::fabs(dz) < 1e-12;
            }
            if (!hit) { std::printf("P - (P - A) is not an atom centre\n"); ++bad; break; }
        }
    }
    // first_order_opt's integral cache (TileOpts): with one entry isolated the pair groups touching it come last, the
    // "subject only" build reproduces exactly that tail (offsets relative), and the free part mirrors under (g,h)<->(h,g)
    {
        const int iso = in.nnd() + in.ndocc - 1;
        TileOpts o1;
        o1.isolate = iso;
        TileSetup f = build(in, "3", false, o1);
        TileOpts o2 = o1;
        o2.only_subject = true; o2.wcut = f.wmax;
        TileSetup sj = build(in, "2", false, o2);
        const int nfree = f.n_free_pg, nall = (int)f.pgs.size();
        if (nfree <= 0 || nfree >= nall) { std::printf("isolate: no free / subject split (%d of %d)\n", nfree, nall); ++bad; }
        const int giso = (int)f.groups.size() - 1;
        for (int x = 0; x < nall; ++x) {
            const bool subj = f.pgs[x].g == giso || f.pgs[x].h == giso;
            if (subj != (x >= nfree)) { std::printf("isolate: subject pair groups are not last\n"); ++bad; break; }
        }
        if (f.groups[giso].entries.size() != 1 || f.groups[giso].entries[0] != iso) { std::printf("isolate: group\n"); ++bad; }
        if ((int)sj.pgs.size() != nall - nfree) { std::printf("only_subject: %zu pair groups, expected %d\n", sj.pgs.size(), nall - nfree); ++bad; }
        else {
            std::vector<PrimPair> tail(f.pps.begin() + f.n_free_pps, f.pps.end());
            std::vector<double> dtail(f.dmat.begin() + f.n_free_d, f.dmat.end());
            std::vector<int> ptail(f.pg_pairs.begin() + 2 * f.n_free_pairs, f.pg_pairs.end());
            if (!same(tail, sj.pps) || !same(dtail, sj.dmat) || !same(ptail, sj.pg_pairs)) { std::printf("only_subject: tables differ from the tail of the full build\n"); ++bad; }
            for (int x = nfree; x < nall; ++x)
                if (f.pgs[x].d_off - (long long)f.n_free_d != sj.pgs[x - nfree].d_off || f.pgs[x].np != sj.pgs[x - nfree].np) { std::printf("only_subject: offsets\n"); ++bad; break; }
        }
        for (int x = 0; x < nfree; ++x) {
            bool found = false;
            for (int y = 0; y < nfree && !found; ++y) found = f.pgs[y].g == f.pgs[x].h && f.pgs[y].h == f.pgs[x].g && f.pgs[y].np == f.pgs[x].np;
            if (!found) { std::printf("free pair group (%d,%d) has no mirror\n", f.pgs[x].g, f.pgs[x].h); ++bad; break; }
        }
    }
    // table build shared by the ranks of one node (TileOpts::shard_mode): every rank publishes its share, then each
    // merges all shares -- bitwise the tables of the plain build, for any number of ranks
    for (int nranks : {2, 3}) {
        const std::string prefix = std::string(argv[1]) + ".share" + std::to_string(nranks) + "_";
        for (int r = 0; r < nranks; ++r) {
            TileOpts o;
            o.shard_mode = 1; o.shard_rank = r; o.shard_nranks = nranks; o.shard_prefix = prefix;
            build(in, "2", true, o);
        }
        TileOpts o;
        o.shard_mode = 2; o.shard_rank = nranks - 1; o.shard_nranks = nranks; o.shard_prefix = prefix;
        TileSetup c = build(in, "3", true, o);
        if (!same(a.pg_pairs, c.pg_pairs) || !same(a.sps, c.sps) || !same(a.pps, c.pps) || !same(a.pps_flat, c.pps_flat) ||
            !same(a.dmat, c.dmat) || !same(a.pgs, c.pgs)) { std::printf("sharded build (%d ranks) differs from the plain build\n", nranks); ++bad; }
        for (int r = 0; r < nranks; ++r) std::remove((prefix + std::to_string(r)).c_str());
        bool threw = false;
        try { build(in, "1", true, o); } catch (const std::exception&) { threw = true; }
        if (!threw) { std::printf("missing shares are not detected\n"); ++bad; }
    }
    // compact Boys table against the exact series
    std::vector<double> tab(BOYS_S_SIZE);
    boys_make_table_small(tab.data());
    double worst = 0.0;
    for (double T = 0.0; T < 60.0; T += 0.0173) {
        double F[5], R[5], G[3];
        boys_s<4>(tab.data(), T, F); boys_s<2>(tab.data(), T, G); boys_reference(4, T, R);
        for (int m = 0; m <= 4; ++m) worst = std::fmax(worst, std::fabs(F[m] - R[m]));
        for (int m = 0; m <= 2; ++m) worst = std::fmax(worst, std::fabs(G[m] - R[m]));
    }
    if (worst > 2e-15) { std::printf("compact Boys table off by %.3e\n", worst); ++bad; }
    std::printf("pair groups %zu, primitive pairs %zu, Boys max abs dev %.2e, %s\n", a.pgs.size(), a.pps.size(), worst, bad ? "FAIL" : "OK");
    return bad ? 1 : 0;
}
