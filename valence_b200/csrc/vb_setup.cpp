#include "vb_setup.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <functional>
#include <cstdint>
#include <stdexcept>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <chrono>
#include <thread>

namespace vb {

double dblfac(int n)   // valence.F90:2314-2322
{
    double r = 1.0;
    for (int i = 3; i <= n; i += 2) r *= (double)i;
    return r;
}

// valence.F90:2264-2305: primitive normalisation for the (l,0,0) component,
// then normalisation of the whole contraction
void norm_prim(int l, int n, const double* exps, const double* raw, double* out)
{
    const double pi32 = std::pow(std::acos(-1.0), 1.5);
    const double fac = pi32 * dblfac(2 * l - 1) * std::pow(2.0, (double)(-l));
    const double fax = -1.5 - l;
    for (int g = 0; g < n; ++g) out[g] = raw[g] * std::pow(fac * std::pow(2.0 * exps[g], fax), -0.5);
    double sovl = 0.0;
    for (int g = 0; g < n; ++g)
        for (int h = 0; h < n; ++h) sovl = sovl + fac * out[g] * out[h] * std::pow(exps[g] + exps[h], fax);
    sovl = std::pow(sovl, -0.5);
    for (int g = 0; g < n; ++g) out[g] = out[g] * sovl;
}

Basis build_basis(const Input& in, const std::vector<double>& xyz)
{
    Basis b;
    // per-type normalised shells
    std::vector<std::vector<int>> type_prim_off(in.natom_t);
    for (int t = 0; t < in.natom_t; ++t)
        for (const ShellDef& sh : in.types[t].shells) {
            if (sh.l > LMAX_SHELL) throw InputError("angular momentum above d is not supported");
            type_prim_off[t].push_back((int)b.exps.size());
            size_t n = sh.exps.size(), o = b.exps.size();
            b.exps.insert(b.exps.end(), sh.exps.begin(), sh.exps.end());
            b.coefs.resize(o + n);
            norm_prim(sh.l, (int)n, sh.exps.data(), sh.raw.data(), b.coefs.data() + o);
        }
    int ao = 0, tsh_base = 0;
    std::vector<int> type_shell_base(in.natom_t);
    for (int t = 0; t < in.natom_t; ++t) { type_shell_base[t] = tsh_base; tsh_base += (int)in.types[t].shells.size(); }
    for (int a = 0; a < in.natom; ++a) {
        int t = in.atom_t[a] - 1;
        b.atom_first_shell.push_back((int)b.shells.size());
        for (size_t s = 0; s < in.types[t].shells.size(); ++s) {
            GShell g;
            g.l = in.types[t].shells[s].l;
            g.atom = a;
            g.nprim = (int)in.types[t].shells[s].exps.size();
            g.ao_off = ao;
            g.prim_off = type_prim_off[t][s];
            g.type_shell = type_shell_base[t] + (int)s;
            for (int d = 0; d < 3; ++d) g.r[d] = xyz[3 * a + d];
            ao += ncart(g.l);
            b.shells.push_back(g);
        }
    }
    b.atom_first_shell.push_back((int)b.shells.size());
    b.nao = ao;
    // setangn, valence.F90:2395-2421
    b.angn.assign(ncum(LMAX_SHELL), 1.0);
    for (int c = 0; c < ncum(LMAX_SHELL); ++c) {
        int L = c_L(c);
        double ashl = std::sqrt(dblfac(2 * L - 1));
        b.angn[c] = ashl * (1.0 / std::sqrt(dblfac(2 * c_lx(c) - 1))) * (1.0 / std::sqrt(dblfac(2 * c_ly(c) - 1))) *
                    (1.0 / std::sqrt(dblfac(2 * c_lz(c) - 1)));
    }
    return b;
}

double nuclear_repulsion(const Input& in, const std::vector<double>& xyz)
{
    double nre = 0.0;
    for (int i = 1; i < in.natom; ++i)
        for (int j = 0; j < i; ++j) {
            double zij = in.types[in.atom_t[i] - 1].charge * in.types[in.atom_t[j] - 1].charge;
            if (std::fabs(zij) > 1.e-12) {
                double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1], dz = xyz[3 * i + 2] - xyz[3 * j + 2];
                double rsq = dx * dx + dy * dy + dz * dz;
                if (rsq > 1.e-5) nre = nre + zij * std::pow(rsq, -0.5);
            }
        }
    return nre;
}

// OBS layout of an orbital: atoms in listed order, shells in basis order,
// cartesians in CCA order
static void obs_layout(const Input& in, const Basis& bas, int orb, std::vector<int>* gsh, std::vector<int>* cmp,
                       std::vector<int>* atom_start)
{
    gsh->clear(); cmp->clear();
    if (atom_start) atom_start->clear();
    for (int a1 : in.orbitals[orb].atoms) {
        int a = a1 - 1;
        if (atom_start) atom_start->push_back((int)gsh->size());
        for (int s = bas.atom_first_shell[a]; s < bas.atom_first_shell[a + 1]; ++s)
            for (int c = 0; c < ncart(bas.shells[s].l); ++c) { gsh->push_back(s); cmp->push_back(c); }
    }
}

bool obs_position(const Input& in, const Basis& bas, int orb, int pos, int* gshell, int* comp)
{
    std::vector<int> gsh, cmp;
    obs_layout(in, bas, orb, &gsh, &cmp, nullptr);
    if (pos < 1 || pos > (int)gsh.size()) return false;
    *gshell = gsh[pos - 1];
    *comp = cmp[pos - 1];
    return true;
}

ExpOrb expand_orbital(const Input& in, const Basis& bas, const std::vector<std::vector<double>>& coeff, int orb)
{
    std::vector<int> gsh, cmp, astart;
    obs_layout(in, bas, orb, &gsh, &cmp, &astart);
    std::vector<double> cf(gsh.size(), 0.0);
    const OrbitalDef& od = in.orbitals[orb];
    const int norbs = in.norbs();
    for (size_t i = 0; i < od.xp.size(); ++i) {
        if (od.xp[i] < 1) {
            int indf = od.xp[i] + norbs - 1;   // 0-based id of the DBF (0 = last one entered)
            if (indf < 0 || indf >= norbs) throw InputError("DBF label out of range");
            std::vector<int> dgsh, dcmp, dstart;
            obs_layout(in, bas, indf, &dgsh, &dcmp, &dstart);
            const OrbitalDef& df = in.orbitals[indf];
            // map each DBF atom to the start of the same atom in this orbital's OBS
            std::vector<int> shift(df.atoms.size(), INT32_MIN);
            for (size_t ja = 0; ja < df.atoms.size(); ++ja)
                for (size_t ia = 0; ia < od.atoms.size(); ++ia)
                    if (od.atoms[ia] == df.atoms[ja]) shift[ja] = astart[ia] - dstart[ja];
            for (size_t j = 0; j < df.xp.size(); ++j) {
                int p = df.xp[j];
                if (p < 1 || p > (int)dgsh.size()) throw InputError("DBF expansion index out of range");
                int ja = (int)df.atoms.size() - 1;
                while (ja > 0 && dstart[ja] > p - 1) --ja;
                if (shift[ja] == INT32_MIN) throw InputError("DBF basis is not a subset of the orbital basis");
                cf[p - 1 + shift[ja]] += coeff[indf][j] * coeff[orb][i];
            }
        } else {
            if (od.xp[i] > (int)cf.size()) throw InputError("orbital expansion index out of range");
            cf[od.xp[i] - 1] += coeff[orb][i];
        }
    }
    ExpOrb e;
    for (size_t k = 0; k < gsh.size();) {
        OrbShell os;
        os.gshell = gsh[k];
        std::memset(os.c, 0, sizeof os.c);
        int n = ncart(bas.shells[gsh[k]].l);
        for (int c = 0; c < n; ++c) os.c[c] = cf[k + c];
        e.sh.push_back(os);
        k += n;
    }
    return e;
}

// ---------------------------------------------------------------------------
// Fold the horizontal recurrence into a cartesian pair density d[a][b]
// (shell A carries la >= lb):  sum_ab d[a][b] (ab| = sum_e dt[e] (e0|,
//   (x-B)^b = sum_k C(b,k) (A-B)^(b-k) (x-A)^k   per cartesian direction.
static void hrr_fold(int la, int lb, const double* AB, const double* d /*[na][nb]*/, double* dt /*[pt_ne]*/)
{
    const int na = ncart(la), nb = ncart(lb), e0 = coff(la);
    if (lb == 0) {                       // nothing to shift: e = a
        for (int a = 0; a < na; ++a) dt[a] += d[a];
        return;
    }
    static const double BIN[3][3] = {{1, 0, 0}, {1, 1, 0}, {1, 2, 1}};
    double pw[3][3];
    for (int k = 0; k < 3; ++k) { pw[k][0] = 1.0; pw[k][1] = AB[k]; pw[k][2] = AB[k] * AB[k]; }
    for (int a = 0; a < na; ++a) {
        int ca = coff(la) + a, ax = c_lx(ca), ay = c_ly(ca), az = c_lz(ca);
        for (int b = 0; b < nb; ++b) {
            double v = d[a * nb + b];
            if (v == 0.0) continue;
            int cb = coff(lb) + b, bx = c_lx(cb), by = c_ly(cb), bz = c_lz(cb);
            for (int kx = 0; kx <= bx; ++kx)
                for (int ky = 0; ky <= by; ++ky)
                    for (int kz = 0; kz <= bz; ++kz) {
                        double f = BIN[bx][kx] * BIN[by][ky] * BIN[bz][kz] * pw[0][bx - kx] * pw[1][by - ky] * pw[2][bz - kz];
                        dt[cidx(ax + kx, ay + ky, az + kz) - e0] += v * f;
                    }
        }
    }
}

static const OrbShell* find_shell(const ExpOrb& o, int gshell)
{
    for (const OrbShell& s : o.sh)
        if (s.gshell == gshell) return &s;
    return nullptr;
}

void build_tiles(const Input& in, const Basis& bas, const Wavefunction& wf, const std::vector<ExpOrb>& orbs,
                 double tau, bool flat, TileSetup* out, const TileOpts& opts)
{
    (void)in;
    TileSetup& ts = *out;
    int seg_len = 64;         // primitives per s / p shell-pair segment (VB_SEG; 64 leaves the 6-31G shell pairs whole: no
                              // measurable difference for k_pclass, profiles/r2_segment_sweep.log)
    if (const char* e = std::getenv("VB_SEG")) seg_len = std::min(64, std::max(1, std::atoi(e)));
    // reset, but keep the storage of the big tables: a TileSetup that is reused across calls is resized in place at
    // the merge (every element is overwritten there), which saves zero-filling and page-faulting ~0.6 GB per call
    ts.groups.clear(); ts.pgs.clear();
    ts.wmax = 0.0;
    ts.max_ne = ts.max_np = ts.max_npp = ts.max_nsp = ts.max_ks = ts.lmax = 0;
    ts.n_free_pg = 0; ts.n_free_pairs = ts.n_free_sps = ts.n_free_pps = ts.n_free_d = 0;
    if (opts.measure_only) { ts.pg_pairs.clear(); ts.sps.clear(); ts.pps.clear(); ts.pps_flat.clear(); ts.dmat.clear(); }
    const int nso = wf.nso;
    // --- entry groups -------------------------------------------------------
    const int NG_MAX = 5, AO_BUDGET = 80;
    auto entry_shells = [&](int s) {
        std::vector<int> v;
        int sl = wf.slot(s, 0);
        for (const OrbShell& x : orbs[wf.bra[sl]].sh) v.push_back(x.gshell);
        for (const OrbShell& x : orbs[wf.ket[sl]].sh) v.push_back(x.gshell);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        return v;
    };
    auto nao_of = [&](const std::vector<int>& sh) { int n = 0; for (int s : sh) n += ncart(bas.shells[s].l); return n; };
    // Entries whose orbital basis sets live on nested atom sets share a group (e.g. the five orbitals
    // of one water molecule, whatever shells the weight screen left them): every AO quartet of a
    // molecule quartet is then generated once for all of their orbital pairs.
    auto atoms_of = [&](const std::vector<int>& sh) {
        std::vector<int> a;
        for (int s : sh) a.push_back(bas.shells[s].atom);
        std::sort(a.begin(), a.end());
        a.erase(std::unique(a.begin(), a.end()), a.end());
        return a;
    };
    std::vector<std::vector<int>> group_atoms;
    for (int s = 0; s < nso; ++s) {
        if (s == opts.isolate) continue;
        std::vector<int> sh = entry_shells(s);
        std::vector<int> at = atoms_of(sh);
        int placed = -1;
        int lo = std::max(0, (int)ts.groups.size() - 16);
        for (int g = (int)ts.groups.size() - 1; g >= lo && placed < 0; --g) {
            EntryGroup& G = ts.groups[g];
            if ((int)G.entries.size() >= NG_MAX) continue;
            std::vector<int> ua;
            std::set_union(group_atoms[g].begin(), group_atoms[g].end(), at.begin(), at.end(), std::back_inserter(ua));
            if (ua.size() != std::max(group_atoms[g].size(), at.size())) continue;   // atom sets must be nested
            std::vector<int> un;
            std::set_union(G.shells.begin(), G.shells.end(), sh.begin(), sh.end(), std::back_inserter(un));
            if (nao_of(un) * ((int)G.entries.size() + 1) > AO_BUDGET) continue;
            G.shells = un;
            G.nao = nao_of(un);
            G.entries.push_back(s);
            group_atoms[g] = ua;
            placed = g;
        }
        if (placed < 0) {
            EntryGroup G;
            G.entries.push_back(s);
            G.shells = sh;
            G.nao = nao_of(sh);
            ts.groups.push_back(G);
            group_atoms.push_back(at);
        }
    }
    int g_iso = -1;
    if (opts.isolate >= 0 && opts.isolate < nso) {
        EntryGroup G;
        G.entries.push_back(opts.isolate);
        G.shells = entry_shells(opts.isolate);
        G.nao = nao_of(G.shells);
        g_iso = (int)ts.groups.size();
        ts.groups.push_back(G);
        group_atoms.push_back(atoms_of(G.shells));
    }
    // --- pair groups --------------------------------------------------------
    const double SQ2PI54 = std::sqrt(2.0) * std::pow(PI, 1.25);
    const int ng = (int)ts.groups.size();
    const int nshell = (int)bas.shells.size();
    // per shell: most diffuse exponent, largest |contraction weight| (cheap far-field bound)
    std::vector<double> sh_emin(nshell), sh_cmax(nshell);
    for (int s = 0; s < nshell; ++s) {
        const GShell& g = bas.shells[s];
        double em = 1e300, cm = 0.0;
        for (int k = 0; k < g.nprim; ++k) { em = std::min(em, bas.exps[g.prim_off + k]); cm = std::max(cm, std::fabs(bas.coefs[g.prim_off + k])); }
        sh_emin[s] = em; sh_cmax[s] = cm;
    }
    // per orbital: largest sum_c |weight * angn| over its shells (bounds any cartesian pair density)
    std::vector<double> orb_cabs(orbs.size(), 0.0);
    for (size_t o = 0; o < orbs.size(); ++o)
        for (const OrbShell& x : orbs[o].sh) {
            const int l = bas.shells[x.gshell].l;
            double a = 0.0;
            for (int c = 0; c < ncart(l); ++c) a += std::fabs(x.c[c] * bas.angn[coff(l) + c]);
            orb_cabs[o] = std::max(orb_cabs[o], a);
        }
    // per group, entry and group shell: the entry's bra / ket orbital on that shell (or null)
    std::vector<std::vector<const OrbShell*>> grp_bra(ng), grp_ket(ng);
    for (int g = 0; g < ng; ++g) {
        const EntryGroup& G = ts.groups[g];
        const size_t nsh = G.shells.size();
        grp_bra[g].assign(G.entries.size() * nsh, nullptr);
        grp_ket[g].assign(G.entries.size() * nsh, nullptr);
        for (size_t ei = 0; ei < G.entries.size(); ++ei)
            for (size_t xi = 0; xi < nsh; ++xi) {
                const int sl = wf.slot(G.entries[ei], 0);
                grp_bra[g][ei * nsh + xi] = find_shell(orbs[wf.bra[sl]], G.shells[xi]);
                grp_ket[g][ei * nsh + xi] = find_shell(orbs[wf.ket[sl]], G.shells[xi]);
            }
    }
    struct TmpSP {
        int type, A, B;
        bool swapped;
        double wmax;
        std::vector<PrimPair> pp;
        std::vector<double> dt;   // folded density [nE][np]
    };
    struct PGOut {                // one pair group with offsets relative to its own arrays
        bool used = false;
        PGDesc pg;
        std::vector<int> pairs;
        std::vector<SPRec> sps;
        std::vector<PrimPair> pps;
        std::vector<double> dmat;
        double wmax = 0.0;
        // what the merge reads: this object's own vectors, or a mapped share published by another rank
        const int* v_pairs = nullptr; const SPRec* v_sps = nullptr; const PrimPair* v_pps = nullptr; const double* v_dmat = nullptr;
        size_t n_pairs = 0, n_sps = 0, n_pps = 0, n_d = 0;
        void view_own() { v_pairs = pairs.data(); v_sps = sps.data(); v_pps = pps.data(); v_dmat = dmat.data();
                          n_pairs = pairs.size(); n_sps = sps.size(); n_pps = pps.size(); n_d = dmat.size(); }
    };
    // one pair group; pass 0 only measures the largest primitive weight, pass 1 emits
    auto do_pair_group = [&](int g, int h, int pass, double wcut, PGOut& out) {
        const EntryGroup& G = ts.groups[g];
        const EntryGroup& H = ts.groups[h];
        std::vector<double> dcart(36), dfold(64);
        std::vector<int>& pairs = out.pairs;
        std::vector<int> pair_ei, pair_ti;   // positions of the pair's entries inside their groups
        pairs.clear();
        double dsum = 0.0;        // bound on sum_ab |d_ab| of any orbital pair of this group pair
        for (size_t ei = 0; ei < G.entries.size(); ++ei)
            for (size_t ti = 0; ti < H.entries.size(); ++ti) {
                const int s = G.entries[ei], t = H.entries[ti];
                if (wf.sym && g == h && t > s) continue;
                pairs.push_back(s);
                pairs.push_back(t);
                pair_ei.push_back((int)ei);
                pair_ti.push_back((int)ti);
                dsum = std::max(dsum, orb_cabs[wf.bra[wf.slot(s, 0)]] * orb_cabs[wf.ket[wf.slot(t, 0)]]);
                dsum = std::max(dsum, orb_cabs[wf.ket[wf.slot(s, 0)]] * orb_cabs[wf.bra[wf.slot(t, 0)]]);
            }
        const int np = (int)pairs.size() / 2;
        if (np == 0) return;
        std::vector<TmpSP> tsp;
        const size_t nshG = G.shells.size(), nshH = H.shells.size();
        tsp.reserve(nshG * nshH);
        if (pass == 1 && g != h) {
            // whole group pair out of range: the shell-pair test below, bounded from above over all shell pairs of
            // (G, H) (most diffuse exponents, largest weights, shortest / longest centre distance)
            double eG = 1e300, eH = 1e300, cG = 0.0, cH = 0.0, d2min = 1e300, d2max = 0.0, dcmax = 0.0;
            int lG = 0, lH = 0;
            for (int X : G.shells) { eG = std::min(eG, sh_emin[X]); cG = std::max(cG, sh_cmax[X]); lG = std::max(lG, bas.shells[X].l); }
            for (int Y : H.shells) { eH = std::min(eH, sh_emin[Y]); cH = std::max(cH, sh_cmax[Y]); lH = std::max(lH, bas.shells[Y].l); }
            for (int X : G.shells)
                for (int Y : H.shells) {
                    const double dx = bas.shells[X].r[0] - bas.shells[Y].r[0], dy = bas.shells[X].r[1] - bas.shells[Y].r[1],
                                 dz = bas.shells[X].r[2] - bas.shells[Y].r[2];
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    d2min = std::min(d2min, d2); d2max = std::max(d2max, d2);
                    dcmax = std::max(dcmax, std::max(std::fabs(dx), std::max(std::fabs(dy), std::fabs(dz))));
                }
            const double pmin = eG + eH;
            const double kb = cG * cH * std::exp(-eG * eH / pmin * d2min) * SQ2PI54 / pmin;
            const double len = std::sqrt(d2max) + 1.0 / std::sqrt(pmin);
            double db = dsum;
            for (int k = 0; k < std::max(lG, lH); ++k) db *= 1.0 + dcmax;
            double poly = 0.0, lp = 1.0;
            for (int L = 0; L <= lG + lH; ++L) { poly += db * lp * ncart(L); lp *= 2.0 * len; }
            if (!(kb * std::pow(2.0 * pmin, -0.25) * poly * wcut * 1.0001 * 1.01 >= tau)) return;
        }
        for (size_t xi = 0; xi < nshG; ++xi)
            for (size_t yi = 0; yi < nshH; ++yi) {
                const int X = G.shells[xi], Y = H.shells[yi];
                TmpSP sp;
                sp.swapped = bas.shells[X].l < bas.shells[Y].l;
                sp.A = sp.swapped ? Y : X;
                sp.B = sp.swapped ? X : Y;
                const GShell& sa = bas.shells[sp.A];
                const GShell& sb = bas.shells[sp.B];
                sp.type = ptype(sa.l, sb.l);
                const int na = ncart(sa.l), nb = ncart(sb.l), nE = pt_ne(sp.type), EA = pt_E(sp.type);
                double AB[3] = {sa.r[0] - sb.r[0], sa.r[1] - sb.r[1], sa.r[2] - sb.r[2]};
                double AB2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
                if (pass == 1) {
                    // far-field rejection before any density is folded: every primitive weight of this
                    // shell pair is below the bound built from the most diffuse primitives
                    const double amin = sh_emin[sp.A], bmin = sh_emin[sp.B], pmin = amin + bmin;
                    const double kb = sh_cmax[sp.A] * sh_cmax[sp.B] * std::exp(-amin * bmin / pmin * AB2) * SQ2PI54 / pmin;
                    const double len = std::sqrt(AB2) + 1.0 / std::sqrt(pmin);
                    const double abm = 1.0 + std::max(std::fabs(AB[0]), std::max(std::fabs(AB[1]), std::fabs(AB[2])));
                    double db = dsum;
                    for (int k = 0; k < sb.l; ++k) db *= abm;
                    double poly = 0.0, lp = 1.0;
                    for (int L = 0; L <= EA; ++L) { if (L >= sa.l) poly += db * lp * ncart(L); lp *= 2.0 * len; }
                    if (!(kb * std::pow(2.0 * pmin, -0.25) * poly * wcut * 1.0001 >= tau)) continue;
                }
                // folded densities of every orbital pair on this shell pair
                sp.dt.assign((size_t)nE * np, 0.0);
                double dmaxL[EMAX + 1] = {0, 0, 0, 0, 0};   // max |folded density| per angular momentum of e
                bool any = false;
                for (int ip = 0; ip < np; ++ip) {
                    // original orientation: X from group g (bra-side orbital of s), Y from group h (ket-side
                    // orbital of t); A is whichever of the two carries the larger angular momentum
                    const OrbShell* cX = grp_bra[g][pair_ei[ip] * nshG + xi];
                    const OrbShell* cY = grp_ket[h][pair_ti[ip] * nshH + yi];
                    if (!cX || !cY) continue;
                    const OrbShell* cA = sp.swapped ? cY : cX;
                    const OrbShell* cB = sp.swapped ? cX : cY;
                    if (sb.l == 0) {
                        // nothing to fold (e = a): same arithmetic as the general branch, without its scaffolding
                        const double fb = bas.angn[coff(0)];
                        bool nz0 = false;
                        for (int a = 0; a < na; ++a) {
                            const double v = cA->c[a] * bas.angn[coff(sa.l) + a] * cB->c[0] * fb;
                            dcart[a] = v;
                            nz0 = nz0 || v != 0.0;
                        }
                        if (!nz0) continue;
                        for (int a = 0; a < na; ++a) {
                            const double v = 0.0 + dcart[a];
                            sp.dt[(size_t)a * np + ip] = v;
                            dmaxL[sa.l] = std::max(dmaxL[sa.l], std::fabs(v));
                            any = any || v != 0.0;
                        }
                        continue;
                    }
                    bool nz = false;
                    for (int a = 0; a < na; ++a)
                        for (int b = 0; b < nb; ++b) {
                            double v = cA->c[a] * bas.angn[coff(sa.l) + a] * cB->c[b] * bas.angn[coff(sb.l) + b];
                            dcart[a * nb + b] = v;
                            nz = nz || v != 0.0;
                        }
                    if (!nz) continue;
                    std::fill(dfold.begin(), dfold.begin() + nE, 0.0);
                    hrr_fold(sa.l, sb.l, AB, dcart.data(), dfold.data());
                    for (int e = 0; e < nE; ++e) {
                        sp.dt[(size_t)e * np + ip] = dfold[e];
                        int L = c_L(coff(sa.l) + e);
                        dmaxL[L] = std::max(dmaxL[L], std::fabs(dfold[e]));
                        any = any || dfold[e] != 0.0;
                    }
                }
                if (!any) continue;   // no orbital pair has density on this shell pair
                sp.wmax = 0.0;
                if (pass == 1) sp.pp.reserve((size_t)sa.nprim * sb.nprim);
                for (int ia = 0; ia < sa.nprim; ++ia)
                    for (int ib = 0; ib < sb.nprim; ++ib) {
                        double a = bas.exps[sa.prim_off + ia], b = bas.exps[sb.prim_off + ib], p = a + b;
                        const double ex = AB2 == 0.0 ? 1.0 : std::exp(-a * b / p * AB2);   // exp(-0) = 1 exactly
                        double K = bas.coefs[sa.prim_off + ia] * bas.coefs[sb.prim_off + ib] * ex * SQ2PI54;
                        if (K == 0.0) continue;
                        PrimPair pp;
                        // P - A = b/p (B - A) formed from the difference of the centres: P - A as a difference of
                        // absolute coordinates loses |coordinate| / |P - A| units in the last place (1e-14 relative for an
                        // atom 40 bohr from the origin); P = A + (P - A) keeps the pair consistent
                        const double sAB = b / p;
                        pp.PAx = sAB * (sb.r[0] - sa.r[0]); pp.PAy = sAB * (sb.r[1] - sa.r[1]); pp.PAz = sAB * (sb.r[2] - sa.r[2]);
                        pp.Px = sa.r[0] + pp.PAx;
                        pp.Py = sa.r[1] + pp.PAy;
                        pp.Pz = sa.r[2] + pp.PAz;
                        pp.p = p;
                        pp.ip = 1.0 / p;
                        pp.Kp = K / p;
                        // Schwarz-type magnitude: the charge cloud sum_e dt[e] [e0| of this primitive pair has
                        // self-repulsion <= w^2, so its share of any (st|uv) is <= w_p w_q.
                        // [ss|ss]_pp = Kp^2 / sqrt(2p); each unit of angular momentum adds a factor
                        // <= |PA| + 1/sqrt(p) (generous: includes the 1/2p and 1/4p terms).
                        double len = std::sqrt(pp.PAx * pp.PAx + pp.PAy * pp.PAy + pp.PAz * pp.PAz) + 1.0 / std::sqrt(p);
                        double poly = 0.0, lp = 1.0;
                        for (int L = 0; L <= EA; ++L) { if (L >= sa.l) poly += dmaxL[L] * lp * ncart(L); lp *= 2.0 * len; }
                        pp.w = std::fabs(pp.Kp) / std::sqrt(std::sqrt(2.0 * p)) * poly;
                        if (pass == 1 && !(pp.w * wcut >= tau)) continue;
                        sp.wmax = std::max(sp.wmax, pp.w);
                        if (pass == 1) sp.pp.push_back(pp);
                    }
                if (pass == 0) { out.wmax = std::max(out.wmax, sp.wmax); continue; }
                if (sp.pp.empty()) continue;
                // by decreasing weight, stable; a few primitives per shell pair on average: insertion sort (std::stable_sort
                // allocates a buffer per call)
                // s / p shell pairs leave as segments of seg_len primitives (below): order by exponent sum first, tightest
                // primitives first, so that a segment holds primitives of similar extent (its far-field radius is that of
                // its most diffuse primitive), then by decreasing weight inside each segment
                auto by_weight = [&](size_t lo, size_t hi) {
                    for (size_t i = lo + 1; i < hi; ++i) {
                        const PrimPair x = sp.pp[i];
                        size_t j = i;
                        for (; j > lo && x.w > sp.pp[j - 1].w; --j) sp.pp[j] = sp.pp[j - 1];
                        sp.pp[j] = x;
                    }
                };
                if (sp.type <= 1 && sp.pp.size() > (size_t)seg_len) {
                    std::stable_sort(sp.pp.begin(), sp.pp.end(), [](const PrimPair& x, const PrimPair& y) { return x.p > y.p; });
                    for (size_t c0 = 0; c0 < sp.pp.size(); c0 += (size_t)seg_len) by_weight(c0, std::min(sp.pp.size(), c0 + (size_t)seg_len));
                } else if (sp.pp.size() <= 12) {
                    by_weight(0, sp.pp.size());
                } else {
                    std::stable_sort(sp.pp.begin(), sp.pp.end(), [](const PrimPair& x, const PrimPair& y) { return x.w > y.w; });
                }
                tsp.push_back(std::move(sp));
            }
        if (pass == 0 || tsp.empty()) return;
        // by type, strongest shell pairs first
        std::stable_sort(tsp.begin(), tsp.end(), [](const TmpSP& a, const TmpSP& b) {
            return a.type != b.type ? a.type < b.type : a.wmax > b.wmax;
        });
        PGDesc& pg = out.pg;
        std::memset(&pg, 0, sizeof pg);
        pg.g = g; pg.h = h; pg.np = np; pg.pair_beg = 0;
        int ne = 0;
        for (const TmpSP& sp : tsp) ne += pt_ne(sp.type);
        pg.ne = ne;
        pg.d_off = 0;
        out.dmat.assign((((size_t)ne * np + 1) & ~(size_t)1), 0.0);   // keep blocks 16-byte aligned
        double* D = out.dmat.data();
        int t_cur = 0, eoff = 0;
        std::vector<SPRec>& recs = out.sps;
        recs.clear();
        out.pps.clear();
        for (size_t k = 0; k < tsp.size(); ++k) {
            const TmpSP& sp = tsp[k];
            while (t_cur <= sp.type) { pg.pp_beg[t_cur] = (int)out.pps.size(); pg.e_beg[t_cur] = eoff; ++t_cur; }
            const int nE = pt_ne(sp.type);
            // s and p shell pairs enter the tables as segments of at most seg_len primitives (all with the pair's e-offset:
            // the integrals are linear in the primitives).  The segment kernels (vb_pseg.cuh) walk (bra segment, ket segment)
            // pairs per lane; bounded, similar lengths keep the lanes of a warp in step.  Primitives are sorted by magnitude,
            // so a segment's first weight is its largest.
            {
                // (splitting the one-to-four pp shell pairs of a pair group over the four bra lanes of a quad was measured as
                // well: 10-25 % slower for the pp classes -- every extra segment is another tensor-core feed)
                const int cnt = (int)sp.pp.size(), base = (int)out.pps.size();
                const int S = sp.type <= 1 ? seg_len : cnt;
                for (int c0 = 0; c0 < cnt; c0 += S) {
                    double ipmax = 0.0;
                    for (int i = c0; i < std::min(cnt, c0 + S); ++i) ipmax = std::max(ipmax, sp.pp[i].ip);
                    recs.push_back({sp.type, eoff, base + c0, std::min(S, cnt - c0), sp.pp[c0].w, ipmax});
                }
            }
            pg.kwmax[sp.type] = std::max(pg.kwmax[sp.type], sp.wmax);
            for (PrimPair pp : sp.pp) { pp.eoff = eoff; pp.pad = 0; pp.wseg = sp.wmax; out.pps.push_back(pp); }
            std::copy(sp.dt.begin(), sp.dt.end(), D + (size_t)eoff * np);
            eoff += nE;
        }
        while (t_cur <= NPTYPE) { pg.pp_beg[t_cur] = (int)out.pps.size(); pg.e_beg[t_cur] = eoff; ++t_cur; }
        // per type, most expensive shell pairs first (they are dealt round-robin to the warps)
        std::stable_sort(recs.begin(), recs.end(), [](const SPRec& a, const SPRec& b) {
            return a.type != b.type ? a.type < b.type : a.pp_cnt > b.pp_cnt;
        });
        {
            int k = 0;
            for (int t = 0; t <= NPTYPE; ++t) {
                while (k < (int)recs.size() && recs[k].type < t) ++k;
                pg.sp_beg[t] = k;
            }
        }
        out.used = true;
        out.view_own();
    };
    const bool dbg_t = std::getenv("VB_DEBUG_SETUP") != nullptr;
    auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tt0 = tnow();
    // pass 0: the largest weights live in the one-group pair groups
    if (opts.wcut > 0.0) ts.wmax = opts.wcut;
    else
        for (int g = 0; g < ng; ++g) {
            if (opts.measure_only && opts.only_subject && g != g_iso) continue;
            PGOut o;
            do_pair_group(g, g, 0, 0.0, o);
            ts.wmax = std::max(ts.wmax, o.wmax);
        }
    if (opts.measure_only) return;
    const double wcut = ts.wmax;
    const double tt1 = tnow();
    // pass 1 over all group pairs, on the host cores; merged in (g,h) order so the layout (and with it
    // every reduction order on the device) does not depend on the thread count
    std::vector<std::pair<int, int>> gh;
    for (int g = 0; g < ng; ++g)
        for (int h = 0; h < (wf.sym ? g + 1 : ng); ++h)
            if (g != g_iso && h != g_iso && !opts.only_subject) gh.emplace_back(g, h);
    const size_t n_free_gh = gh.size();
    if (g_iso >= 0)
        for (int g = 0; g < ng; ++g)
            for (int h = 0; h < (wf.sym ? g + 1 : ng); ++h)
                if (g == g_iso || h == g_iso) gh.emplace_back(g, h);
    std::vector<PGOut> outs(gh.size());
    auto host_threads = [&](size_t nwork, size_t serial_below) {
        int nthr = (int)std::thread::hardware_concurrency();
        if (const char* e = std::getenv("VB_HOST_THREADS")) nthr = std::atoi(e);
        nthr = std::max(1, std::min(nthr, 32));
        if (nwork < serial_below) nthr = 1;
        return nthr;
    };
    auto run_pool = [&](int nthr, const std::function<void()>& work) {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
        work();
        for (std::thread& t : pool) t.join();
    };
    // ownership of the pair groups when the build is shared by the ranks of a node: blocks of 16, round-robin
    const int sh_n = opts.shard_mode ? std::max(1, opts.shard_nranks) : 1, sh_r = opts.shard_rank;
    auto owned = [&](size_t i) { return (int)((i / 16) % (size_t)sh_n) == sh_r; };
    // payload of a record: pps, sps, dmat, pairs -- every section 16-byte aligned (PrimPair / SPRec are)
    struct alignas(16) ShardRec { uint64_t i, off, n_pairs, n_sps, n_pps, n_d; double wmax; PGDesc pg; };
    constexpr size_t SHARD_HEAD = 32;
    constexpr uint64_t SHARD_MAGIC = 0x56425348415244ull;
    if (opts.shard_mode != 2) {
        std::atomic<size_t> next{0};
        run_pool(host_threads(gh.size(), 64), [&]() {
            for (;;) {
                size_t i0 = next.fetch_add(16);
                if (i0 >= gh.size()) break;
                if ((opts.shard_mode == 1 || opts.shard_mode == 3) && !owned(i0)) continue;
                for (size_t i = i0; i < std::min(gh.size(), i0 + 16); ++i) do_pair_group(gh[i].first, gh[i].second, 1, wcut, outs[i]);
            }
        });
    }
    if (opts.shard_mode == 1) {
        // publish this rank's pair groups: [magic, # group pairs, # records] records payload; written to a temporary
        // name and renamed, so a reader never sees a partial file
        std::vector<ShardRec> recs;
        uint64_t off = 0;
        for (size_t i = 0; i < outs.size(); ++i) {
            const PGOut& o = outs[i];
            if (!owned(i) || !o.used) continue;
            ShardRec r;
            std::memset(&r, 0, sizeof r);
            r.i = i; r.off = off; r.n_pairs = o.pairs.size(); r.n_sps = o.sps.size(); r.n_pps = o.pps.size(); r.n_d = o.dmat.size();
            r.wmax = o.wmax; r.pg = o.pg;
            off += r.n_pairs * sizeof(int) + r.n_sps * sizeof(SPRec) + r.n_pps * sizeof(PrimPair) + r.n_d * sizeof(double);
            off = (off + 15) & ~(uint64_t)15;
            recs.push_back(r);
        }
        const uint64_t head[4] = {SHARD_MAGIC, (uint64_t)gh.size(), (uint64_t)recs.size(), 0};
        static_assert(sizeof head == SHARD_HEAD && sizeof(ShardRec) % 16 == 0, "share layout");
        const size_t base = SHARD_HEAD + recs.size() * sizeof(ShardRec), total = base + off;
        const std::string fin = opts.shard_prefix + std::to_string(sh_r), tmp = fin + ".tmp";
        const int fd = open(tmp.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0600);
        char* buf = nullptr;
        if (fd >= 0 && ftruncate(fd, (off_t)total) == 0) {
            void* mp = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (mp != MAP_FAILED) buf = static_cast<char*>(mp);
        }
        if (fd >= 0) close(fd);
        if (!buf) throw std::runtime_error("valence_b200: cannot create the table share " + tmp);
        std::memcpy(buf, head, sizeof head);
        if (!recs.empty()) std::memcpy(buf + SHARD_HEAD, recs.data(), recs.size() * sizeof(ShardRec));
        {
            std::atomic<size_t> nx{0};
            run_pool(host_threads(recs.size(), 256), [&]() {
                for (size_t k; (k = nx.fetch_add(1)) < recs.size();) {
                    const ShardRec& r = recs[k];
                    const PGOut& o = outs[r.i];
                    char* p = buf + base + r.off;
                    std::memcpy(p, o.pps.data(), r.n_pps * sizeof(PrimPair)); p += r.n_pps * sizeof(PrimPair);
                    std::memcpy(p, o.sps.data(), r.n_sps * sizeof(SPRec)); p += r.n_sps * sizeof(SPRec);
                    std::memcpy(p, o.dmat.data(), r.n_d * sizeof(double)); p += r.n_d * sizeof(double);
                    std::memcpy(p, o.pairs.data(), r.n_pairs * sizeof(int));
                }
            });
        }
        munmap(buf, total);
        if (std::rename(tmp.c_str(), fin.c_str()) != 0) throw std::runtime_error("valence_b200: cannot publish the table share " + fin);
        return;
    }
    struct Mapping { void* p; size_t n; };
    struct Mappings { std::vector<Mapping> v; ~Mappings() { for (Mapping& m : v) munmap(m.p, m.n); } } maps;
    if (opts.shard_mode == 2) {
        // the shares are mapped, not copied: the merge below reads them straight from the page cache
        std::vector<char> seen(gh.size(), 0);
        for (int r = 0; r < sh_n; ++r) {
            const std::string fn = opts.shard_prefix + std::to_string(r);
            const int fd = open(fn.c_str(), O_RDONLY);
            if (fd < 0) throw std::runtime_error("valence_b200: table share " + fn + " is missing");
            struct stat sb;
            void* mp = MAP_FAILED;
            if (fstat(fd, &sb) == 0 && sb.st_size >= (off_t)SHARD_HEAD) mp = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            close(fd);
            if (mp == MAP_FAILED) throw std::runtime_error("valence_b200: cannot map the table share " + fn);
            maps.v.push_back({mp, (size_t)sb.st_size});
            const char* buf = static_cast<const char*>(mp);
            const size_t bsz = (size_t)sb.st_size;
            uint64_t head[4];
            std::memcpy(head, buf, sizeof head);
            if (head[0] != SHARD_MAGIC || head[1] != gh.size() || SHARD_HEAD + head[2] * sizeof(ShardRec) > bsz)
                throw std::runtime_error("valence_b200: table share " + fn + " does not belong to this build");
            const ShardRec* recs = reinterpret_cast<const ShardRec*>(buf + SHARD_HEAD);
            const size_t nrec = (size_t)head[2], base = SHARD_HEAD + nrec * sizeof(ShardRec);
            for (size_t k = 0; k < nrec; ++k) {
                const ShardRec& rc = recs[k];
                if (rc.i >= gh.size() || seen[rc.i]) throw std::runtime_error("valence_b200: table shares overlap");
                seen[rc.i] = 1;
                PGOut& o = outs[rc.i];
                const char* p = buf + base + rc.off;
                if (base + rc.off + rc.n_pairs * sizeof(int) + rc.n_sps * sizeof(SPRec) + rc.n_pps * sizeof(PrimPair) + rc.n_d * sizeof(double) > bsz)
                    throw std::runtime_error("valence_b200: table share " + fn + " is truncated");
                o.v_pps = reinterpret_cast<const PrimPair*>(p); o.n_pps = rc.n_pps; p += rc.n_pps * sizeof(PrimPair);
                o.v_sps = reinterpret_cast<const SPRec*>(p); o.n_sps = rc.n_sps; p += rc.n_sps * sizeof(SPRec);
                o.v_dmat = reinterpret_cast<const double*>(p); o.n_d = rc.n_d; p += rc.n_d * sizeof(double);
                o.v_pairs = reinterpret_cast<const int*>(p); o.n_pairs = rc.n_pairs;
                o.pg = rc.pg; o.wmax = rc.wmax; o.used = true;
            }
        }
    }
    const double tt2 = tnow();
    // merge: offsets by a serial scan, the copies (and the magnitude sort of the flat lists) on the host cores
    struct Off { size_t pairs, sps, pps, d; int pg; };
    std::vector<Off> offs(outs.size());
    {
        Off run = {0, 0, 0, 0, 0};
        for (size_t i = 0; i < outs.size(); ++i) {
            offs[i] = run;
            const PGOut& o = outs[i];
            if (!o.used) continue;
            run.pairs += o.n_pairs; run.sps += o.n_sps; run.pps += o.n_pps; run.d += o.n_d; run.pg += 1;
            if (i < n_free_gh) { ts.n_free_pg = run.pg; ts.n_free_pairs = run.pairs / 2; ts.n_free_sps = run.sps; ts.n_free_pps = run.pps; ts.n_free_d = run.d; }
        }
        ts.pg_pairs.resize(run.pairs); ts.sps.resize(run.sps); ts.pps.resize(run.pps); ts.dmat.resize(run.d); ts.pgs.resize(run.pg);
        if (flat) ts.pps_flat.resize(run.pps);
        else ts.pps_flat.clear();
    }
    {
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                const size_t i0 = next.fetch_add(64);
                if (i0 >= outs.size()) break;
                for (size_t i = i0; i < std::min(outs.size(), i0 + 64); ++i) {
                    PGOut& o = outs[i];
                    if (!o.used) continue;
                    const Off& f = offs[i];
                    PGDesc pg = o.pg;
                    const int pp0 = (int)f.pps, sp0 = (int)f.sps;
                    pg.pair_beg = (int)(f.pairs / 2);
                    pg.d_off = (long long)f.d;
                    for (int t = 0; t <= NPTYPE; ++t) { pg.pp_beg[t] += pp0; pg.sp_beg[t] += sp0; }
                    std::copy(o.v_pairs, o.v_pairs + o.n_pairs, ts.pg_pairs.begin() + f.pairs);
                    std::copy(o.v_sps, o.v_sps + o.n_sps, ts.sps.begin() + f.sps);
                    for (size_t k = 0; k < o.n_sps; ++k) ts.sps[f.sps + k].pp_beg += pp0;
                    std::copy(o.v_pps, o.v_pps + o.n_pps, ts.pps.begin() + f.pps);
                    std::copy(o.v_dmat, o.v_dmat + o.n_d, ts.dmat.begin() + f.d);
                    if (flat) {   // one magnitude-sorted list per pair type (each primitive carries its shell pair's e-offset)
                        std::vector<int> idx;
                        for (int t = 0; t < NPTYPE; ++t) {
                            const int b0 = pg.pp_beg[t] - pp0, b1 = pg.pp_beg[t + 1] - pp0;
                            idx.resize(b1 - b0);
                            for (int k = 0; k < b1 - b0; ++k) idx[k] = b0 + k;
                            std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return o.v_pps[x].w > o.v_pps[y].w; });
                            for (int k = 0; k < b1 - b0; ++k) ts.pps_flat[f.pps + b0 + k] = o.v_pps[idx[k]];
                        }
                    }
                    ts.pgs[f.pg] = pg;
                    std::vector<int>().swap(o.pairs); std::vector<double>().swap(o.dmat); std::vector<PrimPair>().swap(o.pps);
                }
            }
        };
        int nthr = (int)std::thread::hardware_concurrency();
        if (const char* e = std::getenv("VB_HOST_THREADS")) nthr = std::atoi(e);
        nthr = std::max(1, std::min(nthr, 32));
        if (outs.size() < 256) nthr = 1;
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
        work();
        for (std::thread& t : pool) t.join();
    }
    for (size_t i = 0; i < outs.size(); ++i) {
        const PGOut& o = outs[i];
        if (!o.used) continue;
        const PGDesc& pg = ts.pgs[offs[i].pg];
        ts.max_npp = std::max(ts.max_npp, pg.pp_beg[NPTYPE] - pg.pp_beg[0]);
        ts.max_nsp = std::max(ts.max_nsp, (int)o.n_sps);
        int ks[NPTYPE] = {0, 0, 0, 0, 0, 0};
        for (size_t k = 0; k < o.n_sps; ++k) ks[o.v_sps[k].type] += pt_ne(o.v_sps[k].type);
        for (int t = 0; t < NPTYPE; ++t) ts.max_ks = std::max(ts.max_ks, ks[t]);
        ts.max_ne = std::max(ts.max_ne, pg.ne);
        ts.max_np = std::max(ts.max_np, pg.np);
    }
    for (const GShell& s : bas.shells) ts.lmax = std::max(ts.lmax, s.l);
    if (dbg_t) std::printf("[setup] groups+pass0 %.1f ms, pass1 %.1f ms, merge %.1f ms\n", tt1 - tt0, tt2 - tt1, tnow() - tt2);
}

}  // namespace vb
