#include "vb_setup.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

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
static double binom(int n, int k)
{
    double r = 1.0;
    for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
    return r;
}

// Fold the horizontal recurrence into a cartesian pair density d[a][b]
// (shell A carries la >= lb):  sum_ab d[a][b] (ab| = sum_e dt[e] (e0|,
//   (x-B)^b = sum_k C(b,k) (A-B)^(b-k) (x-A)^k   per cartesian direction.
static void hrr_fold(int la, int lb, const double* AB, const double* d /*[na][nb]*/, double* dt /*[pt_ne]*/)
{
    const int na = ncart(la), nb = ncart(lb), e0 = coff(la);
    for (int a = 0; a < na; ++a) {
        int ca = coff(la) + a, ax = c_lx(ca), ay = c_ly(ca), az = c_lz(ca);
        for (int b = 0; b < nb; ++b) {
            double v = d[a * nb + b];
            if (v == 0.0) continue;
            int cb = coff(lb) + b, bx = c_lx(cb), by = c_ly(cb), bz = c_lz(cb);
            for (int kx = 0; kx <= bx; ++kx)
                for (int ky = 0; ky <= by; ++ky)
                    for (int kz = 0; kz <= bz; ++kz) {
                        double f = binom(bx, kx) * binom(by, ky) * binom(bz, kz) * std::pow(AB[0], bx - kx) *
                                   std::pow(AB[1], by - ky) * std::pow(AB[2], bz - kz);
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
                 double tau, TileSetup* out)
{
    (void)in;
    TileSetup& ts = *out;
    ts = TileSetup();
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
    // --- pair groups --------------------------------------------------------
    const double SQ2PI54 = std::sqrt(2.0) * std::pow(PI, 1.25);
    const int ng = (int)ts.groups.size();
    struct TmpSP {
        int type, A, B;
        bool swapped;
        double wmax;
        std::vector<PrimPair> pp;
        std::vector<double> dt;   // folded density [nE][np]
    };
    std::vector<double> dcart(36), dfold(64);
    // one pair group; pass 0 only measures the largest primitive weight, pass 1 emits
    auto do_pair_group = [&](int g, int h, int pass, double wcut) {
        const EntryGroup& G = ts.groups[g];
        const EntryGroup& H = ts.groups[h];
        std::vector<int> pairs;
        for (int s : G.entries)
            for (int t : H.entries) {
                if (wf.sym && g == h && t > s) continue;
                pairs.push_back(s);
                pairs.push_back(t);
            }
        const int np = (int)pairs.size() / 2;
        if (np == 0) return;
        std::vector<TmpSP> tsp;
        for (int X : G.shells)
            for (int Y : H.shells) {
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
                // folded densities of every orbital pair on this shell pair
                sp.dt.assign((size_t)nE * np, 0.0);
                double dmaxL[EMAX + 1] = {0, 0, 0, 0, 0};   // max |folded density| per angular momentum of e
                bool any = false;
                for (int ip = 0; ip < np; ++ip) {
                    int s = pairs[2 * ip], t = pairs[2 * ip + 1];
                    const ExpOrb& ob = orbs[wf.bra[wf.slot(s, 0)]];   // bra-side orbital of the pair
                    const ExpOrb& ok = orbs[wf.ket[wf.slot(t, 0)]];   // ket-side orbital
                    // original orientation: X from group g (orbital ob), Y from group h (orbital ok)
                    const OrbShell* cA = find_shell(sp.swapped ? ok : ob, sp.A);
                    const OrbShell* cB = find_shell(sp.swapped ? ob : ok, sp.B);
                    if (!cA || !cB) continue;
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
                for (int ia = 0; ia < sa.nprim; ++ia)
                    for (int ib = 0; ib < sb.nprim; ++ib) {
                        double a = bas.exps[sa.prim_off + ia], b = bas.exps[sb.prim_off + ib], p = a + b;
                        double K = bas.coefs[sa.prim_off + ia] * bas.coefs[sb.prim_off + ib] * std::exp(-a * b / p * AB2) * SQ2PI54;
                        if (K == 0.0) continue;
                        PrimPair pp;
                        pp.Px = (a * sa.r[0] + b * sb.r[0]) / p;
                        pp.Py = (a * sa.r[1] + b * sb.r[1]) / p;
                        pp.Pz = (a * sa.r[2] + b * sb.r[2]) / p;
                        pp.p = p;
                        pp.ip = 1.0 / p;
                        pp.Kp = K / p;
                        pp.PAx = pp.Px - sa.r[0]; pp.PAy = pp.Py - sa.r[1]; pp.PAz = pp.Pz - sa.r[2];
                        // Schwarz-type magnitude: the charge cloud sum_e dt[e] [e0| of this primitive pair has
                        // self-repulsion <= w^2, so its share of any (st|uv) is <= w_p w_q.
                        // [ss|ss]_pp = Kp^2 / sqrt(2p); each unit of angular momentum adds a factor
                        // <= |PA| + 1/sqrt(p) (generous: includes the 1/2p and 1/4p terms).
                        double len = std::sqrt(pp.PAx * pp.PAx + pp.PAy * pp.PAy + pp.PAz * pp.PAz) + 1.0 / std::sqrt(p);
                        double poly = 0.0, lp = 1.0;
                        for (int L = 0; L <= EA; ++L) { if (L >= sa.l) poly += dmaxL[L] * lp * ncart(L); lp *= 2.0 * len; }
                        pp.w = std::fabs(pp.Kp) * std::pow(2.0 * p, -0.25) * poly;
                        if (pass == 1 && !(pp.w * wcut >= tau)) continue;
                        sp.wmax = std::max(sp.wmax, pp.w);
                        if (pass == 1) sp.pp.push_back(pp);
                    }
                if (pass == 0) { ts.wmax = std::max(ts.wmax, sp.wmax); continue; }
                if (sp.pp.empty()) continue;
                std::stable_sort(sp.pp.begin(), sp.pp.end(), [](const PrimPair& x, const PrimPair& y) { return x.w > y.w; });
                tsp.push_back(std::move(sp));
            }
        if (pass == 0 || tsp.empty()) return;
        // by type, strongest shell pairs first
        std::stable_sort(tsp.begin(), tsp.end(), [](const TmpSP& a, const TmpSP& b) {
            return a.type != b.type ? a.type < b.type : a.wmax > b.wmax;
        });
        PGDesc pg;
        std::memset(&pg, 0, sizeof pg);
        pg.g = g; pg.h = h; pg.np = np; pg.pair_beg = (int)ts.pg_pairs.size() / 2;
        ts.pg_pairs.insert(ts.pg_pairs.end(), pairs.begin(), pairs.end());
        int ne = 0;
        for (const TmpSP& sp : tsp) ne += pt_ne(sp.type);
        pg.ne = ne;
        pg.d_off = (long long)ts.dmat.size();
        ts.dmat.resize(ts.dmat.size() + (((size_t)ne * np + 1) & ~(size_t)1), 0.0);   // keep blocks 16-byte aligned
        double* D = ts.dmat.data() + pg.d_off;
        int t_cur = 0, eoff = 0;
        const int sp_base = (int)ts.sps.size();
        std::vector<SPRec> recs;
        for (size_t k = 0; k < tsp.size(); ++k) {
            const TmpSP& sp = tsp[k];
            while (t_cur <= sp.type) { pg.pp_beg[t_cur] = (int)ts.pps.size(); ++t_cur; }
            const int nE = pt_ne(sp.type);
            recs.push_back({sp.type, eoff, (int)ts.pps.size(), (int)sp.pp.size(), sp.wmax, 0.0});
            pg.kwmax[sp.type] = std::max(pg.kwmax[sp.type], sp.wmax);
            for (PrimPair pp : sp.pp) { pp.eoff = eoff; pp.pad = 0; pp.wseg = sp.wmax; ts.pps.push_back(pp); }
            std::copy(sp.dt.begin(), sp.dt.end(), D + (size_t)eoff * np);
            eoff += nE;
        }
        while (t_cur <= NPTYPE) { pg.pp_beg[t_cur] = (int)ts.pps.size(); ++t_cur; }
        // per type, most expensive shell pairs first (they are dealt round-robin to the warps)
        std::stable_sort(recs.begin(), recs.end(), [](const SPRec& a, const SPRec& b) {
            return a.type != b.type ? a.type < b.type : a.pp_cnt > b.pp_cnt;
        });
        ts.sps.insert(ts.sps.end(), recs.begin(), recs.end());
        {
            int k = 0;
            for (int t = 0; t <= NPTYPE; ++t) {
                while (k < (int)recs.size() && recs[k].type < t) ++k;
                pg.sp_beg[t] = sp_base + k;
            }
        }
        ts.max_npp = std::max(ts.max_npp, pg.pp_beg[NPTYPE] - pg.pp_beg[0]);
        ts.max_nsp = std::max(ts.max_nsp, (int)recs.size());
        ts.max_ne = std::max(ts.max_ne, ne);
        ts.max_np = std::max(ts.max_np, np);
        ts.pgs.push_back(pg);
    };
    // pass 0: the largest weights live in the one-group pair groups
    for (int g = 0; g < ng; ++g) do_pair_group(g, g, 0, 0.0);
    const double wcut = ts.wmax;
    for (int g = 0; g < ng; ++g)
        for (int h = 0; h < (wf.sym ? g + 1 : ng); ++h) do_pair_group(g, h, 1, wcut);
    for (const GShell& s : bas.shells) ts.lmax = std::max(ts.lmax, s.l);
}

}  // namespace vb
