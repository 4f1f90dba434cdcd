// Host check of the product's input reader (vb_input.cpp): every file named on the command line is parsed and dumped as one
// canonical token line (integers, doubles as %.17g) that tests/test_host_math.py compares with the Python data model
// (valence_b200/inputs.py, itself pinned to the reference's files by tests/golden/make_golden.py).  A file the reader
// refuses prints "ERROR <message>".
// (reference reader: /root/reference/src/xm_module.F90:41-287)
#include <cstdio>

#include "vb_input.h"

static void pi(long long v) { std::printf(" %lld", v); }
static void pd(double v) { std::printf(" %.17g", v); }

int main(int argc, char** argv)
{
    for (int a = 1; a < argc; ++a) {
        vb::Input in;
        try {
            in = vb::parse_input_file(argv[a]);
        } catch (const std::exception& e) {
            std::printf("ERROR %s\n", e.what());
            continue;
        }
        std::printf("OK");
        for (int v : {in.natom, in.natom_t, in.npair, in.nunpd, in.ndocc, in.totlen, in.xpmax, in.nspinc, in.num_sh, in.num_pr, in.nang,
                      in.ndf, in.nset, in.nxorb, in.mxctr, in.ntol_c, in.ntol_d, in.ntol_i, in.ntol_e_min, in.ntol_e_max, in.max_iter})
            pi(v);
        pd(in.ptbnmax); pd(in.feather);
        pi((long long)in.orbset.size());
        for (int v : in.orbset) pi(v);
        for (int v : in.atom_t) pi(v);
        for (double v : in.coords) pd(v);
        for (const auto& t : in.types) {
            pd(t.charge); pi((long long)t.shells.size());
            for (const auto& s : t.shells) {
                pi(s.l); pi((long long)s.exps.size());
                for (size_t k = 0; k < s.exps.size(); ++k) { pd(s.exps[k]); pd(s.raw[k]); }
            }
        }
        pi((long long)in.coeff_sc.size());
        for (double v : in.coeff_sc) pd(v);
        pi((long long)in.pair_sc.size());
        for (int v : in.pair_sc) pi(v);
        pi((long long)(in.xorb.size() + in.root.size()));
        for (size_t k = 0; k < in.xorb.size(); ++k) { pi(in.xorb[k]); pi(in.root[k]); }
        pi((long long)in.orbitals.size());
        for (const auto& o : in.orbitals) {
            pi((long long)o.atoms.size());
            for (int v : o.atoms) pi(v);
            pi((long long)o.xp.size());
            for (size_t k = 0; k < o.xp.size(); ++k) { pi(o.xp[k]); pd(o.coeff[k]); }
        }
        // derived sizes (valence_initialize_module.F90:89-92)
        pi(in.nelec()); pi(in.norbs()); pi(in.nalpha()); pi(in.nbeta()); pi(in.nnd());
        std::printf("\n");
    }
    return 0;
}
