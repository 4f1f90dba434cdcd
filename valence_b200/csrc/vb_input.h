// Input data model and reader for VALENCE input files.
// Grammar: /root/reference/src/xm_module.F90:41-42 (header), :106-108 (control),
// :119 (geometry), :133-163 (basis), :181-202 (couplings, excitations),
// :212-287 (orbitals); derived sizes valence_initialize_module.F90:89-92,111-113.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

namespace vb {

struct ShellDef {
    int l = 0;
    std::vector<double> exps, raw;   // raw = contraction weights as read
};
struct AtomTypeDef {
    double charge = 0.0;
    std::vector<ShellDef> shells;
};
struct OrbitalDef {
    std::vector<int> atoms;          // 1-based atom indices: the orbital basis set (OBS)
    std::vector<int> xp;             // AO index inside the OBS (1-based); <=0 names a DBF
    std::vector<double> coeff;
};

struct Input {
    int natom = 0, natom_t = 0, npair = 0, nunpd = 0, ndocc = 0, totlen = 0, xpmax = 0, nspinc = 0;
    int num_sh = 0, num_pr = 0, nang = 0, ndf = 0, nset = 0, nxorb = 0, mxctr = 0;
    int ntol_c = 0, ntol_d = 0, ntol_i = 0, ntol_e_min = 0, ntol_e_max = 0, max_iter = 0;
    double ptbnmax = 0.0, feather = 0.0;
    std::vector<int> orbset;                 // 2*nset
    std::vector<int> atom_t;                 // 1-based type per atom
    std::vector<double> coords;              // 3*natom, Angstrom as read
    std::vector<AtomTypeDef> types;
    std::vector<double> coeff_sc;            // nspinc (>=1 entry)
    std::vector<int> pair_sc;                // [isc][pair][2], 1-based orbital labels
    std::vector<int> xorb, root;
    std::vector<OrbitalDef> orbitals;        // 2*npair, nunpd, ndocc, ndf in that order

    int nelec() const { return 2 * npair + 2 * ndocc + nunpd; }
    int norbs() const { return 2 * npair + ndocc + nunpd + ndf; }
    int nalpha() const { return npair + nunpd + ndocc; }
    int nbeta() const { return npair + ndocc; }
    int nnd() const { return 2 * npair + nunpd; }
    int pair(int isc, int k, int s) const { return pair_sc[((size_t)isc * npair + k) * 2 + s]; }
};

struct InputError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

Input parse_input_text(const std::string& text);
Input parse_input_file(const std::string& path);

}  // namespace vb
