// Cofactor ("density") data of a VSVB wavefunction in the inverse / adjugate form that replaces the
// reference's determinant-per-orbital-quartet loop (density, density_sc, dbra, dket, det, givdr:
// /root/reference/src/valence.F90:1535-2144; maths: SURVEY.md appendix B).
//
// For every determinant pair (spin coupling isc x jsc, alpha/beta assignment of each Rumer pair in
// bra and ket: valence.F90:1576-1588, 1688-1760, 1808-1869) and each spin block M = <bra|ket> the
// first- and second-order cofactors are written with a regularised inverse G plus explicit
// null-space terms, which stays finite for singular blocks (symmetry-orthogonal orbitals, the
// substituted rows/columns of first_order_opt):
//   M = U diag(sigma) V^T,  Z = {p : sigma_p ~ 0} (|Z| <= 2),  N = complement,
//   dN = det(U) det(V) prod_{p in N} sigma_p,   piZ = prod_{z in Z} sigma_z,
//   G[r][c] = sum_{p in N} U[r][p] V[c][p] / sigma_p
//   C0 = dN piZ
//   C1(r,c) = dN ( piZ G[r][c] + sum_z piZ\z  U[r][z] V[c][z] )
//   C2(r1 c1, r2 c2) = dN ( piZ (G11 G22 - G12 G21)
//                           + sum_z piZ\z ( Vz[c2]Uz[r2] G11 - Vz[c2]Uz[r1] G21 - Vz[c1]Uz[r2] G12 + Vz[c1]Uz[r1] G22 )
//                           + [|Z|=2] (Vz0[c1]Vz1[c2] - Vz1[c1]Vz0[c2]) (Uz0[r1]Uz1[r2] - Uz1[r1]Uz0[r2]) )
// All of it is bilinear in (row, column) factors, so the sums over the electron slots of a
// wavefunction entry are taken once on the host ("entry level") and the GPU contraction only
// does table look-ups.
#pragma once
#include <vector>

#include "vb_input.h"
#include "vb_setup.h"

namespace vb {

constexpr int COF_HEADER = 12;   // doubles per determinant pair before the arrays

// Packed layout per determinant pair (all doubles), nso = # wavefunction entries:
//   [0] w  [1] dN_a [2] piZ_a [3] pwo_a0 [4] pwo_a1 [5] nz_a  [6] dN_b [7] piZ_b [8] pwo_b0 [9] pwo_b1 [10] nz_b [11] pad
//   Ga[nso*nso] Gb[nso*nso] uza[2][nso] vza[2][nso] uzb[2][nso] vzb[2][nso]
inline size_t cof_stride(int nso) { return COF_HEADER + 2 * (size_t)nso * nso + 8 * (size_t)nso; }

struct CofactorSet {
    int nso = 0, ndp = 0;
    std::vector<double> data;           // ndp * cof_stride(nso)
    double min_sigma_ratio = 1.0;       // smallest kept singular value / largest, over all blocks
    int singular_blocks = 0;
};

// Se: entry-level overlaps <bra_s|ket_t>, row-major nso x nso.
// only_isc/only_jsc >= 0 restrict the sum to one (bra coupling, ket coupling) pair with unit weight
// (the spin-coupling Hamiltonian of spin_opt, valence.F90:1549-1569)
void build_cofactors(const Input& in, const Wavefunction& wf, const std::vector<double>& Se, CofactorSet* out,
                     int only_isc = -1, int only_jsc = -1);

// one- electron numerator and norm from the cofactors (valence.F90:1072-1106):
//   e1 = sum_dp w sum_st He[s][t] C1tot(s,t),  wfnorm = sum_dp w sum_st Se[s][t] C1tot(s,t) / nelec
void one_electron_from_cofactors(const CofactorSet& cs, const std::vector<double>& Se, const std::vector<double>& He,
                                 int nelec, double* e1, double* wfnorm);

}  // namespace vb
