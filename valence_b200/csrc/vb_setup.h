// Host-side set-up of one VSVB energy evaluation: normalised basis, expanded
// orbitals, wavefunction entry lists, orbital-pair groups, shell-pair /
// primitive-pair tables and HRR-folded pair densities for the GPU kernels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vb_eri.cuh"
#include "vb_input.h"

namespace vb {

constexpr double ANGS2BOHR = 1.889725987722;   // /root/reference/src/tools_module.F90:11

struct GShell {
    int l, atom, nprim, ao_off, prim_off, type_shell;   // type_shell: index into the per-type shell tables
    double r[3];
};

struct Basis {
    std::vector<GShell> shells;
    std::vector<int> atom_first_shell;          // natom + 1
    std::vector<double> exps, coefs;            // normalised contraction weights (norm_prim)
    std::vector<double> angn;                   // per cumulative cartesian index, l <= LMAX_SHELL
    int nao = 0;
};

// An orbital expanded over global shells.  c = scattered LCAO weights without
// the angular factor (what the reference holds in coeffi(:), valence.F90:3221-3237)
struct OrbShell {
    int gshell;
    double c[6];
};
struct ExpOrb {
    std::vector<OrbShell> sh;
};

// Device-friendly tables ----------------------------------------------------
struct SPRec {          // one oriented shell pair of a pair group (whole contraction)
    int type;           // ptype(la, lb)
    int eoff;           // first [e0| component of this shell pair inside the group's e-space
    int pp_beg, pp_cnt; // range in the primitive-pair array
    double wmax;        // largest primitive weight of the shell pair
    double pad;         // s / p segments: largest 1/p of the segment (sort key of the list); 32 bytes: staged with cp.async.bulk
};
struct PGDesc {
    int sp_beg[NPTYPE + 1];     // shell pairs sorted by type
    int pp_beg[NPTYPE + 1];     // primitive pairs sorted by type, then shell pair (ket lanes walk these)
    int e_beg[NPTYPE + 1];      // first e-row of each pair type inside the density block (rows of a type are contiguous)
    long long d_off;            // offset of the folded density block [ne][np]
    int ne, np;                 // # e-components, # orbital pairs
    int pair_beg;               // offset into the pair list (s,t)
    int g, h;                   // entry groups
    double smax;                // max Schwarz value over the pairs (filled after the diagonal pass)
    double kwmax[NPTYPE];       // largest primitive weight per pair type
};

struct EntryGroup {
    std::vector<int> entries;   // wavefunction entries (0-based)
    std::vector<int> shells;    // global shells spanned
    int nao = 0;
};

struct Wavefunction {           // bra/ket orbital lists (valence.F90:324-336, 557-605)
    std::vector<int> bra, ket;  // orbital id per electron slot (0-based ids)
    int nnd = 0, nso = 0;       // # single-slot entries, # entries
    bool sym = true;            // bra == ket (ijkl symmetry enabled)
    int subject = -1;           // entry index of the subject orbital slot (first_order_opt), or -1
    // entry -> slots
    int nslots(int s) const { return s < nnd ? 1 : 2; }
    int slot(int s, int k) const { return s < nnd ? s : 2 * s - nnd + k; }   // 0-based
};

struct TileSetup {
    std::vector<EntryGroup> groups;
    std::vector<PGDesc> pgs;
    std::vector<int> pg_pairs;      // 2 ints (s,t) per pair
    std::vector<SPRec> sps;
    std::vector<PrimPair> pps;      // grouped by shell pair
    std::vector<PrimPair> pps_flat; // same ranges per (pair group, type), sorted by magnitude (flat mode only)
    double wmax = 0.0;              // largest primitive-pair magnitude bound (for pruning)
    std::vector<double> dmat;       // folded densities, per pair group [e][p]
    int max_ne = 0, max_np = 0, max_npp = 0, max_nsp = 0;
    int max_ks = 0;                 // widest e-range of one pair type inside a pair group
    int lmax = 0;
    int n_free_pg = 0;              // TileOpts::isolate: pair groups that do not touch the isolated entry (they come first)
    size_t n_free_pairs = 0, n_free_sps = 0, n_free_pps = 0, n_free_d = 0;   // sizes of their tables
};

// Options of build_tiles for first_order_opt's integral cache (vb_engine.cu: first_order).
//   isolate      : this entry forms a group of its own, created last; the pair groups touching it ("subject"
//                  pair groups) are emitted after all the others ("free" pair groups, independent of the
//                  orbitals of that entry)
//   only_subject : emit the subject pair groups only (offsets start at 0; the caller shifts them behind its
//                  resident free tables)
//   wcut         : > 0 -> upper bound of the primitive weights to prune against (instead of this call's own
//                  maximum), so that tables built in separate calls are pruned consistently
struct TileOpts {
    int isolate = -1;
    bool only_subject = false;
    double wcut = 0.0;
    bool measure_only = false;      // only compute TileSetup::wmax (of the isolated group alone with only_subject)
    // Table build shared by the ranks of one node (one process per GPU builds the same tables: N ranks on one host would
    // do the same work N times).  shard_mode 1: build this rank's share of the pair groups and publish it as the file
    // shard_prefix + rank (shared memory, e.g. /dev/shm/...), nothing else; shard_mode 2: take every pair group from the
    // published files of all ranks and merge -- tables bitwise identical to an unsharded build.  The caller
    // synchronises the ranks between the two calls.
    // shard_mode 3: build this rank's share only and merge it alone (local offsets): the shares are exchanged on the DEVICE
    // (NCCL all-gather over NVLink, vb_engine.cu: gather_tables) instead of through host shared memory.
    int shard_mode = 0, shard_rank = 0, shard_nranks = 1;
    std::string shard_prefix;
};

double dblfac(int n);
void norm_prim(int l, int n, const double* exps, const double* raw, double* out);   // valence.F90:2264-2305
Basis build_basis(const Input& in, const std::vector<double>& xyz_bohr);
double nuclear_repulsion(const Input& in, const std::vector<double>& xyz_bohr);      // tools_module.F90:19-40

// scatter of one orbital's weights over its OBS (valence.F90:2919-2932, ndf2obs :2191-2247)
ExpOrb expand_orbital(const Input& in, const Basis& bas, const std::vector<std::vector<double>>& coeff, int orb);
// global shell and component of OBS position `pos` (1-based) of orbital `orb`
bool obs_position(const Input& in, const Basis& bas, int orb, int pos, int* gshell, int* comp);

// tau: primitive pairs whose largest possible contribution to any orbital-level integral stays
// below tau are dropped at set-up (0 keeps everything the reference computes)
// flat: also emit pps_flat, the primitive pairs of a pair group sorted by magnitude inside each pair
// type across shell pairs (ket side of k_ptile)
void build_tiles(const Input& in, const Basis& bas, const Wavefunction& wf, const std::vector<ExpOrb>& orbs2e,
                 double tau, bool flat, TileSetup* out, const TileOpts& opts = TileOpts());

}  // namespace vb
