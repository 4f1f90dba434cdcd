// Tile list and work items of the fused ERI + contraction pass (host side).
#pragma once
#include <utility>
#include <vector>

#include "vb_setup.h"

namespace vb {

struct alignas(8) TilePair { int x, y; };            // (bra pair group, ket pair group): layout of CUDA's int2
struct alignas(16) WorkItem { int x, y, z, w; };     // first tile, # tiles, slot of the first tile among this rank's tiles, -: int4
constexpr int TILES_PER_ITEM_MAX = 8;                // = PT_MAXQ of k_ptile (checked in vb_engine.cu)

// Tiles (a, b), a in avec, b in bvec (ascending pair-group indices), b <= a, smax_a * smax_b > itol (the reference's
// Schwarz screen at pair-group level, valence.F90:1189-1190); ordered in blocks of bra pair groups against chunks of
// ket pair groups (L2 residency), partners by decreasing Schwarz bound.  runs: (first tile, # tiles) of every
// non-empty (a, chunk).  The result does not depend on the number of host threads.
// rank / nranks: only the bra blocks of this rank are generated (blocks dealt in a snake over the ranks, which evens out the
// triangular growth of the block sizes): every rank builds, uploads and walks 1/N of the list.  The union over the ranks is the
// full list; make_items is then called with (0, 1).
void make_tile_list(const std::vector<PGDesc>& pgs, const std::vector<int>& avec, const std::vector<int>& bvec, double itol,
                    std::vector<TilePair>* tl, std::vector<std::pair<long long, int>>* runs, int rank = 0, int nranks = 1);

// Work items: pieces of <= TILES_PER_ITEM_MAX tiles of a run (they share the bra pair group), dealt block-cyclically
// to the ranks (item k -> rank k mod nranks: the replacement of the reference's task farm, valence.F90:1162-1163).
void make_items(const std::vector<std::pair<long long, int>>& runs, long long ntiles, int nsm, int rank, int nranks,
                std::vector<WorkItem>* itl, long long* my_tiles);

}  // namespace vb
