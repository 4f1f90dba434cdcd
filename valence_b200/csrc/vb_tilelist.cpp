// Tile list and work items of the fused ERI + contraction pass: host-only code (no CUDA types), shared by the
// energy pass, the integral cache of first_order_opt and the CPU tests (tests/host/test_tilelist_host.cpp).
#include "vb_tilelist.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <thread>

namespace vb {

// Tiles (a, b), a in avec, b in bvec (both ascending pair-group indices), b <= a, smax_a * smax_b > itol; ordered in
// blocks of bra pair groups against chunks of ket pair groups (L2 residency), partners by decreasing Schwarz bound.
// runs: (first tile, # tiles) of every non-empty (a, chunk).
void make_tile_list(const std::vector<PGDesc>& pgs, const std::vector<int>& avec, const std::vector<int>& bvec, double itol,
                    std::vector<TilePair>* tl, std::vector<std::pair<long long, int>>* runs, int rank, int nranks)
{
    tl->clear(); runs->clear();
    // bra block: 128 pair groups on one GPU (L2 residency of the block's tables); the ranks of a multi-GPU run own whole blocks,
    // so theirs are smaller -- the snake deal evens out the triangular growth, the block count the cost differences between
    // strong and weak pair groups (VB_TILE_PB overrides)
    int PB = nranks > 1 ? 32 : 128;       // 8 x B200, (H2O)_256: tile pass 586 ms with blocks of 64, 550 with 32, 548 with 16, 582 with 8
    if (const char* e = std::getenv("VB_TILE_PB")) PB = std::max(1, std::atoi(e));
    const int QC = 1024;
    const int na = (int)avec.size(), nb = (int)bvec.size();
    const int nchunks = (nb + QC - 1) / QC;
    std::vector<int> order(bvec);
    std::vector<int> cmin(nchunks);
    for (int c = 0; c < nchunks; ++c) {
        cmin[c] = bvec[(size_t)c * QC];
        std::stable_sort(order.begin() + (size_t)c * QC, order.begin() + std::min<size_t>(nb, (size_t)(c + 1) * QC),
                         [&](int x, int y) { return pgs[x].smax > pgs[y].smax; });
    }
    const int nblocks_all = (na + PB - 1) / PB;
    // Owner of every bra block.  One rank: trivial.  Several: the blocks are dealt by estimated cost, heaviest first, each to the
    // rank with the least work so far (every rank computes the same deal from the same descriptors).  Cost of pair group a =
    // (its primitive pairs) x (primitive pairs of the partners that pass the Schwarz bound), the partners counted per ket chunk by a
    // binary search in the chunk's sorted bounds (the chunk that contains a itself pro rata).  A plain snake deal of the blocks left
    // the 8-GPU tile pass of (H2O)_256 8 % above 1/8 of the single-GPU pass.
    std::vector<int> owner_of(nblocks_all, 0);
    if (nranks > 1) {
        auto npp = [&](int x) { return (double)std::max(1, pgs[x].pp_beg[NPTYPE] - pgs[x].pp_beg[0]); };
        std::vector<double> pf((size_t)nb + nchunks + 1, 0.0);     // per chunk: prefix sums of the partners' primitive pairs in sorted order
        std::vector<size_t> pf0(nchunks);
        for (int c = 0, o = 0; c < nchunks; ++c) {
            const int c0 = c * QC, c1 = std::min(nb, c0 + QC);
            pf0[c] = (size_t)o;
            pf[o++] = 0.0;
            for (int j = c0; j < c1; ++j, ++o) pf[o] = pf[o - 1] + npp(order[j]);
        }
        std::vector<double> bcost(nblocks_all, 0.0);
        for (int blk = 0; blk < nblocks_all; ++blk) {
            const int B0 = blk * PB, B1 = std::min(na, B0 + PB);
            double cost = 0.0;
            for (int ai = B0; ai < B1; ++ai) {
                const int a = avec[ai];
                const double sa = pgs[a].smax;
                if (!(sa > 0.0)) continue;
                double part = 0.0;
                for (int c = 0; c < nchunks; ++c) {
                    if (cmin[c] > a) break;
                    const int c0 = c * QC, c1 = std::min(nb, c0 + QC);
                    int lo = c0, hi = c1;                      // first j with sa * smax <= itol
                    while (lo < hi) { const int mid = (lo + hi) / 2; if (sa * pgs[order[mid]].smax > itol) lo = mid + 1; else hi = mid; }
                    double w = pf[pf0[c] + (size_t)(lo - c0)];
                    if (bvec[c1 - 1] > a) w *= (double)(a - cmin[c] + 1) / (double)(bvec[c1 - 1] - cmin[c] + 1);   // a's own chunk
                    part += w;
                }
                cost += npp(a) * part;
            }
            bcost[blk] = cost;
        }
        std::vector<int> byc(nblocks_all);
        for (int i = 0; i < nblocks_all; ++i) byc[i] = i;
        std::stable_sort(byc.begin(), byc.end(), [&](int x, int y) { return bcost[x] > bcost[y]; });
        std::vector<double> load(nranks, 0.0);
        for (int blk : byc) {
            int best = 0;
            for (int r = 1; r < nranks; ++r) if (load[r] < load[best]) best = r;
            owner_of[blk] = best;
            load[best] += bcost[blk];
        }
    }
    auto owner = [&](int blk) { return owner_of[blk]; };
    // one bra block per task on the host cores; the blocks are concatenated in order, so the list does not depend on
    // the thread count
    const int nblocks = (na + PB - 1) / PB;
    std::vector<std::vector<TilePair>> btl(nblocks);
    std::vector<std::vector<std::pair<long long, int>>> bruns(nblocks);
    auto do_block = [&](int blk) {
        if (nranks > 1 && owner(blk) != rank) return;
        const int B0 = blk * PB, B1 = std::min(na, B0 + PB);
        std::vector<TilePair>& t = btl[blk];
        std::vector<std::pair<long long, int>>& r = bruns[blk];
        for (int c = 0; c < nchunks; ++c) {
            if (cmin[c] > avec[B1 - 1]) break;
            const int c0 = c * QC, c1 = std::min(nb, c0 + QC);
            for (int ai = B0; ai < B1; ++ai) {
                const int a = avec[ai];
                if (a < cmin[c]) continue;
                const double sa = pgs[a].smax;
                const long long start = (long long)t.size();
                for (int j = c0; j < c1; ++j) {
                    const int b = order[j];
                    if (!(sa * pgs[b].smax > itol)) break;
                    if (b <= a) t.push_back(TilePair{a, b});
                }
                if ((long long)t.size() > start) r.emplace_back(start, (int)((long long)t.size() - start));
            }
        }
    };
    {
        int nthr = (int)std::thread::hardware_concurrency();
        if (const char* e = std::getenv("VB_HOST_THREADS")) nthr = std::atoi(e);
        nthr = std::max(1, std::min(nthr, 32));
        if (nblocks < 8) nthr = 1;
        std::atomic<int> next{0};
        auto work = [&]() { for (int blk; (blk = next.fetch_add(1)) < nblocks;) do_block(blk); };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
        work();
        for (std::thread& t : pool) t.join();
    }
    size_t total = 0, nruns = 0;
    for (int blk = 0; blk < nblocks; ++blk) { total += btl[blk].size(); nruns += bruns[blk].size(); }
    tl->reserve(total); runs->reserve(nruns);
    for (int blk = 0; blk < nblocks; ++blk) {
        const long long off = (long long)tl->size();
        tl->insert(tl->end(), btl[blk].begin(), btl[blk].end());
        for (const auto& r : bruns[blk]) runs->emplace_back(r.first + off, r.second);
        std::vector<TilePair>().swap(btl[blk]);
    }
    if ((long long)tl->size() > 2000000000LL) throw std::runtime_error("valence_b200: tile list too long");
}

// work items of k_ptile: pieces of <= m tiles of a run, block-cyclic over the ranks; z = slot of the item's first tile
void make_items(const std::vector<std::pair<long long, int>>& runs, long long ntiles, int nsm, int rank, int nranks,
                std::vector<WorkItem>* itl, long long* my_tiles)
{
    itl->clear();
    *my_tiles = 0;
    const long long m = std::max<long long>(1, std::min<long long>(TILES_PER_ITEM_MAX, ntiles / ((long long)nsm * nranks * 16)));
    long long idx = 0;
    for (const auto& run : runs)
        for (long long k = run.first; k < run.first + run.second; k += m, ++idx) {
            if (idx % nranks != rank) continue;
            const int cnt = (int)std::min<long long>(m, run.first + run.second - k);
            itl->push_back(WorkItem{(int)k, cnt, (int)*my_tiles, 0});
            *my_tiles += cnt;
        }
}

}  // namespace vb
