// Host-side check of the tile list and the work-item sharding (vb_tilelist.cpp): the list is exactly the set of
// pair-group pairs that pass the Schwarz screen, it does not depend on the host thread count, and the work items of
// N ranks partition it (every tile owned by exactly one rank, items never straddle a bra pair group).
// Built and run by tests/test_host_math.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <set>
#include <vector>
#include "../../valence_b200/csrc/vb_tilelist.h"
using namespace vb;

int main()
{
    int bad = 0;
    std::mt19937_64 rng(20261017);
    std::uniform_real_distribution<double> U(-9.0, 1.0);
    for (int npg : {1, 7, 300, 2500}) {
        std::vector<PGDesc> pgs(npg);
        for (PGDesc& p : pgs) { std::memset(&p, 0, sizeof p); p.smax = std::pow(10.0, U(rng)); p.pp_beg[NPTYPE] = 1 + (int)(rng() % 400); }
        const double itol = 1e-10;
        std::vector<int> all(npg), odd, even;
        for (int i = 0; i < npg; ++i) { all[i] = i; (i % 2 ? odd : even).push_back(i); }
        struct Case { const std::vector<int>* a; const std::vector<int>* b; };
        for (const Case& cs : {Case{&all, &all}, Case{&odd, &all}, Case{&even, &even}}) {
            if (cs.a->empty()) continue;
            std::vector<TilePair> t1, t5;
            std::vector<std::pair<long long, int>> r1, r5;
            setenv("VB_HOST_THREADS", "1", 1);
            make_tile_list(pgs, *cs.a, *cs.b, itol, &t1, &r1);
            setenv("VB_HOST_THREADS", "5", 1);
            make_tile_list(pgs, *cs.a, *cs.b, itol, &t5, &r5);
            if (t1.size() != t5.size() || r1 != r5 || (!t1.empty() && std::memcmp(t1.data(), t5.data(), t1.size() * sizeof(TilePair)) != 0)) {
                std::printf("tile list depends on the thread count (npg %d)\n", npg); ++bad;
            }
            // brute force: every (a, b), b <= a, passing the screen -- exactly once
            std::set<std::pair<int, int>> want, got;
            for (int a : *cs.a) for (int b : *cs.b) if (b <= a && pgs[a].smax * pgs[b].smax > itol) want.insert({a, b});
            for (const TilePair& t : t1) if (!got.insert({t.x, t.y}).second) { std::printf("duplicate tile\n"); ++bad; break; }
            if (want != got) { std::printf("tile list differs from the screened set: %zu vs %zu (npg %d)\n", got.size(), want.size(), npg); ++bad; }
            // runs tile the list, each within one bra pair group
            long long pos = 0;
            for (const auto& r : r1) {
                if (r.first != pos || r.second <= 0) { std::printf("runs do not tile the list\n"); ++bad; break; }
                for (int k = 1; k < r.second; ++k) if (t1[r.first + k].x != t1[r.first].x) { std::printf("run straddles bra pair groups\n"); ++bad; break; }
                pos += r.second;
            }
            if (pos != (long long)t1.size()) { std::printf("runs do not cover the list\n"); ++bad; }
            // work items of N ranks partition the list
            for (int nranks : {1, 2, 3, 8}) {
                std::vector<int> owner(t1.size(), 0);
                long long total = 0;
                for (int rank = 0; rank < nranks; ++rank) {
                    std::vector<WorkItem> items;
                    long long mine = 0;
                    make_items(r1, (long long)t1.size(), 148, rank, nranks, &items, &mine);
                    long long slot = 0;
                    for (const WorkItem& it : items) {
                        if (it.y < 1 || it.y > TILES_PER_ITEM_MAX || it.z != slot) { std::printf("bad work item\n"); ++bad; break; }
                        for (int k = 0; k < it.y; ++k) {
                            owner[it.x + k]++;
                            if (t1[it.x + k].x != t1[it.x].x) { std::printf("item straddles bra pair groups\n"); ++bad; }
                        }
                        slot += it.y;
                    }
                    if (slot != mine) { std::printf("tile count of a rank is off\n"); ++bad; }
                    total += mine;
                }
                if (total != (long long)t1.size() || std::any_of(owner.begin(), owner.end(), [](int c) { return c != 1; })) {
                    std::printf("ranks do not partition the tile list (npg %d, %d ranks)\n", npg, nranks); ++bad;
                }
            }
            // rank-owned bra blocks (the form the energy pass uses): the ranks' lists are disjoint, their union is the full list,
            // and the cost-based deal of the blocks balances the estimated work (primitive pairs x partners' primitive pairs)
            for (int nranks : {2, 3, 8}) {
                std::set<std::pair<int, int>> uni;
                size_t total = 0;
                std::vector<double> load(nranks, 0.0);
                for (int rank = 0; rank < nranks; ++rank) {
                    std::vector<TilePair> tr;
                    std::vector<std::pair<long long, int>> rr;
                    make_tile_list(pgs, *cs.a, *cs.b, itol, &tr, &rr, rank, nranks);
                    total += tr.size();
                    for (const TilePair& t : tr) {
                        uni.insert({t.x, t.y});
                        load[rank] += (double)std::max(1, pgs[t.x].pp_beg[NPTYPE]) * (double)std::max(1, pgs[t.y].pp_beg[NPTYPE]);
                    }
                }
                if (total != t1.size() || uni.size() != t1.size()) { std::printf("rank-owned blocks do not partition the tile list (npg %d, %d ranks)\n", npg, nranks); ++bad; }
                if (npg >= 2500 && cs.a == &all) {
                    double mx = 0.0, sum = 0.0;
                    for (double l : load) { mx = std::max(mx, l); sum += l; }
                    if (mx > 1.10 * sum / nranks) { std::printf("rank loads out of balance: max %.3g mean %.3g (%d ranks)\n", mx, sum / nranks, nranks); ++bad; }
                }
            }
        }
    }
    std::printf("%s\n", bad ? "FAIL" : "OK");
    return bad ? 1 : 0;
}
