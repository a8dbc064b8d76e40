// Host-side properties of the helpers shared by the plan kernels and the contraction (kernels.cuh: __host__ __device__ inline): what
// k_tile_slices / k_slice_reduce / k_plan_range / k_plan_finalize rely on.  Compiled with g++ against the CUDA headers (no GPU).
//   slice_width: slices of a tile cover its columns exactly once, are multiples of SLICE_COLS wide (except the last), at most S of them,
//                and their number follows the tile's cost;
//   piece_cost : monotone in the active-set sizes, positive for empty tiles;
//   drain_group: non-decreasing along a batch, in [0, DRAIN_GROUPS), first tile of a batch in group 0.
#include <cstdio>
#include <cstdlib>
#include <random>

#include <cuda_runtime_api.h>

#include "../../gimic_b200/csrc/kernels.cuh"

static int failures = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

int main() {
    std::mt19937_64 rng(7);
    auto uni = [&](long long lo, long long hi) { return (long long)(lo + rng() % (unsigned long long)(hi - lo + 1)); };
    for (int it = 0; it < 20000; ++it) {
        gb::TileDesc td{};
        td.nreal = (int)uni(1, 3000); td.nn = (td.nreal + 7) / 8 * 8;
        td.nraw = td.nreal + (int)uni(0, 40); td.nact = (td.nraw + 7) / 8 * 8;
        td.nruns = (int)uni(1, 90); td.npts = (int)uni(1, 128);
        const int S = (int)uni(1, 16);
        const long long item_cost = uni(1, 40000000);
        const int w = gb::slice_width(td, S, item_cost);
        EXPECT(w > 0 && w % gb::SLICE_COLS == 0);
        const int nsl = (td.nn + w - 1) / w;
        EXPECT(nsl >= 1 && nsl <= S);
        int covered = 0;
        for (int s = 0; s < S; ++s) {       // k_tile_slices
            const int c0 = s * w < td.nn ? s * w : td.nn, c1 = c0 + w < td.nn ? c0 + w : td.nn;
            EXPECT(c0 == covered || c1 == c0);
            covered = c1 > covered ? c1 : covered;
            if (s >= nsl) EXPECT(c1 <= c0);
        }
        EXPECT(covered == td.nn);
        const long long c = gb::piece_cost(td.npts, td.nraw, td.nreal, td.nruns);
        if (td.nn >= 2 * gb::SLICE_COLS && S >= 2 && c > 2 * item_cost) EXPECT(nsl >= 2);        // a tile costing more than two items is cut
        if (c <= item_cost) EXPECT(nsl == 1);
        EXPECT(c > 0 && gb::piece_cost(td.npts, td.nraw + 8, td.nreal, td.nruns) > c && gb::piece_cost(td.npts, td.nraw, td.nreal + 8, td.nruns) > c);
    }
    EXPECT(gb::piece_cost(128, 0, 0, 0) > 0);
    for (int it = 0; it < 2000; ++it) {
        const long long pool = uni(1 << 20, 1LL << 32), chunk = uni(1 << 18, pool);
        long long prefix = 0; int last = 0;
        for (int t = 0; t < 400; ++t) {
            const long long b = prefix / pool;
            const int g = gb::drain_group(prefix, pool, chunk);
            EXPECT(g >= 0 && g < gb::DRAIN_GROUPS);
            static long long last_b = -1;
            if (t == 0 || b != last_b) { last = g; last_b = b; if (prefix % pool < chunk) EXPECT(g == 0); }
            EXPECT(g >= last);
            last = g;
            prefix += uni(1, chunk / 2 + 1);
        }
    }
    std::printf("plan props: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
