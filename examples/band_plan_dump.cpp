// Prints the transfer lists of vkpbrt::BandPlan (include/vkpbrt/band_plan.hpp) for tests/test_cpp_layer.py, which compares
// them with the Python plan (vulkanpbrt_b200/multigpu.py).  usage: band_plan_dump W H world taa frames
#include <cstdio>
#include <cstdlib>

#include <vkpbrt/band_plan.hpp>

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]), world = atoi(argv[3]), taa = atoi(argv[4]), frames = atoi(argv[5]);
    try {
        vkpbrt::BandPlan plan(W, H, world, 32, 24, taa != 0);
        for (int g = 0; g <= world; ++g) printf("brow %d\n", plan.brow[g]);
        for (int g = 0; g < world; ++g) {
            const vkpbrt::Rows in = plan.input_rows(g);
            printf("input %d %d %d\n", g, in.lo, in.hi);
        }
        for (int f = 0; f < frames; ++f) {
            for (int g = 0; g < world; ++g) {
                const vkpbrt::Rows o = plan.owned_rows(g, f), a = plan.accumulate_rows(g, f);
                printf("rows %d %d %d %d %d %d\n", f, g, o.lo, o.hi, a.lo, a.hi);
            }
            for (const auto& t : plan.history_transfers(f + 1)) printf("history %d %d %d %s %d %d\n", f + 1, t.src, t.dst, t.plane.c_str(), t.rows.lo, t.rows.hi);
            for (const auto& t : plan.stale_column_transfers(f)) printf("stale %d %d %d %s %d %d\n", f, t.src, t.dst, t.plane.c_str(), t.rows.lo, t.rows.hi);
            for (const auto& t : plan.final_transfers(f)) printf("final %d %d %d %s %d %d\n", f, t.src, t.dst, t.plane.c_str(), t.rows.lo, t.rows.hi);
        }
    } catch (const std::exception& e) {
        printf("error %s\n", e.what());
    }
    return 0;
}
