// Band-sharded multi-GPU runs: integer geometry of the partition (C++ twin of vulkanpbrt_b200/multigpu.py BandPlan,
// tests/test_cpp_layer.py checks the two produce identical transfer lists).
//
// The reference records one device's command graph (VulkanPBRT.cpp:551-618); this is new.  The frame is cut into bands
// of whole block rows of BMFR's jittered grid (bmfrGeneral.comp:36, bmfrPre.comp:16): blocks are independent, so a rank
// needs from its neighbours only
//   * history rows within the maximum reprojection displacement (+1 bilinear row, + the opposite image edge row for
//     the samplers' REPEAT addressing) of prev_depth / accumulated illumination / sample counts ("acc"), the denoised
//     history ("denoised") and the TAA history ("taa"), sent by the rank that owned the row when it was written;
//   * one row of tone-mapped output on each side for TAA's 3x3 neighbourhood ("final", taa.comp:66-83);
//   * column 0 of the rows around a boundary ("final_col0", "denoised_col0"): frames whose x jitter is negative leave
//     column 0 unwritten, and the boundary moves with the y jitter.
// Feed the lists to vkpbrt::HaloExchange (vkpbrt.hpp): one exchange point per list, per (frame % 16, ping-pong parity).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace vkpbrt {

struct Rows {       // [lo, hi)
    int lo = 0, hi = 0;
    bool empty() const { return hi <= lo; }
};

struct Transfer {
    int src, dst;
    std::string plane;
    Rows rows;
};

class BandPlan {
public:
    BandPlan(int width, int height, int world, int block = 32, int max_disp_rows = 24, bool taa = false)
        : W(width), H(height), N(world), b(block), D(max_disp_rows), taa_(taa), nby(height / block + 2)
    {
        for (int g = 0; g <= world; ++g) brow.push_back(round_half_even((double)g * nby / world));
        const int edge_block_rows = (max_disp_rows + 1 + block + block - 1) / block;
        for (int g = 0; g < world; ++g)
            if (brow[g + 1] - brow[g] < 2 * edge_block_rows) throw std::invalid_argument("bands must be at least two edge regions high");
    }

    // ivec2(vec2(b, b) * pixelOffsets[frame % 16]) (bmfrGeneral.comp:36), binary32 product truncated toward zero
    static std::pair<int, int> block_offset(int block, int frame)
    {
        static const float offs[16][2] = {{.7f, .85f}, {.95f, .5f}, {.43f, .76f}, {.97f, .03f}, {.37f, .58f}, {.03f, .36f}, {.81f, .46f},
                                          {0.f, .78f}, {.36f, -.08f}, {-.06f, 0.f}, {.95f, .1f}, {.85f, .61f}, {.06f, .1f}, {.43f, .16f},
                                          {0.f, .5f}, {.73f, .38f}};
        const float* o = offs[((frame % 16) + 16) % 16];
        return {(int)((float)block * o[0]), (int)((float)block * o[1])};
    }

    Rows block_rows(int g) const { return {brow[g], brow[g + 1]}; }

    // first image row of rank g's band at `frame` (rows written by its BMFR blocks)
    int boundary(int g, int frame) const
    {
        if (g <= 0) return 0;
        if (g >= N) return H;
        const int oy = block_offset(b, frame).second;
        return std::min(H, std::max(0, b * brow[g] - oy));
    }
    Rows owned_rows(int g, int frame) const { return {boundary(g, frame), boundary(g + 1, frame)}; }

    // image rows the rank's blocks read through the jitter + mirror footprint
    Rows accumulate_rows(int g, int frame) const
    {
        const int oy = block_offset(b, frame).second;
        const int lo_abs = b * brow[g] - oy, hi_abs = b * brow[g + 1] - oy - 1;
        const int r[4] = {mirror(lo_abs, H), mirror(hi_abs, H), mirror(std::max(lo_abs, 0), H), mirror(std::min(hi_abs, H - 1), H)};
        const int lo = *std::min_element(r, r + 4), hi = *std::max_element(r, r + 4) + 1;
        const Rows o = owned_rows(g, frame);
        return clip({std::min(lo, o.lo), std::max(hi, o.hi)});
    }

    // rows of the producer's planes the rank ever touches (all 16 jitter phases)
    Rows input_rows(int g) const
    {
        Rows out = accumulate_rows(g, 0);
        for (int f = 1; f < 16; ++f) {
            const Rows a = accumulate_rows(g, f);
            out.lo = std::min(out.lo, a.lo);
            out.hi = std::max(out.hi, a.hi);
        }
        return out;
    }

    // rows every rank must receive before running `frame_next`, sent by the owner of the row at frame_next - 1
    std::vector<Transfer> history_transfers(int frame_next) const
    {
        const int f0 = frame_next - 1;
        std::vector<Transfer> out;
        for (int dst = 0; dst < N; ++dst) {
            const Rows have_acc = accumulate_rows(dst, f0), have_own = owned_rows(dst, f0);
            std::vector<std::pair<std::string, std::vector<Rows>>> need;
            need.push_back({"acc", with_disp(accumulate_rows(dst, frame_next))});
            need.push_back({"denoised", with_disp(owned_rows(dst, frame_next))});
            if (taa_) need.push_back({"taa", with_disp(owned_rows(dst, frame_next))});
            for (const auto& pn : need) {
                const Rows have = pn.first == "acc" ? have_acc : have_own;
                for (const Rows& r : pn.second)
                    for (int src = 0; src < N; ++src) {
                        if (src == dst) continue;
                        // drop what the receiver computed itself (identical values)
                        for (const Rows& piece : subtract(intersect(r, owned_rows(src, f0)), have))
                            if (!piece.empty()) out.push_back({src, dst, pn.first, piece});
                    }
            }
        }
        return out;
    }

    // after every frame each rank mirrors column 0 of the rows it wrote inside the window the boundary can move in
    std::vector<Transfer> stale_column_transfers(int frame) const
    {
        std::vector<Transfer> out;
        for (int g = 1; g < N; ++g) {
            const Rows win = clip({b * brow[g] - b, b * brow[g] + 3});
            const int pairs[2][2] = {{g - 1, g}, {g, g - 1}};
            for (const auto& p : pairs) {
                const Rows rows = intersect(win, owned_rows(p[0], frame));
                if (!rows.empty()) {
                    out.push_back({p[0], p[1], "final_col0", rows});
                    out.push_back({p[0], p[1], "denoised_col0", rows});
                }
            }
        }
        return out;
    }

    // one row of the denoiser's tone-mapped output on each side of the owned rows, for TAA
    std::vector<Transfer> final_transfers(int frame) const
    {
        std::vector<Transfer> out;
        if (!taa_) return out;
        for (int dst = 0; dst < N; ++dst) {
            const Rows o = owned_rows(dst, frame);
            const int rows[2] = {o.lo - 1, o.hi};
            for (int row : rows) {
                if (row < 0 || row >= H) continue;
                for (int src = 0; src < N; ++src) {
                    const Rows s = owned_rows(src, frame);
                    if (src != dst && s.lo <= row && row < s.hi) out.push_back({src, dst, "final", {row, row + 1}});
                }
            }
        }
        return out;
    }

    const int W, H, N, b, D;
    std::vector<int> brow;      // first block row of every band, and the end

private:
    static int mirror(int x, int s) { return x < 0 ? -x - 1 : (x >= s ? 2 * s - x - 1 : x); }
    static int round_half_even(double x)      // Python's round()
    {
        const double f = std::floor(x), d = x - f;
        if (d > 0.5) return (int)f + 1;
        if (d < 0.5) return (int)f;
        return ((long long)f % 2 == 0) ? (int)f : (int)f + 1;
    }
    Rows clip(Rows r) const
    {
        const int lo = std::max(0, r.lo), hi = std::min(H, r.hi);
        return {lo, std::max(lo, hi)};
    }
    static Rows intersect(Rows a, Rows c)
    {
        const int lo = std::max(a.lo, c.lo), hi = std::min(a.hi, c.hi);
        return {lo, std::max(lo, hi)};
    }
    static std::vector<Rows> subtract(Rows a, Rows c)      // a \ c
    {
        std::vector<Rows> out;
        if (a.empty()) return out;
        const int lo = std::max(a.lo, c.lo), hi = std::min(a.hi, c.hi);
        if (hi <= lo) {
            out.push_back(a);
            return out;
        }
        if (a.lo < lo) out.push_back({a.lo, lo});
        if (hi < a.hi) out.push_back({hi, a.hi});
        return out;
    }
    std::vector<Rows> with_disp(Rows r) const
    {
        const int lo = std::max(0, r.lo - D - 1), hi = std::min(H, r.hi + D + 1);
        std::vector<Rows> out{{lo, hi}};
        // REPEAT addressing: a tap at row -1 / H wraps to the opposite image edge
        if (lo == 0 && hi < H) out.push_back({H - 1, H});
        if (hi == H && lo > 0) out.push_back({0, 1});
        return out;
    }
    const bool taa_;
    const int nby;
};

}  // namespace vkpbrt
