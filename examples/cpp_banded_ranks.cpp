// cpp_banded_ranks.cpp -- the band-sharded chain driven through vkpbrt::BandedRank (include/vkpbrt/banded.hpp): BandPlan
// decides which rows travel, PeerMemory / HaloExchange move them, the modules' set_*_range restrict each rank to its band.
// BandedRank is the native host of the multi-GPU path (bench.py reaches it through the C ABI, vkpbrt_banded_rank_*).
//
// Here the ranks are THREADS of one process, each with its own Context and module set, so the program runs against
// the test emulator (tests/test_cpp_layer.py) where "peer" memory is the shared address space.  On GPUs the ranks are
// processes, one per device: the only part that changes is all_gather(), which then has to ship the 64-byte handles
// between processes (MPI, a socket, torch.distributed ...).
//
//   cpp_banded_ranks <dir> <width> <height> <frames> <world> <taa 0|1>
//   <dir>/frame_%d.{depth,normal,albedo,illum,cam}  ->  <dir>/final_%d.bgra  (rows of each frame from the rank that owned them)
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#include "vkpbrt/banded.hpp"

using namespace vkpbrt;

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + p);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// ---- what a multi-process host replaces with its own transport ---------------------------------------------------
class Group {
public:
    explicit Group(int world) : world_(world), slots_(world) {}
    // every rank contributes a list of handles and gets everybody's (a collective: all ranks call it in the same order)
    std::vector<std::vector<PeerHandle>> all_gather(int rank, const std::vector<PeerHandle>& mine)
    {
        std::unique_lock<std::mutex> lock(m_);
        slots_[rank] = mine;
        arrive(lock);
        auto out = slots_;
        arrive(lock);
        return out;
    }
    void abort()
    {
        std::lock_guard<std::mutex> lock(m_);
        aborted_ = true;
        cv_.notify_all();
    }

private:
    void arrive(std::unique_lock<std::mutex>& lock)
    {
        const int gen = generation_;
        if (++count_ == world_) {
            count_ = 0;
            ++generation_;
            cv_.notify_all();
        } else {
            cv_.wait(lock, [&] { return generation_ != gen || aborted_; });
        }
        if (aborted_) throw std::runtime_error("another rank failed");
    }
    int world_, count_ = 0, generation_ = 0;
    bool aborted_ = false;
    std::vector<std::vector<PeerHandle>> slots_;
    std::mutex m_;
    std::condition_variable cv_;
};

int main(int argc, char** argv)
{
    if (argc < 7) { std::cerr << "usage: cpp_banded_ranks dir w h frames world taa\n"; return 2; }
    const std::string dir = argv[1];
    const int W = atoi(argv[2]), H = atoi(argv[3]), frames = atoi(argv[4]), world = atoi(argv[5]);
    const bool use_taa = atoi(argv[6]) != 0;
    Group group(world);
    std::vector<std::vector<char>> finals(frames, std::vector<char>((size_t)W * H * 4));
    std::vector<std::string> errors(world);
    std::vector<ref_ptr<BandedRank>> ranks(world);      // destroyed after every thread has finished writing into its peers
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r)
        threads.emplace_back([&, r] {
            try {
                BandedRank::Options opt;
                opt.use_taa = use_taa;
                opt.max_disp_rows = 12;
                opt.timeout_ms = 120000;
                ranks[r] = BandedRank::create(Context::create(0), r, world, W, H, opt,
                                              [&group, r](const std::vector<PeerHandle>& mine) { return group.all_gather(r, mine); });
                BandedRank& rank = *ranks[r];
                std::vector<char> out((size_t)W * H * 4);
                for (int f = 0; f < frames; ++f) {
                    const std::string base = dir + "/frame_" + std::to_string(f);
                    auto depth = slurp(base + ".depth"), normal = slurp(base + ".normal"), albedo = slurp(base + ".albedo"), illum = slurp(base + ".illum");
                    auto cam = slurp(base + ".cam");
                    check(vkpbrt_image_upload(rank.g_buffer->depth->handle, depth.data(), depth.size()));
                    check(vkpbrt_image_upload(rank.g_buffer->normal->handle, normal.data(), normal.size()));
                    check(vkpbrt_image_upload(rank.g_buffer->albedo->handle, albedo.data(), albedo.size()));
                    check(vkpbrt_image_upload(rank.raw_illumination->illumination_images[0]->handle, illum.data(), illum.size()));
                    rank.context->waitForCompletion();
                    rank.run_frame(f, reinterpret_cast<const float*>(cam.data()));
                    check(vkpbrt_image_download(rank.final_image->handle, out.data(), out.size()));
                    rank.context->waitForCompletion();
                    const Rows o = rank.owned_rows(f);
                    memcpy(finals[f].data() + (size_t)o.lo * W * 4, out.data() + (size_t)o.lo * W * 4, (size_t)(o.hi - o.lo) * W * 4);
                }
                rank.flush();
                rank.check_errors();
            } catch (const std::exception& e) {
                errors[r] = e.what();
                group.abort();
            }
        });
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (!errors[r].empty()) { std::cerr << "rank " << r << ": " << errors[r] << std::endl; return 1; }
    for (int f = 0; f < frames; ++f) std::ofstream(dir + "/final_" + std::to_string(f) + ".bgra", std::ios::binary).write(finals[f].data(), finals[f].size());
    return 0;
}
