// cpp_banded_ranks.cpp -- the band-sharded chain written against the C++ layer: vkpbrt::BandPlan (band_plan.hpp) decides
// which rows travel, vkpbrt::PeerMemory / vkpbrt::HaloExchange (vkpbrt.hpp) move them, the modules' set_*_range restrict
// each rank to its band.  It is the C++ twin of BandedPipeline in vulkanpbrt_b200/multigpu.py.
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

#include "vkpbrt/band_plan.hpp"
#include "vkpbrt/vkpbrt.hpp"

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

// ---- one rank ----------------------------------------------------------------------------------------------------
struct PlaneRef {            // an image (or one layer of it) that takes part in an exchange, optionally only its first bytes of a row
    ref_ptr<DescriptorImage> image;
    uint32_t layer = 0;
    uint32_t column_bytes = 0;   // 0: whole rows
};

class Rank {
public:
    Rank(int rank, int world, int width, int height, bool use_taa, Group& group)
        : rank_(rank), world_(world), W(width), H(height), taa_on(use_taa), group_(group), plan(width, height, world, 32, 12, use_taa)
    {
        context = Context::create(0);
        make_current(*context);
        g_buffer = GBuffer::create(*context, W, H);
        raw_illumination = IlluminationBufferDemodulatedFloat::create(*context, W, H);
        g_buffer->compile(*context);
        raw_illumination->compile(*context);
        commands = Commands::create();
        push_constants = PushConstants::create();
        accumulator = Accumulator::create(g_buffer, raw_illumination, true);
        accumulator->compile_images(*context);
        accumulator->add_dispatch_to_command_graph(commands);
        accumulated = accumulator->accumulated_illumination;
        acc = accumulator->accumulation_buffer;
        bmfr = BMFR::create(W, H, 32, 32, g_buffer, accumulated, acc);
        bmfr->compile(*context);
        bmfr->add_dispatch_to_command_graph(commands, push_constants);
        const Rows br = plan.block_rows(rank_);
        bmfr->set_block_row_range(br.lo, br.hi);
        denoiser_final = bmfr->get_final_descriptor_image();
        final_image = denoiser_final;
        if (use_taa) {
            taa = Taa::create(W, H, 16, 16, g_buffer, acc, denoiser_final);
            taa->compile(*context);
            taa->add_dispatch_to_command_graph(commands);
            final_image = taa->get_final_descriptor_image();
            vkpbrt_image_t h;
            check(vkpbrt_taa_history_image(taa->handle, &h));
            taa_history = DescriptorImage::create(h, false);
        }
        acc->copy_to_back_images(commands, g_buffer, accumulated);
        vkpbrt_image_t img;
        check(vkpbrt_accumulation_buffer_image(acc->handle, VKPBRT_ACC_NEXT_DEPTH, &img));
        next_depth = DescriptorImage::create(img, false);
        check(vkpbrt_bmfr_image_get(bmfr->handle, VKPBRT_BMFR_IMAGE_DENOISED, &img));
        denoised = DescriptorImage::create(img, false);

        // flag words: done[group][src] for the groups A, B, F, then ready[dst]
        flags = DescriptorImage::create(*context, (uint32_t)VKPBRT_FORMAT_R32_SFLOAT, (uint32_t)std::max(16, 4 * world), 1u);
        flags->compile(*context);
        context->waitForCompletion();
        const auto everyone = group_.all_gather(rank_, {peer_export(*context, flags->info().data)});
        flag_base_.resize(world);
        for (int r = 0; r < world; ++r) flag_base_[r] = r == rank_ ? static_cast<uint8_t*>(flags->info().data) : map(r, everyone[r][0]);
    }

    void run_frame(int frame, const float* cam /* view, inv_view, proj, inv_proj */)
    {
        auto& pc = push_constants->value();
        CameraMatrices a, b;
        for (int i = 0; i < 16; ++i) pc.view_inverse.m[i] = a.inv_view.m[i] = cam[16 + i];
        a.proj = mat4();
        a.inv_proj = mat4();
        for (int i = 0; i < 16; ++i) { a.proj->m[i] = cam[32 + i]; a.inv_proj->m[i] = pc.proj_inverse.m[i] = cam[48 + i]; }
        pc.frame_number = frame;
        b.view = pc.prev_view;
        accumulator->set_camera_matrices(frame, a, b);
        const Rows ar = plan.accumulate_rows(rank_, frame);
        accumulator->set_row_range(ar.lo, ar.hi);
        auto& c = commands->children;       // accumulate, bmfr, [taa], copy_to_back

        finish(pending_a_);
        c[0](*commands);
        // A: what k_accumulate just wrote (pre-swap handles) is next frame's history; overlaps k_bmfr_block
        pending_a_ = start("A", frame, {{"acc", {{next_depth, 0, 0}, {accumulated->illumination_images[0], 0, 0}, {acc->spp, 0, 0}}}},
                           filter(plan.history_transfers(frame + 1), true));
        finish(pending_b_);
        pending_b_ = {};
        c[1](*commands);
        if (taa) {
            // F: one row of tone-mapped output on each side for TAA's neighbourhood
            finish(start("F", frame, {{"final", {{denoiser_final, 0, 0}}}}, plan.final_transfers(frame)));
            const Rows o = plan.owned_rows(rank_, frame);
            taa->set_row_range(o.lo, o.hi);
            c[2](*commands);
        }
        c.back()(*commands);
        for (int i = 0; i < 16; ++i) pc.prev_view.m[i] = cam[i];
        ++swaps_;
        // B: denoised / TAA history halos and the stale-column strip; overlaps the next frame's k_accumulate
        const uint32_t layer = (uint32_t)((frame & 1) ^ 1);
        std::map<std::string, std::vector<PlaneRef>> images = {{"denoised", {{denoised, layer, 0}}},
                                                               {"final_col0", {{denoiser_final, 0, 4}}},       // 1 BGRA8 texel
                                                               {"denoised_col0", {{denoised, layer, 8}}}};     // 1 rgba16f texel
        if (taa) images["taa"] = {{taa_history, 0, 0}};
        auto transfers = filter(plan.history_transfers(frame + 1), false);
        for (const auto& t : plan.stale_column_transfers(frame)) transfers.push_back(t);
        pending_b_ = start("B", frame, images, transfers);
    }

    void flush()
    {
        finish(pending_a_);
        finish(pending_b_);
        pending_a_ = pending_b_ = {};
    }

    Rows owned_rows(int frame) const { return plan.owned_rows(rank_, frame); }

    ref_ptr<Context> context;
    ref_ptr<GBuffer> g_buffer;
    ref_ptr<IlluminationBuffer> raw_illumination, accumulated;
    ref_ptr<DescriptorImage> final_image;

private:
    struct Pending {
        ref_ptr<HaloExchange> exchange;
        uint32_t value = 0;
    };
    struct Entry {
        ref_ptr<HaloExchange> exchange;
        bool active = false;
    };

    static std::vector<Transfer> filter(const std::vector<Transfer>& in, bool acc_planes)
    {
        std::vector<Transfer> out;
        for (const auto& t : in)
            if ((t.plane == "acc") == acc_planes) out.push_back(t);
        return out;
    }

    // an allocation is opened once; pointers into it differ by the offset their handle carries
    uint8_t* map(int rank, const PeerHandle& h)
    {
        const std::string key = std::to_string(rank) + ":" + std::string(reinterpret_cast<const char*>(h.bytes), sizeof(h.bytes));
        auto it = mapped_.find(key);
        if (it == mapped_.end()) it = mapped_.emplace(key, PeerMemory::create(context, h)).first;
        return static_cast<uint8_t*>(it->second->base()) + h.offset;
    }

    uint32_t* done_word(int owner, int group, int src) { return reinterpret_cast<uint32_t*>(flag_base_[owner] + 4 * (group * world_ + src)); }
    uint32_t* ready_word(int owner, int dst) { return reinterpret_cast<uint32_t*>(flag_base_[owner] + 4 * (3 * world_ + dst)); }

    Pending start(const std::string& kind, int frame, const std::map<std::string, std::vector<PlaneRef>>& images,
                  const std::vector<Transfer>& transfers)
    {
        const int group = kind == "A" ? 0 : (kind == "B" ? 1 : 2);
        const uint32_t value = ++seq_[group];
        if (world_ == 1) return {};
        // buffers alternate with the copy_to_back swaps and the frame parity; row ranges with the jitter phase
        const auto key = std::make_tuple(group, frame % 16, swaps_ & 1, frame & 1);
        auto it = cache_.find(key);
        if (it == cache_.end()) it = cache_.emplace(key, build(kind, group, images, transfers)).first;
        if (!it->second.active) return {};
        it->second.exchange->start(nullptr, nullptr, value);      // emulator: one stream; on GPUs pass a communication stream
        return {it->second.exchange, value};
    }

    void finish(const Pending& p)
    {
        if (p.exchange) p.exchange->wait(nullptr, p.value);
    }

    Entry build(const std::string& kind, int group, const std::map<std::string, std::vector<PlaneRef>>& images, const std::vector<Transfer>& transfers)
    {
        // every rank exports the images of this exchange point, in the same order, and learns everybody's
        std::vector<PeerHandle> mine;
        std::map<std::string, std::vector<size_t>> index;
        for (const auto& kv : images)
            for (const auto& p : kv.second) {
                index[kv.first].push_back(mine.size());
                mine.push_back(peer_export(*context, p.image->info().data));
            }
        const auto everyone = group_.all_gather(rank_, mine);
        std::vector<vkpbrt_halo_copy> copies;
        std::vector<int> send_to, recv_from;
        auto add_unique = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
        for (const auto& t : transfers) {
            if (t.src == t.dst) continue;
            const auto& planes = images.at(t.plane);
            for (size_t k = 0; k < planes.size(); ++k) {
                const auto info = planes[k].image->info();
                const uint64_t offset = planes[k].layer * info.layer_pitch + (uint64_t)t.rows.lo * info.row_pitch;
                if (t.src == rank_) {
                    vkpbrt_halo_copy c{};
                    c.src = static_cast<uint8_t*>(info.data) + offset;
                    c.dst = map(t.dst, everyone[t.dst][index.at(t.plane)[k]]) + offset;
                    c.src_pitch = c.dst_pitch = info.row_pitch;
                    c.row_bytes = planes[k].column_bytes ? planes[k].column_bytes : (uint32_t)info.row_pitch;
                    c.rows = (uint32_t)(t.rows.hi - t.rows.lo);
                    copies.push_back(c);
                    add_unique(send_to, t.dst);
                } else if (t.dst == rank_) {
                    add_unique(recv_from, t.src);
                }
            }
        }
        std::sort(send_to.begin(), send_to.end());
        std::sort(recv_from.begin(), recv_from.end());
        // the end-of-frame group waits for its receivers' frame to be over (they announce it), as posting a receive would
        const bool handshake = kind == "B";
        std::vector<uint32_t*> announce, done;
        std::vector<const uint32_t*> ready, wait;
        for (int s : recv_from) {
            if (handshake) announce.push_back(ready_word(s, rank_));
            wait.push_back(done_word(rank_, group, s));
        }
        for (int d : send_to) {
            if (handshake) ready.push_back(ready_word(rank_, d));
            done.push_back(done_word(d, group, rank_));
        }
        Entry e;
        e.active = !send_to.empty() || !recv_from.empty();
        e.exchange = HaloExchange::create(*context, copies, announce, ready, done, wait, 120000u);
        return e;
    }

    const int rank_, world_, W, H;
    const bool taa_on;
    Group& group_;
    BandPlan plan;
    ref_ptr<Commands> commands;
    ref_ptr<PushConstants> push_constants;
    ref_ptr<Accumulator> accumulator;
    ref_ptr<AccumulationBuffer> acc;
    ref_ptr<BMFR> bmfr;
    ref_ptr<Taa> taa;
    ref_ptr<DescriptorImage> denoiser_final, taa_history, next_depth, denoised, flags;
    std::vector<uint8_t*> flag_base_;
    std::map<std::string, ref_ptr<PeerMemory>> mapped_;
    std::map<std::tuple<int, int, int, int>, Entry> cache_;
    uint32_t seq_[3] = {0, 0, 0};
    int swaps_ = 0;
    Pending pending_a_, pending_b_;
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
    std::vector<std::unique_ptr<Rank>> ranks(world);      // destroyed after every thread has finished writing into its peers
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r)
        threads.emplace_back([&, r] {
            try {
                ranks[r] = std::make_unique<Rank>(r, world, W, H, use_taa, group);
                Rank& rank = *ranks[r];
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
