// banded.hpp -- one rank of the band-sharded denoising chain, in C++ on top of vkpbrt.hpp.
//
// The reference renders on one device; this has no counterpart there.  One process (or thread) per GPU: the frame is
// cut into horizontal bands of whole block rows of the jittered BMFR grid (BandPlan, band_plan.hpp); blocks are
// independent, so the only traffic between neighbours is
//     A  the accumulate planes' history halo (depth history, accumulated illumination, sample counts): pushed right
//        after k_accumulate, overlaps k_bmfr_block, awaited before the NEXT frame's k_accumulate;
//     B  what k_bmfr_block produced and a neighbour reads THIS frame: one row of the tone-mapped output on each side for
//        TAA's 3x3 stencil and column 0 of the rows around the boundary -- a few KB, so that its flag follows the kernel
//        by one launch latency.  Pushed right after k_bmfr_block, overlaps the TAA of the band's inner rows, awaited
//        before its two edge rows;
//     D  what they read NEXT frame: the denoised-history halo (megabytes).  Behind B on the communication stream,
//        awaited before the next frame's k_bmfr_block.  (With both in one push the edge rows waited ~35 us after
//        k_bmfr_block for 3 MB they did not need -- measured at N = 8, 4K.)
//     C  the TAA history halo: pushed after TAA, overlaps the next frame's k_accumulate + k_bmfr_block, awaited before
//        its TAA.
// No rendezvous: that a push may overwrite the receiver's rows follows from what the sender has already waited for
// (a sender at frame f has seen its neighbours' frame f-1 pushes, which they issued after their last read of the
// buffers it writes), with one exception -- B overwrites the stencil row the receiver's TAA of the PREVIOUS frame
// reads, and B is issued before this rank's own TAA -- which is closed by gating B's copies, on the communication
// stream, on the arrival of the receiver's previous C (issued after that TAA).
// Every exchange point is ONE k_halo_push launch on a communication stream (rows stored straight into the receivers'
// HBM over NVLink peer mappings, ordered by flag words, include/vkpbrt_b200.h "Band-sharded multi-GPU runs") and one
// k_halo_wait in front of the consumer; the descriptors are built once per (group, jitter phase, ping-pong parity).
//
// This is the native host of the multi-GPU path: a frame costs one call (run_frame) and ~15 CUDA API calls -- the
// Python twin (vulkanpbrt_b200/multigpu.py BandedPipeline, kept for the NCCL / gloo transports of the tests) spends
// more host time per frame than an 8-way sharded 4K frame takes on the GPUs.
//
// The transport of the 64-byte IPC handles between the ranks is the caller's: AllGather is a collective every rank
// calls in the same order (torch.distributed, MPI, a socket; threads of one process in the emulator tests).
#pragma once

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "band_plan.hpp"
#include "vkpbrt.hpp"

namespace vkpbrt {

// every rank contributes the same number of handles and receives everybody's: result[r] = rank r's list
using AllGather = std::function<std::vector<std::vector<PeerHandle>>(const std::vector<PeerHandle>&)>;

class BandedRank : public Inherit<BandedRank> {
public:
    struct Options {
        bool use_taa = true;
        int max_disp_rows = 24;        // reprojection displacement the halo covers (rows); taps beyond it are counted (check())
        bool external_inputs = false;  // the producer's planes are bound per frame with bind_inputs() (virtual full-frame bases)
        void* comm_stream = nullptr;   // cudaStream_t of the exchange kernels; nullptr: the context's stream (emulator)
        uint32_t timeout_ms = 20000;   // a flag wait that lasts longer sets the error word instead of hanging the GPU
    };

    BandedRank(ref_ptr<Context> ctx, int rank, int world, int width, int height, const Options& opt, AllGather all_gather)
        : context(ctx), plan(width, height, world, 32, opt.max_disp_rows, opt.use_taa), rank_(rank), world_(world), W(width), H(height), opt_(opt),
          all_gather_(std::move(all_gather))
    {
        make_current(*context);
        Context& c = *context;
        if (opt.external_inputs) {
            const uint32_t fmts[4] = {VKPBRT_FORMAT_R32_SFLOAT, VKPBRT_FORMAT_R32G32_SFLOAT, VKPBRT_FORMAT_R8G8B8A8_UNORM, VKPBRT_FORMAT_R32G32B32A32_SFLOAT};
            for (int i = 0; i < 4; ++i) {
                vkpbrt_image_t h;
                check(vkpbrt_image_wrap(c.handle, fmts[i], (uint32_t)W, (uint32_t)H, 1, nullptr, &h));
                ext_[i] = DescriptorImage::create(h, true);
            }
            g_buffer = GBuffer::create(c, ext_[0], ext_[1], ref_ptr<DescriptorImage>(), ext_[2]);
            raw_illumination = IlluminationBufferDemodulatedFloat::create(c, std::vector<ref_ptr<DescriptorImage>>{ext_[3]});
        } else {
            g_buffer = GBuffer::create(c, (uint32_t)W, (uint32_t)H);
            raw_illumination = IlluminationBufferDemodulatedFloat::create(c, (uint32_t)W, (uint32_t)H);
            g_buffer->compile(c);
            raw_illumination->compile(c);
        }
        commands = Commands::create();
        push_constants = PushConstants::create();
        accumulator = Accumulator::create(g_buffer, raw_illumination, /*separate_matrices=*/true);
        accumulator->compile_images(c);
        accumulator->add_dispatch_to_command_graph(commands);
        accumulated = accumulator->accumulated_illumination;
        acc = accumulator->accumulation_buffer;
        bmfr = BMFR::create((uint32_t)W, (uint32_t)H, 32u, 32u, g_buffer, accumulated, acc);
        bmfr->compile(c);
        bmfr->add_dispatch_to_command_graph(commands, push_constants);
        const Rows br = plan.block_rows(rank_);
        bmfr->set_block_row_range(br.lo, br.hi);
        denoiser_final = bmfr->get_final_descriptor_image();
        final_image = denoiser_final;
        if (opt.use_taa) {
            taa = Taa::create((uint32_t)W, (uint32_t)H, 16u, 16u, g_buffer, acc, denoiser_final);
            taa->compile(c);
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
        if (world_ > 1) {
            // a rank holds history rows within max_disp_rows (+1) of the rows it computes: taps beyond that are counted, not
            // silently served from stale rows
            accumulator->set_max_displacement_rows(opt.max_disp_rows);
            // flag words: done[group][src] for the groups A, B, C, D
            flags = DescriptorImage::create(c, (uint32_t)VKPBRT_FORMAT_R32_SFLOAT, (uint32_t)std::max(16, 4 * world), 1u);
            flags->compile(c);
            context->waitForCompletion();
            const auto everyone = all_gather_({peer_export(c, flags->info().data)});
            flag_base_.resize(world);
            for (int r = 0; r < world; ++r) flag_base_[r] = r == rank_ ? static_cast<uint8_t*>(flags->info().data) : map(r, everyone[r][0]);
        }
    }

    // external inputs: this frame's planes as FULL-FRAME base pointers (a band-local buffer holding rows [lo, hi) of
    // input_rows() is passed as  buffer - lo * row_pitch; only rows of input_rows() are ever touched)
    void bind_inputs(void* depth, void* normal, void* albedo, void* illumination)
    {
        void* p[4] = {depth, normal, albedo, illumination};
        for (int i = 0; i < 4; ++i) check(vkpbrt_image_set_data(ext_[i]->handle, p[i]));
    }

    // cam: view, inv_view, proj, inv_proj (4 x 16 floats, column-major) of this frame (VulkanPBRT.cpp:561-563, :578-591)
    void run_frame(int frame, const float* cam)
    {
        make_current(*context);
        auto& pc = push_constants->value();
        CameraMatrices a, b;
        for (int i = 0; i < 16; ++i) pc.view_inverse.m[i] = a.inv_view.m[i] = cam[16 + i];
        a.proj = mat4();
        a.inv_proj = mat4();
        for (int i = 0; i < 16; ++i) { a.proj->m[i] = cam[32 + i]; a.inv_proj->m[i] = pc.proj_inverse.m[i] = cam[48 + i]; }
        pc.frame_number = (uint32_t)frame;
        pc.sample_number = 0;
        b.view = pc.prev_view;
        accumulator->set_camera_matrices(frame, a, b);
        const Rows ar = plan.accumulate_rows(rank_, frame);
        accumulator->set_row_range(ar.lo, ar.hi);
        auto& c = commands->children;       // accumulate, bmfr, [taa], copy_to_back
        const bool multi = world_ > 1;

        finish(pending_a_);
        c[0](*commands);
        if (multi) {
            // A: what k_accumulate just wrote (pre-swap handles) is next frame's history
            pending_a_ = start(0, frame, 0, [&] { return Images{{"acc", {{next_depth, 0, 0}, {accumulated->illumination_images[0], 0, 0}, {acc->spp, 0, 0}}}}; },
                               [&] { return filter(plan.history_transfers(frame + 1), "acc"); });
        }
        finish(pending_d_);              // the neighbours' denoised history rows of the previous frame (bmfrPost.comp:111)
        finish(pending_b_);              // without TAA nothing else waits for B
        pending_b_ = pending_d_ = {};
        c[1](*commands);
        Pending pb;
        if (multi) {
            const uint32_t layer = (uint32_t)((frame & 1) ^ 1);
            // B: what a neighbour reads THIS frame -- the TAA stencil row and column 0 of the tone-mapped image: a few KB,
            // so its flag follows k_bmfr_block by one launch latency.  Gate: the receivers' previous C has arrived here,
            // i.e. their previous TAA no longer reads the stencil row.
            pb = start(1, frame, taa ? seq_[2] : 0,
                       [&] {
                           Images im = {{"final_col0", {{denoiser_final, 0, 4}}}};       // 1 BGRA8 texel
                           if (taa) im["final"] = {{denoiser_final, 0, 0}};
                           return im;
                       },
                       [&] {
                           auto t = filter(plan.stale_column_transfers(frame), "final_col0");
                           if (taa)
                               for (const auto& x : plan.final_transfers(frame)) t.push_back(x);
                           return t;
                       });
            // D: what they read NEXT frame -- the denoised history rows (megabytes): behind B on the communication stream,
            // awaited before the next k_bmfr_block.  No gate: this rank's k_bmfr_block ran after the receivers' previous D
            // arrived, i.e. after their previous k_bmfr_block, the last reader of the layer D overwrites.
            pending_d_ = start(3, frame, 0,
                               [&] {
                                   return Images{{"denoised", {{denoised, layer, 0}}}, {"denoised_col0", {{denoised, layer, 8}}}};     // 1 rgba16f texel
                               },
                               [&] {
                                   auto t = filter(plan.history_transfers(frame + 1), "denoised");
                                   for (const auto& x : filter(plan.stale_column_transfers(frame), "denoised_col0")) t.push_back(x);
                                   return t;
                               });
        }
        if (taa) {
            const Rows o = plan.owned_rows(rank_, frame);
            taa->set_row_range(o.lo, o.hi);
            finish(pending_c_);           // the neighbours' TAA history rows of the previous frame
            pending_c_ = {};
            if (multi && o.hi - o.lo > 2) {
                // the band's first / last row need one row of the neighbour's tone-mapped output (taa.comp:66-83): the rows in
                // between run while that row is in flight, the edge rows after it has landed
                const int i0 = o.lo + (rank_ > 0 ? 1 : 0), i1 = o.hi - (rank_ < world_ - 1 ? 1 : 0);
                taa->record_part(*push_constants, i0, i1, false);
                finish(pb);
                taa->record_parts(*push_constants, o.lo, i0, i1, o.hi, true);   // both edge rows, one launch (either may be empty): hands final -> history
            } else {
                finish(pb);
                c[2](*commands);
            }
        } else {
            pending_b_ = pb;              // awaited before the next frame's k_bmfr_block
        }
        c.back()(*commands);
        for (int i = 0; i < 16; ++i) pc.prev_view.m[i] = cam[i];
        ++swaps_;
        if (multi && taa) {
            pending_c_ = start(2, frame, 0, [&] { return Images{{"taa", {{taa_history, 0, 0}}}}; },
                               [&] { return filter(plan.history_transfers(frame + 1), "taa"); });
        }
    }

    // stream-side wait for the halos in flight: call before reading planes outside the owned rows
    void flush()
    {
        finish(pending_a_);
        finish(pending_b_);
        finish(pending_c_);
        finish(pending_d_);
        pending_a_ = pending_b_ = pending_c_ = pending_d_ = {};
    }

    // synchronises; throws if a flag wait timed out (a peer died or fell out of step) or a reprojection left the rows this rank holds
    void check_errors()
    {
        context->waitForCompletion();
        for (auto& kv : cache_) {
            uint64_t g = 0, w = 0;
            uint32_t e = 0;
            check(vkpbrt_halo_exchange_stats(kv.second.exchange->handle, &g, &w, &e));
            if (e) throw std::runtime_error("halo exchange: a flag wait timed out (peer rank lost or out of step)");
        }
        if (world_ > 1) {
            const uint32_t n = accumulator->displacement_violations();
            if (n)
                throw std::runtime_error("band-sharded run: " + std::to_string(n) + " reprojection taps moved more than max_disp_rows = " +
                                         std::to_string(opt_.max_disp_rows) + " rows; the halo does not cover this camera motion");
        }
    }

    // nanoseconds the streams spent spinning on flag words so far, per group (A, B, C, D): [gate of the push, wait before the consumer]
    std::array<std::array<uint64_t, 2>, 4> spin_ns()
    {
        context->waitForCompletion();
        std::array<std::array<uint64_t, 2>, 4> out{};
        for (auto& kv : cache_) {
            uint64_t g = 0, w = 0;
            uint32_t e = 0;
            check(vkpbrt_halo_exchange_stats(kv.second.exchange->handle, &g, &w, &e));
            out[std::get<0>(kv.first)][0] += g;
            out[std::get<0>(kv.first)][1] += w;
        }
        return out;
    }

    Rows owned_rows(int frame) const { return plan.owned_rows(rank_, frame); }
    Rows input_rows() const { return plan.input_rows(rank_); }
    uint64_t bytes_exchanged() const { return bytes_; }

    ref_ptr<Context> context;
    BandPlan plan;
    ref_ptr<GBuffer> g_buffer;
    ref_ptr<IlluminationBuffer> raw_illumination, accumulated;
    ref_ptr<DescriptorImage> final_image, denoiser_final, denoised;
    ref_ptr<Accumulator> accumulator;
    ref_ptr<BMFR> bmfr;
    ref_ptr<Taa> taa;

private:
    struct PlaneRef {            // an image (or one layer of it) that takes part in an exchange, optionally only the first bytes of its rows
        ref_ptr<DescriptorImage> image;
        uint32_t layer;
        uint32_t column_bytes;   // 0: whole rows
    };
    using Images = std::map<std::string, std::vector<PlaneRef>>;
    struct Pending {
        ref_ptr<HaloExchange> exchange;
        uint32_t value = 0;
    };
    struct Entry {
        ref_ptr<HaloExchange> exchange;
        bool active = false;
        uint64_t bytes = 0;
    };

    static std::vector<Transfer> filter(const std::vector<Transfer>& in, const char* plane)
    {
        std::vector<Transfer> out;
        for (const auto& t : in)
            if (t.plane == plane) out.push_back(t);
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

    // group: 0 = A, 1 = B, 2 = C, 3 = D.  gate_value: see the header comment (0 = no gate)
    template <class ImagesFn, class TransfersFn>
    Pending start(int group, int frame, uint32_t gate_value, ImagesFn images, TransfersFn transfers)
    {
        const uint32_t value = ++seq_[group];
        trace("start", group, frame, value, gate_value);
        // buffers alternate with the copy_to_back swaps and the frame parity; row ranges with the jitter phase
        const auto key = std::make_tuple(group, frame % 16, swaps_ & 1, frame & 1);
        auto it = cache_.find(key);
        if (it == cache_.end()) it = cache_.emplace(key, build(group, frame, images(), transfers())).first;
        if (!it->second.active) return {};
        bytes_ += it->second.bytes;
        it->second.exchange->start_gated(opt_.comm_stream, nullptr, value, gate_value);
        return {it->second.exchange, value};
    }

    void finish(const Pending& p)
    {
        if (p.exchange) {
            trace("wait", -1, -1, p.value, 0);
            p.exchange->wait(nullptr, p.value);
            trace("waited", -1, -1, p.value, 0);
        }
    }
    void trace(const char* what, int group, int frame, uint32_t value, uint32_t gate) const
    {
        static const bool on = std::getenv("VKPBRT_BANDED_TRACE") != nullptr;
        if (on) std::fprintf(stderr, "[rank %d] %s group %d frame %d value %u gate %u\n", rank_, what, group, frame, value, gate);
    }

    Entry build(int group, int /*frame*/, const Images& images, const std::vector<Transfer>& transfers)
    {
        // every rank exports the images of this exchange point, in the same order, and learns everybody's
        std::vector<PeerHandle> mine;
        std::map<std::string, std::vector<size_t>> index;
        for (const auto& kv : images)
            for (const auto& p : kv.second) {
                index[kv.first].push_back(mine.size());
                mine.push_back(peer_export(*context, p.image->info().data));
            }
        const auto everyone = all_gather_(mine);
        std::vector<vkpbrt_halo_copy> copies;
        std::vector<int> send_to, recv_from;
        uint64_t bytes = 0;
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
                    bytes += (uint64_t)c.row_bytes * c.rows;
                    add_unique(send_to, t.dst);
                } else if (t.dst == rank_) {
                    add_unique(recv_from, t.src);
                }
            }
        }
        std::sort(send_to.begin(), send_to.end());
        std::sort(recv_from.begin(), recv_from.end());
        // C always signals both adjacent ranks, rows to copy or not: B's gate (below) counts on one C word per frame
        if (group == 2) {
            for (int d : {rank_ - 1, rank_ + 1})
                if (d >= 0 && d < world_) { add_unique(send_to, d); add_unique(recv_from, d); }
            std::sort(send_to.begin(), send_to.end());
            std::sort(recv_from.begin(), recv_from.end());
        }
        // B's copies are gated on the C words of the ranks whose TAA stencil row it overwrites: the adjacent ranks (local
        // words, set by THEIR C pushes into this rank)
        std::vector<uint32_t*> announce, done;
        std::vector<const uint32_t*> ready, wait;
        for (int s : recv_from) wait.push_back(done_word(rank_, group, s));
        std::vector<int> gate_on;
        if (group == 1 && opt_.use_taa)
            for (const auto& t : transfers)
                if (t.plane == "final" && t.src == rank_ && (t.dst == rank_ - 1 || t.dst == rank_ + 1)) add_unique(gate_on, t.dst);
        for (int d : gate_on) ready.push_back(done_word(rank_, 2, d));
        for (int d : send_to) done.push_back(done_word(d, group, rank_));
        Entry e;
        e.active = !send_to.empty() || !recv_from.empty();
        e.bytes = bytes;
        e.exchange = HaloExchange::create(*context, copies, announce, ready, done, wait, opt_.timeout_ms);
        return e;
    }

    const int rank_, world_, W, H;
    const Options opt_;
    AllGather all_gather_;
    ref_ptr<Commands> commands;
    ref_ptr<PushConstants> push_constants;
    ref_ptr<AccumulationBuffer> acc;
    ref_ptr<DescriptorImage> ext_[4], taa_history, next_depth, flags;
    std::vector<uint8_t*> flag_base_;
    std::map<std::string, ref_ptr<PeerMemory>> mapped_;
    std::map<std::tuple<int, int, int, int>, Entry> cache_;
    uint32_t seq_[4] = {0, 0, 0, 0};
    int swaps_ = 0;
    uint64_t bytes_ = 0;
    Pending pending_a_, pending_b_, pending_c_, pending_d_;
};

}  // namespace vkpbrt
