// cpp_offline_sequence.cpp -- the reference's OFFLINE mode (source/VulkanPBRT.cpp:340-390 import, :566-573 per frame) written
// against include/vkpbrt/io.hpp + vkpbrt.hpp: a pre-rendered sequence (EXR planes + the matrix JSON the reference writes)
// is imported with MatrixIO / GBufferIO / IlluminationBufferIO, staged frame by frame and denoised.
//
//   cpp_offline_sequence <dir> <frames> <position|depth> [export-dir]
//   <dir>/matrices.json, <dir>/{pos|depth}_%d.exr, normal_%d.exr, albedo_%d.exr, illu_%d.exr
//     -> <dir>/final_%d.bgra and the imported G-buffer as <dir>/gbuffer_%d.{depth,normal,albedo}
//   with an export directory, the reference's export flags too (VulkanPBRT.cpp:595-630): every frame's G-buffer and raw
//   illumination are read back after the frame and written as <export-dir>/{pos,depth,normal,albedo,illu}_%d.exr + matrices.json
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "vkpbrt/io.hpp"

using namespace vkpbrt;

static void dump(const std::string& path, const ref_ptr<DescriptorImage>& image, Context& context)
{
    const vkpbrt_image_info i = image->info();
    std::vector<char> host(i.size_bytes);
    check(vkpbrt_image_download(image->handle, host.data(), host.size()));
    context.waitForCompletion();
    std::ofstream(path, std::ios::binary).write(host.data(), (std::streamsize)host.size());
}

int main(int argc, char** argv)
{
    if (argc < 4) { std::cerr << "usage: cpp_offline_sequence dir frames position|depth\n"; return 2; }
    const std::string dir = argv[1];
    const int num_frames = atoi(argv[2]);
    const bool from_position = std::string(argv[3]) == "position";
    const std::string export_dir = argc > 4 ? argv[4] : "";
    try {
        // VulkanPBRT.cpp:340-372: matrices first (positions need them), then the G-buffer and illumination sequences
        const std::vector<CameraMatrices> camera_matrices = MatrixIO::import_matrices(dir + "/matrices.json");
        if ((int)camera_matrices.size() < num_frames) throw std::runtime_error("matrices.json holds fewer frames than requested");
        OfflineGBuffers offline_g_buffers =
            from_position ? GBufferIO::import_g_buffer_position(dir + "/pos_%d.exr", dir + "/normal_%d.exr", "", dir + "/albedo_%d.exr", camera_matrices, num_frames, 0)
                          : GBufferIO::import_g_buffer_depth(dir + "/depth_%d.exr", dir + "/normal_%d.exr", "", dir + "/albedo_%d.exr", num_frames, 0);
        OfflineIlluminations offline_illuminations = IlluminationBufferIO::import_illumination(dir + "/illu_%d.exr", num_frames, 0);
        if (!offline_g_buffers.at(0)->valid()) throw std::runtime_error("frame 0 did not load");
        const uint32_t width = offline_g_buffers[0]->width, height = offline_g_buffers[0]->height;
        const bool separate_matrices = camera_matrices[0].proj.has_value();

        Context context(0);
        make_current(context);
        auto g_buffer = GBuffer::create(context, width, height);
        ref_ptr<IlluminationBuffer> illumination_buffer = IlluminationBufferDemodulatedFloat::create(context, width, height);
        g_buffer->compile(context);
        illumination_buffer->compile(context);
        auto commands = Commands::create();
        auto ray_tracing_push_constants = PushConstants::create();
        auto accumulator = Accumulator::create(g_buffer, illumination_buffer, separate_matrices);
        accumulator->compile_images(context);
        accumulator->add_dispatch_to_command_graph(commands);
        auto raw_illumination = illumination_buffer;
        illumination_buffer = accumulator->accumulated_illumination;
        auto accumulation_buffer = accumulator->accumulation_buffer;
        ref_ptr<DescriptorImage> final_descriptor_image;
        add_denoiser_to_commands(DenoisingType::BMFR, DenoisingBlockSize::X32, commands, context, width, height, ray_tracing_push_constants, g_buffer,
                                 illumination_buffer, accumulation_buffer, final_descriptor_image);
        auto taa = Taa::create(width, height, 16, 16, g_buffer, accumulation_buffer, final_descriptor_image);
        taa->compile(context);
        taa->add_dispatch_to_command_graph(commands);
        final_descriptor_image = taa->get_final_descriptor_image();
        accumulation_buffer->copy_to_back_images(commands, g_buffer, illumination_buffer);

        OfflineGBuffers exported_g_buffers(num_frames);
        OfflineIlluminations exported_illuminations(num_frames);
        for (int frame_index = 0; frame_index < num_frames; ++frame_index) {
            // :566-573: stage the frame, hand the accumulator this frame's and the previous frame's matrices
            offline_g_buffers.at(frame_index)->upload_to_g_buffer(g_buffer, context);
            offline_illuminations.at(frame_index)->upload_to_illumination_buffer(raw_illumination, context);
            const CameraMatrices& cur = camera_matrices[frame_index];
            const CameraMatrices& prev = camera_matrices[frame_index ? frame_index - 1 : frame_index];
            auto& pc = ray_tracing_push_constants->value();
            pc.view_inverse = cur.inv_view;
            if (cur.inv_proj) pc.proj_inverse = *cur.inv_proj;
            pc.frame_number = (uint32_t)frame_index;
            pc.sample_number = 0;
            accumulator->set_camera_matrices(frame_index, cur, prev);
            const std::string n = std::to_string(frame_index);
            dump(dir + "/gbuffer_" + n + ".depth", g_buffer->depth, context);
            dump(dir + "/gbuffer_" + n + ".normal", g_buffer->normal, context);
            dump(dir + "/gbuffer_" + n + ".albedo", g_buffer->albedo, context);
            commands->record();
            pc.prev_view = cur.view;
            dump(dir + "/final_" + n + ".bgra", final_descriptor_image, context);
            if (!export_dir.empty()) {
                // :595-607: wait for the frame, then copy the staged planes into this frame's offline buffers
                exported_g_buffers[frame_index] = OfflineGBuffer::create();
                exported_g_buffers[frame_index]->download_from_g_buffer(g_buffer, context);
                exported_illuminations[frame_index] = OfflineIllumination::create();
                exported_illuminations[frame_index]->download_from_illumination_buffer(raw_illumination, context);
            }
        }
        if (!export_dir.empty()) {
            // :620-634: exporting all images
            const bool fine = GBufferIO::export_g_buffer(separate_matrices ? export_dir + "/pos_%d.exr" : "", export_dir + "/depth_%d.exr", export_dir + "/normal_%d.exr", "",
                                                         export_dir + "/albedo_%d.exr", num_frames, exported_g_buffers, camera_matrices, 0) &&
                              IlluminationBufferIO::export_illumination(export_dir + "/illu_%d.exr", num_frames, exported_illuminations, 0) &&
                              MatrixIO::export_matrices(export_dir + "/matrices.json", camera_matrices);
            if (!fine) throw std::runtime_error("export failed");
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
