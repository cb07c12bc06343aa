// cpp_frame_loop.cpp -- the reference's denoising wiring (source/VulkanPBRT.cpp:337-505 setup, :551-618 frame loop)
// written against include/vkpbrt/vkpbrt.hpp.  Reads a raw synthetic sequence (tests/test_cpp_layer.py writes it),
// replays the recorded command list once per frame and writes the final BGRA8 image of every frame.
//
//   cpp_frame_loop <dir> <width> <height> <frames> <bmfr|bfr> <taa 0|1> [block 8|16|32, or 0 = X8X16X32 + blender]
//   <dir>/frame_%d.{depth,normal,albedo,illum,cam}  ->  <dir>/final_%d.bgra
//
// -DVKPBRT_REFERENCE_WIRING=<file>: instead of this repository's add_denoiser_to_commands, <file> -- the reference's OWN
// source/util/DenoiserUtils.cpp, its #include lines removed (tests/test_cpp_layer.py generates it from the reference tree) --
// is compiled in, behind the three substitutions vkpbrt.hpp documents (vsg::ref_ptr, vsg::Context / CompileTraversal,
// vsg::Commands / PushConstants / DescriptorImage), and called with the reference's signature.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "vkpbrt/vkpbrt.hpp"

#ifdef VKPBRT_REFERENCE_WIRING
namespace vsg {
template <class T> using ref_ptr = vkpbrt::ref_ptr<T>;
using Commands = vkpbrt::Commands;
using PushConstants = vkpbrt::PushConstants;
using DescriptorImage = vkpbrt::DescriptorImage;
struct CompileTraversal { vkpbrt::Context& context; };      // the reference reaches the context through compile.context
}
#define VKPBRT_STRINGIFY_(x) #x
#define VKPBRT_STRINGIFY(x) VKPBRT_STRINGIFY_(x)
#include VKPBRT_STRINGIFY(VKPBRT_REFERENCE_WIRING)
#endif

using namespace vkpbrt;

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + p);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv)
{
    if (argc < 7) { std::cerr << "usage: cpp_frame_loop dir w h frames bmfr|bfr taa\n"; return 2; }
    const std::string dir = argv[1];
    const int width = atoi(argv[2]), height = atoi(argv[3]), num_frames = atoi(argv[4]);
    const DenoisingType denoising_type = std::string(argv[5]) == "bfr" ? DenoisingType::BFR : DenoisingType::BMFR;
    const bool use_taa = atoi(argv[6]) != 0;
    const int block = argc > 7 ? atoi(argv[7]) : 32;
    const DenoisingBlockSize denoising_size = block == 0 ? DenoisingBlockSize::X8X16X32 : block == 8 ? DenoisingBlockSize::X8 : block == 16 ? DenoisingBlockSize::X16 : DenoisingBlockSize::X32;
    try {
        Context context(0);
        make_current(context);
        // VulkanPBRT.cpp:337-339
        auto g_buffer = GBuffer::create(context, width, height);
        ref_ptr<IlluminationBuffer> illumination_buffer = IlluminationBufferDemodulatedFloat::create(context, width, height);
        g_buffer->compile(context);
        illumination_buffer->compile(context);
        auto commands = Commands::create();
        auto ray_tracing_push_constants = PushConstants::create();
        // :426-432
        auto accumulator = Accumulator::create(g_buffer, illumination_buffer, /*separate_matrices=*/true);
        accumulator->compile_images(context);
        accumulator->update_image_layouts(context);
        accumulator->add_dispatch_to_command_graph(commands);
        auto raw_illumination = illumination_buffer;
        illumination_buffer = accumulator->accumulated_illumination;
        auto accumulation_buffer = accumulator->accumulation_buffer;
        // :443
        ref_ptr<DescriptorImage> final_descriptor_image;
#ifdef VKPBRT_REFERENCE_WIRING
        vsg::CompileTraversal compile{context};
        add_denoiser_to_commands(denoising_type, denoising_size, commands, compile, width, height, ray_tracing_push_constants,
                                 g_buffer, illumination_buffer, accumulation_buffer, final_descriptor_image);
#else
        add_denoiser_to_commands(denoising_type, denoising_size, commands, context, width, height, ray_tracing_push_constants,
                                 g_buffer, illumination_buffer, accumulation_buffer, final_descriptor_image);
#endif
        // :448-456 -- a block-local ref_ptr, exactly as in the reference: the command graph keeps the module alive
        if (use_taa) {
            auto taa = Taa::create(width, height, 16, 16, g_buffer, accumulation_buffer, final_descriptor_image);
            taa->compile(context);
            taa->update_image_layouts(context);
            taa->add_dispatch_to_command_graph(commands);
            final_descriptor_image = taa->get_final_descriptor_image();
        }
        // :502-505
        accumulation_buffer->copy_to_back_images(commands, g_buffer, illumination_buffer);

        std::vector<char> out((size_t)width * height * 4);
        for (int frame_index = 0; frame_index < num_frames; ++frame_index) {
            const std::string base = dir + "/frame_" + std::to_string(frame_index);
            auto depth = slurp(base + ".depth"), normal = slurp(base + ".normal"), albedo = slurp(base + ".albedo"), illum = slurp(base + ".illum");
            auto cam = slurp(base + ".cam");   // view, inv_view, proj, inv_proj: 4 x 16 floats
            const float* cm = reinterpret_cast<const float*>(cam.data());
            // :568-569 staging upload
            check(vkpbrt_image_upload(g_buffer->depth->handle, depth.data(), depth.size()));
            check(vkpbrt_image_upload(g_buffer->normal->handle, normal.data(), normal.size()));
            check(vkpbrt_image_upload(g_buffer->albedo->handle, albedo.data(), albedo.size()));
            check(vkpbrt_image_upload(raw_illumination->illumination_images[0]->handle, illum.data(), illum.size()));
            context.waitForCompletion();
            // :561-563, :578-584
            auto& pc = ray_tracing_push_constants->value();
            CameraMatrices a, b;
            for (int i = 0; i < 16; ++i) pc.view_inverse.m[i] = a.inv_view.m[i] = cm[16 + i];
            a.proj = mat4();
            a.inv_proj = mat4();
            for (int i = 0; i < 16; ++i) { a.proj->m[i] = cm[32 + i]; a.inv_proj->m[i] = pc.proj_inverse.m[i] = cm[48 + i]; }
            pc.frame_number = frame_index;
            b.view = pc.prev_view;
            accumulator->set_camera_matrices(frame_index, a, b);
            commands->record();   // :588
            for (int i = 0; i < 16; ++i) pc.prev_view.m[i] = cm[i];   // :591
            check(vkpbrt_image_download(final_descriptor_image->handle, out.data(), out.size()));
            context.waitForCompletion();
            std::ofstream(dir + "/final_" + std::to_string(frame_index) + ".bgra", std::ios::binary).write(out.data(), out.size());
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
