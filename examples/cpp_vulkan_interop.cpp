// cpp_vulkan_interop.cpp -- a raw-Vulkan host sharing its frames with the CUDA denoising path through
// include/vkpbrt/vk_interop.hpp (SURVEY.md section 8(b) "Sync with Vulkan", 8(f)-2).
//
// The Vulkan side plays VulkanPBRT's renderer: it owns TILING_OPTIMAL G-buffer / illumination images (here filled from
// files through a staging buffer instead of by ptRaygen.rgen), copies them into the exported planes of a
// vkpbrt::vk::SharedFrame and signals a timeline semaphore; the CUDA side is the reference's wiring
// (VulkanPBRT.cpp:424-505) with a wait in front and a copy + signal behind; the Vulkan side then copies the denoised
// plane into its "presented" image and reads it back.  No vkQueueWaitIdle orders the two APIs -- only the semaphores.
//
//   cpp_vulkan_interop <dir> <width> <height> <frames> <bmfr|bfr> <taa 0|1>
//   <dir>/frame_%d.{depth,normal,albedo,illum,cam}  ->  <dir>/final_%d.bgra
//   VKPBRT_VULKAN_LIBRARY: the Vulkan loader to dlopen (default libvulkan.so.1; the tests point it at tests/vkmock)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "vkpbrt/vk_interop.hpp"

#if !VKPBRT_HAVE_VULKAN
#error "cpp_vulkan_interop needs Vulkan headers (vulkan/vulkan_core.h on the include path, or -DVKPBRT_VULKAN_HEADER=...)"
#endif

using namespace vkpbrt;

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + p);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// the few raw-Vulkan entry points of the host itself (the interop header loads its own)
struct HostVk {
    PFN_vkGetInstanceProcAddr gipa;
    VkInstance instance = VK_NULL_HANDLE;
    VkPhysicalDevice physical_device = VK_NULL_HANDLE;
    VkDevice device = VK_NULL_HANDLE;
    VkQueue queue = VK_NULL_HANDLE;
    VkPhysicalDeviceMemoryProperties memory_properties{};
    template <typename F> F inst(const char* name) const
    {
        auto f = reinterpret_cast<F>(gipa(instance, name));
        if (!f) throw std::runtime_error(std::string("missing ") + name);
        return f;
    }
    template <typename F> F dev(const char* name) const
    {
        auto f = reinterpret_cast<F>(inst<PFN_vkGetDeviceProcAddr>("vkGetDeviceProcAddr")(device, name));
        if (!f) throw std::runtime_error(std::string("missing ") + name);
        return f;
    }
    uint32_t memory_type(uint32_t bits, VkMemoryPropertyFlags flags) const
    {
        for (uint32_t i = 0; i < memory_properties.memoryTypeCount; ++i)
            if ((bits & (1u << i)) && (memory_properties.memoryTypes[i].propertyFlags & flags) == flags) return i;
        throw std::runtime_error("no suitable memory type");
    }
};

// a TILING_OPTIMAL image of the renderer
struct RenderImage {
    VkImage image = VK_NULL_HANDLE;
    VkDeviceMemory memory = VK_NULL_HANDLE;
};
static RenderImage make_image(const HostVk& vk, VkFormat format, uint32_t w, uint32_t h)
{
    RenderImage r;
    VkImageCreateInfo ci{};
    ci.sType = VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO;
    ci.imageType = VK_IMAGE_TYPE_2D;
    ci.format = format;
    ci.extent = {w, h, 1};
    ci.mipLevels = 1; ci.arrayLayers = 1;
    ci.samples = VK_SAMPLE_COUNT_1_BIT;
    ci.tiling = VK_IMAGE_TILING_OPTIMAL;
    ci.usage = VK_IMAGE_USAGE_STORAGE_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT;
    ci.initialLayout = VK_IMAGE_LAYOUT_UNDEFINED;
    vk::vk_check(vk.dev<PFN_vkCreateImage>("vkCreateImage")(vk.device, &ci, nullptr, &r.image), "vkCreateImage");
    VkMemoryRequirements req;
    vk.dev<PFN_vkGetImageMemoryRequirements>("vkGetImageMemoryRequirements")(vk.device, r.image, &req);
    VkMemoryAllocateInfo ai{};
    ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
    ai.allocationSize = req.size;
    ai.memoryTypeIndex = vk.memory_type(req.memoryTypeBits, VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT);
    vk::vk_check(vk.dev<PFN_vkAllocateMemory>("vkAllocateMemory")(vk.device, &ai, nullptr, &r.memory), "vkAllocateMemory");
    vk::vk_check(vk.dev<PFN_vkBindImageMemory>("vkBindImageMemory")(vk.device, r.image, r.memory, 0), "vkBindImageMemory");
    return r;
}

// a host-visible staging / read-back buffer
struct HostBuffer {
    VkBuffer buffer = VK_NULL_HANDLE;
    VkDeviceMemory memory = VK_NULL_HANDLE;
    char* map = nullptr;
    VkDeviceSize size = 0;
};
static HostBuffer make_host_buffer(const HostVk& vk, VkDeviceSize size)
{
    HostBuffer b;
    b.size = size;
    VkBufferCreateInfo ci{};
    ci.sType = VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO;
    ci.size = size;
    ci.usage = VK_BUFFER_USAGE_TRANSFER_SRC_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT;
    ci.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
    vk::vk_check(vk.dev<PFN_vkCreateBuffer>("vkCreateBuffer")(vk.device, &ci, nullptr, &b.buffer), "vkCreateBuffer");
    VkMemoryRequirements req;
    vk.dev<PFN_vkGetBufferMemoryRequirements>("vkGetBufferMemoryRequirements")(vk.device, b.buffer, &req);
    VkMemoryAllocateInfo ai{};
    ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
    ai.allocationSize = req.size;
    ai.memoryTypeIndex = vk.memory_type(req.memoryTypeBits, VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT | VK_MEMORY_PROPERTY_HOST_COHERENT_BIT);
    vk::vk_check(vk.dev<PFN_vkAllocateMemory>("vkAllocateMemory")(vk.device, &ai, nullptr, &b.memory), "vkAllocateMemory");
    vk::vk_check(vk.dev<PFN_vkBindBufferMemory>("vkBindBufferMemory")(vk.device, b.buffer, b.memory, 0), "vkBindBufferMemory");
    void* p = nullptr;
    vk::vk_check(vk.dev<PFN_vkMapMemory>("vkMapMemory")(vk.device, b.memory, 0, VK_WHOLE_SIZE, 0, &p), "vkMapMemory");
    b.map = static_cast<char*>(p);
    return b;
}

static VkBufferImageCopy whole_image(uint32_t w, uint32_t h)
{
    VkBufferImageCopy r{};
    r.imageSubresource.aspectMask = VK_IMAGE_ASPECT_COLOR_BIT;
    r.imageSubresource.layerCount = 1;
    r.imageExtent = {w, h, 1};
    return r;
}

static void image_barrier(PFN_vkCmdPipelineBarrier barrier, VkCommandBuffer cb, VkImage image, VkImageLayout from, VkImageLayout to)
{
    VkImageMemoryBarrier b{};
    b.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
    b.srcAccessMask = VK_ACCESS_MEMORY_WRITE_BIT;
    b.dstAccessMask = VK_ACCESS_MEMORY_READ_BIT | VK_ACCESS_MEMORY_WRITE_BIT;
    b.oldLayout = from; b.newLayout = to;
    b.srcQueueFamilyIndex = b.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
    b.image = image;
    b.subresourceRange = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, 1};
    barrier(cb, VK_PIPELINE_STAGE_ALL_COMMANDS_BIT, VK_PIPELINE_STAGE_ALL_COMMANDS_BIT, 0, 0, nullptr, 0, nullptr, 1, &b);
}

int main(int argc, char** argv)
{
    if (argc < 7) { std::cerr << "usage: cpp_vulkan_interop dir w h frames bmfr|bfr taa\n"; return 2; }
    const std::string dir = argv[1];
    const uint32_t width = atoi(argv[2]), height = atoi(argv[3]);
    const int num_frames = atoi(argv[4]);
    const DenoisingType denoising_type = std::string(argv[5]) == "bfr" ? DenoisingType::BFR : DenoisingType::BMFR;
    const bool use_taa = atoi(argv[6]) != 0;
    try {
        // ---- Vulkan: instance, the physical device, a device with the interop's extensions (VulkanPBRT.cpp:167-176) ----
        HostVk vk;
        vk.gipa = vk::open_loader(getenv("VKPBRT_VULKAN_LIBRARY"));
        {
            VkApplicationInfo app{};
            app.sType = VK_STRUCTURE_TYPE_APPLICATION_INFO;
            app.pApplicationName = "cpp_vulkan_interop";
            app.apiVersion = VK_API_VERSION_1_2;
            VkInstanceCreateInfo ci{};
            ci.sType = VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO;
            ci.pApplicationInfo = &app;
            auto create_instance = reinterpret_cast<PFN_vkCreateInstance>(vk.gipa(VK_NULL_HANDLE, "vkCreateInstance"));
            if (!create_instance) throw std::runtime_error("the loader has no vkCreateInstance");
            vk::vk_check(create_instance(&ci, nullptr, &vk.instance), "vkCreateInstance");
        }
        {
            uint32_t n = 1;
            const VkResult r = vk.inst<PFN_vkEnumeratePhysicalDevices>("vkEnumeratePhysicalDevices")(vk.instance, &n, &vk.physical_device);
            if ((r != VK_SUCCESS && r != VK_INCOMPLETE) || n < 1) throw std::runtime_error("no Vulkan physical device");
            vk.inst<PFN_vkGetPhysicalDeviceMemoryProperties>("vkGetPhysicalDeviceMemoryProperties")(vk.physical_device, &vk.memory_properties);
        }
        {
            const float priority = 1.f;
            VkDeviceQueueCreateInfo qi{};
            qi.sType = VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO;
            qi.queueFamilyIndex = 0; qi.queueCount = 1; qi.pQueuePriorities = &priority;
            VkPhysicalDeviceVulkan12Features f12{};
            f12.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_VULKAN_1_2_FEATURES;
            f12.timelineSemaphore = VK_TRUE;
            const std::vector<const char*> extensions = vk::required_device_extensions();
            VkDeviceCreateInfo ci{};
            ci.sType = VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO;
            ci.pNext = &f12;
            ci.queueCreateInfoCount = 1; ci.pQueueCreateInfos = &qi;
            ci.enabledExtensionCount = (uint32_t)extensions.size(); ci.ppEnabledExtensionNames = extensions.data();
            vk::vk_check(vk.inst<PFN_vkCreateDevice>("vkCreateDevice")(vk.physical_device, &ci, nullptr, &vk.device), "vkCreateDevice");
            vk.dev<PFN_vkGetDeviceQueue>("vkGetDeviceQueue")(vk.device, 0, 0, &vk.queue);
        }
        const vk::Api api = vk::Api::load(vk.gipa, vk.instance, vk.physical_device, vk.device);

        // ---- CUDA context on the SAME GPU, the shared planes and semaphores ----
        Context context(vk::cuda_device_of(api));
        make_current(context);
        auto frame = vk::SharedFrame::create(api, context, width, height);

        // ---- the reference's wiring (VulkanPBRT.cpp:424-505), between the wait and the signal ----
        auto g_buffer = frame->g_buffer;
        ref_ptr<IlluminationBuffer> illumination_buffer = frame->illumination_buffer;
        auto commands = Commands::create();
        auto ray_tracing_push_constants = PushConstants::create();
        frame->add_wait_to_commands(commands, ray_tracing_push_constants);
        auto accumulator = Accumulator::create(g_buffer, illumination_buffer, /*separate_matrices=*/true);
        accumulator->compile_images(context);
        accumulator->update_image_layouts(context);
        accumulator->add_dispatch_to_command_graph(commands);
        illumination_buffer = accumulator->accumulated_illumination;
        auto accumulation_buffer = accumulator->accumulation_buffer;
        ref_ptr<DescriptorImage> final_descriptor_image;
        add_denoiser_to_commands(denoising_type, DenoisingBlockSize::X32, commands, context, width, height, ray_tracing_push_constants,
                                 g_buffer, illumination_buffer, accumulation_buffer, final_descriptor_image);
        if (use_taa) {
            auto taa = Taa::create(width, height, 16, 16, g_buffer, accumulation_buffer, final_descriptor_image);
            taa->compile(context);
            taa->update_image_layouts(context);
            taa->add_dispatch_to_command_graph(commands);
            final_descriptor_image = taa->get_final_descriptor_image();
        }
        accumulation_buffer->copy_to_back_images(commands, g_buffer, illumination_buffer);
        frame->add_signal_to_commands(commands, ray_tracing_push_constants, final_descriptor_image);

        // ---- the renderer's own resources ----
        RenderImage depth = make_image(vk, VK_FORMAT_R32_SFLOAT, width, height), normal = make_image(vk, VK_FORMAT_R32G32_SFLOAT, width, height),
                    albedo = make_image(vk, VK_FORMAT_R8G8B8A8_UNORM, width, height), illum = make_image(vk, VK_FORMAT_R32G32B32A32_SFLOAT, width, height),
                    presented = make_image(vk, VK_FORMAT_B8G8R8A8_UNORM, width, height);
        const VkDeviceSize px = (VkDeviceSize)width * height;
        HostBuffer s_depth = make_host_buffer(vk, px * 4), s_normal = make_host_buffer(vk, px * 8), s_albedo = make_host_buffer(vk, px * 4),
                   s_illum = make_host_buffer(vk, px * 16), readback = make_host_buffer(vk, px * 4);
        VkCommandPool pool;
        VkCommandPoolCreateInfo pci{};
        pci.sType = VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO;
        pci.flags = VK_COMMAND_POOL_CREATE_RESET_COMMAND_BUFFER_BIT;
        vk::vk_check(vk.dev<PFN_vkCreateCommandPool>("vkCreateCommandPool")(vk.device, &pci, nullptr, &pool), "vkCreateCommandPool");
        VkCommandBuffer cbs[2];
        VkCommandBufferAllocateInfo cai{};
        cai.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO;
        cai.commandPool = pool; cai.level = VK_COMMAND_BUFFER_LEVEL_PRIMARY; cai.commandBufferCount = 2;
        vk::vk_check(vk.dev<PFN_vkAllocateCommandBuffers>("vkAllocateCommandBuffers")(vk.device, &cai, cbs), "vkAllocateCommandBuffers");
        const auto begin = vk.dev<PFN_vkBeginCommandBuffer>("vkBeginCommandBuffer");
        const auto end = vk.dev<PFN_vkEndCommandBuffer>("vkEndCommandBuffer");
        const auto submit = vk.dev<PFN_vkQueueSubmit>("vkQueueSubmit");
        const auto copy_b2i = vk.dev<PFN_vkCmdCopyBufferToImage>("vkCmdCopyBufferToImage");
        const auto copy_i2b = vk.dev<PFN_vkCmdCopyImageToBuffer>("vkCmdCopyImageToBuffer");
        const auto barrier = vk.dev<PFN_vkCmdPipelineBarrier>("vkCmdPipelineBarrier");
        VkCommandBufferBeginInfo bi{};
        bi.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO;
        bi.flags = VK_COMMAND_BUFFER_USAGE_ONE_TIME_SUBMIT_BIT;
        const VkBufferImageCopy region = whole_image(width, height);

        for (int frame_index = 0; frame_index < num_frames; ++frame_index) {
            const std::string base = dir + "/frame_" + std::to_string(frame_index);
            const auto d = slurp(base + ".depth"), n = slurp(base + ".normal"), a = slurp(base + ".albedo"), il = slurp(base + ".illum");
            const auto cam = slurp(base + ".cam");   // view, inv_view, proj, inv_proj: 4 x 16 floats
            const float* cm = reinterpret_cast<const float*>(cam.data());

            // "render": the previous frame's planes must have been consumed before these images' copies overwrite them --
            // that is producer_submit()'s wait; the staging buffers themselves are free once the previous submit has run
            if (frame_index > 0) frame->consumed->host_wait(frame_index);      // (staging reuse only; a real renderer has no staging)
            memcpy(s_depth.map, d.data(), d.size()); memcpy(s_normal.map, n.data(), n.size());
            memcpy(s_albedo.map, a.data(), a.size()); memcpy(s_illum.map, il.data(), il.size());
            VkCommandBuffer cb = cbs[0];
            vk::vk_check(begin(cb, &bi), "vkBeginCommandBuffer");
            struct Up { RenderImage* img; HostBuffer* src; } ups[4] = {{&depth, &s_depth}, {&normal, &s_normal}, {&albedo, &s_albedo}, {&illum, &s_illum}};
            for (auto& u : ups) {
                image_barrier(barrier, cb, u.img->image, VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL);
                copy_b2i(cb, u.src->buffer, u.img->image, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region);
                image_barrier(barrier, cb, u.img->image, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, VK_IMAGE_LAYOUT_GENERAL);    // where ptRaygen.rgen leaves them
            }
            // hand the frame over: G-buffer and illumination into the exported planes
            frame->depth->cmd_copy_from_image(cb, depth.image);
            frame->normal->cmd_copy_from_image(cb, normal.image);
            frame->albedo->cmd_copy_from_image(cb, albedo.image);
            frame->illumination->cmd_copy_from_image(cb, illum.image);
            vk::vk_check(end(cb), "vkEndCommandBuffer");
            {
                vk::TimelineSubmit ts = frame->producer_submit(frame_index);
                VkSubmitInfo si{};
                si.sType = VK_STRUCTURE_TYPE_SUBMIT_INFO;
                si.commandBufferCount = 1; si.pCommandBuffers = &cb;
                ts.apply(si);
                vk::vk_check(submit(vk.queue, 1, &si, VK_NULL_HANDLE), "vkQueueSubmit (producer)");
            }

            // CUDA side: VulkanPBRT.cpp:561-588
            auto& pc = ray_tracing_push_constants->value();
            CameraMatrices cur, prev;
            for (int i = 0; i < 16; ++i) pc.view_inverse.m[i] = cur.inv_view.m[i] = cm[16 + i];
            cur.proj = mat4(); cur.inv_proj = mat4();
            for (int i = 0; i < 16; ++i) { cur.proj->m[i] = cm[32 + i]; cur.inv_proj->m[i] = pc.proj_inverse.m[i] = cm[48 + i]; }
            pc.frame_number = frame_index;
            prev.view = pc.prev_view;
            accumulator->set_camera_matrices(frame_index, cur, prev);
            commands->record();
            for (int i = 0; i < 16; ++i) pc.prev_view.m[i] = cm[i];

            // present: the denoised plane into the displayed image, and (for the test) back to the host
            cb = cbs[1];
            vk::vk_check(begin(cb, &bi), "vkBeginCommandBuffer");
            frame->final_plane->cmd_copy_to_image(cb, presented.image, VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL);
            copy_i2b(cb, presented.image, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, readback.buffer, 1, &region);
            vk::vk_check(end(cb), "vkEndCommandBuffer");
            {
                vk::TimelineSubmit ts = frame->presenter_submit(frame_index);
                VkSubmitInfo si{};
                si.sType = VK_STRUCTURE_TYPE_SUBMIT_INFO;
                si.commandBufferCount = 1; si.pCommandBuffers = &cb;
                ts.apply(si);
                vk::vk_check(submit(vk.queue, 1, &si, VK_NULL_HANDLE), "vkQueueSubmit (presenter)");
            }
            vk::vk_check(vk.dev<PFN_vkQueueWaitIdle>("vkQueueWaitIdle")(vk.queue), "vkQueueWaitIdle");     // the read-back only
            std::ofstream(dir + "/final_" + std::to_string(frame_index) + ".bgra", std::ios::binary).write(readback.map, (std::streamsize)(px * 4));
        }
        if (frame->produced->value() != (uint64_t)num_frames || frame->consumed->value() != (uint64_t)num_frames)
            throw std::runtime_error("timeline payloads do not equal the number of frames");
        context.waitForCompletion();

        // ---- teardown: CUDA objects that alias Vulkan memory first, then the Vulkan objects ----
        commands.reset(); accumulator.reset(); g_buffer.reset(); illumination_buffer.reset(); accumulation_buffer.reset(); final_descriptor_image.reset();
        frame.reset();
        vk.dev<PFN_vkFreeCommandBuffers>("vkFreeCommandBuffers")(vk.device, pool, 2, cbs);
        vk.dev<PFN_vkDestroyCommandPool>("vkDestroyCommandPool")(vk.device, pool, nullptr);
        for (HostBuffer* b : {&s_depth, &s_normal, &s_albedo, &s_illum, &readback}) {
            vk.dev<PFN_vkDestroyBuffer>("vkDestroyBuffer")(vk.device, b->buffer, nullptr);
            vk.dev<PFN_vkFreeMemory>("vkFreeMemory")(vk.device, b->memory, nullptr);
        }
        for (RenderImage* i : {&depth, &normal, &albedo, &illum, &presented}) {
            vk.dev<PFN_vkDestroyImage>("vkDestroyImage")(vk.device, i->image, nullptr);
            vk.dev<PFN_vkFreeMemory>("vkFreeMemory")(vk.device, i->memory, nullptr);
        }
        vk.dev<PFN_vkDestroyDevice>("vkDestroyDevice")(vk.device, nullptr);
        vk.inst<PFN_vkDestroyInstance>("vkDestroyInstance")(vk.instance, nullptr);
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
