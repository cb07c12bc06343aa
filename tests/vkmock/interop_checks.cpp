// interop_checks.cpp -- TEST-ONLY: edge cases of include/vkpbrt/vk_interop.hpp and of the C ABI's import entry points,
// against tests/vkmock (Vulkan side) and the test emulator (CUDA side).  Prints one "ok <case>" line per case; any
// failure throws and exits 1.  Driven by tests/test_vk_interop.py.
#include <dirent.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <thread>
#include <vector>

#include "vkpbrt/vk_interop.hpp"

using namespace vkpbrt;

#define REQUIRE(cond) do { if (!(cond)) throw std::runtime_error(std::string("check failed: ") + #cond + " (line " + std::to_string(__LINE__) + ")"); } while (0)

static int open_fds()
{
    int n = 0;
    DIR* d = opendir("/proc/self/fd");
    while (d && readdir(d)) ++n;
    if (d) closedir(d);
    return n;
}

template <typename F>
static bool throws_with(F&& f, const char* needle)
{
    try { f(); } catch (const std::exception& e) { return std::string(e.what()).find(needle) != std::string::npos; }
    return false;
}

struct Vk {
    PFN_vkGetInstanceProcAddr gipa;
    VkInstance instance = VK_NULL_HANDLE;
    VkPhysicalDevice physical_device = VK_NULL_HANDLE;
    template <typename F> F inst(const char* name) const { return reinterpret_cast<F>(gipa(instance, name)); }
    VkDevice make_device(bool extensions, bool timeline) const
    {
        const float priority = 1.f;
        VkDeviceQueueCreateInfo qi{};
        qi.sType = VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO;
        qi.queueCount = 1; qi.pQueuePriorities = &priority;
        VkPhysicalDeviceVulkan12Features f12{};
        f12.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_VULKAN_1_2_FEATURES;
        f12.timelineSemaphore = timeline ? VK_TRUE : VK_FALSE;
        const std::vector<const char*> ext = vk::required_device_extensions();
        VkDeviceCreateInfo ci{};
        ci.sType = VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO;
        ci.pNext = &f12;
        ci.queueCreateInfoCount = 1; ci.pQueueCreateInfos = &qi;
        if (extensions) { ci.enabledExtensionCount = (uint32_t)ext.size(); ci.ppEnabledExtensionNames = ext.data(); }
        VkDevice device = VK_NULL_HANDLE;
        vk::vk_check(inst<PFN_vkCreateDevice>("vkCreateDevice")(physical_device, &ci, nullptr, &device), "vkCreateDevice");
        return device;
    }
    void destroy_device(VkDevice d) const { reinterpret_cast<PFN_vkDestroyDevice>(inst<PFN_vkGetDeviceProcAddr>("vkGetDeviceProcAddr")(d, "vkDestroyDevice"))(d, nullptr); }
};

int main()
{
    try {
        Vk v;
        v.gipa = vk::open_loader(getenv("VKPBRT_VULKAN_LIBRARY"));
        REQUIRE(throws_with([] { vk::open_loader("/nonexistent/libvulkan.so.1"); }, "cannot load"));
        VkApplicationInfo app{};
        app.sType = VK_STRUCTURE_TYPE_APPLICATION_INFO;
        app.apiVersion = VK_API_VERSION_1_2;
        VkInstanceCreateInfo ici{};
        ici.sType = VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO;
        ici.pApplicationInfo = &app;
        vk::vk_check(reinterpret_cast<PFN_vkCreateInstance>(v.gipa(VK_NULL_HANDLE, "vkCreateInstance"))(&ici, nullptr, &v.instance), "vkCreateInstance");
        uint32_t n = 1;
        v.inst<PFN_vkEnumeratePhysicalDevices>("vkEnumeratePhysicalDevices")(v.instance, &n, &v.physical_device);

        // a device without the *_fd extensions / without timeline semaphores: Api::load says what is missing
        {
            VkDevice bare = v.make_device(false, true);
            REQUIRE(throws_with([&] { vk::Api::load(v.gipa, v.instance, v.physical_device, bare); }, "vkGetMemoryFdKHR"));
            v.destroy_device(bare);
            VkDevice no_timeline = v.make_device(true, false);
            REQUIRE(throws_with([&] { vk::Api::load(v.gipa, v.instance, v.physical_device, no_timeline); }, "vkSignalSemaphore"));
            v.destroy_device(no_timeline);
            REQUIRE(throws_with([&] { vk::Api::load(v.gipa, v.instance, v.physical_device, VK_NULL_HANDLE); }, "null handle"));
            puts("ok api_load_reports_missing_entry_points");
        }

        VkDevice device = v.make_device(true, true);
        const vk::Api api = vk::Api::load(v.gipa, v.instance, v.physical_device, device);
        REQUIRE(vk::cuda_device_of(api) == 0);
        puts("ok device_uuid_match");
        Context context(0);
        make_current(context);

        // C ABI argument checks of the import entry points
        {
            vkpbrt_external_memory_t m = nullptr; void* p = nullptr; vkpbrt_external_semaphore_t s = nullptr;
            REQUIRE(vkpbrt_import_external_memory_fd(context.handle, -1, 4096, 0, 4096, &m, &p) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_import_external_memory_fd(context.handle, 0, 4096, 4096, 1, &m, &p) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_import_external_memory_fd(context.handle, 0, 4096, 0, 0, &m, &p) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_import_external_memory_fd(context.handle, 0, 4096, ~0ull - 8, 64, &m, &p) == VKPBRT_ERR_INVALID_ARGUMENT);   // offset + size wraps
            REQUIRE(vkpbrt_import_external_memory_fd(nullptr, 0, 4096, 0, 64, &m, &p) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_import_external_semaphore_fd(nullptr, 0, 1, &s) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_external_semaphore_wait(nullptr, 1) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_external_memory_destroy(nullptr) == VKPBRT_OK && vkpbrt_external_semaphore_destroy(nullptr) == VKPBRT_OK);
            // an fd that is not an exported allocation: the import fails, and the fd still belongs to the caller
            int pipefd[2];
            REQUIRE(pipe(pipefd) == 0);
            REQUIRE(vkpbrt_import_external_memory_fd(context.handle, pipefd[0], 4096, 0, 4096, &m, &p) == VKPBRT_ERR_CUDA);
            REQUIRE(close(pipefd[0]) == 0 && close(pipefd[1]) == 0);
            puts("ok import_argument_checks");
        }

        // a plane is the same bytes on both sides; nothing leaks when it goes away
        const int fds_before = open_fds();
        {
            const uint32_t W = 37, H = 11;                                   // odd sizes: allocation > plane, rows tight
            auto plane = vk::SharedPlane::create(api, context, VKPBRT_FORMAT_R32G32_SFLOAT, W, H);
            const vkpbrt_image_info info = plane->image->info();
            REQUIRE(info.row_pitch == W * 8 && info.size_bytes == plane->size_bytes && plane->allocation_size >= plane->size_bytes && !info.owned);
            std::vector<float> src(W * H * 2), back(W * H * 2, -1.f);
            for (size_t i = 0; i < src.size(); ++i) src[i] = (float)i * 0.5f;
            check(vkpbrt_image_upload(plane->image->handle, src.data(), src.size() * 4));         // CUDA writes ...
            context.waitForCompletion();
            // ... Vulkan reads: plane -> OPTIMAL image -> host-visible buffer, with the interop's own copy commands
            auto dev = [&](const char* name) { return api.GetDeviceProcAddr(device, name); };
            VkImage image; VkDeviceMemory image_memory, host_memory; VkBuffer host_buffer;
            VkImageCreateInfo ci{};
            ci.sType = VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO;
            ci.imageType = VK_IMAGE_TYPE_2D; ci.format = VK_FORMAT_R32G32_SFLOAT; ci.extent = {W, H, 1}; ci.mipLevels = 1; ci.arrayLayers = 1;
            ci.samples = VK_SAMPLE_COUNT_1_BIT; ci.tiling = VK_IMAGE_TILING_OPTIMAL;
            ci.usage = VK_IMAGE_USAGE_TRANSFER_SRC_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT;
            vk::vk_check(reinterpret_cast<PFN_vkCreateImage>(dev("vkCreateImage"))(device, &ci, nullptr, &image), "vkCreateImage");
            VkMemoryRequirements req;
            reinterpret_cast<PFN_vkGetImageMemoryRequirements>(dev("vkGetImageMemoryRequirements"))(device, image, &req);
            VkMemoryAllocateInfo ai{};
            ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
            ai.allocationSize = req.size; ai.memoryTypeIndex = 1;
            vk::vk_check(api.AllocateMemory(device, &ai, nullptr, &image_memory), "vkAllocateMemory");
            vk::vk_check(reinterpret_cast<PFN_vkBindImageMemory>(dev("vkBindImageMemory"))(device, image, image_memory, 0), "vkBindImageMemory");
            VkBufferCreateInfo bci{};
            bci.sType = VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO;
            bci.size = plane->size_bytes; bci.usage = VK_BUFFER_USAGE_TRANSFER_DST_BIT | VK_BUFFER_USAGE_TRANSFER_SRC_BIT;
            vk::vk_check(api.CreateBuffer(device, &bci, nullptr, &host_buffer), "vkCreateBuffer");
            api.GetBufferMemoryRequirements(device, host_buffer, &req);
            ai.allocationSize = req.size; ai.memoryTypeIndex = 0;
            vk::vk_check(api.AllocateMemory(device, &ai, nullptr, &host_memory), "vkAllocateMemory");
            vk::vk_check(api.BindBufferMemory(device, host_buffer, host_memory, 0), "vkBindBufferMemory");
            void* map = nullptr;
            vk::vk_check(reinterpret_cast<PFN_vkMapMemory>(dev("vkMapMemory"))(device, host_memory, 0, VK_WHOLE_SIZE, 0, &map), "vkMapMemory");
            VkCommandPool pool; VkCommandBuffer cb;
            VkCommandPoolCreateInfo pci{};
            pci.sType = VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO;
            vk::vk_check(reinterpret_cast<PFN_vkCreateCommandPool>(dev("vkCreateCommandPool"))(device, &pci, nullptr, &pool), "vkCreateCommandPool");
            VkCommandBufferAllocateInfo cai{};
            cai.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO;
            cai.commandPool = pool; cai.commandBufferCount = 1;
            vk::vk_check(reinterpret_cast<PFN_vkAllocateCommandBuffers>(dev("vkAllocateCommandBuffers"))(device, &cai, &cb), "vkAllocateCommandBuffers");
            VkCommandBufferBeginInfo bi{};
            bi.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO;
            vk::vk_check(reinterpret_cast<PFN_vkBeginCommandBuffer>(dev("vkBeginCommandBuffer"))(cb, &bi), "vkBeginCommandBuffer");
            plane->cmd_copy_to_image(cb, image, VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL);
            VkBufferImageCopy region{};
            region.imageSubresource.aspectMask = VK_IMAGE_ASPECT_COLOR_BIT; region.imageSubresource.layerCount = 1;
            region.imageExtent = {W, H, 1};
            api.CmdCopyImageToBuffer(cb, image, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, host_buffer, 1, &region);
            vk::vk_check(reinterpret_cast<PFN_vkEndCommandBuffer>(dev("vkEndCommandBuffer"))(cb), "vkEndCommandBuffer");
            VkQueue queue;
            reinterpret_cast<PFN_vkGetDeviceQueue>(dev("vkGetDeviceQueue"))(device, 0, 0, &queue);
            VkSubmitInfo si{};
            si.sType = VK_STRUCTURE_TYPE_SUBMIT_INFO;
            si.commandBufferCount = 1; si.pCommandBuffers = &cb;
            vk::vk_check(reinterpret_cast<PFN_vkQueueSubmit>(dev("vkQueueSubmit"))(queue, 1, &si, VK_NULL_HANDLE), "vkQueueSubmit");
            REQUIRE(memcmp(map, src.data(), src.size() * 4) == 0);
            // and the other way: Vulkan writes the plane (image -> plane), CUDA reads it
            for (size_t i = 0; i < src.size(); ++i) ((float*)map)[i] = -(float)i;
            vk::vk_check(reinterpret_cast<PFN_vkBeginCommandBuffer>(dev("vkBeginCommandBuffer"))(cb, &bi), "vkBeginCommandBuffer");
            {
                VkImageMemoryBarrier b{};
                b.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
                b.oldLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL; b.newLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL;
                b.image = image; b.subresourceRange = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, 1};
                api.CmdPipelineBarrier(cb, VK_PIPELINE_STAGE_TRANSFER_BIT, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, 0, nullptr, 0, nullptr, 1, &b);
                api.CmdCopyBufferToImage(cb, host_buffer, image, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region);
                b.oldLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL; b.newLayout = VK_IMAGE_LAYOUT_GENERAL;
                api.CmdPipelineBarrier(cb, VK_PIPELINE_STAGE_TRANSFER_BIT, VK_PIPELINE_STAGE_ALL_COMMANDS_BIT, 0, 0, nullptr, 0, nullptr, 1, &b);
            }
            plane->cmd_copy_from_image(cb, image);                                                  // GENERAL -> copy -> GENERAL
            vk::vk_check(reinterpret_cast<PFN_vkEndCommandBuffer>(dev("vkEndCommandBuffer"))(cb), "vkEndCommandBuffer");
            vk::vk_check(reinterpret_cast<PFN_vkQueueSubmit>(dev("vkQueueSubmit"))(queue, 1, &si, VK_NULL_HANDLE), "vkQueueSubmit");
            check(vkpbrt_image_download(plane->image->handle, back.data(), back.size() * 4));
            context.waitForCompletion();
            for (size_t i = 0; i < back.size(); ++i) REQUIRE(back[i] == -(float)i);
            reinterpret_cast<PFN_vkFreeCommandBuffers>(dev("vkFreeCommandBuffers"))(device, pool, 1, &cb);
            reinterpret_cast<PFN_vkDestroyCommandPool>(dev("vkDestroyCommandPool"))(device, pool, nullptr);
            api.DestroyBuffer(device, host_buffer, nullptr);
            api.FreeMemory(device, host_memory, nullptr);
            reinterpret_cast<PFN_vkDestroyImage>(dev("vkDestroyImage"))(device, image, nullptr);
            api.FreeMemory(device, image_memory, nullptr);
        }
        REQUIRE(open_fds() == fds_before);
        puts("ok shared_plane_round_trip_no_fd_leak");

        // image_copy_record: what copy_final_image and add_signal_to_commands rest on
        {
            auto a = vk::SharedPlane::create(api, context, VKPBRT_FORMAT_B8G8R8A8_UNORM, 16, 4);
            DescriptorImage rgba(context, VKPBRT_FORMAT_R8G8B8A8_UNORM, 16, 4), wrong_size(context, VKPBRT_FORMAT_R8G8B8A8_UNORM, 16, 5),
                wrong_texel(context, VKPBRT_FORMAT_R16G16B16A16_SFLOAT, 16, 4);
            rgba.compile(context); wrong_size.compile(context); wrong_texel.compile(context);
            std::vector<uint8_t> bytes(16 * 4 * 4), got(16 * 4 * 4);
            for (size_t i = 0; i < bytes.size(); ++i) bytes[i] = (uint8_t)(i * 7);
            check(vkpbrt_image_upload(rgba.handle, bytes.data(), bytes.size()));
            check(vkpbrt_image_copy_record(rgba.handle, a->image->handle));       // size-compatible formats, like vkCmdCopyImage
            check(vkpbrt_image_download(a->image->handle, got.data(), got.size()));
            context.waitForCompletion();
            REQUIRE(got == bytes);
            REQUIRE(vkpbrt_image_copy_record(wrong_size.handle, a->image->handle) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_image_copy_record(wrong_texel.handle, a->image->handle) == VKPBRT_ERR_INVALID_ARGUMENT);
            REQUIRE(vkpbrt_image_copy_record(nullptr, a->image->handle) == VKPBRT_ERR_INVALID_ARGUMENT);
            puts("ok image_copy_record");
        }

        // timeline semaphore seen from both sides, with the waiter really blocked
        {
            auto t = vk::SharedTimeline::create(api, context, 5);
            REQUIRE(t->value() == 5);
            t->cuda_signal(6);
            REQUIRE(t->value() == 6);
            REQUIRE(!t->host_wait(7, 2'000'000));                 // 2 ms: not there yet
            t->host_signal(7);
            REQUIRE(t->host_wait(7, 0));
            t->cuda_wait(7);                                      // already satisfied
            REQUIRE(throws_with([&] { t->cuda_signal(7); }, "vkpbrt"));     // timeline values must increase
            std::atomic<int> stage{0};
            std::thread cuda_side([&] { stage = 1; t->cuda_wait(9); stage = 2; t->cuda_signal(10); });
            while (stage.load() == 0) std::this_thread::yield();
            std::this_thread::sleep_for(std::chrono::milliseconds(30));
            REQUIRE(stage.load() == 1);                           // blocked in the wait
            t->host_signal(8);
            std::this_thread::sleep_for(std::chrono::milliseconds(30));
            REQUIRE(stage.load() == 1);                           // 8 < 9
            t->host_signal(9);
            REQUIRE(t->host_wait(10, 5'000'000'000ull));
            cuda_side.join();
            REQUIRE(stage.load() == 2 && t->value() == 10);
        }
        REQUIRE(open_fds() == fds_before);
        puts("ok shared_timeline_both_sides");

        // the frame object: buffers of the right types and sizes, handshake values
        {
            auto frame = vk::SharedFrame::create(api, context, 64, 48);
            REQUIRE(frame->g_buffer->width == 64 && frame->g_buffer->height == 48);
            REQUIRE(frame->g_buffer->depth->info().data == frame->depth->image->info().data);
            REQUIRE(frame->illumination_buffer->illumination_images.size() == 1);
            REQUIRE(frame->illumination_buffer->illumination_images[0]->info().format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT);
            vk::TimelineSubmit p0 = frame->producer_submit(0), p3 = frame->producer_submit(3), s3 = frame->presenter_submit(3);
            REQUIRE(p0.wait_semaphore == VK_NULL_HANDLE && p0.signal_semaphore == frame->produced->semaphore && p0.signal_value == 1);
            REQUIRE(p3.wait_semaphore == frame->consumed->semaphore && p3.wait_value == 3 && p3.signal_value == 4);
            REQUIRE(s3.wait_semaphore == frame->consumed->semaphore && s3.wait_value == 4 && s3.signal_semaphore == VK_NULL_HANDLE);
            VkSubmitInfo si{};
            si.sType = VK_STRUCTURE_TYPE_SUBMIT_INFO;
            p3.apply(si);
            REQUIRE(si.pNext == &p3.timeline && si.waitSemaphoreCount == 1 && si.signalSemaphoreCount == 1 && *p3.timeline.pWaitSemaphoreValues == 3 &&
                    *p3.timeline.pSignalSemaphoreValues == 4);
        }
        REQUIRE(open_fds() == fds_before);
        puts("ok shared_frame");

        v.destroy_device(device);
        v.inst<PFN_vkDestroyInstance>("vkDestroyInstance")(v.instance, nullptr);
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
