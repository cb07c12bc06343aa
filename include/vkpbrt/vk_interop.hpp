// vk_interop.hpp -- the Vulkan side of the drop-in boundary (SURVEY.md section 8(b) "Sync with Vulkan", 8(f)-2).
//
// VulkanPBRT renders with Vulkan (ray tracing pipeline, vsg); the denoising path of this repository runs as CUDA
// kernels on the same GPU.  The two meet in device memory that Vulkan allocates and exports
// (VK_KHR_external_memory_fd) and CUDA imports (vkpbrt_import_external_memory_fd), ordered by timeline semaphores
// exported the same way (VK_KHR_external_semaphore_fd).  This header is what a VulkanPBRT maintainer adds on the
// Vulkan side:
//
//   vk::Api            function table, loaded through the application's vkGetInstanceProcAddr (no link-time
//                      dependency on libvulkan; vk::open_loader() dlopens it for hosts that do not link it)
//   vk::cuda_device_of the CUDA device that IS the VkPhysicalDevice (UUID match) -- external memory can only be
//                      imported on the exporting device
//   vk::SharedPlane    one tightly packed, pitch-linear plane: an exportable VkBuffer + its DescriptorImage on
//                      the CUDA side, with the copy commands between it and the renderer's (TILING_OPTIMAL)
//                      VkImages -- the kernels need linear memory, an optimally tiled image cannot be viewed as a
//                      pointer (include/vkpbrt_b200.h, "Model")
//   vk::SharedTimeline one timeline semaphore seen from both APIs
//   vk::SharedFrame    the planes of the path (depth, normal, albedo, raw illumination in; final BGRA8 out) as a
//                      GBuffer / IlluminationBufferDemodulatedFloat pair, plus the per-frame handshake
//
// Compile-guarded: without Vulkan headers the header defines VKPBRT_HAVE_VULKAN 0 and nothing else.  Define
// VKPBRT_VULKAN_HEADER to the header to use (e.g. <vulkan/vulkan.h>) if it is not <vulkan/vulkan_core.h>.
// Requires Vulkan 1.2 (timeline semaphores, external memory / semaphore capabilities in core), which is what the
// reference requests (VulkanPBRT.cpp:176), plus the two *_fd device extensions of required_device_extensions(),
// to be appended to window_traits->deviceExtensionNames (VulkanPBRT.cpp:171-175), and the timelineSemaphore feature
// (next to the other VkPhysicalDeviceVulkan12Features the reference enables, VulkanPBRT.cpp:185-192).
//
// tests/test_vk_interop.py compiles examples/cpp_vulkan_interop.cpp (a raw-Vulkan host written against this
// header) and runs it against tests/vkmock, a mock Vulkan implementation whose exported memory and semaphores are
// memfds, with the CUDA side on the test emulator -- neither machine of this project has a Vulkan driver.
#pragma once

#if defined(VKPBRT_VULKAN_HEADER)
#include VKPBRT_VULKAN_HEADER
#define VKPBRT_HAVE_VULKAN 1
#elif defined(VULKAN_CORE_H_)
#define VKPBRT_HAVE_VULKAN 1
#elif defined(__has_include)
#if __has_include(<vulkan/vulkan_core.h>)
#include <vulkan/vulkan_core.h>
#define VKPBRT_HAVE_VULKAN 1
#endif
#endif
#ifndef VKPBRT_HAVE_VULKAN
#define VKPBRT_HAVE_VULKAN 0
#endif

#if VKPBRT_HAVE_VULKAN

#include <dlfcn.h>
#include <unistd.h>

#include <cstring>

#include "vkpbrt.hpp"

namespace vkpbrt {
namespace vk {

inline void vk_check(VkResult r, const char* what)
{
    if (r != VK_SUCCESS) throw std::runtime_error(std::string("vkpbrt::vk: ") + what + " failed (VkResult " + std::to_string((int)r) + ")");
}

// device extensions the interop needs on top of Vulkan 1.2
inline std::vector<const char*> required_device_extensions()
{
    return {VK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME, VK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME};
}

// for hosts that do not link libvulkan: returns its vkGetInstanceProcAddr
inline PFN_vkGetInstanceProcAddr open_loader(const char* library = nullptr)
{
    const char* name = (library && *library) ? library : "libvulkan.so.1";
    void* lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("vkpbrt::vk: cannot load ") + name + ": " + dlerror());
    auto gipa = reinterpret_cast<PFN_vkGetInstanceProcAddr>(dlsym(lib, "vkGetInstanceProcAddr"));
    if (!gipa) throw std::runtime_error(std::string("vkpbrt::vk: ") + name + " has no vkGetInstanceProcAddr");
    return gipa;
}

// The entry points the interop calls, resolved once.  vsg applications pass vkGetInstanceProcAddr of the loader
// they link, *instance / *physicalDevice / *device (vsg::Instance, vsg::PhysicalDevice and vsg::Device convert to
// their Vulkan handles).
struct Api {
    VkInstance instance = VK_NULL_HANDLE;
    VkPhysicalDevice physical_device = VK_NULL_HANDLE;
    VkDevice device = VK_NULL_HANDLE;
#define VKPBRT_VK_INSTANCE_FN(X) X(GetPhysicalDeviceProperties2) X(GetPhysicalDeviceMemoryProperties) X(GetPhysicalDeviceExternalBufferProperties) \
    X(GetPhysicalDeviceExternalSemaphoreProperties) X(GetDeviceProcAddr)
#define VKPBRT_VK_DEVICE_FN(X) X(CreateBuffer) X(DestroyBuffer) X(GetBufferMemoryRequirements) X(AllocateMemory) X(FreeMemory) X(BindBufferMemory) \
    X(GetMemoryFdKHR) X(CreateSemaphore) X(DestroySemaphore) X(GetSemaphoreFdKHR) X(SignalSemaphore) X(WaitSemaphores) X(GetSemaphoreCounterValue) \
    X(CmdPipelineBarrier) X(CmdCopyImageToBuffer) X(CmdCopyBufferToImage)
#define VKPBRT_VK_DECLARE(name) PFN_vk##name name = nullptr;
    VKPBRT_VK_INSTANCE_FN(VKPBRT_VK_DECLARE)
    VKPBRT_VK_DEVICE_FN(VKPBRT_VK_DECLARE)
#undef VKPBRT_VK_DECLARE

    static Api load(PFN_vkGetInstanceProcAddr get_instance_proc_addr, VkInstance instance, VkPhysicalDevice physical_device, VkDevice device)
    {
        if (!get_instance_proc_addr || !instance || !physical_device || !device) throw std::runtime_error("vkpbrt::vk::Api::load: null handle");
        Api a;
        a.instance = instance; a.physical_device = physical_device; a.device = device;
#define VKPBRT_VK_LOAD_I(name) \
        a.name = reinterpret_cast<PFN_vk##name>(get_instance_proc_addr(instance, "vk" #name)); \
        if (!a.name) throw std::runtime_error("vkpbrt::vk: the instance does not provide vk" #name " (Vulkan 1.1+ required)");
        VKPBRT_VK_INSTANCE_FN(VKPBRT_VK_LOAD_I)
#undef VKPBRT_VK_LOAD_I
#define VKPBRT_VK_LOAD_D(name) \
        a.name = reinterpret_cast<PFN_vk##name>(a.GetDeviceProcAddr(device, "vk" #name)); \
        if (!a.name) throw std::runtime_error("vkpbrt::vk: the device does not provide vk" #name " -- create it with Vulkan 1.2, the timelineSemaphore " \
                                              "feature and the extensions of vkpbrt::vk::required_device_extensions()");
        VKPBRT_VK_DEVICE_FN(VKPBRT_VK_LOAD_D)
#undef VKPBRT_VK_LOAD_D
        return a;
    }
#undef VKPBRT_VK_INSTANCE_FN
#undef VKPBRT_VK_DEVICE_FN
};

// The CUDA device ordinal of the Vulkan physical device (VkPhysicalDeviceIDProperties::deviceUUID against
// cudaDeviceProp::uuid, through vkpbrt_device_uuid).  Throws when no CUDA device matches.
inline int cuda_device_of(const Api& api)
{
    VkPhysicalDeviceIDProperties id{};
    id.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_ID_PROPERTIES;
    VkPhysicalDeviceProperties2 props{};
    props.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_PROPERTIES_2;
    props.pNext = &id;
    api.GetPhysicalDeviceProperties2(api.physical_device, &props);
    int count = 0;
    check(vkpbrt_device_count(&count));
    for (int d = 0; d < count; ++d) {
        uint8_t uuid[16];
        check(vkpbrt_device_uuid(d, uuid));
        if (std::memcmp(uuid, id.deviceUUID, VK_UUID_SIZE) == 0) return d;
    }
    throw std::runtime_error(std::string("vkpbrt::vk: no CUDA device has the UUID of the Vulkan device '") + props.properties.deviceName + "'");
}

inline VkFormat to_vk_format(uint32_t format)
{
    switch (format) {
    case VKPBRT_FORMAT_R32_SFLOAT: return VK_FORMAT_R32_SFLOAT;
    case VKPBRT_FORMAT_R32G32_SFLOAT: return VK_FORMAT_R32G32_SFLOAT;
    case VKPBRT_FORMAT_R8G8B8A8_UNORM: return VK_FORMAT_R8G8B8A8_UNORM;
    case VKPBRT_FORMAT_B8G8R8A8_UNORM: return VK_FORMAT_B8G8R8A8_UNORM;
    case VKPBRT_FORMAT_R16G16_SFLOAT: return VK_FORMAT_R16G16_SFLOAT;
    case VKPBRT_FORMAT_R8_UNORM: return VK_FORMAT_R8_UNORM;
    case VKPBRT_FORMAT_R16G16B16A16_SFLOAT: return VK_FORMAT_R16G16B16A16_SFLOAT;
    case VKPBRT_FORMAT_R32G32B32A32_SFLOAT: return VK_FORMAT_R32G32B32A32_SFLOAT;
    case VKPBRT_FORMAT_R16_SFLOAT: return VK_FORMAT_R16_SFLOAT;
    default: return VK_FORMAT_UNDEFINED;
    }
}

// One plane shared by the two APIs.  Vulkan: an exportable VkBuffer (TRANSFER_SRC | TRANSFER_DST | STORAGE), rows
// tightly packed.  CUDA: `image`, a DescriptorImage over the same bytes, usable wherever the modules take one.
class SharedPlane : public Inherit<SharedPlane> {
public:
    SharedPlane(const Api& api, Context& ctx, uint32_t format, uint32_t width, uint32_t height) : _api(api), format(format), width(width), height(height)
    {
        const uint32_t texel = vkpbrt_format_texel_size(format);
        if (!texel || !width || !height) throw std::runtime_error("vkpbrt::vk::SharedPlane: bad format or extent");
        size_bytes = (VkDeviceSize)width * height * texel;
        const VkBufferUsageFlags usage = VK_BUFFER_USAGE_TRANSFER_SRC_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT | VK_BUFFER_USAGE_STORAGE_BUFFER_BIT;

        // can this driver export such a buffer as an opaque fd, and only as a dedicated allocation?
        VkPhysicalDeviceExternalBufferInfo ext_info{};
        ext_info.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_EXTERNAL_BUFFER_INFO;
        ext_info.usage = usage;
        ext_info.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkExternalBufferProperties ext_props{};
        ext_props.sType = VK_STRUCTURE_TYPE_EXTERNAL_BUFFER_PROPERTIES;
        api.GetPhysicalDeviceExternalBufferProperties(api.physical_device, &ext_info, &ext_props);
        const VkExternalMemoryFeatureFlags features = ext_props.externalMemoryProperties.externalMemoryFeatures;
        if (!(features & VK_EXTERNAL_MEMORY_FEATURE_EXPORTABLE_BIT))
            throw std::runtime_error("vkpbrt::vk::SharedPlane: the Vulkan device cannot export buffers as opaque file descriptors");
        dedicated = (features & VK_EXTERNAL_MEMORY_FEATURE_DEDICATED_ONLY_BIT) != 0;

        VkExternalMemoryBufferCreateInfo ext_buffer{};
        ext_buffer.sType = VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO;
        ext_buffer.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkBufferCreateInfo bi{};
        bi.sType = VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO;
        bi.pNext = &ext_buffer;
        bi.size = size_bytes;
        bi.usage = usage;
        bi.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
        vk_check(api.CreateBuffer(api.device, &bi, nullptr, &buffer), "vkCreateBuffer");
        try {
            VkMemoryRequirements req{};
            api.GetBufferMemoryRequirements(api.device, buffer, &req);
            VkPhysicalDeviceMemoryProperties mp{};
            api.GetPhysicalDeviceMemoryProperties(api.physical_device, &mp);
            uint32_t type = UINT32_MAX;
            for (uint32_t i = 0; i < mp.memoryTypeCount; ++i)
                if ((req.memoryTypeBits & (1u << i)) && (mp.memoryTypes[i].propertyFlags & VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT)) { type = i; break; }
            if (type == UINT32_MAX) throw std::runtime_error("vkpbrt::vk::SharedPlane: no device-local memory type for an exportable buffer");
            VkMemoryDedicatedAllocateInfo ded{};
            ded.sType = VK_STRUCTURE_TYPE_MEMORY_DEDICATED_ALLOCATE_INFO;
            ded.buffer = buffer;
            VkExportMemoryAllocateInfo exp{};
            exp.sType = VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO;
            exp.pNext = dedicated ? &ded : nullptr;
            exp.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
            VkMemoryAllocateInfo ai{};
            ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
            ai.pNext = &exp;
            ai.allocationSize = req.size;
            ai.memoryTypeIndex = type;
            vk_check(api.AllocateMemory(api.device, &ai, nullptr, &memory), "vkAllocateMemory (exportable)");
            allocation_size = req.size;
            vk_check(api.BindBufferMemory(api.device, buffer, memory, 0), "vkBindBufferMemory");

            VkMemoryGetFdInfoKHR gi{};
            gi.sType = VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR;
            gi.memory = memory;
            gi.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
            int fd = -1;
            vk_check(api.GetMemoryFdKHR(api.device, &gi, &fd), "vkGetMemoryFdKHR");
            void* ptr = nullptr;
            const int rc = vkpbrt_import_external_memory_fd_ex(ctx.handle, fd, allocation_size, 0, size_bytes, dedicated ? 1 : 0, &_imported, &ptr);
            if (rc != VKPBRT_OK) { ::close(fd); check(rc); }            // a successful import owns the fd, a failed one does not
            vkpbrt_image_t h = nullptr;
            check(vkpbrt_image_wrap(ctx.handle, format, width, height, 1, ptr, &h));
            image = DescriptorImage::create(h, true);
        } catch (...) {
            release();
            throw;
        }
    }
    ~SharedPlane() { release(); }
    SharedPlane(const SharedPlane&) = delete;

    // vkCmdCopyImageToBuffer of the whole of `src` (an image of this plane's extent and a size-compatible format, in
    // `layout`, last written in `producer_stage` with `producer_access`) into the plane; the image is returned to `layout`.
    void cmd_copy_from_image(VkCommandBuffer cb, VkImage src, VkImageLayout layout = VK_IMAGE_LAYOUT_GENERAL,
                             VkPipelineStageFlags producer_stage = VK_PIPELINE_STAGE_ALL_COMMANDS_BIT, VkAccessFlags producer_access = VK_ACCESS_SHADER_WRITE_BIT) const
    {
        image_barrier(cb, src, layout, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, producer_stage, VK_PIPELINE_STAGE_TRANSFER_BIT, producer_access, VK_ACCESS_TRANSFER_READ_BIT);
        const VkBufferImageCopy region = whole();
        _api.CmdCopyImageToBuffer(cb, src, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, buffer, 1, &region);
        image_barrier(cb, src, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, layout, VK_PIPELINE_STAGE_TRANSFER_BIT, producer_stage, VK_ACCESS_TRANSFER_READ_BIT, producer_access);
        // the timeline-semaphore signal of the submission makes the transfer writes available to the importer
    }
    // vkCmdCopyBufferToImage of the plane into the whole of `dst` (e.g. the image that is blitted to the window,
    // VulkanPBRT.cpp:529); `dst` goes from `layout` to `final_layout`.
    void cmd_copy_to_image(VkCommandBuffer cb, VkImage dst, VkImageLayout layout = VK_IMAGE_LAYOUT_UNDEFINED, VkImageLayout final_layout = VK_IMAGE_LAYOUT_GENERAL,
                           VkPipelineStageFlags consumer_stage = VK_PIPELINE_STAGE_ALL_COMMANDS_BIT, VkAccessFlags consumer_access = VK_ACCESS_SHADER_READ_BIT) const
    {
        image_barrier(cb, dst, layout, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, consumer_stage, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, VK_ACCESS_TRANSFER_WRITE_BIT);
        const VkBufferImageCopy region = whole();
        _api.CmdCopyBufferToImage(cb, buffer, dst, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region);
        image_barrier(cb, dst, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, final_layout, VK_PIPELINE_STAGE_TRANSFER_BIT, consumer_stage, VK_ACCESS_TRANSFER_WRITE_BIT, consumer_access);
    }

    VkBuffer buffer = VK_NULL_HANDLE;
    VkDeviceMemory memory = VK_NULL_HANDLE;
    VkDeviceSize size_bytes = 0, allocation_size = 0;
    bool dedicated = false;
    ref_ptr<DescriptorImage> image;      // the CUDA side's view
private:
    VkBufferImageCopy whole() const
    {
        VkBufferImageCopy r{};
        r.bufferOffset = 0;
        r.bufferRowLength = 0;           // tightly packed: what the kernels expect (vkpbrt_image_info::row_pitch)
        r.bufferImageHeight = 0;
        r.imageSubresource.aspectMask = VK_IMAGE_ASPECT_COLOR_BIT;
        r.imageSubresource.mipLevel = 0;
        r.imageSubresource.baseArrayLayer = 0;
        r.imageSubresource.layerCount = 1;
        r.imageExtent.width = width; r.imageExtent.height = height; r.imageExtent.depth = 1;
        return r;
    }
    void image_barrier(VkCommandBuffer cb, VkImage img, VkImageLayout from, VkImageLayout to, VkPipelineStageFlags src_stage, VkPipelineStageFlags dst_stage,
                       VkAccessFlags src_access, VkAccessFlags dst_access) const
    {
        VkImageMemoryBarrier b{};
        b.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
        b.srcAccessMask = src_access;
        b.dstAccessMask = dst_access;
        b.oldLayout = from;
        b.newLayout = to;
        b.srcQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
        b.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
        b.image = img;
        b.subresourceRange.aspectMask = VK_IMAGE_ASPECT_COLOR_BIT;
        b.subresourceRange.levelCount = 1;
        b.subresourceRange.layerCount = 1;
        _api.CmdPipelineBarrier(cb, src_stage, dst_stage, 0, 0, nullptr, 0, nullptr, 1, &b);
    }
    void release()
    {
        image.reset();                                              // the wrapping handle first, then the mapping under it
        if (_imported) { vkpbrt_external_memory_destroy(_imported); _imported = nullptr; }
        if (buffer) { _api.DestroyBuffer(_api.device, buffer, nullptr); buffer = VK_NULL_HANDLE; }
        if (memory) { _api.FreeMemory(_api.device, memory, nullptr); memory = VK_NULL_HANDLE; }
    }
    Api _api;
    vkpbrt_external_memory_t _imported = nullptr;
public:
    const uint32_t format, width, height;
};

// One timeline semaphore, visible to Vulkan queues (`semaphore`) and to the context's CUDA stream.
class SharedTimeline : public Inherit<SharedTimeline> {
public:
    SharedTimeline(const Api& api, Context& ctx, uint64_t initial_value = 0) : _api(api)
    {
        VkPhysicalDeviceExternalSemaphoreInfo qi{};
        qi.sType = VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_EXTERNAL_SEMAPHORE_INFO;
        VkSemaphoreTypeCreateInfo qt{};
        qt.sType = VK_STRUCTURE_TYPE_SEMAPHORE_TYPE_CREATE_INFO;
        qt.semaphoreType = VK_SEMAPHORE_TYPE_TIMELINE;
        qi.pNext = &qt;
        qi.handleType = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkExternalSemaphoreProperties qp{};
        qp.sType = VK_STRUCTURE_TYPE_EXTERNAL_SEMAPHORE_PROPERTIES;
        api.GetPhysicalDeviceExternalSemaphoreProperties(api.physical_device, &qi, &qp);
        if (!(qp.externalSemaphoreFeatures & VK_EXTERNAL_SEMAPHORE_FEATURE_EXPORTABLE_BIT))
            throw std::runtime_error("vkpbrt::vk::SharedTimeline: the Vulkan device cannot export timeline semaphores as opaque file descriptors");

        VkExportSemaphoreCreateInfo exp{};
        exp.sType = VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO;
        exp.handleTypes = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkSemaphoreTypeCreateInfo type{};
        type.sType = VK_STRUCTURE_TYPE_SEMAPHORE_TYPE_CREATE_INFO;
        type.pNext = &exp;
        type.semaphoreType = VK_SEMAPHORE_TYPE_TIMELINE;
        type.initialValue = initial_value;
        VkSemaphoreCreateInfo ci{};
        ci.sType = VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO;
        ci.pNext = &type;
        vk_check(api.CreateSemaphore(api.device, &ci, nullptr, &semaphore), "vkCreateSemaphore (exportable timeline)");
        VkSemaphoreGetFdInfoKHR gi{};
        gi.sType = VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR;
        gi.semaphore = semaphore;
        gi.handleType = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
        int fd = -1;
        VkResult r = api.GetSemaphoreFdKHR(api.device, &gi, &fd);
        int rc = VKPBRT_OK;
        if (r == VK_SUCCESS) {
            rc = vkpbrt_import_external_semaphore_fd(ctx.handle, fd, /*timeline=*/1, &_imported);
            if (rc != VKPBRT_OK) ::close(fd);
        }
        if (r != VK_SUCCESS || rc != VKPBRT_OK) {
            api.DestroySemaphore(api.device, semaphore, nullptr);
            semaphore = VK_NULL_HANDLE;
            vk_check(r, "vkGetSemaphoreFdKHR");
            check(rc);
        }
    }
    ~SharedTimeline()
    {
        if (_imported) vkpbrt_external_semaphore_destroy(_imported);
        if (semaphore) _api.DestroySemaphore(_api.device, semaphore, nullptr);
    }
    SharedTimeline(const SharedTimeline&) = delete;
    // CUDA side, on the context's stream
    void cuda_wait(uint64_t value) { check(vkpbrt_external_semaphore_wait(_imported, value)); }
    void cuda_signal(uint64_t value) { check(vkpbrt_external_semaphore_signal(_imported, value)); }
    // host side
    uint64_t value() const { uint64_t v = 0; vk_check(_api.GetSemaphoreCounterValue(_api.device, semaphore, &v), "vkGetSemaphoreCounterValue"); return v; }
    void host_signal(uint64_t value)
    {
        VkSemaphoreSignalInfo si{};
        si.sType = VK_STRUCTURE_TYPE_SEMAPHORE_SIGNAL_INFO;
        si.semaphore = semaphore;
        si.value = value;
        vk_check(_api.SignalSemaphore(_api.device, &si), "vkSignalSemaphore");
    }
    bool host_wait(uint64_t value, uint64_t timeout_ns = UINT64_MAX)
    {
        VkSemaphoreWaitInfo wi{};
        wi.sType = VK_STRUCTURE_TYPE_SEMAPHORE_WAIT_INFO;
        wi.semaphoreCount = 1;
        wi.pSemaphores = &semaphore;
        wi.pValues = &value;
        const VkResult r = _api.WaitSemaphores(_api.device, &wi, timeout_ns);
        if (r == VK_TIMEOUT) return false;
        vk_check(r, "vkWaitSemaphores");
        return true;
    }
    VkSemaphore semaphore = VK_NULL_HANDLE;
private:
    Api _api;
    vkpbrt_external_semaphore_t _imported = nullptr;
};

// Timeline waits / signals of one vkQueueSubmit.  apply() points the VkSubmitInfo at this object's members: keep it
// alive (and do not move it) until vkQueueSubmit has returned.
struct TimelineSubmit {
    VkSemaphore wait_semaphore = VK_NULL_HANDLE, signal_semaphore = VK_NULL_HANDLE;
    uint64_t wait_value = 0, signal_value = 0;
    VkPipelineStageFlags wait_stage = VK_PIPELINE_STAGE_ALL_COMMANDS_BIT;
    VkTimelineSemaphoreSubmitInfo timeline{};
    void apply(VkSubmitInfo& submit)
    {
        timeline = VkTimelineSemaphoreSubmitInfo{};
        timeline.sType = VK_STRUCTURE_TYPE_TIMELINE_SEMAPHORE_SUBMIT_INFO;
        timeline.pNext = submit.pNext;
        if (wait_semaphore) {
            timeline.waitSemaphoreValueCount = 1; timeline.pWaitSemaphoreValues = &wait_value;
            submit.waitSemaphoreCount = 1; submit.pWaitSemaphores = &wait_semaphore; submit.pWaitDstStageMask = &wait_stage;
        }
        if (signal_semaphore) {
            timeline.signalSemaphoreValueCount = 1; timeline.pSignalSemaphoreValues = &signal_value;
            submit.signalSemaphoreCount = 1; submit.pSignalSemaphores = &signal_semaphore;
        }
        submit.pNext = &timeline;
    }
};

// Everything the path shares with a Vulkan renderer, and the order of one frame:
//
//   Vulkan queue   [render G-buffer + 1-spp illumination] -> copy into the shared planes            signal produced = f + 1
//   CUDA stream    wait produced >= f + 1 -> accumulate, denoise, TAA -> copy final into final_plane  signal consumed = f + 1
//   Vulkan queue   wait consumed >= f + 1 -> final_plane into the presented image
//   and the producer's copies of frame f + 1 wait consumed >= f + 1 as well: by then the kernels have read frame f's planes.
//
// g_buffer / illumination_buffer are what Accumulator::create takes (VulkanPBRT.cpp:426); the history planes of the
// path stay private to the CUDA side.
class SharedFrame : public Inherit<SharedFrame> {
public:
    SharedFrame(const Api& api, Context& ctx, uint32_t width, uint32_t height) : width(width), height(height)
    {
        depth = SharedPlane::create(api, ctx, VKPBRT_FORMAT_R32_SFLOAT, width, height);                       // GBuffer.cpp:60
        normal = SharedPlane::create(api, ctx, VKPBRT_FORMAT_R32G32_SFLOAT, width, height);                   // GBuffer.cpp:77
        albedo = SharedPlane::create(api, ctx, VKPBRT_FORMAT_R8G8B8A8_UNORM, width, height);                  // GBuffer.cpp:111
        illumination = SharedPlane::create(api, ctx, VKPBRT_FORMAT_R32G32B32A32_SFLOAT, width, height);       // IlluminationBuffer.cpp:260-282
        final_plane = SharedPlane::create(api, ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, width, height);             // BMFR.cpp:81, Taa.cpp:42
        g_buffer = GBuffer::create(ctx, depth->image, normal->image, ref_ptr<DescriptorImage>(), albedo->image);   // material is not read by the path
        illumination_buffer = IlluminationBufferDemodulatedFloat::create(ctx, std::vector<ref_ptr<DescriptorImage>>{illumination->image});
        produced = SharedTimeline::create(api, ctx, 0);
        consumed = SharedTimeline::create(api, ctx, 0);
    }
    // command-list entries (recorded once, replayed per frame; the frame index is read from the push constants at replay,
    // like everything else the modules take from them).  wait: before the accumulator; signal: after the last module,
    // `final_image` being what the chain ends in (get_final_descriptor_image() of the denoiser, blender or TAA).
    void add_wait_to_commands(ref_ptr<Commands> commands, ref_ptr<PushConstants> push_constants)
    {
        commands->addChild([self = this->shared_from_this(), push_constants](Commands&) { self->produced->cuda_wait((uint64_t)push_constants->value().frame_number + 1); });
    }
    void add_signal_to_commands(ref_ptr<Commands> commands, ref_ptr<PushConstants> push_constants, ref_ptr<DescriptorImage> final_image)
    {
        commands->addChild([self = this->shared_from_this(), push_constants, final_image](Commands&) {
            check(vkpbrt_image_copy_record(final_image->handle, self->final_plane->image->handle));
            self->consumed->cuda_signal((uint64_t)push_constants->value().frame_number + 1);
        });
    }
    // the renderer's submission of frame `frame`: waits until the kernels are done with the previous frame's planes
    TimelineSubmit producer_submit(uint64_t frame) const
    {
        TimelineSubmit s;
        if (frame > 0) { s.wait_semaphore = consumed->semaphore; s.wait_value = frame; s.wait_stage = VK_PIPELINE_STAGE_TRANSFER_BIT; }
        s.signal_semaphore = produced->semaphore; s.signal_value = frame + 1;
        return s;
    }
    // the submission that shows frame `frame`
    TimelineSubmit presenter_submit(uint64_t frame) const
    {
        TimelineSubmit s;
        s.wait_semaphore = consumed->semaphore; s.wait_value = frame + 1; s.wait_stage = VK_PIPELINE_STAGE_TRANSFER_BIT;
        return s;
    }
    const uint32_t width, height;
    ref_ptr<SharedPlane> depth, normal, albedo, illumination, final_plane;
    ref_ptr<GBuffer> g_buffer;
    ref_ptr<IlluminationBuffer> illumination_buffer;
    ref_ptr<SharedTimeline> produced, consumed;
};

}  // namespace vk
}  // namespace vkpbrt

#endif  // VKPBRT_HAVE_VULKAN
