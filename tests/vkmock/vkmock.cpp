// vkmock.cpp -- TEST-ONLY mock of the Vulkan entry points the interop layer (include/vkpbrt/vk_interop.hpp) and its
// example host (examples/cpp_vulkan_interop.cpp) use.  Neither machine of this project has a Vulkan loader, driver or
// lavapipe, so this plays "the Vulkan implementation" for tests/test_vk_interop.py, with the CUDA side on the test
// emulator (tests/hostsim), and for tools/vk_oracle (tests/test_vk_oracle.py).  It is not a Vulkan implementation: one
// queue that executes a submission synchronously inside vkQueueSubmit, and "compute" that does not run SPIR-V (below).
//
// What it does model, because the interop depends on it:
//   * device memory is a memfd; vkGetMemoryFdKHR hands out a dup() of it (the importer maps the same pages), and only
//     for allocations made with VkExportMemoryAllocateInfo on a device created with VK_KHR_external_memory_fd;
//   * VKMOCK_DEDICATED_ONLY=1: exportable buffers must be dedicated allocations (some drivers' rule);
//   * semaphores are a memfd page {magic, payload, timeline}; timeline waits of a submission block (up to 20 s) until the
//     payload arrives -- a wrong handshake order fails instead of passing silently;
//   * TILING_OPTIMAL images have a padded row pitch, so only vkCmdCopy{Image,Buffer}To{Buffer,Image} can repack them;
//   * images track their layout: a copy whose declared layout is not the current one aborts (the barriers of
//     SharedPlane::cmd_copy_* are checked that way);
//   * entry points of an extension / feature that was not enabled at vkCreateDevice are not returned by
//     vkGetDeviceProcAddr.
//   * compute: a "shader module" is not SPIR-V but a text blob "VKMOCK-SHADER:<name>" naming one of the reference's
//     shaders; vkCmdDispatch gathers the bound descriptors, specialisation and push constants and runs that shader's
//     SOURCE TEXT as compiled by oracle/glsl_shim (oracle/_ref/libref.so, dlopen'ed from $VKMOCK_LIBREF) -- so a raw-Vulkan
//     host (tools/vk_oracle) is checked for the reference's binding numbers, descriptor types, formats, specialisation
//     constants, push-constant blocks and dispatch sizes by the result it produces.  Real SPIR-V is refused.
// The only exported symbol is vkGetInstanceProcAddr, as with a real loader used through vk::open_loader().
#include <vulkan/vulkan_core.h>

#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#define MOCK_FAIL(...) do { fprintf(stderr, "vkmock validation: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

static const uint64_t kSemaphoreMagic = 0x564b4d4f434b5345ull;   // "VKMOCKSE", shared with tests/hostsim/hostsim.cpp
struct SemaphorePage { uint64_t magic; uint64_t payload; uint32_t timeline; };

struct VkInstance_T { int dummy; };
struct VkPhysicalDevice_T { int dummy; };
struct VkQueue_T { int dummy; };
struct VkDevice_T { bool ext_memory_fd = false, ext_semaphore_fd = false, timeline = false; VkQueue_T queue; };
struct VkDeviceMemory_T { int fd; size_t size; char* map; uint32_t type; bool exportable; bool dedicated; };
struct VkBuffer_T { VkDeviceSize size; VkBufferUsageFlags usage; bool external; VkDeviceMemory_T* mem; VkDeviceSize offset; };
struct VkImage_T {
    VkFormat format; uint32_t w, h, texel; size_t pitch; VkImageLayout layout; VkImageUsageFlags usage; VkDeviceMemory_T* mem; VkDeviceSize offset;
    uint32_t layers = 1;
    size_t layer_stride() const { return pitch * h; }
    char* row(uint32_t layer, uint32_t y) const { return mem->map + offset + layer * layer_stride() + y * pitch; }
};
struct VkImageView_T { VkImage_T* image; VkImageViewType type; VkFormat format; uint32_t base_layer, layer_count; };
struct VkSampler_T { VkFilter mag, min; VkSamplerAddressMode u, v; VkBool32 unnormalized; };
struct VkShaderModule_T { std::string name; };
struct VkDescriptorSetLayout_T { std::map<uint32_t, VkDescriptorType> bindings; };
struct VkPipelineLayout_T { std::vector<VkDescriptorSetLayout_T*> sets; uint32_t push_size; };
struct VkPipeline_T { std::string shader; std::map<uint32_t, int32_t> spec; VkPipelineLayout_T* layout; };
struct VkDescriptorSet_T;
struct VkDescriptorPool_T { std::vector<VkDescriptorSet_T*> sets; };      // sets die with their pool
struct Descriptor { VkDescriptorType type; VkImageView_T* view; VkSampler_T* sampler; VkImageLayout layout; };
struct VkDescriptorSet_T { VkDescriptorSetLayout_T* layout; std::map<uint32_t, Descriptor> bound; };
struct VkSemaphore_T { bool timeline; bool exportable; int fd; SemaphorePage* page; };
struct VkCommandBuffer_T;
struct VkCommandPool_T { std::vector<VkCommandBuffer_T*> buffers; };      // command buffers that were not freed die with their pool
struct VkCommandBuffer_T {
    std::vector<std::function<void()>> cmds; bool recording = false;
    VkPipeline_T* pipeline = nullptr; VkDescriptorSet_T* set = nullptr; std::vector<unsigned char> push;     // state while recording
};

static VkInstance_T g_instance;
static VkPhysicalDevice_T g_physical_device;

static int make_memfd(const char* name, size_t size)
{
    int fd = (int)syscall(SYS_memfd_create, name, 0);
    if (fd < 0 || ftruncate(fd, (off_t)size) != 0) MOCK_FAIL("memfd_create/ftruncate failed");
    return fd;
}

static uint32_t texel_size(VkFormat f)
{
    switch (f) {
    case VK_FORMAT_R8_UNORM: return 1;
    case VK_FORMAT_R16_SFLOAT: return 2;
    case VK_FORMAT_R32_SFLOAT: case VK_FORMAT_R8G8B8A8_UNORM: case VK_FORMAT_B8G8R8A8_UNORM: case VK_FORMAT_R16G16_SFLOAT: return 4;
    case VK_FORMAT_R32G32_SFLOAT: case VK_FORMAT_R16G16B16A16_SFLOAT: return 8;
    case VK_FORMAT_R32G32B32A32_SFLOAT: return 16;
    default: return 0;
    }
}

template <typename T>
static const T* find_in_chain(const void* pnext, VkStructureType type)
{
    for (auto* s = (const VkBaseInStructure*)pnext; s; s = s->pNext)
        if (s->sType == type) return (const T*)s;
    return nullptr;
}

// ---- instance level ---------------------------------------------------------------------------------------------
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateInstance(const VkInstanceCreateInfo* ci, const VkAllocationCallbacks*, VkInstance* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO) MOCK_FAIL("vkCreateInstance: bad create info");
    if (!ci->pApplicationInfo || ci->pApplicationInfo->apiVersion < VK_API_VERSION_1_2) MOCK_FAIL("vkCreateInstance: the interop needs apiVersion >= 1.2");
    *out = &g_instance;
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyInstance(VkInstance, const VkAllocationCallbacks*) {}
static VKAPI_ATTR VkResult VKAPI_CALL mock_EnumeratePhysicalDevices(VkInstance, uint32_t* count, VkPhysicalDevice* devices)
{
    if (!devices) { *count = 1; return VK_SUCCESS; }
    if (*count < 1) return VK_INCOMPLETE;
    devices[0] = &g_physical_device; *count = 1;
    return VK_SUCCESS;
}
static void fill_properties(VkPhysicalDeviceProperties* p)
{
    memset(p, 0, sizeof(*p));
    p->apiVersion = VK_API_VERSION_1_2;
    p->deviceType = VK_PHYSICAL_DEVICE_TYPE_DISCRETE_GPU;
    strcpy(p->deviceName, "vkmock (test-only mock device)");
}
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceProperties(VkPhysicalDevice, VkPhysicalDeviceProperties* p) { fill_properties(p); }
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceProperties2(VkPhysicalDevice, VkPhysicalDeviceProperties2* p)
{
    fill_properties(&p->properties);
    for (auto* s = (VkBaseOutStructure*)p->pNext; s; s = s->pNext)
        if (s->sType == VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_ID_PROPERTIES) {
            auto* id = (VkPhysicalDeviceIDProperties*)s;
            const char* uuid = getenv("VKMOCK_DEVICE_UUID");                 // 16 characters; default = the emulator's "CUDA device"
            memcpy(id->deviceUUID, (uuid && strlen(uuid) == 16) ? uuid : "HOSTSIM-DEVICE-0", 16);
            memset(id->driverUUID, 0, 16);
            id->deviceLUIDValid = VK_FALSE;
        }
}
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceFeatures2(VkPhysicalDevice, VkPhysicalDeviceFeatures2* f)
{
    memset(&f->features, 0, sizeof(f->features));
    for (auto* s = (VkBaseOutStructure*)f->pNext; s; s = s->pNext) {
        if (s->sType == VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_VULKAN_1_2_FEATURES) ((VkPhysicalDeviceVulkan12Features*)s)->timelineSemaphore = VK_TRUE;
        if (s->sType == VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_TIMELINE_SEMAPHORE_FEATURES) ((VkPhysicalDeviceTimelineSemaphoreFeatures*)s)->timelineSemaphore = VK_TRUE;
    }
}
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceQueueFamilyProperties(VkPhysicalDevice, uint32_t* count, VkQueueFamilyProperties* props)
{
    if (!props) { *count = 1; return; }
    if (*count < 1) return;
    memset(props, 0, sizeof(*props));
    props[0].queueFlags = VK_QUEUE_GRAPHICS_BIT | VK_QUEUE_COMPUTE_BIT | VK_QUEUE_TRANSFER_BIT;
    props[0].queueCount = 1;
    *count = 1;
}
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceMemoryProperties(VkPhysicalDevice, VkPhysicalDeviceMemoryProperties* p)
{
    memset(p, 0, sizeof(*p));
    p->memoryTypeCount = 2;
    p->memoryTypes[0].propertyFlags = VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT | VK_MEMORY_PROPERTY_HOST_COHERENT_BIT;   // listed first on purpose:
    p->memoryTypes[0].heapIndex = 1;                                                                                  // the interop must pick by flags
    p->memoryTypes[1].propertyFlags = VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT;
    p->memoryTypes[1].heapIndex = 0;
    p->memoryHeapCount = 2;
    p->memoryHeaps[0].size = 1ull << 34; p->memoryHeaps[0].flags = VK_MEMORY_HEAP_DEVICE_LOCAL_BIT;
    p->memoryHeaps[1].size = 1ull << 34;
}
static bool dedicated_only() { const char* e = getenv("VKMOCK_DEDICATED_ONLY"); return e && *e == '1'; }
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceExternalBufferProperties(VkPhysicalDevice, const VkPhysicalDeviceExternalBufferInfo* info, VkExternalBufferProperties* props)
{
    memset(&props->externalMemoryProperties, 0, sizeof(props->externalMemoryProperties));
    if (info->handleType != VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT) return;
    props->externalMemoryProperties.externalMemoryFeatures = VK_EXTERNAL_MEMORY_FEATURE_EXPORTABLE_BIT | VK_EXTERNAL_MEMORY_FEATURE_IMPORTABLE_BIT |
                                                             (dedicated_only() ? VK_EXTERNAL_MEMORY_FEATURE_DEDICATED_ONLY_BIT : 0);
    props->externalMemoryProperties.exportFromImportedHandleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
    props->externalMemoryProperties.compatibleHandleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
}
static VKAPI_ATTR void VKAPI_CALL mock_GetPhysicalDeviceExternalSemaphoreProperties(VkPhysicalDevice, const VkPhysicalDeviceExternalSemaphoreInfo* info, VkExternalSemaphoreProperties* props)
{
    props->exportFromImportedHandleTypes = props->compatibleHandleTypes = 0;
    props->externalSemaphoreFeatures = 0;
    if (info->handleType != VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT) return;
    props->exportFromImportedHandleTypes = props->compatibleHandleTypes = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
    props->externalSemaphoreFeatures = VK_EXTERNAL_SEMAPHORE_FEATURE_EXPORTABLE_BIT | VK_EXTERNAL_SEMAPHORE_FEATURE_IMPORTABLE_BIT;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_EnumerateDeviceExtensionProperties(VkPhysicalDevice, const char*, uint32_t* count, VkExtensionProperties* props)
{
    static const char* names[] = {VK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME, VK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME};
    if (!props) { *count = 2; return VK_SUCCESS; }
    const uint32_t n = *count < 2 ? *count : 2;
    for (uint32_t i = 0; i < n; ++i) { memset(&props[i], 0, sizeof(props[i])); strcpy(props[i].extensionName, names[i]); props[i].specVersion = 1; }
    *count = n;
    return n < 2 ? VK_INCOMPLETE : VK_SUCCESS;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateDevice(VkPhysicalDevice, const VkDeviceCreateInfo* ci, const VkAllocationCallbacks*, VkDevice* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO || ci->queueCreateInfoCount < 1) MOCK_FAIL("vkCreateDevice: bad create info");
    auto* d = new VkDevice_T();
    for (uint32_t i = 0; i < ci->enabledExtensionCount; ++i) {
        const std::string e = ci->ppEnabledExtensionNames[i];
        if (e == VK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME) d->ext_memory_fd = true;
        else if (e == VK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME) d->ext_semaphore_fd = true;
        else { delete d; return VK_ERROR_EXTENSION_NOT_PRESENT; }
    }
    if (auto* f12 = find_in_chain<VkPhysicalDeviceVulkan12Features>(ci->pNext, VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_VULKAN_1_2_FEATURES)) d->timeline |= f12->timelineSemaphore == VK_TRUE;
    if (auto* ft = find_in_chain<VkPhysicalDeviceTimelineSemaphoreFeatures>(ci->pNext, VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_TIMELINE_SEMAPHORE_FEATURES)) d->timeline |= ft->timelineSemaphore == VK_TRUE;
    *out = d;
    return VK_SUCCESS;
}

// ---- device level -----------------------------------------------------------------------------------------------
static VKAPI_ATTR void VKAPI_CALL mock_DestroyDevice(VkDevice d, const VkAllocationCallbacks*) { delete d; }
static VKAPI_ATTR void VKAPI_CALL mock_GetDeviceQueue(VkDevice d, uint32_t family, uint32_t index, VkQueue* q)
{
    if (family != 0 || index != 0) MOCK_FAIL("vkGetDeviceQueue: only queue (0, 0) exists");
    *q = &d->queue;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_DeviceWaitIdle(VkDevice) { return VK_SUCCESS; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_QueueWaitIdle(VkQueue) { return VK_SUCCESS; }

static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateBuffer(VkDevice, const VkBufferCreateInfo* ci, const VkAllocationCallbacks*, VkBuffer* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO || ci->size == 0) MOCK_FAIL("vkCreateBuffer: bad create info");
    auto* ext = find_in_chain<VkExternalMemoryBufferCreateInfo>(ci->pNext, VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO);
    *out = new VkBuffer_T{ci->size, ci->usage, ext && (ext->handleTypes & VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT), nullptr, 0};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyBuffer(VkDevice, VkBuffer b, const VkAllocationCallbacks*) { delete b; }
static VKAPI_ATTR void VKAPI_CALL mock_GetBufferMemoryRequirements(VkDevice, VkBuffer b, VkMemoryRequirements* r)
{
    r->size = (b->size + 4095) / 4096 * 4096;          // larger than the buffer, as real allocations are
    r->alignment = 256;
    r->memoryTypeBits = 0x3;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateImage(VkDevice, const VkImageCreateInfo* ci, const VkAllocationCallbacks*, VkImage* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO || ci->imageType != VK_IMAGE_TYPE_2D || ci->extent.depth != 1 || ci->mipLevels != 1 || ci->arrayLayers < 1)
        MOCK_FAIL("vkCreateImage: only single-level 2-D (array) images");
    const uint32_t texel = texel_size(ci->format);
    if (!texel) MOCK_FAIL("vkCreateImage: format %d is not modelled", (int)ci->format);
    size_t pitch = (size_t)ci->extent.width * texel;
    if (ci->tiling == VK_IMAGE_TILING_OPTIMAL) pitch = (pitch + 63) / 64 * 64 + 64;     // "opaque" layout: rows are not where a linear view expects them
    *out = new VkImage_T{ci->format, ci->extent.width, ci->extent.height, texel, pitch, ci->initialLayout, ci->usage, nullptr, 0};
    (*out)->layers = ci->arrayLayers;
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyImage(VkDevice, VkImage i, const VkAllocationCallbacks*) { delete i; }
static VKAPI_ATTR void VKAPI_CALL mock_GetImageMemoryRequirements(VkDevice, VkImage i, VkMemoryRequirements* r)
{
    r->size = (i->layer_stride() * i->layers + 4095) / 4096 * 4096;
    r->alignment = 1024;
    r->memoryTypeBits = 0x2;                            // device-local only: images cannot be mapped
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_AllocateMemory(VkDevice, const VkMemoryAllocateInfo* ai, const VkAllocationCallbacks*, VkDeviceMemory* out)
{
    if (!ai || ai->sType != VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO || ai->allocationSize == 0 || ai->memoryTypeIndex > 1) MOCK_FAIL("vkAllocateMemory: bad allocate info");
    auto* exp = find_in_chain<VkExportMemoryAllocateInfo>(ai->pNext, VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO);
    auto* ded = find_in_chain<VkMemoryDedicatedAllocateInfo>(ai->pNext, VK_STRUCTURE_TYPE_MEMORY_DEDICATED_ALLOCATE_INFO);
    const bool exportable = exp && (exp->handleTypes & VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT);
    const bool dedicated = ded && (ded->buffer || ded->image);
    if (exportable && dedicated_only() && !dedicated) MOCK_FAIL("vkAllocateMemory: this device exports buffers only from dedicated allocations");
    const int fd = make_memfd("vkmock-memory", ai->allocationSize);
    char* map = (char*)mmap(nullptr, ai->allocationSize, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (map == MAP_FAILED) MOCK_FAIL("mmap failed");
    memset(map, 0xCD, ai->allocationSize);              // fresh device memory is not zero
    *out = new VkDeviceMemory_T{fd, (size_t)ai->allocationSize, map, ai->memoryTypeIndex, exportable, dedicated};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_FreeMemory(VkDevice, VkDeviceMemory m, const VkAllocationCallbacks*)
{
    if (!m) return;
    munmap(m->map, m->size);
    close(m->fd);
    delete m;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_BindBufferMemory(VkDevice, VkBuffer b, VkDeviceMemory m, VkDeviceSize offset)
{
    if (offset % 256 || offset + b->size > m->size) MOCK_FAIL("vkBindBufferMemory: bad offset / size");
    if (b->external && !m->exportable) MOCK_FAIL("vkBindBufferMemory: an external buffer must be bound to memory allocated with VkExportMemoryAllocateInfo");
    b->mem = m; b->offset = offset;
    return VK_SUCCESS;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_BindImageMemory(VkDevice, VkImage i, VkDeviceMemory m, VkDeviceSize offset)
{
    if (offset % 1024 || offset + i->layer_stride() * i->layers > m->size) MOCK_FAIL("vkBindImageMemory: bad offset / size");
    if (m->type != 1) MOCK_FAIL("vkBindImageMemory: images live in device-local memory");
    i->mem = m; i->offset = offset;
    return VK_SUCCESS;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_MapMemory(VkDevice, VkDeviceMemory m, VkDeviceSize offset, VkDeviceSize, VkMemoryMapFlags, void** out)
{
    if (m->type != 0) MOCK_FAIL("vkMapMemory: memory type %u is not HOST_VISIBLE", m->type);
    *out = m->map + offset;
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_UnmapMemory(VkDevice, VkDeviceMemory) {}
static VKAPI_ATTR VkResult VKAPI_CALL mock_GetMemoryFdKHR(VkDevice, const VkMemoryGetFdInfoKHR* gi, int* fd)
{
    if (!gi || gi->sType != VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR || gi->handleType != VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT) MOCK_FAIL("vkGetMemoryFdKHR: bad info");
    if (!gi->memory->exportable) MOCK_FAIL("vkGetMemoryFdKHR: the memory was not allocated with VkExportMemoryAllocateInfo");
    *fd = dup(gi->memory->fd);                          // every call returns a new fd owned by the caller
    return *fd >= 0 ? VK_SUCCESS : VK_ERROR_TOO_MANY_OBJECTS;
}

static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateSemaphore(VkDevice d, const VkSemaphoreCreateInfo* ci, const VkAllocationCallbacks*, VkSemaphore* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO) MOCK_FAIL("vkCreateSemaphore: bad create info");
    auto* type = find_in_chain<VkSemaphoreTypeCreateInfo>(ci->pNext, VK_STRUCTURE_TYPE_SEMAPHORE_TYPE_CREATE_INFO);
    auto* exp = find_in_chain<VkExportSemaphoreCreateInfo>(ci->pNext, VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO);
    const bool timeline = type && type->semaphoreType == VK_SEMAPHORE_TYPE_TIMELINE;
    if (timeline && !d->timeline) MOCK_FAIL("vkCreateSemaphore: the timelineSemaphore feature was not enabled at vkCreateDevice");
    const int fd = make_memfd("vkmock-semaphore", 4096);
    auto* page = (SemaphorePage*)mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (page == MAP_FAILED) MOCK_FAIL("mmap failed");
    page->magic = kSemaphoreMagic;
    page->payload = timeline ? type->initialValue : 0;
    page->timeline = timeline ? 1 : 0;
    *out = new VkSemaphore_T{timeline, exp && (exp->handleTypes & VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT), fd, page};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroySemaphore(VkDevice, VkSemaphore s, const VkAllocationCallbacks*)
{
    if (!s) return;
    munmap(s->page, 4096);
    close(s->fd);
    delete s;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_GetSemaphoreFdKHR(VkDevice, const VkSemaphoreGetFdInfoKHR* gi, int* fd)
{
    if (!gi || gi->sType != VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR || gi->handleType != VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT) MOCK_FAIL("vkGetSemaphoreFdKHR: bad info");
    if (!gi->semaphore->exportable) MOCK_FAIL("vkGetSemaphoreFdKHR: the semaphore was not created with VkExportSemaphoreCreateInfo");
    *fd = dup(gi->semaphore->fd);
    return *fd >= 0 ? VK_SUCCESS : VK_ERROR_TOO_MANY_OBJECTS;
}
static bool wait_payload(VkSemaphore_T* s, uint64_t want, uint64_t timeout_ns)
{
    timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    while (__atomic_load_n(&s->page->payload, __ATOMIC_ACQUIRE) < want) {
        timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
        const uint64_t el = (uint64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ull + (uint64_t)(t1.tv_nsec - t0.tv_nsec + 1000000000l) - 1000000000ull;
        if (el >= timeout_ns) return false;
        sched_yield();
    }
    return true;
}
static void signal_payload(VkSemaphore_T* s, uint64_t v, const char* who)
{
    if (s->timeline && v <= __atomic_load_n(&s->page->payload, __ATOMIC_ACQUIRE)) MOCK_FAIL("%s: timeline value %llu is not greater than the current payload", who, (unsigned long long)v);
    __atomic_store_n(&s->page->payload, s->timeline ? v : 1, __ATOMIC_RELEASE);
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_SignalSemaphore(VkDevice, const VkSemaphoreSignalInfo* si)
{
    if (!si->semaphore->timeline) MOCK_FAIL("vkSignalSemaphore: not a timeline semaphore");
    signal_payload(si->semaphore, si->value, "vkSignalSemaphore");
    return VK_SUCCESS;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_WaitSemaphores(VkDevice, const VkSemaphoreWaitInfo* wi, uint64_t timeout)
{
    for (uint32_t i = 0; i < wi->semaphoreCount; ++i) {
        if (!wi->pSemaphores[i]->timeline) MOCK_FAIL("vkWaitSemaphores: not a timeline semaphore");
        if (!wait_payload(wi->pSemaphores[i], wi->pValues[i], timeout)) return VK_TIMEOUT;
    }
    return VK_SUCCESS;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_GetSemaphoreCounterValue(VkDevice, VkSemaphore s, uint64_t* v)
{
    *v = __atomic_load_n(&s->page->payload, __ATOMIC_ACQUIRE);
    return VK_SUCCESS;
}

// ---- command buffers: closures, executed at submit ------------------------------------------------------------------
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateCommandPool(VkDevice, const VkCommandPoolCreateInfo*, const VkAllocationCallbacks*, VkCommandPool* out) { *out = new VkCommandPool_T(); return VK_SUCCESS; }
static VKAPI_ATTR void VKAPI_CALL mock_DestroyCommandPool(VkDevice, VkCommandPool p, const VkAllocationCallbacks*)
{
    if (!p) return;
    for (VkCommandBuffer_T* cb : p->buffers) delete cb;
    delete p;
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_AllocateCommandBuffers(VkDevice, const VkCommandBufferAllocateInfo* ai, VkCommandBuffer* out)
{
    for (uint32_t i = 0; i < ai->commandBufferCount; ++i) { out[i] = new VkCommandBuffer_T(); ai->commandPool->buffers.push_back(out[i]); }
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_FreeCommandBuffers(VkDevice, VkCommandPool pool, uint32_t n, const VkCommandBuffer* cbs)
{
    for (uint32_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < pool->buffers.size(); ++k)
            if (pool->buffers[k] == cbs[i]) { pool->buffers.erase(pool->buffers.begin() + k); break; }
        delete cbs[i];
    }
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_BeginCommandBuffer(VkCommandBuffer cb, const VkCommandBufferBeginInfo*) { cb->cmds.clear(); cb->recording = true; return VK_SUCCESS; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_EndCommandBuffer(VkCommandBuffer cb) { if (!cb->recording) MOCK_FAIL("vkEndCommandBuffer: not recording"); cb->recording = false; return VK_SUCCESS; }
static VKAPI_ATTR void VKAPI_CALL mock_CmdPipelineBarrier(VkCommandBuffer cb, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags, uint32_t, const VkMemoryBarrier*,
                                                          uint32_t, const VkBufferMemoryBarrier*, uint32_t n_image, const VkImageMemoryBarrier* image_barriers)
{
    if (!cb->recording) MOCK_FAIL("vkCmdPipelineBarrier: not recording");
    std::vector<VkImageMemoryBarrier> copy(image_barriers, image_barriers + n_image);
    cb->cmds.push_back([copy] {
        for (const auto& b : copy) {
            if (b.oldLayout != VK_IMAGE_LAYOUT_UNDEFINED && b.oldLayout != b.image->layout)
                MOCK_FAIL("image barrier: oldLayout %d but the image is in layout %d", (int)b.oldLayout, (int)b.image->layout);
            if (b.oldLayout == VK_IMAGE_LAYOUT_UNDEFINED && b.image->mem) memset(b.image->mem->map + b.image->offset, 0xEE, b.image->layer_stride() * b.image->layers);   // contents discarded
            b.image->layout = b.newLayout;
        }
    });
}
static void check_region(const VkBufferImageCopy& r, const VkBuffer_T* b, const VkImage_T* i, const char* who)
{
    if (r.imageOffset.x || r.imageOffset.y || r.imageOffset.z || r.imageExtent.width != i->w || r.imageExtent.height != i->h || r.imageExtent.depth != 1)
        MOCK_FAIL("%s: only whole-image regions are modelled", who);
    if (r.imageSubresource.aspectMask != VK_IMAGE_ASPECT_COLOR_BIT || r.imageSubresource.layerCount < 1 || r.imageSubresource.mipLevel != 0 ||
        r.imageSubresource.baseArrayLayer + r.imageSubresource.layerCount > i->layers) MOCK_FAIL("%s: bad subresource", who);
    const size_t row = (r.bufferRowLength ? r.bufferRowLength : i->w) * (size_t)i->texel;
    if (r.bufferOffset + row * i->h * r.imageSubresource.layerCount > b->size) MOCK_FAIL("%s: the region does not fit the buffer (%zu > %llu)", who, (size_t)r.bufferOffset + row * i->h, (unsigned long long)b->size);
    if (!b->mem || !i->mem) MOCK_FAIL("%s: unbound resource", who);
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdCopyImageToBuffer(VkCommandBuffer cb, VkImage src, VkImageLayout layout, VkBuffer dst, uint32_t n, const VkBufferImageCopy* regions)
{
    if (!cb->recording || n != 1) MOCK_FAIL("vkCmdCopyImageToBuffer: not recording / one region expected");
    if (!(src->usage & VK_IMAGE_USAGE_TRANSFER_SRC_BIT) || !(dst->usage & VK_BUFFER_USAGE_TRANSFER_DST_BIT)) MOCK_FAIL("vkCmdCopyImageToBuffer: missing TRANSFER usage");
    const VkBufferImageCopy r = regions[0];
    cb->cmds.push_back([=] {
        check_region(r, dst, src, "vkCmdCopyImageToBuffer");
        if (src->layout != layout || (layout != VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL && layout != VK_IMAGE_LAYOUT_GENERAL)) MOCK_FAIL("vkCmdCopyImageToBuffer: image is in layout %d, command says %d", (int)src->layout, (int)layout);
        const size_t row = (r.bufferRowLength ? r.bufferRowLength : src->w) * (size_t)src->texel;
        for (uint32_t l = 0; l < r.imageSubresource.layerCount; ++l)
            for (uint32_t y = 0; y < src->h; ++y)
                memcpy(dst->mem->map + dst->offset + r.bufferOffset + ((size_t)l * src->h + y) * row, src->row(r.imageSubresource.baseArrayLayer + l, y), (size_t)src->w * src->texel);
    });
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdCopyBufferToImage(VkCommandBuffer cb, VkBuffer src, VkImage dst, VkImageLayout layout, uint32_t n, const VkBufferImageCopy* regions)
{
    if (!cb->recording || n != 1) MOCK_FAIL("vkCmdCopyBufferToImage: not recording / one region expected");
    if (!(dst->usage & VK_IMAGE_USAGE_TRANSFER_DST_BIT) || !(src->usage & VK_BUFFER_USAGE_TRANSFER_SRC_BIT)) MOCK_FAIL("vkCmdCopyBufferToImage: missing TRANSFER usage");
    const VkBufferImageCopy r = regions[0];
    cb->cmds.push_back([=] {
        check_region(r, src, dst, "vkCmdCopyBufferToImage");
        if (dst->layout != layout || (layout != VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL && layout != VK_IMAGE_LAYOUT_GENERAL)) MOCK_FAIL("vkCmdCopyBufferToImage: image is in layout %d, command says %d", (int)dst->layout, (int)layout);
        const size_t row = (r.bufferRowLength ? r.bufferRowLength : dst->w) * (size_t)dst->texel;
        for (uint32_t l = 0; l < r.imageSubresource.layerCount; ++l)
            for (uint32_t y = 0; y < dst->h; ++y)
                memcpy(dst->row(r.imageSubresource.baseArrayLayer + l, y), src->mem->map + src->offset + r.bufferOffset + ((size_t)l * dst->h + y) * row, (size_t)dst->w * dst->texel);
    });
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_QueueSubmit(VkQueue, uint32_t n, const VkSubmitInfo* submits, VkFence fence)
{
    if (fence) MOCK_FAIL("vkQueueSubmit: fences are not modelled");
    for (uint32_t s = 0; s < n; ++s) {
        const VkSubmitInfo& si = submits[s];
        auto* tl = find_in_chain<VkTimelineSemaphoreSubmitInfo>(si.pNext, VK_STRUCTURE_TYPE_TIMELINE_SEMAPHORE_SUBMIT_INFO);
        for (uint32_t i = 0; i < si.waitSemaphoreCount; ++i) {
            VkSemaphore_T* sem = si.pWaitSemaphores[i];
            uint64_t want = 1;
            if (sem->timeline) {
                if (!tl || tl->waitSemaphoreValueCount != si.waitSemaphoreCount) MOCK_FAIL("vkQueueSubmit: timeline wait without VkTimelineSemaphoreSubmitInfo values");
                want = tl->pWaitSemaphoreValues[i];
            }
            if (!wait_payload(sem, want, 20ull * 1000000000ull)) {
                fprintf(stderr, "vkmock: vkQueueSubmit waited 20 s for semaphore value %llu (payload %llu): handshake out of order\n", (unsigned long long)want,
                        (unsigned long long)sem->page->payload);
                return VK_ERROR_DEVICE_LOST;
            }
            if (!sem->timeline) __atomic_store_n(&sem->page->payload, 0, __ATOMIC_RELEASE);
        }
        for (uint32_t c = 0; c < si.commandBufferCount; ++c) {
            if (si.pCommandBuffers[c]->recording) MOCK_FAIL("vkQueueSubmit: command buffer still recording");
            for (auto& cmd : si.pCommandBuffers[c]->cmds) cmd();
        }
        for (uint32_t i = 0; i < si.signalSemaphoreCount; ++i) {
            VkSemaphore_T* sem = si.pSignalSemaphores[i];
            uint64_t v = 1;
            if (sem->timeline) {
                if (!tl || tl->signalSemaphoreValueCount != si.signalSemaphoreCount) MOCK_FAIL("vkQueueSubmit: timeline signal without VkTimelineSemaphoreSubmitInfo values");
                v = tl->pSignalSemaphoreValues[i];
            }
            signal_payload(sem, v, "vkQueueSubmit");
        }
    }
    return VK_SUCCESS;
}


// ---- compute: pipelines whose "SPIR-V" names a shader of oracle/_ref/libref.so -----------------------------------------------
struct RefBinding { void* data; int width, height, layers, format; };
typedef int (*ref_dispatch_fn)(const char*, int, int, int, int, int, int, const void*, int, int, int, const RefBinding*, int);
static ref_dispatch_fn ref_dispatch()
{
    static ref_dispatch_fn fn = nullptr;
    if (!fn) {
        const char* path = getenv("VKMOCK_LIBREF");
        void* lib = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
        if (!lib) MOCK_FAIL("compute needs VKMOCK_LIBREF=<oracle/_ref/libref.so> (%s)", path ? dlerror() : "not set");
        fn = (ref_dispatch_fn)dlsym(lib, "ref_dispatch");
        if (!fn) MOCK_FAIL("libref.so has no ref_dispatch");
    }
    return fn;
}
static int shim_format(VkFormat f)      // oracle/ref.py: F_R32F .. F_R16F
{
    switch (f) {
    case VK_FORMAT_R32_SFLOAT: return 1; case VK_FORMAT_R32G32_SFLOAT: return 2; case VK_FORMAT_R8G8B8A8_UNORM: return 3;
    case VK_FORMAT_B8G8R8A8_UNORM: return 4; case VK_FORMAT_R16G16_SFLOAT: return 5; case VK_FORMAT_R8_UNORM: return 6;
    case VK_FORMAT_R16G16B16A16_SFLOAT: return 7; case VK_FORMAT_R32G32B32A32_SFLOAT: return 8; case VK_FORMAT_R16_SFLOAT: return 9;
    default: return 0;
    }
}
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateImageView(VkDevice, const VkImageViewCreateInfo* ci, const VkAllocationCallbacks*, VkImageView* out)
{
    if (!ci || ci->sType != VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO) MOCK_FAIL("vkCreateImageView: bad create info");
    const VkImageSubresourceRange& r = ci->subresourceRange;
    if (r.aspectMask != VK_IMAGE_ASPECT_COLOR_BIT || r.baseMipLevel != 0 || r.levelCount != 1 || r.baseArrayLayer + r.layerCount > ci->image->layers)
        MOCK_FAIL("vkCreateImageView: bad subresource range");
    if (ci->viewType == VK_IMAGE_VIEW_TYPE_2D && r.layerCount != 1) MOCK_FAIL("vkCreateImageView: a 2D view has one layer");
    if (ci->viewType != VK_IMAGE_VIEW_TYPE_2D && ci->viewType != VK_IMAGE_VIEW_TYPE_2D_ARRAY) MOCK_FAIL("vkCreateImageView: view type not modelled");
    if (texel_size(ci->format) != ci->image->texel) MOCK_FAIL("vkCreateImageView: view format is not size-compatible with the image");
    *out = new VkImageView_T{ci->image, ci->viewType, ci->format, r.baseArrayLayer, r.layerCount};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyImageView(VkDevice, VkImageView v, const VkAllocationCallbacks*) { delete v; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateSampler(VkDevice, const VkSamplerCreateInfo* ci, const VkAllocationCallbacks*, VkSampler* out)
{
    *out = new VkSampler_T{ci->magFilter, ci->minFilter, ci->addressModeU, ci->addressModeV, ci->unnormalizedCoordinates};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroySampler(VkDevice, VkSampler s, const VkAllocationCallbacks*) { delete s; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateShaderModule(VkDevice, const VkShaderModuleCreateInfo* ci, const VkAllocationCallbacks*, VkShaderModule* out)
{
    static const char tag[] = "VKMOCK-SHADER:";
    if (!ci || ci->codeSize % 4 || ci->codeSize < sizeof(tag)) MOCK_FAIL("vkCreateShaderModule: bad code size");
    if (ci->pCode[0] == 0x07230203u) MOCK_FAIL("vkCreateShaderModule: this is real SPIR-V; the mock only runs 'VKMOCK-SHADER:<name>' stand-ins");
    std::string text((const char*)ci->pCode, ci->codeSize);
    if (text.compare(0, sizeof(tag) - 1, tag) != 0) MOCK_FAIL("vkCreateShaderModule: neither SPIR-V nor a mock shader");
    text = text.substr(sizeof(tag) - 1);
    text = text.substr(0, text.find_first_of("\n \0", 0, 3));
    *out = new VkShaderModule_T{text};
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyShaderModule(VkDevice, VkShaderModule m, const VkAllocationCallbacks*) { delete m; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateDescriptorSetLayout(VkDevice, const VkDescriptorSetLayoutCreateInfo* ci, const VkAllocationCallbacks*, VkDescriptorSetLayout* out)
{
    auto* l = new VkDescriptorSetLayout_T();
    for (uint32_t i = 0; i < ci->bindingCount; ++i) {
        const VkDescriptorSetLayoutBinding& b = ci->pBindings[i];
        if (b.descriptorCount != 1 || !(b.stageFlags & VK_SHADER_STAGE_COMPUTE_BIT)) MOCK_FAIL("vkCreateDescriptorSetLayout: binding %u: one compute-visible descriptor expected", b.binding);
        if (b.descriptorType != VK_DESCRIPTOR_TYPE_STORAGE_IMAGE && b.descriptorType != VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER) MOCK_FAIL("descriptor type not modelled");
        l->bindings[b.binding] = b.descriptorType;
    }
    *out = l;
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyDescriptorSetLayout(VkDevice, VkDescriptorSetLayout l, const VkAllocationCallbacks*) { delete l; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreatePipelineLayout(VkDevice, const VkPipelineLayoutCreateInfo* ci, const VkAllocationCallbacks*, VkPipelineLayout* out)
{
    auto* l = new VkPipelineLayout_T();
    for (uint32_t i = 0; i < ci->setLayoutCount; ++i) l->sets.push_back(ci->pSetLayouts[i]);
    l->push_size = 0;
    for (uint32_t i = 0; i < ci->pushConstantRangeCount; ++i) {
        if (ci->pPushConstantRanges[i].offset != 0 || !(ci->pPushConstantRanges[i].stageFlags & VK_SHADER_STAGE_COMPUTE_BIT)) MOCK_FAIL("push constant range not modelled");
        l->push_size = ci->pPushConstantRanges[i].size;
    }
    if (l->push_size > 256) MOCK_FAIL("push constants larger than 256 bytes");
    *out = l;
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyPipelineLayout(VkDevice, VkPipelineLayout l, const VkAllocationCallbacks*) { delete l; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateComputePipelines(VkDevice, VkPipelineCache, uint32_t n, const VkComputePipelineCreateInfo* cis, const VkAllocationCallbacks*, VkPipeline* out)
{
    for (uint32_t i = 0; i < n; ++i) {
        const VkPipelineShaderStageCreateInfo& st = cis[i].stage;
        if (st.stage != VK_SHADER_STAGE_COMPUTE_BIT || strcmp(st.pName, "main")) MOCK_FAIL("vkCreateComputePipelines: compute stage 'main' expected");
        auto* p = new VkPipeline_T{st.module->name, {}, cis[i].layout};
        if (st.pSpecializationInfo)
            for (uint32_t e = 0; e < st.pSpecializationInfo->mapEntryCount; ++e) {
                const VkSpecializationMapEntry& m = st.pSpecializationInfo->pMapEntries[e];
                if (m.size != 4 || m.offset + 4 > st.pSpecializationInfo->dataSize) MOCK_FAIL("specialisation constants are 32-bit here");
                int32_t v; memcpy(&v, (const char*)st.pSpecializationInfo->pData + m.offset, 4);
                p->spec[m.constantID] = v;
            }
        out[i] = p;
    }
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyPipeline(VkDevice, VkPipeline p, const VkAllocationCallbacks*) { delete p; }
static VKAPI_ATTR VkResult VKAPI_CALL mock_CreateDescriptorPool(VkDevice, const VkDescriptorPoolCreateInfo*, const VkAllocationCallbacks*, VkDescriptorPool* out) { *out = new VkDescriptorPool_T(); return VK_SUCCESS; }
static VKAPI_ATTR void VKAPI_CALL mock_DestroyDescriptorPool(VkDevice, VkDescriptorPool p, const VkAllocationCallbacks*);
static VKAPI_ATTR VkResult VKAPI_CALL mock_AllocateDescriptorSets(VkDevice, const VkDescriptorSetAllocateInfo* ai, VkDescriptorSet* out)
{
    for (uint32_t i = 0; i < ai->descriptorSetCount; ++i) { out[i] = new VkDescriptorSet_T{ai->pSetLayouts[i], {}}; ai->descriptorPool->sets.push_back(out[i]); }
    return VK_SUCCESS;
}
static VKAPI_ATTR void VKAPI_CALL mock_DestroyDescriptorPool(VkDevice, VkDescriptorPool p, const VkAllocationCallbacks*)
{
    if (!p) return;
    for (VkDescriptorSet_T* s : p->sets) delete s;
    delete p;
}
static VKAPI_ATTR void VKAPI_CALL mock_UpdateDescriptorSets(VkDevice, uint32_t n, const VkWriteDescriptorSet* writes, uint32_t ncopies, const VkCopyDescriptorSet*)
{
    if (ncopies) MOCK_FAIL("vkUpdateDescriptorSets: copies not modelled");
    for (uint32_t i = 0; i < n; ++i) {
        const VkWriteDescriptorSet& w = writes[i];
        if (w.descriptorCount != 1 || w.dstArrayElement != 0 || !w.pImageInfo) MOCK_FAIL("vkUpdateDescriptorSets: one image descriptor per write expected");
        auto it = w.dstSet->layout->bindings.find(w.dstBinding);
        if (it == w.dstSet->layout->bindings.end()) MOCK_FAIL("vkUpdateDescriptorSets: binding %u is not in the set layout", w.dstBinding);
        if (it->second != w.descriptorType) MOCK_FAIL("vkUpdateDescriptorSets: binding %u: descriptor type %d, layout says %d", w.dstBinding, (int)w.descriptorType, (int)it->second);
        if (w.descriptorType == VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER && !w.pImageInfo->sampler) MOCK_FAIL("binding %u: combined image sampler without a sampler", w.dstBinding);
        w.dstSet->bound[w.dstBinding] = Descriptor{w.descriptorType, w.pImageInfo->imageView, w.pImageInfo->sampler, w.pImageInfo->imageLayout};
    }
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdBindPipeline(VkCommandBuffer cb, VkPipelineBindPoint bp, VkPipeline p)
{
    if (bp != VK_PIPELINE_BIND_POINT_COMPUTE) MOCK_FAIL("only compute pipelines");
    cb->pipeline = p;
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdBindDescriptorSets(VkCommandBuffer cb, VkPipelineBindPoint bp, VkPipelineLayout, uint32_t first, uint32_t n, const VkDescriptorSet* sets, uint32_t ndyn, const uint32_t*)
{
    if (bp != VK_PIPELINE_BIND_POINT_COMPUTE || first != 0 || n != 1 || ndyn) MOCK_FAIL("vkCmdBindDescriptorSets: one compute set at index 0 expected");
    cb->set = sets[0];
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdPushConstants(VkCommandBuffer cb, VkPipelineLayout layout, VkShaderStageFlags stages, uint32_t offset, uint32_t size, const void* data)
{
    if (!(stages & VK_SHADER_STAGE_COMPUTE_BIT) || offset + size > layout->push_size) MOCK_FAIL("vkCmdPushConstants: outside the layout's range (%u + %u > %u)", offset, size, layout->push_size);
    if (cb->push.size() < offset + size) cb->push.resize(offset + size);
    memcpy(cb->push.data() + offset, data, size);
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdDispatch(VkCommandBuffer cb, uint32_t gx, uint32_t gy, uint32_t gz)
{
    if (!cb->recording || !cb->pipeline || !cb->set || gz != 1) MOCK_FAIL("vkCmdDispatch: needs a bound pipeline and descriptor set, z = 1");
    VkPipeline_T* pipe = cb->pipeline;
    VkDescriptorSet_T* set = cb->set;
    if (pipe->layout->sets.size() != 1 || pipe->layout->sets[0] != set->layout) MOCK_FAIL("vkCmdDispatch: the bound set's layout is not the pipeline layout's");
    std::vector<unsigned char> push = cb->push;
    if (push.size() < pipe->layout->push_size) MOCK_FAIL("vkCmdDispatch: %zu bytes of push constants pushed, the layout declares %u", push.size(), pipe->layout->push_size);
    cb->cmds.push_back([=] {
        auto spec = [&](uint32_t id) { auto it = pipe->spec.find(id); if (it == pipe->spec.end()) MOCK_FAIL("shader %s: specialisation constant %u not set", pipe->shader.c_str(), id); return (int)it->second; };
        std::string name = pipe->shader;
        int k0, k1, k2 = 0, W = 0, H = 0, radius = 0;
        const bool acc = name.rfind("accumulator", 0) == 0, conv = name == "formatConverter";
        if (acc || conv) { k0 = spec(0); k1 = spec(1); }
        else {
            W = spec(0); H = spec(1); k0 = spec(2); k1 = spec(3);
            if (name.rfind("bmfr", 0) == 0) k2 = spec(4);
            if (name == "bfrBlender") radius = spec(4);
            if ((name == "bmfrPre" || name == "bmfrPost") && pipe->spec.count(5) && pipe->spec.at(5)) name += pipe->spec.at(5) == 1 ? "_w1" : "_w2";
        }
        RefBinding arr[32];
        memset(arr, 0, sizeof arr);
        std::vector<std::vector<char>> tight(32);
        int nb = 0;
        for (auto& kv : set->layout->bindings) {
            const uint32_t b = kv.first;
            if (b >= 32) MOCK_FAIL("binding %u too large", b);
            auto it = set->bound.find(b);
            if (it == set->bound.end()) MOCK_FAIL("shader %s: binding %u was never written", name.c_str(), b);
            const Descriptor& d = it->second;
            VkImage_T* img = d.view->image;
            if (img->layout != VK_IMAGE_LAYOUT_GENERAL || d.layout != VK_IMAGE_LAYOUT_GENERAL) MOCK_FAIL("shader %s: binding %u: image must be in GENERAL layout", name.c_str(), b);
            if (d.type == VK_DESCRIPTOR_TYPE_STORAGE_IMAGE && !(img->usage & VK_IMAGE_USAGE_STORAGE_BIT)) MOCK_FAIL("binding %u: image lacks STORAGE usage", b);
            if (d.type == VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER) {
                if (!(img->usage & VK_IMAGE_USAGE_SAMPLED_BIT)) MOCK_FAIL("binding %u: image lacks SAMPLED usage", b);
                // vsg::Sampler defaults (external/vsg/include/vsg/state/Sampler.h:29-43), which is what the shim's texture() implements
                if (d.sampler->mag != VK_FILTER_LINEAR || d.sampler->min != VK_FILTER_LINEAR || d.sampler->u != VK_SAMPLER_ADDRESS_MODE_REPEAT ||
                    d.sampler->v != VK_SAMPLER_ADDRESS_MODE_REPEAT || d.sampler->unnormalized) MOCK_FAIL("binding %u: sampler is not LINEAR / REPEAT / normalised", b);
            }
            const size_t row = (size_t)img->w * img->texel;
            tight[b].resize(row * img->h * d.view->layer_count);
            for (uint32_t l = 0; l < d.view->layer_count; ++l)
                for (uint32_t y = 0; y < img->h; ++y) memcpy(tight[b].data() + ((size_t)l * img->h + y) * row, img->row(d.view->base_layer + l, y), row);
            arr[b] = RefBinding{tight[b].data(), (int)img->w, (int)img->h, (int)d.view->layer_count, shim_format(d.view->format)};
            if ((int)b + 1 > nb) nb = (int)b + 1;
            if (acc && b == 1) { W = (int)img->w; H = (int)img->h; }
            if (conv && b == 0) { W = (int)img->w; H = (int)img->h; }
        }
        const int rc = ref_dispatch()(name.c_str(), k0, k1, k2, W, H, radius, push.empty() ? nullptr : push.data(), (int)pipe->layout->push_size, (int)gx, (int)gy, arr, nb);
        if (rc) MOCK_FAIL("shader %s with key (%d, %d, %d) is not in libref.so (rc %d)", name.c_str(), k0, k1, k2, rc);
        for (auto& kv : set->bound) {                 // storage images may have been written
            const Descriptor& d = kv.second;
            if (d.type != VK_DESCRIPTOR_TYPE_STORAGE_IMAGE) continue;
            VkImage_T* img = d.view->image;
            const size_t row = (size_t)img->w * img->texel;
            for (uint32_t l = 0; l < d.view->layer_count; ++l)
                for (uint32_t y = 0; y < img->h; ++y) memcpy(img->row(d.view->base_layer + l, y), tight[kv.first].data() + ((size_t)l * img->h + y) * row, row);
        }
    });
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdCopyImage(VkCommandBuffer cb, VkImage src, VkImageLayout sl, VkImage dst, VkImageLayout dl, uint32_t n, const VkImageCopy* regions)
{
    if (!cb->recording || n != 1) MOCK_FAIL("vkCmdCopyImage: not recording / one region expected");
    if (!(src->usage & VK_IMAGE_USAGE_TRANSFER_SRC_BIT) || !(dst->usage & VK_IMAGE_USAGE_TRANSFER_DST_BIT)) MOCK_FAIL("vkCmdCopyImage: missing TRANSFER usage");
    const VkImageCopy r = regions[0];
    cb->cmds.push_back([=] {
        if (src->texel != dst->texel) MOCK_FAIL("vkCmdCopyImage: formats are not size-compatible");
        if (src->layout != sl || dst->layout != dl) MOCK_FAIL("vkCmdCopyImage: layouts (%d, %d), command says (%d, %d)", (int)src->layout, (int)dst->layout, (int)sl, (int)dl);
        if ((sl != VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL && sl != VK_IMAGE_LAYOUT_GENERAL) || (dl != VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL && dl != VK_IMAGE_LAYOUT_GENERAL)) MOCK_FAIL("vkCmdCopyImage: bad layouts");
        if (r.extent.width != src->w || r.extent.height != src->h || r.extent.width != dst->w || r.extent.height != dst->h || r.srcOffset.x || r.srcOffset.y || r.dstOffset.x || r.dstOffset.y)
            MOCK_FAIL("vkCmdCopyImage: only whole-image copies are modelled");
        if (r.srcSubresource.layerCount != r.dstSubresource.layerCount || r.srcSubresource.baseArrayLayer + r.srcSubresource.layerCount > src->layers ||
            r.dstSubresource.baseArrayLayer + r.dstSubresource.layerCount > dst->layers) MOCK_FAIL("vkCmdCopyImage: bad layers");
        for (uint32_t l = 0; l < r.srcSubresource.layerCount; ++l)
            for (uint32_t y = 0; y < src->h; ++y) memcpy(dst->row(r.dstSubresource.baseArrayLayer + l, y), src->row(r.srcSubresource.baseArrayLayer + l, y), (size_t)src->w * src->texel);
    });
}
static VKAPI_ATTR void VKAPI_CALL mock_CmdClearColorImage(VkCommandBuffer cb, VkImage img, VkImageLayout layout, const VkClearColorValue* color, uint32_t n, const VkImageSubresourceRange* ranges)
{
    if (!cb->recording || n != 1) MOCK_FAIL("vkCmdClearColorImage: not recording / one range expected");
    const VkClearColorValue c = *color;
    const VkImageSubresourceRange r = ranges[0];
    cb->cmds.push_back([=] {
        if (img->layout != layout || (layout != VK_IMAGE_LAYOUT_GENERAL && layout != VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL)) MOCK_FAIL("vkCmdClearColorImage: bad layout");
        if (c.uint32[0] || c.uint32[1] || c.uint32[2] || c.uint32[3]) MOCK_FAIL("vkCmdClearColorImage: only clears to zero are modelled");
        const uint32_t layers = r.layerCount == VK_REMAINING_ARRAY_LAYERS ? img->layers - r.baseArrayLayer : r.layerCount;
        for (uint32_t l = 0; l < layers; ++l)
            for (uint32_t y = 0; y < img->h; ++y) memset(img->row(r.baseArrayLayer + l, y), 0, (size_t)img->w * img->texel);
    });
}

// ---- dispatch -------------------------------------------------------------------------------------------------------
static VKAPI_ATTR PFN_vkVoidFunction VKAPI_CALL mock_GetDeviceProcAddr(VkDevice d, const char* name);
extern "C" __attribute__((visibility("default"))) VKAPI_ATTR PFN_vkVoidFunction VKAPI_CALL vkGetInstanceProcAddr(VkInstance instance, const char* name);

struct Entry { const char* name; PFN_vkVoidFunction fn; int need; };     // need: 0 none, 1 external_memory_fd, 2 external_semaphore_fd, 3 timeline feature
#define E(name, need) {"vk" #name, (PFN_vkVoidFunction)mock_##name, need}
static const Entry kEntries[] = {
    E(CreateInstance, 0), E(DestroyInstance, 0), E(EnumeratePhysicalDevices, 0), E(GetPhysicalDeviceProperties, 0), E(GetPhysicalDeviceProperties2, 0),
    E(GetPhysicalDeviceFeatures2, 0), E(GetPhysicalDeviceQueueFamilyProperties, 0), E(GetPhysicalDeviceMemoryProperties, 0),
    E(GetPhysicalDeviceExternalBufferProperties, 0), E(GetPhysicalDeviceExternalSemaphoreProperties, 0), E(EnumerateDeviceExtensionProperties, 0),
    E(CreateDevice, 0), E(GetDeviceProcAddr, 0), E(DestroyDevice, 0), E(GetDeviceQueue, 0), E(DeviceWaitIdle, 0), E(QueueWaitIdle, 0), E(QueueSubmit, 0),
    E(CreateBuffer, 0), E(DestroyBuffer, 0), E(GetBufferMemoryRequirements, 0), E(CreateImage, 0), E(DestroyImage, 0), E(GetImageMemoryRequirements, 0),
    E(AllocateMemory, 0), E(FreeMemory, 0), E(BindBufferMemory, 0), E(BindImageMemory, 0), E(MapMemory, 0), E(UnmapMemory, 0), E(GetMemoryFdKHR, 1),
    E(CreateSemaphore, 0), E(DestroySemaphore, 0), E(GetSemaphoreFdKHR, 2), E(SignalSemaphore, 3), E(WaitSemaphores, 3), E(GetSemaphoreCounterValue, 3),
    E(CreateCommandPool, 0), E(DestroyCommandPool, 0), E(AllocateCommandBuffers, 0), E(FreeCommandBuffers, 0), E(BeginCommandBuffer, 0), E(EndCommandBuffer, 0),
    E(CmdPipelineBarrier, 0), E(CmdCopyImageToBuffer, 0), E(CmdCopyBufferToImage, 0),
    E(CreateImageView, 0), E(DestroyImageView, 0), E(CreateSampler, 0), E(DestroySampler, 0), E(CreateShaderModule, 0), E(DestroyShaderModule, 0),
    E(CreateDescriptorSetLayout, 0), E(DestroyDescriptorSetLayout, 0), E(CreatePipelineLayout, 0), E(DestroyPipelineLayout, 0),
    E(CreateComputePipelines, 0), E(DestroyPipeline, 0), E(CreateDescriptorPool, 0), E(DestroyDescriptorPool, 0), E(AllocateDescriptorSets, 0),
    E(UpdateDescriptorSets, 0), E(CmdBindPipeline, 0), E(CmdBindDescriptorSets, 0), E(CmdPushConstants, 0), E(CmdDispatch, 0), E(CmdCopyImage, 0),
    E(CmdClearColorImage, 0),
};
#undef E

static VKAPI_ATTR PFN_vkVoidFunction VKAPI_CALL mock_GetDeviceProcAddr(VkDevice d, const char* name)
{
    for (const Entry& e : kEntries)
        if (!strcmp(e.name, name)) {
            if ((e.need == 1 && !d->ext_memory_fd) || (e.need == 2 && !d->ext_semaphore_fd) || (e.need == 3 && !d->timeline)) return nullptr;
            return e.fn;
        }
    return nullptr;
}

extern "C" VKAPI_ATTR PFN_vkVoidFunction VKAPI_CALL vkGetInstanceProcAddr(VkInstance, const char* name)
{
    if (!strcmp(name, "vkGetInstanceProcAddr")) return (PFN_vkVoidFunction)vkGetInstanceProcAddr;
    for (const Entry& e : kEntries)
        if (!strcmp(e.name, name)) return e.need ? nullptr : e.fn;      // extension / feature entry points: through the device
    return nullptr;
}
