// vk_oracle.cpp -- the "second oracle" of SURVEY.md section 8(c): a raw-Vulkan compute host (no vsg) that replays the
// reference's denoising command list with the reference's OWN compiled shaders on whatever Vulkan implementation is
// present (meant for Mesa lavapipe on a CPU-only box), so that the hand-written oracle (oracle/vkpbrt_oracle.c) can be
// compared with a real Vulkan execution of shaders/*.comp.  TEST / VALIDATION TOOLING, not product.
//
// It restates the reference's HOST side -- nothing of this repository's product code is used:
//   images and formats      source/buffers/GBuffer.cpp:53-125, IlluminationBuffer.cpp:223-282, AccumulationBuffer.cpp:245-339,
//                           denoisers/BMFR.cpp:56-133, BFR.cpp:33-80, Taa.cpp:21-60
//   samplers                vsg::Sampler defaults (external/vsg/include/vsg/state/Sampler.h:29-43): LINEAR, REPEAT, normalised
//   descriptor bindings     shaders/accumulator.comp:5-18, bmfrGeneral.comp:3-14 (BMFR.hpp:28-30), bfr.comp:5-14, taa.comp:8-11
//   specialisation          Accumulator.cpp:21-24, BMFR.cpp:32-54, BFR.cpp:26-31, BFRBlender.cpp:13-19, Taa.cpp:14-19
//   push constants          Accumulator.hpp:28-33 + Accumulator.cpp:85-117 (separate matrices), PipelineStructs.hpp:6-13 +
//                           VulkanPBRT.cpp:561-563, :591
//   dispatch sizes / order  Accumulator.cpp:72-83, BMFR.cpp:203-230, BFR.cpp:128-138, BFRBlender.cpp:80-90, Taa.cpp:99-107,
//                           util/DenoiserUtils.cpp:8-130, VulkanPBRT.cpp:551-618
//   end-of-frame copies     Taa.cpp:106 (final -> accumulation), AccumulationBuffer.cpp:72-244
//
//   vk_oracle <spv_dir> <frames_dir> <out_dir> <width> <height> <first_frame> <frames> <bmfr|bfr|bmfrx3|bfrx3> <block> <taa 0|1>
//     <spv_dir>/{accumulator_sep,bmfrPre,bmfrFit,bmfrPost,bfr,bfrBlender,taa}.comp.spv   (README.md: how to build them with glslc)
//     <frames_dir>/frame_%d.{depth,normal,albedo,illum,cam}                      (same raw files as examples/cpp_frame_loop.cpp)
//     <frames_dir>/frame_%d.avgsq   optional rgba16f plane loaded into illuminationSquared before the denoisers: no shader ever
//                                   writes that image (accumulator.comp:104), and the blender of the x3 wirings reads it
//     -> <out_dir>/{final_%d.bgra, denoised<b>_%d.rgba16f (2 layers, per block size), motion_%d.rg16f, spp_%d.r8, illum_%d.rgba16f}
//   x3 = DenoisingBlockSize::X8X16X32: three denoisers (b = 8, 16, 32) + BFRBlender (util/DenoiserUtils.cpp:48-70, :106-124)
//   VK_ORACLE_LOADER: Vulkan loader to dlopen (default libvulkan.so.1).
//
// Status: neither machine of this project has a Vulkan loader, an ICD or glslc, so this program has never met a real
// driver.  It is compiled against Khronos' vulkan_core.h and run against tests/vkmock, whose vkCmdDispatch executes the
// reference's shader SOURCE (oracle/glsl_shim) behind the Vulkan API: tests/test_vk_oracle.py requires its output to
// equal the oracle bit for bit, which checks every binding number, descriptor type, format, specialisation constant,
// push-constant block, dispatch size and copy above.
#include <vulkan/vulkan_core.h>

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define VK_FUNCTIONS(X) \
    X(EnumeratePhysicalDevices) X(GetPhysicalDeviceProperties) X(GetPhysicalDeviceMemoryProperties) X(GetPhysicalDeviceQueueFamilyProperties) X(CreateDevice) \
    X(GetDeviceProcAddr)
#define VK_DEVICE_FUNCTIONS(X) \
    X(DestroyDevice) X(GetDeviceQueue) X(QueueSubmit) X(QueueWaitIdle) X(CreateBuffer) X(DestroyBuffer) X(GetBufferMemoryRequirements) X(CreateImage) X(DestroyImage) \
    X(GetImageMemoryRequirements) X(AllocateMemory) X(FreeMemory) X(BindBufferMemory) X(BindImageMemory) X(MapMemory) X(CreateImageView) X(DestroyImageView) \
    X(CreateSampler) X(DestroySampler) X(CreateShaderModule) X(DestroyShaderModule) X(CreateDescriptorSetLayout) X(DestroyDescriptorSetLayout) \
    X(CreatePipelineLayout) X(DestroyPipelineLayout) X(CreateComputePipelines) X(DestroyPipeline) X(CreateDescriptorPool) X(DestroyDescriptorPool) \
    X(AllocateDescriptorSets) X(UpdateDescriptorSets) X(CreateCommandPool) X(DestroyCommandPool) X(AllocateCommandBuffers) X(BeginCommandBuffer) \
    X(EndCommandBuffer) X(CmdPipelineBarrier) X(CmdCopyBufferToImage) X(CmdCopyImageToBuffer) X(CmdCopyImage) X(CmdClearColorImage) X(CmdBindPipeline) \
    X(CmdBindDescriptorSets) X(CmdPushConstants) X(CmdDispatch)

struct Vk {
    PFN_vkGetInstanceProcAddr gipa = nullptr;
    VkInstance instance = VK_NULL_HANDLE;
    VkPhysicalDevice physical_device = VK_NULL_HANDLE;
    VkDevice device = VK_NULL_HANDLE;
    VkQueue queue = VK_NULL_HANDLE;
    uint32_t queue_family = 0;
    VkPhysicalDeviceMemoryProperties memory{};
#define DECLARE(name) PFN_vk##name name = nullptr;
    VK_FUNCTIONS(DECLARE)
    VK_DEVICE_FUNCTIONS(DECLARE)
#undef DECLARE
};

static void vk_check(VkResult r, const char* what)
{
    if (r != VK_SUCCESS) throw std::runtime_error(std::string(what) + " failed (VkResult " + std::to_string((int)r) + ")");
}

static uint32_t memory_type(const Vk& vk, uint32_t bits, VkMemoryPropertyFlags flags)
{
    for (uint32_t i = 0; i < vk.memory.memoryTypeCount; ++i)
        if ((bits & (1u << i)) && (vk.memory.memoryTypes[i].propertyFlags & flags) == flags) return i;
    throw std::runtime_error("no suitable memory type");
}

static void init_vulkan(Vk& vk)
{
    const char* lib_name = getenv("VK_ORACLE_LOADER");
    if (!lib_name || !*lib_name) lib_name = "libvulkan.so.1";
    void* lib = dlopen(lib_name, RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("cannot load ") + lib_name + ": " + dlerror());
    vk.gipa = reinterpret_cast<PFN_vkGetInstanceProcAddr>(dlsym(lib, "vkGetInstanceProcAddr"));
    if (!vk.gipa) throw std::runtime_error("the loader has no vkGetInstanceProcAddr");
    auto create_instance = reinterpret_cast<PFN_vkCreateInstance>(vk.gipa(VK_NULL_HANDLE, "vkCreateInstance"));
    if (!create_instance) throw std::runtime_error("the loader has no vkCreateInstance");
    VkApplicationInfo app{};
    app.sType = VK_STRUCTURE_TYPE_APPLICATION_INFO;
    app.pApplicationName = "vk_oracle";
    app.apiVersion = VK_API_VERSION_1_2;                     // VulkanPBRT.cpp:176
    VkInstanceCreateInfo ici{};
    ici.sType = VK_STRUCTURE_TYPE_INSTANCE_CREATE_INFO;
    ici.pApplicationInfo = &app;
    vk_check(create_instance(&ici, nullptr, &vk.instance), "vkCreateInstance");
#define LOAD(name) \
    vk.name = reinterpret_cast<PFN_vk##name>(vk.gipa(vk.instance, "vk" #name)); \
    if (!vk.name) throw std::runtime_error("missing vk" #name);
    VK_FUNCTIONS(LOAD)
#undef LOAD
    uint32_t n = 0;
    vk_check(vk.EnumeratePhysicalDevices(vk.instance, &n, nullptr), "vkEnumeratePhysicalDevices");
    if (!n) throw std::runtime_error("no Vulkan physical device");
    std::vector<VkPhysicalDevice> devices(n);
    vk.EnumeratePhysicalDevices(vk.instance, &n, devices.data());
    vk.physical_device = devices[0];
    const char* pick = getenv("VK_ORACLE_DEVICE");          // substring of the device name, e.g. "llvmpipe"
    for (VkPhysicalDevice d : devices) {
        VkPhysicalDeviceProperties p;
        vk.GetPhysicalDeviceProperties(d, &p);
        if (pick && strstr(p.deviceName, pick)) vk.physical_device = d;
    }
    VkPhysicalDeviceProperties props;
    vk.GetPhysicalDeviceProperties(vk.physical_device, &props);
    fprintf(stderr, "vk_oracle: device '%s'\n", props.deviceName);
    vk.GetPhysicalDeviceMemoryProperties(vk.physical_device, &vk.memory);
    uint32_t nq = 0;
    vk.GetPhysicalDeviceQueueFamilyProperties(vk.physical_device, &nq, nullptr);
    std::vector<VkQueueFamilyProperties> families(nq);
    vk.GetPhysicalDeviceQueueFamilyProperties(vk.physical_device, &nq, families.data());
    vk.queue_family = UINT32_MAX;
    for (uint32_t i = 0; i < nq; ++i)
        if (families[i].queueFlags & VK_QUEUE_COMPUTE_BIT) { vk.queue_family = i; break; }
    if (vk.queue_family == UINT32_MAX) throw std::runtime_error("no compute queue");
    const float priority = 1.f;
    VkDeviceQueueCreateInfo qi{};
    qi.sType = VK_STRUCTURE_TYPE_DEVICE_QUEUE_CREATE_INFO;
    qi.queueFamilyIndex = vk.queue_family;
    qi.queueCount = 1;
    qi.pQueuePriorities = &priority;
    VkPhysicalDeviceFeatures features{};
    features.shaderStorageImageExtendedFormats = VK_TRUE;     // rg32f / rg16f / r16f / r8 storage images
    features.shaderStorageImageWriteWithoutFormat = VK_TRUE;  // the BGRA8 finals are written through an rgba8 image declaration
    features.shaderStorageImageReadWithoutFormat = VK_TRUE;
    VkDeviceCreateInfo dci{};
    dci.sType = VK_STRUCTURE_TYPE_DEVICE_CREATE_INFO;
    dci.queueCreateInfoCount = 1;
    dci.pQueueCreateInfos = &qi;
    dci.pEnabledFeatures = &features;
    vk_check(vk.CreateDevice(vk.physical_device, &dci, nullptr, &vk.device), "vkCreateDevice");
#define LOAD(name) \
    vk.name = reinterpret_cast<PFN_vk##name>(vk.GetDeviceProcAddr(vk.device, "vk" #name)); \
    if (!vk.name) throw std::runtime_error("missing vk" #name);
    VK_DEVICE_FUNCTIONS(LOAD)
#undef LOAD
    vk.GetDeviceQueue(vk.device, vk.queue_family, 0, &vk.queue);
}

struct Image {
    VkImage image = VK_NULL_HANDLE;
    VkDeviceMemory memory = VK_NULL_HANDLE;
    VkImageView view = VK_NULL_HANDLE;
    VkFormat format = VK_FORMAT_UNDEFINED;
    uint32_t width = 0, height = 0, layers = 1, texel = 0;
    VkDeviceSize bytes() const { return (VkDeviceSize)width * height * layers * texel; }
};

static uint32_t texel_size(VkFormat f)
{
    switch (f) {
    case VK_FORMAT_R8_UNORM: return 1;
    case VK_FORMAT_R16_SFLOAT: return 2;
    case VK_FORMAT_R32_SFLOAT: case VK_FORMAT_R8G8B8A8_UNORM: case VK_FORMAT_B8G8R8A8_UNORM: case VK_FORMAT_R16G16_SFLOAT: return 4;
    case VK_FORMAT_R32G32_SFLOAT: case VK_FORMAT_R16G16B16A16_SFLOAT: return 8;
    case VK_FORMAT_R32G32B32A32_SFLOAT: return 16;
    default: throw std::runtime_error("format");
    }
}

static Image make_image(const Vk& vk, VkFormat format, uint32_t w, uint32_t h, uint32_t layers = 1, bool array_view = false)
{
    Image im;
    im.format = format; im.width = w; im.height = h; im.layers = layers; im.texel = texel_size(format);
    VkImageCreateInfo ci{};
    ci.sType = VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO;
    ci.imageType = VK_IMAGE_TYPE_2D;
    ci.format = format;
    ci.extent = {w, h, 1};
    ci.mipLevels = 1;
    ci.arrayLayers = layers;
    ci.samples = VK_SAMPLE_COUNT_1_BIT;
    ci.tiling = VK_IMAGE_TILING_OPTIMAL;
    ci.usage = VK_IMAGE_USAGE_STORAGE_BIT | VK_IMAGE_USAGE_SAMPLED_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT;
    ci.initialLayout = VK_IMAGE_LAYOUT_UNDEFINED;
    vk_check(vk.CreateImage(vk.device, &ci, nullptr, &im.image), "vkCreateImage");
    VkMemoryRequirements req;
    vk.GetImageMemoryRequirements(vk.device, im.image, &req);
    VkMemoryAllocateInfo ai{};
    ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
    ai.allocationSize = req.size;
    ai.memoryTypeIndex = memory_type(vk, req.memoryTypeBits, VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT);
    vk_check(vk.AllocateMemory(vk.device, &ai, nullptr, &im.memory), "vkAllocateMemory");
    vk_check(vk.BindImageMemory(vk.device, im.image, im.memory, 0), "vkBindImageMemory");
    VkImageViewCreateInfo vi{};
    vi.sType = VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO;
    vi.image = im.image;
    vi.viewType = (layers > 1 || array_view) ? VK_IMAGE_VIEW_TYPE_2D_ARRAY : VK_IMAGE_VIEW_TYPE_2D;
    vi.format = format;
    vi.subresourceRange = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, layers};
    vk_check(vk.CreateImageView(vk.device, &vi, nullptr, &im.view), "vkCreateImageView");
    return im;
}

struct HostBuffer {
    VkBuffer buffer = VK_NULL_HANDLE;
    VkDeviceMemory memory = VK_NULL_HANDLE;
    char* map = nullptr;
};
static HostBuffer make_host_buffer(const Vk& vk, VkDeviceSize size)
{
    HostBuffer b;
    VkBufferCreateInfo ci{};
    ci.sType = VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO;
    ci.size = size;
    ci.usage = VK_BUFFER_USAGE_TRANSFER_SRC_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT;
    ci.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
    vk_check(vk.CreateBuffer(vk.device, &ci, nullptr, &b.buffer), "vkCreateBuffer");
    VkMemoryRequirements req;
    vk.GetBufferMemoryRequirements(vk.device, b.buffer, &req);
    VkMemoryAllocateInfo ai{};
    ai.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;
    ai.allocationSize = req.size;
    ai.memoryTypeIndex = memory_type(vk, req.memoryTypeBits, VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT | VK_MEMORY_PROPERTY_HOST_COHERENT_BIT);
    vk_check(vk.AllocateMemory(vk.device, &ai, nullptr, &b.memory), "vkAllocateMemory");
    vk_check(vk.BindBufferMemory(vk.device, b.buffer, b.memory, 0), "vkBindBufferMemory");
    void* p = nullptr;
    vk_check(vk.MapMemory(vk.device, b.memory, 0, VK_WHOLE_SIZE, 0, &p), "vkMapMemory");
    b.map = static_cast<char*>(p);
    return b;
}

static std::vector<char> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + p);
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// One shader = descriptor set layout + pipeline layout + pipeline + descriptor set
struct Binding { uint32_t binding; VkDescriptorType type; const Image* image; };
struct Stage {
    VkDescriptorSetLayout set_layout = VK_NULL_HANDLE;
    VkPipelineLayout layout = VK_NULL_HANDLE;
    VkDescriptorSet set = VK_NULL_HANDLE;
    std::vector<VkPipeline> pipelines;
    std::vector<VkShaderModule> modules;
};

static Stage make_stage(const Vk& vk, VkDescriptorPool pool, VkSampler sampler, const std::vector<Binding>& bindings, uint32_t push_size)
{
    Stage s;
    std::vector<VkDescriptorSetLayoutBinding> lb;
    for (const Binding& b : bindings) {
        VkDescriptorSetLayoutBinding l{};
        l.binding = b.binding;
        l.descriptorType = b.type;
        l.descriptorCount = 1;
        l.stageFlags = VK_SHADER_STAGE_COMPUTE_BIT;
        lb.push_back(l);
    }
    VkDescriptorSetLayoutCreateInfo li{};
    li.sType = VK_STRUCTURE_TYPE_DESCRIPTOR_SET_LAYOUT_CREATE_INFO;
    li.bindingCount = (uint32_t)lb.size();
    li.pBindings = lb.data();
    vk_check(vk.CreateDescriptorSetLayout(vk.device, &li, nullptr, &s.set_layout), "vkCreateDescriptorSetLayout");
    VkPushConstantRange range{VK_SHADER_STAGE_COMPUTE_BIT, 0, push_size};
    VkPipelineLayoutCreateInfo pi{};
    pi.sType = VK_STRUCTURE_TYPE_PIPELINE_LAYOUT_CREATE_INFO;
    pi.setLayoutCount = 1;
    pi.pSetLayouts = &s.set_layout;
    pi.pushConstantRangeCount = push_size ? 1 : 0;
    pi.pPushConstantRanges = &range;
    vk_check(vk.CreatePipelineLayout(vk.device, &pi, nullptr, &s.layout), "vkCreatePipelineLayout");
    VkDescriptorSetAllocateInfo ai{};
    ai.sType = VK_STRUCTURE_TYPE_DESCRIPTOR_SET_ALLOCATE_INFO;
    ai.descriptorPool = pool;
    ai.descriptorSetCount = 1;
    ai.pSetLayouts = &s.set_layout;
    vk_check(vk.AllocateDescriptorSets(vk.device, &ai, &s.set), "vkAllocateDescriptorSets");
    std::vector<VkDescriptorImageInfo> infos(bindings.size());
    std::vector<VkWriteDescriptorSet> writes(bindings.size());
    for (size_t i = 0; i < bindings.size(); ++i) {
        infos[i].sampler = bindings[i].type == VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER ? sampler : VK_NULL_HANDLE;
        infos[i].imageView = bindings[i].image->view;
        infos[i].imageLayout = VK_IMAGE_LAYOUT_GENERAL;
        writes[i] = VkWriteDescriptorSet{};
        writes[i].sType = VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET;
        writes[i].dstSet = s.set;
        writes[i].dstBinding = bindings[i].binding;
        writes[i].descriptorCount = 1;
        writes[i].descriptorType = bindings[i].type;
        writes[i].pImageInfo = &infos[i];
    }
    vk.UpdateDescriptorSets(vk.device, (uint32_t)writes.size(), writes.data(), 0, nullptr);
    return s;
}

static VkPipeline add_pipeline(const Vk& vk, Stage& s, const std::string& spv_path, const std::vector<int32_t>& spec)
{
    std::vector<char> code = slurp(spv_path);
    code.resize((code.size() + 3) / 4 * 4);
    VkShaderModuleCreateInfo mi{};
    mi.sType = VK_STRUCTURE_TYPE_SHADER_MODULE_CREATE_INFO;
    mi.codeSize = code.size();
    mi.pCode = reinterpret_cast<const uint32_t*>(code.data());
    VkShaderModule module;
    vk_check(vk.CreateShaderModule(vk.device, &mi, nullptr, &module), "vkCreateShaderModule");
    std::vector<VkSpecializationMapEntry> entries;
    for (uint32_t i = 0; i < spec.size(); ++i) entries.push_back({i, i * 4u, 4});          // constant ids 0 .. n-1, as the reference passes them
    VkSpecializationInfo si{};
    si.mapEntryCount = (uint32_t)entries.size();
    si.pMapEntries = entries.data();
    si.dataSize = spec.size() * 4;
    si.pData = spec.data();
    VkComputePipelineCreateInfo ci{};
    ci.sType = VK_STRUCTURE_TYPE_COMPUTE_PIPELINE_CREATE_INFO;
    ci.stage.sType = VK_STRUCTURE_TYPE_PIPELINE_SHADER_STAGE_CREATE_INFO;
    ci.stage.stage = VK_SHADER_STAGE_COMPUTE_BIT;
    ci.stage.module = module;
    ci.stage.pName = "main";
    ci.stage.pSpecializationInfo = &si;
    ci.layout = s.layout;
    VkPipeline p;
    vk_check(vk.CreateComputePipelines(vk.device, VK_NULL_HANDLE, 1, &ci, nullptr, &p), "vkCreateComputePipelines");
    s.pipelines.push_back(p);
    s.modules.push_back(module);
    return p;
}

// Accumulator.hpp:28-33 (std430-compatible: 3 mat4, vec4, int)
struct AccumulatorPush { float view[16], inv_view[16], prev_view[16], prev_pos[4]; int32_t frame_number; };
// PipelineStructs.hpp:6-13
struct RayTracingPush { float view_inverse[16], proj_inverse[16], prev_view[16]; uint32_t frame_number, sample_number; };
static_assert(sizeof(AccumulatorPush) == 212 && sizeof(RayTracingPush) == 200, "push constant layouts");

// vsg::inverse(const mat4&) (external/vsg/src/vsg/maths/maths_transform.cpp:36-156), which Accumulator.cpp:100 calls for
// inverse(prev.view)[3]: affine matrices take t_inverse_4x3, the rest t_inverse_4x4; expressions in the source's order
#define M_(c, r) m[4 * (c) + (r)]
static void vsg_inverse(const float* m, float* o)
{
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (M_(0, 3) == 0.0f && M_(1, 3) == 0.0f && M_(2, 3) == 0.0f && M_(3, 3) == 1.0f) {
        const float det = (M_(0, 0) * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1)) - M_(0, 1) * (M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0))) +
                          M_(0, 2) * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
        if (det == 0.0f) { for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? nan : 0.0f; return; }
        const float A1223 = M_(2, 1) * M_(3, 2) - M_(2, 2) * M_(3, 1), A0223 = M_(2, 0) * M_(3, 2) - M_(2, 2) * M_(3, 0);
        const float A0123 = M_(2, 0) * M_(3, 1) - M_(2, 1) * M_(3, 0), A1213 = M_(1, 1) * M_(3, 2) - M_(1, 2) * M_(3, 1);
        const float A0213 = M_(1, 0) * M_(3, 2) - M_(1, 2) * M_(3, 0), A0113 = M_(1, 0) * M_(3, 1) - M_(1, 1) * M_(3, 0);
        const float id = 1.0f / det;
        o[0] = id * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1));
        o[1] = id * (M_(0, 2) * M_(2, 1) - M_(0, 1) * M_(2, 2));
        o[2] = id * (M_(0, 1) * M_(1, 2) - M_(0, 2) * M_(1, 1));
        o[3] = 0.0f;
        o[4] = id * (M_(1, 2) * M_(2, 0) - M_(1, 0) * M_(2, 2));
        o[5] = id * (M_(0, 0) * M_(2, 2) - M_(0, 2) * M_(2, 0));
        o[6] = id * (M_(0, 2) * M_(1, 0) - M_(0, 0) * M_(1, 2));
        o[7] = 0.0f;
        o[8] = id * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
        o[9] = id * (M_(0, 1) * M_(2, 0) - M_(0, 0) * M_(2, 1));
        o[10] = id * (M_(0, 0) * M_(1, 1) - M_(0, 1) * M_(1, 0));
        o[11] = 0.0f;
        o[12] = id * ((M_(1, 1) * A0223 - M_(1, 2) * A0123) - M_(1, 0) * A1223);
        o[13] = id * ((M_(0, 0) * A1223 - M_(0, 1) * A0223) + M_(0, 2) * A0123);
        o[14] = id * ((M_(0, 1) * A0213 - M_(0, 2) * A0113) - M_(0, 0) * A1213);
        o[15] = 1.0f;
        return;
    }
    const float A2323 = M_(2, 2) * M_(3, 3) - M_(2, 3) * M_(3, 2), A1323 = M_(2, 1) * M_(3, 3) - M_(2, 3) * M_(3, 1);
    const float A1223 = M_(2, 1) * M_(3, 2) - M_(2, 2) * M_(3, 1), A0323 = M_(2, 0) * M_(3, 3) - M_(2, 3) * M_(3, 0);
    const float A0223 = M_(2, 0) * M_(3, 2) - M_(2, 2) * M_(3, 0), A0123 = M_(2, 0) * M_(3, 1) - M_(2, 1) * M_(3, 0);
    const float A2313 = M_(1, 2) * M_(3, 3) - M_(1, 3) * M_(3, 2), A1313 = M_(1, 1) * M_(3, 3) - M_(1, 3) * M_(3, 1);
    const float A1213 = M_(1, 1) * M_(3, 2) - M_(1, 2) * M_(3, 1), A2312 = M_(1, 2) * M_(2, 3) - M_(1, 3) * M_(2, 2);
    const float A1312 = M_(1, 1) * M_(2, 3) - M_(1, 3) * M_(2, 1), A1212 = M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1);
    const float A0313 = M_(1, 0) * M_(3, 3) - M_(1, 3) * M_(3, 0), A0213 = M_(1, 0) * M_(3, 2) - M_(1, 2) * M_(3, 0);
    const float A0312 = M_(1, 0) * M_(2, 3) - M_(1, 3) * M_(2, 0), A0212 = M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0);
    const float A0113 = M_(1, 0) * M_(3, 1) - M_(1, 1) * M_(3, 0), A0112 = M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0);
    const float det = ((M_(0, 0) * ((M_(1, 1) * A2323 - M_(1, 2) * A1323) + M_(1, 3) * A1223) - M_(0, 1) * ((M_(1, 0) * A2323 - M_(1, 2) * A0323) + M_(1, 3) * A0223)) +
                       M_(0, 2) * ((M_(1, 0) * A1323 - M_(1, 1) * A0323) + M_(1, 3) * A0123)) -
                      M_(0, 3) * ((M_(1, 0) * A1223 - M_(1, 1) * A0223) + M_(1, 2) * A0123);
    if (det == 0.0f) { for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? nan : 0.0f; return; }
    const float id = 1.0f / det;
    o[0] = id * ((M_(1, 1) * A2323 - M_(1, 2) * A1323) + M_(1, 3) * A1223);
    o[1] = id * -((M_(0, 1) * A2323 - M_(0, 2) * A1323) + M_(0, 3) * A1223);
    o[2] = id * ((M_(0, 1) * A2313 - M_(0, 2) * A1313) + M_(0, 3) * A1213);
    o[3] = id * -((M_(0, 1) * A2312 - M_(0, 2) * A1312) + M_(0, 3) * A1212);
    o[4] = id * -((M_(1, 0) * A2323 - M_(1, 2) * A0323) + M_(1, 3) * A0223);
    o[5] = id * ((M_(0, 0) * A2323 - M_(0, 2) * A0323) + M_(0, 3) * A0223);
    o[6] = id * -((M_(0, 0) * A2313 - M_(0, 2) * A0313) + M_(0, 3) * A0213);
    o[7] = id * ((M_(0, 0) * A2312 - M_(0, 2) * A0312) + M_(0, 3) * A0212);
    o[8] = id * ((M_(1, 0) * A1323 - M_(1, 1) * A0323) + M_(1, 3) * A0123);
    o[9] = id * -((M_(0, 0) * A1323 - M_(0, 1) * A0323) + M_(0, 3) * A0123);
    o[10] = id * ((M_(0, 0) * A1313 - M_(0, 1) * A0313) + M_(0, 3) * A0113);
    o[11] = id * -((M_(0, 0) * A1312 - M_(0, 1) * A0312) + M_(0, 3) * A0112);
    o[12] = id * -((M_(1, 0) * A1223 - M_(1, 1) * A0223) + M_(1, 2) * A0123);
    o[13] = id * ((M_(0, 0) * A1223 - M_(0, 1) * A0223) + M_(0, 2) * A0123);
    o[14] = id * -((M_(0, 0) * A1213 - M_(0, 1) * A0213) + M_(0, 2) * A0113);
    o[15] = id * ((M_(0, 0) * A1212 - M_(0, 1) * A0212) + M_(0, 2) * A0112);
}
#undef M_

static void barrier(const Vk& vk, VkCommandBuffer cb, VkPipelineStageFlags src, VkPipelineStageFlags dst)
{
    VkMemoryBarrier b{};
    b.sType = VK_STRUCTURE_TYPE_MEMORY_BARRIER;
    b.srcAccessMask = VK_ACCESS_MEMORY_WRITE_BIT;
    b.dstAccessMask = VK_ACCESS_MEMORY_READ_BIT | VK_ACCESS_MEMORY_WRITE_BIT;
    vk.CmdPipelineBarrier(cb, src, dst, 0, 1, &b, 0, nullptr, 0, nullptr);
}
static const VkPipelineStageFlags COMPUTE = VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT, TRANSFER = VK_PIPELINE_STAGE_TRANSFER_BIT;

static VkBufferImageCopy whole(const Image& im, VkDeviceSize offset)
{
    VkBufferImageCopy r{};
    r.bufferOffset = offset;
    r.imageSubresource = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, im.layers};
    r.imageExtent = {im.width, im.height, 1};
    return r;
}

int main(int argc, char** argv)
{
    if (argc < 11) {
        std::cerr << "usage: vk_oracle spv_dir frames_dir out_dir width height first_frame frames bmfr|bfr|bmfrx3|bfrx3 block taa\n";
        return 2;
    }
    const std::string spv = argv[1], in = argv[2], out = argv[3];
    const uint32_t W = atoi(argv[4]), H = atoi(argv[5]);
    const int first = atoi(argv[6]), frames = atoi(argv[7]);
    const std::string kind = argv[8];
    const bool bmfr = kind.rfind("bmfr", 0) == 0, x3 = kind.size() > 2 && kind.compare(kind.size() - 2, 2, "x3") == 0;
    const std::vector<uint32_t> blocks = x3 ? std::vector<uint32_t>{8, 16, 32} : std::vector<uint32_t>{(uint32_t)atoi(argv[9])};
    const bool use_taa = atoi(argv[10]) != 0;
    try {
        Vk vk;
        init_vulkan(vk);
        // ---- images ----
        Image depth = make_image(vk, VK_FORMAT_R32_SFLOAT, W, H), normal = make_image(vk, VK_FORMAT_R32G32_SFLOAT, W, H),          // GBuffer.cpp:60, :77
              material = make_image(vk, VK_FORMAT_R8G8B8A8_UNORM, W, H), albedo = make_image(vk, VK_FORMAT_R8G8B8A8_UNORM, W, H);  // :94, :111
        Image raw = make_image(vk, VK_FORMAT_R32G32B32A32_SFLOAT, W, H);                                                          // IlluminationBuffer.cpp:260-282
        Image illum = make_image(vk, VK_FORMAT_R16G16B16A16_SFLOAT, W, H), illum_sq = make_image(vk, VK_FORMAT_R16G16B16A16_SFLOAT, W, H);   // :223-258
        Image spp = make_image(vk, VK_FORMAT_R8_UNORM, W, H), prev_spp = make_image(vk, VK_FORMAT_R8_UNORM, W, H),                 // AccumulationBuffer.cpp:251, :264
              prev_depth = make_image(vk, VK_FORMAT_R32_SFLOAT, W, H), prev_normal = make_image(vk, VK_FORMAT_R32G32_SFLOAT, W, H),  // :277, :290
              motion = make_image(vk, VK_FORMAT_R16G16_SFLOAT, W, H),                                                              // :303
              prev_illu = make_image(vk, VK_FORMAT_R16G16B16A16_SFLOAT, W, H), prev_illu_sq = make_image(vk, VK_FORMAT_R16G16B16A16_SFLOAT, W, H);   // :316, :329
        struct Denoiser {
            uint32_t B, bx, by;
            Image denoised, final_image, features, weights;
            Stage stage;
            VkPipeline pre = VK_NULL_HANDLE, fit = VK_NULL_HANDLE, post = VK_NULL_HANDLE, bfr = VK_NULL_HANDLE;
        };
        std::vector<std::unique_ptr<Denoiser>> dens;
        for (uint32_t B : blocks) {
            auto d = std::make_unique<Denoiser>();
            d->B = B; d->bx = W / B + 2; d->by = H / B + 2;                                                                        // BMFR.cpp:12-13, BFR.cpp:134
            d->denoised = make_image(vk, VK_FORMAT_R16G16B16A16_SFLOAT, W, H, 2, true);                                            // BMFR.cpp:56-75
            d->final_image = make_image(vk, VK_FORMAT_B8G8R8A8_UNORM, W, H);                                                       // BMFR.cpp:78-93
            d->features = make_image(vk, VK_FORMAT_R16_SFLOAT, d->bx * B, d->by * B, 13, true);                                    // BMFR.cpp:96-113
            d->weights = make_image(vk, VK_FORMAT_R32_SFLOAT, d->bx, d->by, 30, true);                                             // BMFR.cpp:116-133
            dens.push_back(std::move(d));
        }
        Image blend_final = make_image(vk, VK_FORMAT_B8G8R8A8_UNORM, W, H);                                                        // BFRBlender.cpp:21-37
        Image taa_final = make_image(vk, VK_FORMAT_B8G8R8A8_UNORM, W, H), taa_accumulation = make_image(vk, VK_FORMAT_R8G8B8A8_UNORM, W, H);   // Taa.cpp:42, :24
        std::vector<Image*> all = {&depth, &normal, &material, &albedo, &raw, &illum, &illum_sq, &spp, &prev_spp, &prev_depth, &prev_normal, &motion,
                                   &prev_illu, &prev_illu_sq, &blend_final, &taa_final, &taa_accumulation};
        for (auto& d : dens) { all.push_back(&d->denoised); all.push_back(&d->final_image); all.push_back(&d->features); all.push_back(&d->weights); }
        const Image& denoiser_final = x3 ? blend_final : dens[0]->final_image;

        VkSamplerCreateInfo sci{};                            // vsg::Sampler::create() defaults
        sci.sType = VK_STRUCTURE_TYPE_SAMPLER_CREATE_INFO;
        sci.magFilter = sci.minFilter = VK_FILTER_LINEAR;
        sci.mipmapMode = VK_SAMPLER_MIPMAP_MODE_LINEAR;
        sci.addressModeU = sci.addressModeV = sci.addressModeW = VK_SAMPLER_ADDRESS_MODE_REPEAT;
        sci.maxLod = VK_LOD_CLAMP_NONE;
        sci.unnormalizedCoordinates = VK_FALSE;
        VkSampler sampler;
        vk_check(vk.CreateSampler(vk.device, &sci, nullptr, &sampler), "vkCreateSampler");

        VkDescriptorPoolSize sizes[2] = {{VK_DESCRIPTOR_TYPE_STORAGE_IMAGE, 128}, {VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER, 64}};
        VkDescriptorPoolCreateInfo dpi{};
        dpi.sType = VK_STRUCTURE_TYPE_DESCRIPTOR_POOL_CREATE_INFO;
        dpi.maxSets = 16;
        dpi.poolSizeCount = 2;
        dpi.pPoolSizes = sizes;
        VkDescriptorPool pool;
        vk_check(vk.CreateDescriptorPool(vk.device, &dpi, nullptr, &pool), "vkCreateDescriptorPool");

        // ---- pipelines ----
        const VkDescriptorType SI = VK_DESCRIPTOR_TYPE_STORAGE_IMAGE, CIS = VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER;
        // accumulator.comp:5-18
        Stage acc = make_stage(vk, pool, sampler, {{0, CIS, &raw}, {1, SI, &depth}, {2, SI, &normal}, {3, SI, &material}, {4, SI, &albedo}, {5, CIS, &prev_depth},
                                                   {6, CIS, &prev_normal}, {7, SI, &motion}, {8, SI, &spp}, {9, CIS, &prev_spp}, {10, CIS, &prev_illu},
                                                   {11, SI, &illum}, {12, CIS, &prev_illu_sq}, {13, SI, &illum_sq}}, sizeof(AccumulatorPush));
        const VkPipeline p_acc = add_pipeline(vk, acc, spv + "/accumulator_sep.comp.spv", {16, 16});                 // Accumulator.cpp:21-24, work size 16 x 16 (Accumulator.hpp:17)
        // bmfrGeneral.comp:3-14 / bfr.comp:5-14
        for (auto& d : dens) {
            const int32_t B = (int32_t)d->B;
            std::vector<Binding> b = {{0, SI, &depth}, {1, SI, &normal}, {2, SI, &material}, {3, SI, &albedo}, {4, SI, &motion}, {5, SI, &spp},
                                      {6, CIS, &d->denoised}, {7, SI, &d->final_image}, {8, CIS, &illum}, {9, SI, &d->denoised}};
            if (bmfr) { b.push_back({10, SI, &d->features}); b.push_back({11, SI, &d->weights}); }
            d->stage = make_stage(vk, pool, sampler, b, sizeof(RayTracingPush));
            if (bmfr) {
                const int32_t T = B == 8 ? 64 : 256;                                                                  // DenoiserUtils.cpp:78-95 (fitting_kernel)
                d->pre = add_pipeline(vk, d->stage, spv + "/bmfrPre.comp.spv", {(int32_t)W, (int32_t)H, B, B, B});    // BMFR.cpp:32-38
                d->fit = add_pipeline(vk, d->stage, spv + "/bmfrFit.comp.spv", {(int32_t)W, (int32_t)H, T, 1, B});    // :40-46
                d->post = add_pipeline(vk, d->stage, spv + "/bmfrPost.comp.spv", {(int32_t)W, (int32_t)H, B, B, B});  // :48-54
            } else {
                d->bfr = add_pipeline(vk, d->stage, spv + "/bfr.comp.spv", {(int32_t)W, (int32_t)H, B, B});           // BFR.cpp:26-31
            }
        }
        // bfrBlender.comp:4-9; BFRBlender(width, height, illumination_images[0], [1], den8, den16, den32) (DenoiserUtils.cpp:53-55)
        Stage blend;
        VkPipeline p_blend = VK_NULL_HANDLE;
        if (x3) {
            blend = make_stage(vk, pool, sampler, {{0, SI, &illum}, {1, SI, &illum_sq}, {2, SI, &dens[0]->final_image}, {3, SI, &dens[1]->final_image},
                                                   {4, SI, &dens[2]->final_image}, {5, SI, &blend_final}}, 0);
            p_blend = add_pipeline(vk, blend, spv + "/bfrBlender.comp.spv", {(int32_t)W, (int32_t)H, 16, 16, 2});     // BFRBlender.cpp:13-19 (work_height, work_width, radius)
        }
        // taa.comp:8-11
        Stage taa;
        VkPipeline p_taa = VK_NULL_HANDLE;
        if (use_taa) {
            taa = make_stage(vk, pool, sampler, {{0, SI, &motion}, {1, CIS, &denoiser_final}, {2, SI, &taa_final}, {3, CIS, &taa_accumulation}}, sizeof(RayTracingPush));
            p_taa = add_pipeline(vk, taa, spv + "/taa.comp.spv", {(int32_t)W, (int32_t)H, 16, 16});                   // Taa.cpp:14-19, VulkanPBRT.cpp:450
        }

        // ---- staging, command buffer ----
        const VkDeviceSize px = (VkDeviceSize)W * H;
        const VkDeviceSize up_depth = 0, up_normal = up_depth + px * 4, up_albedo = up_normal + px * 8, up_illum = up_albedo + px * 4, up_avgsq = up_illum + px * 16, up_end = up_avgsq + px * 8;
        HostBuffer upload = make_host_buffer(vk, up_end);
        const Image& shown = use_taa ? taa_final : denoiser_final;
        const VkDeviceSize den_bytes = dens[0]->denoised.bytes();
        const VkDeviceSize rb_final = 0, rb_den = rb_final + px * 4, rb_motion = rb_den + den_bytes * dens.size(), rb_spp = rb_motion + px * 4, rb_illum = rb_spp + px,
                           rb_end = rb_illum + px * 8;
        HostBuffer readback = make_host_buffer(vk, (rb_end + 15) / 16 * 16);
        VkCommandPoolCreateInfo cpi{};
        cpi.sType = VK_STRUCTURE_TYPE_COMMAND_POOL_CREATE_INFO;
        cpi.flags = VK_COMMAND_POOL_CREATE_RESET_COMMAND_BUFFER_BIT;
        cpi.queueFamilyIndex = vk.queue_family;
        VkCommandPool cpool;
        vk_check(vk.CreateCommandPool(vk.device, &cpi, nullptr, &cpool), "vkCreateCommandPool");
        VkCommandBufferAllocateInfo cai{};
        cai.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_ALLOCATE_INFO;
        cai.commandPool = cpool;
        cai.level = VK_COMMAND_BUFFER_LEVEL_PRIMARY;
        cai.commandBufferCount = 1;
        VkCommandBuffer cb;
        vk_check(vk.AllocateCommandBuffers(vk.device, &cai, &cb), "vkAllocateCommandBuffers");
        VkCommandBufferBeginInfo bi{};
        bi.sType = VK_STRUCTURE_TYPE_COMMAND_BUFFER_BEGIN_INFO;
        bi.flags = VK_COMMAND_BUFFER_USAGE_ONE_TIME_SUBMIT_BIT;
        auto submit = [&] {
            vk_check(vk.EndCommandBuffer(cb), "vkEndCommandBuffer");
            VkSubmitInfo si{};
            si.sType = VK_STRUCTURE_TYPE_SUBMIT_INFO;
            si.commandBufferCount = 1;
            si.pCommandBuffers = &cb;
            vk_check(vk.QueueSubmit(vk.queue, 1, &si, VK_NULL_HANDLE), "vkQueueSubmit");
            vk_check(vk.QueueWaitIdle(vk.queue), "vkQueueWaitIdle");
        };

        // update_image_layouts (e.g. BMFR.cpp:180-201): UNDEFINED -> GENERAL; plus a clear to zero -- the reference leaves new
        // images undefined, the oracle (and the product) define the initial history as zero
        vk_check(vk.BeginCommandBuffer(cb, &bi), "vkBeginCommandBuffer");
        for (Image* im : all) {
            VkImageMemoryBarrier b{};
            b.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
            b.dstAccessMask = VK_ACCESS_TRANSFER_WRITE_BIT;
            b.oldLayout = VK_IMAGE_LAYOUT_UNDEFINED;
            b.newLayout = VK_IMAGE_LAYOUT_GENERAL;
            b.srcQueueFamilyIndex = b.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
            b.image = im->image;
            b.subresourceRange = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, im->layers};
            vk.CmdPipelineBarrier(cb, VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, TRANSFER, 0, 0, nullptr, 0, nullptr, 1, &b);
            const VkClearColorValue zero{};
            vk.CmdClearColorImage(cb, im->image, VK_IMAGE_LAYOUT_GENERAL, &zero, 1, &b.subresourceRange);
        }
        barrier(vk, cb, TRANSFER, COMPUTE | TRANSFER);
        submit();

        AccumulatorPush apc{};
        RayTracingPush rpc{};
        for (int i = 0; i < 4; ++i) rpc.prev_view[i * 5] = 1.f;          // RayTracingPushConstants' default-constructed (identity) prev_view
        for (int f = first; f < first + frames; ++f) {
            const std::string base = in + "/frame_" + std::to_string(f);
            const auto d = slurp(base + ".depth"), n = slurp(base + ".normal"), a = slurp(base + ".albedo"), il = slurp(base + ".illum"), cam = slurp(base + ".cam");
            if (d.size() != px * 4 || n.size() != px * 8 || a.size() != px * 4 || il.size() != px * 16 || cam.size() != 256) throw std::runtime_error("bad frame file sizes: " + base);
            const float* view = reinterpret_cast<const float*>(cam.data());
            const float *inv_view = view + 16, *inv_proj = view + 48;
            memcpy(upload.map + up_depth, d.data(), d.size());
            memcpy(upload.map + up_normal, n.data(), n.size());
            memcpy(upload.map + up_albedo, a.data(), a.size());
            memcpy(upload.map + up_illum, il.data(), il.size());
            // Accumulator::set_camera_matrices, separate matrices (Accumulator.cpp:87-103): "view" carries the inverse projection
            memcpy(apc.view, inv_proj, 64);
            memcpy(apc.inv_view, inv_view, 64);
            if (f != 0) {
                memcpy(apc.prev_view, rpc.prev_view, 64);                // b.view = pc.prev_view (VulkanPBRT.cpp:578-584)
                float inv[16];
                vsg_inverse(rpc.prev_view, inv);
                apc.prev_pos[0] = inv[12]; apc.prev_pos[1] = inv[13]; apc.prev_pos[2] = inv[14];
                apc.prev_pos[3] = 1.f;
            }
            apc.frame_number = f;
            memcpy(rpc.view_inverse, inv_view, 64);                      // VulkanPBRT.cpp:561-563
            memcpy(rpc.proj_inverse, inv_proj, 64);
            rpc.frame_number = (uint32_t)f;
            rpc.sample_number = 0;

            vk_check(vk.BeginCommandBuffer(cb, &bi), "vkBeginCommandBuffer");
            // the producer: G-buffer and 1-spp illumination arrive (VulkanPBRT.cpp:568-569 in offline mode)
            struct Up { const Image* im; VkDeviceSize off; } ups[4] = {{&depth, up_depth}, {&normal, up_normal}, {&albedo, up_albedo}, {&raw, up_illum}};
            for (const Up& u : ups) {
                const VkBufferImageCopy r = whole(*u.im, u.off);
                vk.CmdCopyBufferToImage(cb, upload.buffer, u.im->image, VK_IMAGE_LAYOUT_GENERAL, 1, &r);
            }
            {
                std::ifstream sq(base + ".avgsq", std::ios::binary);
                if (sq) {
                    sq.read(upload.map + up_avgsq, (std::streamsize)(px * 8));
                    if ((VkDeviceSize)sq.gcount() != px * 8) throw std::runtime_error("bad size: " + base + ".avgsq");
                    const VkBufferImageCopy r = whole(illum_sq, up_avgsq);
                    vk.CmdCopyBufferToImage(cb, upload.buffer, illum_sq.image, VK_IMAGE_LAYOUT_GENERAL, 1, &r);
                }
            }
            barrier(vk, cb, TRANSFER, COMPUTE);
            // Accumulator::add_dispatch_to_command_graph (Accumulator.cpp:72-83)
            vk.CmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_COMPUTE, p_acc);
            vk.CmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_COMPUTE, acc.layout, 0, 1, &acc.set, 0, nullptr);
            vk.CmdPushConstants(cb, acc.layout, VK_SHADER_STAGE_COMPUTE_BIT, 0, sizeof(apc), &apc);
            vk.CmdDispatch(cb, (uint32_t)std::ceil((float)W / 16.f), (uint32_t)std::ceil((float)H / 16.f), 1);
            barrier(vk, cb, COMPUTE, COMPUTE);
            for (auto& d : dens) {
                if (bmfr) {                                               // BMFR.cpp:203-230: pre, fit, post with W_padded / b groups
                    for (VkPipeline p : {d->pre, d->fit, d->post}) {
                        vk.CmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_COMPUTE, p);
                        vk.CmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_COMPUTE, d->stage.layout, 0, 1, &d->stage.set, 0, nullptr);
                        vk.CmdPushConstants(cb, d->stage.layout, VK_SHADER_STAGE_COMPUTE_BIT, 0, sizeof(rpc), &rpc);
                        vk.CmdDispatch(cb, d->bx, d->by, 1);
                        barrier(vk, cb, COMPUTE, COMPUTE);
                    }
                } else {                                                  // BFR.cpp:128-138
                    vk.CmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_COMPUTE, d->bfr);
                    vk.CmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_COMPUTE, d->stage.layout, 0, 1, &d->stage.set, 0, nullptr);
                    vk.CmdPushConstants(cb, d->stage.layout, VK_SHADER_STAGE_COMPUTE_BIT, 0, sizeof(rpc), &rpc);
                    vk.CmdDispatch(cb, d->bx, d->by, 1);
                    barrier(vk, cb, COMPUTE, COMPUTE);
                }
            }
            if (x3) {                                                     // BFRBlender.cpp:80-90
                vk.CmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_COMPUTE, p_blend);
                vk.CmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_COMPUTE, blend.layout, 0, 1, &blend.set, 0, nullptr);
                vk.CmdDispatch(cb, (uint32_t)std::ceil((float)W / 16.f), (uint32_t)std::ceil((float)H / 16.f), 1);
                barrier(vk, cb, COMPUTE, COMPUTE);
            }
            if (use_taa) {
                // Taa.cpp:99-107.  The reference pushes nothing here: taa.comp reads the RayTracingPushConstants the denoiser
                // pushed, which stay valid because the pipeline layouts are push-constant compatible.  Pushed again here.
                vk.CmdBindPipeline(cb, VK_PIPELINE_BIND_POINT_COMPUTE, p_taa);
                vk.CmdBindDescriptorSets(cb, VK_PIPELINE_BIND_POINT_COMPUTE, taa.layout, 0, 1, &taa.set, 0, nullptr);
                vk.CmdPushConstants(cb, taa.layout, VK_SHADER_STAGE_COMPUTE_BIT, 0, sizeof(rpc), &rpc);
                vk.CmdDispatch(cb, (uint32_t)std::ceil((float)W / 16.f), (uint32_t)std::ceil((float)H / 16.f), 1);
                barrier(vk, cb, COMPUTE, TRANSFER);
                VkImageCopy c{};                                          // Taa.cpp:106: raw copy BGRA8 -> RGBA8
                c.srcSubresource = c.dstSubresource = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1};
                c.extent = {W, H, 1};
                vk.CmdCopyImage(cb, taa_final.image, VK_IMAGE_LAYOUT_GENERAL, taa_accumulation.image, VK_IMAGE_LAYOUT_GENERAL, 1, &c);
            }
            barrier(vk, cb, COMPUTE | TRANSFER, TRANSFER);
            // results of the frame, before the history rotates
            {
                VkBufferImageCopy r = whole(shown, rb_final);
                vk.CmdCopyImageToBuffer(cb, shown.image, VK_IMAGE_LAYOUT_GENERAL, readback.buffer, 1, &r);
                for (size_t i = 0; i < dens.size(); ++i) {
                    r = whole(dens[i]->denoised, rb_den + den_bytes * i);
                    vk.CmdCopyImageToBuffer(cb, dens[i]->denoised.image, VK_IMAGE_LAYOUT_GENERAL, readback.buffer, 1, &r);
                }
                r = whole(motion, rb_motion);
                vk.CmdCopyImageToBuffer(cb, motion.image, VK_IMAGE_LAYOUT_GENERAL, readback.buffer, 1, &r);
                r = whole(spp, rb_spp);
                vk.CmdCopyImageToBuffer(cb, spp.image, VK_IMAGE_LAYOUT_GENERAL, readback.buffer, 1, &r);
                r = whole(illum, rb_illum);
                vk.CmdCopyImageToBuffer(cb, illum.image, VK_IMAGE_LAYOUT_GENERAL, readback.buffer, 1, &r);
            }
            // AccumulationBuffer::copy_to_back_images (AccumulationBuffer.cpp:72-244)
            struct Back { const Image* src; const Image* dst; } backs[5] = {{&spp, &prev_spp}, {&depth, &prev_depth}, {&normal, &prev_normal}, {&illum, &prev_illu},
                                                                           {&illum_sq, &prev_illu_sq}};
            for (const Back& bk : backs) {
                VkImageCopy c{};
                c.srcSubresource = c.dstSubresource = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1};
                c.extent = {W, H, 1};
                vk.CmdCopyImage(cb, bk.src->image, VK_IMAGE_LAYOUT_GENERAL, bk.dst->image, VK_IMAGE_LAYOUT_GENERAL, 1, &c);
            }
            barrier(vk, cb, TRANSFER, COMPUTE | TRANSFER | VK_PIPELINE_STAGE_HOST_BIT);
            submit();
            memcpy(rpc.prev_view, view, 64);                             // VulkanPBRT.cpp:591

            auto dump = [&](const char* name, const char* ext, VkDeviceSize off, VkDeviceSize bytes) {
                std::ofstream(out + "/" + name + "_" + std::to_string(f) + "." + ext, std::ios::binary).write(readback.map + off, (std::streamsize)bytes);
            };
            dump("final", "bgra", rb_final, px * 4);
            for (size_t i = 0; i < dens.size(); ++i) dump(("denoised" + std::to_string(dens[i]->B)).c_str(), "rgba16f", rb_den + den_bytes * i, den_bytes);
            dump("motion", "rg16f", rb_motion, px * 4);
            dump("spp", "r8", rb_spp, px);
            dump("illum", "rgba16f", rb_illum, px * 8);
        }

        // ---- teardown ----
        std::vector<Stage*> stages = {&acc, &blend, &taa};
        for (auto& d : dens) stages.push_back(&d->stage);
        for (Stage* s : stages) {
            for (VkPipeline p : s->pipelines) vk.DestroyPipeline(vk.device, p, nullptr);
            for (VkShaderModule m : s->modules) vk.DestroyShaderModule(vk.device, m, nullptr);
            if (s->layout) vk.DestroyPipelineLayout(vk.device, s->layout, nullptr);
            if (s->set_layout) vk.DestroyDescriptorSetLayout(vk.device, s->set_layout, nullptr);
        }
        vk.DestroyDescriptorPool(vk.device, pool, nullptr);
        vk.DestroySampler(vk.device, sampler, nullptr);
        vk.DestroyCommandPool(vk.device, cpool, nullptr);
        for (HostBuffer* b : {&upload, &readback}) { vk.DestroyBuffer(vk.device, b->buffer, nullptr); vk.FreeMemory(vk.device, b->memory, nullptr); }
        for (Image* im : all) { vk.DestroyImageView(vk.device, im->view, nullptr); vk.DestroyImage(vk.device, im->image, nullptr); vk.FreeMemory(vk.device, im->memory, nullptr); }
        vk.DestroyDevice(vk.device, nullptr);
        reinterpret_cast<PFN_vkDestroyInstance>(vk.gipa(vk.instance, "vkDestroyInstance"))(vk.instance, nullptr);
    } catch (const std::exception& e) {
        std::cerr << "vk_oracle: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
