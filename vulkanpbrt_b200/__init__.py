"""vulkanpbrt_b200 -- B200-native (sm_100a) drop-in for VulkanPBRT's denoising hot path:
Accumulator, BMFR, BFR, BFRBlender, Taa.  The kernels live in csrc/ behind the C ABI of
include/vkpbrt_b200.h; this package is the thin host layer that mirrors the reference's
render-module classes.  There is no CPU fallback."""
from ._capi import VkpbrtError  # noqa: F401
from .modules import (BFR, BMFR, Accumulator, AccumulationBuffer, BFRBlender, CameraMatrices, Commands, Context,  # noqa: F401
                      DenoisingBlockSize, DenoisingType, DescriptorImage, GBuffer, IlluminationBuffer,
                      IlluminationBufferDemodulated, IlluminationBufferDemodulatedFloat, IlluminationBufferFinal,
                      IlluminationBufferFinalDemodulated, PushConstants, Taa, add_denoiser_to_commands)
from .matrix_io import export_matrices, import_matrices  # noqa: F401
from .modules import FormatConverter, demodulate  # noqa: F401
from .pipeline import DenoisePipeline  # noqa: F401
from .render_io import GBufferIO, IlluminationBufferIO, OfflineGBuffer, OfflineIllumination  # noqa: F401
