#!/usr/bin/env python
"""Wraps the demodulation statements of the reference's ray-generation shader in a compute shader (TEST INFRASTRUCTURE).

shaders/ptRaygen.rgen cannot go through the shim as a whole (ray tracing pipeline), but what the denoising path
depends on is three statements of it: the radiance clamp and the DEMOD_ILLUMINATION_FLOAT block
(`finalColor = clamp(...)` ... `imageStore(illumination, ...)`).  This script cuts exactly those statements out of the
reference's text, where it lies under the reference tree, and puts them into a main() that supplies the names they use
(finalColor, curAlbedo, rayPayload.position, gl_LaunchIDEXT) from three input images.  The result is written under
oracle/_ref/gen/ (git-ignored) and deleted after compilation; only the wrapper below is this repository's text.

    extract_rgen.py <shaders_dir> <out.comp>
"""
import re
import sys
from pathlib import Path

WRAPPER = """#version 460
#include "ptConstants.glsl"
layout(binding = 0, rgba32f) uniform image2D radianceIn;
layout(binding = 1, rgba32f) uniform image2D albedoIn;
layout(binding = 2, r32f) uniform image2D positionXIn;
layout(binding = 3, rgba32f) uniform image2D illumination;
layout(constant_id = 2) const int IMAGE_WIDTH = 1;
layout(constant_id = 3) const int IMAGE_HEIGHT = 1;
layout (local_size_x_id = 0,local_size_y_id = 1,local_size_z=1) in;
struct RayPayloadStandIn { vec3 position; };
void main(){
    if(gl_GlobalInvocationID.x >= IMAGE_WIDTH || gl_GlobalInvocationID.y >= IMAGE_HEIGHT) return;
    uvec3 gl_LaunchIDEXT = gl_GlobalInvocationID;
    RayPayloadStandIn rayPayload;
    rayPayload.position = vec3(imageLoad(positionXIn, ivec2(gl_LaunchIDEXT.xy)).x, 0, 0);
    vec3 finalColor = imageLoad(radianceIn, ivec2(gl_LaunchIDEXT.xy)).xyz;
    vec3 curAlbedo = imageLoad(albedoIn, ivec2(gl_LaunchIDEXT.xy)).xyz;
@REFERENCE_STATEMENTS@
}
"""


def main():
    shaders, out = Path(sys.argv[1]), Path(sys.argv[2])
    text = (shaders / "ptRaygen.rgen").read_text()
    clamp = re.search(r'^[ \t]*finalColor\s*=\s*clamp\(finalColor[^\n]*;[ \t]*$', text, flags=re.M)
    block = re.search(r'^#if defined DEMOD_ILLUMINATION_FLOAT[ \t]*\n.*?^#endif[ \t]*$', text, flags=re.M | re.S)
    if not clamp or not block or block.start() < clamp.end():
        sys.exit("extract_rgen.py: ptRaygen.rgen does not look like the reference's (clamp / DEMOD_ILLUMINATION_FLOAT block not found)")
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(WRAPPER.replace("@REFERENCE_STATEMENTS@", clamp.group(0) + "\n" + block.group(0)))


if __name__ == "__main__":
    main()
