"""Offline sequence I/O: the reference's way of feeding pre-rendered sequences.

Host-side mirror of `GBufferIO` and `IlluminationBufferIO` (source/io/RenderIO.cpp:6-158 import, :213-310 export,
:312-382 conversions, :501-591 illumination): printf-style per-frame file names, OpenEXR planes, and the conversions
between what is stored (world position / cartesian normal / float albedo, all rgba32f) and what the G-buffer holds
(Euclidean depth r32f, spherical normal rg32f, albedo rgba8).  The EXR codec is OpenCV's bundled OpenEXR (the
reference uses vsgXchange::openexr); planes come back in R, G, B, A channel order, row 0 at the top.

The conversions run in binary32 like the reference's loops.  `acos` / `atan2` / `sin` / `cos` are the platform's libm
in the reference and numpy's here: results agree to an ulp or two, which is below the fp16 / unorm8 storage of
everything downstream -- but it means imported planes are inputs, not something parity is asserted on; parity is
asserted on what the modules compute FROM the imported planes (tests/test_render_io.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .modules import CameraMatrices

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")     # OpenCV ships the codec disabled by default


def _cv2():
    import cv2
    return cv2


@dataclass
class OfflineGBuffer:           # source/io/RenderIO.hpp: depth / normal / material / albedo as uploaded to the GBuffer
    depth: Optional[np.ndarray] = None          # float32 [H][W]
    normal: Optional[np.ndarray] = None         # float32 [H][W][2]  (theta, phi)
    material: Optional[np.ndarray] = None       # uint8   [H][W][4]
    albedo: Optional[np.ndarray] = None         # uint8   [H][W][4]

    def download_from_g_buffer(self, g_buffer) -> "OfflineGBuffer":
        """download_from_g_buffer_command + transfer_staging_data_to (RenderIO.cpp:748-928): the GBuffer's own planes, read back
        once the frame's work on the context's stream is done (Image.download synchronises)"""
        self.depth, self.normal, self.albedo = g_buffer.depth.download(), g_buffer.normal.download(), g_buffer.albedo.download()
        self.material = g_buffer.material.download() if getattr(g_buffer, "material", None) is not None else None
        return self


@dataclass
class OfflineIllumination:
    noisy: Optional[np.ndarray] = None          # float32 [H][W][4]

    def download_from_illumination_buffer(self, illu_buffer) -> "OfflineIllumination":
        """download_from_illumination_buffer_command + transfer_staging_data_to (RenderIO.cpp:401-467): image 0"""
        self.noisy = illu_buffer.illumination_images[0].download()
        return self


# ---- EXR planes --------------------------------------------------------------------------------------------------
def read_exr(path) -> Optional[np.ndarray]:
    """float32 [H][W] or [H][W][C] with channels in R, G, B, A order; None when the file cannot be read"""
    cv2 = _cv2()
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        return None
    img = np.asarray(img, dtype=np.float32)
    if img.ndim == 3 and img.shape[2] >= 3:
        img = img[..., [2, 1, 0] + list(range(3, img.shape[2]))]        # OpenCV hands back B, G, R(, A)
    return np.ascontiguousarray(img)


def write_exr(path, plane: np.ndarray) -> bool:
    cv2 = _cv2()
    a = np.asarray(plane, dtype=np.float32)
    if a.ndim == 3 and a.shape[2] == 2:
        raise ValueError("two-channel planes are converted before they are stored (spherical_to_cartesian)")
    if a.ndim == 3 and a.shape[2] >= 3:
        a = a[..., [2, 1, 0] + list(range(3, a.shape[2]))]
    return bool(cv2.imwrite(str(path), np.ascontiguousarray(a), [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT]))


def _rgba(img: np.ndarray) -> np.ndarray:
    """vec4Array2D view of whatever the file held (3-channel files get w = 1)"""
    if img.ndim == 2:
        img = img[..., None]
    if img.shape[2] == 4:
        return img
    out = np.ones(img.shape[:2] + (4,), np.float32)
    out[..., :min(3, img.shape[2])] = img[..., :3]
    return out


# ---- conversions (RenderIO.cpp:160-211, :312-382) -------------------------------------------------------------------
class GBufferIO:
    @staticmethod
    def convert_normal_to_spherical(normals: np.ndarray) -> np.ndarray:
        """:160-178  (theta, phi) = (acos(n.z), atan2(n.y, n.x))"""
        n = _rgba(normals).astype(np.float32).astype(np.float64)      # the reference's unqualified acos / atan2 are the double routines
        with np.errstate(invalid="ignore"):
            return np.stack([np.arccos(n[..., 2]), np.arctan2(n[..., 1], n[..., 0])], axis=-1).astype(np.float32)

    @staticmethod
    def spherical_to_cartesian(normals: np.ndarray) -> np.ndarray:
        """:312-329"""
        # cos / sin are the C library's double routines in the reference's translation unit (unqualified calls on floats with
        # only <cmath> in scope); the products are rounded once
        t, p = normals[..., 0].astype(np.float64), normals[..., 1].astype(np.float64)
        out = np.empty(normals.shape[:2] + (4,), np.float32)
        out[..., 0] = np.cos(p) * np.sin(t)
        out[..., 1] = np.sin(p) * np.sin(t)
        out[..., 2] = np.cos(t)
        out[..., 3] = 1.0
        return out

    @staticmethod
    def compress_albedo(albedo: np.ndarray) -> np.ndarray:
        """:180-211  float (or half) rgba -> rgba8 by `ubvec4 = vec4 * 255.0F`: the C++ conversion TRUNCATES; integer
        inputs are taken as they are"""
        a = np.asarray(albedo)
        if a.dtype.kind in "ui":
            return _rgba(a.astype(np.float32)).astype(np.uint8) if a.ndim == 3 and a.shape[2] != 4 else a.astype(np.uint8)
        scaled = _rgba(a.astype(np.float32)) * np.float32(255.0)
        return np.clip(np.trunc(scaled), 0, 255).astype(np.uint8)      # out-of-range values are undefined in the reference

    @staticmethod
    def unorm_to_float(array: np.ndarray) -> np.ndarray:
        """:331-346"""
        return (array.astype(np.float32) / np.float32(255.0)).astype(np.float32)

    @staticmethod
    def position_to_depth(position: np.ndarray, matrices: CameraMatrices) -> np.ndarray:
        """:101-118  depth = |camera - p| with the camera position taken as column 2 of inv_view divided by its w: the
        offline matrices are combined view-projections, for which that column of the inverse is the eye point"""
        iv = np.asarray(matrices.inv_view, dtype=np.float32).reshape(4, 4)      # [col][row]
        with np.errstate(divide="ignore", invalid="ignore"):
            cam = (iv[2] * (np.float32(1.0) / iv[2][3]))[:3].astype(np.float32)      # vsg's vec4 /= multiplies by the reciprocal
        d = cam[None, None, :] - _rgba(position)[..., :3].astype(np.float32)
        return np.sqrt((d * d).sum(axis=-1, dtype=np.float32)).astype(np.float32)

    @staticmethod
    def depth_to_position(depth: np.ndarray, matrices: CameraMatrices) -> Optional[np.ndarray]:
        """:348-382  needs separate view / projection matrices"""
        if matrices.proj is None or matrices.inv_proj is None:
            print("GBufferIO::depthToPosition: Camera matrix in wrong layout. Expected camera matrix with separate projection matrix")
            return None
        f = np.float32
        H, W = depth.shape
        ip = np.asarray(matrices.inv_proj, dtype=f).reshape(4, 4)       # [col][row]
        iv = np.asarray(matrices.inv_view, dtype=f).reshape(4, 4)

        def mat_vec(m, v):      # vsg/maths/mat4.h:159-165, the sums in the source's order (binary32 throughout)
            return [((m[0][r] * v[0] + m[1][r] * v[1]) + m[2][r] * v[2]) + m[3][r] * v[3] for r in range(4)]

        x = np.broadcast_to(((np.arange(W, dtype=f) + f(.5)) / f(W) * f(2) - f(1))[None, :], (H, W))
        y = np.broadcast_to(((np.arange(H, dtype=f) + f(.5)) / f(H) * f(2) - f(1))[:, None], (H, W))
        one = np.ones((H, W), f)
        d = mat_vec(ip, [x, y, one, one])
        d[3] = np.zeros((H, W), f)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv_len = f(1.0) / np.sqrt(((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) + d[3] * d[3])      # normalize = v * (1 / length), vec4.h:225-254
            world = mat_vec(iv, [c * inv_len for c in d])
            out = np.ones((H, W, 4), f)
            for c in range(3):
                out[..., c] = iv[3][c] + world[c] * depth.astype(f)
        return out

    # ---- on the device ---------------------------------------------------------------------------------------------
    @staticmethod
    def import_to_device(g_buffer, position: Optional[np.ndarray], matrices: Optional[CameraMatrices], normal: Optional[np.ndarray],
                         albedo: Optional[np.ndarray]) -> None:
        """position_to_depth + convert_normal_to_spherical + compress_albedo of one frame as ONE device launch straight into
        a compiled GBuffer (vkpbrt_gbuffer_import_record): the decoded rgba32f planes are uploaded as they come out of the
        files and converted where the denoiser reads them.  Any plane may be None (that member is left alone)."""
        import ctypes as C

        from . import _capi as capi
        from .modules import DescriptorImage
        ctx = g_buffer.ctx
        imgs = []
        for plane in (position, normal, albedo):
            if plane is None:
                imgs.append(None)
                continue
            a = np.ascontiguousarray(_rgba(np.asarray(plane, np.float32)))
            im = DescriptorImage.create(ctx, capi.FORMAT_R32G32B32A32_SFLOAT, a.shape[1], a.shape[0])
            im.compile()
            im.upload(a)
            imgs.append(im)
        iv = capi.mat16(np.asarray(matrices.inv_view, np.float32)) if (position is not None and matrices is not None) else None
        h = lambda im: im.handle if im is not None else None
        capi.call("vkpbrt_gbuffer_import_record", g_buffer.handle, h(imgs[0]), C.cast(iv, C.c_void_p) if iv is not None else None, h(imgs[1]), h(imgs[2]))
        ctx.synchronize()        # the staging images die with this call

    # ---- files -----------------------------------------------------------------------------------------------------
    @staticmethod
    def import_g_buffer_depth(depth_format: str, normal_format: str, material_format: str, albedo_format: str, num_frames: int,
                              verbosity: int = 1) -> List[OfflineGBuffer]:
        """:6-69  (the reference loads no material plane either)"""
        out = []
        for f in range(num_frames):
            g = OfflineGBuffer()
            out.append(g)
            depth = read_exr(depth_format % f)
            if depth is None:
                print(f"Failed to load image: {depth_format % f}")
                continue
            g.depth = np.ascontiguousarray(depth[..., 0] if depth.ndim == 3 else depth)
            if not GBufferIO._load_normal_albedo(g, normal_format % f, albedo_format % f):
                continue
        return out

    @staticmethod
    def import_g_buffer_position(position_format: str, normal_format: str, material_format: str, albedo_format: str,
                                 matrices: List[CameraMatrices], num_frames: int, verbosity: int = 1) -> List[OfflineGBuffer]:
        """:71-158"""
        out = []
        for f in range(num_frames):
            g = OfflineGBuffer()
            out.append(g)
            pos = read_exr(position_format % f)
            if pos is None:
                print(f"Failed to load image: {position_format % f}")
                continue
            if pos.ndim != 3 or pos.shape[2] < 3:
                print("Unexpected position format")
                continue
            g.depth = GBufferIO.position_to_depth(pos, matrices[f])
            GBufferIO._load_normal_albedo(g, normal_format % f, albedo_format % f)
        return out

    @staticmethod
    def _load_normal_albedo(g: OfflineGBuffer, normal_path: str, albedo_path: str) -> bool:
        normal = read_exr(normal_path)
        if normal is None:
            print(f"Failed to load image: {normal_path}")
            return False
        g.normal = GBufferIO.convert_normal_to_spherical(normal)
        albedo = read_exr(albedo_path)
        if albedo is None:
            print(f"Failed to load image: {albedo_path}")
            return False
        g.albedo = GBufferIO.compress_albedo(albedo)
        g.material = np.zeros(g.albedo.shape, np.uint8)
        return True

    @staticmethod
    def export_g_buffer(position_format: str, depth_format: str, normal_format: str, material_format: str, albedo_format: str,
                        num_frames: int, g_buffers: List[OfflineGBuffer], matrices: List[CameraMatrices], verbosity: int = 1) -> bool:
        """:213-310  empty format strings skip a plane"""
        fine = True
        for f in range(num_frames):
            g = g_buffers[f]
            jobs = []
            if depth_format:
                jobs.append((depth_format % f, g.depth))
            if position_format:
                jobs.append((position_format % f, GBufferIO.depth_to_position(g.depth, matrices[f])))
            if normal_format:
                jobs.append((normal_format % f, GBufferIO.spherical_to_cartesian(g.normal)))
            if material_format:
                jobs.append((material_format % f, GBufferIO.unorm_to_float(g.material)))
            if albedo_format:
                jobs.append((albedo_format % f, GBufferIO.unorm_to_float(g.albedo)))
            for path, plane in jobs:
                if plane is None or not write_exr(path, plane):
                    print(f"Failed to store image: {path}")
                    fine = False
                    break
        return fine


class IlluminationBufferIO:
    @staticmethod
    def import_illumination(illumination_format: str, num_frames: int, verbosity: int = 1) -> List[OfflineIllumination]:
        """:501-548"""
        out = []
        for f in range(num_frames):
            illu = OfflineIllumination()
            out.append(illu)
            img = read_exr(illumination_format % f)
            if img is None:
                print(f"Failed to load image: {illumination_format % f}")
                continue
            illu.noisy = np.ascontiguousarray(_rgba(img))
        return out

    @staticmethod
    def export_illumination(illumination_format: str, num_frames: int, illus: List[OfflineIllumination], verbosity: int = 1) -> bool:
        """:550-591"""
        fine = True
        for f in range(num_frames):
            if not write_exr(illumination_format % f, illus[f].noisy):
                print(f"Faled to store image: {illumination_format % f}")
                fine = False
        return fine
