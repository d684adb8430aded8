"""axiomr_b200 — B200-native (sm_100a) implementation of AxiomR's tiled rasterisation path.

Layout: csrc/ (CUDA kernels + the C ABI of include/axr_b200.h -> libaxr_b200.so), api.py (ctypes binding and the
host-side mirror of the reference's Pipeline / IShader / Mesh / Framebuffer interface), host/ (C++ adapter with the
reference's class names), scenes.py (synthetic inputs), multi.py (screen-band / multi-view sharding over torch.distributed).
"""
from .api import (AxrError, Camera, Color, CutoutShader, Device, FlatShader, Framebuffer, IShader, Material, MaterialGroup, Mesh,  # noqa: F401
                  PBRShader, PhongShader, Pipeline, Texture, TiledPipeline, load_library, render_scene,
                  SAMPLER_BILINEAR, SAMPLER_NEAREST, SHADER_CUTOUT, SHADER_FLAT, SHADER_PBR, SHADER_PHONG)

__all__ = ["AxrError", "Camera", "Color", "CutoutShader", "Device", "FlatShader", "Framebuffer", "IShader", "Material", "MaterialGroup",
           "Mesh", "PBRShader", "PhongShader", "Pipeline", "Texture", "TiledPipeline", "load_library", "render_scene"]
