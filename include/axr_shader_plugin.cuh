// A user-supplied IShader at run time (reference include/IShader.hpp:30-46: any subclass can be handed to Pipeline::setShader).
// On the device a shader is a functor the tile kernels are instantiated with, so a new one has to be compiled — by its author, with
// nvcc, against these headers — into a small shared library that axr_load_shader_plugin() opens at run time:
//
//     #include "axr_shader_plugin.cuh"
//     struct MyShader {
//         static constexpr int NV = 5;                 // floats of the VertexOutput (interpolated as bar.x*v0 + bar.y*v1 + bar.z*v2)
//         static constexpr bool DISCARDS = false;      // true if fragment() may return true (the draw is then depth-peeled)
//         static constexpr bool HAS_FAST = false;      // no fused shade_fast() form
//         static constexpr unsigned TEXTURES = 1u;     // material slots the fragment stage dereferences: bit 0 diffuse, 1 bump, 2 metallic, 3 roughness, 4 ao
//         __device__ static void vertex(const axr::Uniforms& u, axr::v3 pos, axr::v3 n, axr::v3 t, axr::v3 b, float uvx, float uvy, float* o);
//         template <int SMP> __device__ static bool fragment(const axr::Uniforms& u, const axr::Material& m, const float* var, axr::v4& color);
//     };
//     AXR_SHADER_PLUGIN(MyShader)
//
//     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared -cudart static \
//          -I include -I axiomr_b200/csrc -o my_shader.so my_shader.cu                 (tools/build_shader_plugin.py does exactly this)
//
// Uniforms::light_dir / light_color carry the axr_shader_params given to axr_set_shader, Uniforms::user the floats given to
// axr_set_shader_user. The library checks that plugin and library were built from the same kernel headers (layout hash).
#pragma once
#include "axr_kernels.cuh"

namespace axr {
template <typename Shader>
inline int plugin_launch(const void* mv, const void* u, const void* fp, const void* in, int device, int sampler, int which, unsigned gx, unsigned gy, void* stream) {
	if (cudaSetDevice(device) != cudaSuccess) return (int)cudaGetLastError();
	const MeshView& m = *static_cast<const MeshView*>(mv);
	const Uniforms& un = *static_cast<const Uniforms*>(u);
	const FrameParams& f = *static_cast<const FrameParams*>(fp);
	const TileIn& t = *static_cast<const TileIn*>(in);
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	if (which == 0) {
		bool filled = false;
		if constexpr (!Shader::DISCARDS) {  // axr_set_output_fill (multi-GPU composite slots): a second instantiation
			if (t.fill) {
				if (sampler) k_tile_shade<Shader, 1, false, true><<<dim3(gx, gy), TILE_THREADS, 0, s>>>(m, un, f, t);
				else k_tile_shade<Shader, 0, false, true><<<dim3(gx, gy), TILE_THREADS, 0, s>>>(m, un, f, t);
				filled = true;
			}
		}
		if (!filled) {
			if (sampler) k_tile_shade<Shader, 1, false, false><<<dim3(gx, gy), TILE_THREADS, 0, s>>>(m, un, f, t);
			else k_tile_shade<Shader, 0, false, false><<<dim3(gx, gy), TILE_THREADS, 0, s>>>(m, un, f, t);
		}
	} else {
		if (sampler) k_shade_clipped<Shader, 1><<<CLIP_SHADE_CTAS, CLIP_SHADE_THREADS, 0, s>>>(m, un, f, t);
		else k_shade_clipped<Shader, 0><<<CLIP_SHADE_CTAS, CLIP_SHADE_THREADS, 0, s>>>(m, un, f, t);
	}
	return (int)cudaGetLastError();
}
}  // namespace axr

#define AXR_SHADER_PLUGIN(SHADER)                                                                                                         \
	extern "C" unsigned long long axr_shader_plugin_layout(void) { return axr::plugin_layout_hash(); }                                      \
	extern "C" int axr_shader_plugin_discards(void) { return SHADER::DISCARDS ? 1 : 0; }                                                    \
	extern "C" unsigned axr_shader_plugin_textures(void) { return SHADER::TEXTURES; }                                                       \
	extern "C" int axr_shader_plugin_launch(const void* mv, const void* u, const void* fp, const void* in, int device, int sampler, int which, \
	                                        unsigned gx, unsigned gy, void* stream) {                                                      \
		return axr::plugin_launch<SHADER>(mv, u, fp, in, device, sampler, which, gx, gy, stream);                                           \
	}
