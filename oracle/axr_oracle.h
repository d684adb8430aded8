/* TEST INFRASTRUCTURE — CPU restatement of AxiomR's tiled rasterisation path. See axr_oracle.c. */
#ifndef AXR_ORACLE_H
#define AXR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as `struct axr_ref_scene` in ref_harness.cpp so one ctypes struct serves both. */
typedef struct axo_scene {
	int width, height;
	int threads;              /* tile workers (results do not depend on it) */
	int shader_kind;          /* 0 FlatShader, 1 PhongShader, 2 PBRShader, 3 CutoutShader (ref_harness.cpp: discards) */
	float light_dir[3];
	float light_color[3];
	float specular_exponent;
	const uint8_t* tex[5];    /* diffuse, bump, metallic, roughness, ao; RGBA8 row 0 = top */
	int tex_w[5], tex_h[5];
	float view_proj[16];      /* column-major */
	float cam_pos[3];
	float model[16];
	int sampler;              /* 0 nearest (the reference), 1 bilinear (extension; no reference counterpart) */
} axo_scene;

int axo_render(const axo_scene* sc, const float* vertices, uint64_t n_verts, const uint32_t* indices,
               uint64_t n_faces, uint8_t* color_inout, float* depth_inout, double* seconds_out);
int axo_clip_triangle(const float* in3x18, float* out24x18);
int axo_triangle_setup(const float* in3x18, int w, int h, float* out13);
void axo_texture_sample(const uint8_t* rgba, int w, int h, int sampler, const float* uv, int n, float* out_rgba);
void axo_mat4_mul(const float a[16], const float b[16], float out[16]);
void axo_mat4_inverse(const float m[16], float out[16]);

#ifdef __cplusplus
}
#endif
#endif
