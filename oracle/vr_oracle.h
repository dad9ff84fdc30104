/*
 * vr_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement (multithreaded C++, fp32, -ffp-contract=off) of the reference's
 * VolumetricReSTIR shader logic.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (volumetricrestirrelease_b200/csrc) never includes, links or calls it.
 *
 * Parity pin status: the RNG core is pinned against the reference's vendored public-domain xoshiro128** / SplitMix64 C
 * (Source/Externals/xoshiro, compiled into oracle/_ref by oracle/Makefile) and against the KATs in SURVEY.md section 8c.
 * Everything else on this path has NO golden vector or test in the reference: PARITY UNPINNED
 * beyond the RNG; self-made pins (closed-form transmittance, unbiasedness vs the reference's own path tracer mode,
 * bit-packing round trips, HDDA visit dumps) live in tests/.
 *
 * Input structs are the public ones from include/vrestir.h (interface header, no product code).
 */
#ifndef VR_ORACLE_H_
#define VR_ORACLE_H_

#include "../include/vrestir.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vro_pass vro_pass;

typedef struct vro_counters {
    uint64_t density_taps;      /* density evaluations (trilinear = 1 tap of 8 voxels, point = 1 tap of 1 voxel) */
    uint64_t voxels_fetched;    /* individual voxel reads */
    uint64_t voxel_bytes;       /* voxel reads x stored bytes/voxel of the slot */
    uint64_t node_visits;       /* tree node + child-list fetches (36 B each) */
    uint64_t rng_draws;
    uint64_t marches;           /* VolumeTrackingGVDB invocations */
} vro_counters;

const char* vro_last_error(void);
int vro_create(const vrestir_params* params, vro_pass** out);
int vro_destroy(vro_pass* p);
int vro_set_threads(vro_pass* p, int threads);             /* 0 = hardware_concurrency */
int vro_get_threads(const vro_pass* p);
int vro_set_volume(vro_pass* p, const vrestir_grid_desc* grid);      /* caller keeps arrays alive */
int vro_advance_volume(vro_pass* p, const vrestir_grid_desc* grid);
int vro_set_camera(vro_pass* p, const vrestir_camera* cam);
/* importance: optional precomputed importance map chain (same layout as VRESTIR_BUF_ENV_IMPORTANCE); NULL = compute */
int vro_set_envmap(vro_pass* p, const vrestir_envmap_desc* env, const float* importance);
int vro_set_env_alias(vro_pass* p, const float* thresholds, const uint32_t* redirect, const float* pdf, int count);
int vro_set_analytic_lights(vro_pass* p, const vrestir_light* lights, int count);
/* alias table built by the caller (AliasTable.cpp layout): items = {threshold, indexA, indexB, pad} x count */
int vro_set_emissive_triangles(vro_pass* p, const vrestir_emissive_triangle* tris, int count, const uint32_t* alias_items,
                               const float* weights, float weight_sum, float emissiveIntensityMultiplier);
int vro_set_frame(vro_pass* p, int width, int height);
/* restrict work to a pixel rectangle (bounded CPU-baseline samples); buffers stay full-frame */
int vro_set_crop(vro_pass* p, int x0, int y0, int x1, int y1);
int vro_update(vro_pass* p, const char* key, double value);
int vro_set_params(vro_pass* p, const vrestir_params* params);
int vro_get_params(const vro_pass* p, vrestir_params* out);
int vro_set_frame_count(vro_pass* p, int frame_count, int temporal_sample_accumulated);
int vro_get_frame_count(const vro_pass* p, int* frame_count);
int vro_execute(vro_pass* p, float* out_color, float* out_mvec);
int vro_execute_stage(vro_pass* p, int stage, int arg, float* out_color, float* out_mvec);
int vro_buffer_bytes(const vro_pass* p, int buffer, size_t* bytes);
int vro_get_buffer(vro_pass* p, int buffer, void* dst, size_t bytes);
int vro_set_buffer(vro_pass* p, int buffer, const void* src, size_t bytes);
int vro_spatial_input_buffer(const vro_pass* p, int round, int* buffer);
int vro_get_counters(vro_pass* p, vro_counters* out, int reset);
int vro_get_stage_ms(vro_pass* p, vrestir_timings* out);

/* --- unit-level hooks for KATs / property tests --- */
void vro_rng_words(uint32_t px, uint32_t py, uint32_t sample_number, int n, uint32_t* out_words, float* out_floats);
uint32_t vro_morton(uint32_t x, uint32_t y);
/* transmittance of one world-space ray: method = VRESTIR_*_TRACKING; rng seeded by (seed_px, seed_py, seed_n) */
float vro_transmittance(vro_pass* p, const float origin[3], const float dir[3], float tmax, int method, int mip,
                        int linear, float tstep_scale, uint32_t seed_px, uint32_t seed_py, uint32_t seed_n);
float vro_density_world(vro_pass* p, const float pos[3], int mip);
/* HDDA visit dump: records up to max_cells (slot-level, ix,iy,iz, t_enter) for brick visits of one ray; returns count */
int vro_dump_brick_visits(vro_pass* p, const float origin[3], const float dir[3], int mip, int vertex_center,
                          int max_cells, int32_t* out_xyz, float* out_t);
/* bit-packing helpers (VR/ReSTIRHelper.slang:10-89) */
int32_t vro_encode_max_indirect_bounces(int32_t storage, int32_t bounce);
int32_t vro_decode_max_indirect_bounces(int32_t storage, int32_t max_bounces);
int32_t vro_encode_path_tag(int32_t storage, int32_t tag);
int32_t vro_decode_path_tag(int32_t storage);
void vro_encode_wi_dist(const float in4[4], float out3[3]);
void vro_decode_wi_dist(const float in3[3], float out4[4]);
/* env map: evaluate / sample (for light-sampler tests) */
void vro_env_eval(vro_pass* p, const float dir[3], float out_rgb[3]);
int vro_env_sample(vro_pass* p, float u0, float u1, float out_dir[3], float* out_pdf, float out_Le[3]);
/* spatial neighbour offsets of a round (R2 in double / Hammersley), VR/SpatialReuse.cs.slang:64-81,109 */
int vro_record_k1_generator(vro_pass* p, int on);
int vro_get_k1_generator(vro_pass* p, int px, int py, uint32_t out8[8]);
float vro_visibility_state(vro_pass* p, const float o[3], const float d[3], float tmax, int method, int mip, int linear, float tss, int samples,
                           uint32_t spx, uint32_t spy, uint32_t sn, uint32_t out_state[4]);
float vro_sample_supervoxel(vro_pass* p, const float o[3], const float d[3], int mip, uint32_t spx, uint32_t spy, uint32_t sn, uint32_t out_state[4]);
void vro_sample_distances(vro_pass* p, const float o[3], const float d[3], int mip, int linear, int n, uint32_t spx, uint32_t spy, uint32_t sn, float out12[12], uint32_t out_state[4]);
float vro_p_hat(vro_pass* p, int px, int py, float depth, float uvx, float uvy, int lightID);
float vro_phase_hg(float cos_theta, float g);
float vro_sample_phase(float g, const float wo[3], float u0, float u1, float out_wi[3]);
void vro_neighbor_offsets(vro_pass* p, int frame_count, int round, int32_t* out_xy);

#ifdef __cplusplus
}
#endif
#endif
