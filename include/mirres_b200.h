/*
 * mirres_b200.h -- C ABI of libmirres_b200.so, the sm_100a implementation of MIRReS's path tracer with
 * screen-space ReSTIR (the hot path nerf/renderer_restir.py drives in the reference).
 *
 * Conventions
 *   - every entry point is `extern "C" int fn(..., void *stream)`; `stream` is a cudaStream_t (NULL = legacy
 *     default stream).  Return value 0 = enqueued, < 0 = error (MIRRES_ERR_*, or -100 - cudaError for a
 *     launch failure).  Nothing allocates, synchronises, or touches the host after argument checks.
 *   - all pointers are DEVICE pointers into dense row-major tensors owned by the caller (in the product the
 *     caller is PyTorch: `tensor.data_ptr()`), with the shapes/dtypes of the reference tensors they replace;
 *     `[N,k]` means N rows of k fp32/int32 values, N = framedim_x * framedim_y, pixelIndex = y * framedim_x + x.
 *   - numerical contract: include/mirres_fpmath.h.  Integer outputs (BVH topology, hit flags, primitive ids,
 *     light texels, reservoir sample choices) are bit-exact against the oracle; radiance, reservoir weights
 *     and gradients are reproducible to the stated tolerance (float atomics in the env-gradient scatter).
 *   - each declaration cites the reference interface it replaces (file:line under the reference root).
 */
#ifndef MIRRES_B200_H
#define MIRRES_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIRRES_ABI_VERSION 1

#define MIRRES_ERR_NULL (-1)    /* a required pointer is NULL */
#define MIRRES_ERR_SHAPE (-2)   /* a size argument is out of range */
#define MIRRES_ERR_ALIGN (-3)   /* a pointer that must be 16-byte (BVH records: 32-byte, scratch: 256-byte) aligned is not */
#define MIRRES_ERR_SCRATCH (-4) /* scratch buffer smaller than mirres_bvh_scratch_bytes() */
#define MIRRES_ERR_ALIAS (-5)   /* input and output buffers that must differ are the same */

int mirres_abi_version(void);

/* Launch-shape tuning (SURVEY.md 8b allows a tuning context).  The values belong to the CALLING HOST THREAD (thread-local
 * storage): they shape the persistent queue tracers launched by subsequent calls of that thread, so two threads that drive
 * different streams never see each other's settings; value <= 0 restores the default.  Results do not depend on them.
 * The library keeps no other mutable state: device properties (SM count, occupancy) are cached per device, and no
 * environment variable is read.
 *   ANY_BLOCKS / CLOSEST_BLOCKS / MIXED_BLOCKS   blocks per SM of the boolean-ray / closest-hit / combined tracer grids
 *                                                (defaults 3 / 2 / 4: small grids leave room for kernels of other streams)
 *   CLOSEST_SPLIT                                1 (default): idle lanes walk deferred subtrees of closest-hit rays and
 *                                                the result is rebuilt exactly from their hit logs; 2: one lane per ray */
#define MIRRES_TUNE_ANY_BLOCKS 0
#define MIRRES_TUNE_CLOSEST_BLOCKS 1
#define MIRRES_TUNE_MIXED_BLOCKS 2
#define MIRRES_TUNE_CLOSEST_SPLIT 3
#define MIRRES_TUNE_ANY_TOP 4 /* 1: boolean-ray walkers read the first five wide levels of the tree from shared memory */
#define MIRRES_TUNE_COUNT_ 8
int mirres_set_tuning(int key, int value);
int mirres_get_tuning(int key); /* current value of `key` for the calling thread (0 = default), or MIRRES_ERR_SHAPE */

/* ------------------------------------------------------------------------------------------------------------
 * LBVH construction.  Replaces restirbvhWorker.update_bvh, nerf/renderer_restir.py:25-89, and the Slang
 * kernels it launches (nerf/bvhworkers/{get_elements,lbvh_morton_codes,lbvh_single_radixsort,lbvh_hierarchy,
 * lbvh_bounding_boxes}.slang).  vert [V,3] f32, tri [F,3] i32.  Outputs in the reference layout:
 *   info [2F-1,3] i32 (left, right, primitive; leaf <=> left == right == 0), aabb [2F-1,6] f32 (min xyz, max xyz),
 *   internal nodes [0,F-2] (root 0), leaves [F-1,2F-2] in sorted-Morton order  (renderer_restir.py:61-64).
 * packed_nodes / packed_tris (optional, both or neither): traversal records consumed by every ray-casting
 * entry point below; sizes from mirres_bvh_packed_{node,tri}_bytes, 32-byte aligned (256-bit loads).  Opaque to the
 * caller.  (packed_nodes = a table of the first five wide levels, 43 KB, followed by one 128-byte record per internal
 * node: the boxes of its up to four grandchildren in the reference's visit order; packed_tris = one 64-byte record
 * per leaf in sorted-Morton order: v0 | primitive, e1, e2.)
 * The build is nine stream-ordered launches and never synchronises.  Morton codes are 30-bit (lbvh_morton_codes.slang:
 * 24-44): the stable sort runs three 10-bit passes.
 * sorted_codes (optional) [F,2] i32: (Morton code, element index) after the stable sort (renderer_restir.py:48-57).
 * scratch: mirres_bvh_scratch_bytes(F) bytes, 256-byte aligned; contents need not be preserved or cleared between calls
 * (the build's first launch initialises what it needs).  The sort passes wait on counters inside it: scratch must not
 * be written by anything else while a build is in flight.
 */
size_t mirres_bvh_scratch_bytes(int F);
size_t mirres_bvh_packed_node_bytes(int F);
size_t mirres_bvh_packed_tri_bytes(int F);
int mirres_bvh_build(const float *vert, int V, const int *tri, int F, int *info, float *aabb, void *packed_nodes,
                     void *packed_tris, int *sorted_codes, void *scratch, size_t scratch_bytes, void *stream);

/* Granular stages with the argument meaning of the individual reference kernels (used by the slangpy-protocol
 * shim so that the reference's own update_bvh can run unchanged):
 *   generateElements  nerf/bvhworkers/get_elements.slang:3-39          (renderer_restir.py:32-33)
 *   morton_codes      nerf/bvhworkers/lbvh_morton_codes.slang:46-79    (renderer_restir.py:44-51)
 *   radix_sort        nerf/bvhworkers/lbvh_single_radixsort.slang:28-138 (renderer_restir.py:55-57)
 *   hierarchy + get_bvh_height + get_bbox* + set_root
 *                     nerf/bvhworkers/lbvh_hierarchy.slang:111-245, lbvh_bounding_boxes.slang:151-390
 *                                                                       (renderer_restir.py:61-87)            */
int mirres_bvh_elements(const float *vert, const int *tri, int F, int *ele_primitiveIdx, float *ele_aabb, void *stream);
int mirres_bvh_morton(const float *ele_aabb, int F, float min_x, float min_y, float min_z, float max_x, float max_y,
                      float max_z, int *morton_codes_ele, void *stream);
int mirres_bvh_sort(int *pairs, int F, void *scratch, size_t scratch_bytes, void *stream);
int mirres_bvh_hierarchy_refit(const int *sorted_pairs, const float *ele_aabb, int F, int *info, float *aabb,
                               void *scratch, size_t scratch_bytes, void *stream);
/* traversal records from reference-layout tensors built elsewhere */
int mirres_bvh_pack(const int *info, const float *aabb, const float *vert, const int *tri, int F, void *packed_nodes,
                    void *packed_tris, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Standalone rays.  Semantics of bvh_hit / bvh_hit_with_normal, nerf/ScreenSpaceReSTIR/utils/helperDi.slang:197-274,
 * 313-395, with t_min = 0, t_max = 1e7 (the only values any call site uses).  org/dir [n,3]; dir is normalised
 * inside.  hit [n] i32; t [n], pos [n,3], normal [n,3], prim [n] i32 and visits [n,2] u32 (node records fetched x2,
 * triangles tested) are optional.  `prim` and `visits` are extensions the reference does not output.
 */
int mirres_trace_closest(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                         int *hit, float *t, float *pos, float *normal, int *prim, unsigned int *visits, void *stream);
int mirres_trace_any(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                     int *hit, unsigned int *visits, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Environment light.  env_tex [H*W,3] is the vertically flipped, flattened lgt.base (renderer_restir.py:305-311).
 *   mirres_env_build_distribution  replaces make_sampleable(), nerf/ScreenSpaceReSTIR/GenerateLightTiles.py:4-29
 *       (make_sampleable.slang:34-86 + torch sum/cumsum): pdf_ [H*W], cdf_ [H*(W+1)], mpdf_ [H], mcdf_ [H+1];
 *       row_scratch [H].  Scans are sequential fp32 prefix sums (defined order).
 *   mirres_env_weights / mirres_env_distribution2d  the two Slang kernels alone (GenerateLightTiles.py:7-8,20-21).
 *   mirres_neighbor_offsets  createNeighborOffsetTexture, make_sampleable.slang:186-205 (renderer_restir.py:217-219);
 *       out [2*sample_count] raw int8-range values (the host divides by 127).
 *   mirres_light_tiles  process_GenerateLightTiles, GenerateLightTiles.slang:16-62 (GenerateLightTiles.py:42-50):
 *       light_data [T,3] (valid, oct.u, oct.v), light_uv [T,2] i32, light_pdf [T] (the reference names it light_inv_pdf).
 *       light_cache (optional, [T,8] f32, 16-byte aligned): world direction and emitted radiance of every slot, i.e.
 *       what get_light_info (lightDi.slang:291-298) returns for it; mirres_initial_resampling reads it instead of
 *       re-deriving both for each of the 32 candidates of every pixel (identical values, computed once per slot).
 *       frame_offset (optional, device): added to frame_index at run time (see "frame offset" below).
 */
int mirres_env_build_distribution(const float *env_tex, int W, int H, float *pdf_, float *cdf_, float *mpdf_,
                                  float *mcdf_, float *row_scratch, void *stream);
int mirres_env_weights(const float *env_tex, int W, int H, float *weight, void *stream);
int mirres_env_distribution2d(int W, int H, float *pdf_, float *cdf_, void *stream);
int mirres_neighbor_offsets(int sample_count, float *out, void *stream);
int mirres_light_tiles(const float *env_tex, int W, int H, const float *pdf_, const float *cdf_, const float *mpdf_,
                       const float *mcdf_, unsigned int frame_index, int tile_count, int tile_size, float *light_data,
                       int *light_uv, float *light_pdf, float *light_cache, const unsigned int *frame_offset, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Wavefront workspace.  Every ray-casting entry point below runs as  gen (one thread per foreground pixel) ->
 * ray queue -> queue tracer -> resolve  and needs a caller-allocated workspace of mirres_workspace_bytes(N) bytes
 * (256-byte aligned), N = framedim_x * framedim_y.  mirres_workspace_prepare builds the ordered list of foreground
 * pixels (occ >= 0.1, the test every reference kernel starts with, e.g. InitialResampling.slang:166) and must be
 * called whenever the PRIMARY occupancy map changes (once per frame); the bounce kernels, whose `occ` argument is the
 * occupancy of the previous path vertex, reuse the same list (a path vertex only exists where the primary hit does).
 * Frame offset: the 32-bit word at byte offset MIRRES_WORKSPACE_FRAME_OFFSET_BYTES of a workspace is added to the
 * frame_index argument of every entry point that takes that workspace.  The library never writes it; the caller
 * zero-fills the workspace once and may update the word between launches (e.g. before replaying a CUDA graph whose
 * frame indices are baked in).  Zero reproduces the reference's frame-index schedule exactly.
 * Row offset: the word at MIRRES_WORKSPACE_ROW_OFFSET_BYTES is added to a pixel's ROW when its random stream is seeded
 * (Seed_Generator(pixel, frame), nerf/ScreenSpaceReSTIR/utils/random.slang:2-39) -- and to nothing else.  The [N, k]
 * maps are row-major, so a band of rows is a contiguous slice of every tensor: a rank that renders rows [y0, y1) of a
 * larger frame passes the slices with framedim_y = y1 - y0 and sets this word to y0, and every pixel
 * draws the random numbers of its position in the full frame.  Same ownership rules as the frame offset; zero = plain.
 */
#define MIRRES_WORKSPACE_FRAME_OFFSET_BYTES 32
#define MIRRES_WORKSPACE_ROW_OFFSET_BYTES 36
/* Band of the spatial pass: two words [lo, hi) at MIRRES_WORKSPACE_BAND_BYTES.  With hi > lo, mirres_spatial_resampling
 * resamples only the listed pixels of rows [lo, hi) of the frame it is handed; the other listed pixels publish their
 * reservoir sample for those rows to reuse and keep a zero reservoir in the output (row-band rendering: the pass reads
 * neighbours up to 30 rows beyond the band).  hi <= lo (the zero-filled default): every listed pixel. */
#define MIRRES_WORKSPACE_BAND_BYTES 40
/* Error word: set to 1 by a ray-casting entry point whose traversal had to DROP a stack entry because a ray's stack was
 * full (64 entries, the reference's unchecked depth, helperDi.slang:136): results of that launch may miss a subtree.
 * Entry points cannot return it (they never synchronise); the caller reads the word when it synchronises anyway and
 * clears it.  The library never clears it. */
#define MIRRES_WORKSPACE_ERROR_BYTES 48
size_t mirres_workspace_bytes(int n_pixels);
int mirres_workspace_prepare(const float *occ, int n_pixels, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * ReSTIR passes.  Reservoir = (light_data [N,3], light_pdf [N], M [N] i32, weight [N]), res.slang:5-11.
 * G-buffer: occ [N], normal_depth [N,4] (16-byte aligned), brdf_map [N,3] (lum kd, metallic, alpha), ray_dir [N,3],
 * pos_map [N,3]  (renderer_restir.py:279-287).
 *   mirres_initial_resampling   process_InitialResampling_, InitialResampling.slang:151-295 (renderer_restir.py:96-114)
 *   mirres_temporal_resampling  process_TemporalResampling, TemporalResampling.slang:23-135 (Resampling.py:28-44);
 *                               motion [N,2] may be NULL (= zeros, as renderer_restir.py:487 passes)
 *   mirres_spatial_resampling   process_SpatialResampling_, SpatialResampling.slang:178-322 (renderer_restir.py:116-131);
 *                               reads prev_*, writes res_* (must differ); offset_count must be a power of two
 *   mirres_final_visibility     process_EvaluateFinalSamples_get_vis, EvaluateFinalSamples.slang:84-124 (renderer_restir.py:133-146)
 *   mirres_eval_final_fwd/bwd   process_EvaluateFinalSamples_di_ and its Slang-autodiff `.bwd`,
 *                               EvaluateFinalSamples.slang:129-188 (Resampling.py:94-143); bwd ACCUMULATES into grad_env [He*We,3]
 */
int mirres_initial_resampling(const void *packed_nodes, const void *packed_tris, const float *pos_map, float *res_ld,
                              float *res_pdf, int *res_M, float *res_w, const float *env_tex, int env_w, int env_h,
                              int fx, int fy, unsigned int frame_index, const float *occ, const float *normal_depth,
                              const float *brdf_map, const float *ray_dir, const float *pdf_, const float *mpdf_,
                              const float *light_data, const float *light_pdf, const float *light_cache,
                              int tile_count, int tile_size, int screen_tile, int n_light, int n_brdf, void *workspace,
                              size_t workspace_bytes, void *stream);
int mirres_temporal_resampling(float *res_ld, float *res_pdf, int *res_M, float *res_w, const float *prev_ld,
                               const float *prev_pdf, const int *prev_M, const float *prev_w, const float *env_tex,
                               int env_w, int env_h, int fx, int fy, unsigned int frame_index, const float *occ,
                               const float *normal_depth, const float *brdf_map, const float *ray_dir,
                               const float *prev_occ, const float *prev_normal_depth, const float *prev_brdf_map,
                               const float *prev_ray_dir, const float *motion, int max_history, void *workspace, size_t workspace_bytes, void *stream);
int mirres_spatial_resampling(const void *packed_nodes, const void *packed_tris, const float *pos_map, float *res_ld,
                              float *res_pdf, int *res_M, float *res_w, const float *prev_ld, const float *prev_pdf,
                              const int *prev_M, const float *prev_w, const float *neighbor_offsets, const float *env_tex,
                              int env_w, int env_h, int fx, int fy, unsigned int frame_index, const float *occ,
                              const float *normal_depth, const float *brdf_map, const float *ray_dir, int offset_count,
                              int neighbor_count, float gather_radius, void *workspace, size_t workspace_bytes, void *stream);
int mirres_final_visibility(const void *packed_nodes, const void *packed_tris, const float *res_ld, int fx, int fy,
                            const float *pos_map, float *vis_map, void *workspace, size_t workspace_bytes, void *stream);
/* Visibility tags (optional; off until set, per host thread like mirres_set_tuning).  One byte per pixel beside a
 * reservoir set: 1 = the stored sample has already been found UNOCCLUDED from this pixel by the very ray
 * mirres_final_visibility would cast for it (origin pos_map + 0.01 L, direction L, same tree); 0 = not known.  Every
 * sample a pixel ends up with has a provenance that says so: mirres_initial_resampling keeps a candidate only if that ray
 * missed (InitialResampling.slang:255-270); mirres_spatial_resampling selects a neighbour's sample only with a positive
 * weight, i.e. after slot 2k -- the same ray -- missed (SpatialResampling.slang:262-266); a pixel's own sample carried
 * through a pass keeps its tag; a history sample keeps it when mirres_temporal_resampling takes it from the same pixel
 * of a previous reservoir that belongs to the SAME pos_map and tree (the spp loop of one frame,
 * nerf/renderer_restir.py:314-459), and loses it otherwise.  mirres_final_visibility then casts rays only for samples
 * tagged 0; vis_map is bit-identical either way (the skipped rays are repetitions of rays that missed).
 *   res_tag   [N] bytes: the tag of the reservoir set the next pass WRITES (initial: out; temporal: in / out; spatial: out)
 *             or, for mirres_final_visibility, of the set it reads
 *   prev_tag  [N] bytes or NULL: the tag of the prev_* set of the temporal / spatial pass (NULL = nothing known)
 * The caller keeps the tags beside its reservoir sets and passes (NULL, NULL) when it has none: tags are never inferred. */
int mirres_set_visibility_tags(unsigned char *res_tag, const unsigned char *prev_tag);
int mirres_eval_final_fwd(const float *res_ld, const float *res_pdf, const int *res_M, const float *res_w,
                          const float *env_tex, int env_w, int env_h, int fx, int fy, float *fs_dir, float *fs_dist,
                          float *fs_Li, const float *vis_map, void *stream);
int mirres_eval_final_bwd(const float *res_ld, const float *res_pdf, const int *res_M, const float *res_w, int env_w,
                          int env_h, int fx, int fy, const float *vis_map, const float *grad_Li, float *grad_env,
                          void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Shading and the multi-bounce integrator.
 *   mirres_final_shading_fwd/bwd  process_FinalShading and its `.bwd`, FinalShading.slang:14-109 (Resampling.py:145-214);
 *        rough_metal [N,2] = (linear roughness, metallic).  bwd OVERWRITES grad_normal [N,3], grad_diffuse [N,3],
 *        grad_rough_metal [N,2], grad_Li [N,3].
 *   mirres_bounce_first   process_new_dir_for_pt, FinalShading.slang:113-265 (Resampling.py:216-232)
 *   mirres_bounce_shade   process_path_tracing_divided_no_grad, FinalShading.slang:641-1009 (Resampling.py:254-272)
 *        prd [N,5] (throughput rgb, specular flag, stop flag) is read and written; new_* must not alias the inputs.
 *        max_bounce replaces the compile-time MAX_Bounce = 2 (FinalShading.slang:7).
 */
int mirres_final_shading_fwd(const float *fs_dir, const float *fs_dist, const float *fs_Li, const float *env_tex,
                             int env_w, int env_h, int fx, int fy, const float *occ, const float *normal,
                             const float *ray_dir, const float *diffuse_map, const float *rough_metal, float *color,
                             float *diff_light, float *spec_light, void *stream);
int mirres_final_shading_bwd(const float *fs_dir, const float *fs_dist, const float *fs_Li, int fx, int fy,
                             const float *occ, const float *normal, const float *ray_dir, const float *diffuse_map,
                             const float *rough_metal, const float *grad_color, const float *grad_diff_light,
                             const float *grad_spec_light, float *grad_normal, float *grad_diffuse,
                             float *grad_rough_metal, float *grad_Li, void *stream);
int mirres_bounce_first(const void *packed_nodes, const void *packed_tris, unsigned int frame_index,
                        unsigned int bounce_count, int max_bounce, int fx, int fy, const float *occ, const float *pos_map,
                        const float *normal, const float *ray_dir, float *prd, const float *diffuse_map,
                        const float *rough_metal, float *new_pos, float *new_ray_d, float *new_occ, float *new_normal,
                        void *workspace, size_t workspace_bytes, void *stream);
int mirres_bounce_shade(const void *packed_nodes, const void *packed_tris, unsigned int frame_index,
                        unsigned int bounce_count, int max_bounce, int fx, int fy, const float *env_tex, int env_w,
                        int env_h, const float *pdf_, const float *cdf_, const float *mpdf_, const float *mcdf_,
                        const float *occ, const float *pos_map, const float *normal, const float *ray_dir, float *prd,
                        const float *diffuse_map, const float *rough_metal, float *color, float *diff_color,
                        float *spec_color, float *new_pos, float *new_ray_d, float *new_occ, float *new_normal,
                        void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Edge-avoiding a-trous denoiser and the normal-variation AO proxy.
 *   mirres_eaw_fwd   process_EAWDenoise / process_EAWDenoise_no_di, nerf/ScreenSpaceReSTIR/EAWDenoise.slang:50-302
 *                    (nerf/ScreenSpaceReSTIR/Denoising.py:10-61); PHI = (c_phi, n_phi, p_phi); step_width is truncated
 *                    to int as Denoising.py:18 does; out_color must not alias color.
 *   mirres_eaw_bwd   the Slang-autodiff `.bwd` of process_EAWDenoise (Denoising.py:30-48), as a deterministic
 *                    gather; OVERWRITES grad_color / grad_normal / grad_pos [N,3]; cum_w_scratch [N].
 *   mirres_normal_ao process_normal_ao, EAWDenoise.slang:591-647 (nerf/renderer.py:1153-1158); out_ao [N,3].
 */
int mirres_eaw_fwd(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                   const float *color, const float *normal, const float *pos, float *out_color, void *stream);
int mirres_eaw_bwd(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                   const float *color, const float *normal, const float *pos, const float *out_color,
                   const float *grad_out, float *grad_color, float *grad_normal, float *grad_pos, float *cum_w_scratch,
                   void *stream);
int mirres_normal_ao(int fx, int fy, const float *occ, const float *normal, float *out_ao, void *stream);
/* Cross-bilateral denoiser of --use_bi_de: bilateral_denoiser_fwd / _bwd_kernel,
 * nerf/renderutils/c_src/denoising.cu:14-130, reached from nerf/renderer_restir.py:529-541 through
 * nerf/renderutils/ops.py:173-212.  Frame fx x fy, sigma = max(2 factor, 1e-4), filter radius 2 ceil(2.5 sigma) + 1.
 * col [N,3], nrm [N,3] (normalised by the caller, ops.py:197), zdz [N,2] = (depth, depth gradient).
 *   fwd: out [N,4] = (sum_t w col_t, max(sum_t w, 1e-4)),  w = exp(-d^2 / 2 sigma^2) * clamp(n_t . n_c, 1e-4, 1)^128 *
 *        exp(-|z_t - z_c| / max(dz_c d, 1e-4))
 *   bwd: col_grad [N,3] = transposed gather of out_grad [N,4] (first three channels), depth term with the tap's dz. */
int mirres_bilateral_fwd(int fx, int fy, float sigma, const float *col, const float *nrm, const float *zdz, float *out,
                         void *stream);
int mirres_bilateral_bwd(int fx, int fy, float sigma, const float *nrm, const float *zdz, const float *out_grad,
                         float *col_grad, void *stream);
/* Batched a-trous level: n_images (<= 8) colour images filtered over the same occ / normal / pos in one pass -- the five
 * images run_restir_di_with_pt denoises per level (nerf/renderer_restir.py:517-541).  colors / out_colors / ... are HOST
 * arrays of n_images device pointers (read at launch).  Per-image results are bit-identical to mirres_eaw_fwd / _bwd;
 * the guide-buffer loads and the normal / position edge weights are shared.  cum_w (optional in forward, [N] per image)
 * receives the normalisation of every footprint, which mirres_eaw_bwd_multi consumes instead of recomputing it.
 * Backward OVERWRITES grad_colors[m] ([N,3] per image) and grad_normal_sum / grad_pos_sum ([N,3] each, optional): the
 * normal / position gradients summed over the images, which is what autograd forms from them.  The backward carries
 * the gradient tolerance (1e-3), not bit-exactness: it multiplies by reciprocals and uses the hardware exp. */
int mirres_eaw_fwd_multi(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                         const float *normal, const float *pos, int n_images, const float *const *colors,
                         float *const *out_colors, float *const *cum_w, void *stream);
int mirres_eaw_bwd_multi(float c_phi, float n_phi, float p_phi, int fx, int fy, float step_width, const float *occ,
                         const float *normal, const float *pos, int n_images, const float *const *colors,
                         const float *const *out_colors, const float *const *cum_w, const float *const *grad_outs,
                         float *const *grad_colors, float *grad_normal_sum, float *grad_pos_sum, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * G-buffer producer and gradient scatter.
 *   mirres_gbuffer_primary   stands in for the nvdiffrast rasterise + interpolate stage that feeds
 *        run_restir_di_with_pt (nerf/renderer.py:979-1030, 1092-1096): one closest-hit ray per pixel with the
 *        semantics of bvh_hit_with_normal (helperDi.slang:313-395).  org/dir [n,3]; outputs occ [n] (1 hit / 0),
 *        pos [n,3], normal [n,3] (face normal flipped towards the ray, or -- when vnormal [V,3] and tri [F,3] are
 *        given -- the barycentric interpolation of the vertex normals, the role of dr.interpolate over
 *        auto_normals, nerf/meshutils.py:14-39), depth [n] = |pos - org|, prim [n] i32 (-1 miss), bary [n,2] = (u, v)
 *        weights of the triangle's 2nd / 3rd vertex.  prim and bary are optional; so is geom_normal [n,3], which keeps
 *        the face normal when `normal` receives the interpolated one (prepare_shading_normal wants both).  workspace (optional, sized by
 *        mirres_workspace_bytes(n)): the rays go through the persistent queue tracer instead of one thread per ray.
 *   mirres_interpolate_bwd   reverse of that interpolation, i.e. the scatter nvdiffrast / the texture backward do
 *        for the per-pixel gradients the path emits (Resampling.py:193-214):
 *            out[tri[prim[i]][k], c] += w_k(i) * grad[i, c],   w = (1-u-v, u, v)   (bary NULL: 1/3 each)
 *        grad [n,C], C <= 8; out [V,C] is ACCUMULATED into.  Lanes of a warp that see the same triangle are summed
 *        with shuffles first (__match_any_sync on prim), one lane issues the atomics.
 */
int mirres_gbuffer_primary(const void *packed_nodes, const void *packed_tris, const float *org, const float *dir, int n,
                           const float *vnormal, const int *tri, float *occ, float *pos, float *normal, float *depth,
                           int *prim, float *bary, float *geom_normal, void *workspace, size_t workspace_bytes,
                           void *stream);
/* Derived maps of run_restir_di_with_pt in one launch (nerf/renderer_restir.py:279-287 and :484-486; ~25 elementwise torch
 * launches in the reference, same operations in the same order): occ [n] in place (occ <= 0.5 -> 0); normal_depth [n,4]
 * = (normal, depth), 16-byte aligned; brdf_map [n,3] = (0.2126 r + 0.7152 g + 0.0722 b of kd, the same weights summed
 * over metallic, clamp(roughness, 0.01, 1)^2); ray_dir_normalized [n,3] = d / max(|d|, 1e-6). */
int mirres_prepare_maps(int n, float *occ, const float *normal, const float *depth, const float *diffuse_map,
                        const float *rough_metal, const float *ray_dir, float *normal_depth, float *brdf_map,
                        float *ray_dir_normalized, void *stream);
int mirres_interpolate_bwd(const float *grad, int n, int C, const int *prim, const float *bary, const int *tri, int F,
                           float *out, void *stream);
/* Vertex normals of the (offset) mesh: auto_normals (meshutils.py:14-39), which nerf/renderer.py:979-1030 evaluates on
 * the optimised vertices before interpolating them into the G-buffer normal.  Area-weighted face normals summed per
 * vertex (float atomics, like the reference's scatter_add_), normalised; a vertex whose sum has squared length <= 1e-20
 * (unreferenced / degenerate) gets (0, 0, 1).  vsum [V,3] (the un-normalised sums) is written by the forward and read
 * by the backward; the forward zero-fills it.  Triangles with an index outside [0, V) are ignored.
 * Backward: grad_vert [V,3] += d vnrm / d vert (ACCUMULATED into, like mirres_interpolate_bwd) -- with
 * mirres_interpolate_bwd and mirres_shading_normal_bwd this carries the path's grad_normal back to the mesh vertices. */
int mirres_vertex_normals_fwd(const float *vert, int V, const int *tri, int F, float *vsum, float *vnrm, void *stream);
int mirres_vertex_normals_bwd(const float *vert, int V, const int *tri, int F, const float *vsum, const float *grad_vnrm,
                              float *grad_vert, void *stream);
/* Shading-normal set-up of the G-buffer stage: prepare_shading_normal (nerf/renderutils/ops.py:129-163) =
 * PrepareShadingNormalFwdKernel / BwdKernel (nerf/renderutils/c_src/normal.cu:95-178), called at nerf/renderer.py:1013
 * on the rasterised maps before run_restir_di_with_pt.  Tangent-frame perturbation, two-sided flip, bending of
 * back-facing normals towards the eye (threshold 0.1).  Six inputs of 3 floats per pixel, each with a row stride in
 * floats: 3 = dense [n,3], 0 = one row broadcast (the reference passes view_pos and perturbed_nrm as [1,1,1,3]), any
 * other non-negative stride for a view into a wider tensor.  out [n,3].  The backward writes (not accumulates) full
 * [n,3] gradients for the inputs whose pointer is non-NULL -- like the reference it leaves the reduction over a
 * broadcast input to the caller. */
int mirres_shading_normal_fwd(int n, const float *pos, int pos_rs, const float *view_pos, int view_rs,
                              const float *perturbed_nrm, int perturbed_rs, const float *smooth_nrm, int smooth_nrm_rs,
                              const float *smooth_tng, int smooth_tng_rs, const float *geom_nrm, int geom_rs, int two_sided,
                              int opengl, float *out, void *stream);
int mirres_shading_normal_bwd(int n, const float *pos, int pos_rs, const float *view_pos, int view_rs,
                              const float *perturbed_nrm, int perturbed_rs, const float *smooth_nrm, int smooth_nrm_rs,
                              const float *smooth_tng, int smooth_tng_rs, const float *geom_nrm, int geom_rs, int two_sided,
                              int opengl, const float *grad_out, float *grad_pos, float *grad_view_pos,
                              float *grad_perturbed_nrm, float *grad_smooth_nrm, float *grad_smooth_tng, float *grad_geom_nrm,
                              void *stream);


/* ------------------------------------------------------------------------------------------------------------
 * Host-side chains of the spp loop as single launches.  Same operations in the same order as the
 * torch expressions they replace, so results are bit-identical.
 *   mirres_material_procedural   the synthetic stand-in of the `mlp_mat` query between bounces
 *        (nerf/renderer_restir.py:398-408, 428-438): kd_c = 0.1 + 0.8 tri(1.7 pos_c), roughness = 0.08 + 0.92
 *        tri(0.8 pos_x + 0.5 pos_y), tri(x) = |2 frac(x) - 1|, metallic constant.  mode 0: kd / rough_metal = value * occ
 *        (occ NULL: value) -- the G-buffer materials; mode 1: the torch.where merge -- pixels with occ >= 0.5 take the new
 *        value (kd times scale_xyz when given), the others keep theirs; with scale_xyz (HOST pointer to 3 floats, read at
 *        launch) kd is clamped to [0,1] afterwards (:405-408).
 *   mirres_sum_images            dst[i] = ((accumulate ? dst[i] : 0) + src_0[i] + src_1[i] + ...) [/ divisor if != 0]:
 *        the running sums of the per-iteration outputs and `total / mFrameIndex` (:443-459, :505-515).  src: HOST
 *        array of n_src <= 32 device pointers.  The division is evaluated the way torch evaluates tensor / python
 *        scalar on a CUDA tensor: multiplication by the fp32 reciprocal of the divisor.
 *   mirres_composite_fwd / _bwd  final_color = nan_to_num(where(occ <= 0.1, 1, kd (1 - metallic) dd + ds + di))
 *        (:543-549) and its reverse mode w.r.t. kd, (roughness, metallic), dd, ds (what torch autograd derives).
 *   mirres_final_shading_bwd_multi   mirres_final_shading_bwd for the n_passes <= 16 shading passes of an spp loop in one
 *        launch: the passes share surface inputs and upstream gradients (grad_color may be NULL = zeros; divided by
 *        grad_divisor first when it is != 0, the backward of `total / mFrameIndex`) and differ in their final samples;
 *        fs_dir / fs_dist / fs_Li / grad_Li are HOST arrays of n_passes device pointers.
 *        grad_normal / grad_diffuse / grad_rough_metal receive the sum over passes (last pass first, the autograd
 *        engine's order; accumulate != 0 continues an earlier call).  grad_Li[k] is written per pass, or -- with
 *        sum_grad_Li -- grad_Li[0] receives the sum (for passes that evaluated one shared reservoir buffer).  When in
 *        addition every pass names the SAME fs_dir / fs_dist buffers, the passes are evaluated once on the summed fs_Li
 *        (all gradients but grad_Li are linear in Li); the result differs from per-pass evaluation by rounding only.
 */
int mirres_material_procedural(int n, const float *pos, const float *occ, int mode, float metallic, const float *scale_xyz,
                               float *kd, float *rough_metal, void *stream);
int mirres_sum_images(int n_floats, int n_src, const float *const *src, float divisor, int accumulate, float *dst, void *stream);
int mirres_composite_fwd(int n, const float *occ, const float *diffuse_map, const float *rough_metal, const float *denoised_diffuse,
                         const float *denoised_spec, const float *denoised_indirect, float *final_color, void *stream);
int mirres_composite_bwd(int n, const float *occ, const float *diffuse_map, const float *rough_metal, const float *denoised_diffuse,
                         const float *denoised_spec, const float *denoised_indirect, const float *grad_final_color,
                         float *grad_diffuse_map, float *grad_rough_metal, float *grad_denoised_diffuse,
                         float *grad_denoised_spec, void *stream);
int mirres_final_shading_bwd_multi(int n_passes, const float *const *fs_dir, const float *const *fs_dist,
                                   const float *const *fs_Li, int fx, int fy, const float *occ, const float *normal,
                                   const float *ray_dir, const float *diffuse_map, const float *rough_metal,
                                   const float *grad_color, const float *grad_diff_light, const float *grad_spec_light,
                                   float grad_divisor, int accumulate, float *grad_normal, float *grad_diffuse, float *grad_rough_metal,
                                   int sum_grad_Li, float *const *grad_Li, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MIRRES_B200_H */
