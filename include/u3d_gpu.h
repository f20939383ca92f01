/* u3d_gpu.h -- C-ABI of the B200-native visibility path (libu3dgpu.so).
 *
 * The reference has no FFI layer for this path: it is a C++ class API.  Each entry point below therefore names the
 * reference member function it replaces ("OB" = Source/Urho3D/Graphics/OcclusionBuffer, "Octree" = Source/Urho3D/Graphics/Octree).
 * The C++ host module urho3d_b200/GPUCompute/ wraps these in classes with the reference's method names; INTEGRATION.md shows the
 * edits that re-point View.cpp / Renderer.cpp at them.
 *
 * Conventions
 *   - plain pointers and sizes only; host pointers unless a parameter says "device".
 *   - functions that mirror a reference `bool` return 1 (true) / 0 (false); every function returns a NEGATIVE value
 *     (-U3D_ERR_*) on failure and u3d_last_error() then holds a message for the calling thread.  Nothing throws.
 *   - there is NO CPU fallback: without a usable sm_100 device u3d_ctx_create fails with -U3D_ERR_NO_DEVICE.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context's own stream).  Calls that return results to host
 *     memory synchronise that stream before returning; `_async` / run calls do not.
 *   - matrices are row-major as in the reference's Matrix3x4 / Matrix4 (Matrix3x4::Data(), Matrix4::Data()).
 */
#ifndef U3D_GPU_H
#define U3D_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U3D_API __attribute__((visibility("default")))

enum
{
    U3D_OK = 0,
    U3D_ERR_INVALID = 1,    /* bad argument */
    U3D_ERR_CUDA = 2,       /* a CUDA call failed; message has cudaGetErrorString */
    U3D_ERR_NOMEM = 3,
    U3D_ERR_CAPACITY = 4,   /* an output or pool capacity was exceeded; message says which */
    U3D_ERR_NO_DEVICE = 5,  /* no CUDA device / not sm_100: the product has no CPU path */
    U3D_ERR_UNSUPPORTED = 6
};

/* CullMode, GraphicsDefs.h */
enum { U3D_CULL_NONE = 0, U3D_CULL_CCW = 1, U3D_CULL_CW = 2 };
/* which OctreeQuery the cull evaluates */
enum
{
    U3D_QUERY_FRUSTUM = 0,       /* FrustumOctreeQuery, OctreeQuery.h:140-158 / OctreeQuery.cpp:98-118 */
    U3D_QUERY_OCCLUDED = 1,      /* OccludedFrustumOctreeQuery, View.cpp:116-158 */
    U3D_QUERY_SHADOWCASTER = 2,  /* ShadowCasterOctreeQuery, View.cpp:59-84 */
    /* shape queries (OctreeQuery.h:72-138, OctreeQuery.cpp:32-96); their parameters travel in the 24 floats of `planes`: */
    U3D_QUERY_SPHERE = 3,        /* SphereOctreeQuery: planes[0..3] = centre.xyz, radius (Sphere::IsInside / IsInsideFast, Math/Sphere.cpp:136-256) */
    U3D_QUERY_BOX = 4,           /* BoxOctreeQuery:    planes[0..5] = min.xyz, max.xyz */
    U3D_QUERY_POINT = 5,         /* PointOctreeQuery:  planes[0..2] = point */
    /* ray queries (Octree.cpp:257-309): planes[0..6] = origin.xyz, direction.xyz, maxDistance; every octant incl. the root is tested */
    U3D_QUERY_RAY = 6,           /* RayOctreeQuery walk of Octree::Raycast: drawables with Ray::HitDistance(worldBox) < maxDistance */
    U3D_QUERY_RAY_CANDIDATES = 7,/* RaycastSingle's first pass (GetDrawablesOnlyInternal): all drawables of the octants the ray touches */
    U3D_QUERY_ZONE_OCCLUDER = 8  /* ZoneOccluderOctreeQuery, View.cpp:87-113: frustum walk; a drawable passes when its flags are exactly DRAWABLE_ZONE,
                                    or exactly DRAWABLE_GEOMETRY with IsOccluder() set (u3d_*_set_occluder_flags); drawable_flags is not consulted */
};
/* what is emitted */
enum
{
    U3D_STAGE_QUERY = 0,    /* Octree::GetDrawables result_ (Octree.cpp:495-499), DFS order */
    U3D_STAGE_VISIBLE = 1,  /* ... filtered by CheckVisibilityWork's predicate !occludee || IsVisible(box) (View.cpp:177), same order */
    U3D_STAGE_INVIEW = 2    /* ... and by the draw-distance test (View.cpp:179-186, Drawable::UpdateBatches distance Drawable.cpp:134-137,
                               Camera::GetDistance Camera.cpp:487-496); also reduces the view-space Z range of the geometries that stay
                               (View.cpp:198-211, 926-951): read it with u3d_batch_read_z_ranges */
};

typedef struct u3d_ctx u3d_ctx;             /* one per GPU */
typedef struct u3d_octree u3d_octree;       /* host octree mirror (structure of the reference's Octree) */
typedef struct u3d_scene u3d_scene;         /* flattened octree + drawable AABBs resident in HBM */
typedef struct u3d_mesh u3d_mesh;           /* vertex/index data resident in HBM */
typedef struct u3d_occluders u3d_occluders; /* occluder table resident in HBM */
typedef struct u3d_ob u3d_ob;               /* one occlusion buffer: drop-in for class OcclusionBuffer */
typedef struct u3d_batch u3d_batch;         /* many-view pipeline: raster + hierarchy + cull for V views per call */

/* One view of the batched entry point = the inputs OcclusionBuffer::SetView (OB.cpp:115-127) takes from the Camera plus the
 * query's frustum and filters (OctreeQuery.h:40-70, 140-158). */
typedef struct u3d_view
{
    float view[12];          /* Camera::GetView(), 3x4 */
    float projection[16];    /* Camera::GetProjection(), 4x4 */
    float planes[24];        /* Camera::GetFrustum(): 6 x (normal.xyz, d), order NEAR,LEFT,RIGHT,UP,DOWN,FAR (Frustum.h:37-45) */
    float near_clip, far_clip;
    uint32_t drawable_flags; /* OctreeQuery::drawableFlags_ */
    uint32_t view_mask;      /* OctreeQuery::viewMask_ */
    int32_t reverse_culling; /* Camera::GetReverseCulling() */
    int32_t query_mode;      /* U3D_QUERY_* */
    int32_t use_occlusion;   /* 0: no occlusion buffer for this view (frustum-only query; predicate passes everything) */
    int32_t orthographic;    /* Camera::IsOrthographic(): selects the distance formula of Camera::GetDistance (U3D_STAGE_INVIEW only) */
    float camera_position[3];/* Camera node world position, used by Camera::GetDistance for perspective cameras (U3D_STAGE_INVIEW only) */
    int32_t reserved;
} u3d_view;

U3D_API const char* u3d_last_error(void);
/* Version of this header; u3d_abi_version() returns the library's.  2 = round 2: u3d_batch_set_scene added, U3D_NUM_PHASES 8 -> 10. */
#define U3D_ABI_VERSION 2
U3D_API int u3d_abi_version(void);
/* Test hook / tuning.  "force_irregular" = 1 routes every set-up triangle of the batched path through the general (global-memory) fill.
   "blocking_wait" = 1 makes u3d_batch_cull_views wait on a blocking event instead of spinning in cudaStreamSynchronize: for hosts that run
   more calling threads than cores (8 ranks x 12 batch slots on a 32-core box: 3.5 -> 5.1 M views/s end to end).  "tile_target" = tiles a
   full batch should launch at least (default 3 per SM); buffers / batches created afterwards pick their tile height (64 / 32 / 16 rows)
   from it.  The other names are A/B switches of the kernels (csrc/capi.cu), each documented where it is read. */
U3D_API int u3d_debug_set_option(const char* name, int value);
/* Diagnostic: after u3d_debug_set_option("cull_prof", 1), clock64 sums of the cull kernel's phases over all CTAs (8 slots). */
U3D_API int u3d_debug_read_cull_prof(unsigned long long* out8);

/* ---- context -------------------------------------------------------------------------------------------------------------- */
U3D_API int u3d_ctx_create(int device, u3d_ctx** out);
U3D_API void u3d_ctx_destroy(u3d_ctx* ctx);
U3D_API int u3d_ctx_device(const u3d_ctx* ctx);
U3D_API int u3d_ctx_synchronize(u3d_ctx* ctx, void* stream);

/* ---- host octree mirror (no GPU needed) ----------------------------------------------------------------------------------- */
/* Octree::Octree / SetSize (Octree.cpp:296-353); defaults are +-1000 and 8 levels (Octree.cpp:46-47). */
U3D_API int u3d_octree_create(const float min3[3], const float max3[3], uint32_t levels, u3d_octree** out);
U3D_API void u3d_octree_destroy(u3d_octree* t);
/* Octree::InsertDrawable for n new drawables in order (Octree.cpp:134-166).  aabb6 = n x (min.xyz, max.xyz) world boxes;
 * NULL arrays mean DRAWABLE_GEOMETRY / DEFAULT_VIEWMASK / occludee / no shadows.  Ids are assigned consecutively; *first_id = first. */
U3D_API int u3d_octree_add(u3d_octree* t, uint32_t n, const float* aabb6, const uint8_t* drawable_flags, const uint32_t* view_mask,
    const uint8_t* occludee, const uint8_t* cast_shadows, uint32_t* first_id);
/* Drawable::SetDrawDistance for n drawables starting at first_id (0 = unlimited, the default; used by U3D_STAGE_INVIEW). */
U3D_API int u3d_octree_set_draw_distances(u3d_octree* t, uint32_t first_id, uint32_t n, const float* draw_distance);
/* Drawable::SetOccluder for n drawables starting at first_id (default: not an occluder; used by U3D_QUERY_ZONE_OCCLUDER). */
U3D_API int u3d_octree_set_occluder_flags(u3d_octree* t, uint32_t first_id, uint32_t n, const uint8_t* is_occluder);
/* A moved/resized drawable: the reinsertion rule of Octree::Update (Octree.cpp:440-472).  Returns 1 when the flattened layout did not
 * change (the drawable stayed in its octant: u3d_scene_update_boxes is enough), 0 when it did (re-create the scene). */
U3D_API int u3d_octree_update(u3d_octree* t, uint32_t id, const float aabb6[6]);
/* Octree::RemoveDrawable path (Octree.h:66-74, DecDrawableCount :123-136). */
U3D_API int u3d_octree_remove(u3d_octree* t, uint32_t id);
/* Flatten in the visiting order of Octant::GetDrawablesInternal (Octree.cpp:229-255). */
U3D_API int u3d_octree_counts(const u3d_octree* t, uint32_t* num_octants, uint32_t* num_drawables_in_tree);
U3D_API int u3d_octree_flatten(const u3d_octree* t, int32_t* parent, int32_t* level, float* culling_box6, int32_t* first, int32_t* count,
    int32_t* order);

/* ---- scene ---------------------------------------------------------------------------------------------------------------- */
/* Upload the flattened octree + per-drawable records (32-byte BoundingBox layout, BoundingBox.h:326-330, meta in the pad words). */
U3D_API int u3d_scene_create(u3d_ctx* ctx, const u3d_octree* t, u3d_scene** out);
/* Same from caller-flattened arrays (what an engine integration produces from the real Octree).  Per-drawable arrays are indexed
 * by drawable id; order[i] = id of the drawable in DFS slot i. */
U3D_API int u3d_scene_create_flat(u3d_ctx* ctx, uint32_t num_octants, const int32_t* parent, const int32_t* level, const float* culling_box6,
    const int32_t* first, const int32_t* count, uint32_t num_slots, const int32_t* order, const float* aabb6, const uint8_t* drawable_flags,
    const uint32_t* view_mask, const uint8_t* occludee, const uint8_t* cast_shadows, u3d_scene** out);
/* Per-drawable draw distances (indexed by drawable id, 0 = unlimited) for a scene made with u3d_scene_create_flat. */
U3D_API int u3d_scene_set_draw_distances(u3d_scene* s, uint32_t n, const float* draw_distance);
/* Drawable::IsOccluder() per drawable id for a scene made with u3d_scene_create_flat (in place: one bit of each record's meta word). */
U3D_API int u3d_scene_set_occluder_flags(u3d_scene* s, uint32_t n, const uint8_t* is_occluder);
/* Dirty-range maintenance: new world boxes for n drawables whose octant did not change (u3d_octree_update returned 1); only their
 * 2 x 12 bytes are rewritten in HBM, the layout and the meta words stay.  An id may appear several times (the drawable moved more than
 * once since the last upload): the LAST box of the call wins, as in the host tree. */
U3D_API int u3d_scene_update_boxes(u3d_scene* s, uint32_t n, const uint32_t* ids, const float* aabb6, void* stream);
U3D_API void u3d_scene_destroy(u3d_scene* s);
U3D_API int u3d_scene_counts(const u3d_scene* s, uint32_t* num_octants, uint32_t* num_drawables);

/* ---- occluder geometry ---------------------------------------------------------------------------------------------------- */
/* Vertex data: first 12 bytes of every vertex_size-byte vertex are the position (callers verify SEM_POSITION at offset 0,
 * StaticModel.cpp:216).  index_size 2 or 4, or 0 with indices == NULL for non-indexed geometry. */
U3D_API int u3d_mesh_create(u3d_ctx* ctx, const void* vertices, uint32_t vertex_size, uint32_t vertex_count, const void* indices,
    uint32_t index_size, uint32_t index_count, u3d_mesh** out);
U3D_API void u3d_mesh_destroy(u3d_mesh* m);
/* Occluder table for the batched path: per occluder its world box (selection test), model matrix (3x4), mesh and draw range
 * (= the arguments its DrawOcclusion would pass to AddTriangles, StaticModel.cpp:189-228).  index_start/index_count may be NULL
 * (whole mesh).  cull_mode is what DrawOcclusion passes to SetCullMode; the reference reads a single cullMode_ at draw time
 * (OB.cpp:634), so one mode applies to every batch of a DrawTriangles call. */
U3D_API int u3d_occluders_create(u3d_ctx* ctx, uint32_t n, const float* world_aabb6, const float* model12, u3d_mesh* const* meshes,
    uint32_t num_meshes, const uint32_t* mesh_of_occluder, const uint32_t* index_start, const uint32_t* index_count, int cull_mode,
    u3d_occluders** out);
U3D_API void u3d_occluders_destroy(u3d_occluders* o);

/* ---- OcclusionBuffer drop-in (OB.h:101-158) ------------------------------------------------------------------------------- */
U3D_API int u3d_ob_create(u3d_ctx* ctx, u3d_ob** out);                        /* OcclusionBuffer::OcclusionBuffer */
U3D_API void u3d_ob_destroy(u3d_ob* ob);
U3D_API int u3d_ob_set_size(u3d_ob* ob, int width, int height, int threaded); /* SetSize, OB.cpp:61-113 (1/0) */
U3D_API int u3d_ob_set_view(u3d_ob* ob, const float view12[12], const float projection16[16], float near_clip, float far_clip,
    int reverse_culling);                                                      /* SetView, OB.cpp:115-127 */
U3D_API int u3d_ob_set_max_triangles(u3d_ob* ob, uint32_t triangles);         /* SetMaxTriangles, OB.cpp:129-132 */
U3D_API int u3d_ob_set_cull_mode(u3d_ob* ob, int mode);                       /* SetCullMode, OB.cpp:134-144 */
U3D_API int u3d_ob_reset(u3d_ob* ob);                                         /* Reset, OB.cpp:146-150 */
U3D_API int u3d_ob_clear(u3d_ob* ob);                                         /* Clear, OB.cpp:152-162 */
/* AddTriangles (non-indexed / indexed), OB.cpp:164-198.  Like the reference, only the pointers are stored until DrawTriangles.
 * Counts that are not a multiple of three behave as in DrawBatch (OB.cpp:464-541): non-indexed geometry draws floor(count / 3) triangles;
 * indexed geometry draws ceil(count / 3), reading up to two indices past the stated range (so they must be readable, as for the reference);
 * numTriangles_ grows by count / 3 either way.  The batched occluder table refuses such indexed ranges (-U3D_ERR_UNSUPPORTED). */
U3D_API int u3d_ob_add_triangles(u3d_ob* ob, const float model12[12], const void* vertex_data, uint32_t vertex_size, uint32_t vertex_start,
    uint32_t vertex_count);
U3D_API int u3d_ob_add_triangles_indexed(u3d_ob* ob, const float model12[12], const void* vertex_data, uint32_t vertex_size,
    const void* index_data, uint32_t index_size, uint32_t index_start, uint32_t index_count);
/* Same budget/queue semantics with geometry already resident in HBM (extension; avoids the per-frame upload). */
U3D_API int u3d_ob_add_triangles_mesh(u3d_ob* ob, const float model12[12], const u3d_mesh* mesh, uint32_t start, uint32_t count);
U3D_API int u3d_ob_draw_triangles(u3d_ob* ob);                                /* DrawTriangles, OB.cpp:200-232 */
U3D_API int u3d_ob_build_depth_hierarchy(u3d_ob* ob);                         /* BuildDepthHierarchy, OB.cpp:234-329 */
U3D_API int u3d_ob_is_visible(u3d_ob* ob, const float aabb6[6]);              /* IsVisible, OB.cpp:336-456 (1/0) */
U3D_API int u3d_ob_is_visible_many(u3d_ob* ob, const float* aabb6, uint32_t n, uint8_t* out);
U3D_API int u3d_ob_get_buffer(u3d_ob* ob, int32_t* out_depth);                /* GetBuffer, OB.h:126: width*height ints */
U3D_API int u3d_ob_get_mip(u3d_ob* ob, int level, int32_t* out_minmax);       /* mipBuffers_[level]: (min,max) pairs */
U3D_API int u3d_ob_num_levels(const u3d_ob* ob);
U3D_API int u3d_ob_mip_size(const u3d_ob* ob, int level, int* width, int* height);
U3D_API int u3d_ob_get_width(const u3d_ob* ob);                               /* GetWidth */
U3D_API int u3d_ob_get_height(const u3d_ob* ob);                              /* GetHeight */
U3D_API uint32_t u3d_ob_get_num_triangles(const u3d_ob* ob);                  /* GetNumTriangles */
U3D_API uint32_t u3d_ob_get_max_triangles(const u3d_ob* ob);                  /* GetMaxTriangles */
U3D_API int u3d_ob_get_cull_mode(const u3d_ob* ob);                           /* GetCullMode */
U3D_API int u3d_ob_is_threaded(const u3d_ob* ob);                             /* IsThreaded */
U3D_API int u3d_ob_get_view_proj(const u3d_ob* ob, float out16[16]);          /* viewProj_ (for the 1e-5 screen-bounds check) */

/* Octree::GetDrawables(OctreeQuery&) (Octree.cpp:495-499) with a Frustum / OccludedFrustum / ShadowCaster query, optionally followed by
 * the CheckVisibilityWork predicate.  ob may be NULL (no occlusion).  out_ids receives at most `capacity` drawable ids in result_ order;
 * *out_query_count = size of the query result, *out_count = number emitted for `stage` (may exceed capacity: -U3D_ERR_CAPACITY). */
U3D_API int u3d_octree_query(u3d_scene* scene, u3d_ob* ob, const float planes24[24], uint32_t drawable_flags, uint32_t view_mask, int query_mode,
    int stage, uint32_t* out_ids, uint32_t capacity, uint32_t* out_query_count, uint32_t* out_count);

/* ---- ray queries ----------------------------------------------------------------------------------------- */
/* RayQueryResult (OctreeQuery.h:189-224) at RAY_AABB level, i.e. what Drawable::ProcessRayQuery fills in (Drawable.cpp:118-132):
 * position = origin + distance * direction, normal = -direction, subObject = M_MAX_UNSIGNED (not stored). */
typedef struct u3d_ray_hit
{
    float position[3];
    float normal[3];
    float distance;
    uint32_t drawable;
} u3d_ray_hit;
/* Octree::Raycast (single = 0, Octree.cpp:501-508) or Octree::RaycastSingle (single = 1, Octree.cpp:510-548) for n_rays rays at once, with
 * a RayOctreeQuery of level RAY_AABB.  rays7 = n x (origin.xyz, direction.xyz, maxDistance); direction is used as given (Ray's constructor
 * normalises, Ray::direction_ is what is meant here); maxDistance may be INFINITY.  The octree walk and the hit distances run on the GPU
 * (one CTA per ray); the final ordering uses the reference's own Sort procedure (Container/Sort.h) so ties come out in its order.
 * hits = n x capacity records, counts[i] = result size of ray i.  With single = 1 `capacity` must also hold every candidate drawable of
 * the octants a ray touches (they are sorted to find the winner); -U3D_ERR_CAPACITY otherwise.  Synchronises `stream`. */
U3D_API int u3d_scene_raycast(u3d_scene* scene, uint32_t n_rays, const float* rays7, uint32_t drawable_flags, uint32_t view_mask, int single,
    uint32_t capacity, uint32_t* counts, u3d_ray_hit* hits, void* stream);
/* The reference's Sort (Container/Sort.h:70-141) over float keys with a payload, compare = `key <` (CompareDrawables, Drawable.h:419-422;
 * CompareRayQueryResults, Octree.cpp:66-69).  Host code, no GPU needed; exposed because its tie order is part of observable results. */
U3D_API int u3d_sort_by_key(uint32_t n, float* keys, uint32_t* values);

/* ---- batched many-view entry point (new) ----------------------------------------------------------------------- */
/* Buffers for up to max_views views of width x height.  occluders may be NULL (frustum-only views, BASELINE config 4).
 * out_capacity = result slots per view.  tri_pool = capacity of the set-up triangle pool in triangles (0 = automatic). */
U3D_API int u3d_batch_create(u3d_ctx* ctx, u3d_scene* scene, u3d_occluders* occluders, int width, int height, uint32_t max_views,
    uint32_t out_capacity, uint64_t tri_pool, u3d_batch** out);
U3D_API void u3d_batch_destroy(u3d_batch* b);
/* Re-points the batch at another scene of the same context (the scene object is replaced whenever the flattened octree changes:
 * Octree::InsertDrawable / RemoveDrawable / a move across octants, Octree.cpp:134-190, 440-472).  The batch keeps no pointer into the old
 * scene afterwards, so that one may be destroyed.  Call before u3d_batch_run; per-view buffers and occluders are kept. */
U3D_API int u3d_batch_set_scene(u3d_batch* b, u3d_scene* scene);
/* Host -> device copy of n views (asynchronous on `stream`; the host array is consumed before the call returns). */
U3D_API int u3d_batch_set_views(u3d_batch* b, const u3d_view* views, uint32_t n, void* stream);
/* The whole per-view pipeline for the views last set, all on `stream`, no host synchronisation:
 *   View::DrawOccluders threaded branch (View.cpp:2258-2273): Clear, select + AddTriangles, DrawTriangles, BuildDepthHierarchy;
 *   Octree::GetDrawables(query) (View.cpp:873-878); CheckVisibilityWork predicate (View.cpp:177) when stage == U3D_STAGE_VISIBLE. */
U3D_API int u3d_batch_run(u3d_batch* b, int stage, void* stream);
/* Parts of the pipeline, for measurement: 1 = raster+hierarchy only, 2 = cull only (buffers from a previous run). */
U3D_API int u3d_batch_run_part(u3d_batch* b, int part, int stage, void* stream);
/* u3d_batch_run with a CUDA event after each phase; synchronises and returns the device time of each phase in milliseconds:
 * 0 clear, 1 occluder select (+ near / far split), 2 region scan, 3 triangle set-up (first pass), 4 depth fill + fused hierarchy (first
 * pass), 5 remaining hierarchy levels + the second pass of the two-phase occluder drawing (box test, set-up, fill),
 * 6 octree walk (octant tests), 7 drawable query tests, 8 occlusion predicate over the query's survivors, 9 ordered result emission.
 * The one-CTA-per-view cull kernel (small batches) reports its whole time under 6. */
#define U3D_NUM_PHASES 10
U3D_API int u3d_batch_run_timed(u3d_batch* b, int stage, void* stream, float* phase_ms);
/* Results.  counts = n x 2 uint32: (query result size, emitted count).  Synchronises `stream`; reports pool/output overflow. */
U3D_API int u3d_batch_read_counts(u3d_batch* b, uint32_t* counts, void* stream);
/* Emitted ids of every view packed back to back; offsets has n+1 entries.  capacity in ids.  Synchronises. */
U3D_API int u3d_batch_read_packed(u3d_batch* b, uint32_t* offsets, uint32_t* ids, uint64_t capacity, void* stream);
/* After a run with U3D_STAGE_INVIEW: n x 2 floats (minZ, maxZ) = View::minZ_/maxZ_ of every view (View.cpp:926-951). */
U3D_API int u3d_batch_read_z_ranges(u3d_batch* b, float* minmax, void* stream);
U3D_API int u3d_batch_read_view(u3d_batch* b, uint32_t view, uint32_t* ids, uint32_t capacity, uint32_t* count, void* stream);
U3D_API int u3d_batch_read_depth(u3d_batch* b, uint32_t view, int32_t* out_depth, void* stream);
U3D_API int u3d_batch_read_mip(u3d_batch* b, uint32_t view, int level, int32_t* out_minmax, void* stream);
/* Per view: (submitted triangles, numTriangles_ as the reference counts it = submitted + source triangles that drew, set-up triangles). */
U3D_API int u3d_batch_read_tri_stats(u3d_batch* b, uint32_t* stats3, void* stream);
/* Diagnostic: the set-up triangle records of one view (16 ints each: y0,y1,y2,dzdx, top half lx,lxs,lz,lzs,rx,rxs, bottom half the same),
 * in arbitrary order.  *count = records held. */
U3D_API int u3d_batch_read_setup_records(u3d_batch* b, uint32_t view, int32_t* out16, uint32_t capacity, uint32_t* count, void* stream);
/* One call, host buffers in and out (the e2e path): set_views + run + read_counts + read_packed. */
U3D_API int u3d_batch_cull_views(u3d_batch* b, const u3d_view* views, uint32_t n, int stage, uint32_t* counts, uint32_t* offsets, uint32_t* ids,
    uint64_t capacity, void* stream);
/* Device pointers of the result arrays (for NCCL gathers): counts = max_views x 2 uint32, ids = max_views x out_capacity uint32. */
U3D_API int u3d_batch_device_results(u3d_batch* b, void** counts_dev, void** ids_dev, uint32_t* out_capacity);
/* Pack the emitted ids of all views back to back ON THE DEVICE (two launches on `stream`, no synchronisation) and return the device
 * pointers: offsets = n+1 uint32, packed = sum of counts uint32.  This is what a rank hands to the NCCL gather. */
U3D_API int u3d_batch_pack_device(u3d_batch* b, void* stream, void** offsets_dev, void** packed_dev);
/* Occluder selection by the reference's heuristics instead of "every occluder whose box is in the frustum":
 * View::UpdateOccluders (View.cpp:2171-2229: draw-distance test on Drawable::UpdateBatches' distance, screen-size threshold
 * Renderer::GetOccluderSizeThreshold(), sort key density / size, Sort) followed by the triangle budget of View::DrawOccluders' threaded
 * branch (View.cpp:2258-2270 with OcclusionBuffer::AddTriangles' return value, OB.cpp:164-198: submission stops after the first
 * occluder that takes numTriangles_ past max_triangles = View::maxOccluderTriangles_).  Candidates are the occluder table's entries
 * passing IsInsideFast, in table order (the reference: ZoneOccluderOctreeQuery result order, View.cpp:87-113).
 * camera3 = n x (Camera::GetHalfViewSize(), Camera::GetOrthoSize(), Camera::GetFarClip()) for the views last set; call again after
 * every u3d_batch_set_views.  Where the budget cut separates occluders with EQUAL sort keys the reference's own (unstable) Sort
 * procedure is replayed on the device, so the selected set is the reference's.  At most 4096 candidates per view survive the size
 * threshold (-U3D_ERR_CAPACITY at the next read otherwise).  enable = 0 returns to the default rule.
 * enable = 2 selects the NON-threaded branch of View::DrawOccluders instead (View.cpp:2236-2256; the engine's default, since
 * Renderer::threadedOcclusion_ is off): occluders are drawn one after the other in sorted order, each one after the first only if
 * OcclusionBuffer::IsVisible(its box) holds against what has been drawn so far (pixel scan, the hierarchy being dirty), and the budget sees
 * numTriangles_ = submitted triangles + source triangles that drew something (OB.cpp:196-197, 681-682).  Serial per view, parallel over
 * views; slower than mode 1 but identical to a stock engine configuration.  Equal sort keys always replay the reference's Sort here. */
U3D_API int u3d_batch_set_occluder_policy(u3d_batch* b, int enable, float size_threshold, uint32_t max_triangles, const float* camera3, void* stream);
/* Drawable::GetDrawDistance() of every occluder (0 = unlimited, the default), used by the policy above (View.cpp:2186-2188). */
U3D_API int u3d_occluders_set_draw_distances(u3d_occluders* o, uint32_t n, const float* draw_distance);
/* View::activeOccluders_ of every view after a run in mode 2 (occluders that passed the IsVisible gate and were submitted). */
U3D_API int u3d_batch_read_active_occluders(u3d_batch* b, uint32_t* active, void* stream);
/* Occluders submitted for `view` by the last run, in submission order (diagnostic / tests; synchronises). */
U3D_API int u3d_batch_read_occluder_selection(u3d_batch* b, uint32_t view, uint32_t* occluder_ids, uint32_t capacity, uint32_t* count, void* stream);
/* ---- shadow casters ----
 * One shadow split of View::ProcessShadowCasters (View.cpp:2379-2407).  The frusta are computed by the host with the engine's own code:
 * light_view = shadowCamera->GetView(); shadow_camera_frustum = shadowCamera->GetFrustum() planes (point lights: the cube face);
 * light_view_frustum = planes of cullCamera->GetSplitFrustum(...).Transformed(light_view) (View.cpp:2396-2401);
 * light_view_frustum_max_z = BoundingBox(lightViewFrustum).max_.z_; skip = 1 for a degenerate split frustum (View.cpp:2406-2407). */
typedef struct u3d_shadow_split
{
    float light_view[12];
    float shadow_camera_frustum[24];
    float light_view_frustum[24];
    float light_view_frustum_max_z;
    float shadow_far_clip;        /* shadowCamera->GetFarClip() */
    int32_t shadow_orthographic;  /* shadowCamera->IsOrthographic() */
    int32_t light_type;           /* LightType (Light.h): 0 LIGHT_DIRECTIONAL, 1 LIGHT_SPOT, 2 LIGHT_POINT */
    uint32_t light_mask;          /* light->GetLightMask() */
    int32_t skip;
    int32_t reserved[2];
} u3d_shadow_split;
/* Per-drawable View::GetShadowMask(drawable) (drawable & zone shadow mask, View.h:297-300) and Drawable::GetShadowDistance(), indexed by
 * drawable id; NULL = 0xFFFFFFFF / 0 (unlimited).  Draw distances come from u3d_*_set_draw_distances. */
U3D_API int u3d_scene_set_shadow_params(u3d_scene* s, uint32_t n, const uint32_t* shadow_mask, const float* shadow_distance);
/* The per-drawable loop of View::ProcessShadowCasters + View::IsShadowCasterVisible (View.cpp:2409-2489) over the list every view of the batch
 * emitted in its last run (directional cascades: a U3D_QUERY_SHADOWCASTER view per split, View.cpp:2365-2366; spot / point lights: the
 * light's lit-geometry query, one view per split or cube face), splits[v] describing view v.  Lists are filtered in place, order kept; read
 * them with u3d_batch_read_counts / read_packed.  main_* = the cull camera (Drawable::UpdateBatches' distance for the shadow / draw distance
 * test); in_view[id] != 0 = Drawable::IsInView(frame_) after the main view's visibility pass (perspective lights' shortcut, View.cpp:2464);
 * ids >= n_in_view count as not in view.  Synchronises `stream`. */
U3D_API int u3d_batch_process_shadow_casters(u3d_batch* b, const u3d_shadow_split* splits, const float main_view12[12],
    const float main_camera_position[3], int main_orthographic, const uint8_t* in_view, uint32_t n_in_view, void* stream);
/* Kernel launches issued by the last u3d_batch_run / run_part on this batch. */
U3D_API uint32_t u3d_batch_last_launches(const u3d_batch* b);
/* Hint that the caller keeps several batches in flight on different streams (e.g. one per worker thread, or frame k+1 submitted while frame k
 * drains).  Results are unchanged; kernels pick the launch shape with the best aggregate throughput instead of the lowest latency of a lone
 * batch.  No reference counterpart (the reference runs one view at a time on the main thread, View.cpp:779-809). */
U3D_API int u3d_batch_set_pipelined(u3d_batch* b, int on);

/* ---------------------------------------------------------------------------------------------------------------------------
 * Multi-GPU result exchange.  Views shard over the GPUs of one box, one process (rank) per GPU, the scene is replicated and
 * raster / cull need no communication; the only exchange is the gather of per-view counts and visible sets.  No reference counterpart
 * (the reference renders every view on one CPU, Renderer.cpp:673-686).  Instead of handing packed lists to a library all-gather,
 * u3d_batch_push_results packs the emitted ids of the batch and stores them, in the same kernel, straight into the receive block of
 * EVERY rank through NVLink peer memory, then raises a per-rank arrival flag; consumers wait on the flags on their own stream.  No host
 * synchronisation anywhere.  All ranks must call push / wait / release the same number of times, in the same order.
 *
 * Receive block of a rank (device pointers from u3d_gather_buffers), identical on every rank once an epoch has arrived:
 *   counts  [world][views_per_rank][2]   uint32   (query result size, emitted count) of view slot s of rank r
 *   offsets [world][views_per_rank + 1]  uint32   packed offset of each view inside its rank's ids region; [n] = total
 *   ids     [world][ids_capacity]        uint32   packed emitted ids of rank r
 * With interleaved sharding (view v on rank v % world, slot v / world) view v is found at counts[v % world][v / world]. */
typedef struct u3d_gather u3d_gather;
#define U3D_IPC_HANDLE_BYTES 64
/* Which rank culls which view.  The interleaved rule (view v on rank v % world) leaves the ranks several per cent apart when views differ
 * widely in cost; this gives every rank the same number of views as that rule but equal shares of every cost column at once.
 * costs: n_views x n_costs doubles, row-major -- per-view cost proxies, normally the previous frame's statistics (u3d_batch_read_tri_stats,
 * u3d_batch_read_counts); non-finite or non-positive entries count as 0.  owner_out[v] = rank of view v.  Deterministic, so every rank
 * derives the same answer from the same statistics.  Host code, no GPU needed. */
U3D_API int u3d_balance_views(uint32_t n_views, uint32_t n_costs, const double* costs, uint32_t world, uint32_t* owner_out);
U3D_API int u3d_gather_create(u3d_ctx* ctx, int world, int rank, uint32_t views_per_rank, uint64_t ids_capacity_per_rank, u3d_gather** out);
U3D_API void u3d_gather_destroy(u3d_gather* g);
/* Inter-process set-up: every rank exports a 64-byte handle of its block, the host exchanges them by any means (the Python host uses
 * torch.distributed.all_gather_object) and hands all `world` handles, in rank order, to connect. */
U3D_API int u3d_gather_export(u3d_gather* g, void* handle64);
U3D_API int u3d_gather_connect(u3d_gather* g, const void* handles);
/* Same for exchange objects that live in ONE process (one host thread per GPU, or several logical ranks on one GPU in tests). */
U3D_API int u3d_gather_connect_local(u3d_gather* g, u3d_gather* const* all_ranks);
/* Pack + store to every rank + signal, on `stream` after the batch's run (3 launches: release wait, offset scan, fused push). */
U3D_API int u3d_batch_push_results(u3d_batch* b, u3d_gather* g, void* stream);
/* Make `stream` wait (on the device) until every rank's results of the last pushed epoch have arrived in this rank's block. */
U3D_API int u3d_gather_wait(u3d_gather* g, void* stream);
/* Tell every producer, on `stream`, that this rank has consumed the epoch it last waited for (its regions may be overwritten). */
U3D_API int u3d_gather_release(u3d_gather* g, void* stream);
U3D_API int u3d_gather_buffers(u3d_gather* g, void** counts_dev, void** offsets_dev, void** ids_dev, uint32_t* views_per_rank,
    uint64_t* ids_capacity_per_rank);
/* Host copies of the receive block (any pointer may be NULL); synchronises `stream` and reports like u3d_gather_status. */
U3D_API int u3d_gather_read(u3d_gather* g, uint32_t* counts, uint32_t* offsets, uint32_t* ids, void* stream);
/* Synchronises `stream`; -U3D_ERR_CAPACITY if some rank's ids region overflowed, -U3D_ERR_CUDA if a wait timed out (a peer died). */
U3D_API int u3d_gather_status(u3d_gather* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* U3D_GPU_H */
