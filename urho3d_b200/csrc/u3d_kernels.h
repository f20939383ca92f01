// Host-visible declarations of the kernel launchers (raster.cu, cull.cu) used by the C-ABI layer (capi.cu).
#pragma once
#include "u3d_device.cuh"

namespace u3d
{

constexpr int kItemTris = 256;      // triangles per draw item; bigger batches are split so no block owns a huge mesh
constexpr int kSetupItems = 64;     // draw items per set-up slice (64 boxes = 768 triangles = 3 full sweeps of 256 threads)
constexpr int kSetupThreads = 256;
constexpr int kSetupQueueBytes = 2 * kSetupThreads * 11 * 4; // survivor queue of k_setup_triangles
constexpr int kClipEntryFloats = 16;  // clipper queue entry: three clip-space vertices, plane mask, view
constexpr int kCullSmallBatch = 640;  // views per launch up to which the one-CTA-per-SM cull variant is used

// bits of the per-batch device flag word
constexpr uint32_t kFlagTriOverflow = 1u;   // a view's triangle region was too small
constexpr uint32_t kFlagPoolOverflow = 2u;  // the triangle pool cannot hold all regions
constexpr uint32_t kFlagOutOverflow = 4u;   // a view produced more results than the output capacity
constexpr uint32_t kFlagIrrOverflow = 8u;   // more irregular triangles than the general-path list holds
constexpr uint32_t kFlagRankOverflow = 32u;  // a view has more occluder candidates than the ranked selection holds (kRankCap)
constexpr int kRankCap = 4096;               // occluder candidates per view in k_select_ranked's shared-memory sort
constexpr uint32_t kFlagClipOverflow = 16u;  // more plane-crossing triangles than the clipper queue holds

// tile path
constexpr int kTileH = 64;          // tile = min(width, 256) x (at most) 64 pixels of int depth in shared memory (64 KiB for 256 wide)
constexpr int kTileThreads = 512;
constexpr uint32_t kClsLane = 0;      // bbox area <= 64: one lane rasterises it
constexpr uint32_t kClsWarp = 1;      // bbox area <= 16384: one warp, laid out as rows x lanes-per-row
constexpr uint32_t kClsCta = 2;       // larger: the whole CTA, rows interleaved over the warps
constexpr int kNumCls = 3;            // every tile has one bin per class
constexpr int kLaneAreaMax = 256;     // bbox area up to which one lane rasterises a triangle
constexpr int kWarpAreaMax = 2048;    // ... up to which one warp does; larger ones take the whole CTA
constexpr int kWarpGroup = 8;         // warp-class triangles handed to a warp at a time

struct TileGeom
{
    int tileW, tilesX, tilesY, numTiles;
    int tileH, tileHLog; // rows per tile: kTileH, or 32 / 16 for batches of few views (make_tile_geom)
    int tileWLog; // tileW is a power of two (min(width, 256), widths are powers of two): index arithmetic by shifts
    int fused;   // hierarchy levels reduced inside the tile kernel
};

/// Device pointers of the rasteriser's per-batch state.
struct RasterDev
{
    TriSetup* pool;
    const uint64_t* triBase;     // per view: first record of its region in pool
    const uint32_t* triCap;      // per view: region capacity
    uint32_t* triCount;          // per view: records emitted
    uint32_t* drawnSources;      // per view: source triangles that drew (numTriangles_ bookkeeping)
    uint32_t* flags;
    uint32_t* binIdx;            // per view: numTiles x kNumCls x triCap entries (slot | hint << 29); nullptr = general path only
    uint32_t* binCount;          // numViews x numTiles x kNumCls
    unsigned long long* irrList; // (view << 32 | slot) of triangles the tile path refuses
    uint32_t irrCap;
    uint32_t* irrTotal;
    uint32_t* irrOfView;         // per view
    float* clipQueue;            // kClipEntryFloats per triangle that needs the clipper
    uint32_t clipCap;
    uint32_t* clipCount;
    int writeMips;
    int forceIrregular;          // test hook: send every triangle through the general path
    int laneAreaMax, warpAreaMax; // bbox-area limits of the lane and warp size classes
    // two-phase occluder drawing (launch_items_near_far / launch_items_far_test): a set-up launch handles the items whose itemPass equals
    // `pass` (itemPass == nullptr: all of them)
    const unsigned char* itemPass; // per (view, item): kPassNear, kPassFarVisible, kPassFarHidden
    int pass;
    // ... or, when set, exactly the items listed (positions in the view's item array): the near / far-visible lists are short, and a block
    // that sweeps all items to find a handful is what the set-up launch would otherwise spend its time on
    const uint32_t* itemList;    // per view: itemCap positions
    const uint32_t* listCount;   // per view
    int countOnly;               // set-up decides drawOk per source triangle (numTriangles_, OB.cpp:681-682) but emits nothing
    int secondPass;              // the tile kernel composites on top of the plane it finds and leaves untouched tiles alone
};
constexpr unsigned char kPassNear = 1, kPassFarVisible = 2, kPassFarHidden = 3;

/// Flattened octree + drawables resident in HBM (DFS pre-order = the order Octant::GetDrawablesInternal visits, Octree.cpp:229-255).
struct SceneDev
{
    int numOctants;
    int numDrawables;
    int numLevels;                 // deepest octant level + 1
    const float4* octMin;          // xyz = cullingBox_.min_, w = parent (int bits; -1 for the root)
    const float4* octMax;          // xyz = cullingBox_.max_, w = first drawable (int bits)
    const int* octCount;           // drawables owned by the octant
    const int2* octChild;          // x = position of the first child in bfs[], y = number of children (children are contiguous there)
    const int* bfs;                // octant indices sorted by level (stable)
    const int* levelStart;         // numLevels + 1 offsets into bfs
    const float4* drwMin;          // xyz = world AABB min, w = drawableFlags | kMetaOccludee | kMetaCastShadows (uint bits)
    const float4* drwMax;          // xyz = world AABB max, w = viewMask (uint bits)
    const uint32_t* drwId;         // caller's drawable id for each DFS slot
    const float* drwDrawDist;      // Drawable::GetDrawDistance per DFS slot (0 = unlimited)
    // Static summaries of the drawables' filter words, per "word" = 32 consecutive drawables of one octant (word k of octant o has index
    // octWordStart[o] + k): x = drawableFlags if all of the word's drawables have the same flags AND view mask, else 0xFFFFFFFF; y = that view
    // mask; z = occludee bits; w = castShadows bits.  A word of an octant that is INSIDE the query volume needs nothing else from HBM.
    const uint4* wordInfo;
    const uint32_t* octWordStart;  // numOctants + 1
};

void launch_clear(int* depth, size_t nInts, cudaStream_t s);
void launch_select(const ViewConst* views, int numViews, int numOcc, const float* occAabb6, const uint32_t* occMesh, const uint32_t* occStart,
    const uint32_t* occCount, DrawItem* items, uint32_t itemCap, uint32_t* itemCount, uint32_t* submitted, cudaStream_t s);
/// View::UpdateOccluders heuristics + DrawOccluders' triangle budget per view (camParams: halfViewSize, 1 / orthoSize, farClip per view).
cudaError_t init_rank_kernel();
void launch_select_ranked(const ViewConst* views, const float4* camParams, int numViews, int numOcc, const float* occAabb6, const float* occDrawDist,
    const uint32_t* occMesh, const uint32_t* occStart, const uint32_t* occCount, float sizeThreshold, uint32_t maxTriangles, DrawItem* items,
    uint32_t itemCap, uint32_t* itemCount, uint32_t* submitted, uint32_t* flags, uint32_t* rankList /* null: threaded-branch budget + draw items;
    else the full sorted list, kRankCap entries per view */, uint32_t* rankCount, cudaStream_t s);
/// View::DrawOccluders' non-threaded branch over the sorted lists of launch_select_ranked (one CTA per view, general-path fill).
void launch_draw_occluders_serial(const BufGeom& g, const ViewConst* views, int numViews, const uint32_t* rankList, const uint32_t* rankCount,
    const float* occAabb6, const float* models, const uint32_t* occMesh, const uint32_t* occStart, const uint32_t* occCount, const MeshDev* meshes,
    uint32_t maxTriangles, int* depth, TriSetup* scratch, uint32_t scratchCap, uint32_t* submitted, uint32_t* drawnSources, uint32_t* activeOccluders,
    uint32_t* flags, cudaStream_t s);
void launch_scan_regions(int numViews, const uint32_t* submitted, uint32_t slack, uint64_t poolCap, uint64_t* triBase, uint32_t* triCap, uint32_t* flags,
    cudaStream_t s);
void launch_setup(const BufGeom& g, const TileGeom& tg, const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, uint32_t maxItems,
    const uint32_t* itemCount, const float* models, const MeshDev* meshes, const RasterDev& rd, cudaStream_t s);
void launch_fill_global(const BufGeom& g, int numViews, int* depth, const TriSetup* pool, const uint64_t* triBase, const uint32_t* triCap,
    const uint32_t* triCount, int blocksPerView, cudaStream_t s);
TileGeom make_tile_geom(const BufGeom& g, uint32_t maxViews, uint32_t targetTiles);
cudaError_t init_tile_kernel();
/// Two-phase occluder drawing.  An occluder whose box fails OcclusionBuffer::IsVisible against a buffer that holds only SOME of the
/// occluders cannot change the final plane either (every pixel of its rect is already nearer than its nearest point), so a view first
/// draws its nearest occluders, tests the boxes of the others against that buffer, and draws only those that pass: same plane, same
/// hierarchy, a fraction of the set-up and fill work (on the benchmark scene ~11 % of the selected occluders are drawn).
/// near_far: per view, the items of the nearest occluders (by clip-space w of the box centre; at least minNear items or 1 / 2^shift of
/// them) get kPassNear, the rest kPassFarHidden.  far_test: items marked kPassFarHidden whose occluder box is visible in the buffers drawn
/// so far become kPassFarVisible.
void launch_items_near_far(const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, const uint32_t* itemCount, const float* occAabb6,
    unsigned char* itemPass, uint32_t* nearList /* near items from the front, far items from the back */, uint32_t* nearCount, uint32_t* farCount,
    uint32_t* farVisCount /* zeroed here */, uint32_t minNear, int shift, cudaStream_t s);
void launch_items_far_test(const BufGeom& g, const ViewConst* views, int numViews, const DrawItem* items, uint32_t itemCap, uint32_t maxItems,
    const uint32_t* nearList, const uint32_t* farCount, const float* occAabb6, const int* depth, const int2* mips, unsigned char* itemPass,
    uint32_t* farVisList, uint32_t* farVisCount, cudaStream_t s);
/// irregular pre-pass (2 launches) + tile kernel (1 launch)
void launch_fill_tiles(const BufGeom& g, const TileGeom& tg, int numViews, int* depth, int2* mips, const RasterDev& rd, cudaStream_t s);
void launch_hierarchy(const BufGeom& g, int numViews, const int* depth, int2* mips, int firstLevel, cudaStream_t s);

cudaError_t init_cull_kernels();
/// Octree query (+ optional CheckVisibility predicate) for every view, one CTA per view.  octState: numViews x stateStride bytes of scratch.
/// onlyIf != nullptr: the kernel runs only when *onlyIf is non-zero on the device (fallback behind launch_cull_dense); octState is then
/// not cleared here (what the dense path left in it is what this kernel would write itself).
void launch_cull(const SceneDev& sc, const BufGeom& g, const ViewConst* views, int numViews, const int* depth, const int2* mips, int mipsValid,
    unsigned char* octState, int stateStride, int stage, uint32_t* outIdx, uint32_t outCap, uint32_t* outCounts /* [V][2]: query result size, emitted */,
    float* zRange /* [V][2] minZ, maxZ (STAGE_INVIEW; may be null) */, uint32_t* flags,
    bool pipelined /* other launches share the GPU: prefer the shape with the best aggregate throughput */, const uint32_t* onlyIf, cudaStream_t s);

/// Scratch of the batch-wide cull path (cull_dense.cu).  ctr = counter block: [0..31] queued (view, octant) pairs per octree level,
/// [32] groups with work, [33] words handed out, [34] overflow -> fall back to launch_cull; the per-view alive bits follow it.
struct DenseDev
{
    uint2* queue;            // (view, octant | inside << 31), level after level
    uint32_t queueCap;
    uint32_t* ctr;
    uint32_t* aliveBits;     // per view: one bit per group of 32 octants that holds a surviving octant with drawables
    int aliveWords;          // words per view = ceil(numOctants / 1024)
    uint32_t* groupCount;    // per view and group: result words (32 drawables each) its surviving octants need
    uint32_t* groupOff;      // exclusive scan of groupCount within the view
    int groupsPerView;       // ceil(numOctants / 32)
    uint2* groupList;        // (view, group) of every group with work
    uint32_t* wordBase;      // per view: first word of its region in the pool
    uint32_t* viewWords;     // per view: words in use
    uint32_t* visWord;       // pool: one bit per drawable that is in the result
    uint32_t* wordSlot;      // pool: DFS slot of a word's bit 0
    uint32_t* wordMeta;      // pool: view | (valid drawables - 1) << 16 | inside << 21
    uint32_t* testMask;      // pool: drawables that passed the query and still need the occlusion predicate
    uint32_t* wordStatic;    // pool: index of the word's static summary (SceneDev::wordInfo)
    uint32_t poolCap;
    int stateStride;         // bytes of octant state per view (groupsPerView * 32)
    uint32_t epoch;          // 1..65535, tags the queue entries of this run (k_dense_tree)
};
size_t dense_counter_bytes(int numViews, int numOctants);
int dense_launch_count(const SceneDev& sc);
void dense_set_heavy_extent(int px);
void dense_set_walk_pool(int on);
void dense_set_walker(int on);
void dense_set_tree(int on);
void dense_set_frustum_bulk(int on);
void dense_set_tree_blocks(int perSm);
void dense_set_walk_budget(int busy, int idle);
void dense_prof_enable(int on);
void dense_prof_read(unsigned long long* out8);
void dense_prof_read_levels(unsigned long long* out128);
/// Same results as launch_cull, all views pooled into dense launches (no per-view CTA, no block barriers).  Needs numLevels <= 30.
/// If a queue or the word pool overflows, ctr[34] is raised on the device and every later kernel of the sequence returns at once:
/// the caller queues launch_cull(..., onlyIf = ctr + 34) behind it.
void launch_cull_dense(const SceneDev& sc, const BufGeom& g, const ViewConst* views, int numViews, const int* depth, const int2* mips, int mipsValid,
    unsigned char* octState, int stage, uint32_t* outIdx, uint32_t outCap, uint32_t* outCounts, float* zRange, uint32_t* flags, const DenseDev& dd,
    cudaStream_t s, cudaEvent_t* phaseEvents = nullptr /* 3 events: after the octree walk, after the drawable tests, after the occlusion pass */,
    bool pipelined = false /* other batches are in flight on other streams: leave them room */);

/// Ray::HitDistance of every drawable a QUERY_RAY / QUERY_RAY_CANDIDATES run emitted (slotOfId: drawable id -> DFS slot).
void launch_ray_distances(const SceneDev& sc, const ViewConst* views, int numRays, const uint32_t* slotOfId, const uint32_t* outIdx, uint32_t outCap,
    const uint32_t* outCounts, float* outDist, cudaStream_t s);

/// View::ProcessShadowCasters over the lists the views emitted (one split per view), in place.
void launch_shadow_casters(const SceneDev& sc, const ShadowSplitDev* splits, int numViews, const uint32_t* slotOfId, const uint32_t* shadowMask,
    const float* shadowDistance, const unsigned char* inView, uint32_t numInView, float4 mainPos, float4 mainViewZ, uint32_t* outIdx, uint32_t outCap,
    uint32_t* outCounts, cudaStream_t s);
void launch_set_occluder_bits(float4* drwMin, const unsigned char* bySlot, uint32_t n, cudaStream_t s);
void launch_update_boxes(float4* drwMin, float4* drwMax, const uint32_t* slots, const float* boxes, uint32_t n, cudaStream_t s);

/// Diagnostic: in-kernel phase timers of k_cull_views (clock64 sums over all CTAs; slots: 0 prologue, 1 octant sweeps, 2 octant tests,
/// 3 chunk set-up, 4 frustum rounds, 5 drawable occlusion tests + emission).
void cull_prof_enable(int on);
void cull_set_heavy_extent(int px);
void cull_set_walk_pool(int on);
void cull_set_ring(int entries);
void cull_prof_read(unsigned long long* out8);

/// OcclusionBuffer::IsVisible for n boxes against view 0's buffer.
void launch_is_visible(const BufGeom& g, const ViewConst* view, const int* depth, const int2* mips, int mipsValid, const float* aabb6, int n,
    unsigned char* out, cudaStream_t s);

} // namespace u3d
