// C-ABI layer of libu3dgpu.so (include/u3d_gpu.h): object lifetimes, device memory, host<->device staging and kernel sequencing.
// No CPU fallback anywhere: every compute entry point launches the sm_100a kernels of raster.cu / cull.cu or fails.
#include "../../include/u3d_gpu.h"
#include "u3d_host.h"
#include "u3d_kernels.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <map>
#include <new>
#include <string>
#include <vector>

using namespace u3d;

static thread_local std::string g_err;
static int g_forceIrregular = 0; // u3d_debug_set_option("force_irregular")
static int g_laneAreaMax = kLaneAreaMax, g_warpAreaMax = kWarpAreaMax; // tuning hooks "lane_area_max" / "warp_area_max"
// Which cull path a batch takes ("cull_dense"): 0 = one CTA per view always, 1 = the batch-wide dense path from kDenseMinViews views on,
// 2 = the dense path always.  "dense_queue_cap" / "dense_pool_cap" shrink its scratch (test hook for the on-device fallback).
static int g_cullDense = 1;
static uint32_t g_denseQueueCap = 0, g_densePoolCap = 0;
// "two_phase": 1 = the batched path draws a view's nearest occluders first and the others only if their box shows in that buffer
// (launch_items_near_far / launch_items_far_test), 0 = every selected occluder is set up and filled.  "two_phase_min" / "two_phase_shift":
// the near set holds at least that many draw items, or 1 / 2^shift of the view's items.
static int g_twoPhase = 1, g_twoPhaseMin = 32, g_twoPhaseShift = 4;
static int g_blockingWait = 0; // u3d_batch_cull_views waits on a blocking event instead of spinning in cudaStreamSynchronize ("blocking_wait")
static int g_tileTarget = 148 * 3; // tiles a full batch should launch at least (make_tile_geom halves the tile height below it); "tile_target"
static int g_readDenseProf = 0; // "read_dense_prof": u3d_debug_read_cull_prof returns the dense path's timers
constexpr int kDenseMinViews = 16;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return -code;
}

#define CU_CHECK(expr)                                                                                                   \
    do                                                                                                                   \
    {                                                                                                                    \
        cudaError_t e_ = (expr);                                                                                         \
        if (e_ != cudaSuccess)                                                                                           \
            return fail(U3D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                               \
    } while (0)

namespace
{

/// Owning device buffer (grow-only helper).
template <class T> struct DevBuf
{
    T* p{};
    size_t n{};
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n)
            return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess)
            n = count;
        return e;
    }
};

template <class T> struct PinnedBuf
{
    T* p{};
    size_t n{};
    ~PinnedBuf()
    {
        if (p)
            cudaFreeHost(p);
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess)
            n = count;
        return e;
    }
};

BufGeom make_geom(int w, int h)
{
    BufGeom g{};
    g.w = w;
    g.h = h;
    int off = 0;
    int lw = w, lh = h;
    for (;;) // OB.cpp:97-106
    {
        lw = (lw + 1) / 2;
        lh = (lh + 1) / 2;
        g.mw[g.nlev] = lw;
        g.mh[g.nlev] = lh;
        g.moff[g.nlev] = off;
        off += lw * lh;
        ++g.nlev;
        if ((lw <= kMinMip && lh <= kMinMip) || g.nlev == kMaxLevels)
            break;
    }
    g.mipTexels = off;
    g.sx = 0.5f * w; // CalculateViewport, OB.cpp:578-587
    g.sy = -0.5f * h;
    g.ox = 0.5f * w + 0.5f;
    g.oy = 0.5f * h + 0.5f;
    return g;
}

/// viewProj_ = projection_ * view_  (Matrix4 * Matrix3x4, Matrix4.cpp:43-63).  Host code is built with -ffp-contract=off.
void mul_proj_view(const float* a, const float* b, float* o)
{
    for (int r = 0; r < 4; ++r)
    {
        const float* ar = a + 4 * r;
        for (int c = 0; c < 3; ++c)
            o[4 * r + c] = ar[0] * b[c] + ar[1] * b[4 + c] + ar[2] * b[8 + c];
        o[4 * r + 3] = ar[0] * b[3] + ar[1] * b[7] + ar[2] * b[11] + ar[3];
    }
}

int flip_cull(int mode, bool reverse) // SetCullMode, OB.cpp:134-144
{
    if (reverse)
    {
        if (mode == CULL_CW)
            return CULL_CCW;
        if (mode == CULL_CCW)
            return CULL_CW;
    }
    return mode;
}

} // namespace

struct u3d_ctx
{
    int device{};
    cudaStream_t stream{};
    cudaStream_t pick(void* s) const { return s ? (cudaStream_t)s : stream; }
};

struct u3d_octree
{
    HostOctree tree;
    u3d_octree(const float* mn, const float* mx, unsigned lv) : tree(mn, mx, lv) {}
};

struct u3d_scene
{
    u3d_ctx* ctx{};
    SceneDev dev{};
    DevBuf<float4> octMin, octMax, drwMin, drwMax;
    DevBuf<int> octCount, bfs, levelStart;
    DevBuf<int2> octChild;
    DevBuf<uint32_t> drwId;
    DevBuf<float> drwDrawDist;       // per DFS slot, 0 = unlimited
    DevBuf<uint4> wordInfo;          // static filter summaries per 32 consecutive drawables of an octant (SceneDev::wordInfo)
    DevBuf<uint32_t> octWordStart;
    DevBuf<uint32_t> updSlots;       // staging of u3d_scene_update_boxes
    DevBuf<float> updBoxes;
    std::vector<uint32_t> slotOfId;  // drawable id -> DFS slot (for late per-drawable uploads)
    // View::ProcessShadowCasters inputs (u3d_scene_set_shadow_params / u3d_batch_process_shadow_casters)
    DevBuf<uint32_t> shadowMask;      // per DFS slot; empty = 0xFFFFFFFF everywhere
    DevBuf<float> shadowDistance;     // per DFS slot; empty = 0 (unlimited)
    bool hasShadowParams{};
    DevBuf<unsigned char> inView;     // per drawable id
    DevBuf<ShadowSplitDev> splits;
    // ray queries (u3d_scene_raycast): device copy of slotOfId + scratch, all grow-only
    DevBuf<uint32_t> slotOfIdDev;
    bool slotOfIdUploaded{};
    DevBuf<ViewConst> rayViews;
    DevBuf<unsigned char> rayOctState;
    DevBuf<uint32_t> rayIdx, rayCounts, rayFlags;
    DevBuf<float> rayDist;
    std::vector<ViewConst> rayHost;
    std::vector<uint32_t> rayHostIdx, rayHostCounts;
    std::vector<float> rayHostDist;
};

struct u3d_mesh
{
    u3d_ctx* ctx{};
    DevBuf<unsigned char> vtx;
    DevBuf<unsigned char> idx;
    uint32_t stride{}, vertexCount{}, idxSize{}, idxCount{};
    MeshDev dev() const { return MeshDev{vtx.p, idxSize ? (const void*)idx.p : nullptr, stride, idxSize}; }
};

struct u3d_occluders
{
    u3d_ctx* ctx{};
    uint32_t n{};
    int cullMode{CULL_CCW};
    uint32_t totalItems{};
    uint64_t totalTris{};
    DevBuf<float> aabb6, model12;
    DevBuf<float> drawDist;           // Drawable::GetDrawDistance per occluder (u3d_occluders_set_draw_distances); empty = unlimited
    DevBuf<uint32_t> mesh, start, count;
    DevBuf<MeshDev> meshTable;
};

/// Device state shared by the single-buffer object and the batched pipeline.
struct ViewSet
{
    u3d_ctx* ctx{};
    BufGeom g{};
    uint32_t maxViews{};
    DevBuf<ViewConst> views;
    DevBuf<int> depth;
    DevBuf<int2> mips;
    // Counters that start every run at zero live in ONE block (one cudaMemsetAsync per run instead of a dozen): the members below point into it.
    struct Sub
    {
        uint32_t* p{};
        size_t n{};
    };
    DevBuf<uint32_t> zeroBlock;
    size_t zeroWords{}, passLocalOffset{};
    Sub itemCount, submitted, triCount, drawnSources, flags, binCount, irrTotal, irrOfView, clipCount;
    DevBuf<uint32_t> triCap;
    DevBuf<uint64_t> triBase;
    DevBuf<TriSetup> pool;
    DevBuf<DrawItem> items;
    uint32_t itemCap{};
    // tile path
    TileGeom tg{};
    bool useTiles{};
    DevBuf<uint32_t> binIdx;
    DevBuf<unsigned long long> irrList;
    uint32_t irrCap{};
    DevBuf<float> clipQueue;
    uint32_t clipCap{};
    cudaError_t clear_counters(cudaStream_t st) { return cudaMemsetAsync(zeroBlock.p, 0, zeroWords * 4, st); }
    cudaError_t clear_pass_counters(cudaStream_t st) { return cudaMemsetAsync(zeroBlock.p + passLocalOffset, 0, (zeroWords - passLocalOffset) * 4, st); }
    RasterDev raster_dev(bool tiles, int writeMips)
    {
        RasterDev rd{};
        rd.pool = pool.p;
        rd.triBase = triBase.p;
        rd.triCap = triCap.p;
        rd.triCount = triCount.p;
        rd.drawnSources = drawnSources.p;
        rd.flags = flags.p;
        rd.binIdx = tiles ? binIdx.p : nullptr;
        rd.binCount = tiles ? binCount.p : nullptr;
        rd.irrList = irrList.p;
        rd.irrCap = irrCap;
        rd.irrTotal = irrTotal.p;
        rd.irrOfView = irrOfView.p;
        rd.clipQueue = clipQueue.p;
        rd.clipCap = clipCap;
        rd.clipCount = clipCount.p;
        rd.writeMips = writeMips;
        rd.forceIrregular = g_forceIrregular;
        rd.laneAreaMax = g_laneAreaMax;
        rd.warpAreaMax = g_warpAreaMax;
        return rd;
    }
    int alloc(u3d_ctx* c, int w, int h, uint32_t nviews, uint32_t itemCapPerView, uint64_t poolTris, bool tiles = false)
    {
        ctx = c;
        g = make_geom(w, h);
        tg = make_tile_geom(g, nviews, (uint32_t)std::max(g_tileTarget, 0));
        useTiles = tiles && g.w >= 4;
        maxViews = nviews;
        itemCap = itemCapPerView;
        if (useTiles)
        {
            irrCap = 1u << 20;
            CU_CHECK(binIdx.ensure(std::max<uint64_t>(poolTris, 1) * tg.numTiles * kNumCls));
            CU_CHECK(irrList.ensure(irrCap));
        }
        {
            // [flags, pad x3 | itemCount | submitted | triCount | drawnSources || clipCount, irrTotal, pad x2 | irrOfView | binCount]: the part
            // behind || belongs to one set-up + fill pass and is cleared again between the two passes of the two-phase occluder drawing
            const size_t nBins = useTiles ? (size_t)nviews * tg.numTiles * kNumCls : 0;
            passLocalOffset = 4 + 4 * (size_t)nviews;
            zeroWords = passLocalOffset + 4 + (size_t)nviews + nBins;
            CU_CHECK(zeroBlock.ensure(zeroWords));
            uint32_t* q = zeroBlock.p;
            flags = Sub{q, 1};
            q += 4;
            itemCount = Sub{q, nviews};
            submitted = Sub{q + nviews, nviews};
            triCount = Sub{q + 2 * (size_t)nviews, nviews};
            drawnSources = Sub{q + 3 * (size_t)nviews, nviews};
            q = zeroBlock.p + passLocalOffset;
            clipCount = Sub{q, 1};
            irrTotal = Sub{q + 1, 1};
            q += 4;
            irrOfView = Sub{q, nviews};
            binCount = Sub{q + nviews, nBins};
        }
        CU_CHECK(views.ensure(nviews));
        CU_CHECK(depth.ensure((size_t)nviews * w * h));
        CU_CHECK(mips.ensure((size_t)nviews * g.mipTexels));
        CU_CHECK(triCap.ensure(nviews));
        CU_CHECK(triBase.ensure(nviews));
        CU_CHECK(items.ensure((size_t)nviews * std::max<uint32_t>(itemCap, 1)));
        CU_CHECK(pool.ensure(std::max<uint64_t>(poolTris, 1)));
        // triangles that cross a clip plane are a fraction of the submitted ones; an overflow is reported, not dropped silently
        clipCap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(poolTris / 4, 4096), 0x7FFFFFFFu);
        CU_CHECK(clipQueue.ensure((size_t)clipCap * kClipEntryFloats));
        return 0;
    }
};

struct HostBatchRec
{
    float model[12];
    const void* vtx;
    uint32_t vtxSize;
    const void* idx;
    uint32_t idxSize;
    uint32_t start, count;
    const u3d_mesh* mesh;
};

struct u3d_ob
{
    u3d_ctx* ctx{};
    ViewSet vs;
    bool allocated{};
    int width{}, height{};
    bool threaded{};
    float view[12]{}, proj[16]{}, viewProj[16]{};
    float nearClip{}, farClip{};
    bool reverseCulling{};
    int cullMode{CULL_CCW};
    uint32_t numTriangles{}, maxTriangles{5000}; // OCCLUSION_DEFAULT_MAX_TRIANGLES, OB.h:84
    bool hierarchyDirty{true};
    std::vector<HostBatchRec> batches;
    // per-draw staging
    DevBuf<unsigned char> geom;
    DevBuf<float> models;
    DevBuf<MeshDev> meshTable;
    DevBuf<float> boxes;
    DevBuf<unsigned char> visOut;
    // query scratch
    DevBuf<unsigned char> octState;
    DevBuf<uint32_t> outIdx, outCounts, queryFlags;
};

/// Scratch of the batch-wide cull path (cull_dense.cu); sized from the scene and the batch's view capacity, grow-only.
struct DenseScratch
{
    DevBuf<uint2> queue, groupList;
    DevBuf<uint32_t> ctr, groupCount, groupOff, wordBase, viewWords, visWord, wordSlot, wordMeta, testMask, wordStatic;
    DenseDev dev{};
    uint32_t epochCounter{};
    /// Epoch of the next run's queue entries; the queue is wiped whenever the 16-bit tag starts over (and before its first use), so an
    /// entry left by an earlier run can never carry the current tag.
    cudaError_t next_epoch(cudaStream_t st)
    {
        cudaError_t e = cudaSuccess;
        if (epochCounter % 65535u == 0)
            e = cudaMemsetAsync(queue.p, 0, queue.n * sizeof(uint2), st);
        dev.epoch = epochCounter % 65535u + 1u;
        ++epochCounter;
        return e;
    }
    int ensure(uint32_t maxViews, const SceneDev& sc, uint32_t queueCapOverride, uint32_t poolCapOverride)
    {
        const uint64_t V = maxViews, M = (uint64_t)sc.numOctants, N = (uint64_t)sc.numDrawables;
        const int groups = (int)((M + 31) / 32), aliveWords = (int)((M + 1023) / 1024);
        // worst cases: every octant of every view queued once; every drawable of every view in a word of its own octant.
        // The defaults hold a good part of that; an overflow is caught on the device and answered by the one-CTA-per-view kernel.
        const uint64_t queueWorst = V * M, poolWorst = V * ((N + 31) / 32 + std::min(M, N));
        uint64_t queueCap = std::min<uint64_t>(queueWorst, std::max<uint64_t>(1u << 20, queueWorst / 2));
        uint64_t poolCap = std::min<uint64_t>(poolWorst, std::max<uint64_t>(1u << 20, poolWorst / 4));
        queueCap = std::min<uint64_t>(queueCap, 0x7FFFFFFFu);
        poolCap = std::min<uint64_t>(poolCap, 0x7FFFFFFFu);
        if (queueCapOverride)
            queueCap = queueCapOverride;
        if (poolCapOverride)
            poolCap = poolCapOverride;
        const uint2* oldQueue = queue.p;
        CU_CHECK(queue.ensure(queueCap));
        if (queue.p != oldQueue)
            epochCounter = 0; // fresh memory: wiped below before its first use
        CU_CHECK(ctr.ensure(dense_counter_bytes((int)V, (int)M) / 4));
        CU_CHECK(groupCount.ensure(V * groups));
        CU_CHECK(groupOff.ensure(V * groups));
        CU_CHECK(groupList.ensure(V * groups));
        CU_CHECK(wordBase.ensure(V));
        CU_CHECK(viewWords.ensure(V));
        CU_CHECK(visWord.ensure(poolCap));
        CU_CHECK(wordSlot.ensure(poolCap));
        CU_CHECK(wordMeta.ensure(poolCap));
        CU_CHECK(testMask.ensure(poolCap));
        CU_CHECK(wordStatic.ensure(poolCap));
        dev.queue = queue.p;
        dev.queueCap = (uint32_t)queueCap;
        dev.ctr = ctr.p;
        dev.aliveBits = ctr.p + 128;
        dev.aliveWords = aliveWords;
        dev.groupCount = groupCount.p;
        dev.groupOff = groupOff.p;
        dev.groupsPerView = groups;
        dev.groupList = groupList.p;
        dev.wordBase = wordBase.p;
        dev.viewWords = viewWords.p;
        dev.visWord = visWord.p;
        dev.wordSlot = wordSlot.p;
        dev.wordMeta = wordMeta.p;
        dev.testMask = testMask.p;
        dev.wordStatic = wordStatic.p;
        dev.poolCap = (uint32_t)poolCap;
        dev.stateStride = groups * 32;
        return 0;
    }
};

struct u3d_batch
{
    u3d_ctx* ctx{};
    u3d_scene* scene{};
    u3d_occluders* occ{};
    ViewSet vs;
    uint32_t numViews{};
    uint32_t outCap{};
    uint64_t poolTris{};
    DevBuf<unsigned char> octState;
    DevBuf<unsigned char> itemPass;   // two-phase occluder drawing: per (view, item) kPassNear / kPassFarVisible / kPassFarHidden
    DevBuf<uint32_t> nearList, farVisList, listCounts; // ... and the near / far-visible items' positions, compacted; listCounts = [2][maxViews]
    bool countPending{};              // numTriangles_ of the occluders the second pass skipped is still to be counted (u3d_batch_read_tri_stats)
    DenseScratch dense;
    DevBuf<uint32_t> outIdx, outCounts, packOffsets, packed;
    DevBuf<float> zRange; // per view (minZ, maxZ)
    PinnedBuf<ViewConst> hostViews;
    PinnedBuf<uint32_t> hostCounts;
    cudaEvent_t viewsCopied{};        // the last host -> device copy out of hostViews (whatever stream it ran on)
    bool haveViewsEvent{};
    cudaEvent_t done{};               // blocking wait of u3d_batch_cull_views (option "blocking_wait")
    bool haveDoneEvent{};
    // u3d_batch_cull_views: flags | counts | packed offsets leave the device as ONE block, and the packed ids are copied in the same
    // breath up to the size the previous call needed (+ 1/8), so that the call has a single synchronisation in the steady state
    DevBuf<uint32_t> hdr;
    PinnedBuf<uint32_t> hostHdr;
    uint64_t idsGuess{};
    uint32_t lastLaunches{};
    bool hasBuffers{};
    cudaEvent_t ev[U3D_NUM_PHASES + 1]{};
    // u3d_batch_set_occluder_policy: View::UpdateOccluders heuristics + triangle budget instead of "every occluder in the frustum"
    bool rankOccluders{};
    bool serialOccluders{};           // non-threaded branch of View::DrawOccluders (policy mode 2)
    DevBuf<uint32_t> rankList, rankCount, activeOccluders;
    DevBuf<TriSetup> serialScratch;
    float sizeThreshold{};
    uint32_t maxTriangles{};
    DevBuf<float4> camParams;
    std::vector<float4> camHost;
    bool pipelined{};  // u3d_batch_set_pipelined
    bool timed{};      // record an event after every phase of the next run
    bool haveEvents{};
    ~u3d_batch()
    {
        if (haveEvents)
            for (auto& e : ev)
                cudaEventDestroy(e);
        if (haveViewsEvent)
            cudaEventDestroy(viewsCopied);
        if (haveDoneEvent)
            cudaEventDestroy(done);
    }
};

// -----------------------------------------------------------------------------------------------------------------------------
extern "C"
{

const char* u3d_last_error(void) { return g_err.c_str(); }
int u3d_abi_version(void) { return U3D_ABI_VERSION; }

int u3d_debug_set_option(const char* name, int value)
{
    if (name && std::strcmp(name, "force_irregular") == 0)
    {
        g_forceIrregular = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "lane_area_max") == 0)
    {
        g_laneAreaMax = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "warp_area_max") == 0)
    {
        g_warpAreaMax = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "heavy_extent") == 0)
    {
        cull_set_heavy_extent(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "two_phase") == 0)
    {
        g_twoPhase = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "two_phase_min") == 0)
    {
        g_twoPhaseMin = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "blocking_wait") == 0)
    {
        g_blockingWait = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "tile_target") == 0)
    {
        g_tileTarget = value; // takes effect for buffers / batches created afterwards
        return U3D_OK;
    }
    if (name && std::strcmp(name, "two_phase_shift") == 0)
    {
        g_twoPhaseShift = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "cull_dense") == 0)
    {
        g_cullDense = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_heavy") == 0)
    {
        dense_set_heavy_extent(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "walk_budget_busy") == 0)
    {
        dense_set_walk_budget(value, -1);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "walk_budget_idle") == 0)
    {
        dense_set_walk_budget(-1, value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_tree_blocks") == 0)
    {
        dense_set_tree_blocks(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "frustum_bulk") == 0)
    {
        dense_set_frustum_bulk(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_tree") == 0)
    {
        dense_set_tree(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_walker") == 0)
    {
        dense_set_walker(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_pool") == 0)
    {
        dense_set_walk_pool(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_queue_cap") == 0)
    {
        g_denseQueueCap = (uint32_t)value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_pool_cap") == 0)
    {
        g_densePoolCap = (uint32_t)value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "cull_ring") == 0)
    {
        cull_set_ring(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "walk_pool") == 0)
    {
        cull_set_walk_pool(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "cull_prof") == 0)
    {
        cull_prof_enable(value);
        return U3D_OK;
    }
    if (name && std::strcmp(name, "read_dense_prof") == 0)
    {
        g_readDenseProf = value;
        return U3D_OK;
    }
    if (name && std::strcmp(name, "dense_prof") == 0)
    {
        dense_prof_enable(value);
        return U3D_OK;
    }
    return fail(U3D_ERR_INVALID, "u3d_debug_set_option: unknown option");
}

int u3d_debug_read_cull_prof(unsigned long long* out8)
{
    if (!out8)
        return fail(U3D_ERR_INVALID, "NULL argument");
    if (g_readDenseProf >= 2 && g_readDenseProf < 34)
    {
        unsigned long long all[128];
        dense_prof_read_levels(all); // "read_dense_prof" = 2 + level: that level's four numbers
        for (int i = 0; i < 8; ++i)
            out8[i] = i < 4 ? all[4 * (g_readDenseProf - 2) + i] : 0ull;
    }
    else if (g_readDenseProf)
        dense_prof_read(out8);
    else
        cull_prof_read(out8);
    return U3D_OK;
}

int u3d_ctx_create(int device, u3d_ctx** out)
{
    if (!out)
        return fail(U3D_ERR_INVALID, "u3d_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(U3D_ERR_NO_DEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= count)
        return fail(U3D_ERR_INVALID, "u3d_ctx_create: bad device index");
    cudaDeviceProp prop{};
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(U3D_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; kernels are built for sm_100a only");
    CU_CHECK(cudaSetDevice(device));
    u3d_ctx* c = new (std::nothrow) u3d_ctx();
    if (!c)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    c->device = device;
    e = init_tile_kernel();
    if (e == cudaSuccess)
        e = init_cull_kernels();
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess)
    {
        delete c;
        return fail(U3D_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    *out = c;
    return U3D_OK;
}

void u3d_ctx_destroy(u3d_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int u3d_ctx_device(const u3d_ctx* ctx) { return ctx ? ctx->device : -U3D_ERR_INVALID; }

int u3d_ctx_synchronize(u3d_ctx* ctx, void* stream)
{
    if (!ctx)
        return fail(U3D_ERR_INVALID, "ctx is NULL");
    CU_CHECK(cudaSetDevice(ctx->device));
    CU_CHECK(cudaStreamSynchronize(ctx->pick(stream)));
    return U3D_OK;
}

// ---- host octree ------------------------------------------------------------------------------------------------------------
int u3d_octree_create(const float min3[3], const float max3[3], uint32_t levels, u3d_octree** out)
{
    if (!min3 || !max3 || !out)
        return fail(U3D_ERR_INVALID, "u3d_octree_create: NULL argument");
    *out = new (std::nothrow) u3d_octree(min3, max3, levels);
    return *out ? U3D_OK : fail(U3D_ERR_NOMEM, "out of host memory");
}

void u3d_octree_destroy(u3d_octree* t) { delete t; }

int u3d_octree_add(u3d_octree* t, uint32_t n, const float* aabb6, const uint8_t* drawable_flags, const uint32_t* view_mask, const uint8_t* occludee,
    const uint8_t* cast_shadows, uint32_t* first_id)
{
    if (!t || (n && !aabb6))
        return fail(U3D_ERR_INVALID, "u3d_octree_add: NULL argument");
    uint32_t first = (uint32_t)t->tree.Drawables().size();
    for (uint32_t i = 0; i < n; ++i)
        t->tree.AddDrawable(aabb6 + 6 * (size_t)i, drawable_flags ? drawable_flags[i] : (uint8_t)1, view_mask ? view_mask[i] : 0xFFFFFFFFu,
            occludee ? occludee[i] != 0 : true, cast_shadows ? cast_shadows[i] != 0 : false);
    if (first_id)
        *first_id = first;
    return U3D_OK;
}

int u3d_octree_set_draw_distances(u3d_octree* t, uint32_t first_id, uint32_t n, const float* draw_distance)
{
    if (!t || (n && !draw_distance))
        return fail(U3D_ERR_INVALID, "u3d_octree_set_draw_distances: NULL argument");
    if (!t->tree.SetDrawDistances(first_id, n, draw_distance))
        return fail(U3D_ERR_INVALID, "u3d_octree_set_draw_distances: drawable id out of range");
    return U3D_OK;
}

int u3d_octree_set_occluder_flags(u3d_octree* t, uint32_t first_id, uint32_t n, const uint8_t* is_occluder)
{
    if (!t || (n && !is_occluder))
        return fail(U3D_ERR_INVALID, "u3d_octree_set_occluder_flags: NULL argument");
    if (!t->tree.SetOccluderFlags(first_id, n, is_occluder))
        return fail(U3D_ERR_INVALID, "u3d_octree_set_occluder_flags: drawable id out of range");
    return U3D_OK;
}

int u3d_octree_update(u3d_octree* t, uint32_t id, const float aabb6[6])
{
    if (!t || !aabb6)
        return fail(U3D_ERR_INVALID, "u3d_octree_update: NULL argument");
    const uint64_t before = t->tree.StructureVersion();
    if (!t->tree.UpdateDrawable(id, aabb6))
        return fail(U3D_ERR_INVALID, "u3d_octree_update: unknown drawable id");
    return t->tree.StructureVersion() == before ? 1 : 0;
}

int u3d_octree_remove(u3d_octree* t, uint32_t id)
{
    if (!t)
        return fail(U3D_ERR_INVALID, "u3d_octree_remove: NULL argument");
    return t->tree.RemoveDrawable(id) ? U3D_OK : fail(U3D_ERR_INVALID, "u3d_octree_remove: unknown drawable id");
}

int u3d_octree_counts(const u3d_octree* t, uint32_t* num_octants, uint32_t* num_drawables_in_tree)
{
    if (!t)
        return fail(U3D_ERR_INVALID, "u3d_octree_counts: NULL argument");
    FlatOctree f;
    t->tree.Flatten(f);
    if (num_octants)
        *num_octants = (uint32_t)f.parent.size();
    if (num_drawables_in_tree)
        *num_drawables_in_tree = (uint32_t)f.order.size();
    return U3D_OK;
}

int u3d_octree_flatten(const u3d_octree* t, int32_t* parent, int32_t* level, float* culling_box6, int32_t* first, int32_t* count, int32_t* order)
{
    if (!t)
        return fail(U3D_ERR_INVALID, "u3d_octree_flatten: NULL argument");
    FlatOctree f;
    t->tree.Flatten(f);
    const size_t m = f.parent.size();
    if (parent) std::memcpy(parent, f.parent.data(), m * sizeof(int));
    if (level) std::memcpy(level, f.level.data(), m * sizeof(int));
    if (culling_box6) std::memcpy(culling_box6, f.box6.data(), m * 6 * sizeof(float));
    if (first) std::memcpy(first, f.first.data(), m * sizeof(int));
    if (count) std::memcpy(count, f.count.data(), m * sizeof(int));
    if (order) std::memcpy(order, f.order.data(), f.order.size() * sizeof(int));
    return U3D_OK;
}

// ---- scene ------------------------------------------------------------------------------------------------------------------
int u3d_scene_create_flat(u3d_ctx* ctx, uint32_t M, const int32_t* parent, const int32_t* level, const float* box6, const int32_t* first,
    const int32_t* count, uint32_t S, const int32_t* order, const float* aabb6, const uint8_t* drawable_flags, const uint32_t* view_mask,
    const uint8_t* occludee, const uint8_t* cast_shadows, u3d_scene** out)
{
    if (!ctx || !out || !M || !parent || !level || !box6 || !first || !count || (S && (!order || !aabb6)))
        return fail(U3D_ERR_INVALID, "u3d_scene_create_flat: NULL/empty argument");
    *out = nullptr;
    CU_CHECK(cudaSetDevice(ctx->device));
    int maxLevel = 0;
    for (uint32_t i = 0; i < M; ++i)
    {
        if (i && (parent[i] < 0 || parent[i] >= (int)i))
            return fail(U3D_ERR_INVALID, "u3d_scene_create_flat: octants are not in DFS pre-order (parent must precede child)");
        if (first[i] < 0 || count[i] < 0 || (uint64_t)first[i] + count[i] > S)
            return fail(U3D_ERR_INVALID, "u3d_scene_create_flat: octant drawable range out of bounds");
        maxLevel = std::max(maxLevel, level[i]);
    }
    std::vector<float4> omin(M), omax(M), dmin(std::max<uint32_t>(S, 1)), dmax(std::max<uint32_t>(S, 1));
    std::vector<uint32_t> ids(std::max<uint32_t>(S, 1));
    for (uint32_t i = 0; i < M; ++i)
    {
        const float* b = box6 + 6 * (size_t)i;
        const int p = i ? parent[i] : -1;
        omin[i] = make_float4(b[0], b[1], b[2], 0.f);
        omax[i] = make_float4(b[3], b[4], b[5], 0.f);
        std::memcpy(&omin[i].w, &p, 4);
        std::memcpy(&omax[i].w, &first[i], 4);
    }
    for (uint32_t s = 0; s < S; ++s)
    {
        const uint32_t id = (uint32_t)order[s];
        const float* b = aabb6 + 6 * (size_t)id;
        uint32_t meta = drawable_flags ? drawable_flags[id] : 1u;
        if (!occludee || occludee[id])
            meta |= kMetaOccludee;
        if (cast_shadows && cast_shadows[id])
            meta |= kMetaCastShadows;
        const uint32_t vm = view_mask ? view_mask[id] : 0xFFFFFFFFu;
        dmin[s] = make_float4(b[0], b[1], b[2], 0.f);
        dmax[s] = make_float4(b[3], b[4], b[5], 0.f);
        std::memcpy(&dmin[s].w, &meta, 4);
        std::memcpy(&dmax[s].w, &vm, 4);
        ids[s] = id;
    }
    // octants sorted by level (stable): a child's parent is always resolved one sweep earlier
    std::vector<int> levelStart(maxLevel + 2, 0), bfs(M);
    for (uint32_t i = 0; i < M; ++i)
        ++levelStart[level[i] + 1];
    for (int l = 0; l <= maxLevel; ++l)
        levelStart[l + 1] += levelStart[l];
    {
        std::vector<int> cur(levelStart.begin(), levelStart.end() - 1);
        for (uint32_t i = 0; i < M; ++i)
            bfs[cur[level[i]]++] = (int)i;
    }
    for (uint32_t i = 1; i < M; ++i)
        if (level[parent[i]] + 1 != level[i])
            return fail(U3D_ERR_INVALID, "u3d_scene_create_flat: a child's level must be its parent's level + 1");
    if (M > (256u * 32u << 9))
        return fail(U3D_ERR_UNSUPPORTED, "u3d_scene_create_flat: more than 4M octants");
    // children of an octant are contiguous in the level-sorted list (stable sort of a DFS order): record where
    std::vector<int2> child(M, make_int2(0, 0));
    for (uint32_t p = 0; p < M; ++p)
    {
        const int o = bfs[p];
        if (o == 0)
            continue;
        int2& c = child[parent[o]];
        if (c.y == 0)
            c.x = (int)p;
        else if (c.x + c.y != (int)p)
            return fail(U3D_ERR_INVALID, "u3d_scene_create_flat: octants are not in DFS pre-order (children not contiguous per level)");
        ++c.y;
    }

    u3d_scene* s = new (std::nothrow) u3d_scene();
    if (!s)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    s->ctx = ctx;
    auto up = [&](auto& buf, const auto& vec) -> cudaError_t {
        cudaError_t e = buf.ensure(vec.size());
        if (e != cudaSuccess)
            return e;
        return cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice);
    };
    std::vector<int> cnt(count, count + M);
    // static filter summaries (SceneDev::wordInfo)
    std::vector<uint32_t> wordStart(M + 1, 0);
    for (uint32_t i = 0; i < M; ++i)
        wordStart[i + 1] = wordStart[i] + ((uint32_t)count[i] + 31u) / 32u;
    std::vector<uint4> wordInfo(std::max<uint32_t>(wordStart[M], 1), make_uint4(0xFFFFFFFFu, 0, 0, 0));
    for (uint32_t i = 0; i < M; ++i)
        for (uint32_t k = 0; k < (uint32_t)count[i]; k += 32)
        {
            const uint32_t s0 = (uint32_t)first[i] + k, nv = std::min<uint32_t>(32u, (uint32_t)count[i] - k);
            uint32_t meta0, vm0;
            std::memcpy(&meta0, &dmin[s0].w, 4);
            std::memcpy(&vm0, &dmax[s0].w, 4);
            bool uniform = true;
            uint32_t occ = 0, cast = 0;
            for (uint32_t j = 0; j < nv; ++j)
            {
                uint32_t meta, vm;
                std::memcpy(&meta, &dmin[s0 + j].w, 4);
                std::memcpy(&vm, &dmax[s0 + j].w, 4);
                uniform = uniform && (meta & 0xFFu) == (meta0 & 0xFFu) && vm == vm0;
                if (meta & kMetaOccludee)
                    occ |= 1u << j;
                if (meta & kMetaCastShadows)
                    cast |= 1u << j;
            }
            wordInfo[wordStart[i] + k / 32] = make_uint4(uniform ? (meta0 & 0xFFu) : 0xFFFFFFFFu, vm0, occ, cast);
        }
    cudaError_t e = cudaSuccess;
    if ((e = up(s->octMin, omin)) || (e = up(s->octMax, omax)) || (e = up(s->octCount, cnt)) || (e = up(s->bfs, bfs)) || (e = up(s->octChild, child)) ||
        (e = up(s->levelStart, levelStart)) || (e = up(s->drwMin, dmin)) || (e = up(s->drwMax, dmax)) || (e = up(s->drwId, ids)) ||
        (e = up(s->wordInfo, wordInfo)) || (e = up(s->octWordStart, wordStart)))
    {
        delete s;
        return fail(U3D_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e));
    }
    s->dev.numOctants = (int)M;
    s->dev.numDrawables = (int)S;
    s->dev.numLevels = maxLevel + 1;
    s->dev.octMin = s->octMin.p;
    s->dev.octMax = s->octMax.p;
    s->dev.octCount = s->octCount.p;
    s->dev.bfs = s->bfs.p;
    s->dev.octChild = s->octChild.p;
    s->dev.levelStart = s->levelStart.p;
    s->dev.drwMin = s->drwMin.p;
    s->dev.drwMax = s->drwMax.p;
    s->dev.drwId = s->drwId.p;
    s->dev.wordInfo = s->wordInfo.p;
    s->dev.octWordStart = s->octWordStart.p;
    {
        cudaError_t e2 = s->drwDrawDist.ensure(std::max<uint32_t>(S, 1));
        if (e2 == cudaSuccess)
            e2 = cudaMemset(s->drwDrawDist.p, 0, (size_t)std::max<uint32_t>(S, 1) * 4);
        if (e2 != cudaSuccess)
        {
            delete s;
            return fail(U3D_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e2));
        }
        s->dev.drwDrawDist = s->drwDrawDist.p;
        uint32_t maxId = 0;
        for (uint32_t i = 0; i < S; ++i)
            maxId = std::max(maxId, ids[i]);
        s->slotOfId.assign(S ? maxId + 1 : 0, 0xFFFFFFFFu);
        for (uint32_t i = 0; i < S; ++i)
            s->slotOfId[ids[i]] = i;
    }
    *out = s;
    return U3D_OK;
}

int u3d_scene_create(u3d_ctx* ctx, const u3d_octree* t, u3d_scene** out)
{
    if (!ctx || !t || !out)
        return fail(U3D_ERR_INVALID, "u3d_scene_create: NULL argument");
    FlatOctree f;
    t->tree.Flatten(f);
    const auto& dr = t->tree.Drawables();
    const size_t n = dr.size();
    std::vector<float> aabb(6 * std::max<size_t>(n, 1));
    std::vector<uint8_t> fl(std::max<size_t>(n, 1)), oc(std::max<size_t>(n, 1)), cs(std::max<size_t>(n, 1));
    std::vector<uint32_t> vm(std::max<size_t>(n, 1));
    for (size_t i = 0; i < n; ++i)
    {
        std::memcpy(&aabb[6 * i], dr[i].box, sizeof(dr[i].box));
        fl[i] = dr[i].flags;
        vm[i] = dr[i].viewMask;
        oc[i] = dr[i].occludee;
        cs[i] = dr[i].castShadows;
    }
    int rc = u3d_scene_create_flat(ctx, (uint32_t)f.parent.size(), f.parent.data(), f.level.data(), f.box6.data(), f.first.data(), f.count.data(),
        (uint32_t)f.order.size(), f.order.data(), aabb.data(), fl.data(), vm.data(), oc.data(), cs.data(), out);
    if (rc < 0)
        return rc;
    std::vector<float> dd(std::max<size_t>(n, 1), 0.0f);
    bool any = false;
    for (size_t i = 0; i < n; ++i)
    {
        dd[i] = dr[i].drawDistance;
        any |= dd[i] != 0.0f;
    }
    if (any)
    {
        rc = u3d_scene_set_draw_distances(*out, (uint32_t)n, dd.data());
        if (rc < 0)
        {
            u3d_scene_destroy(*out); // never hand out a half-initialised scene
            *out = nullptr;
            return rc;
        }
    }
    std::vector<uint8_t> occl(std::max<size_t>(n, 1), 0);
    any = false;
    for (size_t i = 0; i < n; ++i)
    {
        occl[i] = dr[i].occluder;
        any |= dr[i].occluder;
    }
    rc = any ? u3d_scene_set_occluder_flags(*out, (uint32_t)n, occl.data()) : U3D_OK;
    if (rc < 0)
    {
        u3d_scene_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int u3d_scene_set_occluder_flags(u3d_scene* s, uint32_t n, const uint8_t* is_occluder)
{
    if (!s || (n && !is_occluder))
        return fail(U3D_ERR_INVALID, "u3d_scene_set_occluder_flags: NULL argument");
    CU_CHECK(cudaSetDevice(s->ctx->device));
    const uint32_t S = (uint32_t)s->dev.numDrawables;
    std::vector<unsigned char> bySlot(std::max<uint32_t>(S, 1), 0);
    for (uint32_t id = 0; id < n && id < s->slotOfId.size(); ++id)
        if (s->slotOfId[id] != 0xFFFFFFFFu)
            bySlot[s->slotOfId[id]] = is_occluder[id] ? 1 : 0;
    DevBuf<unsigned char> staging;
    CU_CHECK(staging.ensure(std::max<uint32_t>(S, 1)));
    CU_CHECK(cudaMemcpy(staging.p, bySlot.data(), std::max<uint32_t>(S, 1), cudaMemcpyHostToDevice));
    launch_set_occluder_bits(s->drwMin.p, staging.p, S, s->ctx->stream);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaStreamSynchronize(s->ctx->stream));
    return U3D_OK;
}

int u3d_scene_set_draw_distances(u3d_scene* s, uint32_t n, const float* draw_distance)
{
    if (!s || (n && !draw_distance))
        return fail(U3D_ERR_INVALID, "u3d_scene_set_draw_distances: NULL argument");
    CU_CHECK(cudaSetDevice(s->ctx->device));
    const uint32_t S = (uint32_t)s->dev.numDrawables;
    std::vector<float> bySlot(std::max<uint32_t>(S, 1), 0.0f);
    for (uint32_t id = 0; id < n && id < s->slotOfId.size(); ++id)
        if (s->slotOfId[id] != 0xFFFFFFFFu)
            bySlot[s->slotOfId[id]] = draw_distance[id];
    CU_CHECK(cudaMemcpy(s->drwDrawDist.p, bySlot.data(), (size_t)std::max<uint32_t>(S, 1) * 4, cudaMemcpyHostToDevice));
    return U3D_OK;
}

int u3d_scene_update_boxes(u3d_scene* s, uint32_t n, const uint32_t* ids, const float* aabb6, void* stream)
{
    if (!s || (n && (!ids || !aabb6)))
        return fail(U3D_ERR_INVALID, "u3d_scene_update_boxes: NULL argument");
    if (!n)
        return U3D_OK;
    CU_CHECK(cudaSetDevice(s->ctx->device));
    cudaStream_t st = s->ctx->pick(stream);
    // A drawable may appear several times (it moved more than once since the last upload): the LAST box wins, as in the host tree.  The
    // kernel writes one slot per thread without ordering, so duplicates are folded here.
    std::vector<uint32_t> slots;
    std::vector<float> boxes;
    slots.reserve(n);
    boxes.reserve((size_t)n * 6);
    std::map<uint32_t, uint32_t> where; // slot -> position in `slots`
    for (uint32_t i = 0; i < n; ++i)
    {
        if (ids[i] >= s->slotOfId.size() || s->slotOfId[ids[i]] == 0xFFFFFFFFu)
            return fail(U3D_ERR_INVALID, "u3d_scene_update_boxes: drawable id is not in the scene");
        const uint32_t slot = s->slotOfId[ids[i]];
        auto it = where.find(slot);
        if (it == where.end())
        {
            where.emplace(slot, (uint32_t)slots.size());
            slots.push_back(slot);
            boxes.insert(boxes.end(), aabb6 + 6 * (size_t)i, aabb6 + 6 * (size_t)i + 6);
        }
        else
            std::memcpy(&boxes[6 * (size_t)it->second], aabb6 + 6 * (size_t)i, 24);
    }
    n = (uint32_t)slots.size();
    CU_CHECK(s->updSlots.ensure(n));
    CU_CHECK(s->updBoxes.ensure((size_t)n * 6));
    CU_CHECK(cudaMemcpyAsync(s->updSlots.p, slots.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(s->updBoxes.p, boxes.data(), (size_t)n * 24, cudaMemcpyHostToDevice, st));
    launch_update_boxes(s->drwMin.p, s->drwMax.p, s->updSlots.p, s->updBoxes.p, n, st);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaStreamSynchronize(st)); // host staging vectors go out of scope
    return U3D_OK;
}

void u3d_scene_destroy(u3d_scene* s)
{
    if (!s)
        return;
    cudaSetDevice(s->ctx->device);
    delete s;
}

int u3d_scene_counts(const u3d_scene* s, uint32_t* num_octants, uint32_t* num_drawables)
{
    if (!s)
        return fail(U3D_ERR_INVALID, "u3d_scene_counts: NULL argument");
    if (num_octants)
        *num_octants = (uint32_t)s->dev.numOctants;
    if (num_drawables)
        *num_drawables = (uint32_t)s->dev.numDrawables;
    return U3D_OK;
}

// ---- meshes / occluders -----------------------------------------------------------------------------------------------------
int u3d_mesh_create(u3d_ctx* ctx, const void* vertices, uint32_t vertex_size, uint32_t vertex_count, const void* indices, uint32_t index_size,
    uint32_t index_count, u3d_mesh** out)
{
    if (!ctx || !out || !vertices || vertex_size < 12 || !vertex_count)
        return fail(U3D_ERR_INVALID, "u3d_mesh_create: bad vertex arguments (vertex_size must be >= 12)");
    if (indices && index_size != 2 && index_size != 4)
        return fail(U3D_ERR_INVALID, "u3d_mesh_create: index_size must be 2 or 4");
    if ((vertex_size & 3) != 0)
        return fail(U3D_ERR_UNSUPPORTED, "u3d_mesh_create: vertex_size must be a multiple of 4 (float-aligned positions)");
    *out = nullptr;
    CU_CHECK(cudaSetDevice(ctx->device));
    u3d_mesh* m = new (std::nothrow) u3d_mesh();
    if (!m)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    m->ctx = ctx;
    m->stride = vertex_size;
    m->vertexCount = vertex_count;
    // the reference reads 12 bytes per vertex; the last vertex need not be a full stride long
    const size_t vbytes = (size_t)(vertex_count - 1) * vertex_size + 12;
    cudaError_t e = m->vtx.ensure((size_t)vertex_count * vertex_size);
    if (e == cudaSuccess)
        e = cudaMemcpy(m->vtx.p, vertices, vbytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && indices)
    {
        m->idxSize = index_size;
        m->idxCount = index_count;
        e = m->idx.ensure((size_t)index_count * index_size);
        if (e == cudaSuccess)
            e = cudaMemcpy(m->idx.p, indices, (size_t)index_count * index_size, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess)
    {
        delete m;
        return fail(U3D_ERR_CUDA, std::string("mesh upload: ") + cudaGetErrorString(e));
    }
    *out = m;
    return U3D_OK;
}

void u3d_mesh_destroy(u3d_mesh* m)
{
    if (!m)
        return;
    cudaSetDevice(m->ctx->device);
    delete m;
}

int u3d_occluders_create(u3d_ctx* ctx, uint32_t n, const float* world_aabb6, const float* model12, u3d_mesh* const* meshes, uint32_t num_meshes,
    const uint32_t* mesh_of_occluder, const uint32_t* index_start, const uint32_t* index_count, int cull_mode, u3d_occluders** out)
{
    if (!ctx || !out || (n && (!world_aabb6 || !model12 || !meshes || !num_meshes)))
        return fail(U3D_ERR_INVALID, "u3d_occluders_create: NULL argument");
    if (cull_mode < 0 || cull_mode > 2)
        return fail(U3D_ERR_INVALID, "u3d_occluders_create: bad cull mode");
    *out = nullptr;
    CU_CHECK(cudaSetDevice(ctx->device));
    std::vector<MeshDev> table(std::max<uint32_t>(num_meshes, 1));
    for (uint32_t i = 0; i < num_meshes; ++i)
    {
        if (!meshes[i])
            return fail(U3D_ERR_INVALID, "u3d_occluders_create: NULL mesh");
        table[i] = meshes[i]->dev();
    }
    std::vector<uint32_t> mesh(std::max<uint32_t>(n, 1)), start(std::max<uint32_t>(n, 1)), count(std::max<uint32_t>(n, 1));
    uint32_t totalItems = 0;
    uint64_t totalTris = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t mi = mesh_of_occluder ? mesh_of_occluder[i] : 0;
        if (mi >= num_meshes)
            return fail(U3D_ERR_INVALID, "u3d_occluders_create: mesh index out of range");
        const u3d_mesh* m = meshes[mi];
        const uint32_t avail = m->idxSize ? m->idxCount : m->vertexCount;
        const uint32_t st = index_start ? index_start[i] : 0;
        const uint32_t ct = index_count ? index_count[i] : avail;
        if ((uint64_t)st + ct > avail)
            return fail(U3D_ERR_INVALID, "u3d_occluders_create: draw range exceeds the mesh");
        if (m->idxSize && ct % 3)
            // DrawBatch (OB.cpp:507-519, 525-537) would draw ceil(count / 3) triangles and read past the range; the batched kernels draw
            // count / 3, so such a range is refused rather than drawn differently (the drop-in u3d_ob_* path follows the reference)
            return fail(U3D_ERR_UNSUPPORTED, "u3d_occluders_create: an indexed draw range must hold whole triangles (count % 3 == 0)");
        mesh[i] = mi;
        start[i] = st;
        count[i] = ct;
        const uint32_t tris = ct / 3;
        totalTris += tris;
        totalItems += (tris + kItemTris - 1) / kItemTris;
    }
    u3d_occluders* o = new (std::nothrow) u3d_occluders();
    if (!o)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    o->ctx = ctx;
    o->n = n;
    o->cullMode = cull_mode;
    o->totalItems = totalItems;
    o->totalTris = totalTris;
    cudaError_t e = cudaSuccess;
    auto up = [&](auto& buf, const void* src, size_t count_, size_t elem) -> cudaError_t {
        cudaError_t e2 = buf.ensure(std::max<size_t>(count_, 1));
        if (e2 != cudaSuccess || !count_)
            return e2;
        return cudaMemcpy(buf.p, src, count_ * elem, cudaMemcpyHostToDevice);
    };
    if ((e = up(o->aabb6, world_aabb6, (size_t)n * 6, 4)) || (e = up(o->model12, model12, (size_t)n * 12, 4)) || (e = up(o->mesh, mesh.data(), n, 4)) ||
        (e = up(o->start, start.data(), n, 4)) || (e = up(o->count, count.data(), n, 4)) ||
        (e = up(o->meshTable, table.data(), table.size(), sizeof(MeshDev))))
    {
        delete o;
        return fail(U3D_ERR_CUDA, std::string("occluder upload: ") + cudaGetErrorString(e));
    }
    *out = o;
    return U3D_OK;
}

void u3d_occluders_destroy(u3d_occluders* o)
{
    if (!o)
        return;
    cudaSetDevice(o->ctx->device);
    delete o;
}

// ---- OcclusionBuffer drop-in ------------------------------------------------------------------------------------------------
int u3d_ob_create(u3d_ctx* ctx, u3d_ob** out)
{
    if (!ctx || !out)
        return fail(U3D_ERR_INVALID, "u3d_ob_create: NULL argument");
    u3d_ob* ob = new (std::nothrow) u3d_ob();
    if (!ob)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    ob->ctx = ctx;
    *out = ob;
    return U3D_OK;
}

void u3d_ob_destroy(u3d_ob* ob)
{
    if (!ob)
        return;
    cudaSetDevice(ob->ctx->device);
    delete ob;
}

static int ob_upload_view(u3d_ob* ob)
{
    if (!ob->allocated)
        return U3D_OK;
    ViewConst vc{};
    std::memcpy(vc.vp, ob->viewProj, sizeof(vc.vp));
    vc.drawableFlags = 0xFF;
    vc.viewMask = 0xFFFFFFFFu;
    vc.cullMode = ob->cullMode;
    vc.queryMode = QUERY_OCCLUDED;
    vc.useOcclusion = 1;
    CU_CHECK(cudaMemcpyAsync(ob->vs.views.p, &vc, sizeof(vc), cudaMemcpyHostToDevice, ob->ctx->stream));
    CU_CHECK(cudaStreamSynchronize(ob->ctx->stream)); // vc lives on this stack frame
    return U3D_OK;
}

int u3d_ob_set_size(u3d_ob* ob, int width, int height, int threaded)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "u3d_ob_set_size: NULL argument");
    if (height & 1) // OB.cpp:63-65
        ++height;
    if (width == ob->width && height == ob->height && ob->allocated)
        return 1;
    if (width <= 0 || height <= 0)
        return 0;
    if (width & (width - 1))
    {
        g_err = "Requested occlusion buffer width is not a power of two"; // OB.cpp:73-77
        return 0;
    }
    if (width < 2 || width > 8192 || height > 16384)
        return fail(U3D_ERR_UNSUPPORTED, "u3d_ob_set_size: supported sizes are 2..8192 x 2..16384");
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    ob->width = width;
    ob->height = height;
    ob->threaded = threaded != 0;
    int rc = ob->vs.alloc(ob->ctx, width, height, 1, 1, 1);
    if (rc < 0)
        return rc;
    ob->allocated = true;
    ob->hierarchyDirty = true;
    return 1;
}

int u3d_ob_set_view(u3d_ob* ob, const float view12[12], const float projection16[16], float near_clip, float far_clip, int reverse_culling)
{
    if (!ob || !view12 || !projection16)
        return fail(U3D_ERR_INVALID, "u3d_ob_set_view: NULL argument");
    std::memcpy(ob->view, view12, sizeof(ob->view));
    std::memcpy(ob->proj, projection16, sizeof(ob->proj));
    mul_proj_view(ob->proj, ob->view, ob->viewProj);
    ob->nearClip = near_clip;
    ob->farClip = far_clip;
    ob->reverseCulling = reverse_culling != 0;
    return U3D_OK;
}

int u3d_ob_set_max_triangles(u3d_ob* ob, uint32_t triangles)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "NULL argument");
    ob->maxTriangles = triangles;
    return U3D_OK;
}

int u3d_ob_set_cull_mode(u3d_ob* ob, int mode)
{
    if (!ob || mode < 0 || mode > 2)
        return fail(U3D_ERR_INVALID, "u3d_ob_set_cull_mode: bad argument");
    ob->cullMode = flip_cull(mode, ob->reverseCulling);
    return U3D_OK;
}

int u3d_ob_reset(u3d_ob* ob)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "NULL argument");
    ob->numTriangles = 0;
    ob->batches.clear();
    return U3D_OK;
}

int u3d_ob_clear(u3d_ob* ob)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "NULL argument");
    ob->numTriangles = 0;
    ob->batches.clear();
    if (ob->allocated)
    {
        CU_CHECK(cudaSetDevice(ob->ctx->device));
        launch_clear(ob->vs.depth.p, (size_t)ob->width * ob->height, ob->ctx->stream);
        CU_CHECK(cudaGetLastError());
    }
    ob->hierarchyDirty = true;
    return U3D_OK;
}

static int ob_push(u3d_ob* ob, const float* model, const void* vtx, uint32_t vsize, const void* idx, uint32_t isize, uint32_t start, uint32_t count,
    const u3d_mesh* mesh)
{
    HostBatchRec r{};
    std::memcpy(r.model, model, sizeof(r.model));
    r.vtx = vtx;
    r.vtxSize = vsize;
    r.idx = idx;
    r.idxSize = isize;
    r.start = start;
    r.count = count;
    r.mesh = mesh;
    ob->batches.push_back(r);
    ob->numTriangles += count / 3;
    return ob->numTriangles <= ob->maxTriangles ? 1 : 0;
}

int u3d_ob_add_triangles(u3d_ob* ob, const float model12[12], const void* vertex_data, uint32_t vertex_size, uint32_t vertex_start, uint32_t vertex_count)
{
    if (!ob || !model12 || !vertex_data)
        return fail(U3D_ERR_INVALID, "u3d_ob_add_triangles: NULL argument");
    if (vertex_size < 12 || (vertex_size & 3))
        return fail(U3D_ERR_UNSUPPORTED, "u3d_ob_add_triangles: vertex_size must be >= 12 and a multiple of 4");
    return ob_push(ob, model12, vertex_data, vertex_size, nullptr, 0, vertex_start, vertex_count, nullptr);
}

int u3d_ob_add_triangles_indexed(u3d_ob* ob, const float model12[12], const void* vertex_data, uint32_t vertex_size, const void* index_data,
    uint32_t index_size, uint32_t index_start, uint32_t index_count)
{
    if (!ob || !model12 || !vertex_data || !index_data)
        return fail(U3D_ERR_INVALID, "u3d_ob_add_triangles_indexed: NULL argument");
    if (vertex_size < 12 || (vertex_size & 3))
        return fail(U3D_ERR_UNSUPPORTED, "u3d_ob_add_triangles_indexed: vertex_size must be >= 12 and a multiple of 4");
    if (index_size != 2 && index_size != 4)
        return fail(U3D_ERR_INVALID, "u3d_ob_add_triangles_indexed: index_size must be 2 or 4");
    return ob_push(ob, model12, vertex_data, vertex_size, index_data, index_size, index_start, index_count, nullptr);
}

int u3d_ob_add_triangles_mesh(u3d_ob* ob, const float model12[12], const u3d_mesh* mesh, uint32_t start, uint32_t count)
{
    if (!ob || !model12 || !mesh)
        return fail(U3D_ERR_INVALID, "u3d_ob_add_triangles_mesh: NULL argument");
    const uint32_t avail = mesh->idxSize ? mesh->idxCount : mesh->vertexCount;
    // indexed draws touch whole triangles (see u3d_ob_draw_triangles): the last one may reach up to two indices past `count`
    const uint64_t touched = mesh->idxSize ? ((uint64_t)count + 2) / 3 * 3 : count;
    if ((uint64_t)start + touched > avail)
        return fail(U3D_ERR_INVALID, "u3d_ob_add_triangles_mesh: range exceeds the mesh (indexed ranges are read in whole triangles, as the reference's DrawBatch does)");
    return ob_push(ob, model12, nullptr, mesh->stride, nullptr, mesh->idxSize, start, count, mesh);
}

int u3d_ob_draw_triangles(u3d_ob* ob)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "NULL argument");
    if (!ob->allocated || ob->batches.empty())
    {
        ob->batches.clear(); // OB.cpp:231
        return U3D_OK;
    }
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    cudaStream_t st = ob->ctx->stream;
    const size_t nb = ob->batches.size();

    // Stage the caller's geometry: one contiguous blob = [vertex ranges][index ranges], deduplicated by host pointer.
    struct VRange { size_t bytes; size_t offset; };
    std::map<const void*, VRange> vranges;
    std::vector<size_t> idxOffset(nb, 0);
    size_t total = 0;
    std::vector<uint32_t> tris(nb);
    for (size_t i = 0; i < nb; ++i)
    {
        const HostBatchRec& r = ob->batches[i];
        // DrawBatch, OB.cpp:464-541: non-indexed geometry draws floor(count / 3) triangles (`while (index + 2 < drawCount)`), indexed geometry
        // loops `while (indices < indicesEnd)` in steps of three, i.e. ceil(count / 3) triangles, reading up to two indices past the stated
        // range when the count is not a multiple of three.  Followed as is (numTriangles_ still counts count / 3, OB.cpp:196).
        const bool indexed = r.mesh ? r.mesh->idxSize != 0 : r.idx != nullptr;
        tris[i] = indexed ? (r.count + 2) / 3 : r.count / 3;
        if (r.mesh || !tris[i])
            continue;
        uint32_t maxVertex = 0;
        if (!r.idx)
            maxVertex = r.start + 3 * tris[i] - 1;
        else if (r.idxSize == 2)
        {
            const uint16_t* ix = (const uint16_t*)r.idx + r.start;
            for (uint32_t k = 0; k < 3 * tris[i]; ++k)
                maxVertex = std::max<uint32_t>(maxVertex, ix[k]);
        }
        else
        {
            const uint32_t* ix = (const uint32_t*)r.idx + r.start;
            for (uint32_t k = 0; k < 3 * tris[i]; ++k)
                maxVertex = std::max<uint32_t>(maxVertex, ix[k]);
        }
        const size_t need = (size_t)maxVertex * r.vtxSize + 12;
        auto it = vranges.find(r.vtx);
        if (it == vranges.end())
            vranges[r.vtx] = VRange{need, 0};
        else
            it->second.bytes = std::max(it->second.bytes, need);
    }
    for (auto& kv : vranges)
    {
        kv.second.offset = total;
        total += (kv.second.bytes + 15) & ~(size_t)15;
    }
    for (size_t i = 0; i < nb; ++i)
    {
        const HostBatchRec& r = ob->batches[i];
        if (r.mesh || !r.idx || !tris[i])
            continue;
        idxOffset[i] = total;
        total += ((size_t)3 * tris[i] * r.idxSize + 15) & ~(size_t)15;
    }
    std::vector<unsigned char> blob(std::max<size_t>(total, 16));
    for (auto& kv : vranges)
        std::memcpy(blob.data() + kv.second.offset, kv.first, kv.second.bytes);
    for (size_t i = 0; i < nb; ++i)
    {
        const HostBatchRec& r = ob->batches[i];
        if (r.mesh || !r.idx || !tris[i])
            continue;
        std::memcpy(blob.data() + idxOffset[i], (const unsigned char*)r.idx + (size_t)r.start * r.idxSize, (size_t)3 * tris[i] * r.idxSize);
    }
    CU_CHECK(ob->geom.ensure(blob.size()));
    CU_CHECK(cudaMemcpyAsync(ob->geom.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));

    std::vector<float> models(nb * 12);
    std::vector<MeshDev> table(nb);
    std::vector<DrawItem> items;
    uint64_t totalTris = 0;
    for (size_t i = 0; i < nb; ++i)
    {
        const HostBatchRec& r = ob->batches[i];
        std::memcpy(&models[12 * i], r.model, sizeof(r.model));
        uint32_t start = r.start;
        if (r.mesh)
            table[i] = r.mesh->dev();
        else if (tris[i])
        {
            table[i].vtx = ob->geom.p + vranges[r.vtx].offset;
            table[i].stride = r.vtxSize;
            table[i].idxSize = r.idxSize;
            table[i].idx = r.idx ? (const void*)(ob->geom.p + idxOffset[i]) : nullptr;
            if (r.idx)
                start = 0; // staged index range begins at the batch's first index
        }
        for (uint32_t t0 = 0; t0 < tris[i]; t0 += kItemTris)
        {
            DrawItem it;
            it.model = (uint32_t)i;
            it.mesh = (uint32_t)i;
            it.start = start + 3 * t0;
            it.count = 3 * std::min<uint32_t>(tris[i] - t0, kItemTris);
            items.push_back(it);
        }
        totalTris += tris[i];
    }
    ob->batches.clear(); // OB.cpp:231
    if (items.empty())
    {
        CU_CHECK(cudaStreamSynchronize(st));
        return U3D_OK;
    }
    ViewSet& vs = ob->vs;
    CU_CHECK(ob->models.ensure(models.size()));
    CU_CHECK(ob->meshTable.ensure(table.size()));
    CU_CHECK(vs.items.ensure(items.size()));
    vs.itemCap = (uint32_t)items.size();
    // any submitted triangle may need the clipper
    vs.clipCap = (uint32_t)std::min<uint64_t>(totalTris + 1, 0x7FFFFFFFu);
    CU_CHECK(vs.clipQueue.ensure((size_t)vs.clipCap * kClipEntryFloats));
    // every source triangle can be clipped into several; grow on overflow
    uint64_t poolTris = totalTris + totalTris / 2 + 256;
    for (int attempt = 0; attempt < 8; ++attempt)
    {
        CU_CHECK(vs.pool.ensure(poolTris));
        CU_CHECK(cudaMemcpyAsync(ob->models.p, models.data(), models.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        CU_CHECK(cudaMemcpyAsync(ob->meshTable.p, table.data(), table.size() * sizeof(MeshDev), cudaMemcpyHostToDevice, st));
        CU_CHECK(cudaMemcpyAsync(vs.items.p, items.data(), items.size() * sizeof(DrawItem), cudaMemcpyHostToDevice, st));
        const uint32_t nItems = (uint32_t)items.size();
        const uint32_t cap = (uint32_t)std::min<uint64_t>(poolTris, 0xFFFFFFFFu);
        const uint64_t zero64 = 0;
        CU_CHECK(vs.clear_counters(st));
        CU_CHECK(cudaMemcpyAsync(vs.itemCount.p, &nItems, 4, cudaMemcpyHostToDevice, st));
        CU_CHECK(cudaMemcpyAsync(vs.triCap.p, &cap, 4, cudaMemcpyHostToDevice, st));
        CU_CHECK(cudaMemcpyAsync(vs.triBase.p, &zero64, 8, cudaMemcpyHostToDevice, st));
        int rc = ob_upload_view(ob); // cullMode_ is read at draw time (OB.cpp:634)
        if (rc < 0)
            return rc;
        launch_setup(vs.g, vs.tg, vs.views.p, 1, vs.items.p, vs.itemCap, nItems, vs.itemCount.p, ob->models.p, ob->meshTable.p, vs.raster_dev(false, 0), st);
        CU_CHECK(cudaGetLastError());
        uint32_t res[3] = {0, 0, 0};
        CU_CHECK(cudaMemcpyAsync(&res[0], vs.triCount.p, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(&res[1], vs.drawnSources.p, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(&res[2], vs.flags.p, 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        if (res[2] & kFlagTriOverflow)
        {
            poolTris = (uint64_t)res[0] + 256; // exact requirement is known now
            continue;
        }
        launch_fill_global(vs.g, 1, vs.depth.p, vs.pool.p, vs.triBase.p, vs.triCap.p, vs.triCount.p, 148 * 2, st);
        CU_CHECK(cudaGetLastError());
        CU_CHECK(cudaStreamSynchronize(st));
        ob->numTriangles += res[1]; // ++numTriangles_ per drawn source triangle, OB.cpp:681-682
        ob->hierarchyDirty = true;
        return U3D_OK;
    }
    return fail(U3D_ERR_CAPACITY, "u3d_ob_draw_triangles: triangle pool kept overflowing");
}

int u3d_ob_build_depth_hierarchy(u3d_ob* ob)
{
    if (!ob)
        return fail(U3D_ERR_INVALID, "NULL argument");
    if (!ob->allocated || !ob->hierarchyDirty) // OB.cpp:236
        return U3D_OK;
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    launch_hierarchy(ob->vs.g, 1, ob->vs.depth.p, ob->vs.mips.p, 0, ob->ctx->stream);
    CU_CHECK(cudaGetLastError());
    ob->hierarchyDirty = false;
    return U3D_OK;
}

int u3d_ob_is_visible_many(u3d_ob* ob, const float* aabb6, uint32_t n, uint8_t* out)
{
    if (!ob || (n && (!aabb6 || !out)))
        return fail(U3D_ERR_INVALID, "u3d_ob_is_visible_many: NULL argument");
    if (!ob->allocated) // OB.cpp:338-339
    {
        std::memset(out, 1, n);
        return U3D_OK;
    }
    if (!n)
        return U3D_OK;
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    cudaStream_t st = ob->ctx->stream;
    int rc = ob_upload_view(ob);
    if (rc < 0)
        return rc;
    CU_CHECK(ob->boxes.ensure((size_t)n * 6));
    CU_CHECK(ob->visOut.ensure(n));
    CU_CHECK(cudaMemcpyAsync(ob->boxes.p, aabb6, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    launch_is_visible(ob->vs.g, ob->vs.views.p, ob->vs.depth.p, ob->vs.mips.p, ob->hierarchyDirty ? 0 : 1, ob->boxes.p, (int)n, ob->visOut.p, st);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(out, ob->visOut.p, n, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return U3D_OK;
}

int u3d_ob_is_visible(u3d_ob* ob, const float aabb6[6])
{
    uint8_t r = 1;
    int rc = u3d_ob_is_visible_many(ob, aabb6, 1, &r);
    return rc < 0 ? rc : (int)r;
}

int u3d_ob_get_buffer(u3d_ob* ob, int32_t* out_depth)
{
    if (!ob || !out_depth || !ob->allocated)
        return fail(U3D_ERR_INVALID, "u3d_ob_get_buffer: NULL argument or no buffer");
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    CU_CHECK(cudaMemcpyAsync(out_depth, ob->vs.depth.p, (size_t)ob->width * ob->height * 4, cudaMemcpyDeviceToHost, ob->ctx->stream));
    CU_CHECK(cudaStreamSynchronize(ob->ctx->stream));
    return U3D_OK;
}

int u3d_ob_get_mip(u3d_ob* ob, int level, int32_t* out_minmax)
{
    if (!ob || !out_minmax || !ob->allocated || level < 0 || level >= ob->vs.g.nlev)
        return fail(U3D_ERR_INVALID, "u3d_ob_get_mip: bad argument");
    CU_CHECK(cudaSetDevice(ob->ctx->device));
    const BufGeom& g = ob->vs.g;
    CU_CHECK(cudaMemcpyAsync(out_minmax, ob->vs.mips.p + g.moff[level], (size_t)g.mw[level] * g.mh[level] * 8, cudaMemcpyDeviceToHost, ob->ctx->stream));
    CU_CHECK(cudaStreamSynchronize(ob->ctx->stream));
    return U3D_OK;
}

int u3d_ob_num_levels(const u3d_ob* ob) { return ob && ob->allocated ? ob->vs.g.nlev : 0; }

int u3d_ob_mip_size(const u3d_ob* ob, int level, int* width, int* height)
{
    if (!ob || !ob->allocated || level < 0 || level >= ob->vs.g.nlev)
        return fail(U3D_ERR_INVALID, "u3d_ob_mip_size: bad argument");
    if (width) *width = ob->vs.g.mw[level];
    if (height) *height = ob->vs.g.mh[level];
    return U3D_OK;
}

int u3d_ob_get_width(const u3d_ob* ob) { return ob ? ob->width : 0; }
int u3d_ob_get_height(const u3d_ob* ob) { return ob ? ob->height : 0; }
uint32_t u3d_ob_get_num_triangles(const u3d_ob* ob) { return ob ? ob->numTriangles : 0; }
uint32_t u3d_ob_get_max_triangles(const u3d_ob* ob) { return ob ? ob->maxTriangles : 0; }
int u3d_ob_get_cull_mode(const u3d_ob* ob) { return ob ? ob->cullMode : 0; }
int u3d_ob_is_threaded(const u3d_ob* ob) { return ob ? (int)ob->threaded : 0; }

int u3d_ob_get_view_proj(const u3d_ob* ob, float out16[16])
{
    if (!ob || !out16)
        return fail(U3D_ERR_INVALID, "NULL argument");
    std::memcpy(out16, ob->viewProj, sizeof(ob->viewProj));
    return U3D_OK;
}

int u3d_octree_query(u3d_scene* scene, u3d_ob* ob, const float planes24[24], uint32_t drawable_flags, uint32_t view_mask, int query_mode, int stage,
    uint32_t* out_ids, uint32_t capacity, uint32_t* out_query_count, uint32_t* out_count)
{
    if (!scene || !planes24 || (capacity && !out_ids))
        return fail(U3D_ERR_INVALID, "u3d_octree_query: NULL argument");
    if (query_mode < 0 || (query_mode > 5 && query_mode != QUERY_ZONE_OCCLUDER) || stage < 0 || stage > 1)
        return fail(U3D_ERR_INVALID, "u3d_octree_query: bad mode/stage");
    u3d_ctx* ctx = scene->ctx;
    CU_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool useOb = ob && ob->allocated;
    if (query_mode == QUERY_OCCLUDED && !useOb)
        return fail(U3D_ERR_INVALID, "u3d_octree_query: occluded query needs an occlusion buffer with a size");
    // scratch lives in the ob when there is one, else in a temporary
    u3d_ob tmp;
    u3d_ob* holder = ob ? ob : &tmp;
    if (ob && ob->ctx != ctx)
        return fail(U3D_ERR_INVALID, "u3d_octree_query: the occlusion buffer belongs to another context than the scene");
    holder->ctx = ctx;
    ViewConst vc{};
    if (useOb)
        std::memcpy(vc.vp, ob->viewProj, sizeof(vc.vp));
    std::memcpy(vc.planes, planes24, sizeof(vc.planes));
    vc.drawableFlags = drawable_flags;
    vc.viewMask = view_mask;
    vc.cullMode = 0;
    vc.queryMode = query_mode;
    vc.useOcclusion = useOb ? 1 : 0;
    BufGeom g = useOb ? ob->vs.g : make_geom(2, 2);
    CU_CHECK(holder->vs.views.ensure(1));
    CU_CHECK(holder->queryFlags.ensure(1));
    CU_CHECK(holder->octState.ensure((size_t)scene->dev.numOctants));
    CU_CHECK(holder->outIdx.ensure(std::max<uint32_t>(capacity, 1)));
    CU_CHECK(holder->outCounts.ensure(2));
    CU_CHECK(cudaMemcpyAsync(holder->vs.views.p, &vc, sizeof(vc), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemsetAsync(holder->queryFlags.p, 0, 4, st));
    launch_cull(scene->dev, g, holder->vs.views.p, 1, useOb ? ob->vs.depth.p : nullptr, useOb ? ob->vs.mips.p : nullptr,
        useOb && !ob->hierarchyDirty ? 1 : 0, holder->octState.p, scene->dev.numOctants, stage, holder->outIdx.p, capacity, holder->outCounts.p, nullptr,
        holder->queryFlags.p, false, nullptr, st);
    CU_CHECK(cudaGetLastError());
    uint32_t counts[2] = {0, 0};
    CU_CHECK(cudaMemcpyAsync(counts, holder->outCounts.p, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    const uint32_t n = std::min(counts[1], capacity);
    if (n)
    {
        CU_CHECK(cudaMemcpyAsync(out_ids, holder->outIdx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    if (out_query_count)
        *out_query_count = counts[0];
    if (out_count)
        *out_count = counts[1];
    if (counts[1] > capacity)
        return fail(U3D_ERR_CAPACITY, "u3d_octree_query: result larger than capacity");
    return U3D_OK;
}

// ---- ray queries ------------------------------------------------------------------------------------------------------------
} // extern "C"

namespace
{
struct RayKey
{
    float key;
    uint32_t val;
};

/// The reference's Sort (Container/Sort.h:70-141) for (key, payload) records ordered by `key <`: quicksort passes with a median-of-three
/// pivot and Hoare partitioning while a range is longer than 16, then one insertion sort over everything.  Not stable: where keys tie, the
/// result order is whatever this procedure leaves, and Octree::Raycast / RaycastSingle expose that order, so it is followed step by step.
void urho_quick_pass(RayKey* a, long lo, long hi)
{
    while (hi - lo > 16)
    {
        long p = lo + (hi - lo) / 2;
        if (a[lo].key < a[p].key && a[hi - 1].key < a[lo].key)
            p = lo;
        else if (a[hi - 1].key < a[p].key && a[lo].key < a[hi - 1].key)
            p = hi - 1;
        const RayKey pivot = a[p];
        long i = lo - 1, j = hi;
        for (;;)
        {
            do
                --j;
            while (pivot.key < a[j].key);
            do
                ++i;
            while (a[i].key < pivot.key);
            if (i >= j)
                break;
            std::swap(a[i], a[j]);
        }
        urho_quick_pass(a, lo, j + 1);
        lo = j + 1;
    }
}

void urho_sort(RayKey* a, long n)
{
    urho_quick_pass(a, 0, n);
    for (long i = 1; i < n; ++i)
    {
        const RayKey t = a[i];
        long j = i;
        for (; j > 0 && t.key < a[j - 1].key; --j)
            a[j] = a[j - 1];
        a[j] = t;
    }
}
} // namespace

extern "C"
{

} // extern "C"

/// Device copy of the drawable id -> DFS slot map (uploaded on first use).
static int ensure_slot_map(u3d_scene* s, cudaStream_t st)
{
    if (!s->slotOfIdUploaded)
    {
        CU_CHECK(s->slotOfIdDev.ensure(std::max<size_t>(s->slotOfId.size(), 1)));
        if (!s->slotOfId.empty())
            CU_CHECK(cudaMemcpyAsync(s->slotOfIdDev.p, s->slotOfId.data(), s->slotOfId.size() * 4, cudaMemcpyHostToDevice, st));
        s->slotOfIdUploaded = true;
    }
    return U3D_OK;
}

extern "C"
{

int u3d_scene_set_shadow_params(u3d_scene* s, uint32_t n, const uint32_t* shadow_mask, const float* shadow_distance)
{
    if (!s)
        return fail(U3D_ERR_INVALID, "u3d_scene_set_shadow_params: NULL argument");
    CU_CHECK(cudaSetDevice(s->ctx->device));
    const uint32_t S = std::max<uint32_t>((uint32_t)s->dev.numDrawables, 1);
    std::vector<uint32_t> m(S, 0xFFFFFFFFu);
    std::vector<float> d(S, 0.0f);
    for (uint32_t id = 0; id < n && id < s->slotOfId.size(); ++id)
        if (s->slotOfId[id] != 0xFFFFFFFFu)
        {
            if (shadow_mask)
                m[s->slotOfId[id]] = shadow_mask[id];
            if (shadow_distance)
                d[s->slotOfId[id]] = shadow_distance[id];
        }
    CU_CHECK(s->shadowMask.ensure(S));
    CU_CHECK(s->shadowDistance.ensure(S));
    CU_CHECK(cudaMemcpy(s->shadowMask.p, m.data(), (size_t)S * 4, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMemcpy(s->shadowDistance.p, d.data(), (size_t)S * 4, cudaMemcpyHostToDevice));
    s->hasShadowParams = true;
    return U3D_OK;
}

int u3d_batch_process_shadow_casters(u3d_batch* b, const u3d_shadow_split* splits, const float main_view12[12], const float main_camera_position[3],
    int main_orthographic, const uint8_t* in_view, uint32_t n_in_view, void* stream)
{
    static_assert(sizeof(u3d_shadow_split) == sizeof(ShadowSplitDev), "u3d_shadow_split and ShadowSplitDev must have the same layout");
    if (!b || !splits || !main_view12 || !main_camera_position || (n_in_view && !in_view))
        return fail(U3D_ERR_INVALID, "u3d_batch_process_shadow_casters: NULL argument");
    if (!b->numViews)
        return U3D_OK;
    u3d_scene* s = b->scene;
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    int rc = ensure_slot_map(s, st);
    if (rc < 0)
        return rc;
    CU_CHECK(s->splits.ensure(b->numViews));
    CU_CHECK(cudaMemcpyAsync(s->splits.p, splits, (size_t)b->numViews * sizeof(ShadowSplitDev), cudaMemcpyHostToDevice, st));
    if (n_in_view)
    {
        CU_CHECK(s->inView.ensure(n_in_view));
        CU_CHECK(cudaMemcpyAsync(s->inView.p, in_view, n_in_view, cudaMemcpyHostToDevice, st));
    }
    const float4 mainPos = make_float4(main_camera_position[0], main_camera_position[1], main_camera_position[2], main_orthographic ? 1.0f : 0.0f);
    const float4 mainViewZ = make_float4(main_view12[8], main_view12[9], main_view12[10], main_view12[11]);
    launch_shadow_casters(s->dev, s->splits.p, (int)b->numViews, s->slotOfIdDev.p, s->hasShadowParams ? s->shadowMask.p : nullptr,
        s->hasShadowParams ? s->shadowDistance.p : nullptr, n_in_view ? s->inView.p : nullptr, n_in_view, mainPos, mainViewZ, b->outIdx.p, b->outCap,
        b->outCounts.p, st);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaStreamSynchronize(st)); // `splits` / `in_view` are the caller's host memory
    return U3D_OK;
}

int u3d_sort_by_key(uint32_t n, float* keys, uint32_t* values)
{
    if (n && (!keys || !values))
        return fail(U3D_ERR_INVALID, "u3d_sort_by_key: NULL argument");
    std::vector<RayKey> a(n);
    for (uint32_t i = 0; i < n; ++i)
        a[i] = RayKey{keys[i], values[i]};
    urho_sort(a.data(), (long)n);
    for (uint32_t i = 0; i < n; ++i)
    {
        keys[i] = a[i].key;
        values[i] = a[i].val;
    }
    return U3D_OK;
}

// Multi-GPU host logic: cost-balanced assignment of views to ranks.  Same procedure as urho3d_b200/sharding.py:balance_views:
// columns normalised to sum 1, views taken in decreasing order of their normalised total and given to the rank whose worst column stays
// lowest among the ranks with room left (every rank keeps the view count of the interleaved rule), then a few swaps between the rank holding
// the worst column and the rank lightest in that column.
int u3d_balance_views(uint32_t n_views, uint32_t n_costs, const double* costs, uint32_t world, uint32_t* owner_out)
{
    if (!world || (n_views && (!owner_out || (n_costs && !costs))))
        return fail(U3D_ERR_INVALID, "u3d_balance_views: NULL argument or zero ranks");
    const size_t n = n_views, D0 = n_costs, W = world;
    std::vector<double> tot(D0, 0.0);
    auto sane = [](double v) { return std::isfinite(v) && v > 0.0 ? v : 0.0; };
    for (size_t v = 0; v < n; ++v)
        for (size_t c = 0; c < D0; ++c)
            tot[c] += sane(costs[v * D0 + c]);
    std::vector<size_t> cols;
    for (size_t c = 0; c < D0; ++c)
        if (tot[c] > 0.0)
            cols.push_back(c);
    const size_t D = cols.size();
    if (!D)
    {
        for (size_t v = 0; v < n; ++v)
            owner_out[v] = (uint32_t)(v % W); // no information: the interleaved rule
        return U3D_OK;
    }
    std::vector<double> c(n * D), key(n, 0.0);
    for (size_t v = 0; v < n; ++v)
        for (size_t k = 0; k < D; ++k)
        {
            c[v * D + k] = sane(costs[v * D0 + cols[k]]) / tot[cols[k]];
            key[v] += c[v * D + k];
        }
    std::vector<uint32_t> order(n);
    for (size_t v = 0; v < n; ++v)
        order[v] = (uint32_t)v;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] > key[b]; });
    std::vector<size_t> quota(W), held(W, 0);
    for (size_t r = 0; r < W; ++r)
        quota[r] = n / W + (r < n % W ? 1 : 0); // len(range(r, n, W))
    std::vector<double> load(W * D, 0.0);
    auto worst_of = [&](size_t r) { return *std::max_element(load.begin() + r * D, load.begin() + (r + 1) * D); };
    for (uint32_t v : order)
    {
        size_t best = W;
        double bestWorst = 0.0;
        for (size_t r = 0; r < W; ++r)
        {
            if (held[r] >= quota[r])
                continue;
            double w = 0.0;
            for (size_t k = 0; k < D; ++k)
                w = std::max(w, load[r * D + k] + c[(size_t)v * D + k]);
            if (best == W || w < bestWorst)
            {
                best = r;
                bestWorst = w;
            }
        }
        owner_out[v] = (uint32_t)best;
        ++held[best];
        for (size_t k = 0; k < D; ++k)
            load[best * D + k] += c[(size_t)v * D + k];
    }
    for (size_t it = 0; it < 4 * W; ++it)
    {
        const size_t top = (size_t)(std::max_element(load.begin(), load.end()) - load.begin());
        const size_t a = top / D, d = top % D;
        size_t b = 0;
        for (size_t r = 1; r < W; ++r)
            if (load[r * D + d] < load[b * D + d])
                b = r;
        if (a == b)
            break;
        std::vector<uint32_t> ia, ib;
        for (size_t v = 0; v < n; ++v)
            if (owner_out[v] == a)
                ia.push_back((uint32_t)v);
            else if (owner_out[v] == b)
                ib.push_back((uint32_t)v);
        if (ia.empty() || ib.empty())
            break;
        const size_t sa = std::max<size_t>(1, ia.size() / 512), sb = std::max<size_t>(1, ib.size() / 512); // a bounded number of candidate pairs
        const double before = std::max(worst_of(a), worst_of(b));
        double bestAfter = before;
        long bi = -1, bj = -1;
        for (size_t i = 0; i < ia.size(); i += sa)
            for (size_t j = 0; j < ib.size(); j += sb)
            {
                double after = 0.0;
                for (size_t k = 0; k < D; ++k)
                {
                    const double delta = c[(size_t)ib[j] * D + k] - c[(size_t)ia[i] * D + k]; // what rank a gains by trading view i for view j
                    after = std::max(after, std::max(load[a * D + k] + delta, load[b * D + k] - delta));
                }
                if (after < bestAfter)
                {
                    bestAfter = after;
                    bi = (long)i;
                    bj = (long)j;
                }
            }
        if (bi < 0 || bestAfter >= before * (1.0 - 1e-12))
            break;
        for (size_t k = 0; k < D; ++k)
        {
            const double delta = c[(size_t)ib[bj] * D + k] - c[(size_t)ia[bi] * D + k];
            load[a * D + k] += delta;
            load[b * D + k] -= delta;
        }
        owner_out[ia[bi]] = (uint32_t)b;
        owner_out[ib[bj]] = (uint32_t)a;
    }
    return U3D_OK;
}

int u3d_scene_raycast(u3d_scene* s, uint32_t n_rays, const float* rays7, uint32_t drawable_flags, uint32_t view_mask, int single, uint32_t capacity,
    uint32_t* counts, u3d_ray_hit* hits, void* stream)
{
    if (!s || (n_rays && (!rays7 || !counts)) || (capacity && !hits))
        return fail(U3D_ERR_INVALID, "u3d_scene_raycast: NULL argument");
    if (!n_rays)
        return U3D_OK;
    if (n_rays > 65535)
        return fail(U3D_ERR_UNSUPPORTED, "u3d_scene_raycast: at most 65535 rays per call");
    u3d_ctx* ctx = s->ctx;
    CU_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->pick(stream);
    // RaycastSingle needs EVERY candidate of the touched octants (its winner is decided by the sort); size the device lists accordingly
    const uint32_t devCap = std::max<uint32_t>(capacity, 1);
    int rcMap = ensure_slot_map(s, st);
    if (rcMap < 0)
        return rcMap;
    s->rayHost.assign(n_rays, ViewConst{});
    for (uint32_t i = 0; i < n_rays; ++i)
    {
        ViewConst& vc = s->rayHost[i];
        std::memcpy(vc.planes, rays7 + 7 * (size_t)i, 7 * sizeof(float));
        vc.drawableFlags = drawable_flags;
        vc.viewMask = view_mask;
        vc.queryMode = single ? QUERY_RAY_CANDIDATES : QUERY_RAY;
    }
    CU_CHECK(s->rayViews.ensure(n_rays));
    CU_CHECK(s->rayOctState.ensure((size_t)n_rays * s->dev.numOctants));
    CU_CHECK(s->rayIdx.ensure((size_t)n_rays * devCap));
    CU_CHECK(s->rayDist.ensure((size_t)n_rays * devCap));
    CU_CHECK(s->rayCounts.ensure((size_t)n_rays * 2));
    CU_CHECK(s->rayFlags.ensure(1));
    CU_CHECK(cudaMemcpyAsync(s->rayViews.p, s->rayHost.data(), (size_t)n_rays * sizeof(ViewConst), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemsetAsync(s->rayFlags.p, 0, 4, st));
    const BufGeom g = make_geom(2, 2);
    launch_cull(s->dev, g, s->rayViews.p, (int)n_rays, nullptr, nullptr, 0, s->rayOctState.p, s->dev.numOctants, STAGE_QUERY, s->rayIdx.p, devCap,
        s->rayCounts.p, nullptr, s->rayFlags.p, false, nullptr, st);
    launch_ray_distances(s->dev, s->rayViews.p, (int)n_rays, s->slotOfIdDev.p, s->rayIdx.p, devCap, s->rayCounts.p, s->rayDist.p, st);
    CU_CHECK(cudaGetLastError());
    s->rayHostCounts.resize((size_t)n_rays * 2);
    CU_CHECK(cudaMemcpyAsync(s->rayHostCounts.data(), s->rayCounts.p, (size_t)n_rays * 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    bool overflow = false;
    for (uint32_t i = 0; i < n_rays; ++i)
        overflow |= s->rayHostCounts[2 * i + 1] > devCap;
    s->rayHostIdx.resize((size_t)n_rays * devCap);
    s->rayHostDist.resize((size_t)n_rays * devCap);
    CU_CHECK(cudaMemcpyAsync(s->rayHostIdx.data(), s->rayIdx.p, (size_t)n_rays * devCap * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(s->rayHostDist.data(), s->rayDist.p, (size_t)n_rays * devCap * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    std::vector<RayKey> keys;
    for (uint32_t i = 0; i < n_rays; ++i)
    {
        const float* ray = rays7 + 7 * (size_t)i;
        const uint32_t m = std::min(s->rayHostCounts[2 * i + 1], devCap);
        keys.resize(m);
        for (uint32_t k = 0; k < m; ++k)
            keys[k] = RayKey{s->rayHostDist[(size_t)i * devCap + k], s->rayHostIdx[(size_t)i * devCap + k]};
        // Octree::Raycast: Sort(result_, CompareRayQueryResults), Octree.cpp:507; RaycastSingle: Sort(drawables, CompareDrawables), Octree.cpp:525
        urho_sort(keys.data(), (long)m);
        uint32_t nOut = m;
        if (single)
            // the first sorted candidate is the only one that can pass `sortValue < Min(closestHit, maxDistance)` (Octree.cpp:528-541)
            nOut = (m && keys[0].key < ray[6]) ? 1u : 0u;
        counts[i] = nOut;
        for (uint32_t k = 0; k < nOut && k < capacity; ++k)
        {
            u3d_ray_hit& h = hits[(size_t)i * capacity + k];
            const float dist = keys[k].key;
            // RayQueryResult of Drawable::ProcessRayQuery, Drawable.cpp:123-130
            h.position[0] = ray[0] + dist * ray[3];
            h.position[1] = ray[1] + dist * ray[4];
            h.position[2] = ray[2] + dist * ray[5];
            h.normal[0] = -ray[3];
            h.normal[1] = -ray[4];
            h.normal[2] = -ray[5];
            h.distance = dist;
            h.drawable = keys[k].val;
        }
    }
    if (overflow)
        return fail(U3D_ERR_CAPACITY, "u3d_scene_raycast: a ray produced more hits/candidates than capacity");
    return U3D_OK;
}

// ---- batched pipeline -------------------------------------------------------------------------------------------------------
int u3d_batch_create(u3d_ctx* ctx, u3d_scene* scene, u3d_occluders* occluders, int width, int height, uint32_t max_views, uint32_t out_capacity,
    uint64_t tri_pool, u3d_batch** out)
{
    if (!ctx || !scene || !out || !max_views)
        return fail(U3D_ERR_INVALID, "u3d_batch_create: NULL/zero argument");
    if (max_views > 65535)
        return fail(U3D_ERR_UNSUPPORTED, "u3d_batch_create: at most 65535 views per batch object");
    *out = nullptr;
    const bool hasBuffers = occluders != nullptr;
    if (hasBuffers)
    {
        if (height & 1)
            ++height;
        if (width < 2 || height < 2 || (width & (width - 1)) || width > 8192 || height > 16384)
            return fail(U3D_ERR_INVALID, "u3d_batch_create: width must be a power of two in 2..8192, height in 2..16384");
    }
    else
    {
        width = 2;
        height = 2;
    }
    CU_CHECK(cudaSetDevice(ctx->device));
    u3d_batch* b = new (std::nothrow) u3d_batch();
    if (!b)
        return fail(U3D_ERR_NOMEM, "out of host memory");
    b->ctx = ctx;
    b->scene = scene;
    b->occ = occluders;
    b->outCap = out_capacity;
    b->hasBuffers = hasBuffers;
    const uint32_t itemCap = hasBuffers ? std::max<uint32_t>(occluders->totalItems, 1) : 1;
    uint64_t pool = 1;
    if (hasBuffers)
    {
        // worst case = every occluder of every view; default caps the pool at 2 GiB of records and relies on the overflow report
        const uint64_t worst = (uint64_t)max_views * (occluders->totalTris + 64);
        pool = tri_pool ? tri_pool : std::min<uint64_t>(worst, (2ull << 30) / sizeof(TriSetup));
    }
    b->poolTris = pool;
    int rc = b->vs.alloc(ctx, width, height, hasBuffers ? max_views : 1, itemCap, pool, hasBuffers);
    cudaError_t e = cudaSuccess;
    if (rc >= 0)
    {
        // ViewSet::alloc sized views for `nviews`; frustum-only batches still need max_views view constants
        if ((e = b->vs.views.ensure(max_views)) || (e = b->octState.ensure((size_t)max_views * ((scene->dev.numOctants + 31) / 32 * 32))) ||
            (e = b->outIdx.ensure((size_t)max_views * std::max<uint32_t>(out_capacity, 1))) || (e = b->outCounts.ensure((size_t)max_views * 2)) ||
            (e = b->packOffsets.ensure((size_t)max_views + 1)) || (e = b->zRange.ensure((size_t)max_views * 2)) || (e = b->hostViews.ensure(max_views)) || (e = b->hostCounts.ensure((size_t)max_views * 2 + 2)))
            rc = fail(U3D_ERR_CUDA, std::string("u3d_batch_create: ") + cudaGetErrorString(e));
    }
    if (rc < 0)
    {
        delete b;
        return rc;
    }
    b->vs.maxViews = max_views;
    *out = b;
    return U3D_OK;
}

int u3d_batch_set_scene(u3d_batch* b, u3d_scene* scene)
{
    if (!b || !scene)
        return fail(U3D_ERR_INVALID, "u3d_batch_set_scene: NULL argument");
    if (scene->ctx != b->ctx)
        return fail(U3D_ERR_INVALID, "u3d_batch_set_scene: the scene belongs to another context");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    b->scene = scene;
    // the per-view octant states (and, at the next run, the dense path's scratch) follow the new scene's octant count
    CU_CHECK(b->octState.ensure((size_t)b->vs.maxViews * ((scene->dev.numOctants + 31) / 32 * 32)));
    return U3D_OK;
}

void u3d_batch_destroy(u3d_batch* b)
{
    if (!b)
        return;
    cudaSetDevice(b->ctx->device);
    delete b;
}

int u3d_batch_set_views(u3d_batch* b, const u3d_view* views, uint32_t n, void* stream)
{
    if (!b || (n && !views))
        return fail(U3D_ERR_INVALID, "u3d_batch_set_views: NULL argument");
    if (n > b->vs.maxViews)
        return fail(U3D_ERR_CAPACITY, "u3d_batch_set_views: more views than max_views");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    // the pinned staging array may still be in flight from the previous call (on this or another stream)
    if (!b->haveViewsEvent)
    {
        CU_CHECK(cudaEventCreateWithFlags(&b->viewsCopied, cudaEventDisableTiming));
        b->haveViewsEvent = true;
    }
    else
        CU_CHECK(cudaEventSynchronize(b->viewsCopied));
    for (uint32_t i = 0; i < n; ++i)
    {
        const u3d_view& v = views[i];
        ViewConst& vc = b->hostViews.p[i];
        mul_proj_view(v.projection, v.view, vc.vp);
        std::memcpy(vc.planes, v.planes, sizeof(vc.planes));
        vc.drawableFlags = v.drawable_flags;
        vc.viewMask = v.view_mask;
        vc.cullMode = flip_cull(b->occ ? b->occ->cullMode : CULL_CCW, v.reverse_culling != 0);
        vc.queryMode = v.query_mode;
        vc.useOcclusion = (v.use_occlusion && b->hasBuffers) ? 1 : 0;
        if (vc.queryMode == QUERY_OCCLUDED && !vc.useOcclusion)
            vc.queryMode = QUERY_FRUSTUM; // View.cpp:873-883: without a buffer the plain frustum query is used
        vc.ortho = v.orthographic != 0;
        vc.viewZ[0] = v.view[8];
        vc.viewZ[1] = v.view[9];
        vc.viewZ[2] = v.view[10];
        vc.viewZ[3] = v.view[11];
        vc.camPos[0] = v.camera_position[0];
        vc.camPos[1] = v.camera_position[1];
        vc.camPos[2] = v.camera_position[2];
        vc.pad = 0;
    }
    b->numViews = n;
    if (n)
        CU_CHECK(cudaMemcpyAsync(b->vs.views.p, b->hostViews.p, (size_t)n * sizeof(ViewConst), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaEventRecord(b->viewsCopied, st));
    return U3D_OK;
}

int u3d_batch_run_part(u3d_batch* b, int part, int stage, void* stream)
{
    if (!b || stage < 0 || stage > 2 || part < 0 || part > 2)
        return fail(U3D_ERR_INVALID, "u3d_batch_run: bad argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const int V = (int)b->numViews;
    b->lastLaunches = 0;
    if (!V)
        return U3D_OK;
    ViewSet& vs = b->vs;
    const bool timed = b->timed;
    b->timed = false;
    auto mark = [&](int phase) -> cudaError_t { return timed ? cudaEventRecord(b->ev[phase], st) : cudaSuccess; };
    CU_CHECK(mark(0));
    if (part != 2)
        CU_CHECK(vs.clear_counters(st)); // flags, per-view triangle / item counters, tile bins: one block, one memset
    if (b->hasBuffers && part != 2)
    {
        const u3d_occluders* oc = b->occ;
        if (!vs.useTiles)
        {
            launch_clear(vs.depth.p, (size_t)V * vs.g.w * vs.g.h, st);
            b->lastLaunches += 1;
        }
        if (b->rankOccluders && b->serialOccluders)
        {
            // View::DrawOccluders' non-threaded branch: sorted list -> serial per-view draw loop (general-path fill) -> whole hierarchy
            constexpr uint32_t kSerialScratch = 2048; // set-up triangles one 256-triangle chunk of an occluder may produce
            CU_CHECK(b->rankList.ensure((size_t)vs.maxViews * kRankCap));
            CU_CHECK(b->rankCount.ensure(vs.maxViews));
            CU_CHECK(b->activeOccluders.ensure(vs.maxViews));
            CU_CHECK(b->serialScratch.ensure((size_t)vs.maxViews * kSerialScratch));
            if (vs.useTiles)
            {
                launch_clear(vs.depth.p, (size_t)V * vs.g.w * vs.g.h, st);
                b->lastLaunches += 1;
            }
            CU_CHECK(mark(1));
            launch_select_ranked(vs.views.p, b->camParams.p, V, (int)oc->n, oc->aabb6.p, oc->drawDist.n ? oc->drawDist.p : nullptr, oc->mesh.p, oc->start.p,
                oc->count.p, b->sizeThreshold, b->maxTriangles, vs.items.p, vs.itemCap, vs.itemCount.p, vs.submitted.p, vs.flags.p, b->rankList.p,
                b->rankCount.p, st);
            CU_CHECK(mark(2));
            CU_CHECK(mark(3));
            launch_draw_occluders_serial(vs.g, vs.views.p, V, b->rankList.p, b->rankCount.p, oc->aabb6.p, oc->model12.p, oc->mesh.p, oc->start.p, oc->count.p,
                oc->meshTable.p, b->maxTriangles, vs.depth.p, b->serialScratch.p, kSerialScratch, vs.submitted.p, vs.drawnSources.p, b->activeOccluders.p,
                vs.flags.p, st);
            CU_CHECK(mark(4));
            CU_CHECK(mark(5));
            launch_hierarchy(vs.g, V, vs.depth.p, vs.mips.p, 0, st);
            b->lastLaunches += 2 + (uint32_t)vs.g.nlev;
            CU_CHECK(cudaGetLastError());
        }
        else
        {
        CU_CHECK(mark(1));
        if (b->rankOccluders)
            launch_select_ranked(vs.views.p, b->camParams.p, V, (int)oc->n, oc->aabb6.p, oc->drawDist.n ? oc->drawDist.p : nullptr, oc->mesh.p, oc->start.p,
                oc->count.p, b->sizeThreshold, b->maxTriangles, vs.items.p, vs.itemCap, vs.itemCount.p, vs.submitted.p, vs.flags.p, nullptr, nullptr, st);
        else
            launch_select(vs.views.p, V, (int)oc->n, oc->aabb6.p, oc->mesh.p, oc->start.p, oc->count.p, vs.items.p, vs.itemCap, vs.itemCount.p, vs.submitted.p, st);
        CU_CHECK(mark(2));
        launch_scan_regions(V, vs.submitted.p, 64, b->poolTris, vs.triBase.p, vs.triCap.p, vs.flags.p, st);
        const bool twoPhase = g_twoPhase && vs.useTiles;
        b->countPending = false;
        RasterDev rd = vs.raster_dev(vs.useTiles, 1);
        if (twoPhase)
        {
            CU_CHECK(b->itemPass.ensure((size_t)vs.maxViews * std::max<uint32_t>(vs.itemCap, 1)));
            CU_CHECK(b->nearList.ensure((size_t)vs.maxViews * std::max<uint32_t>(vs.itemCap, 1)));
            CU_CHECK(b->farVisList.ensure((size_t)vs.maxViews * std::max<uint32_t>(vs.itemCap, 1)));
            CU_CHECK(b->listCounts.ensure((size_t)vs.maxViews * 3)); // near, far-visible, far
            launch_items_near_far(vs.views.p, V, vs.items.p, vs.itemCap, vs.itemCount.p, oc->aabb6.p, b->itemPass.p, b->nearList.p, b->listCounts.p,
                b->listCounts.p + 2 * (size_t)vs.maxViews, b->listCounts.p + vs.maxViews, (uint32_t)g_twoPhaseMin, g_twoPhaseShift, st);
            rd.itemList = b->nearList.p;
            rd.listCount = b->listCounts.p;
            b->lastLaunches += 1;
        }
        CU_CHECK(mark(3));
        launch_setup(vs.g, vs.tg, vs.views.p, V, vs.items.p, vs.itemCap, oc->totalItems, vs.itemCount.p, oc->model12.p, oc->meshTable.p, rd, st);
        CU_CHECK(mark(4));
        int firstLevel = 0;
        if (vs.useTiles)
        {
            launch_fill_tiles(vs.g, vs.tg, V, vs.depth.p, vs.mips.p, rd, st);
            firstLevel = vs.tg.fused;
            b->lastLaunches += 3;
        }
        else
        {
            launch_fill_global(vs.g, V, vs.depth.p, vs.pool.p, vs.triBase.p, vs.triCap.p, vs.triCount.p, 8, st);
            b->lastLaunches += 1;
        }
        CU_CHECK(mark(5));
        launch_hierarchy(vs.g, V, vs.depth.p, vs.mips.p, firstLevel, st);
        b->lastLaunches += 4 + (uint32_t)(vs.g.nlev - firstLevel); // select, region scan, set-up, clipper + one launch per unfused level
        if (twoPhase)
        {
            // second pass: the occluders that were left out, as far as their boxes show in the first pass's buffer (and hierarchy)
            launch_items_far_test(vs.g, vs.views.p, V, vs.items.p, vs.itemCap, oc->totalItems, b->nearList.p, b->listCounts.p + 2 * (size_t)vs.maxViews,
                oc->aabb6.p, vs.depth.p, vs.mips.p, b->itemPass.p, b->farVisList.p, b->listCounts.p + vs.maxViews, st);
            CU_CHECK(vs.clear_pass_counters(st));
            rd.itemList = b->farVisList.p;
            rd.listCount = b->listCounts.p + vs.maxViews;
            rd.secondPass = 1;
            launch_setup(vs.g, vs.tg, vs.views.p, V, vs.items.p, vs.itemCap, oc->totalItems, vs.itemCount.p, oc->model12.p, oc->meshTable.p, rd, st);
            launch_fill_tiles(vs.g, vs.tg, V, vs.depth.p, vs.mips.p, rd, st);
            launch_hierarchy(vs.g, V, vs.depth.p, vs.mips.p, firstLevel, st);
            b->lastLaunches += 5 + (uint32_t)(vs.g.nlev - firstLevel);
            b->countPending = true;
        }
        CU_CHECK(cudaGetLastError());
        }
    }
    else
    {
        for (int p = 1; p <= 5; ++p)
            CU_CHECK(mark(p));
    }
    CU_CHECK(mark(6));
    if (part != 1)
    {
        const SceneDev& sc = b->scene->dev;
        const int stateStride = (sc.numOctants + 31) / 32 * 32;
        CU_CHECK(b->octState.ensure((size_t)vs.maxViews * stateStride));
        const bool dense = (g_cullDense == 2 || (g_cullDense == 1 && V >= kDenseMinViews)) && sc.numLevels <= 30;
        if (dense)
        {
            int rc = b->dense.ensure(vs.maxViews, sc, g_denseQueueCap, g_densePoolCap);
            if (rc < 0)
                return rc;
            CU_CHECK(b->dense.next_epoch(st));
            launch_cull_dense(sc, vs.g, vs.views.p, V, vs.depth.p, vs.mips.p, 1, b->octState.p, stage, b->outIdx.p, b->outCap, b->outCounts.p, b->zRange.p,
                vs.flags.p, b->dense.dev, st, timed ? b->ev + 7 : nullptr, b->pipelined);
            // answered on the device if the dense path's scratch overflowed (returns at once otherwise)
            launch_cull(sc, vs.g, vs.views.p, V, vs.depth.p, vs.mips.p, 1, b->octState.p, stateStride, stage, b->outIdx.p, b->outCap, b->outCounts.p,
                b->zRange.p, vs.flags.p, b->pipelined, b->dense.dev.ctr + 34, st);
            b->lastLaunches += (uint32_t)dense_launch_count(sc) + 1;
        }
        else
        {
            launch_cull(sc, vs.g, vs.views.p, V, vs.depth.p, vs.mips.p, 1, b->octState.p, stateStride, stage, b->outIdx.p, b->outCap, b->outCounts.p,
                b->zRange.p, vs.flags.p, b->pipelined, nullptr, st);
            b->lastLaunches += 1;
            for (int p = 7; p <= 9; ++p)
                CU_CHECK(mark(p));
        }
        CU_CHECK(cudaGetLastError());
    }
    else
    {
        for (int p = 7; p <= 9; ++p)
            CU_CHECK(mark(p));
    }
    CU_CHECK(mark(10));
    return U3D_OK;
}

int u3d_batch_run_timed(u3d_batch* b, int stage, void* stream, float* phase_ms)
{
    if (!b || !phase_ms)
        return fail(U3D_ERR_INVALID, "u3d_batch_run_timed: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    if (!b->haveEvents)
    {
        for (auto& e : b->ev)
            CU_CHECK(cudaEventCreate(&e));
        b->haveEvents = true;
    }
    b->timed = true;
    int rc = u3d_batch_run_part(b, 0, stage, stream);
    if (rc < 0)
        return rc;
    CU_CHECK(cudaStreamSynchronize(b->ctx->pick(stream)));
    for (int p = 0; p < U3D_NUM_PHASES; ++p)
        CU_CHECK(cudaEventElapsedTime(&phase_ms[p], b->ev[p], b->ev[p + 1]));
    return U3D_OK;
}

int u3d_batch_run(u3d_batch* b, int stage, void* stream) { return u3d_batch_run_part(b, 0, stage, stream); }

static int flags_to_error(uint32_t flags);

static int batch_check_flags(u3d_batch* b, cudaStream_t st)
{
    uint32_t flags = 0;
    CU_CHECK(cudaMemcpyAsync(&flags, b->vs.flags.p, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return flags_to_error(flags);
}

int u3d_batch_read_counts(u3d_batch* b, uint32_t* counts, void* stream)
{
    if (!b || !counts)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_counts: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    if (b->numViews)
        CU_CHECK(cudaMemcpyAsync(counts, b->outCounts.p, (size_t)b->numViews * 8, cudaMemcpyDeviceToHost, st));
    return batch_check_flags(b, st);
}

} // extern "C"

namespace u3d
{
__global__ void k_pack_offsets(int numViews, const uint32_t* __restrict__ counts, uint32_t cap, uint32_t* __restrict__ offsets)
{
    // single block exclusive scan of min(count, cap)
    __shared__ uint32_t carry;
    __shared__ uint32_t warpSum[32];
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < numViews; base += blockDim.x)
    {
        const int v = base + threadIdx.x;
        const uint32_t c = v < numViews ? min(counts[2 * v + 1], cap) : 0;
        uint32_t x = c;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
        }
        if (lane == 31)
            warpSum[wid] = x;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i)
        {
            if (i < wid)
                pre += warpSum[i];
            tot += warpSum[i];
        }
        if (v < numViews)
            offsets[v] = carry + pre + x - c;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        offsets[numViews] = carry;
}

/// k_pack_offsets + a copy of everything the host wants to see first, as one block: hdr = flags | counts[V][2] | offsets[V + 1].
__global__ void k_pack_header(int numViews, const uint32_t* __restrict__ counts, uint32_t cap, uint32_t* __restrict__ offsets, const uint32_t* __restrict__ flags,
    uint32_t* __restrict__ hdr)
{
    __shared__ uint32_t carry;
    __shared__ uint32_t warpSum[32];
    if (threadIdx.x == 0)
    {
        carry = 0;
        hdr[0] = *flags;
    }
    __syncthreads();
    uint32_t* hCounts = hdr + 1;
    uint32_t* hOffsets = hdr + 1 + 2 * (size_t)numViews;
    for (int base = 0; base < numViews; base += blockDim.x)
    {
        const int v = base + threadIdx.x;
        uint32_t c = 0;
        if (v < numViews)
        {
            const uint32_t q = counts[2 * v], e = counts[2 * v + 1];
            hCounts[2 * v] = q;
            hCounts[2 * v + 1] = e;
            c = min(e, cap);
        }
        uint32_t x = c;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o)
                x += y;
        }
        if (lane == 31)
            warpSum[wid] = x;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i)
        {
            if (i < wid)
                pre += warpSum[i];
            tot += warpSum[i];
        }
        if (v < numViews)
        {
            const uint32_t off = carry + pre + x - c;
            offsets[v] = off;
            hOffsets[v] = off;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        offsets[numViews] = carry;
        hOffsets[numViews] = carry;
    }
}

__global__ void k_pack_ids(const uint32_t* __restrict__ counts, uint32_t cap, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ids,
    uint32_t* __restrict__ packed)
{
    const int v = blockIdx.x;
    const uint32_t n = min(counts[2 * v + 1], cap);
    const uint32_t* src = ids + (size_t)v * cap;
    uint32_t* dst = packed + offsets[v];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        dst[i] = src[i];
}
} // namespace u3d

extern "C"
{

int u3d_batch_read_packed(u3d_batch* b, uint32_t* offsets, uint32_t* ids, uint64_t capacity, void* stream)
{
    if (!b || !offsets)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_packed: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const int V = (int)b->numViews;
    if (!V)
    {
        offsets[0] = 0;
        return U3D_OK;
    }
    CU_CHECK(b->packed.ensure((size_t)V * std::max<uint32_t>(b->outCap, 1)));
    k_pack_offsets<<<1, 1024, 0, st>>>(V, b->outCounts.p, b->outCap, b->packOffsets.p);
    k_pack_ids<<<V, 256, 0, st>>>(b->outCounts.p, b->outCap, b->packOffsets.p, b->outIdx.p, b->packed.p);
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(offsets, b->packOffsets.p, (size_t)(V + 1) * 4, cudaMemcpyDeviceToHost, st));
    int rc = batch_check_flags(b, st); // synchronises
    if (rc < 0)
        return rc;
    const uint64_t total = offsets[V];
    if (total > capacity)
        return fail(U3D_ERR_CAPACITY, "u3d_batch_read_packed: ids capacity too small");
    if (total && ids)
    {
        CU_CHECK(cudaMemcpyAsync(ids, b->packed.p, total * 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    return U3D_OK;
}

int u3d_batch_pack_device(u3d_batch* b, void* stream, void** offsets_dev, void** packed_dev)
{
    if (!b)
        return fail(U3D_ERR_INVALID, "u3d_batch_pack_device: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const int V = (int)b->numViews;
    CU_CHECK(b->packed.ensure((size_t)std::max(V, 1) * std::max<uint32_t>(b->outCap, 1)));
    if (V)
    {
        k_pack_offsets<<<1, 1024, 0, st>>>(V, b->outCounts.p, b->outCap, b->packOffsets.p);
        k_pack_ids<<<V, 256, 0, st>>>(b->outCounts.p, b->outCap, b->packOffsets.p, b->outIdx.p, b->packed.p);
        CU_CHECK(cudaGetLastError());
    }
    if (offsets_dev)
        *offsets_dev = b->packOffsets.p;
    if (packed_dev)
        *packed_dev = b->packed.p;
    return U3D_OK;
}

} // extern "C" (result exchange kernels follow)

// -----------------------------------------------------------------------------------------------------------------------------
// Multi-GPU result exchange over NVLink peer memory (the one exchange step of the path).
//
// Every rank owns one symmetric block per exchange object:
//   arrive[kMaxRanks] | ack[kMaxRanks] | status | pad | counts[world][per][2] | offsets[world][per + 1] | ids[world][cap]
// A producer packs the visible ids of its views back to back and STORES them straight into region [rank] of every rank's block
// (plain P2P stores through NVLink / NVSwitch; its own block is just another destination), then raises arrive[rank] = epoch in
// every block with a system-scope release.  A consumer waits until all arrive[] flags of its own block reached the epoch, reads,
// and acknowledges by writing ack[rank] = epoch into every producer's block; a producer does not overwrite epoch e before every
// consumer has acknowledged it.  No NCCL call, no host synchronisation, one kernel for pack + transfer + signal.
// -----------------------------------------------------------------------------------------------------------------------------
namespace u3d
{
constexpr int kMaxRanks = 64;
constexpr uint32_t kGatherTimeout = 1u, kGatherOverflow = 2u;

struct GatherLayout
{
    uint32_t per;       // views per rank (upper bound)
    uint64_t cap;       // packed ids per rank
    int world;
    __host__ __device__ size_t countsOff() const { return (size_t)(2 * kMaxRanks + 64) * 4; }
    __host__ __device__ size_t offsetsOff() const { return countsOff() + (size_t)world * per * 8; }
    __host__ __device__ size_t idsOff() const { return (offsetsOff() + (size_t)world * (per + 1) * 4 + 255) & ~(size_t)255; }
    __host__ __device__ size_t bytes() const { return idsOff() + (size_t)world * cap * 4; }
};

struct PeerTable
{
    unsigned char* block[kMaxRanks];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/// One thread per rank: spin until flags[r] >= epoch (epochs only grow).  Gives up after `timeoutNs` and records it in *status
/// instead of hanging the device when a peer died.
__global__ void k_gather_wait(const uint32_t* __restrict__ flags, int world, uint32_t epoch, uint32_t* __restrict__ status, unsigned long long timeoutNs)
{
    const int r = threadIdx.x;
    if (r >= world)
        return;
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(flags + r) - epoch) < 0)
    {
        __nanosleep(200);
        if (global_timer_ns() - t0 > timeoutNs)
        {
            atomicOr(status, kGatherTimeout);
            return;
        }
    }
}

/// One thread per rank: flag[myRank] = epoch in every rank's block.
__global__ void k_gather_signal(PeerTable peers, int world, int myRank, size_t flagOff, uint32_t epoch)
{
    const int r = threadIdx.x;
    if (r >= world)
        return;
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.block[r] + flagOff) + myRank, epoch);
}

/// The whole producer side of the exchange in ONE launch, one block per view slot of this rank (all `per` of them, so that the
/// slots this batch does not use read as empty on every rank):
///   1. wait until every consumer has released the previous epoch's contents of this rank's regions (ack flags of the own block);
///   2. packed offset of the view = sum of the emitted counts of the views before it (every block sums for itself: a few KB of L2 reads
///      instead of a scan kernel in front);
///   3. copy the view's emitted ids into region [myRank] of EVERY rank's block: each element is loaded once and stored with 16-byte
///      peer stores (NVLink packets of 4 ids; the destination is aligned by a scalar head, the source is read as 4 scalars);
///   4. counts and offsets beside them; the last block to finish raises this rank's arrival flag in every block.
__global__ void __launch_bounds__(256) k_push_results(PeerTable peers, GatherLayout lay, int myRank, int numViews, const uint32_t* __restrict__ counts,
    uint32_t outCap, const uint32_t* __restrict__ ids, uint32_t epoch, uint32_t* __restrict__ doneCounter, unsigned long long timeoutNs)
{
    __shared__ uint32_t red[2][8];
    __shared__ bool last;
    const int v = blockIdx.x;
    const int world = lay.world;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* ownStatus = reinterpret_cast<uint32_t*>(peers.block[myRank]) + 2 * kMaxRanks;
    // 1. acknowledgements of the previous epoch
    if (epoch > 1 && tid < world)
    {
        const uint32_t* ack = reinterpret_cast<const uint32_t*>(peers.block[myRank]) + kMaxRanks + tid;
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys(ack) - (epoch - 1)) < 0)
        {
            __nanosleep(200);
            if (global_timer_ns() - t0 > timeoutNs)
            {
                atomicOr(ownStatus, kGatherTimeout);
                break;
            }
        }
    }
    // 2. offsets
    uint32_t before = 0, total = 0;
    for (int u = tid; u < numViews; u += 256)
    {
        const uint32_t c = min(counts[2 * u + 1], outCap);
        total += c;
        if (u < v)
            before += c;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
    {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0)
    {
        red[0][wid] = before;
        red[1][wid] = total;
    }
    __syncthreads(); // also orders the acknowledgement wait before any store below
    before = total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
    {
        before += red[0][w];
        total += red[1][w];
    }
    const bool used = v < numViews;
    const uint32_t off = used ? before : total;
    const uint32_t n = used ? min(counts[2 * v + 1], outCap) : 0u;
    // 3. ids
    uint32_t m = n;
    if ((uint64_t)off + n > lay.cap)
    {
        m = off < lay.cap ? (uint32_t)(lay.cap - off) : 0u;
        if (tid < world) // every consumer must learn that this rank's region is truncated
            atomicOr(reinterpret_cast<uint32_t*>(peers.block[tid]) + 2 * kMaxRanks, kGatherOverflow);
    }
    if (m)
    {
        const uint32_t* src = ids + (size_t)v * outCap;
        const size_t d0 = (size_t)myRank * lay.cap + off;          // first destination element inside the ids region
        const uint32_t head = min(m, (uint32_t)((4 - (d0 & 3)) & 3)); // scalar elements until the destination is 16-byte aligned
        const uint32_t nBody = (m - head) >> 2, tail = head + (nBody << 2);
        for (uint32_t k = tid; k < nBody; k += 256)
        {
            const uint32_t* p = src + head + (k << 2);
            const uint4 x = make_uint4(p[0], p[1], p[2], p[3]);
            for (int r = 0; r < world; ++r)
            {
                const int dst = (myRank + r) % world; // start with the own block, spread the ranks over the links
                *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(peers.block[dst] + lay.idsOff()) + d0 + head + (k << 2)) = x;
            }
        }
        for (uint32_t i = tid; i < head + (m - tail); i += 256)
        {
            const uint32_t e = i < head ? i : tail + (i - head);
            const uint32_t x = src[e];
            for (int r = 0; r < world; ++r)
            {
                const int dst = (myRank + r) % world;
                reinterpret_cast<uint32_t*>(peers.block[dst] + lay.idsOff())[d0 + e] = x;
            }
        }
    }
    // 4. counts, offsets, arrival
    if (tid < world)
    {
        const int dst = tid;
        uint32_t* c = reinterpret_cast<uint32_t*>(peers.block[dst] + lay.countsOff()) + ((size_t)myRank * lay.per + v) * 2;
        c[0] = used ? counts[2 * v] : 0u;
        c[1] = used ? counts[2 * v + 1] : 0u;
        uint32_t* o = reinterpret_cast<uint32_t*>(peers.block[dst] + lay.offsetsOff()) + (size_t)myRank * (lay.per + 1);
        o[v] = off;
        if (v == (int)lay.per - 1)
            o[lay.per] = total;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0)
        last = atomicAdd(doneCounter, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (last)
    {
        if (tid == 0)
            *doneCounter = 0;
        if (tid < world)
        {
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t*>(peers.block[tid]) + myRank, epoch);
        }
    }
}
} // namespace u3d

struct u3d_gather
{
    u3d_ctx* ctx{};
    int world{}, rank{};
    GatherLayout lay{};
    unsigned char* block{};       // own block (cudaMalloc)
    PeerTable peers{};
    bool opened[kMaxRanks]{};     // peers.block[r] came from cudaIpcOpenMemHandle
    DevBuf<uint32_t> doneCounter;
    uint32_t epoch{};             // last epoch pushed by this rank
    uint32_t waited{};            // last epoch waited for
    bool connected{};
    unsigned long long timeoutNs{5000000000ull};
};

extern "C"
{

int u3d_gather_create(u3d_ctx* ctx, int world, int rank, uint32_t views_per_rank, uint64_t ids_capacity_per_rank, u3d_gather** out)
{
    if (!ctx || !out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
        return fail(U3D_ERR_INVALID, "u3d_gather_create: bad argument");
    CU_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<u3d_gather> g(new u3d_gather);
    g->ctx = ctx;
    g->world = world;
    g->rank = rank;
    g->lay.per = std::max<uint32_t>(views_per_rank, 1);
    g->lay.cap = std::max<uint64_t>(ids_capacity_per_rank, 1);
    g->lay.world = world;
    CU_CHECK(cudaMalloc((void**)&g->block, g->lay.bytes()));
    cudaError_t ge = cudaMemset(g->block, 0, g->lay.idsOff());
    if (ge == cudaSuccess)
        ge = g->doneCounter.ensure(1);
    if (ge == cudaSuccess)
        ge = cudaMemset(g->doneCounter.p, 0, 4);
    if (ge == cudaSuccess)
        ge = cudaDeviceSynchronize();
    if (ge != cudaSuccess)
    {
        cudaFree(g->block); // u3d_gather has no destructor of its own
        return fail(U3D_ERR_CUDA, std::string("u3d_gather_create: ") + cudaGetErrorString(ge));
    }
    g->peers.block[rank] = g->block;
    *out = g.release();
    return U3D_OK;
}

void u3d_gather_destroy(u3d_gather* g)
{
    if (!g)
        return;
    cudaSetDevice(g->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < g->world; ++r)
        if (g->opened[r])
            cudaIpcCloseMemHandle(g->peers.block[r]);
    cudaFree(g->block);
    delete g;
}

int u3d_gather_export(u3d_gather* g, void* handle)
{
    if (!g || !handle)
        return fail(U3D_ERR_INVALID, "u3d_gather_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == U3D_IPC_HANDLE_BYTES, "IPC handle size");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    cudaIpcMemHandle_t h;
    CU_CHECK(cudaIpcGetMemHandle(&h, g->block));
    std::memcpy(handle, &h, sizeof(h));
    return U3D_OK;
}

int u3d_gather_connect(u3d_gather* g, const void* handles)
{
    if (!g || !handles)
        return fail(U3D_ERR_INVALID, "u3d_gather_connect: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    for (int r = 0; r < g->world; ++r)
    {
        if (r == g->rank)
            continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const unsigned char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        CU_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g->peers.block[r] = (unsigned char*)p;
        g->opened[r] = true;
    }
    g->connected = true;
    return U3D_OK;
}

int u3d_gather_connect_local(u3d_gather* g, u3d_gather* const* all)
{
    if (!g || !all)
        return fail(U3D_ERR_INVALID, "u3d_gather_connect_local: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    for (int r = 0; r < g->world; ++r)
    {
        if (!all[r] || all[r]->world != g->world || all[r]->rank != r || all[r]->lay.per != g->lay.per || all[r]->lay.cap != g->lay.cap)
            return fail(U3D_ERR_INVALID, "u3d_gather_connect_local: exchange objects do not match");
        if (all[r]->ctx->device != g->ctx->device)
        {
            int can = 0;
            CU_CHECK(cudaDeviceCanAccessPeer(&can, g->ctx->device, all[r]->ctx->device));
            if (!can)
                return fail(U3D_ERR_UNSUPPORTED, "u3d_gather_connect_local: no peer access between the devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                CU_CHECK(e);
            cudaGetLastError();
        }
        g->peers.block[r] = all[r]->block;
    }
    g->connected = true;
    return U3D_OK;
}

int u3d_batch_push_results(u3d_batch* b, u3d_gather* g, void* stream)
{
    if (!b || !g)
        return fail(U3D_ERR_INVALID, "u3d_batch_push_results: NULL argument");
    if (!g->connected && g->world > 1)
        return fail(U3D_ERR_INVALID, "u3d_batch_push_results: exchange object is not connected");
    if (b->numViews > g->lay.per)
        return fail(U3D_ERR_CAPACITY, "u3d_batch_push_results: more views than the exchange object was created for");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const int V = (int)b->numViews;
    const uint32_t epoch = ++g->epoch;
    // one launch: acknowledgement wait + packed offsets + peer stores + arrival flags (k_push_results)
    k_push_results<<<g->lay.per, 256, 0, st>>>(g->peers, g->lay, g->rank, V, b->outCounts.p, b->outCap, b->outIdx.p, epoch, g->doneCounter.p, g->timeoutNs);
    CU_CHECK(cudaGetLastError());
    return U3D_OK;
}

int u3d_gather_wait(u3d_gather* g, void* stream)
{
    if (!g)
        return fail(U3D_ERR_INVALID, "u3d_gather_wait: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    cudaStream_t st = g->ctx->pick(stream);
    g->waited = g->epoch;
    k_gather_wait<<<1, kMaxRanks, 0, st>>>(reinterpret_cast<uint32_t*>(g->block), g->world, g->waited, reinterpret_cast<uint32_t*>(g->block) + 2 * kMaxRanks, g->timeoutNs);
    CU_CHECK(cudaGetLastError());
    return U3D_OK;
}

int u3d_gather_release(u3d_gather* g, void* stream)
{
    if (!g)
        return fail(U3D_ERR_INVALID, "u3d_gather_release: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    cudaStream_t st = g->ctx->pick(stream);
    k_gather_signal<<<1, kMaxRanks, 0, st>>>(g->peers, g->world, g->rank, (size_t)kMaxRanks * 4, g->waited);
    CU_CHECK(cudaGetLastError());
    return U3D_OK;
}

int u3d_gather_buffers(u3d_gather* g, void** counts_dev, void** offsets_dev, void** ids_dev, uint32_t* views_per_rank, uint64_t* ids_capacity_per_rank)
{
    if (!g)
        return fail(U3D_ERR_INVALID, "u3d_gather_buffers: NULL argument");
    if (counts_dev)
        *counts_dev = g->block + g->lay.countsOff();
    if (offsets_dev)
        *offsets_dev = g->block + g->lay.offsetsOff();
    if (ids_dev)
        *ids_dev = g->block + g->lay.idsOff();
    if (views_per_rank)
        *views_per_rank = g->lay.per;
    if (ids_capacity_per_rank)
        *ids_capacity_per_rank = g->lay.cap;
    return U3D_OK;
}

int u3d_gather_read(u3d_gather* g, uint32_t* counts, uint32_t* offsets, uint32_t* ids, void* stream)
{
    if (!g)
        return fail(U3D_ERR_INVALID, "u3d_gather_read: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    cudaStream_t st = g->ctx->pick(stream);
    if (counts)
        CU_CHECK(cudaMemcpyAsync(counts, g->block + g->lay.countsOff(), (size_t)g->world * g->lay.per * 8, cudaMemcpyDeviceToHost, st));
    if (offsets)
        CU_CHECK(cudaMemcpyAsync(offsets, g->block + g->lay.offsetsOff(), (size_t)g->world * (g->lay.per + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (ids)
        CU_CHECK(cudaMemcpyAsync(ids, g->block + g->lay.idsOff(), (size_t)g->world * g->lay.cap * 4, cudaMemcpyDeviceToHost, st));
    return u3d_gather_status(g, stream);
}

int u3d_gather_status(u3d_gather* g, void* stream)
{
    if (!g)
        return fail(U3D_ERR_INVALID, "u3d_gather_status: NULL argument");
    CU_CHECK(cudaSetDevice(g->ctx->device));
    cudaStream_t st = g->ctx->pick(stream);
    uint32_t status = 0;
    CU_CHECK(cudaMemcpyAsync(&status, reinterpret_cast<uint32_t*>(g->block) + 2 * kMaxRanks, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (status)
    {
        // reported once: the word is cleared so that one bad epoch does not fail every later read of this object
        CU_CHECK(cudaMemsetAsync(reinterpret_cast<uint32_t*>(g->block) + 2 * kMaxRanks, 0, 4, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    if (status & kGatherTimeout)
        return fail(U3D_ERR_CUDA, "result exchange timed out waiting for a peer rank");
    if (status & kGatherOverflow)
        return fail(U3D_ERR_CAPACITY, "result exchange: ids_capacity_per_rank too small for this batch");
    return U3D_OK;
}

int u3d_occluders_set_draw_distances(u3d_occluders* o, uint32_t n, const float* draw_distance)
{
    if (!o || !draw_distance || n != o->n)
        return fail(U3D_ERR_INVALID, "u3d_occluders_set_draw_distances: bad argument");
    CU_CHECK(cudaSetDevice(o->ctx->device));
    CU_CHECK(o->drawDist.ensure(std::max<uint32_t>(n, 1)));
    CU_CHECK(cudaMemcpy(o->drawDist.p, draw_distance, (size_t)n * 4, cudaMemcpyHostToDevice));
    return U3D_OK;
}

int u3d_batch_set_occluder_policy(u3d_batch* b, int enable, float size_threshold, uint32_t max_triangles, const float* camera3, void* stream)
{
    if (!b || (enable && !camera3 && b->numViews))
        return fail(U3D_ERR_INVALID, "u3d_batch_set_occluder_policy: NULL argument");
    if (enable < 0 || enable > 2)
        return fail(U3D_ERR_INVALID, "u3d_batch_set_occluder_policy: enable must be 0, 1 or 2");
    if (enable && !b->hasBuffers)
        return fail(U3D_ERR_INVALID, "u3d_batch_set_occluder_policy: the batch has no occluders");
    b->rankOccluders = enable != 0;
    b->serialOccluders = enable == 2;
    if (!enable)
        return U3D_OK;
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    CU_CHECK(init_rank_kernel()); // per device context; cheap
    b->sizeThreshold = size_threshold;
    b->maxTriangles = max_triangles;
    const uint32_t V = b->numViews;
    b->camHost.resize(std::max<uint32_t>(V, 1));
    for (uint32_t v = 0; v < V; ++v)
        // halfViewSize, invOrthoSize = 1.0f / camera->GetOrthoSize() (View.cpp:2175), farClip
        b->camHost[v] = make_float4(camera3[3 * v], 1.0f / camera3[3 * v + 1], camera3[3 * v + 2], 0.0f);
    CU_CHECK(b->camParams.ensure(std::max<uint32_t>(b->vs.maxViews, 1)));
    if (V)
    {
        CU_CHECK(cudaMemcpyAsync(b->camParams.p, b->camHost.data(), (size_t)V * sizeof(float4), cudaMemcpyHostToDevice, st));
        CU_CHECK(cudaStreamSynchronize(st)); // camHost is reused by the next call
    }
    return U3D_OK;
}

int u3d_batch_read_active_occluders(u3d_batch* b, uint32_t* active, void* stream)
{
    if (!b || !active)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_active_occluders: NULL argument");
    if (!b->serialOccluders || !b->activeOccluders.n)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_active_occluders: only after a run with occluder policy 2");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    if (b->numViews)
    {
        CU_CHECK(cudaMemcpyAsync(active, b->activeOccluders.p, (size_t)b->numViews * 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    return U3D_OK;
}

int u3d_batch_read_occluder_selection(u3d_batch* b, uint32_t view, uint32_t* occluder_ids, uint32_t capacity, uint32_t* count, void* stream)
{
    if (!b || view >= b->numViews || !count || (capacity && !occluder_ids))
        return fail(U3D_ERR_INVALID, "u3d_batch_read_occluder_selection: bad argument");
    if (!b->hasBuffers)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_occluder_selection: the batch has no occluders");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    uint32_t n = 0;
    CU_CHECK(cudaMemcpyAsync(&n, b->vs.itemCount.p + view, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    n = std::min(n, b->vs.itemCap);
    std::vector<DrawItem> items(n);
    if (n)
    {
        CU_CHECK(cudaMemcpyAsync(items.data(), b->vs.items.p + (size_t)view * b->vs.itemCap, (size_t)n * sizeof(DrawItem), cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (i && items[i].model == items[i - 1].model)
            continue; // later <= 256-triangle slices of the same occluder (an occluder is submitted at most once per view)
        if (m < capacity)
            occluder_ids[m] = items[i].model;
        ++m;
    }
    *count = m;
    if (m > capacity)
        return fail(U3D_ERR_CAPACITY, "u3d_batch_read_occluder_selection: capacity too small");
    return U3D_OK;
}

int u3d_batch_read_z_ranges(u3d_batch* b, float* minmax, void* stream)
{
    if (!b || !minmax)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_z_ranges: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    if (b->numViews)
    {
        CU_CHECK(cudaMemcpyAsync(minmax, b->zRange.p, (size_t)b->numViews * 8, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    return U3D_OK;
}

int u3d_batch_read_view(u3d_batch* b, uint32_t view, uint32_t* ids, uint32_t capacity, uint32_t* count, void* stream)
{
    if (!b || view >= b->numViews)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_view: bad argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    uint32_t c[2];
    CU_CHECK(cudaMemcpyAsync(c, b->outCounts.p + 2 * (size_t)view, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    if (count)
        *count = c[1];
    const uint32_t n = std::min(std::min(c[1], capacity), b->outCap);
    if (n && ids)
    {
        CU_CHECK(cudaMemcpyAsync(ids, b->outIdx.p + (size_t)view * b->outCap, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    return U3D_OK;
}

int u3d_batch_read_depth(u3d_batch* b, uint32_t view, int32_t* out_depth, void* stream)
{
    if (!b || !out_depth || view >= b->numViews || !b->hasBuffers)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_depth: bad argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const size_t n = (size_t)b->vs.g.w * b->vs.g.h;
    CU_CHECK(cudaMemcpyAsync(out_depth, b->vs.depth.p + view * n, n * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return U3D_OK;
}

int u3d_batch_read_mip(u3d_batch* b, uint32_t view, int level, int32_t* out_minmax, void* stream)
{
    if (!b || !out_minmax || view >= b->numViews || !b->hasBuffers || level < 0 || level >= b->vs.g.nlev)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_mip: bad argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const BufGeom& g = b->vs.g;
    CU_CHECK(cudaMemcpyAsync(out_minmax, b->vs.mips.p + (size_t)view * g.mipTexels + g.moff[level], (size_t)g.mw[level] * g.mh[level] * 8,
        cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    return U3D_OK;
}

int u3d_batch_read_tri_stats(u3d_batch* b, uint32_t* stats3, void* stream)
{
    if (!b || !stats3)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_tri_stats: NULL argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    const uint32_t V = b->numViews;
    if (!b->hasBuffers)
    {
        std::memset(stats3, 0, (size_t)V * 12);
        return U3D_OK;
    }
    if (b->countPending && V)
    {
        // numTriangles_ counts every source triangle that passes clipping and culling, drawn into a visible place or not (OB.cpp:681-682).
        // The occluders the second pass left out never went through set-up; their triangles are counted here, on demand, by the same
        // kernels in count-only mode (the count is a statistic of the batched path: its budget is checked on submitted triangles).
        ViewSet& vs = b->vs;
        const u3d_occluders* oc = b->occ;
        RasterDev rd = vs.raster_dev(true, 1);
        rd.itemPass = b->itemPass.p;
        rd.pass = kPassFarHidden;
        rd.countOnly = 1;
        CU_CHECK(cudaMemsetAsync(vs.clipCount.p, 0, 4, st));
        launch_setup(vs.g, vs.tg, vs.views.p, (int)V, vs.items.p, vs.itemCap, oc->totalItems, vs.itemCount.p, oc->model12.p, oc->meshTable.p, rd, st);
        CU_CHECK(cudaGetLastError());
        b->countPending = false;
    }
    std::vector<uint32_t> sub(V), drawn(V), setup(V);
    CU_CHECK(cudaMemcpyAsync(sub.data(), b->vs.submitted.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(drawn.data(), b->vs.drawnSources.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(setup.data(), b->vs.triCount.p, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    for (uint32_t v = 0; v < V; ++v)
    {
        stats3[3 * v] = sub[v];
        stats3[3 * v + 1] = sub[v] + drawn[v];
        stats3[3 * v + 2] = setup[v];
    }
    return U3D_OK;
}

int u3d_batch_read_setup_records(u3d_batch* b, uint32_t view, int32_t* out16, uint32_t capacity, uint32_t* count, void* stream)
{
    if (!b || view >= b->numViews || !b->hasBuffers)
        return fail(U3D_ERR_INVALID, "u3d_batch_read_setup_records: bad argument");
    CU_CHECK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->pick(stream);
    uint32_t n = 0, cap = 0;
    uint64_t base = 0;
    CU_CHECK(cudaMemcpyAsync(&n, b->vs.triCount.p + view, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(&cap, b->vs.triCap.p + view, 4, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(&base, b->vs.triBase.p + view, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    n = std::min(n, cap);
    if (count)
        *count = n;
    const uint32_t m = std::min(n, capacity);
    if (m && out16)
    {
        CU_CHECK(cudaMemcpyAsync(out16, b->vs.pool.p + base, (size_t)m * sizeof(TriSetup), cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    return U3D_OK;
}

static int flags_to_error(uint32_t flags)
{
    if (flags & kFlagClipOverflow)
        return fail(U3D_ERR_CAPACITY, "clipper queue too small for this batch: recreate the batch with a larger tri_pool");
    if (flags & kFlagIrrOverflow)
        return fail(U3D_ERR_CAPACITY, "more than 2^20 irregular triangles in one batch (degenerate input)");
    if (flags & (kFlagTriOverflow | kFlagPoolOverflow))
        return fail(U3D_ERR_CAPACITY, "triangle pool too small for this batch: recreate the batch with a larger tri_pool");
    if (flags & kFlagOutOverflow)
        return fail(U3D_ERR_CAPACITY, "a view produced more results than out_capacity");
    if (flags & kFlagRankOverflow)
        return fail(U3D_ERR_CAPACITY, "a view has more than 4096 occluder candidates after the size threshold (ranked selection)");
    return U3D_OK;
}

int u3d_batch_cull_views(u3d_batch* b, const u3d_view* views, uint32_t n, int stage, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t capacity,
    void* stream)
{
    int rc = u3d_batch_set_views(b, views, n, stream);
    if (rc < 0)
        return rc;
    rc = u3d_batch_run(b, stage, stream);
    if (rc < 0)
        return rc;
    if (!counts && !offsets)
        return U3D_OK;
    cudaStream_t st = b->ctx->pick(stream);
    const int V = (int)n;
    if (!V)
    {
        if (offsets)
            offsets[0] = 0;
        return U3D_OK;
    }
    // One synchronisation: the header (flags | counts | packed offsets) and the packed ids leave in the same breath, the ids up to what
    // the previous call needed plus 1/8 (frames are coherent); only a batch that outgrew that pays a second copy.
    const size_t hdrWords = 1 + 2 * (size_t)V + (size_t)V + 1;
    CU_CHECK(b->hdr.ensure(1 + 3 * (size_t)b->vs.maxViews + 1));
    CU_CHECK(b->hostHdr.ensure(1 + 3 * (size_t)b->vs.maxViews + 1));
    CU_CHECK(b->packed.ensure((size_t)b->vs.maxViews * std::max<uint32_t>(b->outCap, 1)));
    k_pack_header<<<1, 1024, 0, st>>>(V, b->outCounts.p, b->outCap, b->packOffsets.p, b->vs.flags.p, b->hdr.p);
    const bool wantIds = offsets && ids && capacity;
    if (wantIds)
        k_pack_ids<<<V, 256, 0, st>>>(b->outCounts.p, b->outCap, b->packOffsets.p, b->outIdx.p, b->packed.p);
    CU_CHECK(cudaGetLastError());
    b->lastLaunches += wantIds ? 2 : 1;
    CU_CHECK(cudaMemcpyAsync(b->hostHdr.p, b->hdr.p, hdrWords * 4, cudaMemcpyDeviceToHost, st));
    const uint64_t packedCap = (uint64_t)V * std::max<uint32_t>(b->outCap, 1);
    const uint64_t first = wantIds ? std::min<uint64_t>(std::min<uint64_t>(capacity, packedCap), b->idsGuess) : 0;
    if (first)
        CU_CHECK(cudaMemcpyAsync(ids, b->packed.p, first * 4, cudaMemcpyDeviceToHost, st));
    if (g_blockingWait)
    {
        // more calling threads than host cores (8 ranks x 12 slots on a 16-core box): a spinning wait takes the cores the other threads
        // need to launch their batches; a blocking event gives them back
        if (!b->haveDoneEvent)
        {
            CU_CHECK(cudaEventCreateWithFlags(&b->done, cudaEventBlockingSync | cudaEventDisableTiming));
            b->haveDoneEvent = true;
        }
        CU_CHECK(cudaEventRecord(b->done, st));
        CU_CHECK(cudaEventSynchronize(b->done));
    }
    else
        CU_CHECK(cudaStreamSynchronize(st));
    rc = flags_to_error(b->hostHdr.p[0]);
    if (rc < 0)
        return rc;
    if (counts)
        std::memcpy(counts, b->hostHdr.p + 1, (size_t)V * 8);
    if (offsets)
    {
        std::memcpy(offsets, b->hostHdr.p + 1 + 2 * (size_t)V, ((size_t)V + 1) * 4);
        const uint64_t total = offsets[V];
        if (total > capacity)
            return fail(U3D_ERR_CAPACITY, "u3d_batch_cull_views: ids capacity too small");
        if (wantIds && total > first)
        {
            CU_CHECK(cudaMemcpyAsync(ids + first, b->packed.p + first, (total - first) * 4, cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaStreamSynchronize(st));
        }
        b->idsGuess = total + total / 8 + 1024;
    }
    return U3D_OK;
}

int u3d_batch_device_results(u3d_batch* b, void** counts_dev, void** ids_dev, uint32_t* out_capacity)
{
    if (!b)
        return fail(U3D_ERR_INVALID, "NULL argument");
    if (counts_dev) *counts_dev = b->outCounts.p;
    if (ids_dev) *ids_dev = b->outIdx.p;
    if (out_capacity) *out_capacity = b->outCap;
    return U3D_OK;
}

uint32_t u3d_batch_last_launches(const u3d_batch* b) { return b ? b->lastLaunches : 0; }

int u3d_batch_set_pipelined(u3d_batch* b, int on)
{
    if (!b)
        return fail(U3D_ERR_INVALID, "u3d_batch_set_pipelined: NULL argument");
    b->pipelined = on != 0;
    return U3D_OK;
}

} // extern "C"
