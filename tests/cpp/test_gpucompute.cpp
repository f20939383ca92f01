// C++ test of the GPUCompute host module (urho3d_b200/GPUCompute/GPUCompute.h), written the way engine code uses the reference
// classes (View::DrawOccluders / View::GetDrawables, View.cpp:2231-2274, 873-878): SetSize, SetView, Clear, AddTriangles, DrawTriangles,
// BuildDepthHierarchy, Octree::GetDrawables(query), IsVisible.  Every result is compared bit-for-bit with the C oracle
// (oracle/u3d_oracle.c, linked as the checker only).  Built and run by tests/test_gpu_cpp_host.py.
#include "../../urho3d_b200/GPUCompute/GPUCompute.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C"
{
struct OrcOB;
typedef struct
{
    int m;
    const int* parent;
    const float* box6;
    const int* first;
    const int* count;
    int n;
    const float* aabb6;
    const unsigned char* flags;
    const unsigned* view_mask;
    const unsigned char* occludee;
    const unsigned char* cast_shadows;
} OrcTree;
OrcOB* orc_ob_create(void);
void orc_ob_destroy(OrcOB*);
int orc_ob_set_size(OrcOB*, int, int);
void orc_ob_set_view(OrcOB*, const float*, const float*, int);
void orc_ob_set_max_triangles(OrcOB*, unsigned);
void orc_ob_set_cull_mode(OrcOB*, int);
void orc_ob_clear(OrcOB*);
int orc_ob_draw_batch(OrcOB*, const float*, const void*, unsigned, const void*, unsigned, unsigned, unsigned);
void orc_ob_build_hierarchy(OrcOB*);
int orc_ob_is_visible(const OrcOB*, const float*);
void orc_ob_read_depth(const OrcOB*, int*);
void orc_ob_read_mip(const OrcOB*, int, int*);
unsigned orc_ob_num_triangles(const OrcOB*);
int orc_query(const OrcTree*, int, const float*, const OrcOB*, unsigned, unsigned, int*);
int orc_check_visibility(const OrcTree*, const OrcOB*, const int*, int, int*);
int orc_raycast(const OrcTree*, const float*, unsigned, unsigned, int, int*, float*);
}

using namespace Urho3D::GPUCompute;

static unsigned g_seed = 12345;
static float Rand(float lo, float hi)
{
    g_seed = g_seed * 1664525u + 1013904223u;
    return lo + (hi - lo) * (float)((g_seed >> 8) & 0xFFFFFF) / 16777216.0f;
}

#define EXPECT(cond)                                                                                                     \
    do                                                                                                                   \
    {                                                                                                                    \
        if (!(cond))                                                                                                     \
        {                                                                                                                \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);                                                  \
            return 1;                                                                                                    \
        }                                                                                                                \
    } while (0)

int main()
{
    const int W = 256, H = 160;
    GpuContext ctx(0);

    // --- a scene: boxes on a ground slab, wall occluders (unit cube, 36 indices, stride 24 to exercise the stride) ---
    const int N = 5000, K = 96;
    std::vector<float> boxes(6 * N);
    std::vector<unsigned char> flags(N, 1), occludee(N, 1), shadows(N, 0);
    std::vector<unsigned> masks(N, 0xFFFFFFFFu);
    for (int i = 0; i < N; ++i)
    {
        const float cx = Rand(-80, 80), cz = Rand(-80, 80), hx = Rand(0.3f, 1.2f), hy = Rand(0.3f, 1.2f), hz = Rand(0.3f, 1.2f);
        float* b = &boxes[6 * i];
        b[0] = cx - hx; b[1] = 0.0f; b[2] = cz - hz; b[3] = cx + hx; b[4] = 2 * hy; b[5] = cz + hz;
        occludee[i] = (i % 97) != 0;
    }
    static const float cubeP[8][3] = {{-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {.5f, .5f, -.5f}, {-.5f, .5f, -.5f}, {-.5f, -.5f, .5f}, {.5f, -.5f, .5f},
        {.5f, .5f, .5f}, {-.5f, .5f, .5f}};
    static const unsigned short cubeI[36] = {0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 3, 6, 2, 3, 7, 6, 0, 7, 3, 0, 4, 7, 1, 2, 6, 1, 6, 5};
    std::vector<float> vtx(8 * 6, 7.0f);
    for (int i = 0; i < 8; ++i)
        for (int a = 0; a < 3; ++a)
            vtx[6 * i + a] = cubeP[i][a];
    std::vector<float> models(12 * K, 0.0f);
    for (int k = 0; k < K; ++k)
    {
        float* m = &models[12 * k];
        m[0] = Rand(6, 18); m[5] = 8.0f; m[10] = 1.0f;
        m[3] = Rand(-80, 80); m[7] = 4.0f; m[11] = Rand(-80, 80);
    }

    GpuOctree octree(ctx, (const float[3]){-1000, -1000, -1000}, (const float[3]){1000, 1000, 1000}, 8);
    octree.InsertDrawables(N, boxes.data(), flags.data(), masks.data(), occludee.data(), shadows.data());
    octree.Sync();

    // flatten for the oracle
    uint32_t M = 0, S = 0;
    u3d_octree_counts(octree.Tree(), &M, &S);
    std::vector<int> parent(M), level(M), first(M), count(M), order(S);
    std::vector<float> cbox(6 * M);
    u3d_octree_flatten(octree.Tree(), parent.data(), level.data(), cbox.data(), first.data(), count.data(), order.data());
    std::vector<float> dBoxes(6 * S);
    std::vector<unsigned char> dFlags(S), dOcc(S), dSh(S);
    std::vector<unsigned> dMask(S);
    for (uint32_t s = 0; s < S; ++s)
    {
        const int id = order[s];
        for (int a = 0; a < 6; ++a)
            dBoxes[6 * s + a] = boxes[6 * id + a];
        dFlags[s] = flags[id]; dOcc[s] = occludee[id]; dSh[s] = shadows[id]; dMask[s] = masks[id];
    }
    OrcTree otree = {(int)M, parent.data(), cbox.data(), first.data(), count.data(), (int)S, dBoxes.data(), dFlags.data(), dMask.data(), dOcc.data(), dSh.data()};

    GpuOcclusionBuffer buffer(ctx);
    float probe[6] = {0, 0, 0, 1, 1, 1};
    EXPECT(buffer.IsVisible(probe));      // no buffer yet: true, OcclusionBuffer.cpp:338-339
    EXPECT(!buffer.SetSize(100, 64, false));
    EXPECT(buffer.SetSize(W, H - 1, false));
    EXPECT(buffer.GetHeight() == H);
    OrcOB* ob = orc_ob_create();
    orc_ob_set_size(ob, W, H);

    for (int viewNo = 0; viewNo < 4; ++viewNo)
    {
        // camera at (px, 2, pz) looking down +z: view = inverse translation, D3D-style projection (Camera.cpp:648-691)
        const float px = Rand(-40, 40), pz = Rand(-90, -30), nearClip = 0.1f, farClip = 300.0f, fov = 45.0f, aspect = (float)W / H;
        float view[12] = {1, 0, 0, -px, 0, 1, 0, -2.0f, 0, 0, 1, -pz};
        const float h = 1.0f / std::tan(fov * 3.14159265f / 360.0f), w = h / aspect, q = farClip / (farClip - nearClip);
        float proj[16] = {w, 0, 0, 0, 0, h, 0, 0, 0, 0, q, -q * nearClip, 0, 0, 1, 0};
        const float t = std::tan(fov * 3.14159265f / 360.0f), tx = t * aspect;
        const float il = 1.0f / std::sqrt(1 + tx * tx), iu = 1.0f / std::sqrt(1 + t * t);
        float planes[24] = {0, 0, 1, -(pz + nearClip),
            il, 0, tx * il, -(il * px + tx * il * pz), -il, 0, tx * il, -(-il * px + tx * il * pz),
            0, -iu, t * iu, -(-iu * 2.0f + t * iu * pz), 0, iu, t * iu, -(iu * 2.0f + t * iu * pz),
            0, 0, -1, pz + farClip};

        // View::DrawOccluders, threaded branch (View.cpp:2258-2273)
        buffer.SetView(view, proj, nearClip, farClip, false);
        buffer.SetMaxTriangles(1u << 30);
        buffer.Clear();
        orc_ob_set_view(ob, view, proj, 0);
        orc_ob_set_max_triangles(ob, 1u << 30);
        orc_ob_clear(ob);
        for (int k = 0; k < K; ++k)
        {
            buffer.SetCullMode(GPU_CULL_CCW);
            orc_ob_set_cull_mode(ob, 1);
            const bool a = buffer.AddTriangles(&models[12 * k], vtx.data(), 24, cubeI, sizeof(unsigned short), 0, 36);
            const bool b = orc_ob_draw_batch(ob, &models[12 * k], vtx.data(), 24, cubeI, 2, 0, 36) != 0;
            EXPECT(a == b);
        }
        buffer.DrawTriangles();
        buffer.BuildDepthHierarchy();
        orc_ob_build_hierarchy(ob);
        EXPECT(buffer.GetNumTriangles() == orc_ob_num_triangles(ob));

        std::vector<int> want((size_t)W * H);
        orc_ob_read_depth(ob, want.data());
        const int* got = buffer.GetBuffer();
        EXPECT(got != nullptr);
        EXPECT(std::memcmp(got, want.data(), want.size() * sizeof(int)) == 0);
        for (int lv = 0; lv < buffer.GetNumLevels(); ++lv)
        {
            std::vector<int> mip;
            int mw = 0, mh = 0;
            EXPECT(buffer.GetMip(lv, mip, mw, mh));
            std::vector<int> wantMip(mip.size());
            orc_ob_read_mip(ob, lv, wantMip.data());
            EXPECT(mip == wantMip);
        }

        // Octree::GetDrawables(OccludedFrustumOctreeQuery) (View.cpp:873-878), then the CheckVisibilityWork predicate (View.cpp:177)
        std::vector<unsigned> result;
        GpuOccludedFrustumOctreeQuery query(result, planes, &buffer, 0xFF, 0xFFFFFFFFu);
        octree.GetDrawables(query);
        std::vector<int> cand(S), vis(S);
        const int nc = orc_query(&otree, 1, planes, ob, 0xFF, 0xFFFFFFFFu, cand.data());
        EXPECT((int)result.size() == nc);
        for (int i = 0; i < nc; ++i)
            EXPECT((int)result[i] == order[cand[i]]);
        unsigned querySize = 0;
        octree.GetVisibleDrawables(query, &querySize);
        const int nv = orc_check_visibility(&otree, ob, cand.data(), nc, vis.data());
        EXPECT((int)querySize == nc && (int)result.size() == nv);
        for (int i = 0; i < nv; ++i)
            EXPECT((int)result[i] == order[vis[i]]);
        // plain frustum query and per-box IsVisible
        GpuFrustumOctreeQuery fq(result, planes);
        octree.GetDrawables(fq);
        const int nf = orc_query(&otree, 0, planes, nullptr, 0xFF, 0xFFFFFFFFu, cand.data());
        EXPECT((int)result.size() == nf);
        for (int i = 0; i < 50; ++i)
            EXPECT(buffer.IsVisible(&boxes[6 * (i * 37 % N)]) == (orc_ob_is_visible(ob, &boxes[6 * (i * 37 % N)]) != 0));
        std::printf("view %d: %u triangles, query %d, visible %d, frustum-only %d\n", viewNo, buffer.GetNumTriangles(), nc, nv, nf);
    }

    // Octree::Raycast / RaycastSingle with a RAY_AABB-level RayOctreeQuery (Octree.cpp:501-548)
    {
        std::vector<GpuRayQueryResult> hits;
        const float origin[3] = {-90.0f, 1.0f, -85.0f};
        const float direction[3] = {0.70710678f, 0.0f, 0.70710678f};
        GpuRayOctreeQuery rq(hits, origin, direction, 400.0f);
        octree.Raycast(rq);
        const float ray7[7] = {origin[0], origin[1], origin[2], direction[0], direction[1], direction[2], 400.0f};
        std::vector<int> pos(S);
        std::vector<float> dist(S);
        const int nr = orc_raycast(&otree, ray7, 0xFF, 0xFFFFFFFFu, 0, pos.data(), dist.data());
        EXPECT((int)hits.size() == nr && nr > 0);
        for (int i = 0; i < nr && i < (int)hits.size(); ++i)
            EXPECT((int)hits[i].drawable == order[pos[i]] && std::memcmp(&hits[i].distance, &dist[i], 4) == 0);
        octree.RaycastSingle(rq);
        const int ns = orc_raycast(&otree, ray7, 0xFF, 0xFFFFFFFFu, 1, pos.data(), dist.data());
        EXPECT((int)hits.size() == ns && ns == 1 && (int)hits[0].drawable == order[pos[0]]);
        std::printf("raycast: %d hits, closest drawable %u at %f\n", nr, hits.empty() ? 0u : hits[0].drawable, hits.empty() ? 0.0 : (double)hits[0].distance);
    }

    // a drawable moves: host mirror re-flattened on the next query
    float moved[6] = {500, 0, 500, 501, 1, 501};
    octree.UpdateDrawable(3, moved);
    std::vector<unsigned> result;
    const float all[24] = {0, 0, 1, 2000, 1, 0, 0, 2000, -1, 0, 0, 2000, 0, -1, 0, 2000, 0, 1, 0, 2000, 0, 0, -1, 2000};
    GpuFrustumOctreeQuery everything(result, all);
    octree.GetDrawables(everything);
    EXPECT((int)result.size() == N);
    // The batched entry point across structural changes of the octree (u3d_batch_set_scene through GpuViewBatch): GpuOctree::Sync replaces
    // the device scene after an insert / a removal / a move across octants; the batch must follow it.
    {
        GpuViewBatch batch(ctx, octree, nullptr, 2, 2, 4, N + 64);
        u3d_view views[2];
        std::memset(views, 0, sizeof(views));
        for (int k = 0; k < 2; ++k)
        {
            views[k].view[0] = views[k].view[5] = views[k].view[10] = 1.0f;
            views[k].projection[0] = views[k].projection[5] = views[k].projection[10] = views[k].projection[15] = 1.0f;
            std::memcpy(views[k].planes, all, sizeof(all));
            views[k].drawable_flags = 0xFF;
            views[k].view_mask = 0xFFFFFFFFu;
            views[k].query_mode = U3D_QUERY_FRUSTUM;
        }
        views[1].planes[3] = -40.0f; // near plane z >= 40: a part of the scene only
        std::vector<unsigned> ids;
        EXPECT(batch.CullViews(views, 2, ids, false));
        EXPECT(batch.Offsets()[1] == (unsigned)N);
        const unsigned partBefore = batch.Offsets()[2] - batch.Offsets()[1];
        // insert, move across octants (twice: the last box wins), remove
        float extra[6] = {10, 0, 60, 11, 1, 61};
        const unsigned newId = octree.InsertDrawable(extra, 1, 0xFFFFFFFFu, true, false);
        float far1[6] = {-70, 0, -70, -69, 1, -69}, far2[6] = {70, 0, 70, 71, 1, 71};
        octree.UpdateDrawable(5, far1);
        octree.UpdateDrawable(5, far2);
        octree.RemoveDrawable(7);
        EXPECT(batch.CullViews(views, 2, ids, false));
        EXPECT(batch.Offsets()[1] == (unsigned)N); // + 1 inserted - 1 removed
        std::vector<unsigned> single;
        GpuFrustumOctreeQuery q0(single, views[0].planes), q1(single, views[1].planes);
        octree.GetDrawables(q0);
        EXPECT(single.size() == batch.Offsets()[1] && std::equal(single.begin(), single.end(), ids.begin()));
        octree.GetDrawables(q1);
        EXPECT(single.size() == batch.Offsets()[2] - batch.Offsets()[1] && std::equal(single.begin(), single.end(), ids.begin() + batch.Offsets()[1]));
        bool sawNew = false, saw5 = false, saw7 = false;
        for (unsigned i = batch.Offsets()[1]; i < batch.Offsets()[2]; ++i)
        {
            sawNew |= ids[i] == newId;
            saw5 |= ids[i] == 5;
            saw7 |= ids[i] == 7;
        }
        EXPECT(sawNew && saw5 && !saw7);
        std::printf("batch across structural changes: %u -> %u drawables behind z = 40\n", partBefore, batch.Offsets()[2] - batch.Offsets()[1]);
    }
    orc_ob_destroy(ob);
    std::printf("OK\n");
    return 0;
}
