/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, scalar, one thread) of the reference's per-frame visibility
 * path.  It is the checker for the CUDA path; nothing in urho3d_b200/ may link, load or call it.
 *
 * PARITY PINNING: the reference ships no golden vectors for this path.  This restatement is
 * pinned differentially against the UNMODIFIED reference sources compiled into oracle/_ref/libu3dref.so
 * (oracle/build_ref.sh) by tests/test_oracle_vs_reference.py, and against tests/golden/ fixtures generated from that
 * library by tests/golden/make_golden.py.
 *
 * Float mode: compile with -ffp-contract=off (no FMA contraction), no -ffast-math: SURVEY.md hard part 1.
 * "OB.cpp" = Source/Urho3D/Graphics/OcclusionBuffer.cpp of the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define Z_SCALE 16777216.0f    /* OCCLUSION_Z_SCALE, OcclusionBuffer.h:86 */
#define X_SCALE 65536.0f       /* OCCLUSION_X_SCALE, OcclusionBuffer.h:85 */
#define MIN_MIP 8              /* OCCLUSION_MIN_SIZE, OcclusionBuffer.h:83 */
#define FIXED_BIAS 16          /* OCCLUSION_FIXED_BIAS, OcclusionBuffer.h:87 */
#define REL_BIAS 0.00001f      /* OCCLUSION_RELATIVE_BIAS, OcclusionBuffer.h:88 */
#define MAX_LEVELS 32

enum { CULL_NONE = 0, CULL_CCW = 1, CULL_CW = 2 };
enum { OUTSIDE = 0, INTERSECTS = 1, INSIDE = 2 }; /* MathDefs.h Intersection */

typedef struct { float x, y, z, w; } V4;
typedef struct { float x, y, z; } V3;

typedef struct OrcOB
{
    int w, h;
    int nlev;
    int* base;                 /* w*(h+2)+2 ints, OB.cpp:89 */
    int* depth;                /* base + w + 1, OB.cpp:90 */
    int* mip[MAX_LEVELS];      /* interleaved (min,max) */
    int mw[MAX_LEVELS], mh[MAX_LEVELS];
    float vp[16];              /* viewProj_ row-major */
    float sx, sy, ox, oy;
    int cull, reverse;
    unsigned num_tris, max_tris;
    int dirty;
    unsigned long long px_tested, px_written; /* statistics only */
    unsigned long long ref_undefined;         /* writes the reference would have made outside its padded allocation (undefined behaviour there) */
} OrcOB;

/* x86 cvttss2si/cvttsd2si semantics made explicit: out-of-range and NaN give INT_MIN (SURVEY hard part 2). */
static int cvtt(double v)
{
    if (!(v >= -2147483648.0 && v < 2147483648.0))
        return INT32_MIN;
    return (int)v;
}
static int trunc_f(float v) { return cvtt((double)v); }
static int round_to_int(float v) { return cvtt(round((double)v)); } /* MathDefs.h:238 */
static int wrap_add(int a, int b) { return (int)((uint32_t)a + (uint32_t)b); }

OrcOB* orc_ob_create(void)
{
    OrcOB* ob = (OrcOB*)calloc(1, sizeof(OrcOB));
    ob->cull = CULL_CCW;
    ob->max_tris = 5000; /* OCCLUSION_DEFAULT_MAX_TRIANGLES, OcclusionBuffer.h:84 */
    ob->dirty = 1;
    return ob;
}

static void free_buffers(OrcOB* ob)
{
    free(ob->base);
    ob->base = NULL;
    for (int i = 0; i < ob->nlev; ++i)
        free(ob->mip[i]);
    ob->nlev = 0;
}

void orc_ob_destroy(OrcOB* ob)
{
    free_buffers(ob);
    free(ob);
}

static void calc_viewport(OrcOB* ob) /* OB.cpp:578-587 */
{
    ob->sx = 0.5f * ob->w;
    ob->sy = -0.5f * ob->h;
    ob->ox = 0.5f * ob->w + 0.5f;
    ob->oy = 0.5f * ob->h + 0.5f;
}

/* OB.cpp:61-113.  Returns 1 on success like the reference's bool. */
int orc_ob_set_size(OrcOB* ob, int w, int h)
{
    if (h & 1)
        ++h;
    if (w == ob->w && h == ob->h)
        return 1;
    if (w <= 0 || h <= 0)
        return 0;
    if (w & (w - 1))
        return 0;
    free_buffers(ob);
    ob->w = w;
    ob->h = h;
    ob->base = (int*)malloc(sizeof(int) * ((size_t)w * (h + 2) + 2));
    ob->depth = ob->base + w + 1;
    for (;;)
    {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        ob->mw[ob->nlev] = w;
        ob->mh[ob->nlev] = h;
        ob->mip[ob->nlev] = (int*)malloc(sizeof(int) * 2 * (size_t)w * h);
        ++ob->nlev;
        if (w <= MIN_MIP && h <= MIN_MIP)
            break;
    }
    calc_viewport(ob);
    return 1;
}

int orc_ob_width(const OrcOB* ob) { return ob->w; }
int orc_ob_height(const OrcOB* ob) { return ob->h; }
int orc_ob_num_levels(const OrcOB* ob) { return ob->nlev; }
unsigned orc_ob_num_triangles(const OrcOB* ob) { return ob->num_tris; }
int orc_ob_cull_mode(const OrcOB* ob) { return ob->cull; }
void orc_ob_pixel_stats(const OrcOB* ob, unsigned long long* tested, unsigned long long* written) { *tested = ob->px_tested; *written = ob->px_written; }
/* Since creation: how often the geometry drawn so far would have made the reference write outside its allocation (one row + one int of
 * padding each side of the plane, OB.cpp:87-90) -- counted pixel by pixel for the rows walked here, once per half-triangle whose row range
 * had to be cut short.  Non-zero = the reference's own result for that geometry is undefined (it usually corrupts its heap). */
unsigned long long orc_ob_ref_undefined(const OrcOB* ob) { return ob->ref_undefined; }

/* Matrix4 * Matrix3x4, Matrix4.cpp:43-63 (a = 4x4 row-major, b = 3x4 row-major). */
static void mul44_34(const float* a, const float* b, float* o)
{
    for (int r = 0; r < 4; ++r)
    {
        const float* ar = a + 4 * r;
        for (int c = 0; c < 3; ++c)
            o[4 * r + c] = ar[0] * b[c] + ar[1] * b[4 + c] + ar[2] * b[8 + c];
        o[4 * r + 3] = ar[0] * b[3] + ar[1] * b[7] + ar[2] * b[11] + ar[3];
    }
}

/* OB.cpp:115-127 with the Camera getters replaced by their values. */
void orc_ob_set_view(OrcOB* ob, const float* view12, const float* proj16, int reverse_culling)
{
    mul44_34(proj16, view12, ob->vp);
    ob->reverse = reverse_culling != 0;
    calc_viewport(ob);
}

void orc_ob_get_viewproj(const OrcOB* ob, float* out16) { memcpy(out16, ob->vp, sizeof(ob->vp)); }
void orc_ob_set_max_triangles(OrcOB* ob, unsigned n) { ob->max_tris = n; }

void orc_ob_set_cull_mode(OrcOB* ob, int mode) /* OB.cpp:134-144 */
{
    if (ob->reverse)
    {
        if (mode == CULL_CW)
            mode = CULL_CCW;
        else if (mode == CULL_CCW)
            mode = CULL_CW;
    }
    ob->cull = mode;
}

void orc_ob_reset(OrcOB* ob) { ob->num_tris = 0; }

void orc_ob_clear(OrcOB* ob) /* OB.cpp:152-162, 1028-1039 */
{
    ob->num_tris = 0;
    if (ob->base)
    {
        int fill = (int)Z_SCALE;
        size_t n = (size_t)ob->w * ob->h;
        for (size_t i = 0; i < n; ++i)
            ob->depth[i] = fill;
    }
    ob->dirty = 1;
}

static V4 xform(const float* m, const float* p) /* ModelTransform, OB.cpp:543-551 */
{
    V4 r;
    r.x = m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3];
    r.y = m[4] * p[0] + m[5] * p[1] + m[6] * p[2] + m[7];
    r.z = m[8] * p[0] + m[9] * p[1] + m[10] * p[2] + m[11];
    r.w = m[12] * p[0] + m[13] * p[1] + m[14] * p[2] + m[15];
    return r;
}

static V3 to_screen(const OrcOB* ob, V4 v) /* ViewportTransform, OB.cpp:553-561 */
{
    float inv_w = 1.0f / v.w;
    V3 r;
    r.x = inv_w * v.x * ob->sx + ob->ox;
    r.y = inv_w * v.y * ob->sy + ob->oy;
    r.z = inv_w * v.z * Z_SCALE;
    return r;
}

static V4 clip_lerp(V4 a, V4 b, float da, float db) /* ClipEdge, OB.cpp:563-567 */
{
    float t = da / (da - db);
    V4 r;
    r.x = a.x + t * (b.x - a.x);
    r.y = a.y + t * (b.y - a.y);
    r.z = a.z + t * (b.z - a.z);
    r.w = a.w + t * (b.w - a.w);
    return r;
}

/* Store an int into the depth plane by linear offset from data_.  The reference writes through a raw pointer into
 * an allocation padded by one row + 1 int each side (OB.cpp:87-90); anything beyond that is undefined behaviour there.
 * Writes outside [0, w*h) never reach the visible plane, so they are dropped here. */
static void depth_min(OrcOB* ob, int64_t off, int z)
{
    if (off < 0 || off >= (int64_t)ob->w * ob->h)
    {
        if (off < -((int64_t)ob->w + 1) || off >= (int64_t)ob->w * ob->h + ob->w + 1)
            ++ob->ref_undefined;
        return;
    }
    ++ob->px_tested;
    if (z < ob->depth[off])
    {
        ob->depth[off] = z;
        ++ob->px_written;
    }
}

typedef struct { int x, xstep, z, zstep; } EdgeI;

/* Edge constructor, OB.cpp:792-803 */
static EdgeI make_edge(float dzdx, float dzdy, V3 top, V3 bot, int top_y)
{
    EdgeI e;
    float height = bot.y - top.y;
    float slope = (height != 0.0f) ? (bot.x - top.x) / height : 0.0f;
    float ypre = (float)(top_y + 1) - top.y;
    float xpre = slope * ypre;
    e.x = round_to_int((xpre + top.x) * X_SCALE);
    e.xstep = round_to_int(slope * X_SCALE);
    e.z = round_to_int(top.z + xpre * dzdx + ypre * dzdy);
    e.zstep = round_to_int(slope * dzdx + dzdy);
    return e;
}

/* One half of the triangle: rows [y0, y1), span from `left` to `right`, depth walks along `left`. OB.cpp:897-1001 */
static void edge_skip(EdgeI* e, int64_t rows)
{
    /* `rows` incremental steps at once; additions wrap, so a 32-bit modular multiply gives the same state */
    e->x = (int)((uint32_t)e->x + (uint32_t)rows * (uint32_t)e->xstep);
    e->z = (int)((uint32_t)e->z + (uint32_t)rows * (uint32_t)e->zstep);
}

static void fill_half(OrcOB* ob, int y0, int y1, EdgeI* left, EdgeI* right, int dzdx_i)
{
    /* Rows whose linear span could touch [0,w*h): |x>>16| <= 32768, so rows far outside cannot change the visible plane.
     * They are skipped (float garbage can make the row range billions long); the edges are advanced as if stepped. */
    int64_t guard = 32768 / ob->w + 2;
    int64_t lo = y0, hi = y1;
    if (hi <= lo)
        return;
    if (lo < -guard)
        lo = -guard;
    if (hi > (int64_t)ob->h + guard)
        hi = (int64_t)ob->h + guard;
    if (hi < lo)
        hi = lo = (y1 < lo ? y1 : lo);
    if (lo > y1)
        lo = hi = y1;
    if (lo != y0 || hi != y1)
        ++ob->ref_undefined; /* rows far outside the plane: the reference walks (and, unless their spans are empty, writes) them */
    edge_skip(left, lo - y0);
    edge_skip(right, lo - y0);
    for (int64_t y = lo; y < hi; ++y)
    {
        int z = left->z;
        int64_t xs = left->x >> 16;
        int64_t xe = right->x >> 16;
        int64_t row = y * ob->w;
        for (int64_t x = xs; x < xe; ++x)
        {
            depth_min(ob, row + x, z);
            z = wrap_add(z, dzdx_i);
        }
        left->x = wrap_add(left->x, left->xstep);
        left->z = wrap_add(left->z, left->zstep);
        right->x = wrap_add(right->x, right->xstep);
    }
    edge_skip(left, (int64_t)y1 - hi);
    edge_skip(right, (int64_t)y1 - hi);
}

/* DrawTriangle2D, OB.cpp:815-1002 */
static void raster_tri(OrcOB* ob, const V3* v, int clockwise)
{
    int top, mid, bot, mid_right;
    if (v[0].y < v[1].y)
    {
        if (v[2].y < v[0].y) { top = 2; mid = 0; bot = 1; mid_right = 1; }
        else
        {
            top = 0;
            if (v[1].y < v[2].y) { mid = 1; bot = 2; mid_right = 1; }
            else { mid = 2; bot = 1; mid_right = 0; }
        }
    }
    else
    {
        if (v[2].y < v[1].y) { top = 2; mid = 1; bot = 0; mid_right = 0; }
        else
        {
            top = 1;
            if (v[0].y < v[2].y) { mid = 0; bot = 2; mid_right = 0; }
            else { mid = 2; bot = 0; mid_right = 1; }
        }
    }
    int ty = trunc_f(v[top].y), my = trunc_f(v[mid].y), by = trunc_f(v[bot].y);
    if (ty == by)
        return;
    if (!clockwise)
        mid_right = !mid_right;

    /* Gradients, OB.cpp:762-778 */
    float inv_dx = 1.0f / (((v[1].x - v[2].x) * (v[0].y - v[2].y)) - ((v[0].x - v[2].x) * (v[1].y - v[2].y)));
    float inv_dy = -inv_dx;
    float dzdx = inv_dx * (((v[1].z - v[2].z) * (v[0].y - v[2].y)) - ((v[0].z - v[2].z) * (v[1].y - v[2].y)));
    float dzdy = inv_dy * (((v[1].z - v[2].z) * (v[0].x - v[2].x)) - ((v[0].z - v[2].z) * (v[1].x - v[2].x)));
    int dzdx_i = trunc_f(dzdx);

    EdgeI lng = make_edge(dzdx, dzdy, v[top], v[bot], ty);
    if (ty != my)
    {
        EdgeI e = make_edge(dzdx, dzdy, v[top], v[mid], ty);
        if (mid_right)
            fill_half(ob, ty, my, &lng, &e, dzdx_i);
        else
            fill_half(ob, ty, my, &e, &lng, dzdx_i);
    }
    if (my != by)
    {
        EdgeI e = make_edge(dzdx, dzdy, v[mid], v[bot], my);
        if (mid_right)
            fill_half(ob, my, by, &lng, &e, dzdx_i);
        else
            fill_half(ob, my, by, &e, &lng, dzdx_i);
    }
}

static int project_and_fill(OrcOB* ob, const V4* c)
{
    V3 p[3];
    p[0] = to_screen(ob, c[0]);
    p[1] = to_screen(ob, c[1]);
    p[2] = to_screen(ob, c[2]);
    /* SignedArea, OB.cpp:569-576 */
    float ax = p[0].x - p[1].x, ay = p[0].y - p[1].y, bx = p[2].x - p[1].x, by = p[2].y - p[1].y;
    int clockwise = (ax * by - ay * bx) < 0.0f;
    if (ob->cull == CULL_NONE || (ob->cull == CULL_CCW && clockwise) || (ob->cull == CULL_CW && !clockwise))
    {
        raster_tri(ob, p, clockwise);
        return 1;
    }
    return 0;
}

static float plane_dot(const float* pl, V4 v) { return pl[0] * v.x + pl[1] * v.y + pl[2] * v.z + pl[3] * v.w; }

/* ClipVertices, OB.cpp:685-753: every live triangle is cut by one plane; a triangle with one vertex behind spawns a second. */
static void clip_plane(const float* pl, V4* vv, unsigned char* live, unsigned* count)
{
    unsigned n = *count;
    for (unsigned i = 0; i < n; ++i)
    {
        if (!live[i])
            continue;
        V4* t = vv + 3 * i;
        float d0 = plane_dot(pl, t[0]), d1 = plane_dot(pl, t[1]), d2 = plane_dot(pl, t[2]);
        int b0 = d0 < 0.0f, b1 = d1 < 0.0f, b2 = d2 < 0.0f;
        if (b0 && b1 && b2)
            live[i] = 0;
        else if (b0 && b1)
        {
            t[0] = clip_lerp(t[0], t[2], d0, d2);
            t[1] = clip_lerp(t[1], t[2], d1, d2);
        }
        else if (b0 && b2)
        {
            t[0] = clip_lerp(t[0], t[1], d0, d1);
            t[2] = clip_lerp(t[2], t[1], d2, d1);
        }
        else if (b1 && b2)
        {
            t[1] = clip_lerp(t[1], t[0], d1, d0);
            t[2] = clip_lerp(t[2], t[0], d2, d0);
        }
        else if (b0)
        {
            V4* nt = vv + 3 * (*count);
            live[(*count)++] = 1;
            nt[0] = clip_lerp(t[0], t[2], d0, d2);
            nt[1] = t[0] = clip_lerp(t[0], t[1], d0, d1);
            nt[2] = t[2];
        }
        else if (b1)
        {
            V4* nt = vv + 3 * (*count);
            live[(*count)++] = 1;
            nt[1] = clip_lerp(t[1], t[0], d1, d0);
            nt[2] = t[1] = clip_lerp(t[1], t[2], d1, d2);
            nt[0] = t[0];
        }
        else if (b2)
        {
            V4* nt = vv + 3 * (*count);
            live[(*count)++] = 1;
            nt[2] = clip_lerp(t[2], t[1], d2, d1);
            nt[0] = t[2] = clip_lerp(t[2], t[0], d2, d0);
            nt[1] = t[1];
        }
    }
}

/* DrawTriangle, OB.cpp:589-683.  vv has room for 64 triangles; vv[0..2] hold the clip-space input. */
static void draw_tri(OrcOB* ob, V4* vv)
{
    unsigned any = 0, all = 0;
    for (int i = 0; i < 3; ++i)
    {
        unsigned m = 0;
        if (vv[i].x > vv[i].w) m |= 1;
        if (vv[i].x < -vv[i].w) m |= 2;
        if (vv[i].y > vv[i].w) m |= 4;
        if (vv[i].y < -vv[i].w) m |= 8;
        if (vv[i].z > vv[i].w) m |= 16;
        if (vv[i].z < 0.0f) m |= 32;
        any |= m;
        all = i ? (all & m) : m;
    }
    if (all)
        return;
    int drawn = 0;
    if (!any)
        drawn = project_and_fill(ob, vv);
    else
    {
        static const float planes[6][4] = {
            {-1, 0, 0, 1}, {1, 0, 0, 1}, {0, -1, 0, 1}, {0, 1, 0, 1}, {0, 0, -1, 1}, {0, 0, 1, 0}};
        unsigned char live[64];
        unsigned count = 1;
        live[0] = 1;
        for (int p = 0; p < 6; ++p)
            if (any & (1u << p))
                clip_plane(planes[p], vv, live, &count);
        for (unsigned i = 0; i < count; ++i)
            if (live[i])
                drawn |= project_and_fill(ob, vv + 3 * i);
    }
    if (drawn)
        ++ob->num_tris;
}

/* AddTriangles (OB.cpp:164-198) immediately followed by DrawBatch (OB.cpp:464-541) for that one batch.
 * idx == NULL -> non-indexed (start/count in vertices).  Returns AddTriangles' bool (budget not exceeded). */
int orc_ob_draw_batch(OrcOB* ob, const float* model12, const void* vtx, unsigned stride, const void* idx, unsigned idx_size,
    unsigned start, unsigned count)
{
    ob->num_tris += count / 3;
    int ok = ob->num_tris <= ob->max_tris;
    if (!ob->base)
        return ok;
    float mvp[16];
    mul44_34(ob->vp, model12, mvp);
    const unsigned char* vb = (const unsigned char*)vtx;
    V4 vv[64 * 3];
    if (!idx)
    {
        const unsigned char* src = vb + (size_t)start * stride;
        for (unsigned i = 0; i + 2 < count; i += 3)
        {
            for (int k = 0; k < 3; ++k)
                vv[k] = xform(mvp, (const float*)(src + (size_t)(i + k) * stride));
            draw_tri(ob, vv);
        }
    }
    else
    {
        for (unsigned i = 0; i < count; i += 3)
        {
            for (int k = 0; k < 3; ++k)
            {
                unsigned id = idx_size == 2 ? ((const uint16_t*)idx)[start + i + k] : ((const uint32_t*)idx)[start + i + k];
                vv[k] = xform(mvp, (const float*)(vb + (size_t)id * stride));
            }
            draw_tri(ob, vv);
        }
    }
    ob->dirty = 1;
    return ok;
}

/* OB.cpp:234-329 */
void orc_ob_build_hierarchy(OrcOB* ob)
{
    if (!ob->base || !ob->dirty)
        return;
    int pw = ob->w, ph = ob->h;
    for (int lv = 0; lv < ob->nlev; ++lv)
    {
        int w = ob->mw[lv], h = ob->mh[lv];
        int* dst = ob->mip[lv];
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
            {
                int mn = INT32_MAX, mx = INT32_MIN;
                int rows = (2 * y + 1 < ph) ? 2 : 1;
                for (int r = 0; r < rows; ++r)
                    for (int c = 0; c < 2; ++c)
                    {
                        /* the reference always reads two texels per row (src[0], src[1]); with a power-of-two width >= 2
                         * the second one always exists.  For a 1-texel-wide level it reads the next row's first texel. */
                        size_t o = (size_t)(2 * y + r) * pw + 2 * x + c;
                        int lo, hi;
                        if (lv == 0) { lo = hi = ob->depth[o]; }
                        else { lo = ob->mip[lv - 1][2 * o]; hi = ob->mip[lv - 1][2 * o + 1]; }
                        if (lo < mn) mn = lo;
                        if (hi > mx) mx = hi;
                    }
                dst[2 * ((size_t)y * w + x)] = mn;
                dst[2 * ((size_t)y * w + x) + 1] = mx;
            }
        pw = w;
        ph = h;
    }
    ob->dirty = 0;
}

/* OB.cpp:336-456 */
int orc_ob_is_visible(const OrcOB* ob, const float* b /* minx,miny,minz,maxx,maxy,maxz */)
{
    if (!ob->base)
        return 1;
    float min_x = 0, max_x = 0, min_y = 0, max_y = 0, min_z = 0;
    for (int i = 0; i < 8; ++i)
    {
        float p[3] = {b[(i & 1) ? 3 : 0], b[(i & 2) ? 4 : 1], b[(i & 4) ? 5 : 2]};
        V4 v = xform(ob->vp, p);
        v.z -= REL_BIAS;
        /* The reference biases all 8 first and then walks them in order, returning at the first z <= 0; since the answer is
         * "visible" whichever corner trips, checking while transforming gives the same result. */
        if (v.z <= 0.0f)
            return 1;
        V3 s = to_screen(ob, v);
        if (i == 0)
        {
            min_x = max_x = s.x;
            min_y = max_y = s.y;
            min_z = s.z;
        }
        else
        {
            if (s.x < min_x) min_x = s.x;
            if (s.x > max_x) max_x = s.x;
            if (s.y < min_y) min_y = s.y;
            if (s.y > max_y) max_y = s.y;
            if (s.z < min_z) min_z = s.z;
        }
    }
    int left = trunc_f(min_x - 1.5f), top = trunc_f(min_y - 1.5f), right = round_to_int(max_x), bottom = round_to_int(max_y);
    if (right < 0 || bottom < 0)
        return 1;
    if (left >= ob->w || top >= ob->h)
        return 1;
    if (left < 0) left = 0;
    if (top < 0) top = 0;
    if (right >= ob->w) right = ob->w - 1;
    if (bottom >= ob->h) bottom = ob->h - 1;
    int z = (int)((uint32_t)round_to_int(min_z) - FIXED_BIAS);

    if (!ob->dirty)
    {
        for (int lv = ob->nlev - 1; lv >= 0; --lv)
        {
            int sh = lv + 1;
            int roww = ob->w >> sh; /* the reference indexes rows with width_ >> shift, OB.cpp:410 */
            int all_occluded = 1;
            for (int y = top >> sh; y <= (bottom >> sh); ++y)
                for (int x = left >> sh; x <= (right >> sh); ++x)
                {
                    const int* t = ob->mip[lv] + 2 * ((size_t)y * roww + x);
                    if (z <= t[0])
                        return 1;
                    if (z <= t[1])
                        all_occluded = 0;
                }
            if (all_occluded)
                return 0;
        }
    }
    for (int y = top; y <= bottom; ++y)
        for (int x = left; x <= right; ++x)
            if (z <= ob->depth[(size_t)y * ob->w + x])
                return 1;
    return 0;
}

void orc_ob_is_visible_many(const OrcOB* ob, int n, const float* aabb6, unsigned char* out)
{
    for (int i = 0; i < n; ++i)
        out[i] = (unsigned char)orc_ob_is_visible(ob, aabb6 + 6 * (size_t)i);
}

void orc_ob_read_depth(const OrcOB* ob, int* out) { memcpy(out, ob->depth, sizeof(int) * (size_t)ob->w * ob->h); }
void orc_ob_read_mip(const OrcOB* ob, int lv, int* out) { memcpy(out, ob->mip[lv], sizeof(int) * 2 * (size_t)ob->mw[lv] * ob->mh[lv]); }
void orc_ob_mip_size(const OrcOB* ob, int lv, int* w, int* h) { *w = ob->mw[lv]; *h = ob->mh[lv]; }

/* Frustum::IsInside / IsInsideFast, Math/Frustum.h:123-159.  planes = 6 x (nx,ny,nz,d); absNormal = |n| (Plane.h:83-88). */
int orc_frustum_test(const float* planes24, const float* b, int fast)
{
    float cx = (b[3] + b[0]) * 0.5f, cy = (b[4] + b[1]) * 0.5f, cz = (b[5] + b[2]) * 0.5f; /* BoundingBox::Center, BoundingBox.h:265 */
    float ex = cx - b[0], ey = cy - b[1], ez = cz - b[2];
    int all_inside = 1;
    for (int i = 0; i < 6; ++i)
    {
        const float* p = planes24 + 4 * i;
        float dist = p[0] * cx + p[1] * cy + p[2] * cz + p[3];
        float ad = fabsf(p[0]) * ex + fabsf(p[1]) * ey + fabsf(p[2]) * ez;
        if (dist < -ad)
            return OUTSIDE;
        if (!fast && dist < ad)
            all_inside = 0;
    }
    return (fast || all_inside) ? INSIDE : INTERSECTS;
}

/* Flattened octree in DFS pre-order (the visiting order of Octant::GetDrawablesInternal, Octree.cpp:229-255):
 * octant i has parent[i] (< i, -1 for the root), culling box box6[i], and owns drawables [first[i], first[i]+count[i])
 * of the DFS-ordered drawable arrays.  num_desc[i] = number of octants in i's subtree excluding i. */
typedef struct
{
    int m;
    const int* parent;
    const float* box6;
    const int* first;
    const int* count;
    int n;
    const float* aabb6;          /* DFS order */
    const unsigned char* flags;  /* drawableFlags */
    const unsigned* view_mask;
    const unsigned char* occludee;
    const unsigned char* cast_shadows;
} OrcTree;

enum { Q_FRUSTUM = 0, Q_OCCLUDED = 1, Q_SHADOWCASTER = 2, Q_SPHERE = 3, Q_BOX = 4, Q_POINT = 5 };

/* Sphere::IsInside / IsInsideFast(BoundingBox), Math/Sphere.cpp:136-256.  s = centre.xyz, radius. */
static int sphere_test(const float* s, const float* b, int fast)
{
    float r2 = s[3] * s[3], d2 = 0.0f, t;
    for (int a = 0; a < 3; ++a)
    {
        if (s[a] < b[a]) { t = s[a] - b[a]; d2 += t * t; }
        else if (s[a] > b[3 + a]) { t = s[a] - b[3 + a]; d2 += t * t; }
    }
    if (d2 >= r2)
        return OUTSIDE;
    if (fast)
        return INSIDE;
    float lo[3] = {b[0] - s[0], b[1] - s[1], b[2] - s[2]}, hi[3] = {b[3] - s[0], b[4] - s[1], b[5] - s[2]};
    /* corner walk of the reference: ---, +--, ++-, -+-, -++, --+, +-+, +++ */
    static const int sel[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 1, 1}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}};
    for (int c = 0; c < 8; ++c)
    {
        float x = sel[c][0] ? hi[0] : lo[0], y = sel[c][1] ? hi[1] : lo[1], z = sel[c][2] ? hi[2] : lo[2];
        if (x * x + y * y + z * z >= r2)
            return INTERSECTS;
    }
    return INSIDE;
}

/* BoundingBox::IsInside / IsInsideFast(BoundingBox) with q as `this`, BoundingBox.h:295-315. */
static int box_test(const float* q, const float* b, int fast)
{
    if (b[3] < q[0] || b[0] > q[3] || b[4] < q[1] || b[1] > q[4] || b[5] < q[2] || b[2] > q[5])
        return OUTSIDE;
    if (!fast && (b[0] < q[0] || b[3] > q[3] || b[1] < q[1] || b[4] > q[4] || b[2] < q[2] || b[5] > q[5]))
        return INTERSECTS;
    return INSIDE;
}

/* BoundingBox::IsInside(Vector3) with b as `this`, BoundingBox.h:285-292. */
static int point_test(const float* p, const float* b)
{
    if (p[0] < b[0] || p[0] > b[3] || p[1] < b[1] || p[1] > b[4] || p[2] < b[2] || p[2] > b[5])
        return OUTSIDE;
    return INSIDE;
}

static int shape_test(int mode, const float* params, const float* b, int fast)
{
    if (mode == Q_SPHERE)
        return sphere_test(params, b, fast);
    if (mode == Q_BOX)
        return box_test(params, b, fast);
    if (mode == Q_POINT)
        return point_test(params, b);
    return orc_frustum_test(params, b, fast);
}

/* Octree::GetDrawables(query) for FrustumOctreeQuery (OctreeQuery.cpp:98-118), OccludedFrustumOctreeQuery (View.cpp:116-158),
 * ShadowCasterOctreeQuery (View.cpp:59-84) and the Sphere / Box / Point queries (OctreeQuery.cpp:32-96; parameters in planes24).  Writes DFS positions of the result; returns their number. */
int orc_query(const OrcTree* t, int mode, const float* planes24, const OrcOB* ob, unsigned flags, unsigned view_mask, int* out)
{
    /* iterative pre-order walk: state[i] 0 = culled, 1 = visited with inside=false, 2 = visited with inside=true */
    unsigned char* state = (unsigned char*)malloc((size_t)t->m);
    int nout = 0;
    for (int i = 0; i < t->m; ++i)
    {
        int inside = 0;
        if (i > 0)
        {
            unsigned char ps = state[t->parent[i]];
            if (!ps)
            {
                state[i] = 0;
                continue;
            }
            inside = ps == 2;
            const float* box = t->box6 + 6 * (size_t)i;
            int res;
            if (mode == Q_OCCLUDED)
            {
                if (inside)
                    res = orc_ob_is_visible(ob, box) ? INSIDE : OUTSIDE;
                else
                {
                    res = orc_frustum_test(planes24, box, 0);
                    if (res != OUTSIDE && !orc_ob_is_visible(ob, box))
                        res = OUTSIDE;
                }
            }
            else
                res = inside ? INSIDE : shape_test(mode, planes24, box, 0);
            if (res == OUTSIDE)
            {
                state[i] = 0;
                continue;
            }
            if (res == INSIDE)
                inside = 1;
        }
        state[i] = inside ? 2 : 1;
        for (int d = t->first[i]; d < t->first[i] + t->count[i]; ++d)
        {
            if (mode == Q_SHADOWCASTER && !t->cast_shadows[d])
                continue;
            if ((t->flags[d] & flags) && (t->view_mask[d] & view_mask))
                if (inside || shape_test(mode, planes24, t->aabb6 + 6 * (size_t)d, 1) != OUTSIDE)
                    out[nout++] = d;
        }
    }
    free(state);
    return nout;
}

/* CheckVisibilityWork's occlusion predicate, View.cpp:177, over a query result (DFS positions); keeps order. */
int orc_check_visibility(const OrcTree* t, const OrcOB* ob, const int* cand, int ncand, int* out)
{
    int n = 0;
    for (int i = 0; i < ncand; ++i)
    {
        int d = cand[i];
        if (!ob || !t->occludee[d] || orc_ob_is_visible(ob, t->aabb6 + 6 * (size_t)d))
            out[n++] = d;
    }
    return n;
}

/* The rest of CheckVisibilityWork that decides what is "in view" (View.cpp:177-211): occlusion predicate, draw-distance test on
 * Drawable::UpdateBatches' distance (Drawable.cpp:134-137, Camera::GetDistance Camera.cpp:487-496), and the view-space Z range of the
 * geometries that stay, merged as View::GetDrawables does (View.cpp:893-894, 926-951).  draw_dist is in DFS order (NULL = all 0).
 * The orthographic distance follows the operation order of the reference's default URHO3D_SSE build of Matrix3x4 * Vector3
 * (Matrix3x4.h:256-274): (m20*x + m22*z) + (m21*y + m23). */
int orc_check_in_view(const OrcTree* t, const OrcOB* ob, const float* view12, const float* cam_pos3, int ortho, const float* draw_dist,
    const int* cand, int ncand, int* out, float* minmax2)
{
    int n = 0;
    float min_z = (float)HUGE_VAL, max_z = 0.0f;
    const float vz0 = view12[8], vz1 = view12[9], vz2 = view12[10], vz3 = view12[11];
    for (int i = 0; i < ncand; ++i)
    {
        int d = cand[i];
        const float* b = t->aabb6 + 6 * (size_t)d;
        if (!(!ob || !t->occludee[d] || orc_ob_is_visible(ob, b)))
            continue;
        float cx = (b[3] + b[0]) * 0.5f, cy = (b[4] + b[1]) * 0.5f, cz = (b[5] + b[2]) * 0.5f;
        float max_distance = draw_dist ? draw_dist[d] : 0.0f;
        if (max_distance > 0.0f)
        {
            float dist;
            if (!ortho)
            {
                float dx = cx - cam_pos3[0], dy = cy - cam_pos3[1], dz = cz - cam_pos3[2];
                dist = sqrtf(dx * dx + dy * dy + dz * dz);
            }
            else
                dist = fabsf((vz0 * cx + vz2 * cz) + (vz1 * cy + vz3));
            if (dist > max_distance)
                continue;
        }
        if (t->flags[d] & 1) /* DRAWABLE_GEOMETRY */
        {
            float ex = (b[3] - b[0]) * 0.5f, ey = (b[4] - b[1]) * 0.5f, ez = (b[5] - b[2]) * 0.5f;
            if (ex * ex + ey * ey + ez * ez < 100000000.0f * 100000000.0f)
            {
                float vc = vz0 * cx + vz1 * cy + vz2 * cz + vz3;
                float ve = fabsf(vz0) * ex + fabsf(vz1) * ey + fabsf(vz2) * ez;
                float lo = vc - ve, hi = vc + ve;
                min_z = lo < min_z ? lo : min_z;
                max_z = hi > max_z ? hi : max_z;
            }
        }
        out[n++] = d;
    }
    if (min_z == (float)HUGE_VAL)
        min_z = 0.0f;
    minmax2[0] = min_z;
    minmax2[1] = max_z;
    return n;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * Ray queries: Octree::Raycast / RaycastSingle at RAY_AABB level (Octree.cpp:257-282, 284-309, 501-548; Drawable::ProcessRayQuery,
 * Drawable.cpp:118-132; Ray::HitDistance(BoundingBox), Math/Ray.cpp:71-148).  TEST INFRASTRUCTURE ONLY.
 * ------------------------------------------------------------------------------------------------------------------------- */
static int in_range(float v, float lo, float hi) { return v >= lo && v <= hi; }

/* Ray::HitDistance(const BoundingBox&), Math/Ray.cpp:71-148.  ray6 = origin.xyz, direction.xyz; b = min.xyz, max.xyz. */
float orc_ray_hit_distance(const float* ray6, const float* b)
{
    const float* o = ray6;
    const float* d = ray6 + 3;
    if (b[0] == INFINITY) /* !box.Defined(), BoundingBox.h:258-261 */
        return INFINITY;
    if (!(o[0] < b[0] || o[0] > b[3] || o[1] < b[1] || o[1] > b[4] || o[2] < b[2] || o[2] > b[5]))
        return 0.0f; /* origin inside the box */
    float dist = INFINITY;
    for (int axis = 0; axis < 3; ++axis)
    {
        const int u = (axis + 1) % 3, w = (axis + 2) % 3;
        for (int side = 0; side < 2; ++side)
        {
            /* min face when coming from below it, max face when coming from above it */
            const float face = side == 0 ? b[axis] : b[3 + axis];
            const int towards = side == 0 ? (o[axis] < face && d[axis] > 0.0f) : (o[axis] > face && d[axis] < 0.0f);
            if (!towards)
                continue;
            const float x = (face - o[axis]) / d[axis];
            if (x < dist)
            {
                const float pu = o[u] + x * d[u], pw = o[w] + x * d[w];
                if (in_range(pu, b[u], b[3 + u]) && in_range(pw, b[w], b[3 + w]))
                    dist = x;
            }
        }
    }
    return dist;
}

/* Sort() of Container/Sort.h:70-141 applied to (key, payload) pairs with `a.key < b.key` as the compare: median-of-three quicksort
 * (Hoare partition) down to partitions of 16, then one insertion sort over the whole range.  Not stable; the order of equal keys is a
 * property of this exact procedure, which is why it is restated instead of calling qsort. */
typedef struct { float key; int val; } KeyVal;

static void kv_quick(KeyVal* a, long begin, long end)
{
    while (end - begin > 16)
    {
        long pivot = begin + (end - begin) / 2;
        if (a[begin].key < a[pivot].key && a[end - 1].key < a[begin].key)
            pivot = begin;
        else if (a[end - 1].key < a[pivot].key && a[begin].key < a[end - 1].key)
            pivot = end - 1;
        long i = begin - 1, j = end;
        const KeyVal pv = a[pivot];
        for (;;)
        {
            while (pv.key < a[--j].key) {}
            while (a[++i].key < pv.key) {}
            if (i < j)
            {
                const KeyVal t = a[i];
                a[i] = a[j];
                a[j] = t;
            }
            else
                break;
        }
        kv_quick(a, begin, j + 1);
        begin = j + 1;
    }
}

static void kv_sort(KeyVal* a, long n)
{
    kv_quick(a, 0, n);
    for (long i = 1; i < n; ++i)
    {
        const KeyVal t = a[i];
        long j = i;
        while (j > 0 && t.key < a[j - 1].key)
        {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = t;
    }
}

/* Exposed for the tests of the product's own host-side sort. */
void orc_sort_by_key(int n, float* keys, int* vals)
{
    KeyVal* a = (KeyVal*)malloc(sizeof(KeyVal) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i)
    {
        a[i].key = keys[i];
        a[i].val = vals[i];
    }
    kv_sort(a, n);
    for (int i = 0; i < n; ++i)
    {
        keys[i] = a[i].key;
        vals[i] = a[i].val;
    }
    free(a);
}

/* Octree::Raycast (single = 0) / RaycastSingle (single = 1) with a RAY_AABB-level query.  ray7 = origin.xyz, direction.xyz, maxDistance.
 * Writes DFS positions and distances in result_ order; returns the result size. */
int orc_raycast(const OrcTree* t, const float* ray7, unsigned flags, unsigned view_mask, int single, int* out_pos, float* out_dist)
{
    const float maxd = ray7[6];
    unsigned char* state = (unsigned char*)malloc((size_t)t->m);
    KeyVal* hits = (KeyVal*)malloc(sizeof(KeyVal) * (size_t)(t->n > 0 ? t->n : 1));
    long nh = 0;
    for (int i = 0; i < t->m; ++i)
    {
        /* every octant is tested, the root included (Octree.cpp:257-261) */
        if (i > 0 && !state[t->parent[i]])
        {
            state[i] = 0;
            continue;
        }
        const float od = orc_ray_hit_distance(ray7, t->box6 + 6 * (size_t)i);
        if (od >= maxd)
        {
            state[i] = 0;
            continue;
        }
        state[i] = 1;
        for (int d = t->first[i]; d < t->first[i] + t->count[i]; ++d)
        {
            if (!((t->flags[d] & flags) && (t->view_mask[d] & view_mask)))
                continue;
            const float dist = orc_ray_hit_distance(ray7, t->aabb6 + 6 * (size_t)d);
            if (single || dist < maxd) /* RaycastSingle collects first (GetDrawablesOnlyInternal), Raycast tests at once (ProcessRayQuery) */
            {
                hits[nh].key = dist;
                hits[nh].val = d;
                ++nh;
            }
        }
    }
    kv_sort(hits, nh);
    int n = 0;
    if (!single)
    {
        for (long i = 0; i < nh; ++i, ++n)
        {
            out_pos[n] = hits[i].val;
            out_dist[n] = hits[i].key;
        }
    }
    else if (nh && hits[0].key < maxd)
    {
        /* the first sorted candidate passes `sortValue < Min(closestHit, maxDistance)` and sets closestHit to its own distance; every later
         * one has sortValue >= closestHit and ends the loop (Octree.cpp:528-541) */
        out_pos[0] = hits[0].val;
        out_dist[0] = hits[0].key;
        n = 1;
    }
    free(hits);
    free(state);
    return n;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * Occluder selection: View::UpdateOccluders (View.cpp:2171-2229) + the triangle budget of View::DrawOccluders' threaded branch
 * (View.cpp:2258-2270, OcclusionBuffer::AddTriangles' return value OB.cpp:178-179, 196-197).  TEST INFRASTRUCTURE ONLY.
 * cam = view12 (3x4), cam_pos3, ortho, half_view_size, ortho_size, far_clip.  Candidates: table entries passing IsInsideFast, table order.
 * Writes the sorted occluder indices / sort values; *submitted_n = how many of them the budget lets through.  Returns the sorted count.
 * ------------------------------------------------------------------------------------------------------------------------- */
int orc_select_occluders(int num_occ, const float* occ_aabb6, const unsigned* num_tris, const float* draw_dist, const float* planes24,
    const float* view12, const float* cam_pos3, int ortho, float half_view_size, float ortho_size, float far_clip, float size_threshold,
    unsigned max_triangles, int* out_idx, float* out_keys, int* submitted_n)
{
    KeyVal* a = (KeyVal*)malloc(sizeof(KeyVal) * (size_t)(num_occ > 0 ? num_occ : 1));
    long n = 0;
    const float inv_ortho = 1.0f / ortho_size;
    const float vz0 = view12[8], vz1 = view12[9], vz2 = view12[10], vz3 = view12[11];
    for (int i = 0; i < num_occ; ++i)
    {
        const float* b = occ_aabb6 + 6 * (size_t)i;
        if (orc_frustum_test(planes24, b, 1) == OUTSIDE)
            continue;
        float cx = (b[3] + b[0]) * 0.5f, cy = (b[4] + b[1]) * 0.5f, cz = (b[5] + b[2]) * 0.5f;
        float distance;
        if (!ortho)
        {
            float dx = cx - cam_pos3[0], dy = cy - cam_pos3[1], dz = cz - cam_pos3[2];
            distance = sqrtf(dx * dx + dy * dy + dz * dz);
        }
        else
            distance = fabsf((vz0 * cx + vz2 * cz) + (vz1 * cy + vz3));
        float max_distance = draw_dist ? draw_dist[i] : 0.0f;
        if (!(max_distance <= 0.0f || distance <= max_distance))
            continue;
        float sx = b[3] - b[0], sy = b[4] - b[1], sz = b[5] - b[2];
        float diagonal = sqrtf(sx * sx + sy * sy + sz * sz);
        float compare;
        if (!ortho)
        {
            float fraction = distance / far_clip;
            compare = diagonal * half_view_size / (distance * fraction);
            const float* p = cam_pos3;
            if (!(p[0] < b[0] || p[0] > b[3] || p[1] < b[1] || p[1] > b[4] || p[2] < b[2] || p[2] > b[5]))
                compare *= diagonal;
        }
        else
            compare = diagonal * inv_ortho;
        if (compare < size_threshold)
            continue;
        float density = (float)num_tris[i] / diagonal;
        a[n].key = density / compare;
        a[n].val = i;
        ++n;
    }
    kv_sort(a, n);
    unsigned long long total = 0;
    int sub = 0;
    for (long i = 0; i < n; ++i)
    {
        out_idx[i] = a[i].val;
        out_keys[i] = a[i].key;
    }
    for (long i = 0; i < n; ++i)
    {
        ++sub;
        total += num_tris[a[i].val];
        if (total > max_triangles)
            break;
    }
    *submitted_n = sub;
    free(a);
    return (int)n;
}

/* Octree::GetDrawables with ZoneOccluderOctreeQuery (View.cpp:87-113): FrustumOctreeQuery's octant test (OctreeQuery.cpp:98-104); a drawable
 * passes when its flags are exactly DRAWABLE_ZONE (4), or exactly DRAWABLE_GEOMETRY (1) with IsOccluder(), and its view mask matches, and it
 * is inside the frustum (octant inside, or IsInsideFast).  occluder[] is per DFS slot.  TEST INFRASTRUCTURE ONLY. */
int orc_query_zone_occluder(const OrcTree* t, const float* planes24, const unsigned char* occluder, unsigned view_mask, int* out)
{
    unsigned char* state = (unsigned char*)malloc((size_t)t->m);
    int nout = 0;
    for (int i = 0; i < t->m; ++i)
    {
        int inside = 0;
        if (i > 0)
        {
            unsigned char ps = state[t->parent[i]];
            if (!ps)
            {
                state[i] = 0;
                continue;
            }
            inside = ps == 2;
            int res = inside ? INSIDE : orc_frustum_test(planes24, t->box6 + 6 * (size_t)i, 0);
            if (res == OUTSIDE)
            {
                state[i] = 0;
                continue;
            }
            if (res == INSIDE)
                inside = 1;
        }
        state[i] = inside ? 2 : 1;
        for (int d = t->first[i]; d < t->first[i] + t->count[i]; ++d)
        {
            const unsigned char f = t->flags[d];
            if ((f == 4 || (f == 1 && occluder[d])) && (t->view_mask[d] & view_mask))
                if (inside || orc_frustum_test(planes24, t->aabb6 + 6 * (size_t)d, 1) != OUTSIDE)
                    out[nout++] = d;
        }
    }
    free(state);
    return nout;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * View::ProcessShadowCasters' per-drawable filter + View::IsShadowCasterVisible (View.cpp:2409-2489).  TEST INFRASTRUCTURE ONLY.
 * Arithmetic follows the reference's default (URHO3D_SSE) build: BoundingBox::Transformed (BoundingBox.cpp:131-156) sums each row as
 * (m0*cx + m2*cz) + (m1*cy + m3), Matrix3x4 * Vector3 likewise (Matrix3x4.h:256-274).  All per-drawable arrays are indexed by drawable id.
 * ------------------------------------------------------------------------------------------------------------------------- */
static void box_transformed(const float* b, const float* m, float* out6)
{
    const float cx = (b[0] + b[3]) * 0.5f, cy = (b[1] + b[4]) * 0.5f, cz = (b[2] + b[5]) * 0.5f;
    const float hx = cx - b[0], hy = cy - b[1], hz = cz - b[2];
    for (int r = 0; r < 3; ++r)
    {
        const float* row = m + 4 * r;
        const float c = (row[0] * cx + row[2] * cz) + (row[1] * cy + row[3]);
        const float d = (fabsf(row[0] * hx) + fabsf(row[2] * hz)) + (fabsf(row[1] * hy) + 0.0f);
        out6[r] = c - d;
        out6[3 + r] = c + d;
    }
}

int orc_process_shadow_casters(int num_in, const int* in_ids, const float* aabb6, const unsigned char* cast_shadows, const float* light_view12,
    const float* shadow_planes24, const float* lvf_planes24, float lvf_max_z, float shadow_far_clip, int shadow_ortho, int light_type,
    unsigned light_mask, const unsigned* shadow_mask, const float* shadow_distance, const float* draw_distance, const unsigned char* in_view,
    const float* main_view12, const float* main_pos3, int main_ortho, int* out_ids)
{
    int n = 0;
    const float vz0 = main_view12[8], vz1 = main_view12[9], vz2 = main_view12[10], vz3 = main_view12[11];
    for (int k = 0; k < num_in; ++k)
    {
        const int id = in_ids[k];
        const float* b = aabb6 + 6 * (size_t)id;
        if (!cast_shadows[id])
            continue;
        if (!(shadow_mask[id] & light_mask))
            continue;
        if (light_type == 2 && orc_frustum_test(shadow_planes24, b, 1) == OUTSIDE)
            continue;
        float max_shadow = shadow_distance[id];
        const float draw = draw_distance[id];
        if (draw > 0.0f && (max_shadow <= 0.0f || draw < max_shadow))
            max_shadow = draw;
        const float cx = (b[3] + b[0]) * 0.5f, cy = (b[4] + b[1]) * 0.5f, cz = (b[5] + b[2]) * 0.5f;
        float distance;
        if (!main_ortho)
        {
            const float dx = cx - main_pos3[0], dy = cy - main_pos3[1], dz = cz - main_pos3[2];
            distance = sqrtf(dx * dx + dy * dy + dz * dz);
        }
        else
            distance = fabsf((vz0 * cx + vz2 * cz) + (vz1 * cy + vz3));
        if (max_shadow > 0.0f && distance > max_shadow)
            continue;
        float lv[6];
        box_transformed(b, light_view12, lv);
        int visible;
        if (shadow_ortho)
        {
            lv[5] = lv[5] > lvf_max_z ? lv[5] : lvf_max_z;
            visible = orc_frustum_test(lvf_planes24, lv, 1) != OUTSIDE;
        }
        else if (in_view[id])
            visible = 1;
        else
        {
            const float ccx = (lv[3] + lv[0]) * 0.5f, ccy = (lv[4] + lv[1]) * 0.5f, ccz = (lv[5] + lv[2]) * 0.5f;
            const float len2 = ccx * ccx + ccy * ccy + ccz * ccz;
            float dirx = ccx, diry = ccy, dirz = ccz; /* Ray(center, center): direction_ = center.Normalized(), Vector3.h:425-435 */
            if (!(fabsf(len2 - 1.0f) <= 0.000001f) && len2 > 0.0f)
            {
                const float inv = 1.0f / sqrtf(len2);
                dirx = ccx * inv;
                diry = ccy * inv;
                dirz = ccz * inv;
            }
            const float extrusion = shadow_far_clip;
            float original = sqrtf(len2);
            if (original < 0.000001f)
                original = 0.000001f;
            else if (original > extrusion)
                original = extrusion;
            const float factor = extrusion / original;
            const float ncx = dirx * extrusion, ncy = diry * extrusion, ncz = dirz * extrusion;
            const float hx = (lv[3] - lv[0]) * factor * 0.5f, hy = (lv[4] - lv[1]) * factor * 0.5f, hz = (lv[5] - lv[2]) * factor * 0.5f;
            const float e[6] = {ncx - hx, ncy - hy, ncz - hz, ncx + hx, ncy + hy, ncz + hz};
            for (int a = 0; a < 3; ++a)
            {
                lv[a] = lv[a] < e[a] ? lv[a] : e[a];
                lv[3 + a] = lv[3 + a] > e[3 + a] ? lv[3 + a] : e[3 + a];
            }
            visible = orc_frustum_test(lvf_planes24, lv, 1) != OUTSIDE;
        }
        if (visible)
            out_ids[n++] = id;
    }
    return n;
}
