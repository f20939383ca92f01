// Host-side loose octree that keeps exactly the structure the reference's Octree would hold for the same sequence of
// inserts / moves / removals, and flattens it into the DFS-ordered arrays the GPU scene consumes.
//
// Mirrors (behaviour, not code): Octant::Initialize (Octree.cpp:221-227), GetOrCreateChild (:82-107), InsertDrawable (:134-166),
// CheckDrawableFit (:168-190), AddDrawable / RemoveDrawable / Inc/DecDrawableCount (Octree.h:56-136), the reinsertion rule of
// Octree::Update (Octree.cpp:440-472) and the visiting order of GetDrawablesInternal (:229-255).
// In an engine integration the real Octree stays the owner and only the flatten step is needed.
// Compiled with -ffp-contract=off: centre / half-size / culling-box floats must match the reference's.
#include "u3d_host.h"

#include <algorithm>
#include <cstring>

namespace u3d
{

void HostOctree::InitOctant(HOctant& o, const float* mn, const float* mx, unsigned level, int parent, unsigned index)
{
    for (int a = 0; a < 3; ++a)
    {
        o.bmin[a] = mn[a];
        o.bmax[a] = mx[a];
        o.center[a] = (mx[a] + mn[a]) * 0.5f;        // BoundingBox::Center, BoundingBox.h:265
        o.half[a] = 0.5f * (mx[a] - mn[a]);          // 0.5f * box.Size(), Octree.cpp:225
        o.cmin[a] = mn[a] - o.half[a];               // cullingBox_, Octree.cpp:226
        o.cmax[a] = mx[a] + o.half[a];
    }
    o.level = level;
    o.parent = parent;
    o.index = index;
    o.numDrawables = 0;
    o.alive = true;
    o.drawables.clear();
    for (int& c : o.children)
        c = -1;
}

HostOctree::HostOctree(const float* mn, const float* mx, unsigned levels)
{
    SetSize(mn, mx, levels);
}

void HostOctree::SetSize(const float* mn, const float* mx, unsigned levels)
{
    // Octree::SetSize (Octree.cpp:340-353): drawables fall back to the root and are re-queued; here callers rebuild.
    ++version_;
    octants_.clear();
    freeList_.clear();
    octants_.emplace_back();
    InitOctant(octants_[0], mn, mx, 0, -1, 0xFFFFFFFFu);
    numLevels_ = std::max(levels, 1u); // Octree.cpp:351
    for (auto& d : drawables_)
        d.octant = -1;
    for (size_t i = 0; i < drawables_.size(); ++i)
        if (drawables_[i].present)
            Insert(0, (uint32_t)i);
}

int HostOctree::NewOctant()
{
    if (!freeList_.empty())
    {
        int i = freeList_.back();
        freeList_.pop_back();
        return i;
    }
    octants_.emplace_back();
    return (int)octants_.size() - 1;
}

int HostOctree::GetOrCreateChild(int oi, unsigned index)
{
    if (octants_[oi].children[index] >= 0)
        return octants_[oi].children[index];
    float mn[3], mx[3];
    {
        const HOctant& o = octants_[oi];
        for (int a = 0; a < 3; ++a)
        {
            mn[a] = o.bmin[a];
            mx[a] = o.bmax[a];
            const float c = (o.bmax[a] + o.bmin[a]) * 0.5f; // oldCenter
            if (index & (1u << a))
                mn[a] = c;
            else
                mx[a] = c;
        }
    }
    const unsigned level = octants_[oi].level + 1;
    ++version_;
    const int ci = NewOctant(); // may reallocate octants_
    InitOctant(octants_[ci], mn, mx, level, oi, index);
    octants_[oi].children[index] = ci;
    return ci;
}

bool HostOctree::CheckFit(const HOctant& o, const float* b) const
{
    const float sx = b[3] - b[0], sy = b[4] - b[1], sz = b[5] - b[2];
    if (o.level >= numLevels_ || sx >= o.half[0] || sy >= o.half[1] || sz >= o.half[2])
        return true;
    if (b[0] <= o.bmin[0] - 0.5f * o.half[0] || b[3] >= o.bmax[0] + 0.5f * o.half[0] || b[1] <= o.bmin[1] - 0.5f * o.half[1] ||
        b[4] >= o.bmax[1] + 0.5f * o.half[1] || b[2] <= o.bmin[2] - 0.5f * o.half[2] || b[5] >= o.bmax[2] + 0.5f * o.half[2])
        return true;
    return false;
}

static bool BoxFullyInside(const float* cmin, const float* cmax, const float* b) // BoundingBox::IsInside(box) == INSIDE, BoundingBox.h:295-305
{
    if (b[3] < cmin[0] || b[0] > cmax[0] || b[4] < cmin[1] || b[1] > cmax[1] || b[5] < cmin[2] || b[2] > cmax[2])
        return false;
    if (b[0] < cmin[0] || b[3] > cmax[0] || b[1] < cmin[1] || b[4] > cmax[1] || b[2] < cmin[2] || b[5] > cmax[2])
        return false;
    return true;
}

void HostOctree::IncCount(int oi)
{
    for (; oi >= 0; oi = octants_[oi].parent)
        ++octants_[oi].numDrawables;
}

void HostOctree::DecCount(int oi)
{
    // Octant::DecDrawableCount: an octant whose subtree becomes empty is deleted by its parent
    while (oi >= 0)
    {
        HOctant& o = octants_[oi];
        const int parent = o.parent;
        --o.numDrawables;
        if (!o.numDrawables && parent >= 0)
        {
            octants_[parent].children[o.index] = -1;
            o.alive = false;
            o.drawables.clear();
            freeList_.push_back(oi);
        }
        oi = parent;
    }
}

void HostOctree::Insert(int oi, uint32_t id)
{
    for (;;)
    {
        HDrawable& d = drawables_[id];
        const HOctant& o = octants_[oi];
        bool here;
        if (oi == 0)
            here = !d.occludee || !BoxFullyInside(o.cmin, o.cmax, d.box) || CheckFit(o, d.box);
        else
            here = CheckFit(o, d.box);
        if (here)
        {
            const int old = d.octant;
            if (old != oi)
            {
                // add first, then remove (Octree.cpp:148-153)
                ++version_;
                d.octant = oi;
                octants_[oi].drawables.push_back(id);
                IncCount(oi);
                if (old >= 0)
                {
                    auto& v = octants_[old].drawables;
                    auto it = std::find(v.begin(), v.end(), id);
                    if (it != v.end())
                    {
                        v.erase(it);
                        DecCount(old);
                    }
                }
            }
            return;
        }
        const float cx = (d.box[3] + d.box[0]) * 0.5f, cy = (d.box[4] + d.box[1]) * 0.5f, cz = (d.box[5] + d.box[2]) * 0.5f;
        const unsigned x = cx < o.center[0] ? 0 : 1;
        const unsigned y = cy < o.center[1] ? 0 : 2;
        const unsigned z = cz < o.center[2] ? 0 : 4;
        oi = GetOrCreateChild(oi, x + y + z);
    }
}

uint32_t HostOctree::AddDrawable(const float* box6, uint8_t flags, uint32_t viewMask, bool occludee, bool castShadows)
{
    HDrawable d;
    std::memcpy(d.box, box6, sizeof(d.box));
    d.flags = flags;
    d.viewMask = viewMask;
    d.occludee = occludee;
    d.castShadows = castShadows;
    d.octant = -1;
    d.present = true;
    d.drawDistance = 0.0f;
    drawables_.push_back(d);
    const uint32_t id = (uint32_t)drawables_.size() - 1;
    Insert(0, id);
    return id;
}

bool HostOctree::UpdateDrawable(uint32_t id, const float* box6)
{
    if (id >= drawables_.size() || !drawables_[id].present)
        return false;
    HDrawable& d = drawables_[id];
    std::memcpy(d.box, box6, sizeof(d.box));
    // Octree::Update reinsertion rule, Octree.cpp:452-458
    if (d.octant >= 0)
    {
        const HOctant& o = octants_[d.octant];
        if (d.occludee && BoxFullyInside(o.cmin, o.cmax, d.box) && CheckFit(o, d.box))
            return true;
    }
    Insert(0, id);
    return true;
}

bool HostOctree::RemoveDrawable(uint32_t id)
{
    if (id >= drawables_.size() || !drawables_[id].present)
        return false;
    HDrawable& d = drawables_[id];
    ++version_;
    if (d.octant >= 0)
    {
        auto& v = octants_[d.octant].drawables;
        auto it = std::find(v.begin(), v.end(), id);
        if (it != v.end())
        {
            v.erase(it);
            DecCount(d.octant);
        }
    }
    d.octant = -1;
    d.present = false;
    return true;
}

void HostOctree::FlattenRec(int oi, int parent, FlatOctree& f) const
{
    const HOctant& o = octants_[oi];
    const int me = (int)f.parent.size();
    f.parent.push_back(parent);
    f.level.push_back((int)o.level);
    for (int a = 0; a < 3; ++a)
        f.box6.push_back(o.cmin[a]);
    for (int a = 0; a < 3; ++a)
        f.box6.push_back(o.cmax[a]);
    f.first.push_back((int)f.order.size());
    f.count.push_back((int)o.drawables.size());
    for (uint32_t id : o.drawables)
        f.order.push_back((int)id);
    for (int c = 0; c < 8; ++c)
        if (o.children[c] >= 0)
            FlattenRec(o.children[c], me, f);
}

void HostOctree::Flatten(FlatOctree& f) const
{
    f = FlatOctree();
    FlattenRec(0, -1, f);
}

} // namespace u3d
