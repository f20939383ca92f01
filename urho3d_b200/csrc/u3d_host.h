// Host-side structures behind the C-ABI (include/u3d_gpu.h).
#pragma once
#include <cstdint>
#include <vector>

namespace u3d
{

struct FlatOctree
{
    std::vector<int> parent, level, first, count, order;
    std::vector<float> box6;
};

struct HOctant
{
    float bmin[3], bmax[3], center[3], half[3], cmin[3], cmax[3];
    unsigned level;
    int parent;
    unsigned index;
    int children[8];
    unsigned numDrawables;
    bool alive;
    std::vector<uint32_t> drawables;
};

struct HDrawable
{
    float box[6];
    uint8_t flags;
    uint32_t viewMask;
    bool occludee, castShadows, present;
    int octant;
    float drawDistance{0.0f};
    bool occluder{false};
};

/// See octree_host.cpp.
class HostOctree
{
public:
    HostOctree(const float* mn, const float* mx, unsigned levels);
    void SetSize(const float* mn, const float* mx, unsigned levels);
    uint32_t AddDrawable(const float* box6, uint8_t flags, uint32_t viewMask, bool occludee, bool castShadows);
    bool UpdateDrawable(uint32_t id, const float* box6);
    bool RemoveDrawable(uint32_t id);
    bool SetDrawDistances(uint32_t first, uint32_t n, const float* d)
    {
        if ((uint64_t)first + n > drawables_.size())
            return false;
        for (uint32_t i = 0; i < n; ++i)
            drawables_[first + i].drawDistance = d[i];
        return true;
    }
    bool SetOccluderFlags(uint32_t first, uint32_t n, const uint8_t* f)
    {
        if ((uint64_t)first + n > drawables_.size())
            return false;
        for (uint32_t i = 0; i < n; ++i)
            drawables_[first + i].occluder = f[i] != 0;
        return true;
    }
    void Flatten(FlatOctree& f) const;
    const std::vector<HDrawable>& Drawables() const { return drawables_; }
    unsigned NumLevels() const { return numLevels_; }
    /// Bumped whenever an octant's drawable list changes or an octant is created / deleted (i.e. when the flattened layout changes).
    uint64_t StructureVersion() const { return version_; }

private:
    void InitOctant(HOctant& o, const float* mn, const float* mx, unsigned level, int parent, unsigned index);
    int NewOctant();
    int GetOrCreateChild(int oi, unsigned index);
    bool CheckFit(const HOctant& o, const float* b) const;
    void IncCount(int oi);
    void DecCount(int oi);
    void Insert(int oi, uint32_t id);
    void FlattenRec(int oi, int parent, FlatOctree& f) const;

    std::vector<HOctant> octants_;
    std::vector<int> freeList_;
    std::vector<HDrawable> drawables_;
    unsigned numLevels_{8};
    uint64_t version_{0};
};

} // namespace u3d
