// Nearest-neighbour search shared by chamfer.cu (operator-level C ABI) and fit_prox.cu (fused PROX driver).
#pragma once
#include "common.cuh"

namespace lemo {

// brute force, both clouds dynamic: dist/idx of the nearest target for every query (first minimum wins)
int chamfer_nn_launch(const float* q, long long q_bs, int nq, const float* t, long long t_bs, int nt, int B, float* dist, int* idx,
                      cudaStream_t st);

// Static scene (the PROX scene mesh is fixed for a recording, fit_temp_loadprox_slide.py:366-372): the points are k-d sorted once (median splits of the longest axis)
// into tiles of SG_TILE with an axis-aligned box per tile; a query scans only the tiles whose box can still contain a closer (or
// equally close, lower-index) point.  Results are IDENTICAL to the brute-force scan -- same pinned distance arithmetic, same
// first-minimum rule -- at a few percent of its pair evaluations.
constexpr int SG_TILE = 128;    // points per tile
constexpr int SG_GROUP = 32;    // consecutive (tree-adjacent) tiles per group: one lane per tile when a group is opened
struct SceneGrid {
    int device = 0, n = 0, ntile = 0, ngroup = 0;
    float4* pts = nullptr;     // [ntile*SG_TILE]  x, y, z, original index (int bits); padding = far away
    float4* box = nullptr;     // [ntile pad SG_GROUP][2]   (min xyz, _), (max xyz, _); padding tiles = empty boxes far away
    float4* gbox = nullptr;    // [ngroup][2]      box of the group's tiles
};
int scene_grid_create(const float* scene_dev, int n, SceneGrid** out);      // synchronises (create time only)
void scene_grid_free(SceneGrid* g);
// q: queries of batch b start at q + b*q_bs (floats), nq per batch -> dist [B,nq], idx [B,nq] (original indices)
int scene_grid_query(const SceneGrid* g, const float* q, long long q_bs, int nq, int B, float* dist, int* idx, cudaStream_t st);

}  // namespace lemo
