// Chamfer nearest neighbour, forward + backward.
// Replaces the external `chamfer` CUDA extension (ChamferDistancePytorch@719b0f1, chamfer.cu) behind
// temp_prox/dist_chamfer.py:10-53: exact brute-force squared-L2 nearest neighbour both ways, first minimum
// wins (strict <), int32 indices; backward scatters 2*g*(x1-x2) to both clouds with atomics.
// Differences from the reference kernel: the target cloud may be SHARED across
// the batch (batch stride 0) instead of being replicated B times (fitting_temp_slide.py:748), targets are
// staged as float4 through 16 KB of shared memory, and each thread keeps 4 queries in registers so one
// broadcast LDS.128 feeds 4 distance evaluations.
#include "common.cuh"
#include "chamfer.cuh"
#include "../../include/lemo_b200.h"
#include <vector>
#include <algorithm>

namespace lemo {

constexpr int CH_TILE = 1024;   // targets per shared-memory tile
constexpr int CH_QT = 4;        // queries per thread

__global__ void __launch_bounds__(256) k_chamfer_nn(const float* __restrict__ q, long long q_bs, int nq, const float* __restrict__ t,
                                                    long long t_bs, int nt, float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ float4 s_t[CH_TILE];
    const int b = blockIdx.y;
    const float* qb = q + (size_t)b * q_bs;
    const float* tb = t + (size_t)b * t_bs;
    float qx[CH_QT], qy[CH_QT], qz[CH_QT], best[CH_QT];
    int bi[CH_QT];
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * CH_QT;
#pragma unroll
    for (int k = 0; k < CH_QT; ++k) {
        const int i = min(i0 + k, nq - 1);
        qx[k] = qb[i * 3]; qy[k] = qb[i * 3 + 1]; qz[k] = qb[i * 3 + 2];
        best[k] = 3.4e38f; bi[k] = 0;
    }
    for (int j0 = 0; j0 < nt; j0 += CH_TILE) {
        const int nj = min(CH_TILE, nt - j0);
        for (int j = threadIdx.x; j < nj; j += blockDim.x) {
            const float* p = tb + (size_t)(j0 + j) * 3;
            s_t[j] = make_float4(p[0], p[1], p[2], 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const float4 p = s_t[j];
#pragma unroll
            for (int k = 0; k < CH_QT; ++k) {
                const float dx = qx[k] - p.x, dy = qy[k] - p.y, dz = qz[k] - p.z;
                // pinned evaluation order (= what -fmad=true makes of the reference kernel's dx*dx + dy*dy + dz*dz, and what
                // oracle/csrc/chamfer_ref.c computes with fmaf): indices are bit-exact against the oracle, ties included
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (d < best[k]) { best[k] = d; bi[k] = j0 + j; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < CH_QT; ++k) {
        const int i = i0 + k;
        if (i < nq) { dist[(size_t)b * nq + i] = best[k]; idx[(size_t)b * nq + i] = bi[k]; }
    }
}

// gradient of sum_i g[i] * |a_i - c_{idx[i]}|^2 : da_i += 2g(a-c), dc_idx -= 2g(a-c)
__global__ void k_chamfer_bwd(const float* __restrict__ a, long long a_bs, int na, const float* __restrict__ c, long long c_bs,
                              const float* __restrict__ g, const int* __restrict__ idx, float* __restrict__ da, float* __restrict__ dc, int B) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * na) return;
    const int b = (int)(i / na), r = (int)(i - (long long)b * na);
    const float* ap = a + (size_t)b * a_bs + (size_t)r * 3;
    const int j = idx[i];
    const float* cp = c + (size_t)b * c_bs + (size_t)j * 3;
    const float gg = 2.f * g[i];
    float* dap = da + (size_t)b * a_bs + (size_t)r * 3;
    float* dcp = dc + (size_t)b * c_bs + (size_t)j * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = gg * (ap[k] - cp[k]);
        atomicAdd(&dap[k], v);
        atomicAdd(&dcp[k], -v);
    }
}

int chamfer_nn_launch(const float* q, long long q_bs, int nq, const float* t, long long t_bs, int nt, int B, float* dist, int* idx,
                      cudaStream_t st) {
    k_chamfer_nn<<<dim3(cdiv(nq, 256 * CH_QT), B), 256, 0, st>>>(q, q_bs, nq, t, t_bs, nt, dist, idx);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- static scene, tiled + pruned
// One warp per query; a tile (128 points) is scanned 4 points per lane with the pinned distance arithmetic of k_chamfer_nn, and
// (distance, original index) pairs are compared lexicographically, so the result is the brute-force scan's first minimum.
// A box bound is shrunk by 0.1 % before it is compared with `best`, so fp32 rounding can never hide an equal-distance point.
__device__ __forceinline__ void sg_scan_tile(const float4* __restrict__ tp, float qx, float qy, float qz, float& best, int& bi) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < SG_TILE / 32; ++k) {
        const float4 p = tp[k * 32 + lane];
        const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));       // pinned order, see k_chamfer_nn
        const int id = __float_as_int(p.w);
        if (d < best || (d == best && id < bi)) { best = d; bi = id; }
    }
}
__device__ __forceinline__ void sg_warp_min(float& best, int& bi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
}
__device__ __forceinline__ float sg_box_lb(const float4* __restrict__ b, float qx, float qy, float qz) {
    const float4 lo = __ldg(b), hi = __ldg(b + 1);
    const float ex = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f), ey = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f),
                ez = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
    return ex * ex + ey * ey + ez * ez;
}
__device__ __forceinline__ void sg_warp_argmin(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
// Two-level search, one warp per query.  Groups of 32 tree-adjacent tiles carry their own box, so a query touches
// ngroup + 32 * (groups that can still matter) boxes instead of every tile box.
//   pass 1: nearest group box -> its nearest tile box -> scan that tile: a near-optimal `best`;
//   pass 2: every group whose box bound (shrunk by 0.1 %) does not exceed `best` is opened (one lane per tile), and every tile whose
//           bound does not exceed `best` is scanned.  Bounds only ever prune boxes that cannot hold a point at distance <= best,
//           and (distance, original index) pairs are compared lexicographically: the result is the brute-force first minimum.
__global__ void __launch_bounds__(256) k_scene_query(const float4* __restrict__ pts, const float4* __restrict__ box, const float4* __restrict__ gbox,
                                                     int ntile, int ngroup, const float* __restrict__ q, long long q_bs, int nq, int B,
                                                     float* __restrict__ dist, int* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const long long wq = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wq >= (long long)B * nq) return;
    const int b = (int)(wq / nq), i = (int)(wq - (long long)b * nq);
    const float* qp = q + (size_t)b * q_bs + (size_t)i * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    // ---- pass 1
    float gmin = 3.4e38f;
    int gsel = 0;
    for (int g = lane; g < ngroup; g += 32) {
        const float lb = sg_box_lb(gbox + 2 * (size_t)g, qx, qy, qz);
        if (lb < gmin) { gmin = lb; gsel = g; }
    }
    sg_warp_argmin(gmin, gsel);
    float tl = sg_box_lb(box + 2 * ((size_t)gsel * SG_GROUP + lane), qx, qy, qz);
    int tsel = gsel * SG_GROUP + lane;
    sg_warp_argmin(tl, tsel);
    float best = 3.4e38f;
    int bi = 0x7fffffff;
    sg_scan_tile(pts + (size_t)tsel * SG_TILE, qx, qy, qz, best, bi);
    sg_warp_min(best, bi);
    // ---- pass 2
    for (int g0 = 0; g0 < ngroup; g0 += 32) {
        const int g = g0 + lane;
        const bool gneed = g < ngroup && sg_box_lb(gbox + 2 * (size_t)g, qx, qy, qz) * 0.999f <= best;
        unsigned gm = __ballot_sync(0xffffffffu, gneed);
        while (gm) {
            const int gl = g0 + __ffs(gm) - 1;
            gm &= gm - 1;
            const int t = gl * SG_GROUP + lane;
            const bool need = t != tsel && sg_box_lb(box + 2 * (size_t)t, qx, qy, qz) * 0.999f <= best;
            unsigned m = __ballot_sync(0xffffffffu, need);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                sg_scan_tile(pts + (size_t)(gl * SG_GROUP + l) * SG_TILE, qx, qy, qz, best, bi);
                sg_warp_min(best, bi);
            }
        }
    }
    if (lane == 0) { dist[wq] = best; idx[wq] = bi; }
}

// k-d ordering: split the longest axis at the median until a node holds <= SG_TILE points.  Leaves in tree order are the tiles
// (compact boxes: 2.4 tiles scanned per query on the config-4 scene; a Morton sort, tried first, leaves Z-curve jumps inside the
// 128-point runs -- boxes up to the size of the scene, 76 tiles per query), and SG_GROUP consecutive leaves are a subtree.
static void kd_split(const std::vector<float>& h, std::vector<int>& idx, int lo, int hi, std::vector<std::pair<int, int>>& leaves) {
    if (hi - lo <= SG_TILE) { leaves.push_back({lo, hi}); return; }
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = lo; i < hi; ++i)
        for (int a = 0; a < 3; ++a) { const float v = h[(size_t)idx[i] * 3 + a]; mn[a] = std::min(mn[a], v); mx[a] = std::max(mx[a], v); }
    int ax = 0;
    for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    const int mid = lo + (hi - lo) / 2;
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                     [&](int a, int b) { const float va = h[(size_t)a * 3 + ax], vb = h[(size_t)b * 3 + ax]; return va < vb || (va == vb && a < b); });
    kd_split(h, idx, lo, mid, leaves);
    kd_split(h, idx, mid, hi, leaves);
}

int scene_grid_create(const float* scene_dev, int n, SceneGrid** out) {
    LEMO_CHECK(scene_dev && n > 0 && out, "bad arguments");
    std::vector<float> h((size_t)n * 3);
    LEMO_CUDA(cudaMemcpy(h.data(), scene_dev, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::vector<std::pair<int, int>> leaves;
    kd_split(h, idx, 0, n, leaves);
    const int ntile = (int)leaves.size();
    const int ngroup = (ntile + SG_GROUP - 1) / SG_GROUP;
    std::vector<float4> pts((size_t)ntile * SG_TILE);
    const float4 far_lo = make_float4(1e18f, 1e18f, 1e18f, 0.f), far_hi = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    std::vector<float4> box((size_t)ngroup * SG_GROUP * 2), gbox((size_t)ngroup * 2);
    for (size_t t = 0; t < (size_t)ngroup * SG_GROUP; ++t) { box[2 * t] = far_lo; box[2 * t + 1] = far_hi; }   // padding tiles: never needed
    for (int t = 0; t < ntile; ++t) {
        float bl[3] = {3.4e38f, 3.4e38f, 3.4e38f}, bh[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        const int cnt = leaves[t].second - leaves[t].first;
        for (int k = 0; k < SG_TILE; ++k) {
            float4 p;
            if (k < cnt) {
                const int id = idx[leaves[t].first + k];
                p.x = h[(size_t)id * 3]; p.y = h[(size_t)id * 3 + 1]; p.z = h[(size_t)id * 3 + 2];
                int ii = id;
                memcpy(&p.w, &ii, 4);
                const float v[3] = {p.x, p.y, p.z};
                for (int a = 0; a < 3; ++a) { bl[a] = std::min(bl[a], v[a]); bh[a] = std::max(bh[a], v[a]); }
            } else {
                p.x = p.y = p.z = 1e18f;               // padding: (1e18)^2 is finite and never the minimum
                int ii = 0x7fffffff;
                memcpy(&p.w, &ii, 4);
            }
            pts[(size_t)t * SG_TILE + k] = p;
        }
        box[2 * (size_t)t] = make_float4(bl[0], bl[1], bl[2], 0.f);
        box[2 * (size_t)t + 1] = make_float4(bh[0], bh[1], bh[2], 0.f);
    }
    for (int g = 0; g < ngroup; ++g) {
        float4 lo4 = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 0.f), hi4 = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, 0.f);
        for (int t = g * SG_GROUP; t < std::min(ntile, (g + 1) * SG_GROUP); ++t) {
            lo4.x = std::min(lo4.x, box[2 * (size_t)t].x); lo4.y = std::min(lo4.y, box[2 * (size_t)t].y); lo4.z = std::min(lo4.z, box[2 * (size_t)t].z);
            hi4.x = std::max(hi4.x, box[2 * (size_t)t + 1].x); hi4.y = std::max(hi4.y, box[2 * (size_t)t + 1].y);
            hi4.z = std::max(hi4.z, box[2 * (size_t)t + 1].z);
        }
        gbox[2 * (size_t)g] = lo4; gbox[2 * (size_t)g + 1] = hi4;
    }
    SceneGrid* g = new SceneGrid();
    g->n = n; g->ntile = ntile; g->ngroup = ngroup;
    cudaGetDevice(&g->device);
    LEMO_CUDA(cudaMalloc((void**)&g->pts, pts.size() * sizeof(float4)));
    LEMO_CUDA(cudaMalloc((void**)&g->box, box.size() * sizeof(float4)));
    LEMO_CUDA(cudaMalloc((void**)&g->gbox, gbox.size() * sizeof(float4)));
    LEMO_CUDA(cudaMemcpy(g->pts, pts.data(), pts.size() * sizeof(float4), cudaMemcpyHostToDevice));
    LEMO_CUDA(cudaMemcpy(g->box, box.data(), box.size() * sizeof(float4), cudaMemcpyHostToDevice));
    LEMO_CUDA(cudaMemcpy(g->gbox, gbox.data(), gbox.size() * sizeof(float4), cudaMemcpyHostToDevice));
    *out = g;
    return 0;
}

void scene_grid_free(SceneGrid* g) {
    if (!g) return;
    cudaFree(g->pts); cudaFree(g->box); cudaFree(g->gbox);
    delete g;
}

int scene_grid_query(const SceneGrid* g, const float* q, long long q_bs, int nq, int B, float* dist, int* idx, cudaStream_t st) {
    LEMO_CHECK(g && q && dist && idx && nq > 0 && B > 0, "bad arguments");
    const long long threads = (long long)B * nq * 32;
    k_scene_query<<<cdiv(threads, 256), 256, 0, st>>>(g->pts, g->box, g->gbox, g->ntile, g->ngroup, q, q_bs, nq, B, dist, idx);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo

using namespace lemo;
struct LemoScene { lemo::SceneGrid* g; };
extern "C" {

int lemo_scene_create(const float* scene_points, int32_t n, LemoScene** out) {
    LEMO_CHECK(out, "null argument");
    SceneGrid* g = nullptr;
    LEMO_TRY(scene_grid_create(scene_points, n, &g));
    *out = new LemoScene{g};
    return 0;
}
int lemo_scene_destroy(LemoScene* s) {
    if (s) { scene_grid_free(s->g); delete s; }
    return 0;
}
int lemo_scene_query(const LemoScene* s, const float* xyz1, int32_t B, int32_t n, float* dist1, int32_t* idx1, void* stream) {
    LEMO_NVTX("lemo_scene_query");
    LEMO_CHECK(s && s->g, "null scene");
    return scene_grid_query(s->g, xyz1, (long long)n * 3, n, B, dist1, idx1, (cudaStream_t)stream);
}

int lemo_chamfer_forward(const float* xyz1, int32_t B, int32_t n, const float* xyz2, int32_t m, int64_t xyz2_batch_stride, float* dist1,
                         float* dist2, int32_t* idx1, int32_t* idx2, void* stream) {
    LEMO_NVTX("lemo_chamfer_forward");
    LEMO_CHECK(xyz1 && xyz2 && dist1 && idx1 && B > 0 && n > 0 && m > 0, "bad arguments");
    LEMO_CHECK((dist2 == nullptr) == (idx2 == nullptr), "dist2 and idx2 must both be given or both be NULL");
    LEMO_CHECK(xyz2_batch_stride == 0 || xyz2_batch_stride >= (int64_t)m * 3, "xyz2_batch_stride must be 0 (shared) or >= 3*m");
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_TRY(chamfer_nn_launch(xyz1, (long long)n * 3, n, xyz2, xyz2_batch_stride, m, B, dist1, idx1, st));
    // the scene -> body direction is skipped when the caller does not consume it (the PROX contact term reads dist1 only,
    // fitting_temp_slide.py:749-753: 100 000 x 1121 x B pair evaluations saved)
    if (dist2) k_chamfer_nn<<<dim3(cdiv(m, 256 * CH_QT), B), 256, 0, st>>>(xyz2, xyz2_batch_stride, m, xyz1, (long long)n * 3, n, dist2, idx2);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int lemo_chamfer_backward(const float* xyz1, int32_t B, int32_t n, const float* xyz2, int32_t m, int64_t xyz2_batch_stride,
                          const float* g_dist1, const float* g_dist2, const int32_t* idx1, const int32_t* idx2, float* d_xyz1,
                          float* d_xyz2, void* stream) {
    LEMO_NVTX("lemo_chamfer_backward");
    LEMO_CHECK(xyz1 && xyz2 && g_dist1 && idx1 && d_xyz1 && d_xyz2, "null argument");
    LEMO_CHECK((g_dist2 == nullptr) == (idx2 == nullptr), "g_dist2 and idx2 must both be given or both be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CUDA(cudaMemsetAsync(d_xyz1, 0, (size_t)B * n * 3 * sizeof(float), st));
    const size_t n2 = xyz2_batch_stride == 0 ? (size_t)m * 3 : (size_t)(B - 1) * xyz2_batch_stride + (size_t)m * 3;
    LEMO_CUDA(cudaMemsetAsync(d_xyz2, 0, n2 * sizeof(float), st));
    k_chamfer_bwd<<<cdiv((long long)B * n, 256), 256, 0, st>>>(xyz1, (long long)n * 3, n, xyz2, xyz2_batch_stride, g_dist1, idx1, d_xyz1, d_xyz2, B);
    if (g_dist2) k_chamfer_bwd<<<cdiv((long long)B * m, 256), 256, 0, st>>>(xyz2, xyz2_batch_stride, m, xyz1, (long long)n * 3, g_dist2, idx2, d_xyz2, d_xyz1, B);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
}
