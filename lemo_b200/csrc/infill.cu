// Infill pre-stage on the device: the clip-level steps either side of the AE fine-tune that the reference runs in numpy float64 on the
// host (one GPU -> numpy -> GPU bounce per clip):
//   * body representation   get_local_markers_4chan          utils/utils.py:209-265 (+ the loader's normalisation and [4,d,T] layout,
//                                                             loader/optimize_loader_amass_new.py:359-361,376-377)
//   * mask + reflect pad    opt_amass_temp.py:164-184
//   * crop, contact labels, de-normalisation, reconstruct_global_body   opt_amass_temp.py:205-325, utils/utils.py:180-203
// The arithmetic stays in double wherever the reference is float64 (a clip is 120 x 68 points: nothing here is throughput-bound; the point
// is to keep the clip on the device between the per-frame stage, the AE and the temporal stage).  float32 roundings happen exactly where
// the reference's dtypes force them (floor shift on the float32 input; de-normalised values assigned into a float32 array).
#include "common.cuh"
#include "../../include/lemo_b200.h"

namespace lemo {

constexpr int IF_P = 68;             // pelvis + 67 SSM2 markers
constexpr int IF_D = 208;            // 68*3 + 4 contact rows
constexpr int IF_PADT = 8;           // reflect pad in time (left = right)

struct Q { double w, x, y, z; };
__device__ __forceinline__ Q q_mul(const Q& q, const Q& r) {             // Quaternions.__mul__ (utils/Quaternions.py:93-104)
    return Q{r.w * q.w - r.x * q.x - r.y * q.y - r.z * q.z,
             r.w * q.x + r.x * q.w - r.y * q.z + r.z * q.y,
             r.w * q.y + r.x * q.z + r.y * q.w - r.z * q.x,
             r.w * q.z - r.x * q.y + r.y * q.x + r.z * q.w};
}
__device__ __forceinline__ Q q_conj(const Q& q) { return Q{q.w, -q.x, -q.y, -q.z}; }
__device__ __forceinline__ void q_rot(const Q& q, const double v[3], double o[3]) {    // imaginary part of q (0,v) q*
    const Q t = q_mul(q, q_mul(Q{0.0, v[0], v[1], v[2]}, q_conj(q)));
    o[0] = t.x; o[1] = t.y; o[2] = t.z;
}
__device__ __forceinline__ Q q_yaxis(double angle) {                       // from_angle_axis(angle, (0,1,0)), axis / (|axis| + 1e-10)
    const double a = 1.0 / (1.0 + 1e-10);
    return Q{cos(angle / 2.0), 0.0 * a * sin(angle / 2.0), a * sin(angle / 2.0), 0.0 * a * sin(angle / 2.0)};
}
__device__ __forceinline__ double q_pivot(const Q& q) {                    // Pivots.from_quaternions: atan2 of the rotated +z axis (x, z)
    const double z[3] = {0.0, 0.0, 1.0};
    double d[3];
    q_rot(q, z, d);
    return atan2(d[0], d[2]);
}

// ---------------------------------------------------------------------------------------------- body representation
__global__ void __launch_bounds__(256) k_repr_floor_min(const float* __restrict__ body, int n_pts, float* __restrict__ out) {
    __shared__ float s[32];
    float m = 3.4e38f;
    for (int i = threadIdx.x; i < n_pts; i += blockDim.x) m = fminf(m, body[(size_t)i * 3 + 2]);
    for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 3.4e38f;
        for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[0] = m;
    }
}
// swapped, floor-shifted coordinates of point p at frame t: (x, z - min [float32 subtraction], y)
__device__ __forceinline__ void repr_point(const float* __restrict__ body, float zmin, int t, int p, double o[3]) {
    const float* b = body + ((size_t)t * IF_P + p) * 3;
    o[0] = (double)b[0];
    o[1] = (double)(b[2] - zmin);
    o[2] = (double)b[1];
}
__global__ void k_repr_forward(const float* __restrict__ body, const float* __restrict__ zmin, int T, double* __restrict__ fwd) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    // shoulders / hips: indices 28,58 / 29,59 of the array WITH the reference joint in front (utils/utils.py:233) = 27,57 / 28,58 here
    double sl[3], sr[3], hl[3], hr[3], a[3];
    repr_point(body, zmin[0], t, 27, sl); repr_point(body, zmin[0], t, 57, sr);
    repr_point(body, zmin[0], t, 28, hl); repr_point(body, zmin[0], t, 58, hr);
    for (int k = 0; k < 3; ++k) a[k] = (sr[k] - sl[k]) + (hr[k] - hl[k]);
    const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    for (int k = 0; k < 3; ++k) a[k] /= n;
    fwd[t * 3 + 0] = a[1] * 0.0 - a[2] * 1.0;              // np.cross(across, (0,1,0))
    fwd[t * 3 + 1] = a[2] * 0.0 - a[0] * 0.0;
    fwd[t * 3 + 2] = a[0] * 1.0 - a[1] * 0.0;
}
// gaussian_filter1d(forward, sigma 20, mode 'nearest', truncate 4 => radius 80) + normalise + Quaternions.between(forward, (0,0,1))
__global__ void k_repr_rotation(const double* __restrict__ fwd, int T, double* __restrict__ rot, double* __restrict__ rot0_pivot) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int R = 80;
    double ksum = 0.0;
    for (int j = -R; j <= R; ++j) ksum += exp(-0.5 / 400.0 * (double)(j * j));
    double f[3] = {0.0, 0.0, 0.0};
    for (int j = -R; j <= R; ++j) {
        const double w = exp(-0.5 / 400.0 * (double)(j * j)) / ksum;
        const int u = min(max(t + j, 0), T - 1);
        for (int k = 0; k < 3; ++k) f[k] += w * fwd[u * 3 + k];
    }
    const double n = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int k = 0; k < 3; ++k) f[k] /= n;
    // a = f x (0,0,1); w = sqrt(|f|^2 |t|^2) + f.t
    const double ax = f[1] * 1.0 - f[2] * 0.0, ay = f[2] * 0.0 - f[0] * 1.0, az = f[0] * 0.0 - f[1] * 0.0;
    const double w = sqrt((f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) * 1.0) + f[2];
    const double qn = sqrt(w * w + ax * ax + ay * ay + az * az);
    const Q q{w / qn, ax / qn, ay / qn, az / qn};
    rot[t * 4 + 0] = q.w; rot[t * 4 + 1] = q.x; rot[t * 4 + 2] = q.y; rot[t * 4 + 3] = q.z;
    if (t == 0) rot0_pivot[0] = q_pivot(q);
}
// one thread per (frame t < T-1, row r < 208): channel 0 row r, and (r == 0) the three trajectory channels broadcast over all rows later
__global__ void k_repr_emit(const float* __restrict__ body, const float* __restrict__ contact, const float* __restrict__ zmin,
                            const double* __restrict__ rot, const double* __restrict__ stats, int T, float* __restrict__ repr) {
    const int Tm = T - 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Tm * IF_D) return;
    const int r = i / Tm, t = i - r * Tm;
    const Q q{rot[t * 4], rot[t * 4 + 1], rot[t * 4 + 2], rot[t * 4 + 3]};
    double ref[3];
    repr_point(body, zmin[0], t, 0, ref);                  // reference joint = pelvis * (1, 0, 1)
    double v0;
    if (r < IF_P * 3) {
        const int p = r / 3, c = r - p * 3;
        double v[3], o[3];
        repr_point(body, zmin[0], t, p, v);
        v[0] -= ref[0]; v[2] -= ref[2];
        q_rot(q, v, o);
        v0 = c == 0 ? o[0] : (c == 1 ? o[2] : o[1]);       // back to (x, y, z) order
    } else {
        v0 = (double)contact[t * 4 + (r - IF_P * 3)];
    }
    // trajectory channels: velocity of the reference joint rotated by q[t+1], heading change between t and t+1
    double ref1[3];
    repr_point(body, zmin[0], t + 1, 0, ref1);
    const Q q1{rot[(t + 1) * 4], rot[(t + 1) * 4 + 1], rot[(t + 1) * 4 + 2], rot[(t + 1) * 4 + 3]};
    const double vel[3] = {ref1[0] - ref[0], 0.0 - 0.0, ref1[2] - ref[2]};
    double vr[3];
    q_rot(q1, vel, vr);
    const double rv = q_pivot(q_mul(q1, q_conj(q)));
    double c0 = v0, c1 = vr[0], c2 = vr[2], c3 = rv;
    if (stats) {
        c0 = (c0 - stats[r]) / stats[IF_D + r];
        c1 = (c1 - stats[2 * IF_D]) / stats[2 * IF_D + 1];
        c2 = (c2 - stats[2 * IF_D]) / stats[2 * IF_D + 1];
        c3 = (c3 - stats[2 * IF_D + 2]) / stats[2 * IF_D + 3];
    }
    const size_t plane = (size_t)IF_D * Tm;
    repr[(size_t)r * Tm + t] = (float)c0;
    repr[plane + (size_t)r * Tm + t] = (float)c1;
    repr[2 * plane + (size_t)r * Tm + t] = (float)c2;
    repr[3 * plane + (size_t)r * Tm + t] = (float)c3;
}

// ---------------------------------------------------------------------------------------------- mask + reflect pad
__device__ __forceinline__ bool infill_row_masked(int r) {       // rows of channel 0 blanked for the AE (opt_amass_temp.py:167-180)
    if (r >= IF_D - 4) return true;
    const int p = r / 3;                                          // point index (0 = pelvis, 1.. = markers)
    const int m = p - 1;
    constexpr int ids[22] = {14, 15, 18, 19, 29, 2, 20, 21, 30, 25, 16, 45, 46, 48, 49, 59, 32, 50, 51, 55, 60, 47};
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 22; ++k) hit |= (m == ids[k]);
    return hit;
}
__global__ void k_infill_prepare(const float* __restrict__ clip, int T, float* __restrict__ xpad, float* __restrict__ row_mask) {
    const int Wd = T + 2 * IF_PADT, Hd = IF_D + 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Hd && row_mask) {
        // fine-tune loss rows: every padded row that is not a masked row (+1 for the pad), minus the last five (opt_amass_temp.py:196-200)
        const int r = i - 1;
        const bool masked = r >= 0 && r < IF_D - 4 && infill_row_masked(r);
        row_mask[i] = (!masked && i < Hd - 5) ? 1.f : 0.f;
    }
    if (i >= 4 * Hd * Wd) return;
    const int c = i / (Hd * Wd), rem = i - c * Hd * Wd, y = rem / Wd, x = rem - y * Wd;
    int r = y - 1, t = x - IF_PADT;
    if (r < 0) r = -r; else if (r >= IF_D) r = 2 * (IF_D - 1) - r;              // reflect (no edge repeat)
    if (t < 0) t = -t; else if (t >= T) t = 2 * (T - 1) - t;
    float v = clip[((size_t)c * IF_D + r) * T + t];
    if (c == 0 && infill_row_masked(r)) v = 0.f;
    xpad[i] = v;
}

// ---------------------------------------------------------------------------------------------- finalize
// sequential part of reconstruct_global_body (utils/utils.py:190-198): heading and floor translation in force at every frame
__global__ void k_infill_scan(const float* __restrict__ clip, const double* __restrict__ stats, const double* __restrict__ rot0_pivot, int T,
                              double* __restrict__ ws /* [T][7]: quaternion + translation */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const size_t plane = (size_t)IF_D * T;
    Q rotation{1.0, 0.0, 0.0, 0.0};
    double tr[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < T; ++i) {
        if (i == 0) rotation = q_mul(q_yaxis(-rot0_pivot[0]), rotation);
        double* o = ws + (size_t)i * 7;
        o[0] = rotation.w; o[1] = rotation.x; o[2] = rotation.y; o[3] = rotation.z; o[4] = tr[0]; o[5] = tr[1]; o[6] = tr[2];
        // the de-normalised trajectory is stored through a float32 array by the reference (in-place slice assignment)
        const double rx = (double)(float)((double)clip[plane + i] * stats[2 * IF_D + 1] + stats[2 * IF_D]);
        const double rz = (double)(float)((double)clip[2 * plane + i] * stats[2 * IF_D + 1] + stats[2 * IF_D]);
        const double rr = (double)(float)((double)clip[3 * plane + i] * stats[2 * IF_D + 3] + stats[2 * IF_D + 2]);
        rotation = q_mul(q_yaxis(-rr), rotation);
        const double v[3] = {rx, 0.0, rz};
        double d[3];
        q_rot(rotation, v, d);
        tr[0] += d[0]; tr[1] += d[1]; tr[2] += d[2];
    }
}
__global__ void k_infill_emit(const float* __restrict__ rec_pad, const float* __restrict__ clip, const double* __restrict__ stats,
                              const double* __restrict__ ws, int T, float* __restrict__ markers_rec, float* __restrict__ contact_lbl,
                              float* __restrict__ markers_input) {
    const int Wd = T + 2 * IF_PADT;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T * 4) {                                        // contact labels: sigmoid(rec) > 0.5 (opt_amass_temp.py:268-271)
        const int t = i / 4, p = i - t * 4;
        const float x = rec_pad[(size_t)(IF_D - 4 + p + 1) * Wd + t + IF_PADT];
        const float sg = 1.f / (1.f + expf(-x));
        contact_lbl[i] = sg > 0.5f ? 1.f : 0.f;
    }
    if (i >= T * 67 * 2) return;
    const int which = i / (T * 67), j = i - which * T * 67, t = j / 67, m = j - t * 67;
    if (which == 1 && !markers_input) return;
    const double* w = ws + (size_t)t * 7;
    const Q q{w[0], w[1], w[2], w[3]};
    double v[3], o[3];
    for (int c = 0; c < 3; ++c) {
        const int r = (m + 1) * 3 + c;                      // skip the pelvis rows
        const float raw = which == 0 ? rec_pad[(size_t)(r + 1) * Wd + t + IF_PADT] : clip[(size_t)r * T + t];
        v[c] = (double)(float)((double)raw * stats[IF_D + r] + stats[r]);
    }
    const double sw[3] = {v[0], v[2], v[1]};                // (x, z, y)
    q_rot(q, sw, o);
    o[0] += w[4]; o[2] += w[6];
    float* out = (which == 0 ? markers_rec : markers_input) + ((size_t)t * 67 + m) * 3;
    out[0] = (float)o[0]; out[1] = (float)o[2]; out[2] = (float)o[1];
}


// stand-alone reconstruct_global_body on the packed [T, 1+68+1, 3] layout (zero reference, local pelvis + markers, trajectory)
__global__ void k_rgb_scan(const float* __restrict__ packed, const double* __restrict__ rot0_pivot, int T, double* __restrict__ ws) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Q rotation{1.0, 0.0, 0.0, 0.0};
    double tr[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < T; ++i) {
        if (i == 0) rotation = q_mul(q_yaxis(-rot0_pivot[0]), rotation);
        double* o = ws + (size_t)i * 7;
        o[0] = rotation.w; o[1] = rotation.x; o[2] = rotation.y; o[3] = rotation.z; o[4] = tr[0]; o[5] = tr[1]; o[6] = tr[2];
        const float* root = packed + ((size_t)i * 70 + 69) * 3;             // (root_x, root_z, root_r) = columns 0, 1, 2
        rotation = q_mul(q_yaxis(-(double)root[2]), rotation);
        const double v[3] = {(double)root[0], 0.0, (double)root[1]};
        double d[3];
        q_rot(rotation, v, d);
        tr[0] += d[0]; tr[1] += d[1]; tr[2] += d[2];
    }
}
__global__ void k_rgb_emit(const float* __restrict__ packed, const double* __restrict__ ws, int T, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * 68) return;
    const int t = i / 68, p = i - t * 68;
    const double* w = ws + (size_t)t * 7;
    const Q q{w[0], w[1], w[2], w[3]};
    const float* s = packed + ((size_t)t * 70 + 1 + p) * 3;
    const double sw[3] = {(double)s[0], (double)s[2], (double)s[1]};
    double o[3];
    q_rot(q, sw, o);
    o[0] += w[4]; o[2] += w[6];
    out[(size_t)i * 3] = (float)o[0]; out[(size_t)i * 3 + 1] = (float)o[2]; out[(size_t)i * 3 + 2] = (float)o[1];
}

}  // namespace lemo

using namespace lemo;

extern "C" {

int lemo_repr_local_markers_4chan(const float* body, const float* contact, int32_t T, const double* d_stats, float* repr,
                                  double* rot_0_pivot, double* workspace, void* stream) {
    LEMO_NVTX("lemo_repr_local_markers_4chan");
    LEMO_CHECK(body && contact && repr && rot_0_pivot && workspace && T >= 2, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double* fwd = workspace;                 // [T][3]
    double* rot = workspace + 3 * (size_t)T; // [T][4]
    float* zmin = reinterpret_cast<float*>(workspace + 7 * (size_t)T);
    k_repr_floor_min<<<1, 256, 0, st>>>(body, T * IF_P, zmin);
    k_repr_forward<<<cdiv(T, 128), 128, 0, st>>>(body, zmin, T, fwd);
    k_repr_rotation<<<cdiv(T, 64), 64, 0, st>>>(fwd, T, rot, rot_0_pivot);
    k_repr_emit<<<cdiv((T - 1) * IF_D, 256), 256, 0, st>>>(body, contact, zmin, rot, d_stats, T, repr);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int lemo_reconstruct_global_body(const float* packed, const double* rot_0_pivot, int32_t T, float* out, double* workspace, void* stream) {
    LEMO_CHECK(packed && rot_0_pivot && out && workspace && T >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    k_rgb_scan<<<1, 32, 0, st>>>(packed, rot_0_pivot, T, workspace);
    k_rgb_emit<<<cdiv(T * 68, 256), 256, 0, st>>>(packed, workspace, T, out);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int lemo_infill_prepare_input(const float* clip, int32_t d, int32_t T, float* x_pad, float* row_mask, int32_t* n_rows_selected,
                              void* stream) {
    LEMO_CHECK(clip && x_pad && d == IF_D && T > IF_PADT, "the infill prior is built for d = 208 rows and clips longer than the 8-frame pad");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = 4 * (IF_D + 2) * (T + 2 * IF_PADT);
    k_infill_prepare<<<cdiv(n, 256), 256, 0, st>>>(clip, T, x_pad, row_mask);
    LEMO_CUDA(cudaGetLastError());
    if (n_rows_selected) *n_rows_selected = (IF_D + 2) - 66 - 5;           // 22 masked markers x 3 rows, minus the last five rows
    return 0;
}

int lemo_infill_finalize(const float* rec_pad, const float* clip, const double* d_stats, const double* rot_0_pivot, int32_t d, int32_t T,
                         float* markers_rec, float* contact_lbl, float* markers_input, double* workspace, void* stream) {
    LEMO_NVTX("lemo_infill_finalize");
    LEMO_CHECK(rec_pad && clip && d_stats && rot_0_pivot && markers_rec && contact_lbl && workspace && d == IF_D && T >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    k_infill_scan<<<1, 32, 0, st>>>(clip, d_stats, rot_0_pivot, T, workspace);
    k_infill_emit<<<cdiv(T * 67 * 2, 256), 256, 0, st>>>(rec_pad, clip, d_stats, workspace, T, markers_rec, contact_lbl, markers_input);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
}
