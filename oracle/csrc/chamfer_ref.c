/* oracle/csrc/chamfer_ref.c -- CPU restatement of the Chamfer nearest-neighbour search.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows the call contract of /root/reference/temp_prox/dist_chamfer.py:10-28 around the external `chamfer` CUDA extension
 * (ChamferDistancePytorch@719b0f1, chamfer.cu NmDistanceKernel -- third-party, absent from /root/reference: parity unpinned,
 * restated from its published algorithm): for every point of cloud 1 the squared L2 distance to, and the int32 index of, its
 * nearest point in cloud 2; the scan runs over ascending target index with a strict `<`, so the FIRST minimum wins.
 *
 * Arithmetic pinned here and mirrored by lemo_b200/csrc/chamfer.cu: the distance is evaluated as
 *      d = fma(dz, dz, fma(dy, dy, dx * dx)),   dx = q.x - p.x  (single-rounded fp32 throughout)
 * which is what nvcc's default -fmad=true contraction makes of the reference kernel's `dx*dx + dy*dy + dz*dz`.
 * Built with -ffp-contract=off so the compiler adds no contraction of its own; fmaf() is exact by C99.
 *
 * Single-threaded C (this image's gcc has no libgomp): oracle/ref_chamfer.py fans batches out over host threads.
 * xyz2_batch_stride (in floats) = 0 means one scene shared by the whole batch (fitting_temp_slide.py:748 repeats it B times).
 */
#include <math.h>
#include <stdint.h>

void chamfer_ref_nn(const float* q, int64_t q_bs, int32_t nq, const float* t, int64_t t_bs, int32_t nt, int32_t B, float* dist,
                    int32_t* idx) {
    for (int32_t b = 0; b < B; ++b)
        for (int32_t i = 0; i < nq; ++i) {
            const float* qp = q + (int64_t)b * q_bs + (int64_t)i * 3;
            const float* tb = t + (int64_t)b * t_bs;
            const float qx = qp[0], qy = qp[1], qz = qp[2];
            float best = 3.4e38f;
            int32_t bi = 0;
            for (int32_t j = 0; j < nt; ++j) {
                const float dx = qx - tb[3 * j], dy = qy - tb[3 * j + 1], dz = qz - tb[3 * j + 2];
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d < best) { best = d; bi = j; }
            }
            dist[(int64_t)b * nq + i] = best;
            idx[(int64_t)b * nq + i] = bi;
        }
}
