#pragma once
#include "common.cuh"
namespace lemo {
struct VPoser {
    int device = 0, maxB = 0;
    float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr, *W3 = nullptr, *b3 = nullptr;   // nn.Linear [out,in]
    float *h1 = nullptr, *h2 = nullptr, *o = nullptr;       // saved activations [B,512],[B,512],[B,126]
    float *d_o = nullptr, *dh2 = nullptr, *dh1 = nullptr;
    // tensor-core path (TF32, generic GEMM of blend_tc.cu): weights rounded once to TF32 (+ transposed copies for the adjoint),
    // activations kept additionally as (hi|lo) splits that are the A operands of the next GEMM
    bool has_tc = false;
    float *W1r = nullptr, *W2r = nullptr, *W3r = nullptr, *W1t = nullptr, *W2t = nullptr, *W3t = nullptr;      // rn_tf32(W)
    float *L1r = nullptr, *L2r = nullptr, *L3r = nullptr, *L1t = nullptr, *L2t = nullptr, *L3t = nullptr;      // W - rn_tf32(W)
    float *zs = nullptr, *h1s = nullptr, *h2s = nullptr, *dos = nullptr, *dh2s = nullptr, *dh1s = nullptr;
    alignas(64) unsigned char m_w1[128], m_w2[128], m_w3[128], m_w1t[128], m_w2t[128], m_w3t[128];
    alignas(64) unsigned char l_w1[128], l_w2[128], l_w3[128], l_w1t[128], l_w2t[128], l_w3t[128];
    alignas(64) unsigned char m_zs[128], m_h1s[128], m_h2s[128], m_dos[128], m_dh2s[128], m_dh1s[128];
};
int vposer_create(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                  int maxB, int device, VPoser** out);
void vposer_free(VPoser* v);
int vposer_decode(VPoser* v, const float* z, int B, float* R_body, float* aa, cudaStream_t st);
int vposer_decode_backward(VPoser* v, const float* z, int B, const float* dR_body, float* dz, cudaStream_t st);
}  // namespace lemo
