#pragma once
#include "common.cuh"
namespace lemo {
struct VPoser {
    int device = 0, maxB = 0;
    float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr, *W3 = nullptr, *b3 = nullptr;   // nn.Linear [out,in]
    float *h1 = nullptr, *h2 = nullptr, *o = nullptr;       // saved activations [B,512],[B,512],[B,126]
    float *d_o = nullptr, *dh2 = nullptr, *dh1 = nullptr;
};
int vposer_create(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                  int maxB, int device, VPoser** out);
void vposer_free(VPoser* v);
int vposer_decode(VPoser* v, const float* z, int B, float* R_body, float* aa, cudaStream_t st);
int vposer_decode_backward(VPoser* v, const float* z, int B, const float* dR_body, float* dz, cudaStream_t st);
}  // namespace lemo
