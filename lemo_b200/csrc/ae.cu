// AE (motion-infilling prior, reference models/AE.py:79-108): placeholder translation unit, filled in below.
#include "handles.cuh"
#include "../../include/lemo_b200.h"
namespace lemo {
int ae_create(int in_ch, const float* h_weights, long long n_weights, int maxN, int H, int W, bool with_backward, int device, ConvNet** out) {
    (void)in_ch; (void)h_weights; (void)n_weights; (void)maxN; (void)H; (void)W; (void)with_backward; (void)device; (void)out;
    set_error("AE convnet not built in this revision");
    return 3;
}
}
extern "C" {
int lemo_ae_forward(LemoConvNet* net, const float* x, int32_t N, float* rec, float* z, void* stream) {
    (void)net; (void)x; (void)N; (void)rec; (void)z; (void)stream;
    lemo::set_error("AE convnet not built in this revision");
    return 3;
}
int lemo_ae_backward_weights(LemoConvNet* net, const float* d_rec, int32_t N, float* d_weights, void* stream) {
    (void)net; (void)d_rec; (void)N; (void)d_weights; (void)stream;
    lemo::set_error("AE convnet not built in this revision");
    return 3;
}
}
