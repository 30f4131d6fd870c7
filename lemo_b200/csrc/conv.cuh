// 3x3 convolution stack of the motion priors (reference models/AE_sep.py, models/AE.py) on sm_100a.
//
// Activation layout ("padded pitch-linear planes"): a [N,C,H,W] tensor is stored as [N][C][Hp*Wp] with
// Hp = H+2, Wp = roundup(W+1, 8); pixel (y,x) lives at (y+1)*Wp + (x+1) and every border element is 0.  ONE pad column is enough: in the
// linear pixel order the right neighbour of a row's last pixel is the next row's column 0, which is that row's (zero) left pad.  (Round 1
// used roundup(W+2, 8) = 144 for the 135-column velocity image; 136 computes 5.6 % fewer padded pixels per layer.)
// With the zero border physically present a 3x3/pad-1 convolution is a pure 1-D correlation over the
// linear index  out[q] = sum in[q + (ky-1)*Wp + (kx-1)] * w[ky][kx], so CTAs tile the LINEAR pixel range
// in 256-pixel tiles: no 2-D tile-edge waste on the 245x134 (or any T) image, 16-byte aligned vector
// smem reads for every (ky,kx) shift, coalesced 512-byte output rows.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include <vector>

namespace lemo {

struct PlaneGeom {
    int H, W, Hp, Wp, PS;   // PS = Hp*Wp
};
inline PlaneGeom make_geom(int H, int W) {
    PlaneGeom g;
    static const int pad = []() { const char* e = getenv("LEMO_PLANE_PAD"); return (e && e[0] == '2') ? 2 : 1; }();   // 2 = the round-1 layout (A/B measurements)
    g.H = H; g.W = W; g.Hp = H + 2; g.Wp = (W + pad + 7) / 8 * 8; g.PS = g.Hp * g.Wp;
    return g;
}

enum ConvEpi { EPI_BIAS_LRELU = 0, EPI_MASK = 1, EPI_BIAS = 2, EPI_NONE = 3 };

// out[n][oc] = epi( sum_ic in[n][ic] (*) wk[ic][ky][kx][oc] ),  wk is [Cin][9][Cout] (oc fastest)
// scratch (optional): split-K workspace of >= conv3x3_splitk_floats(...) floats; with it, layers whose plane is too small to fill the
// GPU are contracted in input-channel slices by separate CTAs and summed in slice order by a second kernel (deterministic).
int conv3x3_launch(const float* in, const float* wk, const float* bias, const float* aux, float* out,
                   int N, int Cin, int Cout, const PlaneGeom& g, ConvEpi epi, cudaStream_t st, float* scratch = nullptr,
                   size_t scratch_floats = 0);
int conv3x3_splitk(int N, int Cin, int Cout, const PlaneGeom& g);
size_t conv3x3_splitk_floats(int N, int Cin, int Cout, const PlaneGeom& g);
// dW[oc][ic][ky][kx] (+)= sum_{n,pix} dpre[n][oc][q] * in[n][ic][q + shift];  db[oc] (+)= sum dpre
// scratch: >= conv3x3_wgrad_floats(...) floats of partials (fixed-order reduction; dW / db are overwritten, not accumulated)
int conv3x3_wgrad_launch(const float* in, const float* dpre, float* dW_oikk, float* db, int N, int Cin, int Cout,
                         const PlaneGeom& g, bool transpose_io, cudaStream_t st, float* scratch, size_t scratch_floats);
size_t conv3x3_wgrad_floats(int N, int Cin, int Cout, const PlaneGeom& g);

int pack_planes(const float* dense, float* planes, int NC, const PlaneGeom& g, cudaStream_t st);
int unpack_planes(const float* planes, float* dense, int NC, const PlaneGeom& g, cudaStream_t st);

struct ConvLayer {
    int Cin = 0, Cout = 0;
    bool transposed = false;   // nn.ConvTranspose2d weight layout [Cin,Cout,3,3]
    long long w_off = 0, b_off = 0;   // offsets into the flat state_dict-ordered weight vector
    float* wk_f = nullptr;     // forward kernel weights   [Cin][9][Cout]
    float* wk_b = nullptr;     // input-gradient weights   [Cout][9][Cin]
};

struct ConvNet {
    int device = 0;
    int kind = 0;              // 0 Enc (AE_sep, no pooling), 1 AE
    int in_ch = 1, maxN = 0;
    bool with_backward = false;
    long long n_weights = 0;
    float* w_flat = nullptr;   // state_dict order (the tensor Adam updates during the AE fine-tune)
    std::vector<ConvLayer> layers;
    std::vector<PlaneGeom> geom;       // geometry per resolution level (Enc: 1 level; AE: 6 levels)
    std::vector<float*> act;           // saved activations (planes) per layer output, act[0] = packed input
    std::vector<float*> grad;          // gradient planes (two ping-pong buffers per level)
    std::vector<float*> pool_in;       // AE: pre-pool activations
    std::vector<int*> pool_idx;        // AE: argmax indices
    std::vector<float*> up;            // AE: zero-upsampled planes
    float* d_wflat = nullptr;
    float* sk_scratch = nullptr;       // split-K workspace of the small-plane layers (AE)
    size_t sk_floats = 0;
    float* wg_scratch = nullptr;       // weight-gradient partials (fixed-order reduction instead of float atomics)
    size_t wg_floats = 0;
    // AE backward: the weight gradients of a layer run on two side streams beside the input-gradient chain (fork/join with events; captured
    // into the fine-tune graph as parallel branches); three rotating gradient buffers give every weight gradient two layers of slack
    float* wg_scratch2 = nullptr;
    std::vector<float*> glev;          // AE: AE_GB gradient buffers per resolution level (zeroed once; producers write interiors only, so no per-step clears)
    void* bw_side[2] = {nullptr, nullptr};
    void* bw_ready = nullptr;
    void* bw_done[2] = {nullptr, nullptr};
    // fine-tune driver (lemo_ae_finetune_run): device-side step schedule + one captured step
    void *ft_sched = nullptr, *ft_graph = nullptr, *ft_gexec = nullptr, *ft_stream = nullptr, *ft_ev_in = nullptr, *ft_ev_out = nullptr;
    const void *ft_x = nullptr, *ft_mask = nullptr;
    float* ft_losses = nullptr;
    int ft_N = 0, ft_rows = 0;
    long long launches = 0;
    void* tc = nullptr;                // EncTC (conv_tc.cu): tensor-core path state, kind 0 only
};

int convnet_create(int kind, int in_ch, const float* h_weights, long long n_weights, int maxN, int H, int W,
                   bool with_backward, int device, ConvNet** out);
void convnet_free(ConvNet* n);
int convnet_refresh_weights(ConvNet* n, cudaStream_t st);      // rebuild wk_f / wk_b from w_flat
// Enc on planes: x_planes [N][1][PS] -> act.back() [N][64][PS]
int enc_forward_planes(ConvNet* n, const float* x_planes, int N, cudaStream_t st);
// dpre of the last layer must be in n->grad[0] (already multiplied by LeakyReLU'(z)); result dx_planes [N][1][PS]
int enc_backward_planes(ConvNet* n, int N, float* dx_planes, cudaStream_t st);
// tensor-core Enc path (conv_tc.cu)
bool conv_tc_enabled();
void conv_tc_set(int on);
int enc_tc_create(ConvNet* n);
void enc_tc_free(ConvNet* n);
int enc_tc_refresh_weights(ConvNet* n, cudaStream_t st);
int enc_tc_forward(ConvNet* n, const float* x_planes, int N, cudaStream_t st);
int enc_tc_smooth_loss(ConvNet* n, int N, float w, int acc_stride, int acc_slot, float* acc, cudaStream_t st);
int enc_tc_backward(ConvNet* n, int N, float* dx_planes, cudaStream_t st);
int enc_tc_profile_layer(ConvNet* n, int layer, int N, int backward, int reps, cudaStream_t st);
int enc_tc_unpack_z(ConvNet* n, int N, float* z_dense, cudaStream_t st);
int enc_tc_pack_dz(ConvNet* n, int N, const float* dz_dense, cudaStream_t st);
inline bool enc_uses_tc(const ConvNet* n) { return n->tc != nullptr && conv_tc_enabled(); }
inline float* enc_z_planes(ConvNet* n) { return n->act.back(); }
inline float* enc_gz_planes(ConvNet* n) { return n->grad[0]; }

}  // namespace lemo
