// Parameter block of the backward kernel (shared by abi.cu and render_bwd.cu).
#pragma once
#include "common.cuh"

namespace bts {
struct BwdParams {
  KernelParams k;
  // upstream gradients (nullable), shapes as the forward outputs
  const float* g_rgb;      // (n*R, Vc*Cc)
  const float* g_depth;    // (n*R)
  const float* g_weights;  // (n*R, K)
  const float* g_alphas;   // (n*R, K)
  const float* saved_sigma;  // (n*R, K) from the forward, or null -> sweep 1 re-evaluates the MLP
  // accumulated outputs (caller zero-initialises)
  float* d_G;              // (n, H, W, Hd)
  float* d_g_empty;        // (Hd)
  float* d_w_pe;           // (Hd, P)
  float* d_w_out;          // (Hd)
  float* d_b_out;          // (1)
  float *d_fc0_w, *d_fc0_b, *d_fc1_w, *d_fc1_b;   // block 0 of ResnetFC (n_blocks = 1, d_hidden = 32), nullable
};
}  // namespace bts
