// Launch parameters of the patch photometric loss kernels (patch_loss.cu).
#pragma once
#include <cstdint>

namespace bts {

struct PatchLossParams {
  const float* pred;          // (P, h, w, nv, c)
  const float* gt;            // (P, h, w, c)
  const float* invalid_ray;   // (P, h, w) or null
  const float* invalid_mass;  // (P, h, w, nv) or null: invalid = all_v(mass_v > mass_thresh)   (loss.py:107)
  float mass_thresh;
  int64_t P;
  int h, w, nv, c;
  float* per_ray;             // (P, h, w)   min_v(err) * (1 - invalid)
  int32_t* argmin;            // (P, h, w)   view that attains the minimum
  float* errors;              // (P, h, w, nv) or null: the un-reduced criterion (loss.py:154)
  // backward
  const float* g_per_ray;     // (P, h, w) upstream gradient of per_ray
  float* d_pred;              // (P, h, w, nv, c)
};

}  // namespace bts
