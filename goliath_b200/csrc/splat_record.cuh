// goliath_b200/csrc/splat_record.cuh — the 48-byte blend record of the packed blend path, one definition
// shared by csrc/splat_blend_packed.cu (pack kernels) and csrc/splat_bin_tiles.cu (sort+pack kernel):
//   x y ex ey | A B C opacity | c0 c1 c2 c3
// (ex, ey) = half extents of the axis-aligned box outside which alpha < 1/255 for every pixel.
#pragma once
#include "common.cuh"

namespace gb {

// alpha >= 1/255  <=>  sigma <= log(255 * o) =: s.  Box of the ellipse {sigma <= s}: ex = sqrt(2 s cov_xx).
__device__ __forceinline__ void cull_box(float A, float B, float Cc, float o, float& ex, float& ey) {
  // 2x2 determinant with the cancellation error recovered (Kahan): the box must bound the ellipse of THIS conic
  const float bb = B * B;
  const float det = fmaf(A, Cc, -bb) + fmaf(-B, B, bb);
  const float s = __logf(255.f * o) + 1e-3f;  // margin keeps the box conservative w.r.t. ex2.approx / rounding
  if (!(o >= 0.f) || !(det > 0.f) || !(A > 0.f) || !(Cc > 0.f)) {
    ex = ey = 3.0e38f;  // malformed conic / opacity: never cull, let the per-pixel test decide
  } else if (s <= 0.f) {
    ex = ey = -1.f;     // opacity below 1/255: contributes nowhere
  } else {
    const float inv = 1.f / det;
    ex = fmaf(sqrtf(2.f * s * Cc * inv), 1.0005f, 1e-3f);
    ey = fmaf(sqrtf(2.f * s * A * inv), 1.0005f, 1e-3f);
  }
}

// Fused-render record of Gaussian g: opacity = opacity * compensation and 4th colour channel = view-space
// depth, i.e. exactly what ca_code/utils/render_gsplat.py:72 and :97 feed to the two rasterise calls.
// colors3 == nullptr: the colour quarter is left to a later pass (csrc/splat_bin_tiles.cu gathers it by Gaussian id).
__device__ __forceinline__ void pack_record_fused(int g, const float2* __restrict__ xys,
                                                  const float* __restrict__ conics,
                                                  const float* __restrict__ colors3,
                                                  const float* __restrict__ depths,
                                                  const float* __restrict__ opacity,
                                                  const float* __restrict__ comp, float4* __restrict__ out3) {
  const float2 xy = xys[g];
  const float A = conics[3 * g], B = conics[3 * g + 1], Cc = conics[3 * g + 2];
  const float o = opacity[g] * comp[g];
  float ex, ey;
  cull_box(A, B, Cc, o, ex, ey);
  out3[0] = make_float4(xy.x, xy.y, ex, ey);
  out3[1] = make_float4(A, B, Cc, o);
  if (colors3) out3[2] = make_float4(colors3[3 * g], colors3[3 * g + 1], colors3[3 * g + 2], depths[g]);
}

}  // namespace gb
