// goliath_b200/csrc/deconv_wnub.cu — stride-2 4x4 transposed convolution with weight-norm scale, untied
// (per-pixel) bias and LeakyReLU fused into the epilogue (sm_100a), forward.
//
// Replaces, for the layer type every RGCA / hand-MVP decoder tower is made of
// (make_conv_trans(.., 4, 2, 1, "wn", LeakyReLU(0.2), ub=(H,W)), ca_code/models/rgca.py:408-456,
// ca_code/nn/layers.py:27-47), the reference's three kernels per layer: cuDNN conv_transpose2d
// (layers.py:380-391), the `output + bias[None]` add (layers.py:392-396) and the in-place LeakyReLU, plus the
// weight-norm reparametrisation w = g * v / ||v||_F (whole-tensor norm, layers.py:200-204), which is
// folded into a per-output-channel scale applied to the accumulator.
//
// Round-1 implementation: fp32 SIMT with the sub-pixel (4-phase) decomposition — each thread owns one input
// position, i.e. a 2x2 output quad, for 8 output channels (32 accumulators); the 3x3 input neighbourhood and the
// 8x16 weights of the current input channel come from shared memory (weights as broadcast float4).  The tcgen05
// phase-GEMM version (TF32x3, TMA-fed) is the planned successor; the last layers are bound by the
// untied-bias + output traffic (1.07 GB for 16->125 @1024^2), not by FLOPs.
#include "common.cuh"

namespace {

constexpr int TQ = 16;        // quads (input positions) per CTA edge -> 32x32 output tile
constexpr int CO_T = 8;       // output channels per CTA pass
constexpr int CI_CHUNK = 8;   // input channels staged per step
constexpr int HALO = TQ + 2;

__global__ void __launch_bounds__(TQ* TQ) deconv4x4s2_fwd_kernel(
    int Cin, int Cout, int Hi, int Wi, const float* __restrict__ x /* [B,Cin,Hi,Wi] */,
    const float* __restrict__ v /* [Cin,Cout,4,4] */, const float* __restrict__ scale /* [Cout] */,
    const float* __restrict__ bias /* [Cout,2Hi,2Wi] or null */, float slope, int apply_act,
    float* __restrict__ out /* [B,Cout,2Hi,2Wi] */) {
  __shared__ float s_x[CI_CHUNK][HALO][HALO + 1];
  __shared__ __align__(16) float s_w[CI_CHUNK][CO_T][16];
  const int tiles_x = (Wi + TQ - 1) / TQ;
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * TQ, tx0 = (tile % tiles_x) * TQ;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;
  const int tid = threadIdx.x;
  const int qy = tid / TQ, qx = tid % TQ;
  const int m = ty0 + qy, n = tx0 + qx;  // input position / output quad
  const int Ho = 2 * Hi, Wo = 2 * Wi;

  float acc[CO_T][4];
#pragma unroll
  for (int c = 0; c < CO_T; ++c) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; }

  const float* xb = x + (size_t)b * Cin * Hi * Wi;
  for (int ci0 = 0; ci0 < Cin; ci0 += CI_CHUNK) {
    __syncthreads();
    // stage the input tile with a 1-pixel halo (zeros outside the image / beyond Cin)
    for (int i = tid; i < CI_CHUNK * HALO * HALO; i += TQ * TQ) {
      const int ci = i / (HALO * HALO), r = (i / HALO) % HALO, c = i % HALO;
      const int yy = ty0 - 1 + r, xx = tx0 - 1 + c;
      float val = 0.f;
      if (ci0 + ci < Cin && yy >= 0 && yy < Hi && xx >= 0 && xx < Wi) val = xb[((size_t)(ci0 + ci) * Hi + yy) * Wi + xx];
      s_x[ci][r][c] = val;
    }
    // stage the weights of (ci chunk) x (co block): v[ci][co][ky][kx]
    for (int i = tid; i < CI_CHUNK * CO_T * 16; i += TQ * TQ) {
      const int ci = i / (CO_T * 16), co = (i / 16) % CO_T, k = i % 16;
      float val = 0.f;
      if (ci0 + ci < Cin && co0 + co < Cout) val = v[((size_t)(ci0 + ci) * Cout + (co0 + co)) * 16 + k];
      s_w[ci][co][k] = val;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CI_CHUNK; ++ci) {
      // 3x3 neighbourhood of the input position (rows m-1..m+1, cols n-1..n+1)
      const float a00 = s_x[ci][qy][qx], a01 = s_x[ci][qy][qx + 1], a02 = s_x[ci][qy][qx + 2];
      const float a10 = s_x[ci][qy + 1][qx], a11 = s_x[ci][qy + 1][qx + 1], a12 = s_x[ci][qy + 1][qx + 2];
      const float a20 = s_x[ci][qy + 2][qx], a21 = s_x[ci][qy + 2][qx + 1], a22 = s_x[ci][qy + 2][qx + 2];
#pragma unroll
      for (int c = 0; c < CO_T; ++c) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[ci][c][0]);   // ky = 0: kx 0..3
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[ci][c][4]);   // ky = 1
        const float4 w2 = *reinterpret_cast<const float4*>(&s_w[ci][c][8]);   // ky = 2
        const float4 w3 = *reinterpret_cast<const float4*>(&s_w[ci][c][12]);  // ky = 3
        // out(2m  ,2n  ) = x(m,n) w11 + x(m,n-1) w13 + x(m-1,n) w31 + x(m-1,n-1) w33
        acc[c][0] += a11 * w1.y + a10 * w1.w + a01 * w3.y + a00 * w3.w;
        // out(2m  ,2n+1) = x(m,n+1) w10 + x(m,n) w12 + x(m-1,n+1) w30 + x(m-1,n) w32
        acc[c][1] += a12 * w1.x + a11 * w1.z + a02 * w3.x + a01 * w3.z;
        // out(2m+1,2n  ) = x(m+1,n) w01 + x(m+1,n-1) w03 + x(m,n) w21 + x(m,n-1) w23
        acc[c][2] += a21 * w0.y + a20 * w0.w + a11 * w2.y + a10 * w2.w;
        // out(2m+1,2n+1) = x(m+1,n+1) w00 + x(m+1,n) w02 + x(m,n+1) w20 + x(m,n) w22
        acc[c][3] += a22 * w0.x + a21 * w0.z + a12 * w2.x + a11 * w2.z;
      }
    }
  }
  if (m >= Hi || n >= Wi) return;
#pragma unroll
  for (int c = 0; c < CO_T; ++c) {
    const int co = co0 + c;
    if (co >= Cout) break;
    const float sc = scale[co];
    float o00 = acc[c][0] * sc, o01 = acc[c][1] * sc, o10 = acc[c][2] * sc, o11 = acc[c][3] * sc;
    const size_t row0 = ((size_t)co * Ho + 2 * m) * Wo + 2 * n;
    if (bias) {
      const float2 b0 = *reinterpret_cast<const float2*>(bias + row0);
      const float2 b1 = *reinterpret_cast<const float2*>(bias + row0 + Wo);
      o00 += b0.x; o01 += b0.y; o10 += b1.x; o11 += b1.y;
    }
    if (apply_act) {
      o00 = o00 > 0.f ? o00 : o00 * slope; o01 = o01 > 0.f ? o01 : o01 * slope;
      o10 = o10 > 0.f ? o10 : o10 * slope; o11 = o11 > 0.f ? o11 : o11 * slope;
    }
    float* ob = out + (size_t)b * Cout * Ho * Wo;
    *reinterpret_cast<float2*>(ob + row0) = make_float2(o00, o01);
    *reinterpret_cast<float2*>(ob + row0 + Wo) = make_float2(o10, o11);
  }
}

}  // namespace

// Fused ConvTranspose2dWNUB(k=4, s=2, p=1) [+ LeakyReLU] forward.  x [B,Cin,Hi,Wi], v = weight_v [Cin,Cout,4,4],
// scale [Cout] = weight_g / ||weight_v||_F, bias [Cout,2Hi,2Wi] or NULL, out [B,Cout,2Hi,2Wi].
// Replaces layers.py:380-396 (+ the activation that follows it in make_conv_trans, layers.py:27-47).
GB_API int gb_deconv4x4s2_wnub_fwd(int B, int Cin, int Cout, int Hi, int Wi, const float* x, const float* v,
                                   const float* scale, const float* bias, float slope, int apply_act, float* out,
                                   void* stream) {
  if (B <= 0 || Cin <= 0 || Cout <= 0 || Hi <= 0 || Wi <= 0) return 0;
  const int tiles = gb::cdiv(Hi, TQ) * gb::cdiv(Wi, TQ);
  dim3 grid(tiles, gb::cdiv(Cout, CO_T), B);
  deconv4x4s2_fwd_kernel<<<grid, TQ * TQ, 0, (cudaStream_t)stream>>>(Cin, Cout, Hi, Wi, x, v, scale, bias, slope, apply_act, out);
  gb::count_launches(1);
  GB_CHECK_LAUNCH();
  return 0;
}
