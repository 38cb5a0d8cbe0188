/*
 * oracle/splat_oracle.c — CPU restatement of the Gaussian-splat path RGCA renders through.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under goliath_b200/ may import, link or execute this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party dependency
 * gsplat==0.1.11 (reference requirements.txt:7), which is NOT under /root/reference, not installed
 * and not fetchable.  This file restates the published gsplat v0.1.11 algorithm
 * (gsplat/cuda/csrc/{forward,backward}.cu + helpers.cuh, gsplat/{project_gaussians,rasterize,
 * utils}.py) anchored on the reference's own call sites:
 *   ca_code/utils/render_gsplat.py:49-63   project_gaussians(means3D, scales, 1.0, quats, Rt, fx,fy,cx,cy,H,W,16,z_near)
 *   ca_code/utils/render_gsplat.py:65-78   rasterize_gaussians(..., colors, opacity*compensation, H, W, 16, bg, return_alpha=True)
 *   ca_code/utils/render_gsplat.py:90-104  second rasterize with depth as colour
 *   ca_code/models/rgca.py:112-151         per-view loop, depth / alpha.clamp(0.05,1)
 * and on SURVEY.md Appendix A (every constant there is a contract item).
 *
 * Arithmetic contract (so that tile binning can be compared BIT-EXACTLY with the CUDA path):
 * projection is evaluated in IEEE fp32, one rounding per operation, no FMA contraction, in exactly
 * the operation order written here (compile with -ffp-contract=off; the CUDA projection kernel is
 * compiled with -fmad=false and no fast-math and follows the same order).  The blend uses libm
 * expf; the CUDA path uses the hardware ex2 approximation like the reference (`__expf`), so pixels
 * and gradients are compared to 1e-4 relative, not bitwise.
 *
 * Layouts: all arrays contiguous fp32 / int32 / int64 exactly as the reference passes them.
 * quats are (w,x,y,z).  viewmat = first 12 floats of a row-major [R|t] (reference passes [3,4]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---------- helpers -------------------------------------------------------------------------- */

static inline int f2i_sat(float v) { /* CUDA cvt.rzi.s32.f32 semantics: saturate, NaN -> 0 */
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int)v;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* row-major 3x3 product, left-to-right accumulation */
static inline void mm3(const float a[9], const float b[9], float o[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      o[r * 3 + c] = a[r * 3 + 0] * b[0 * 3 + c] + a[r * 3 + 1] * b[1 * 3 + c] + a[r * 3 + 2] * b[2 * 3 + c];
}
static inline void tr3(const float a[9], float o[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[c * 3 + r];
}

/* gsplat helpers.cuh quat_to_rotmat: normalises internally; returns row-major R */
static inline void quat_to_rotmat(const float q[4], float R[9]) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  float n = sqrtf(w * w + x * x + y * y + z * z);
  float s = 1.0f / n;
  w = w * s; x = x * s; y = y * s; z = z * s;
  R[0] = 1.f - 2.f * (y * y + z * z);
  R[1] = 2.f * (x * y - w * z);
  R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z);
  R[4] = 1.f - 2.f * (x * x + z * z);
  R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y);
  R[7] = 2.f * (y * z + w * x);
  R[8] = 1.f - 2.f * (x * x + y * y);
}

/* tile bbox (helpers.cuh get_tile_bbox / get_bbox): inclusive min, exclusive max, in tile units */
static inline void tile_bbox(float cx, float cy, float radius, int tbx, int tby, int bw,
                             int* x0, int* y0, int* x1, int* y1) {
  float fb = (float)bw;
  float tcx = cx / fb, tcy = cy / fb, tr = radius / fb;
  *x0 = imin(imax(0, f2i_sat(tcx - tr)), tbx);
  *x1 = imin(imax(0, f2i_sat(tcx + tr + 1.f)), tbx);
  *y0 = imin(imax(0, f2i_sat(tcy - tr)), tby);
  *y1 = imin(imax(0, f2i_sat(tcy + tr + 1.f)), tby);
}

/* ---------- project_gaussians forward (Appendix A steps 1-6) --------------------------------- */

ORC_API void orc_project_fwd(int G, const float* means3d, const float* scales, float glob_scale,
                             const float* quats, const float* viewmat, float fx, float fy, float cx,
                             float cy, int img_h, int img_w, int block_width, float clip_thresh,
                             float* cov3d, float* xys, float* depths, int32_t* radii, float* conics,
                             float* compensation, int32_t* num_tiles_hit) {
  const int tbx = (img_w + block_width - 1) / block_width;
  const int tby = (img_h + block_width - 1) / block_width;
  const float tan_fovx = 0.5f * (float)img_w / fx;
  const float tan_fovy = 0.5f * (float)img_h / fy;
  const float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
  const float* V = viewmat;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < G; ++i) {
    /* outputs are pre-zeroed by the Python side; culled Gaussians keep zeros */
    radii[i] = 0; num_tiles_hit[i] = 0;
    xys[2 * i] = xys[2 * i + 1] = 0.f; depths[i] = 0.f; compensation[i] = 0.f;
    conics[3 * i] = conics[3 * i + 1] = conics[3 * i + 2] = 0.f;
    for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = 0.f;

    const float px = means3d[3 * i], py = means3d[3 * i + 1], pz = means3d[3 * i + 2];
    /* 1. view transform + near clip */
    const float vx = V[0] * px + V[1] * py + V[2] * pz + V[3];
    const float vy = V[4] * px + V[5] * py + V[6] * pz + V[7];
    const float vz = V[8] * px + V[9] * py + V[10] * pz + V[11];
    if (vz <= clip_thresh) continue;

    /* 2. cov3d = M M^T, M = R(q) diag(glob_scale * s) */
    float R[9]; quat_to_rotmat(quats + 4 * i, R);
    const float sx = glob_scale * scales[3 * i], sy = glob_scale * scales[3 * i + 1], sz = glob_scale * scales[3 * i + 2];
    float M[9] = {R[0] * sx, R[1] * sy, R[2] * sz, R[3] * sx, R[4] * sy, R[5] * sz, R[6] * sx, R[7] * sy, R[8] * sz};
    float Mt[9], S3[9]; tr3(M, Mt); mm3(M, Mt, S3);
    float* c3 = cov3d + 6 * i;
    c3[0] = S3[0]; c3[1] = S3[1]; c3[2] = S3[2]; c3[3] = S3[4]; c3[4] = S3[5]; c3[5] = S3[8];

    /* 3. EWA: clamp x/z,y/z to 1.3*tan_fov; J; cov2d = (J W) cov3d (J W)^T */
    const float tz = vz;
    float tx = tz * fminf(lim_x, fmaxf(-lim_x, vx / tz));
    float ty = tz * fminf(lim_y, fmaxf(-lim_y, vy / tz));
    const float rz = 1.f / tz, rz2 = rz * rz;
    const float j00 = fx * rz, j02 = -fx * tx * rz2, j11 = fy * rz, j12 = -fy * ty * rz2;
    /* T = J W, rows 0,1 only (row 2 of J is zero) */
    const float t00 = j00 * V[0] + j02 * V[8], t01 = j00 * V[1] + j02 * V[9], t02 = j00 * V[2] + j02 * V[10];
    const float t10 = j11 * V[4] + j12 * V[8], t11 = j11 * V[5] + j12 * V[9], t12 = j11 * V[6] + j12 * V[10];
    /* TV = T * cov3d (symmetric V) */
    const float a0 = t00 * c3[0] + t01 * c3[1] + t02 * c3[2];
    const float a1 = t00 * c3[1] + t01 * c3[3] + t02 * c3[4];
    const float a2 = t00 * c3[2] + t01 * c3[4] + t02 * c3[5];
    const float b0 = t10 * c3[0] + t11 * c3[1] + t12 * c3[2];
    const float b1 = t10 * c3[1] + t11 * c3[3] + t12 * c3[4];
    const float b2 = t10 * c3[2] + t11 * c3[4] + t12 * c3[5];
    const float c00 = a0 * t00 + a1 * t01 + a2 * t02;
    const float c01 = a0 * t10 + a1 * t11 + a2 * t12;
    const float c11 = b0 * t10 + b1 * t11 + b2 * t12;

    /* 4. blur + compensation */
    const float det0 = c00 * c11 - c01 * c01;
    const float A = c00 + 0.3f, B = c01, C = c11 + 0.3f;
    const float det1 = A * C - B * B;
    const float comp = sqrtf(fmaxf(0.f, det0 / det1));

    /* 5. conic + radius */
    if (det1 == 0.f) continue;
    const float inv_det = 1.f / det1;
    const float q0 = C * inv_det, q1 = -B * inv_det, q2 = A * inv_det;
    const float mid = 0.5f * (A + C);
    const float disc = sqrtf(fmaxf(0.1f, mid * mid - det1));
    const float v1 = mid + disc, v2 = mid - disc;
    const float radius = ceilf(3.f * sqrtf(fmaxf(v1, v2)));
    conics[3 * i] = q0; conics[3 * i + 1] = q1; conics[3 * i + 2] = q2;

    /* 6. pixel centre + tile bbox */
    const float rw = 1.f / (vz + 1e-6f);
    const float ctr_x = vx * rw * fx + cx, ctr_y = vy * rw * fy + cy;
    int x0, y0, x1, y1; tile_bbox(ctr_x, ctr_y, radius, tbx, tby, block_width, &x0, &y0, &x1, &y1);
    const int area = (x1 - x0) * (y1 - y0);
    if (area <= 0) continue;
    num_tiles_hit[i] = area; depths[i] = vz; radii[i] = f2i_sat(radius);
    xys[2 * i] = ctr_x; xys[2 * i + 1] = ctr_y; compensation[i] = comp;
  }
}

/* ---------- project_gaussians backward -------------------------------------------------------
 * gsplat backward.cu project_gaussians_backward_kernel: Gaussians with radii<=0 get zero grads.
 * Quirks kept: the 1.3*tan_fov clamp is NOT applied in the backward; quaternion gradient is taken
 * w.r.t. the normalised quaternion without projecting through the normalisation. */
ORC_API void orc_project_bwd(int G, const float* means3d, const float* scales, float glob_scale,
                             const float* quats, const float* viewmat, float fx, float fy,
                             const float* cov3d, const int32_t* radii, const float* conics,
                             const float* compensation, const float* v_xy, const float* v_depth,
                             const float* v_conic, const float* v_compensation, float* v_cov2d,
                             float* v_cov3d, float* v_mean3d, float* v_scale, float* v_quat) {
  const float* V = viewmat;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < G; ++i) {
    for (int k = 0; k < 3; ++k) { v_cov2d[3 * i + k] = 0.f; v_mean3d[3 * i + k] = 0.f; v_scale[3 * i + k] = 0.f; }
    for (int k = 0; k < 6; ++k) v_cov3d[6 * i + k] = 0.f;
    for (int k = 0; k < 4; ++k) v_quat[4 * i + k] = 0.f;
    if (radii[i] <= 0) continue;
    const float px = means3d[3 * i], py = means3d[3 * i + 1], pz = means3d[3 * i + 2];
    const float vx = V[0] * px + V[1] * py + V[2] * pz + V[3];
    const float vy = V[4] * px + V[5] * py + V[6] * pz + V[7];
    const float vz = V[8] * px + V[9] * py + V[10] * pz + V[11];
    /* project_pix_vjp then R^T */
    const float rw = 1.f / (vz + 1e-6f);
    const float vpx = fx * v_xy[2 * i], vpy = fy * v_xy[2 * i + 1];
    const float gvx = vpx * rw, gvy = vpy * rw, gvz = -(vpx * vx + vpy * vy) * rw * rw;
    float vm[3];
    vm[0] = V[0] * gvx + V[4] * gvy + V[8] * gvz;
    vm[1] = V[1] * gvx + V[5] * gvy + V[9] * gvz;
    vm[2] = V[2] * gvx + V[6] * gvy + V[10] * gvz;
    const float vzg = v_depth[i];
    vm[0] += V[8] * vzg; vm[1] += V[9] * vzg; vm[2] += V[10] * vzg;

    /* conic -> cov2d vjp: v_Sigma = -X G X */
    const float X00 = conics[3 * i], X01 = conics[3 * i + 1], X11 = conics[3 * i + 2];
    const float G00 = v_conic[3 * i], G01 = 0.5f * v_conic[3 * i + 1], G11 = v_conic[3 * i + 2];
    /* XG */
    const float xg00 = X00 * G00 + X01 * G01, xg01 = X00 * G01 + X01 * G11;
    const float xg10 = X01 * G00 + X11 * G01, xg11 = X01 * G01 + X11 * G11;
    const float s00 = -(xg00 * X00 + xg01 * X01), s01 = -(xg00 * X01 + xg01 * X11);
    const float s10 = -(xg10 * X00 + xg11 * X01), s11 = -(xg10 * X01 + xg11 * X11);
    float vc2[3] = {s00, s01 + s10, s11};
    /* compensation vjp */
    {
      const float comp = compensation[i];
      const float inv_det = X00 * X11 - X01 * X01;
      const float omc = 1.f - comp * comp;
      const float vsq = v_compensation[i] * 0.5f / (comp + 1e-6f);
      vc2[0] += vsq * (omc * X00 - 0.3f * inv_det);
      vc2[1] += 2.f * vsq * (omc * X01);
      vc2[2] += vsq * (omc * X11 - 0.3f * inv_det);
    }
    v_cov2d[3 * i] = vc2[0]; v_cov2d[3 * i + 1] = vc2[1]; v_cov2d[3 * i + 2] = vc2[2];

    /* project_cov3d_ewa_vjp (no fov clamp) */
    const float* c3 = cov3d + 6 * i;
    const float tx = vx, ty = vy, tz = vz;
    const float rz = 1.f / tz, rz2 = rz * rz, rz3 = rz2 * rz;
    float W[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]}; /* row-major rotation */
    float J[9] = {fx * rz, 0.f, -fx * tx * rz2, 0.f, fy * rz, -fy * ty * rz2, 0.f, 0.f, 0.f};
    float T[9]; mm3(J, W, T);
    float Vm[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
    float Gc[9] = {vc2[0], 0.5f * vc2[1], 0.f, 0.5f * vc2[1], vc2[2], 0.f, 0.f, 0.f, 0.f};
    float Tt[9], tmp[9], vV[9], vT[9], tmp2[9];
    tr3(T, Tt);
    mm3(Tt, Gc, tmp); mm3(tmp, T, vV); /* v_V = T^T G T */
    mm3(Gc, T, tmp); mm3(tmp, Vm, vT); /* G T V^T (V symmetric) */
    mm3(Gc, T, tmp); mm3(tmp, Vm, tmp2); /* G^T T V == same since G symmetric */
    for (int k = 0; k < 9; ++k) vT[k] = vT[k] + tmp2[k];
    float* vc3 = v_cov3d + 6 * i;
    vc3[0] = vV[0]; vc3[1] = vV[1] + vV[3]; vc3[2] = vV[2] + vV[6];
    vc3[3] = vV[4]; vc3[4] = vV[5] + vV[7]; vc3[5] = vV[8];
    /* v_J = v_T W^T ; only rows 0,1 matter */
    float Wt[9], vJ[9]; tr3(W, Wt); mm3(vT, Wt, vJ);
    /* J[0][0]=fx/z, J[0][2]=-fx tx/z^2, J[1][1]=fy/z, J[1][2]=-fy ty/z^2 (row-major) */
    const float vt0 = -fx * rz2 * vJ[2];
    const float vt1 = -fy * rz2 * vJ[5];
    const float vt2 = -fx * rz2 * vJ[0] + 2.f * fx * tx * rz3 * vJ[2] - fy * rz2 * vJ[4] + 2.f * fy * ty * rz3 * vJ[5];
    /* v_mean += W^T v_t */
    vm[0] += vt0 * W[0] + vt1 * W[3] + vt2 * W[6];
    vm[1] += vt0 * W[1] + vt1 * W[4] + vt2 * W[7];
    vm[2] += vt0 * W[2] + vt1 * W[5] + vt2 * W[8];
    v_mean3d[3 * i] = vm[0]; v_mean3d[3 * i + 1] = vm[1]; v_mean3d[3 * i + 2] = vm[2];

    /* scale_rot_to_cov3d_vjp */
    float vVs[9] = {vc3[0], 0.5f * vc3[1], 0.5f * vc3[2], 0.5f * vc3[1], vc3[3], 0.5f * vc3[4], 0.5f * vc3[2], 0.5f * vc3[4], vc3[5]};
    float R[9]; quat_to_rotmat(quats + 4 * i, R);
    const float sx = glob_scale * scales[3 * i], sy = glob_scale * scales[3 * i + 1], sz = glob_scale * scales[3 * i + 2];
    float Mm[9] = {R[0] * sx, R[1] * sy, R[2] * sz, R[3] * sx, R[4] * sy, R[5] * sz, R[6] * sx, R[7] * sy, R[8] * sz};
    float vM[9]; mm3(vVs, Mm, vM);
    for (int k = 0; k < 9; ++k) vM[k] = 2.f * vM[k];
    /* v_scale_j = <R[:,j], vM[:,j]> * glob_scale */
    v_scale[3 * i + 0] = (R[0] * vM[0] + R[3] * vM[3] + R[6] * vM[6]) * glob_scale;
    v_scale[3 * i + 1] = (R[1] * vM[1] + R[4] * vM[4] + R[7] * vM[7]) * glob_scale;
    v_scale[3 * i + 2] = (R[2] * vM[2] + R[5] * vM[5] + R[8] * vM[8]) * glob_scale;
    /* v_R = vM S */
    float vR[9] = {vM[0] * sx, vM[1] * sy, vM[2] * sz, vM[3] * sx, vM[4] * sy, vM[5] * sz, vM[6] * sx, vM[7] * sy, vM[8] * sz};
    /* quat_to_rotmat_vjp, in terms of row-major vR[r][c] (glm v_R[c][r]) */
    const float* q = quats + 4 * i;
    float qn = 1.f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float w = q[0] * qn, x = q[1] * qn, y = q[2] * qn, z = q[3] * qn;
#define VR(r, c) vR[(r) * 3 + (c)]
    v_quat[4 * i + 0] = 2.f * (x * (VR(2, 1) - VR(1, 2)) + y * (VR(0, 2) - VR(2, 0)) + z * (VR(1, 0) - VR(0, 1)));
    v_quat[4 * i + 1] = 2.f * (-2.f * x * (VR(1, 1) + VR(2, 2)) + y * (VR(1, 0) + VR(0, 1)) + z * (VR(2, 0) + VR(0, 2)) + w * (VR(2, 1) - VR(1, 2)));
    v_quat[4 * i + 2] = 2.f * (x * (VR(1, 0) + VR(0, 1)) - 2.f * y * (VR(0, 0) + VR(2, 2)) + z * (VR(2, 1) + VR(1, 2)) + w * (VR(0, 2) - VR(2, 0)));
    v_quat[4 * i + 3] = 2.f * (x * (VR(2, 0) + VR(0, 2)) + y * (VR(2, 1) + VR(1, 2)) - 2.f * z * (VR(0, 0) + VR(1, 1)) + w * (VR(1, 0) - VR(0, 1)));
#undef VR
  }
}

/* ---------- binning: cumsum -> isect ids -> stable sort -> tile bins ------------------------- */

/* returns number of intersections; cum_tiles_hit (int32) inclusive cumsum */
ORC_API int64_t orc_cumsum(int G, const int32_t* num_tiles_hit, int32_t* cum_tiles_hit) {
  int64_t acc = 0;
  for (int i = 0; i < G; ++i) { acc += num_tiles_hit[i]; cum_tiles_hit[i] = (int32_t)acc; }
  return acc;
}

ORC_API void orc_map_to_intersects(int G, const float* xys, const float* depths, const int32_t* radii,
                                   const int32_t* cum_tiles_hit, int img_h, int img_w, int block_width,
                                   int64_t* isect_ids, int32_t* gaussian_ids) {
  const int tbx = (img_w + block_width - 1) / block_width;
  const int tby = (img_h + block_width - 1) / block_width;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < G; ++i) {
    if (radii[i] <= 0) continue;
    int x0, y0, x1, y1;
    tile_bbox(xys[2 * i], xys[2 * i + 1], (float)radii[i], tbx, tby, block_width, &x0, &y0, &x1, &y1);
    int32_t cur = (i == 0) ? 0 : cum_tiles_hit[i - 1];
    int32_t dbits; memcpy(&dbits, depths + i, 4);
    const int64_t depth_id = (int64_t)dbits; /* sign-extending like (int64_t)*(int32_t*)&depth */
    for (int ty = y0; ty < y1; ++ty)
      for (int tx = x0; tx < x1; ++tx) {
        const int64_t tile_id = (int64_t)ty * tbx + tx;
        isect_ids[cur] = (tile_id << 32) | depth_id;
        gaussian_ids[cur] = i;
        ++cur;
      }
  }
}

typedef struct { int64_t key; int32_t val; } kv_t;
static int kv_cmp(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->val > y->val) - (x->val < y->val); /* stable: intersections are emitted in id order */
}

/* ascending int64 sort, ties keep emission order (== ascending gaussian id) */
ORC_API void orc_sort_intersects(int64_t n, const int64_t* isect_ids, const int32_t* gaussian_ids,
                                 int64_t* isect_sorted, int32_t* gids_sorted) {
  kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) { kv[i].key = isect_ids[i]; kv[i].val = gaussian_ids[i]; }
  qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
  for (int64_t i = 0; i < n; ++i) { isect_sorted[i] = kv[i].key; gids_sorted[i] = kv[i].val; }
  free(kv);
}

/* tile_bins [T,2] pre-zeroed by the caller (untouched tiles stay (0,0)) */
ORC_API void orc_tile_bin_edges(int64_t n, const int64_t* isect_sorted, int32_t* tile_bins) {
  for (int64_t i = 0; i < n; ++i) {
    const int32_t cur = (int32_t)(isect_sorted[i] >> 32);
    if (i == 0) tile_bins[2 * cur] = 0;
    if (i == n - 1) tile_bins[2 * cur + 1] = (int32_t)n;
    if (i == 0) continue;
    const int32_t prev = (int32_t)(isect_sorted[i - 1] >> 32);
    if (prev != cur) { tile_bins[2 * prev + 1] = (int32_t)i; tile_bins[2 * cur] = (int32_t)i; }
  }
}

/* ---------- blend forward --------------------------------------------------------------------
 * C colour channels (reference uses 3; the fused rgb+depth entry point of the product uses 4).
 * Per pixel, front to back over the tile's sorted list; semantics of rasterize_forward:
 *   sigma = .5(A dx^2 + C dy^2) + B dx dy ; alpha = min(.999, o exp(-sigma))
 *   skip if sigma<0 or alpha<1/255 ; next_T = T(1-alpha) ; stop (before adding) if next_T<=1e-4
 * final_idx = sorted-list index of the last Gaussian blended (0 if none). */
ORC_API void orc_rasterize_fwd(int img_h, int img_w, int block_width, int C, const int32_t* gids_sorted,
                               const int32_t* tile_bins, const float* xys, const float* conics,
                               const float* colors, const float* opacities, const float* background,
                               float* out_img, float* final_Ts, int32_t* final_idx) {
  const int tbx = (img_w + block_width - 1) / block_width;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < img_h; ++i) {
    for (int j = 0; j < img_w; ++j) {
      const int tile = (i / block_width) * tbx + (j / block_width);
      const int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
      const float px = (float)j + 0.5f, py = (float)i + 0.5f;
      float T = 1.f; int cur_idx = 0;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = lo; k < hi; ++k) {
        const int g = gids_sorted[k];
        const float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
        const float A = conics[3 * g], B = conics[3 * g + 1], Cc = conics[3 * g + 2];
        const float sigma = 0.5f * (A * dx * dx + Cc * dy * dy) + B * dx * dy;
        const float alpha = fminf(0.999f, opacities[g] * expf(-sigma));
        if (sigma < 0.f || alpha < 1.f / 255.f) continue;
        const float next_T = T * (1.f - alpha);
        if (next_T <= 1e-4f) break;
        const float vis = alpha * T;
        for (int c = 0; c < C; ++c) acc[c] = acc[c] + colors[(size_t)C * g + c] * vis;
        T = next_T; cur_idx = k;
      }
      const size_t pix = (size_t)i * img_w + j;
      final_Ts[pix] = T; final_idx[pix] = cur_idx;
      for (int c = 0; c < C; ++c) out_img[pix * C + c] = acc[c] + T * background[c];
    }
  }
}

/* ---------- blend backward -------------------------------------------------------------------
 * rasterize_backward_kernel of gsplat 0.1.x.  Contract constant kept from the published source:
 * the BACKWARD clamps alpha at 0.99 (forward: 0.999) — see ORC_BWD_ALPHA_CLAMP. */
#define ORC_BWD_ALPHA_CLAMP 0.99f
ORC_API void orc_rasterize_bwd(int G, int img_h, int img_w, int block_width, int C,
                               const int32_t* gids_sorted, const int32_t* tile_bins, const float* xys,
                               const float* conics, const float* colors, const float* opacities,
                               const float* background, const float* final_Ts, const int32_t* final_idx,
                               const float* v_output, const float* v_output_alpha, float* v_xy,
                               float* v_conic, float* v_colors, float* v_opacity) {
  /* per-Gaussian sums are accumulated in double (order-independent to ~1e-16) so the fp32 atomics
   * of the GPU path are compared against a clean value; per-pixel arithmetic stays fp32. */
  const int tbx = (img_w + block_width - 1) / block_width;
  const int S = C + 6; /* [colors C | conic 3 | xy 2 | opacity 1] */
  double* acc = (double*)calloc((size_t)(G > 0 ? G : 1) * S, sizeof(double));
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < img_h; ++i) {
    for (int j = 0; j < img_w; ++j) {
      const int tile = (i / block_width) * tbx + (j / block_width);
      const int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
      if (hi <= lo) continue;
      const size_t pix = (size_t)i * img_w + j;
      const float px = (float)j + 0.5f, py = (float)i + 0.5f;
      const float T_final = final_Ts[pix];
      float T = T_final;
      float buffer[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const int bin_final = final_idx[pix];
      const float* vo = v_output + pix * C;
      const float voa = v_output_alpha[pix];
      for (int k = imin(bin_final, hi - 1); k >= lo; --k) {
        const int g = gids_sorted[k];
        const float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
        const float A = conics[3 * g], B = conics[3 * g + 1], Cc = conics[3 * g + 2];
        const float sigma = 0.5f * (A * dx * dx + Cc * dy * dy) + B * dx * dy;
        const float vis = expf(-sigma);
        const float opac = opacities[g];
        const float alpha = fminf(ORC_BWD_ALPHA_CLAMP, opac * vis);
        if (sigma < 0.f || alpha < 1.f / 255.f) continue;
        const float ra = 1.f / (1.f - alpha);
        T *= ra;
        const float fac = alpha * T;
        float v_alpha = 0.f;
        double* a = acc + (size_t)g * S;
        for (int c = 0; c < C; ++c) {
          const float col = colors[(size_t)C * g + c];
          const float vc = fac * vo[c];
#pragma omp atomic
          a[c] += (double)vc;
          v_alpha += (col * T - buffer[c] * ra) * vo[c];
        }
        v_alpha += T_final * ra * voa;
        for (int c = 0; c < C; ++c) v_alpha += -T_final * ra * background[c] * vo[c];
        for (int c = 0; c < C; ++c) buffer[c] += colors[(size_t)C * g + c] * fac;
        const float v_sigma = -opac * vis * v_alpha;
        const float g0 = 0.5f * v_sigma * dx * dx, g1 = v_sigma * dx * dy, g2 = 0.5f * v_sigma * dy * dy;
        const float gx = v_sigma * (A * dx + B * dy), gy = v_sigma * (B * dx + Cc * dy);
        const float go = vis * v_alpha;
#pragma omp atomic
        a[C + 0] += (double)g0;
#pragma omp atomic
        a[C + 1] += (double)g1;
#pragma omp atomic
        a[C + 2] += (double)g2;
#pragma omp atomic
        a[C + 3] += (double)gx;
#pragma omp atomic
        a[C + 4] += (double)gy;
#pragma omp atomic
        a[C + 5] += (double)go;
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int g = 0; g < G; ++g) {
    const double* a = acc + (size_t)g * S;
    for (int c = 0; c < C; ++c) v_colors[(size_t)C * g + c] = (float)a[c];
    v_conic[3 * g] = (float)a[C]; v_conic[3 * g + 1] = (float)a[C + 1]; v_conic[3 * g + 2] = (float)a[C + 2];
    v_xy[2 * g] = (float)a[C + 3]; v_xy[2 * g + 1] = (float)a[C + 4];
    v_opacity[g] = (float)a[C + 5];
  }
  free(acc);
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
