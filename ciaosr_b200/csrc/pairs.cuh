// Per-(query, neighbour) geometry: the local-ensemble coordinate encoding of
// ciaosr_net.py:148-193, shared by both engines.  All fp32 steps are explicitly
// rounded in the reference's order (see oracle/ciaosr_oracle.py::query_rgb).
#pragma once
#include "common.cuh"

namespace ciaosr {

struct PairInfo {
  int pix;        // b*H*W + iy*W + ix of the neighbour latent code, -1 if outside (zero padding)
  int gidx;       // (b*H*W + query pixel)*9 + (dy+1)*3 + (dx+1), -1 if the query pixel is outside
  float rel_y, rel_x, sc_y, sc_x;
};

struct PairConsts {
  int H, W, Q, eval_bsize, local_size;
  float cy0, cy1, cx0, cx1;
};

// neighbour n of local_size -> (vx, vy) in {-1,0,1}; order of ciaosr_net.py:152-155
__device__ __forceinline__ void neighbour_offset(int local_size, int n, int& vx, int& vy) {
  if (local_size == 1) { vx = 0; vy = 0; }
  else if (local_size == 2) { vx = (n >> 1) * 2 - 1; vy = (n & 1) * 2 - 1; }
  else { vx = n / 3 - 1; vy = n % 3 - 1; }
}

__device__ __forceinline__ PairInfo compute_pair(const PairConsts& k, const float* __restrict__ coord,
                                                 const float* __restrict__ cell, long long g, int n) {
  const int b = (int)(g / k.Q), q = (int)(g % k.Q);
  const float cy = coord[g * 2 + 0], cx = coord[g * 2 + 1];
  const float ly = cell[g * 2 + 0], lx = cell[g * 2 + 1];
  // tx, ty come from the first cell of the reference's eval_bsize chunk (ciaosr_net.py:162-163, 243)
  const int q0 = k.eval_bsize > 0 ? (q / k.eval_bsize) * k.eval_bsize : 0;
  const long long g0 = (long long)b * k.Q + q0;
  const float c0y = cell[g0 * 2 + 0], c0x = cell[g0 * 2 + 1];
  int vx, vy;
  neighbour_offset(k.local_size, n, vx, vy);
  float sy = cy, sx = cx;
  const float lo = (float)(-1.0 + 1e-6), hi = (float)(1.0 - 1e-6), eps = (float)1e-6;
  if (vx != 0) {
    // (H-1) / (1 - cell)  is  reciprocal(1 - cell) * (H-1)  in torch (int.__truediv__(Tensor))
    const float tx = __fmul_rn(__fdiv_rn(1.0f, __fsub_rn(1.0f, c0y)), (float)(k.H - 1));
    const float rx = __fdiv_rn(1.0f, tx);
    sy = __fadd_rn(sy, __fadd_rn(vx > 0 ? rx : -rx, eps));
  }
  if (vy != 0) {
    const float ty = __fmul_rn(__fdiv_rn(1.0f, __fsub_rn(1.0f, c0x)), (float)(k.W - 1));
    const float ry = __fdiv_rn(1.0f, ty);
    sx = __fadd_rn(sx, __fadd_rn(vy > 0 ? ry : -ry, eps));
  }
  sy = fminf(fmaxf(sy, lo), hi);
  sx = fminf(fmaxf(sx, lo), hi);
  const int iy = nearest_index(sy, k.H), ix = nearest_index(sx, k.W);
  const int qy = nearest_index(cy, k.H), qx = nearest_index(cx, k.W);
  PairInfo p;
  const bool nb_ok = iy >= 0 && iy < k.H && ix >= 0 && ix < k.W;
  const bool q_ok = qy >= 0 && qy < k.H && qx >= 0 && qx < k.W;
  p.pix = nb_ok ? (b * k.H + iy) * k.W + ix : -1;
  int dy = iy - qy, dx = ix - qx;
  dy = max(-1, min(1, dy));   // |d| <= 1 always holds for an in-range query (shift < 1 LR pixel)
  dx = max(-1, min(1, dx));
  p.gidx = (q_ok && nb_ok) ? ((b * k.H + qy) * k.W + qx) * 9 + (dy + 1) * 3 + (dx + 1) : -1;
  // centre of the neighbour latent code = make_coord((H,W))[iy,ix]; zero when out of range
  const float ky = nb_ok ? __fadd_rn(k.cy0, __fmul_rn(k.cy1, (float)iy)) : 0.0f;
  const float kx = nb_ok ? __fadd_rn(k.cx0, __fmul_rn(k.cx1, (float)ix)) : 0.0f;
  p.rel_y = __fmul_rn(__fsub_rn(cy, ky), (float)k.H);
  p.rel_x = __fmul_rn(__fsub_rn(cx, kx), (float)k.W);
  p.sc_y = __fmul_rn(ly, (float)k.H);
  p.sc_x = __fmul_rn(lx, (float)k.W);
  return p;
}

// value / key vector element in tap-major order: cp < 9C -> feat[pix + tap][ch], else non-local
__device__ __forceinline__ float value_at(const float* __restrict__ featT,
                                          const float* __restrict__ nlT, int pix, int cp, int H,
                                          int W, int C, int Cn) {
  if (pix < 0) return 0.0f;
  if (cp >= 9 * C) return nlT[(long long)pix * Cn + (cp - 9 * C)];
  const int t = cp / C, ch = cp % C;
  const int hw = pix % (H * W), b = pix / (H * W);
  const int y = hw / W + t / 3 - 1, x = hw % W + t % 3 - 1;
  if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
  return featT[(((long long)b * H + y) * W + x) * C + ch];
}

// bilinear, border padding, align_corners=False (ciaosr_net.py:107-108)
__device__ __forceinline__ float bilinear_border(const float* __restrict__ img, int H, int W,
                                                 float cy, float cx) {
  float uy = fminf(fmaxf(unnormalize(cy, H), 0.0f), (float)(H - 1));
  float ux = fminf(fmaxf(unnormalize(cx, W), 0.0f), (float)(W - 1));
  const float y0f = floorf(uy), x0f = floorf(ux);
  const int y0 = (int)y0f, x0 = (int)x0f;
  const float fy = uy - y0f, fx = ux - x0f, gy = (y0f + 1.0f) - uy, gx = (x0f + 1.0f) - ux;
  float acc = img[y0 * W + x0] * (gx * gy);
  if (x0 + 1 < W) acc += img[y0 * W + x0 + 1] * (fx * gy);
  if (y0 + 1 < H) acc += img[(y0 + 1) * W + x0] * (gx * fy);
  if (y0 + 1 < H && x0 + 1 < W) acc += img[(y0 + 1) * W + x0 + 1] * (fx * fy);
  return acc;
}

}  // namespace ciaosr
