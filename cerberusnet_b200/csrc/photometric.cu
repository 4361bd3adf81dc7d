// photometric.cu -- the photometric term of unFlowLoss, fused: flow-warp + L1 + SSIM(3x3) + mean, forward and
// backward, sm_100a.  (the second-largest consumer of flow_warp in a training step.)
//
// Replaces, per scale and direction (reference paths relative to the reference checkout):
//   flow_warp(im_src, flow)                               nnet_training/loss_functions/UnFlowLoss.py:83-94, 279-283
//   unFlowLoss.loss_photometric with the occlusion mask   UnFlowLoss.py:225-241 (the mask is all ones: the reference
//                                                         disabled its occlusion estimate, :285-297)
//   SSIM (ReflectionPad2d(1) + five AvgPool2d(3,1))       nnet_training/loss_functions/loss_functions.py:47-78
//   autograd of all of the above with respect to the flow
// i.e. ~30 PyTorch kernels, the warped image, five pooled maps and their gradients per call become two launches each
// way and nothing but the scalar loss / the flow gradient is written.
//
//   rec   = flow_warp(im_src, flow)
//   loss  = mean_{n,c,y,x} ( l1_w * |im_orig - rec| + ssim_w * clamp((1 - ssim_n / ssim_d) / 2, 0, 1) )
//   x = rec, y = im_orig, both reflection-padded by 1;  mu = avgpool3(.), sigma_x = avgpool3(x^2) - mu_x^2, ...
//   ssim_n = (2 mu_x mu_y + C1)(2 sigma_xy + C2),  ssim_d = (mu_x^2 + mu_y^2 + C1)(sigma_x + sigma_y + C2)
// Images are inputs of the loss (no gradient); the backward returns d loss / d flow only.
#include "costvolume_common.cuh"
#include "costvolume_launch.h"

namespace cerb {

constexpr int PT = 16;                 // tile edge (pixels)
constexpr float kC1 = 0.01f * 0.01f;   // loss_functions.py:60-61
constexpr float kC2 = 0.03f * 0.03f;

__device__ __forceinline__ int reflect1(int i, int n) {   // ReflectionPad2d(1): -1 -> 1, n -> n-2
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct WarpTap {
  int off[4];
  float w[4];
  float dwx[2], dwy[2];   // weights of the x / y finite differences (backward)
  bool in_x, in_y, bx1, by1;
};

__device__ __forceinline__ WarpTap warp_tap(const float* __restrict__ flow_n, long long plane, int y, int x, int H, int W,
                                            const AxisConst& ax, const AxisConst& ay, int mode) {
  WarpTap t;
  const float* fp = flow_n + (long long)y * W + x;
  const float sx = sample_pos(x, __ldg(fp), ax, mode, t.in_x);
  const float sy = sample_pos(y, __ldg(fp + plane), ay, mode, t.in_y);
  const Taps tp = make_taps(sx, sy, H, W, W);
#pragma unroll
  for (int k = 0; k < 4; ++k) { t.off[k] = tp.off[k]; t.w[k] = tp.w[k]; }
  const float fx = floorf(sx), fy = floorf(sy);
  t.dwy[0] = fy + 1.f - sy; t.dwy[1] = sy - fy;   // wy1, wy0
  t.dwx[0] = fx + 1.f - sx; t.dwx[1] = sx - fx;   // wx1, wx0
  t.bx1 = (int)fx + 1 < W; t.by1 = (int)fy + 1 < H;
  return t;
}

__device__ __forceinline__ float warp_value(const float* __restrict__ img_c, const WarpTap& t) {
  Taps tp;
#pragma unroll
  for (int k = 0; k < 4; ++k) { tp.off[k] = t.off[k]; tp.w[k] = t.w[k]; }
  return blend(__ldg(img_c + t.off[0]), __ldg(img_c + t.off[1]), __ldg(img_c + t.off[2]), __ldg(img_c + t.off[3]), tp);
}

// SSIM statistics of the 3x3 window centred at (wy, wx) of two smem tiles with row pitch `pitch`
struct SsimStat { float mu_x, mu_y, sxx, syy, sxy; };
__device__ __forceinline__ SsimStat ssim_stat(const float* xs, const float* ys, int pitch, int wy, int wx) {
  float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const float a = xs[(wy + dy) * pitch + wx + dx], b = ys[(wy + dy) * pitch + wx + dx];
      sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
    }
  SsimStat s;
  const float inv9 = 1.f / 9.f;   // avg_pool2d: sum / 9
  s.mu_x = sx * inv9; s.mu_y = sy * inv9; s.sxx = sxx * inv9; s.syy = syy * inv9; s.sxy = sxy * inv9;
  return s;
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(PT * PT) photometric_fwd_kernel(const float* __restrict__ orig, const float* __restrict__ src,
                                                                  const float* __restrict__ flow, double* __restrict__ partial,
                                                                  int B, int C, int H, int W, float l1_w, float ssim_w, int mode) {
  constexpr int HP = PT + 2;   // tile + reflection halo
  __shared__ float xs[HP * HP], ys[HP * HP];
  __shared__ double red[PT * PT / 32];
  const int tid = threadIdx.x, n = blockIdx.z;
  const int ty0 = blockIdx.y * PT, tx0 = blockIdx.x * PT;
  const long long plane = (long long)H * W;
  const AxisConst ax = make_axis(W), ay = make_axis(H);
  const float* flow_n = flow + (long long)n * 2 * plane;
  // sampling data of the (up to two) halo positions this thread stages, shared by all channels
  WarpTap tp[2];
  int hpos[2], opos[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + j * PT * PT;
    hpos[j] = i < HP * HP ? i : -1;
    const int hy = i / HP, hx = i - hy * HP;
    const int y = reflect1(min(ty0 + hy - 1, H), H), x = reflect1(min(tx0 + hx - 1, W), W);   // past the image: any valid pixel
    opos[j] = y * W + x;
    tp[j] = warp_tap(flow_n, plane, y, x, H, W, ax, ay, mode);
  }
  const int py = tid / PT, px = tid - py * PT;
  const bool live = ty0 + py < H && tx0 + px < W;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* oc = orig + ((long long)n * C + c) * plane;
    const float* sc = src + ((long long)n * C + c) * plane;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (hpos[j] >= 0) { xs[hpos[j]] = warp_value(sc, tp[j]); ys[hpos[j]] = __ldg(oc + opos[j]); }
    __syncthreads();
    if (live) {
      const SsimStat s = ssim_stat(xs, ys, HP, py + 1, px + 1);
      const float sig_x = s.sxx - s.mu_x * s.mu_x, sig_y = s.syy - s.mu_y * s.mu_y, sig_xy = s.sxy - s.mu_x * s.mu_y;
      const float nn = (2.f * s.mu_x * s.mu_y + kC1) * (2.f * sig_xy + kC2);
      const float dd = (s.mu_x * s.mu_x + s.mu_y * s.mu_y + kC1) * (sig_x + sig_y + kC2);
      const float v = fminf(fmaxf((1.f - nn / dd) * 0.5f, 0.f), 1.f);
      const float x0 = xs[(py + 1) * HP + px + 1], y0 = ys[(py + 1) * HP + px + 1];
      acc += l1_w * fabsf(y0 - x0) + ssim_w * v;
    }
  }
  double d = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((tid & 31) == 0) red[tid >> 5] = d;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < PT * PT / 32; ++w) s += red[w];
    partial[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
  }
}

// fixed-order sum of the per-tile partials (bit-reproducible run to run), scaled to the mean
__global__ void __launch_bounds__(256) photometric_finalize_kernel(const double* __restrict__ partial, int n, double inv_count,
                                                                   float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] * inv_count);
}

// ---------------------------------------------------------------- backward
// d ssim_q / d x_i = alpha_q + beta_q x_i + gamma_q y_i for every sample i of window q (0 outside the clamp range):
//   beta  =  n (mu_x^2 + mu_y^2 + C1) / (9 d^2)
//   gamma = -(2 mu_x mu_y + C1) / (9 d)
//   alpha = -1/2 [ dn_const / d - n dd_const / d^2 ],   dn_const = (2 mu_y / 9) ((2 sig_xy + C2) - (2 mu_x mu_y + C1)),
//                                                       dd_const = (2 mu_x / 9) ((sig_x + sig_y + C2) - (mu_x^2 + mu_y^2 + C1))
__global__ void __launch_bounds__(PT * PT) photometric_bwd_kernel(const float* __restrict__ orig, const float* __restrict__ src,
                                                                  const float* __restrict__ flow, const float* __restrict__ gloss,
                                                                  float* __restrict__ gflow, int B, int C, int H, int W,
                                                                  float l1_w, float ssim_w, int mode, float inv_count) {
  constexpr int HP = PT + 4, CP = PT + 2;   // samples: tile + 2; windows: tile + 1
  __shared__ float xs[HP * HP], ys[HP * HP];
  __shared__ float ca[CP * CP], cb[CP * CP], cg[CP * CP];
  const int tid = threadIdx.x, n = blockIdx.z;
  const int ty0 = blockIdx.y * PT, tx0 = blockIdx.x * PT;
  const long long plane = (long long)H * W;
  const AxisConst ax = make_axis(W), ay = make_axis(H);
  const float* flow_n = flow + (long long)n * 2 * plane;
  WarpTap tp[2];
  int hpos[2], opos[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = tid + j * PT * PT;
    hpos[j] = i < HP * HP ? i : -1;
    const int hy = i / HP, hx = i - hy * HP;
    const int y = reflect1(min(max(ty0 + hy - 2, -1), H), H), x = reflect1(min(max(tx0 + hx - 2, -1), W), W);
    opos[j] = y * W + x;
    tp[j] = warp_tap(flow_n, plane, y, x, H, W, ax, ay, mode);
  }
  const int py = tid / PT, px = tid - py * PT;
  const int gy = ty0 + py, gx = tx0 + px;
  const bool live = gy < H && gx < W;
  WarpTap me;
  if (live) me = warp_tap(flow_n, plane, gy, gx, H, W, ax, ay, mode);
  float gu = 0.f, gv = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* oc = orig + ((long long)n * C + c) * plane;
    const float* sc = src + ((long long)n * C + c) * plane;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (hpos[j] >= 0) { xs[hpos[j]] = warp_value(sc, tp[j]); ys[hpos[j]] = __ldg(oc + opos[j]); }
    __syncthreads();
    // coefficient maps of the windows centred on tile + 1 (only windows whose centre is a pixel of the image exist)
    for (int i = tid; i < CP * CP; i += PT * PT) {
      const int wy = i / CP, wx = i - wy * CP;
      const int qy = ty0 + wy - 1, qx = tx0 + wx - 1;
      float a = 0.f, b = 0.f, g = 0.f;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const SsimStat s = ssim_stat(xs, ys, HP, wy + 1, wx + 1);
        const float sig_x = s.sxx - s.mu_x * s.mu_x, sig_y = s.syy - s.mu_y * s.mu_y, sig_xy = s.sxy - s.mu_x * s.mu_y;
        const float n1 = 2.f * s.mu_x * s.mu_y + kC1, n2 = 2.f * sig_xy + kC2;
        const float d1 = s.mu_x * s.mu_x + s.mu_y * s.mu_y + kC1, d2 = sig_x + sig_y + kC2;
        const float nn = n1 * n2, dd = d1 * d2;
        const float v = (1.f - nn / dd) * 0.5f;
        if (v > 0.f && v < 1.f) {   // gradient of clamp
          const float inv9 = 1.f / 9.f, idd = 1.f / dd;
          const float dn_const = 2.f * s.mu_y * inv9 * (n2 - n1);
          const float dd_const = 2.f * s.mu_x * inv9 * (d2 - d1);
          a = -0.5f * (dn_const * idd - nn * dd_const * idd * idd);
          b = nn * d1 * 2.f * inv9 * idd * idd * 0.5f;
          g = -n1 * 2.f * inv9 * idd * 0.5f;
        }
      }
      ca[i] = a; cb[i] = b; cg[i] = g;
    }
    __syncthreads();
    if (live) {
      const float x0 = xs[(py + 2) * HP + px + 2], y0 = ys[(py + 2) * HP + px + 2];
      // every padded sample position that reflects onto this pixel: itself, plus the mirror images across a border
      float A = 0.f, Bc = 0.f, G = 0.f;
      int iy[2] = {gy, gy}, ix[2] = {gx, gx}, ny = 1, nx = 1;   // H, W >= 4: a coordinate mirrors across at most one border
      if (gy == 1) { iy[1] = -1; ny = 2; } else if (gy == H - 2) { iy[1] = H; ny = 2; }
      if (gx == 1) { ix[1] = -1; nx = 2; } else if (gx == W - 2) { ix[1] = W; nx = 2; }
      for (int a_ = 0; a_ < ny; ++a_)
        for (int b_ = 0; b_ < nx; ++b_)
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int qy = iy[a_] + dy, qx = ix[b_] + dx;
              if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
                const int ci = (qy - ty0 + 1) * CP + (qx - tx0 + 1);   // inside tile + 1 by construction
                A += ca[ci]; Bc += cb[ci]; G += cg[ci];
              }
            }
      const float d_rec = l1_w * (x0 > y0 ? 1.f : (x0 < y0 ? -1.f : 0.f)) + ssim_w * (A + Bc * x0 + G * y0);
      const float v_nw = __ldg(sc + me.off[0]);
      const float v_ne = me.bx1 ? __ldg(sc + me.off[1]) : 0.f;
      const float v_sw = me.by1 ? __ldg(sc + me.off[2]) : 0.f;
      const float v_se = (me.bx1 && me.by1) ? __ldg(sc + me.off[3]) : 0.f;
      gu += d_rec * ((v_ne - v_nw) * me.dwy[0] + (v_se - v_sw) * me.dwy[1]);
      gv += d_rec * ((v_sw - v_nw) * me.dwx[0] + (v_se - v_ne) * me.dwx[1]);
    }
  }
  if (live) {
    const float s = __ldg(gloss) * inv_count;
    float* gf = gflow + (long long)n * 2 * plane + (long long)gy * W + gx;
    gf[0] = me.in_x ? gu * pos_scale(W, mode) * s : 0.f;
    gf[plane] = me.in_y ? gv * pos_scale(H, mode) * s : 0.f;
  }
}

size_t photometric_workspace_bytes(int B, int H, int W) {
  return sizeof(double) * (size_t)B * ((H + PT - 1) / PT) * ((W + PT - 1) / PT);
}

cudaError_t launch_photometric_forward(const float* orig, const float* src, const float* flow, float* loss, void* workspace, int B,
                                       int C, int H, int W, float l1_w, float ssim_w, int mode, cudaStream_t stream) {
  dim3 grid((W + PT - 1) / PT, (H + PT - 1) / PT, B);
  const int nblk = (int)(grid.x * grid.y * grid.z);
  photometric_fwd_kernel<<<grid, PT * PT, 0, stream>>>(orig, src, flow, (double*)workspace, B, C, H, W, l1_w, ssim_w, mode);
  photometric_finalize_kernel<<<1, 256, 0, stream>>>((const double*)workspace, nblk, 1.0 / ((double)B * C * H * W), loss);
  count_launches(2);
  return cudaGetLastError();
}

cudaError_t launch_photometric_backward(const float* orig, const float* src, const float* flow, const float* gloss, float* gflow,
                                        int B, int C, int H, int W, float l1_w, float ssim_w, int mode, cudaStream_t stream) {
  dim3 grid((W + PT - 1) / PT, (H + PT - 1) / PT, B);
  photometric_bwd_kernel<<<grid, PT * PT, 0, stream>>>(orig, src, flow, gloss, gflow, B, C, H, W, l1_w, ssim_w, mode,
                                                       (float)(1.0 / ((double)B * C * H * W)));
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace cerb
