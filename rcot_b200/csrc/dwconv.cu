// dwconv.cu -- lean depthwise 3x3 kernels for the aligned case (W % 4 == 0, H % ROWS == 0), which is
// every feature map of a training patch.  Each thread owns a 4-wide x ROWS-high output patch of one
// (image, channel) plane: ROWS+2 input rows of 6 values are fetched once (one aligned float4 + two edge
// scalars per row) and reused by all taps; patches that do not touch the plane border take a path without
// any predicate.  Thread/plane mapping costs one integer division per thread: large planes put the plane
// in the grid (y = channel, z = image); small planes pack several planes into one 256-thread CTA.
//   plain / transposed (+ row sums of squares for MDTA's q,k norms)   Net_Restormer.py:26
//   GELU gate and its backward                                          Net_Restormer.py:75,81-83
//   fused backward: din = dw^T(dout), dW += corr(in, dout)
// The generic kernels in elem.cu remain the fallback for odd sizes (whole-image inference).
#include "../../include/rcot_b200.h"
#include "common.cuh"

namespace rcot {

__device__ __forceinline__ float gelu_erf_d(float a) { return 0.5f * a * (1.f + erff(a * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad_d(float a) {
  return 0.5f * (1.f + erff(a * 0.70710678118654752f)) + a * 0.39894228040143268f * __expf(-0.5f * a * a);
}

template <int ROWS>
struct Patch {
  float v[ROWS + 2][6];  // rows y-1 .. y+ROWS, columns x0-1 .. x0+4
};

template <int ROWS>
__device__ __forceinline__ void load_patch(Patch<ROWS>& P, const float* __restrict__ plane, int y, int x0, int H, int W) {
  const float* p = plane + (y - 1) * W + x0;
  if (y > 0 && y + ROWS < H && x0 > 0 && x0 + 4 < W) {   // interior: no predicates
#pragma unroll
    for (int r = 0; r < ROWS + 2; ++r) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(p + r * W));
      P.v[r][0] = __ldg(p + r * W - 1);
      P.v[r][1] = m.x;
      P.v[r][2] = m.y;
      P.v[r][3] = m.z;
      P.v[r][4] = m.w;
      P.v[r][5] = __ldg(p + r * W + 4);
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < ROWS + 2; ++r) {
    const int yy = y - 1 + r;
    if ((unsigned)yy < (unsigned)H) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(p + r * W));
      P.v[r][0] = x0 > 0 ? __ldg(p + r * W - 1) : 0.f;
      P.v[r][1] = m.x;
      P.v[r][2] = m.y;
      P.v[r][3] = m.z;
      P.v[r][4] = m.w;
      P.v[r][5] = x0 + 4 < W ? __ldg(p + r * W + 4) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) P.v[r][i] = 0.f;
    }
  }
}

template <int ROWS>
__device__ __forceinline__ void conv_patch(const Patch<ROWS>& P, const float (&w)[9], float (&o)[ROWS][4]) {
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc = fmaf(P.v[r + ky][j + kx], w[ky * 3 + kx], acc);
      o[r][j] = acc;
    }
}

// Which (image, plane, patch) a thread owns.  pshift < 0: large planes, the plane is (blockIdx.y, blockIdx.z);
// otherwise 2^pshift (>= patches per plane) threads per plane and 256 >> pshift planes per CTA.
struct DwGeom {
  int H, W, pw, ppp, planes, pshift, ROWSv;
};
struct DwThread {
  int b, ch, y, x0;
  bool active;
};
__device__ __forceinline__ DwThread dw_map(const DwGeom& g, int rows) {
  DwThread t;
  int patch;
  t.b = blockIdx.z;
  if (g.pshift < 0) {
    t.ch = blockIdx.y;
    patch = blockIdx.x * blockDim.x + threadIdx.x;
    t.active = patch < g.ppp;
  } else {
    t.ch = blockIdx.y * (256 >> g.pshift) + (threadIdx.x >> g.pshift);
    patch = threadIdx.x & ((1 << g.pshift) - 1);
    t.active = patch < g.ppp && t.ch < g.planes;
  }
  if (!t.active) {
    patch = 0;
    t.ch = 0;
  }
  const int py = patch / g.pw;
  t.y = py * rows;
  t.x0 = (patch - py * g.pw) * 4;
  return t;
}
__device__ __forceinline__ float warp_sum_dw(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ plain / transposed
template <int ROWS>
__global__ void __launch_bounds__(256)
    dw_plain_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ w, float* __restrict__ out,
                    int64_t out_bs, int flip, float* __restrict__ sumsq, int nsq, const DwGeom g) {
  const DwThread t = dw_map(g, ROWS);
  const int HW = g.H * g.W;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + t.ch * 9 + (flip ? 8 - i : i));
  float o[ROWS][4];
  float sq = 0.f;
  if (t.active) {
    Patch<ROWS> P;
    load_patch<ROWS>(P, in + (size_t)t.b * in_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
    conv_patch<ROWS>(P, wk, o);
    float* op = out + (size_t)t.b * out_bs + (size_t)t.ch * HW + t.y * g.W + t.x0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      *reinterpret_cast<float4*>(op + r * g.W) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) sq = fmaf(o[r][j], o[r][j], sq);
    }
  }
  if (sumsq) {   // uniform
    const bool mine = t.active && t.ch < nsq;
    if (g.pshift < 0 || g.pshift >= 5) {   // every warp lies inside one plane
      sq = warp_sum_dw(mine ? sq : 0.f);
      if ((threadIdx.x & 31) == 0 && mine) atomicAdd(sumsq + (size_t)t.b * nsq + t.ch, sq);
    } else if (mine) {
      atomicAdd(sumsq + (size_t)t.b * nsq + t.ch, sq);
    }
  }
}

// ------------------------------------------------------------------ GELU gate: forward / backward
// mode 1: out[j] = gelu(dw(in[j])) * dw(in[j+hid])
// mode 2: a = dw(in[j]), b = dw(in[j+hid]); out[j] = dg*b*gelu'(a); out[j+hid] = dg*gelu(a); g_out[j] = gelu(a)*b
template <int MODE>
__global__ void __launch_bounds__(256)
    dw_gate_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ w, float* __restrict__ out,
                   int64_t out_bs, int hid, const float* __restrict__ dg, int64_t dg_bs, float* __restrict__ g_out,
                   int64_t g_bs, const DwGeom g) {
  constexpr int ROWS = 2;
  const DwThread t = dw_map(g, ROWS);
  if (!t.active) return;
  const int HW = g.H * g.W;
  float w0[9], w1[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    w0[i] = __ldg(w + t.ch * 9 + i);
    w1[i] = __ldg(w + (t.ch + hid) * 9 + i);
  }
  const float* inb = in + (size_t)t.b * in_bs;
  float a[ROWS][4], gt[ROWS][4];
  {
    Patch<ROWS> P;
    load_patch<ROWS>(P, inb + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
    conv_patch<ROWS>(P, w0, a);
  }
  {
    Patch<ROWS> P;
    load_patch<ROWS>(P, inb + (size_t)(t.ch + hid) * HW, t.y, t.x0, g.H, g.W);
    conv_patch<ROWS>(P, w1, gt);
  }
  const int pix = t.y * g.W + t.x0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int pr = pix + r * g.W;
    if (MODE == 1) {
      *reinterpret_cast<float4*>(out + (size_t)t.b * out_bs + (size_t)t.ch * HW + pr) =
          make_float4(gelu_erf_d(a[r][0]) * gt[r][0], gelu_erf_d(a[r][1]) * gt[r][1], gelu_erf_d(a[r][2]) * gt[r][2],
                      gelu_erf_d(a[r][3]) * gt[r][3]);
    } else {
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(dg + (size_t)t.b * dg_bs + (size_t)t.ch * HW + pr));
      const float d[4] = {d4.x, d4.y, d4.z, d4.w};
      float da[4], db[4], gg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ga = gelu_erf_d(a[r][j]);
        da[j] = d[j] * gt[r][j] * gelu_erf_grad_d(a[r][j]);
        db[j] = d[j] * ga;
        gg[j] = ga * gt[r][j];
      }
      float* ob = out + (size_t)t.b * out_bs;
      *reinterpret_cast<float4*>(ob + (size_t)t.ch * HW + pr) = make_float4(da[0], da[1], da[2], da[3]);
      *reinterpret_cast<float4*>(ob + (size_t)(t.ch + hid) * HW + pr) = make_float4(db[0], db[1], db[2], db[3]);
      if (g_out)
        *reinterpret_cast<float4*>(g_out + (size_t)t.b * g_bs + (size_t)t.ch * HW + pr) =
            make_float4(gg[0], gg[1], gg[2], gg[3]);
    }
  }
}

// ------------------------------------------------------------------ fused backward: din and dW
__global__ void __launch_bounds__(256)
    dw_bwd2_kernel(const float* __restrict__ in, int64_t in_bs, const float* __restrict__ dout, int64_t dout_bs,
                   const float* __restrict__ w, float* __restrict__ din, int64_t din_bs, float* __restrict__ dw,
                   const DwGeom g, const int ppt) {
  constexpr int ROWS = 2;
  DwThread t = dw_map(g, ROWS);
  const int HW = g.H * g.W;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  float wf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wf[i] = __ldg(w + t.ch * 9 + 8 - i);   // flipped taps
  // large planes: a thread walks `ppt` patches (stride = patches covered by the grid) before the tap sums
  // are reduced, so the 9 warp reductions are amortised
  for (int it = 0; it < ppt; ++it) {
    if (it > 0) {
      const int patch = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
      t.active = patch < g.ppp;
      const int py = (t.active ? patch : 0) / g.pw;
      t.y = py * ROWS;
      t.x0 = ((t.active ? patch : 0) - py * g.pw) * 4;
    }
    if (!t.active) continue;
    float d[ROWS][4];
    {
      Patch<ROWS> P;
      load_patch<ROWS>(P, dout + (size_t)t.b * dout_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
      float o[ROWS][4];
      conv_patch<ROWS>(P, wf, o);
      float* dp = din + (size_t)t.b * din_bs + (size_t)t.ch * HW + t.y * g.W + t.x0;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        *reinterpret_cast<float4*>(dp + r * g.W) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[r][j] = P.v[r + 1][j + 1];   // centre rows of dout
      }
    }
    Patch<ROWS> Q;
    load_patch<ROWS>(Q, in + (size_t)t.b * in_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[ky * 3 + kx] = fmaf(d[r][j], Q.v[r + ky][j + kx], acc[ky * 3 + kx]);
  }
  // reduce the 9 tap sums over the threads that share a channel, one atomicAdd per group
  if (g.pshift < 0) {   // whole CTA = one channel
    __shared__ float red[9][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float s = warp_sum_dw(acc[i]);
      if (lane == 0) red[i][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
      float s = 0.f;
      for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
      atomicAdd(dw + blockIdx.y * 9 + threadIdx.x, s);
    }
  } else {
    const int width = g.pshift >= 5 ? 32 : (1 << g.pshift);   // lanes per channel inside a warp
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float s = acc[i];
      for (int o = width >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & (width - 1)) == 0 && t.ch < g.planes && (threadIdx.x >> g.pshift) < (256 >> g.pshift))
        atomicAdd(dw + t.ch * 9 + i, s);
    }
  }
}

// ------------------------------------------------------------------ launch geometry
static bool dw_geom(DwGeom& g, dim3& grid, int B, int planes, int H, int W, int rows) {
  if (W % 4 != 0 || H % rows != 0 || B > 65535) return false;
  g.H = H;
  g.W = W;
  g.pw = W / 4;
  g.ppp = g.pw * (H / rows);
  g.planes = planes;
  g.ROWSv = rows;
  if (g.ppp >= 256) {
    if (planes > 65535) return false;
    g.pshift = -1;
    grid = dim3(cdiv(g.ppp, 256), planes, B);
  } else {
    int sh = 0;
    while ((1 << sh) < g.ppp) ++sh;
    g.pshift = sh;
    const int ppb = 256 >> sh;
    grid = dim3(1, cdiv(planes, ppb), B);
    if (grid.y > 65535) return false;
  }
  return true;
}

// Returns 1 if the aligned fast path handled the call, 0 if the caller must use the generic kernels.
int dwconv_fast(const rcot_dw_params& p, int planes, cudaStream_t st) {
  const bool al = p.in_bs % 4 == 0 && p.out_bs % 4 == 0 && ((uintptr_t)p.in % 16 == 0) && ((uintptr_t)p.out % 16 == 0);
  if (!al) return 0;
  DwGeom g;
  dim3 grid;
  if (p.mode == 0) {
    if (dw_geom(g, grid, p.B, planes, p.H, p.W, 4)) {
      dw_plain_kernel<4><<<grid, 256, 0, st>>>(p.in, p.in_bs, p.w, p.out, p.out_bs, p.flip, p.sumsq, p.nsq, g);
      return 1;
    }
    if (dw_geom(g, grid, p.B, planes, p.H, p.W, 2)) {
      dw_plain_kernel<2><<<grid, 256, 0, st>>>(p.in, p.in_bs, p.w, p.out, p.out_bs, p.flip, p.sumsq, p.nsq, g);
      return 1;
    }
    return 0;
  }
  if (!dw_geom(g, grid, p.B, planes, p.H, p.W, 2)) return 0;
  if (p.mode == 1) {
    dw_gate_kernel<1><<<grid, 256, 0, st>>>(p.in, p.in_bs, p.w, p.out, p.out_bs, p.hid, nullptr, 0, nullptr, 0, g);
  } else {
    if (p.dg_bs % 4 != 0 || (p.g_out && p.g_bs % 4 != 0)) return 0;
    dw_gate_kernel<2><<<grid, 256, 0, st>>>(p.in, p.in_bs, p.w, p.out, p.out_bs, p.hid, p.dg, p.dg_bs, p.g_out, p.g_bs, g);
  }
  return 1;
}

int dwconv_bwd_fast(const float* in, int64_t in_bs, const float* dout, int64_t dout_bs, const float* w, float* din,
                    int64_t din_bs, float* dw, int B, int Cn, int H, int W, cudaStream_t st) {
  DwGeom g;
  dim3 grid;
  if (!dw_geom(g, grid, B, Cn, H, W, 2)) return 0;
  int ppt = 1;
  if (g.pshift < 0) {   // large planes: up to 4 patches per thread
    ppt = g.ppp / 256;
    if (ppt > 4) ppt = 4;
    if (ppt < 1) ppt = 1;
    grid.x = cdiv(g.ppp, 256 * ppt);
  }
  dw_bwd2_kernel<<<grid, 256, 0, st>>>(in, in_bs, dout, dout_bs, w, din, din_bs, dw, g, ppt);
  return 1;
}

}  // namespace rcot
