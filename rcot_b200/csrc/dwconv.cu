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
#include "gelu.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace rcot {

// Element access for the two storage types of the hidden tensors: fp32, or bf16 in the bf16-storage mode (arithmetic
// is fp32 either way).  4-element row segments are one 16-byte resp. 8-byte access.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
// Row segments of PW output pixels per thread and row.  PW = 4 for both storage types: 8-pixel rows for bf16 (16-byte
// accesses) were measured SLOWER (gate forward 0.51 vs 0.45 ms, fused backward 1.08 vs 0.79 ms at C=96, 128x128, B=32:
// the patch doubles the live registers and halves the CTAs in flight) -- these stencils sit at the CUDA-core issue limit
// once the bytes are halved, so the bf16 mode buys memory, not time, here.
template <typename T> struct RowW { static constexpr int PW = 4; };
__device__ __forceinline__ void ldrow(const float* p, float (&v)[4]) {
  const float4 m = ld4(p);
  v[0] = m.x; v[1] = m.y; v[2] = m.z; v[3] = m.w;
}
__device__ __forceinline__ void ldrow(const __nv_bfloat16* p, float (&v)[4]) {
  const float4 m = ld4(p);
  v[0] = m.x; v[1] = m.y; v[2] = m.z; v[3] = m.w;
}
__device__ __forceinline__ void strow(float* p, const float (&v)[4]) { st4(p, make_float4(v[0], v[1], v[2], v[3])); }
__device__ __forceinline__ void strow(__nv_bfloat16* p, const float (&v)[4]) { st4(p, make_float4(v[0], v[1], v[2], v[3])); }

// the value a store of v will leave in memory (so that sums of squares match what later kernels read)
__device__ __forceinline__ float stored(const float*, float v) { return v; }
__device__ __forceinline__ float stored(const __nv_bfloat16*, float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int ROWS, int PW>
struct Patch {
  float v[ROWS + 2][PW + 2];  // rows y-1 .. y+ROWS, columns x0-1 .. x0+PW
};

template <int ROWS, typename T>
__device__ __forceinline__ void load_patch(Patch<ROWS, RowW<T>::PW>& P, const T* __restrict__ plane, int y, int x0, int H,
                                           int W) {
  constexpr int PW = RowW<T>::PW;
  const T* p = plane + (y - 1) * W + x0;
  if (y > 0 && y + ROWS < H && x0 > 0 && x0 + PW < W) {   // interior: no predicates
#pragma unroll
    for (int r = 0; r < ROWS + 2; ++r) {
      float m[PW];
      ldrow(p + r * W, m);
      P.v[r][0] = ld1(p + r * W - 1);
#pragma unroll
      for (int i = 0; i < PW; ++i) P.v[r][1 + i] = m[i];
      P.v[r][PW + 1] = ld1(p + r * W + PW);
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < ROWS + 2; ++r) {
    const int yy = y - 1 + r;
    if ((unsigned)yy < (unsigned)H) {
      float m[PW];
      ldrow(p + r * W, m);
      P.v[r][0] = x0 > 0 ? ld1(p + r * W - 1) : 0.f;
#pragma unroll
      for (int i = 0; i < PW; ++i) P.v[r][1 + i] = m[i];
      P.v[r][PW + 1] = x0 + PW < W ? ld1(p + r * W + PW) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < PW + 2; ++i) P.v[r][i] = 0.f;
    }
  }
}

template <int ROWS, int PW>
__device__ __forceinline__ void conv_patch(const Patch<ROWS, PW>& P, const float (&w)[9], float (&o)[ROWS][PW]) {
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int j = 0; j < PW; ++j) {
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
  int H, W, pw, ppp, planes, pshift, ROWSv, PWv;   // PWv: pixels per patch row (4: fp32, 8: bf16)
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
  t.x0 = (patch - py * g.pw) * g.PWv;
  return t;
}
__device__ __forceinline__ float warp_sum_dw(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ plain / transposed
template <int ROWS, typename T>
__global__ void __launch_bounds__(256)
    dw_plain_kernel(const T* __restrict__ in, int64_t in_bs, const float* __restrict__ w, T* __restrict__ out,
                    int64_t out_bs, int flip, float* __restrict__ sumsq, int nsq, const DwGeom g) {
  const DwThread t = dw_map(g, ROWS);
  const int HW = g.H * g.W;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = __ldg(w + t.ch * 9 + (flip ? 8 - i : i));
  constexpr int PW = RowW<T>::PW;
  float o[ROWS][PW];
  float sq = 0.f;
  if (t.active) {
    Patch<ROWS, PW> P;
    load_patch<ROWS, T>(P, in + (size_t)t.b * in_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
    conv_patch<ROWS, PW>(P, wk, o);
    T* op = out + (size_t)t.b * out_bs + (size_t)t.ch * HW + t.y * g.W + t.x0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      strow(op + r * g.W, o[r]);
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const float v = stored(op, o[r][j]);
        sq = fmaf(v, v, sq);
      }
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
template <int MODE, int ROWS, typename T, int MINB = 0, bool EARLY = false>
__global__ void __launch_bounds__(256, MINB)
    dw_gate_kernel(const T* __restrict__ in, int64_t in_bs, const float* __restrict__ w, T* __restrict__ out,
                   int64_t out_bs, int hid, const T* __restrict__ dg, int64_t dg_bs, T* __restrict__ g_out,
                   int64_t g_bs, const DwGeom g) {
  const DwThread t = dw_map(g, ROWS);
  if (!t.active) return;
  const int HW = g.H * g.W;
  float w0[9], w1[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    w0[i] = __ldg(w + t.ch * 9 + i);
    w1[i] = __ldg(w + (t.ch + hid) * 9 + i);
  }
  constexpr int PW = RowW<T>::PW;
  const T* inb = in + (size_t)t.b * in_bs;
  float a[ROWS][PW], gt[ROWS][PW];
  const int pix = t.y * g.W + t.x0;
  float dgr[ROWS][PW];
  if (EARLY) {
    // every global request of the thread first (both patches, then the dg rows): more bytes in flight per warp
    Patch<ROWS, PW> P, Q;
    load_patch<ROWS, T>(P, inb + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
    load_patch<ROWS, T>(Q, inb + (size_t)(t.ch + hid) * HW, t.y, t.x0, g.H, g.W);
    if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) ldrow(dg + (size_t)t.b * dg_bs + (size_t)t.ch * HW + pix + r * g.W, dgr[r]);
    }
    conv_patch<ROWS, PW>(P, w0, a);
    conv_patch<ROWS, PW>(Q, w1, gt);
  } else {
    {
      Patch<ROWS, PW> P;
      load_patch<ROWS, T>(P, inb + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
      conv_patch<ROWS, PW>(P, w0, a);
    }
    {
      Patch<ROWS, PW> P;
      load_patch<ROWS, T>(P, inb + (size_t)(t.ch + hid) * HW, t.y, t.x0, g.H, g.W);
      conv_patch<ROWS, PW>(P, w1, gt);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int pr = pix + r * g.W;
    if (MODE == 1) {
      float gv[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) gv[j] = gelu_fast(a[r][j]) * gt[r][j];
      strow(out + (size_t)t.b * out_bs + (size_t)t.ch * HW + pr, gv);
    } else {
      float d[PW];
      if (EARLY) {
#pragma unroll
        for (int j = 0; j < PW; ++j) d[j] = dgr[r][j];
      } else {
        ldrow(dg + (size_t)t.b * dg_bs + (size_t)t.ch * HW + pr, d);
      }
      float da[PW], db[PW], gg[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        float ga, dga;
        gelu_pair(a[r][j], ga, dga);
        da[j] = d[j] * gt[r][j] * dga;
        db[j] = d[j] * ga;
        gg[j] = ga * gt[r][j];
      }
      T* ob = out + (size_t)t.b * out_bs;
      strow(ob + (size_t)t.ch * HW + pr, da);
      strow(ob + (size_t)(t.ch + hid) * HW + pr, db);
      if (g_out) strow(g_out + (size_t)t.b * g_bs + (size_t)t.ch * HW + pr, gg);
    }
  }
}

// ------------------------------------------------------------------ fused backward: din and dW
template <int ROWS, typename T, int MINB = 0, bool EARLYQ = false>
__global__ void __launch_bounds__(256, MINB)
    dw_bwd2_kernel(const T* __restrict__ in, int64_t in_bs, const T* __restrict__ dout, int64_t dout_bs,
                   const float* __restrict__ w, T* __restrict__ din, int64_t din_bs, float* __restrict__ dw,
                   const DwGeom g, const int ppt, const int ipc, const int B) {
  constexpr int PW = RowW<T>::PW;
  DwThread t = dw_map(g, ROWS);
  const int HW = g.H * g.W;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.f;
  float wf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wf[i] = __ldg(w + t.ch * 9 + 8 - i);   // flipped taps
  // The 9 tap sums are reduced once per thread, after it has walked `ppt` patches of a large plane (stride = patches
  // covered by the grid) or, on small planes (ppt == 1), the same patch of `ipc` consecutive images: the warp
  // reductions and atomics cost as much as one patch of work and are amortised that way.
  for (int bi = 0; bi < ipc; ++bi) {
  const int b = blockIdx.z * ipc + bi;
  if (b >= B) break;
  for (int it = 0; it < ppt; ++it) {
    if (it > 0) {
      const int patch = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
      t.active = patch < g.ppp;
      const int py = (t.active ? patch : 0) / g.pw;
      t.y = py * ROWS;
      t.x0 = ((t.active ? patch : 0) - py * g.pw) * PW;
    }
    if (!t.active) continue;
    float d[ROWS][PW];
    Patch<ROWS, PW> Q;
    if (EARLYQ) load_patch<ROWS, T>(Q, in + (size_t)b * in_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
    {
      Patch<ROWS, PW> P;
      load_patch<ROWS, T>(P, dout + (size_t)b * dout_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
      float o[ROWS][PW];
      conv_patch<ROWS, PW>(P, wf, o);
      T* dp = din + (size_t)b * din_bs + (size_t)t.ch * HW + t.y * g.W + t.x0;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        strow(dp + r * g.W, o[r]);
#pragma unroll
        for (int j = 0; j < PW; ++j) d[r][j] = P.v[r + 1][j + 1];   // centre rows of dout
      }
    }
    if (!EARLYQ) load_patch<ROWS, T>(Q, in + (size_t)b * in_bs + (size_t)t.ch * HW, t.y, t.x0, g.H, g.W);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < PW; ++j) acc[ky * 3 + kx] = fmaf(d[r][j], Q.v[r + ky][j + kx], acc[ky * 3 + kx]);
  }
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

// ------------------------------------------------------------------ GDFN middle backward, fused
// Everything between the two 1x1 GEMMs of GDFN's backward in ONE pass over the hidden tensor:
//     a = dw(u_j), b = dw(u_{j+hid});  da = dg*b*gelu'(a), db = dg*gelu(a);
//     du_j = dw^T(da), du_{j+hid} = dw^T(db);  dW_dw += corr(u, [da; db]);  optionally g = gelu(a)*b.
// The two-kernel form (dw_gate<2> then dw_bwd2) writes [da; db] to HBM and reads it and u again; here a CTA owns a
// 32-row band of one channel pair of one image and walks it in 32x32 tiles: u with a 2-pixel halo goes to shared
// memory once (both channels), [da; db] is produced on the tile plus a 1-pixel halo into shared memory (4-pixel
// strips per thread: float4 shared-memory reads, one float4 of dg), and the transposed stencil and the 9-tap
// weight-gradient sums read both from there.  HBM traffic: u and dg in, du out.
constexpr int GF_T = 32;            // tile edge
constexpr int GF_LD = GF_T + 8;     // row stride of both regions: image columns x0-4 .. x0+35 (index = x - x0 + 4)
constexpr int GF_UH = GF_T + 4;     // u region rows y0-2 .. y0+33
constexpr int GF_DH = GF_T + 2;     // [da; db] region rows y0-1 .. y0+32 (columns x0-1 .. x0+32 are used)

__global__ void __launch_bounds__(256, 4)
    gdfn_mid_bwd_kernel(const float* __restrict__ u, int64_t u_bs, const float* __restrict__ dg, int64_t dg_bs,
                        const float* __restrict__ w, float* __restrict__ du, int64_t du_bs, float* __restrict__ dw,
                        float* __restrict__ g_out, int64_t g_bs, int hid, int H, int W) {
  __shared__ __align__(16) float su[2][GF_UH][GF_LD];
  __shared__ __align__(16) float sd[2][GF_DH][GF_LD];
  __shared__ float red[18][8];
  const int tid = threadIdx.x;
  const int ch = blockIdx.y, b = blockIdx.z;
  const int y0 = blockIdx.x * GF_T;
  const int th = min(GF_T, H - y0);
  const int HW = H * W;
  const float* ua = u + (size_t)b * u_bs + (size_t)ch * HW;
  const float* ub = ua + (size_t)hid * HW;
  const float* dgp = dg + (size_t)b * dg_bs + (size_t)ch * HW;
  float* dua = du + (size_t)b * du_bs + (size_t)ch * HW;
  float* dub = dua + (size_t)hid * HW;
  float* gp = g_out ? g_out + (size_t)b * g_bs + (size_t)ch * HW : nullptr;
  float wk[2][9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    wk[0][i] = __ldg(w + ch * 9 + i);
    wk[1][i] = __ldg(w + (ch + hid) * 9 + i);
  }
  float acc[2][9];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[c][i] = 0.f;
  const int ty = tid >> 3, k4 = (tid & 7) * 4;   // stencil role: row ty, columns k4 .. k4+3 of the tile
  const int nrow_u = th + 4, nrow_d = th + 2;
  const int n_strip = nrow_d * 8;                // step 2: eight 4-pixel strips per row ...
  const int edge0 = (n_strip + 31) & ~31;        // ... then (warp aligned) one task per row for the two halo columns
  // Six values (columns x-1 .. x+4) around this lane's 4-pixel strip of one region row.  The eight lanes of a row
  // hold neighbouring strips, so the two edge values come from the neighbours' float4 by warp shuffle (a scalar
  // shared-memory read at stride 4 words is a 4-way bank conflict); only strips 0 and 7 read their outer edge.
  // Must be called by all 32 lanes.
  const int ls = tid & 7;
  auto row6 = [&](const float* row, int sx, float (&v)[6]) {
    const float4 m = *reinterpret_cast<const float4*>(row + sx + 4);
    float l = __shfl_up_sync(0xffffffffu, m.w, 1);
    float r = __shfl_down_sync(0xffffffffu, m.x, 1);
    if (ls == 0) l = row[3];
    if (ls == 7) r = row[GF_T + 4];
    v[0] = l; v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w; v[5] = r;
  };
  for (int x0 = 0; x0 < W; x0 += GF_T) {
    __syncthreads();                             // previous tile fully consumed
    // ---- 1: u (both channels) with a 2-pixel halo, zero outside the image
    for (int e = tid; e < nrow_u * (GF_LD / 4); e += 256) {
      const int r = e / (GF_LD / 4), q = e - r * (GF_LD / 4);
      const int y = y0 - 2 + r, x = x0 - 4 + 4 * q;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
        va = __ldg(reinterpret_cast<const float4*>(ua + y * W + x));
        vb = __ldg(reinterpret_cast<const float4*>(ub + y * W + x));
      }
      *reinterpret_cast<float4*>(&su[0][r][4 * q]) = va;
      *reinterpret_cast<float4*>(&su[1][r][4 * q]) = vb;
    }
    __syncthreads();
    // ---- 2: [da; db] on the tile plus a 1-pixel halo (zero outside the image)
    for (int t = tid; t < edge0 + nrow_d; t += 256) {
      if ((t & ~31) < n_strip) {                 // warp-uniform: whole warps run the strip code (shuffles inside)
        const bool mine = t < n_strip;
        const int r = mine ? (t >> 3) : nrow_d - 1, sx = ls * 4;   // region row, tile column of the strip
        const int y = y0 - 1 + r;
        const bool in = mine && (unsigned)y < (unsigned)H;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float v0[6], v1[6];
          row6(&su[0][r + ky][0], sx, v0);
          row6(&su[1][r + ky][0], sx, v1);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a[j] = fmaf(v0[j + kx], wk[0][ky * 3 + kx], a[j]);
              bb[j] = fmaf(v1[j + kx], wk[1][ky * 3 + kx], bb[j]);
            }
        }
        float4 da4 = make_float4(0.f, 0.f, 0.f, 0.f), db4 = da4;
        if (in) {
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(dgp + y * W + x0 + sx));
          const float d[4] = {d4.x, d4.y, d4.z, d4.w};
          float da[4], db[4], gg[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float ge, dge;
            gelu_pair(a[j], ge, dge);
            da[j] = d[j] * bb[j] * dge;
            db[j] = d[j] * ge;
            gg[j] = ge * bb[j];
          }
          da4 = make_float4(da[0], da[1], da[2], da[3]);
          db4 = make_float4(db[0], db[1], db[2], db[3]);
          if (gp && r >= 1 && r <= th)
            *reinterpret_cast<float4*>(gp + y * W + x0 + sx) = make_float4(gg[0], gg[1], gg[2], gg[3]);
        }
        if (mine) {
          *reinterpret_cast<float4*>(&sd[0][r][sx + 4]) = da4;
          *reinterpret_cast<float4*>(&sd[1][r][sx + 4]) = db4;
        }
      } else if (t >= edge0) {
        const int r = t - edge0;
        const int y = y0 - 1 + r;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int ci = side == 0 ? 3 : GF_T + 4;        // region column index of x0-1 / x0+32
          const int x = x0 - 4 + ci;
          float da = 0.f, db = 0.f;
          if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) {
            float a = 0.f, bb = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                a = fmaf(su[0][r + ky][ci - 1 + kx], wk[0][ky * 3 + kx], a);
                bb = fmaf(su[1][r + ky][ci - 1 + kx], wk[1][ky * 3 + kx], bb);
              }
            const float d = __ldg(dgp + y * W + x);
            float ge, dge;
            gelu_pair(a, ge, dge);
            da = d * bb * dge;
            db = d * ge;
          }
          sd[0][r][ci] = da;
          sd[1][r][ci] = db;
        }
      }
    }
    __syncthreads();
    // ---- 3: du = dw^T([da; db]) and the tap sums of dW, a 4-pixel strip per thread and channel
    // (th % 4 == 0 is checked by the launcher, so the four rows of a warp are all inside or all outside)
    if (ty < th) {
      const int y = y0 + ty;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float D[3][6];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) row6(&sd[c][ty + ky][0], k4, D[ky]);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) s = fmaf(D[ky][j + kx], wk[c][8 - (ky * 3 + kx)], s);
          o[j] = s;
        }
        *reinterpret_cast<float4*>((c == 0 ? dua : dub) + y * W + x0 + k4) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          float Q[6];
          row6(&su[c][ty + 1 + ky][0], k4, Q);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[c][ky * 3 + kx] = fmaf(D[1][j + 1], Q[j + kx], acc[c][ky * 3 + kx]);
        }
      }
    }
  }
  // ---- 18 tap sums of the band: warp shuffles, then one atomicAdd per tap
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float s = warp_sum_dw(acc[c][i]);
      if (lane == 0) red[c * 9 + i][wid] = s;
    }
  __syncthreads();
  if (tid < 18) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[tid][k];
    atomicAdd(dw + (tid < 9 ? ch * 9 + tid : (ch + hid) * 9 + tid - 9), s);
  }
}

// Returns 1 if the fused kernel handled the call, 0 if the geometry does not qualify.
int gdfn_mid_bwd_fast(const float* u, int64_t u_bs, const float* dg, int64_t dg_bs, const float* w, float* du,
                      int64_t du_bs, float* dw, float* g_out, int64_t g_bs, int B, int hid, int H, int W,
                      cudaStream_t st) {
  if (W % GF_T != 0 || H % 4 != 0 || B > 65535 || hid > 65535) return 0;
  if (u_bs % 4 != 0 || du_bs % 4 != 0 || dg_bs % 4 != 0 || ((uintptr_t)u % 16 != 0) || ((uintptr_t)du % 16 != 0) ||
      ((uintptr_t)dg % 16 != 0) || (g_out && (g_bs % 4 != 0 || (uintptr_t)g_out % 16 != 0)))
    return 0;
  dim3 grid(cdiv(H, GF_T), hid, B);
  gdfn_mid_bwd_kernel<<<grid, 256, 0, st>>>(u, u_bs, dg, dg_bs, w, du, du_bs, dw, g_out, g_bs, hid, H, W);
  return 1;
}

// ------------------------------------------------------------------ launch geometry
static int dw_rows_forced() {   // 0: per-kernel default
  static int pref = -1;
  if (pref < 0) {
    const char* e = getenv("RCOT_DW_ROWS");
    pref = !e ? 0 : (e[0] == '2' ? 2 : (e[0] == '4' ? 4 : 0));
  }
  return pref;
}
static bool dw_geom(DwGeom& g, dim3& grid, int B, int planes, int H, int W, int rows, int PW = 4) {
  if (W % PW != 0 || H % rows != 0 || B > 65535) return false;
  g.H = H;
  g.W = W;
  g.PWv = PW;
  g.pw = W / PW;
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

template <typename T>
static int dwconv_fast_t(const rcot_dw_params& p, int planes, cudaStream_t st) {
  constexpr int PWT = RowW<T>::PW;
  const bool al = p.in_bs % PWT == 0 && p.out_bs % PWT == 0 && ((uintptr_t)p.in % 16 == 0) && ((uintptr_t)p.out % 16 == 0) &&
                  (PWT == 4 || (p.dg_bs % PWT == 0 && p.g_bs % PWT == 0));
  if (!al) return 0;
  const T* in = reinterpret_cast<const T*>(p.in);
  T* out = reinterpret_cast<T*>(p.out);
  const T* dg = reinterpret_cast<const T*>(p.dg);
  T* g_out = reinterpret_cast<T*>(p.g_out);
  DwGeom g;
  dim3 grid;
  if (p.mode == 0) {
    if (dw_geom(g, grid, p.B, planes, p.H, p.W, 4, RowW<T>::PW)) {
      dw_plain_kernel<4, T><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.flip, p.sumsq, p.nsq, g);
      return 1;
    }
    if (dw_geom(g, grid, p.B, planes, p.H, p.W, 2, RowW<T>::PW)) {
      dw_plain_kernel<2, T><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.flip, p.sumsq, p.nsq, g);
      return 1;
    }
    return 0;
  }
  // 4x4 output patches (halo rows re-read 1.5x instead of 2x) pay off for the gate forward (measured 374 vs 422 us at
  // C=96, 128x128, B=32); the gate backward holds three more planes in registers and is faster with 4x2 patches
  // (557 vs 695 us).  RCOT_DW_ROWS=2|4 forces one shape for both (A/B switch).
  const int want = dw_rows_forced() ? dw_rows_forced() : (p.mode == 1 ? 4 : 2);
  const int rows = (want == 4 && p.H % 4 == 0 && dw_geom(g, grid, p.B, planes, p.H, p.W, 4, RowW<T>::PW)) ? 4 : 2;
  if (rows == 2 && !dw_geom(g, grid, p.B, planes, p.H, p.W, 2, RowW<T>::PW)) return 0;
  if (p.mode == 1) {
    // gate forward: four CTAs per SM (64 registers) 376 -> 366 us at 510x128x128; requesting both patches up front LOSES
    // (448 us: more registers, fewer resident warps) -- scripts/bench_dw.py
    static int variant1 = -1;         // A/B switch: RCOT_DW_GATE1_VAR = 10 * MINB + EARLY (default 40)
    if (variant1 < 0) {
      const char* e = getenv("RCOT_DW_GATE1_VAR");
      variant1 = e ? atoi(e) : 40;
    }
#define RCOT_GATE1(MB, EA) dw_gate_kernel<1, 4, T, MB, EA><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.hid, nullptr, 0, nullptr, 0, g)
    if (rows == 4) {
      switch (variant1) {
        case 11: RCOT_GATE1(1, true); break;
        case 21: RCOT_GATE1(2, true); break;
        case 30: RCOT_GATE1(3, false); break;
        case 31: RCOT_GATE1(3, true); break;
        case 10: RCOT_GATE1(0, false); break;
        default: RCOT_GATE1(4, false); break;
      }
    }
#undef RCOT_GATE1
    else
      dw_gate_kernel<1, 2, T><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.hid, nullptr, 0, nullptr, 0, g);
  } else {
    if (p.dg_bs % 4 != 0 || (p.g_out && p.g_bs % 4 != 0)) return 0;
    static int variant = -1;          // A/B switch: RCOT_DW_GATE_VAR = 10 * MINB + EARLY (default 10)
    if (variant < 0) {
      const char* e = getenv("RCOT_DW_GATE_VAR");
      variant = e ? atoi(e) : 10;
    }
#define RCOT_GATE2(MB, EA) dw_gate_kernel<2, 2, T, MB, EA><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.hid, dg, p.dg_bs, g_out, p.g_bs, g)
    if (rows == 4)
      dw_gate_kernel<2, 4, T><<<grid, 256, 0, st>>>(in, p.in_bs, p.w, out, p.out_bs, p.hid, dg, p.dg_bs, g_out, p.g_bs, g);
    else {
      switch (variant) {
        case 11: RCOT_GATE2(1, true); break;
        case 21: RCOT_GATE2(2, true); break;
        case 31: RCOT_GATE2(3, true); break;
        case 41: RCOT_GATE2(4, true); break;
        case 50: RCOT_GATE2(5, false); break;
        case 51: RCOT_GATE2(5, true); break;
        default: RCOT_GATE2(0, false); break;
      }
    }
#undef RCOT_GATE2
  }
  return 1;
}

// Returns 1 if the aligned fast path handled the call, 0 if the caller must use the generic kernels.
int dwconv_fast(const rcot_dw_params& p, int planes, cudaStream_t st) {
  return p.bf16 ? dwconv_fast_t<__nv_bfloat16>(p, planes, st) : dwconv_fast_t<float>(p, planes, st);
}

template <typename T>
static int dwconv_bwd_fast_t(const T* in, int64_t in_bs, const T* dout, int64_t dout_bs, const float* w, T* din,
                             int64_t din_bs, float* dw, int B, int Cn, int H, int W, cudaStream_t st) {
  DwGeom g;
  dim3 grid;
  const int want = dw_rows_forced() ? dw_rows_forced() : 4;   // 4x4 patches: 669 vs 755 us at C=96, 128x128, B=32
  const int rows = (want == 4 && H % 4 == 0 && dw_geom(g, grid, B, Cn, H, W, 4, RowW<T>::PW)) ? 4 : 2;
  if (rows == 2 && !dw_geom(g, grid, B, Cn, H, W, 2, RowW<T>::PW)) return 0;
  int ppt = 1;
  if (g.pshift < 0) {   // large planes: up to 4 patches per thread
    ppt = g.ppp / 256;
    if (ppt > 4) ppt = 4;
    if (ppt < 1) ppt = 1;
    grid.x = cdiv(g.ppp, 256 * ppt);
  }
  int ipc = 1;          // small planes: up to 4 images per thread while >= 8 CTAs per SM remain
  while (ppt == 1 && ipc < 4 && (long)grid.x * grid.y * cdiv(B, ipc * 2) >= 8 * 148) ipc *= 2;
  grid.z = cdiv(B, ipc);
  // Measured at 510 x 128 x 128, B=32 (scripts/bench_dw.py): the kernel is latency-bound (ncu: 24 % warps active at 95
  // registers, 41 % long-scoreboard stalls), so occupancy pays: __launch_bounds__(256, 3) = 80 registers, three CTAs
  // per SM: 662 -> 593 us (5.4 TB/s); four CTAs (64 registers, spills) 691 us; loading both patches up front 860 us.
  static int variant = -1;            // A/B switch: RCOT_DW_BWD_VAR = 10 * MINB + EARLYQ  (default 30)
  if (variant < 0) {
    const char* e = getenv("RCOT_DW_BWD_VAR");
    variant = e ? atoi(e) : 30;
  }
#define RCOT_BWD2(MB, EQ) dw_bwd2_kernel<4, T, MB, EQ><<<grid, 256, 0, st>>>(in, in_bs, dout, dout_bs, w, din, din_bs, dw, g, ppt, ipc, B)
  if (rows == 4) {
    switch (variant) {
      case 11: RCOT_BWD2(1, true); break;
      case 20: RCOT_BWD2(2, false); break;
      case 21: RCOT_BWD2(2, true); break;
      case 31: RCOT_BWD2(3, true); break;
      case 40: RCOT_BWD2(4, false); break;
      case 41: RCOT_BWD2(4, true); break;
      case 10: RCOT_BWD2(0, false); break;
      default: RCOT_BWD2(3, false); break;
    }
  }
#undef RCOT_BWD2
  else
    dw_bwd2_kernel<2, T><<<grid, 256, 0, st>>>(in, in_bs, dout, dout_bs, w, din, din_bs, dw, g, ppt, ipc, B);
  return 1;
}

int dwconv_bwd_fast(const void* in, int64_t in_bs, const void* dout, int64_t dout_bs, const float* w, void* din,
                    int64_t din_bs, float* dw, int B, int Cn, int H, int W, int bf16, cudaStream_t st) {
  if (bf16)
    return dwconv_bwd_fast_t<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(in), in_bs,
                                            reinterpret_cast<const __nv_bfloat16*>(dout), dout_bs, w,
                                            reinterpret_cast<__nv_bfloat16*>(din), din_bs, dw, B, Cn, H, W, st);
  return dwconv_bwd_fast_t<float>(reinterpret_cast<const float*>(in), in_bs, reinterpret_cast<const float*>(dout), dout_bs,
                                  w, reinterpret_cast<float*>(din), din_bs, dw, B, Cn, H, W, st);
}

}  // namespace rcot
