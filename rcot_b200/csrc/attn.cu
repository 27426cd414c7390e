// attn.cu -- the small-matrix steps of MDTA channel attention (Net_Restormer.py:39-49), one CTA per
// (head, image), everything c x c (c = C/heads in {24,48,96}) in shared memory:
//   forward : Gt = G / (|q| |k|^T), A = softmax_rows(Gt * temperature), M = W_out * blockdiag(A)
//             -> M (and M^T) written as per-image packed tcgen05 B operands, so that
//             project_out(attn @ v) becomes ONE pixel-as-M GEMM  y = x + M v.
//   backward: from P = dy v^T  ->  dW_out, dtemperature, and the per-image packed matrix
//             W12 = [[diag(c_q), B_q], [B_q^T, diag(c_k)]] with  [dq; dk] = W12 [q; k]
//             (softmax backward, L2-normalisation backward folded into c_q, c_k).
#include "../../include/rcot_b200.h"
#include "common.cuh"
#include "tc.cuh"
#include <stdlib.h>

namespace rcot {

__device__ __forceinline__ float warp_sum_a(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_a(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Packed-operand geometry of an [N x K] matrix (tc.cuh packed_offset), with the divisions by runtime values done
// once per kernel instead of once per store: passes of BN columns, nk K-chunks, `tile` bytes per (pass, chunk, term).
struct PackGeom {
  int BN, nk;
  uint32_t tile;
  __device__ PackGeom(int N, int K) {
    const int nst = (N + 255) / 256;
    BN = ((N + nst - 1) / nst + 15) / 16 * 16;
    nk = (K + KC - 1) / KC;
    tile = op_tile_bytes(BN);
  }
  // byte offset of element (n, k) in the hi image; the lo image follows at +tile.  N <= 768 here: at most 3 passes.
  __device__ __forceinline__ size_t hi(int n, int k) const {
    const int pass = (n >= 2 * BN) ? 2 : (n >= BN ? 1 : 0);
    const int np = n - pass * BN, kc = k >> 5, kp = k & 31;
    return ((size_t)(pass * nk + kc) * 2) * tile + op_offset(np, kp);
  }
};
__device__ __forceinline__ void store_split(uint8_t* base, const PackGeom& pg, int n, int k, float w) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t o = pg.hi(n, k);
  *reinterpret_cast<__nv_bfloat16*>(base + o) = hi;
  *reinterpret_cast<__nv_bfloat16*>(base + o + pg.tile) = lo;
}

// 8 consecutive k (k0 % 8 == 0) of row n: one 16-byte core-matrix row in the hi image and one in the lo image.
__device__ __forceinline__ void store_split8(uint8_t* base, const PackGeom& pg, int n, int k0, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  const size_t o = pg.hi(n, k0);
  *reinterpret_cast<uint4*>(base + o) = hi;
  *reinterpret_cast<uint4*>(base + o + pg.tile) = lo;
}

// Forward.  grid = (heads, B, nsplit): every CTA of a (head, image) recomputes the c x c softmax (cheap) and
// produces `nco` rows of M = W_out * blockdiag(A); split 0 also writes A and the normalised Gram.
__global__ void __launch_bounds__(256) attn_fwd_kernel(const rcot_attn_params p, const int nco) {
  extern __shared__ __align__(16) float sm[];
  const int h = blockIdx.x, b = blockIdx.y, c = p.C / p.heads, C = p.C;
  float* sA = sm;            // [c*c]
  float* snq = sm + c * c;   // [c]
  float* snk = snq + c;      // [c]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int co_begin = blockIdx.z * nco;
  const int co_end = min(C, co_begin + nco);
  if (co_begin >= co_end) return;
  const bool first = blockIdx.z == 0;
  const float* Gb = p.G + ((size_t)b * p.heads + h) * c * c;
  const float* ss = p.sumsq + (size_t)b * 2 * C;
  for (int i = tid; i < c; i += blockDim.x) {
    snq[i] = fmaxf(sqrtf(ss[h * c + i]), 1e-12f);
    snk[i] = fmaxf(sqrtf(ss[C + h * c + i]), 1e-12f);
  }
  __syncthreads();
  const float tau = __ldg(p.temperature + h);
  float* Ab = p.A + ((size_t)b * p.heads + h) * c * c;
  float* Gtb = p.Gt + ((size_t)b * p.heads + h) * c * c;
  for (int i = warp; i < c; i += nwarp) {  // one warp per row
    float mx = -INFINITY;
    for (int j = lane; j < c; j += 32) {
      const float gt = __ldg(Gb + i * c + j) / (snq[i] * snk[j]);
      if (first) Gtb[i * c + j] = gt;
      const float l = gt * tau;
      sA[i * c + j] = l;
      mx = fmaxf(mx, l);
    }
    mx = warp_max_a(mx);
    float sum = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float e = expf(sA[i * c + j] - mx);
      sA[i * c + j] = e;
      sum += e;
    }
    sum = warp_sum_a(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < c; j += 32) {
      const float a = sA[i * c + j] * inv;
      sA[i * c + j] = a;
      if (first) Ab[i * c + j] = a;
    }
  }
  __syncthreads();
  // M[co, (h,j)] = sum_i W_out[co, (h,i)] * A[i,j] for this split's rows -> shared memory (four j per thread:
  // one float4 of A per W value), then 16-byte packed stores
  float* sM = snk + c;       // [nco * c]
  const int c4 = c >> 2, rows = co_end - co_begin;
  for (int t = tid; t < rows * c4; t += blockDim.x) {
    const int col = t / c4, j0 = (t - col * c4) * 4;
    const float* wrow = p.w_out + (size_t)(co_begin + col) * C + h * c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < c; i += 4) {                       // c % 8 == 0: one 16-byte load of W_out per four rows of A
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + i));
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(sA + (i + k) * c + j0);
        acc.x = fmaf(wv[k], a.x, acc.x);
        acc.y = fmaf(wv[k], a.y, acc.y);
        acc.z = fmaf(wv[k], a.z, acc.z);
        acc.w = fmaf(wv[k], a.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(sM + col * c + j0) = acc;
  }
  __syncthreads();
  uint8_t* Mp = reinterpret_cast<uint8_t*>(p.Mpack) + (size_t)b * p.pack_bs;
  uint8_t* MTp = p.MTpack ? reinterpret_cast<uint8_t*>(p.MTpack) + (size_t)b * p.pack_bs : nullptr;
  const int c8 = c >> 3;
  const PackGeom pgM(C, C);
  for (int t = tid; t < rows * c8; t += blockDim.x) {       // M: row n = co, 8 consecutive k = h*c + j
    const int col = t / c8, j0 = (t - col * c8) * 8;
    store_split8(Mp, pgM, co_begin + col, h * c + j0, sM + col * c + j0);
  }
  if (MTp) {
    const int r8 = rows >> 3;                               // rows is a multiple of 8 (launcher)
    for (int t = tid; t < c * r8; t += blockDim.x) {        // M^T: row n = h*c + j, 8 consecutive k = co
      const int col0 = (t / c) * 8, j = t - (t / c) * c;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = sM[(col0 + i) * c + j];
      store_split8(MTp, pgM, h * c + j, co_begin + col0, v);
    }
  }
}

constexpr int AB_CH = 32;  // rows of W_out / P per phase-1 CTA
constexpr int AB_T = 6;    // largest register tile edge: ceil(96 / 16)

// Backward, phase 1.  grid = (heads, B, ceil(C / AB_CH)): a CTA stages AB_CH rows of W_out / P (head block) and runs
// two register-tiled products:  dA[i,j] += sum_co W_out[co,(h,i)] P[co,(h,j)]  (T x T outputs per thread,
// T = ceil(c/16); partial over this CTA's rows -> atomics into the zeroed scratch dA) and
// dW_out[co,(h,i)] += sum_j P[co,(h,j)] A[i,j]  (2 x T outputs per thread, complete for these rows).
// One CTA per (head, image) for all of it ran at < 1 IPC on 32 SMs; split by rows it fills the machine.
template <int AB_TT>   // register tile edge ceil(c / 16), compile time: 2 (c <= 32), 3 (c <= 48), 6 (c <= 96)
__global__ void __launch_bounds__(256) attn_bwd_p1_kernel(const rcot_attn_params p, const int ipc) {
  extern __shared__ __align__(16) float sm[];
  const int h = blockIdx.x, c = p.C / p.heads, C = p.C;
  const int b0 = blockIdx.y * ipc;        // this CTA's images b0 .. b0+ipc-1 (W_out rows staged once for all of them)
  const int co0 = blockIdx.z * AB_CH;
  const int ca = c + 1;         // padded row stride of sA: threads that differ in the row hit different banks
  float* sA = sm;               // [c*ca] softmax probabilities A[i][j] of the current image
  float* sW = sA + c * ca;      // [AB_CH*c] rows of W_out[:, head block]
  float* sP = sW + AB_CH * c;   // [AB_CH*c] rows of P[:, head block] of the current image
  const int tid = threadIdx.x;
  constexpr int T = AB_TT;               // register tile edge: T * 16 >= c
  const int nt = (c + T - 1) / T;         // tiles per dimension (<= 16)
  // dA role: thread (ti, tj) owns rows ti*T.., columns tj*T.. of dA
  const int ti = tid / nt, tj = tid - ti * nt;
  const bool da_on = ti < nt;
  int ia[AB_TT], ja[AB_TT];
#pragma unroll
  for (int q = 0; q < AB_TT; ++q) {
    ia[q] = min(ti * T + q, c - 1);       // clamped: duplicates are never stored
    ja[q] = min(tj * T + q, c - 1);
  }
  // dW_out role: thread (rg, ig) owns staged rows 2rg, 2rg+1 and head columns ig*T..
  const int rg = tid >> 4, ig = tid & 15;
  int iw[AB_TT];
#pragma unroll
  for (int q = 0; q < AB_TT; ++q) iw[q] = min(ig * T + q, c - 1) * ca;
  for (int r = tid / c, i = tid - (tid / c) * c; r < AB_CH;) {   // (r, i) walk without a division per element
    sW[r * c + i] = (co0 + r < C) ? __ldg(p.w_out + (size_t)(co0 + r) * C + h * c + i) : 0.f;   // partial chunk: zeros
    i += 256;
    while (i >= c) {
      i -= c;
      ++r;
    }
  }
  float aw[2][AB_TT];            // dW_out partial sums, accumulated over this CTA's images: one atomic per entry
#pragma unroll
  for (int q = 0; q < AB_TT; ++q) aw[0][q] = aw[1][q] = 0.f;
  for (int bi = 0; bi < ipc; ++bi) {
    const int b = b0 + bi;
    if (b >= p.B) break;
    const size_t hb = ((size_t)b * p.heads + h) * c * c;
    const float* Pm = p.P + (size_t)b * C * C;  // [co, ci]
    __syncthreads();                          // previous image's sA / sP fully consumed
    for (int i = tid / c, j = tid - (tid / c) * c; i < c;) {
      sA[i * ca + j] = __ldg(p.A + hb + i * c + j);
      j += 256;
      while (j >= c) {
        j -= c;
        ++i;
      }
    }
    for (int r = tid / c, i = tid - (tid / c) * c; r < AB_CH;) {
      sP[r * c + i] = (co0 + r < C) ? __ldg(Pm + (size_t)(co0 + r) * C + h * c + i) : 0.f;
      i += 256;
      while (i >= c) {
        i -= c;
        ++r;
      }
    }
    __syncthreads();
    if (da_on) {
      float acc[AB_TT][AB_TT];
#pragma unroll
      for (int x = 0; x < AB_TT; ++x)
#pragma unroll
        for (int y = 0; y < AB_TT; ++y) acc[x][y] = 0.f;
#pragma unroll 4
      for (int r = 0; r < AB_CH; ++r) {
        float wv[AB_TT], pv[AB_TT];
#pragma unroll
        for (int q = 0; q < AB_TT; ++q)
          if (q < T) {
            wv[q] = sW[r * c + ia[q]];
            pv[q] = sP[r * c + ja[q]];
          }
#pragma unroll
        for (int x = 0; x < AB_TT; ++x)
#pragma unroll
          for (int y = 0; y < AB_TT; ++y)
            if (x < T && y < T) acc[x][y] = fmaf(wv[x], pv[y], acc[x][y]);
      }
      float* dA = p.dA + hb;
#pragma unroll
      for (int x = 0; x < AB_TT; ++x)
#pragma unroll
        for (int y = 0; y < AB_TT; ++y)
          if (x < T && y < T && ti * T + x < c && tj * T + y < c) atomicAdd(dA + (ti * T + x) * c + tj * T + y, acc[x][y]);
    }
    const float* p0 = sP + (2 * rg) * c;
    const float* p1 = p0 + c;
#pragma unroll 4
    for (int j = 0; j < c; ++j) {
      const float x0 = p0[j], x1 = p1[j];
#pragma unroll
      for (int q = 0; q < AB_TT; ++q)
        if (q < T) {
          const float a = sA[iw[q] + j];
          aw[0][q] = fmaf(x0, a, aw[0][q]);
          aw[1][q] = fmaf(x1, a, aw[1][q]);
        }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int co = co0 + 2 * rg + k;
    if (co < C) {
#pragma unroll
      for (int q = 0; q < AB_TT; ++q)
        if (q < T && ig * T + q < c) atomicAdd(p.dw_out + (size_t)co * C + h * c + ig * T + q, aw[k][q]);
    }
  }
}

// Backward, phase 2.  One CTA per (head, image): softmax / normalisation backward on the c x c tile (dA complete in
// global memory) and the packed W12 stores.
__global__ void __launch_bounds__(256) attn_bwd_p2_kernel(const rcot_attn_params p) {
  extern __shared__ __align__(16) float sm[];
  const int h = blockIdx.x, b = blockIdx.y, c = p.C / p.heads, C = p.C;
  float* sA = sm;               // [c*c] softmax probabilities
  float* sG = sA + c * c;       // [c*c] normalised Gram
  float* sD = sG + c * c;       // [c*c] dA -> dGt -> B_q
  float* snq = sD + c * c;      // [c]
  float* snk = snq + c;
  float* srq = snk + c;
  float* srk = srq + c;
  __shared__ float s_dtau;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const size_t hb = ((size_t)b * p.heads + h) * c * c;
  const float* ss = p.sumsq + (size_t)b * 2 * C;
  for (int e = tid; e < c * c; e += blockDim.x) {
    sA[e] = __ldg(p.A + hb + e);
    sG[e] = __ldg(p.Gt + hb + e);
    sD[e] = p.dA[hb + e];
  }
  for (int i = tid; i < c; i += blockDim.x) {
    snq[i] = sqrtf(ss[h * c + i]);
    snk[i] = sqrtf(ss[C + h * c + i]);
  }
  if (tid == 0) s_dtau = 0.f;
  __syncthreads();
  const float tau = __ldg(p.temperature + h);
  float dtau = 0.f;
  for (int i = warp; i < c; i += nwarp) {
    float rd = 0.f;
    for (int j = lane; j < c; j += 32) rd = fmaf(sD[i * c + j], sA[i * c + j], rd);
    rd = warp_sum_a(rd);
    float rq = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float dS = sA[i * c + j] * (sD[i * c + j] - rd);
      dtau = fmaf(dS, sG[i * c + j], dtau);
      const float dG = dS * tau;
      sD[i * c + j] = dG;
      rq = fmaf(dG, sG[i * c + j], rq);
    }
    rq = warp_sum_a(rq);
    if (lane == 0) srq[i] = rq;
  }
  dtau = warp_sum_a(dtau);
  if (lane == 0) atomicAdd(&s_dtau, dtau);
  __syncthreads();
  if (tid == 0) atomicAdd(p.dtemperature + h, s_dtau);
  for (int j = tid; j < c; j += blockDim.x) {               // column sums r_k[j] = sum_i dGt[i,j] * Gt[i,j]
    float t = 0.f;
    for (int i = 0; i < c; ++i) t = fmaf(sD[i * c + j], sG[i * c + j], t);
    srk[j] = t;
  }
  __syncthreads();
  // packed W12: rows [0,C) produce dq, rows [C,2C) produce dk; K runs over [q channels ; k channels].
  // Only the head-diagonal blocks are ever written: the caller keeps one zero-initialised pack
  // buffer per (C, heads) configuration, so every other entry stays zero.
  uint8_t* Wp = reinterpret_cast<uint8_t*>(p.W12pack) + (size_t)b * p.pack12_bs;
  const PackGeom pgW(2 * C, 2 * C);
  for (int i = warp; i < c; i += nwarp) {                   // B_q = dGt / (|q_i| |k_j|), in place
    const float qi = fmaxf(snq[i], 1e-12f);
    for (int j = lane; j < c; j += 32) sD[i * c + j] = sD[i * c + j] / (qi * fmaxf(snk[j], 1e-12f));
  }
  __syncthreads();
  const int c8 = c >> 3;
  for (int t = tid; t < c * c8; t += blockDim.x) {
    {   // dq rows: n = h*c + i, 8 consecutive k = C + h*c + j
      const int i = t / c8, j0 = (t - i * c8) * 8;
      store_split8(Wp, pgW, h * c + i, C + h * c + j0, sD + i * c + j0);
    }
    {   // dk rows: n = C + h*c + j, 8 consecutive k = h*c + i
      const int j = t % c, i0 = (t / c) * 8;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = sD[(i0 + q) * c + j];
      store_split8(Wp, pgW, C + h * c + j, h * c + i0, v);
    }
  }
  // (the diagonal entries below live in the q->dq and k->dk blocks, which nothing above writes)
  for (int i = tid; i < c; i += blockDim.x) {
    // d/dq of q/max(|q|,eps): the projection term vanishes when the clamp is active
    const float cq = snq[i] >= 1e-12f ? -srq[i] / (snq[i] * snq[i]) : 0.f;
    const float ck = snk[i] >= 1e-12f ? -srk[i] / (snk[i] * snk[i]) : 0.f;
    store_split(Wp, pgW, h * c + i, h * c + i, cq);
    store_split(Wp, pgW, C + h * c + i, C + h * c + i, ck);
  }
}

}  // namespace rcot

using namespace rcot;

static int attn_check(const rcot_attn_params& p) {
  RCOT_REQUIRE(p.B > 0 && p.B <= 65535 && p.C > 0 && p.heads > 0 && p.C % p.heads == 0, "attn: bad sizes");
  RCOT_REQUIRE(p.C / p.heads <= 96, "attn: per-head channels must be <= 96 (got %d)", p.C / p.heads);
  RCOT_REQUIRE(p.C <= 384, "attn: at most 384 channels (packed [2C x 2C] operand of <= 3 passes), got %d", p.C);
  RCOT_REQUIRE(p.sumsq && p.temperature && p.w_out && p.A && p.Gt, "attn: null pointer");
  return RCOT_OK;
}

extern "C" int rcot_attn_fwd(const rcot_attn_params* pp, rcot_stream_t st) {
  RCOT_REQUIRE(pp != nullptr, "attn_fwd: null params");
  const rcot_attn_params& p = *pp;
  int rc = attn_check(p);
  if (rc) return rc;
  RCOT_REQUIRE(p.G && p.Mpack, "attn_fwd: null pointer");
  const int c = p.C / p.heads;
  RCOT_REQUIRE(c % 8 == 0 && p.C % 8 == 0, "attn_fwd: channels per head must be a multiple of 8");
  // rows of M per CTA: enough CTAs to cover the 148 SMs, a multiple of 8 rows each (packed 16-byte stores)
  static int target = -1;             // A/B switch RCOT_ATTN_FWD_CTAS: CTAs the grid should reach (default 592 = 4 per SM)
  if (target < 0) {
    const char* e = getenv("RCOT_ATTN_FWD_CTAS");
    target = e ? atoi(e) : 592;
  }
  int nsplit = cdiv(target, (long)p.heads * p.B);
  if (nsplit > p.C / 8) nsplit = p.C / 8;
  if (nsplit < 1) nsplit = 1;
  const int nco = round_up(cdiv(p.C, nsplit), 8);
  nsplit = cdiv(p.C, nco);
  const size_t smem = ((size_t)c * c + 2 * c + (size_t)nco * c) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid(p.heads, p.B, nsplit);
  attn_fwd_kernel<<<grid, 256, smem, (cudaStream_t)st>>>(p, nco);
  return check_launch("attn_fwd");
}

extern "C" int rcot_attn_bwd(const rcot_attn_params* pp, rcot_stream_t st) {
  RCOT_REQUIRE(pp != nullptr, "attn_bwd: null params");
  const rcot_attn_params& p = *pp;
  int rc = attn_check(p);
  if (rc) return rc;
  RCOT_REQUIRE(p.P && p.dw_out && p.dtemperature && p.W12pack && p.dA, "attn_bwd: null pointer");
  const int c = p.C / p.heads;
  RCOT_REQUIRE(c % 8 == 0, "attn_bwd: channels per head must be a multiple of 8");
  const size_t smem1 = ((size_t)c * (c + 1) + 2 * AB_CH * c) * sizeof(float);
  const size_t smem2 = ((size_t)3 * c * c + 4 * c) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_p1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_p1_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_p1_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) {
      set_error("attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return RCOT_ERR_CUDA;
    }
    attr_set = true;
  }
  // images per CTA (RCOT_ATTN_IPC, default 1): with more, W_out rows are staged once and the dW_out sums leave as one
  // atomic per entry and CTA instead of one per image -- measured neutral at C = 384 (235 us for 1, 2, 4) and slower
  // elsewhere, so the atomics are not what bounds this kernel; one image per CTA keeps the most CTAs in flight.
  static int ipc_env = -1;
  if (ipc_env < 0) {
    const char* e = getenv("RCOT_ATTN_IPC");
    ipc_env = e ? atoi(e) : 0;
  }
  const int chunks = cdiv(p.C, AB_CH);
  const int ipc = ipc_env > 0 ? ipc_env : 1;
  dim3 grid1(p.heads, cdiv(p.B, ipc), chunks);
  // the tile edge is a compile-time constant: at c = 48 (every level but two block kinds) the 6 x 6 register tile of the
  // generic form carried 27 predicated-off FMAs per step and 128 registers
  if (c <= 32)
    attn_bwd_p1_kernel<2><<<grid1, 256, smem1, (cudaStream_t)st>>>(p, ipc);
  else if (c <= 48)
    attn_bwd_p1_kernel<3><<<grid1, 256, smem1, (cudaStream_t)st>>>(p, ipc);
  else
    attn_bwd_p1_kernel<6><<<grid1, 256, smem1, (cudaStream_t)st>>>(p, ipc);
  rc = check_launch("attn_bwd(p1)");
  if (rc) return rc;
  dim3 grid2(p.heads, p.B);
  attn_bwd_p2_kernel<<<grid2, 256, smem2, (cudaStream_t)st>>>(p);
  return check_launch("attn_bwd(p2)");
}
