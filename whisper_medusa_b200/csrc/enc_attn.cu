// Encoder self-attention: full (non-causal) softmax(Q K^T) V over S = 1500 positions, head_dim 64
// (HF modeling_whisper.py:284-357; the head_dim^-0.5 query scaling is applied to the scores here).
// Flash-style: one CTA = 64 query rows of one head, 4 warps x 16 rows, 64-key blocks streamed through a
// double-buffered cp.async pipeline, S and P kept in registers, online softmax in fp32, P rounded
// to fp16 for the P*V MMA (mma.sync m16n8k16).
#include "common.cuh"
#include "engine.h"

namespace wm {

#define EA_BQ 64
#define EA_BK 64
#define EA_THREADS 128

// 64-half (128 B) rows, 8 chunks of 16 B, physical chunk = c ^ (row & 7)
__device__ __forceinline__ int ea_swz(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

__global__ void __launch_bounds__(EA_THREADS) enc_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                              int S, int d) {
  __shared__ __align__(128) __half sQ[EA_BQ * 64];
  __shared__ __align__(128) __half sK[2][EA_BK * 64];
  __shared__ __align__(128) __half sV[2][EA_BK * 64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y, q0 = blockIdx.x * EA_BQ;
  const int ld = 3 * d;
  const __half* Qg = qkv + (size_t)q0 * ld + h * 64;
  const __half* Kg = qkv + d + h * 64;
  const __half* Vg = qkv + 2 * d + h * 64;
  const int nkb = (S + EA_BK - 1) / EA_BK;

  auto load_kv = [&](int buf, int kb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * EA_THREADS;  // 512 chunks
      const int row = c >> 3, ch = c & 7;
      const size_t goff = (size_t)(kb * EA_BK + row) * ld + ch * 8;
      cp_async16(&sK[buf][ea_swz(row, ch)], Kg + goff);
      cp_async16(&sV[buf][ea_swz(row, ch)], Vg + goff);
    }
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + i * EA_THREADS;
    const int row = c >> 3, ch = c & 7;
    cp_async16(&sQ[ea_swz(row, ch)], Qg + (size_t)row * ld + ch * 8);
  }
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[4][4];
  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o_acc[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float L2E = 1.4426950408889634f * 0.125f;  // log2(e) * head_dim^-0.5 (q is NOT pre-scaled)
  const int g = lane >> 2, t = lane & 3;

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) load_kv(buf ^ 1, kb + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = kk * 2 + (lane >> 4);
        ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], &sQ[ea_swz(row, ch)]);
      }
    }
    // S = Q K^T   (16 x 64 per warp)
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s_acc[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int ch = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(b0, b1, b2, b3, &sK[buf][ea_swz(row, ch)]);
        mma_16816(s_acc[np * 2], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_16816(s_acc[np * 2 + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }
    // mask keys >= S (only the last block can be partial)
    if ((kb + 1) * EA_BK > S) {
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb * EA_BK + ni * 8 + 2 * t + (e & 1);
          if (key >= S) s_acc[ni][e] = -INFINITY;
        }
    }
    // online softmax (rows g and g+8)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      mx[0] = fmaxf(mx[0], fmaxf(s_acc[ni][0], s_acc[ni][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s_acc[ni][2], s_acc[ni][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], mscaled[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f((m_run[r] - m_new) * L2E);   // first block: exp2(-inf) = 0
      m_run[r] = m_new;
      mscaled[r] = m_new * L2E;
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      o_acc[ni][0] *= alpha[0]; o_acc[ni][1] *= alpha[0];
      o_acc[ni][2] *= alpha[1]; o_acc[ni][3] *= alpha[1];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const float p0 = exp2f(s_acc[ni][0] * L2E - mscaled[0]);
      const float p1 = exp2f(s_acc[ni][1] * L2E - mscaled[0]);
      const float p2 = exp2f(s_acc[ni][2] * L2E - mscaled[1]);
      const float p3 = exp2f(s_acc[ni][3] * L2E - mscaled[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int kk = ni >> 1;
      if ((ni & 1) == 0) { pf[kk][0] = pack_half2(p0, p1); pf[kk][1] = pack_half2(p2, p3); }
      else               { pf[kk][2] = pack_half2(p0, p1); pf[kk][3] = pack_half2(p2, p3); }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {      // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {    // 16 dims per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = dp * 2 + (lane >> 4);
        ldmatrix_x4_trans(b0, b1, b2, b3, &sV[buf][ea_swz(row, ch)]);
        mma_16816(o_acc[dp * 2], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_16816(o_acc[dp * 2 + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    __syncthreads();   // everyone done with `buf` before it is refilled two iterations later
  }
  cp_async_wait<0>();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row >= S) continue;
    const float inv = 1.0f / l_run[r];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      const int col = h * 64 + ni * 8 + 2 * t;
      *reinterpret_cast<__half2*>(out + (size_t)row * d + col) =
          __floats2half2_rn(o_acc[ni][r * 2] * inv, o_acc[ni][r * 2 + 1] * inv);
    }
  }
}

cudaError_t enc_attention(const __half* qkv, __half* out, int S, int d, int H, cudaStream_t s, int64_t* n_launch) {
  dim3 grid((S + EA_BQ - 1) / EA_BQ, H);
  enc_attn_kernel<<<grid, EA_THREADS, 0, s>>>(qkv, out, S, d);
  if (n_launch) ++*n_launch;
  return cudaGetLastError();
}

}  // namespace wm
