// Encoder-side dense GEMM on the 5th-generation tensor cores (sm_100a):
//   C[M,N] = A[M,K] * W[N,K]^T   fp16 operands (both K-major), fp32 accumulation in TMEM.
//
// Structure: one (128 MB) x BN output tile per CTA, K swept in 64-column blocks.  At M = 1500 these GEMMs are bound by
// the rate at which ONE SM can pull operands out of L2 (about 80 GB/s per SM, measured), not by the tensor pipe: a tile
// costs (BM + BN) * K * 2 bytes, so the host picks, per GEMM, the tile shape with the fewest operand bytes on the
// busiest SM (see pick_tile and DESIGN.md):
//   MB = 1, BN = 128 (or 64), 3-stage ring, two CTAs per SM (one tile's epilogue overlaps the other's main loop);
//   MB = 1, BN = 128, 6-stage ring, one CTA per SM: outputs too narrow to give every SM two tiles;
//   MB = 2, BN = 128 / 192 / 256, one CTA per SM: two 128-row accumulators share every W tile (up to all 512 TMEM
//   columns) -- wide outputs in ONE wave of <= 148 tiles.
//   warp 0  : TMA producer   -- cp.async.bulk.tensor.2d (SWIZZLE_128B) of the A and W tiles into a
//                               shared-memory ring, completion on `full` mbarriers
//   warp 1  : MMA issuer     -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                               (M=128, N=BN, K=16) x4 (x MB) per stage, accumulators = MB * BN TMEM columns;
//                               tcgen05.commit releases the stage (`empty`) / signals the epilogue
//   warps 2.. : epilogue     -- 2 (MB = 1) or 4 (MB = 2) warps per TMEM lane quarter, 32-column blocks dealt round-robin: tcgen05.ld
//                               (32 lanes x 32 columns per instruction) -> registers ->
//                               bias / GELU / residual / position epilogue -> global
// The A operand may be an "implicit im2col" view: conv1/conv2 read a time-major activation whose
// GEMM rows overlap (row stride 80 resp. 2d elements) -- the TMA tensor map simply carries that
// stride, no im2col buffer exists.
#include <cuda.h>

#include <cstring>
#include <map>
#include <tuple>

#include "common.cuh"
#include "engine.h"
#include "tc_common.cuh"

namespace wm {

#define TC_BM 128
#define TC_BN 128
#define TC_BK 64
#ifndef WM_TC_NARROW_BN
#define WM_TC_NARROW_BN 128   /* tile width for outputs of <= 1536 columns (64: twice the tiles, 3 stages, 2 CTAs / SM) */
#endif
#ifndef WM_TC_W192_STAGES
#define WM_TC_W192_STAGES 4   /* ring depth of the 256 x 192 tiles (225 KB; 3 stages: 5.54 vs 5.51 ms per clip) */
#endif
#ifndef WM_TC_SMEM_BIAS
#define WM_TC_SMEM_BIAS 1     /* 0: bias values read from global memory inside the epilogue (A/B) */
#endif
#define TC_THREADS 320       /* MB = 1: TMA warp + MMA warp + 8 epilogue warps */
#define TC_THREADS_WIDE 576  /* MB = 2: ... + 16 epilogue warps (the epilogue of the single wave is exposed) */
__host__ __device__ constexpr int tc_stage_bytes(int mb, int bn) {
  return mb == 1 ? (TC_BM + TC_BN) * TC_BK * 2 /* any width: A at 0, W at 16 KB */ : (mb * TC_BM + bn) * TC_BK * 2;
}
__host__ __device__ constexpr int tc_tmem_cols(int mb, int bn) {
  return mb * bn <= 128 ? 128 : (mb * bn <= 256 ? 256 : 512);
}

struct TcArgs {
  int M, N, K;
  int epi;
  const float* bias;
  __half* out16; int ldo16;
  float* out32; int ldo32;
  const float* pos;
  __half* vt; int vt_col0, vt_ld;   // columns >= vt_col0 are ALSO written transposed: vt[col - vt_col0][row] (null: off)
  __half* ck; __half* cv; int kv_spad;   // non-null: the output [pos][k | v] goes to the decode layout [head][kv_spad][72]
};

template <int EPI, int MB, int BN, int STAGES>
__global__ void __launch_bounds__(MB == 1 ? TC_THREADS : TC_THREADS_WIDE, MB == 1 ? 2 : 1)
enc_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // stage s: A tile [128 MB rows][64 halfs] at s * STAGE_BYTES, W tile [BN rows][64 halfs] behind it (both 1024-B
  // aligned, SW128)
  constexpr int STAGE_BYTES = tc_stage_bytes(MB, BN);
  constexpr int A_BYTES = MB * TC_BM * TC_BK * 2;
  constexpr int TMEM_COLS = tc_tmem_cols(MB, BN);
  constexpr int NEPI_Q = (MB == 1 ? TC_THREADS / 32 - 2 : TC_THREADS_WIDE / 32 - 2) / 4;   // epilogue warps per lane quarter
  __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES], s_tmem_full;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_bias[BN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * (MB * TC_BM), n0 = blockIdx.x * BN;
  const int KT = a.K / TC_BK;
  const uint32_t smem_base = tc_smem_u32(smem_raw);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      tc_mbar_init(tc_smem_u32(&s_full[s]), 1);
      tc_mbar_init(tc_smem_u32(&s_empty[s]), 1);
    }
    tc_mbar_init(tc_smem_u32(&s_tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  tc_grid_dep_launch();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // the weight tiles of the first ring round do not depend on the predecessor kernel: they are requested before the
      // grid-dependency wait (programmatic dependent launch, tc_common.cuh), the activation tiles after it
      const int pre = KT < STAGES ? KT : STAGES;
      for (int kt = 0; kt < pre; ++kt) {
        const uint32_t full = tc_smem_u32(&s_full[kt]);
        tc_mbar_expect_tx(full, (MB * TC_BM + BN) * TC_BK * 2);
        tc_tma_load_2d(smem_base + kt * STAGE_BYTES + A_BYTES, &map_w, kt * TC_BK, n0, full);
      }
      tc_grid_dep_wait();
      for (int kt = 0; kt < pre; ++kt)
        tc_tma_load_2d(smem_base + kt * STAGE_BYTES, &map_a, kt * TC_BK, m0, tc_smem_u32(&s_full[kt]));   // (one box of 128 MB rows)
      for (int kt = pre; kt < KT; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        tc_mbar_wait(tc_smem_u32(&s_empty[s]), ph ^ 1);
        const uint32_t full = tc_smem_u32(&s_full[s]);
        tc_mbar_expect_tx(full, (MB * TC_BM + BN) * TC_BK * 2);
        tc_tma_load_2d(smem_base + s * STAGE_BYTES, &map_a, kt * TC_BK, m0, full);
        tc_tma_load_2d(smem_base + s * STAGE_BYTES + A_BYTES, &map_w, kt * TC_BK, n0, full);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tc_instr_desc(BN);
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        tc_mbar_wait(tc_smem_u32(&s_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = tc_smem_desc(smem_base + s * STAGE_BYTES);
        const uint64_t bdesc = tc_smem_desc(smem_base + s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 halfs = 32 B inside the 128-B swizzle atom: +2 in the (>>4) start address; the second 128-row
          // block of A lies 16 KB further, its accumulator BN columns further
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
            tc_mma(tmem_base + (uint32_t)(mb * BN), adesc + (uint64_t)(mb * ((TC_BM * TC_BK * 2) >> 4) + k * 2),
                   bdesc + (uint64_t)(k * 2), idesc, (kt > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(tc_smem_u32(&s_empty[s]));      // stage reusable once these MMAs have read it
      }
      tc_commit(tc_smem_u32(&s_tmem_full));       // accumulator complete
    }
  } else {
    // ===== epilogue: a warp reads the TMEM lane quarter (warp % 4) =====
    // The bias row of the tile goes to shared memory while the main loop runs (these warps idle until the accumulator
    // is complete).  (A/B, rejected: also pulling the residual / position values of the fp32 epilogues into registers
    // before the accumulator is complete -- 64 more live registers, 5.75 vs 5.54 ms per clip.)
    const int q = warp & 3;
    constexpr int NEPI_T = NEPI_Q * 4 * 32;
    constexpr int NBLK = MB * (BN / 32) / NEPI_Q;   // 32 x 32 blocks per warp
    constexpr bool F32_OUT = (EPI == ENC_EPI_BIAS_RES_F32 || EPI == ENC_EPI_BIAS_GELU_POS_F32);
    for (int i = (int)threadIdx.x - 64; i < BN; i += NEPI_T) s_bias[i] = a.bias[n0 + i];
    tc_grid_dep_wait();   // (the epilogue reads and writes activations)
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_T) : "memory");   // (epilogue warps only) s_bias complete
    tc_mbar_wait(tc_smem_u32(&s_tmem_full), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto block = [&](const int blk) {
      const int mb = blk / (BN / 32), cb = blk % (BN / 32);
      const int row = m0 + mb * TC_BM + q * 32 + lane;
      const bool row_ok = row < a.M;
      uint32_t v[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * BN + cb * 32), v);
      const int col0 = n0 + cb * 32;
      if constexpr (!F32_OUT) {
        if (!row_ok) return;
        __half* dst = a.out16 + (size_t)row * a.ldo16 + col0;
        if (EPI == ENC_EPI_BIAS_F16 && a.ck != nullptr) {
          // cross-attention K/V straight into the layout the decode kernels read: [head][position][64 dims + 8 pad]
          // (pad halfs and positions >= M stay zero from the allocation); 64 contiguous bytes per thread as above
          const int dm = a.N >> 1;
          const int cc = col0 >= dm ? col0 - dm : col0;
          dst = (col0 >= dm ? a.cv : a.ck) + ((size_t)(cc >> 6) * a.kv_spad + row) * 72 + (cc & 63);
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const float* bsrc = WM_TC_SMEM_BIAS ? &s_bias[cb * 32 + c8 * 8] : a.bias + col0 + c8 * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(bsrc);
          const float4 b1 = *reinterpret_cast<const float4*>(bsrc + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0 = __uint_as_float(v[c8 * 8 + 2 * e]) + bb[2 * e];
            float x1 = __uint_as_float(v[c8 * 8 + 2 * e + 1]) + bb[2 * e + 1];
            if (EPI == ENC_EPI_BIAS_GELU_F16) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); }
            pk[e] = pack_half2(x0, x1);
          }
          *reinterpret_cast<uint4*>(dst + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (EPI == ENC_EPI_BIAS_F16 && a.vt != nullptr && col0 >= a.vt_col0) {
            // V^T for the tcgen05 attention: thread = position, so for a fixed column the 32 lanes write 32
            // consecutive halfs (64 contiguous bytes)
            __half* vcol = a.vt + (size_t)(col0 - a.vt_col0 + c8 * 8) * a.vt_ld + row;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 hh = *reinterpret_cast<const __half2*>(&pk[e]);
              vcol[(size_t)(2 * e) * a.vt_ld] = __low2half(hh);
              vcol[(size_t)(2 * e + 1) * a.vt_ld] = __high2half(hh);
            }
          }
        }
      } else {
        // fp32 read-modify-write epilogues (residual stream / conv2 + positions).  tcgen05.ld hands every thread one ROW
        // of the tile; row-wise global accesses would touch 32 different lines per instruction (the kernel then spends
        // most of its time waiting for the residual loads -- ncu, round 2).  The 32 x 32 block is transposed through
        // shared memory (the pipeline stages are idle once the accumulator is complete): lane = column, so every
        // load / store of the residual stream is one contiguous 128-byte segment.
        float* sC = reinterpret_cast<float*>(smem_raw) + (size_t)(warp - 2) * (32 * 33);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) sC[lane * 33 + c] = __uint_as_float(v[c]);
        __syncwarp();
        const float bcol = WM_TC_SMEM_BIAS ? s_bias[cb * 32 + lane] : a.bias[col0 + lane];
        const int row_base = m0 + mb * TC_BM + q * 32;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float in[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = row_base + r0 + i;
            in[i] = 0.f;
            if (rr < a.M)
              in[i] = (EPI == ENC_EPI_BIAS_RES_F32) ? a.out32[(size_t)rr * a.ldo32 + col0 + lane]
                                                    : a.pos[(size_t)rr * a.N + col0 + lane];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = row_base + r0 + i;
            if (rr < a.M) {
              const float x = sC[(r0 + i) * 33 + lane] + bcol;
              a.out32[(size_t)rr * a.ldo32 + col0 + lane] = (EPI == ENC_EPI_BIAS_RES_F32) ? in[i] + x : gelu_erf(x) + in[i];
            }
          }
        }
      }
    };
#pragma unroll 1
    for (int b = 0; b < NBLK; ++b) block(((warp - 2) >> 2) + b * NEPI_Q);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps (cached per operand view) and launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 view [rows][K] with row stride `ld` elements, boxes of [128 rows][64 columns], SWIZZLE_128B
static bool make_map(CUtensorMap* map, const __half* base, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {ld * sizeof(__half)};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static constexpr size_t tc_smem(int mb, int bn, int stages) { return (size_t)stages * tc_stage_bytes(mb, bn) + 1024; }

// every instantiation: X(EPI, MB, BN, STAGES)
#define WM_TC_NARROW_TILES(X, EPI) X(EPI, 1, 128, 3) X(EPI, 1, 128, 6) X(EPI, 1, 64, 3)
#define WM_TC_WIDE_TILES(X, EPI) X(EPI, 2, 128, 4) X(EPI, 2, 192, WM_TC_W192_STAGES) X(EPI, 2, 256, 3)
#define WM_TC_ALL(X)                                                                                             \
  WM_TC_NARROW_TILES(X, ENC_EPI_BIAS_F16) WM_TC_NARROW_TILES(X, ENC_EPI_BIAS_GELU_F16)                           \
  WM_TC_NARROW_TILES(X, ENC_EPI_BIAS_RES_F32) WM_TC_NARROW_TILES(X, ENC_EPI_BIAS_GELU_POS_F32)                   \
  WM_TC_WIDE_TILES(X, ENC_EPI_BIAS_F16) WM_TC_WIDE_TILES(X, ENC_EPI_BIAS_GELU_F16)

cudaError_t enc_gemm_tc_configure() {
  cudaError_t e;
#define WM_SET(EPI, MB, BN, ST)                                                                                          \
  e = cudaFuncSetAttribute(enc_gemm_tc_kernel<EPI, MB, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                           (int)tc_smem(MB, BN, ST));                                                                    \
  if (e != cudaSuccess) return e;
  WM_TC_ALL(WM_SET)
#undef WM_SET
  return get_encode_fn() ? cudaSuccess : cudaErrorNotSupported;
}

// Tile shape of one GEMM.  The operand bytes the busiest SM pulls out of L2 decide the time: tiles are dealt to the
// 148 SMs in ceil(tiles / 148) rounds of K * (BM + BN) * 2 bytes each.  Candidates with two 128-row blocks per CTA exist
// for the fp16-output epilogues only (the fp32 read-modify-write epilogues feed N = d outputs: too few tiles already).
struct TcTile { int mb, bn, stages; };
static TcTile pick_tile(const EncGemmArgs& g, int n_sm) {
  const bool narrow = g.N <= 1536;   // fewer than one 128 x 128 tile per SM and round
  TcTile best = {1, (narrow && WM_TC_NARROW_BN == 64 && g.N % 64 == 0) ? 64 : TC_BN, 3};
  if (narrow && best.bn == TC_BN) best.stages = 6;
  const bool f16_out = g.epi == ENC_EPI_BIAS_F16 || g.epi == ENC_EPI_BIAS_GELU_F16;
  if (g.tile == 1 || !f16_out) return best;
  auto cost = [&](int mb, int bn) {
    const long tiles = (long)((g.M + mb * TC_BM - 1) / (mb * TC_BM)) * (g.N / bn);
    return ((tiles + n_sm - 1) / n_sm) * (long)(mb * TC_BM + bn);
  };
  long c = cost(1, best.bn);
  const TcTile wide[3] = {{2, 128, 4}, {2, 192, WM_TC_W192_STAGES}, {2, 256, 3}};
  for (const TcTile& t : wide) {
    if (g.N % t.bn != 0) continue;
    const long ct = cost(t.mb, t.bn);
    if (ct < c) { c = ct; best = t; }
  }
  return best;
}

void enc_gemm_tc_tile(int M, int N, int K, bool fp16_out, int n_sm, int out[3]) {
  EncGemmArgs g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = N; g.K = K; g.epi = fp16_out ? ENC_EPI_BIAS_F16 : ENC_EPI_BIAS_RES_F32;
  const TcTile t = pick_tile(g, n_sm > 0 ? n_sm : 148);
  out[0] = t.mb * TC_BM; out[1] = t.bn; out[2] = t.stages;
}

// `a_rows` = rows of the A view that may be touched (the allocation is padded accordingly)
cudaError_t enc_gemm_tc(const EncGemmArgs& g, int a_rows, cudaStream_t s, int64_t* n_launch) {
  if (g.N % TC_BN != 0 || g.K % TC_BK != 0 || (g.lda % 8) != 0) return cudaErrorInvalidValue;
  typedef std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t> Key;
  static thread_local std::map<Key, CUtensorMap> cache;
  auto get = [&](const __half* base, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows, CUtensorMap* out) -> bool {
    Key k(base, rows, K, ld, box_rows);
    auto it = cache.find(k);
    if (it == cache.end()) {
      CUtensorMap m;
      if (!make_map(&m, base, rows, K, ld, box_rows)) return false;
      it = cache.emplace(k, m).first;
    }
    *out = it->second;
    return true;
  };
  static thread_local int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
      n_sm = 148;
  }
  const TcTile t = pick_tile(g, n_sm);
  CUtensorMap ma, mw;
  if (!get(g.A, (uint64_t)a_rows, (uint64_t)g.K, (uint64_t)g.lda, (uint32_t)(t.mb * TC_BM), &ma)) return cudaErrorInvalidValue;
  if (!get(g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.K, (uint32_t)t.bn, &mw)) return cudaErrorInvalidValue;
  TcArgs a;
  a.M = g.M; a.N = g.N; a.K = g.K; a.epi = g.epi; a.bias = g.bias; a.out16 = g.out16; a.ldo16 = g.ldo16;
  a.out32 = g.out32; a.ldo32 = g.ldo32; a.pos = g.pos;
  a.vt = g.vt; a.vt_col0 = g.vt_col0; a.vt_ld = g.vt_ld;
  a.ck = g.ck; a.cv = g.cv; a.kv_spad = g.kv_spad;
  if (a.ck != nullptr && (g.epi != ENC_EPI_BIAS_F16 || (g.N >> 1) % 64 != 0)) return cudaErrorInvalidValue;
  dim3 grid(g.N / t.bn, (g.M + t.mb * TC_BM - 1) / (t.mb * TC_BM));
  bool launched = false;
  cudaError_t le = cudaSuccess;
#define WM_LAUNCH(EPI, MB, BN, ST)                                                                                       \
  if (!launched && g.epi == EPI && t.mb == MB && t.bn == BN && t.stages == ST) {                                          \
    le = tc_launch(enc_gemm_tc_kernel<EPI, MB, BN, ST>, grid, dim3(MB == 1 ? TC_THREADS : TC_THREADS_WIDE),               \
                   tc_smem(MB, BN, ST), s, g.pdl != 0, ma, mw, a);                                                        \
    launched = true;                                                                                                     \
  }
  WM_TC_ALL(WM_LAUNCH)
#undef WM_LAUNCH
  if (!launched) return cudaErrorInvalidValue;
  if (le != cudaSuccess) return le;
  if (n_launch) ++*n_launch;
  return cudaGetLastError();
}

}  // namespace wm
