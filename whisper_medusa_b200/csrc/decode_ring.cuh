// Persistent decode kernel with a weight ring (included from decode.cu, inside namespace wm).
//
// The stage chain of one speculative iteration is latency-bound if every stage first waits for
// the grid barrier and only then starts pulling its weights from HBM: a d x d GEMV stage moves
// 3.3 MB -- far less than the bandwidth-delay product of the chip.  But WHICH weight bytes a CTA
// needs is static: for every GEMM stage it owns a fixed, contiguous range of rows of W.  So each
// CTA streams its rows for the upcoming stages through a shared-memory ring with bulk async
// copies (cp.async.bulk -> mbarrier complete_tx; SASS UBLKCP), independently of the activation
// dependency chain: the copies for stages s+1, s+2, ... are in flight while the compute warps sit
// in the grid barrier of stage s.  When a stage's activations finally arrive, its weights are
// already on-chip and the stage costs: X staging (L2) + a few MMAs out of shared memory + epilogue.
//
// Roles: warps 0..14 = compute (WM_DEC_THREADS threads, named barrier 1), warp 15 = producer.
// The producer walks a per-CTA chunk table that the host builds once per model (no pointer
// chasing on the device), waits on `empty[slot]`, arms `full[slot]` with the byte count and lets
// lanes 0..nrows-1 issue one bulk copy per weight row.
//
// Control data never waits on L2 after a barrier (every grid barrier invalidates L1): the model
// description lives in a shared-memory copy, the pass geometry in registers, and the next stage
// instruction is loaded before the barrier it follows.
//
// Stages with K > d (FC2) are split over CTAs along K as well: CTA = (row block, k segment); the
// segment partials go to a global scratch and the CTA that arrives last for a row block folds
// them in segment order (deterministic) and runs the epilogue.
//
// Ring geometry: a "chunk" = up to 16 weight rows x d columns (fp16), row stride d*2 + 64 B
// (bank-conflict-free LDS.128 of the B fragments); WM_RING_G chunks are resident; chunks are
// consumed in program order.
#pragma once

#define WM_RING_G 3
#define WM_RING_THREADS (WM_DEC_THREADS + 32)

struct StageInstr { int stage, mode, layer; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__host__ __device__ __forceinline__ bool is_gemm_stage(int st) {
  return st == ST_QKV || st == ST_OPROJ || st == ST_CROSS_Q || st == ST_CROSS_O || st == ST_FC1 || st == ST_FC2 ||
         st == ST_HEADS || st == ST_VOCAB;
}

// static part of a GEMM stage (what the producer needs): weights and shape
struct WDesc { const __half* W; int N, K; };
__host__ __device__ inline WDesc stage_weights(const DecModel* m, int stage, int mode, int layer) {
  const DecLayer& L = m->layers[layer];
  const int d = m->d;
  WDesc w;
  switch (stage) {
    case ST_QKV: w.W = L.qkv_w; w.N = 3 * d; w.K = d; break;
    case ST_OPROJ: w.W = L.o_w; w.N = d; w.K = d; break;
    case ST_CROSS_Q: w.W = L.cq_w; w.N = d; w.K = d; break;
    case ST_CROSS_O: w.W = L.co_w; w.N = d; w.K = d; break;
    case ST_FC1: w.W = L.fc1_w; w.N = m->ffn; w.K = d; break;
    case ST_FC2: w.W = L.fc2_w; w.N = d; w.K = m->ffn; break;
    case ST_HEADS:
      w.W = m->heads_w; w.K = d;
      w.N = (mode == MODE_A) ? (m->has_block ? m->K : m->K + 1) * d : d;
      break;
    default: w.W = m->embed; w.N = m->V; w.K = d; break;
  }
  return w;
}

__host__ __device__ __forceinline__ void cta_rows(int N, int part, int nparts, int& n_begin, int& n_rows) {
  const int rows_per = N / nparts, rem = N % nparts;
  n_begin = part * rows_per + (part < rem ? part : rem);
  n_rows = rows_per + (part < rem ? 1 : 0);
}

// Work of one CTA in a GEMM stage: rows [n_begin, n_begin + n_rows) of W, k segment `seg` of `segs`
// (segs > 1 <=> K > d: the stage is split along K over CTAs; `block` = row block shared by `segs` CTAs).
struct GemmWork { int n_begin, n_rows, seg, segs, block; };
__host__ __device__ __forceinline__ GemmWork gemm_work(int N, int K, int d, int cta, int ncta) {
  GemmWork w;
  w.segs = K / d;
  if (w.segs <= 1) {
    w.segs = 1; w.seg = 0; w.block = cta;
    cta_rows(N, cta, ncta, w.n_begin, w.n_rows);
    return w;
  }
  const int nb = ncta / w.segs;   // host guarantees nb >= 1
  if (cta >= nb * w.segs) { w.n_begin = 0; w.n_rows = 0; w.seg = 0; w.block = 0; return w; }
  w.block = cta / w.segs;
  w.seg = cta - w.block * w.segs;
  cta_rows(N, w.block, nb, w.n_begin, w.n_rows);
  return w;
}

struct RingCtx {
  unsigned char* ring;     // WM_RING_G slots
  uint64_t* full;          // WM_RING_G mbarriers (producer -> compute)
  uint64_t* empty;         // WM_RING_G mbarriers (compute -> producer)
  int row_stride;          // bytes
  int slot_bytes;
  int d;
  int cta, ncta;
  unsigned int consumed;   // chunks consumed so far in this launch (uniform across the compute warps)
};

// ---------------------------------------------------------------------------------------------
// producer warp: stream the chunk table through the ring
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void ring_producer(const RingCtx& rc, const ChunkDesc* __restrict__ tab, int first, int last) {
  const int lane = threadIdx.x & 31;
  if (first >= last) return;
  ChunkDesc nxt = tab[first];
  for (int c = first; c < last; ++c) {
    const unsigned int k = (unsigned int)(c - first);
    const int slot = k % WM_RING_G;
    const unsigned int round = k / WM_RING_G;
    const ChunkDesc dsc = nxt;
    if (c + 1 < last) nxt = tab[c + 1];                 // next descriptor is in flight while we wait
    if (lane == 0) {
      while (!mbar_try_wait(rc.empty + slot, (round & 1) ^ 1)) { }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the slot vs async writes
      mbar_expect_tx(rc.full + slot, (uint32_t)(dsc.nrows * rc.d * 2));
    }
    __syncwarp();
    if (lane < dsc.nrows) {
      bulk_g2s(rc.ring + (size_t)slot * rc.slot_bytes + (size_t)lane * rc.row_stride,
               reinterpret_cast<const unsigned char*>(dsc.src) + (size_t)lane * dsc.row_bytes, (uint32_t)(rc.d * 2),
               rc.full + slot);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// activation staging (fp32 -> fp16 hi/lo in shared memory), all loads in flight before first use
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_hilo4(__half* hi, __half* lo, float4 y) {
  const __half h0 = __float2half_rn(y.x), h1 = __float2half_rn(y.y), h2 = __float2half_rn(y.z), h3 = __float2half_rn(y.w);
  __half2 a = __halves2half2(h0, h1), b = __halves2half2(h2, h3);
  __half2 c = __floats2half2_rn(y.x - __half2float(h0), y.y - __half2float(h1));
  __half2 e = __floats2half2_rn(y.z - __half2float(h2), y.w - __half2float(h3));
  uint2 vh, vl;
  vh.x = *reinterpret_cast<uint32_t*>(&a); vh.y = *reinterpret_cast<uint32_t*>(&b);
  vl.x = *reinterpret_cast<uint32_t*>(&c); vl.y = *reinterpret_cast<uint32_t*>(&e);
  *reinterpret_cast<uint2*>(hi) = vh;
  *reinterpret_cast<uint2*>(lo) = vl;
}

#define WM_LN_MAXV 10   // float4 per lane: d <= 1280 (every Whisper size)

__device__ __forceinline__ void ring_stage_x(const GemmDesc& g, int seg, int d, __half* xhi, __half* xlo, int xstride, int& rows_dirty) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = WM_DEC_THREADS >> 5;
  const int T = g.x_rows;
  const int nv = d >> 7;   // float4 per lane
  const bool ln = (g.xsrc == XS_LN);
  const float4* g4 = reinterpret_cast<const float4*>(g.ln_g);
  const float4* b4 = reinterpret_cast<const float4*>(g.ln_b);
  if (ln && warp < T) {
    // the affine parameters are only needed after two reductions: pull them into L1 meanwhile
    for (int i = 0; i < nv; ++i) { prefetch_l1(g4 + i * 32 + lane); prefetch_l1(b4 + i * 32 + lane); }
  }
  // one warp per row: lane l holds float4 columns l, l+32, ...  (all loads in flight before first use)
  for (int r = warp; r < T; r += nwarps) {
    const float4* x4 = reinterpret_cast<const float4*>(g.X + (size_t)(g.x_row0 + r) * g.K + (size_t)seg * d);
    float4 v[WM_LN_MAXV];
#pragma unroll
    for (int i = 0; i < WM_LN_MAXV; ++i)
      if (i < nv) v[i] = x4[i * 32 + lane];
    float mean = 0.f, rstd = 1.f;
    if (ln) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      mean = warp_sum(s) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
          q += (a * a + b * b) + (c * c + e * e);
        }
      rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
    }
    __half* hi = xhi + (size_t)r * xstride + lane * 4;
    __half* lo = xlo + (size_t)r * xstride + lane * 4;
#pragma unroll
    for (int i = 0; i < WM_LN_MAXV; ++i)
      if (i < nv) {
        float4 y = v[i];
        if (ln) {
          const float4 gg = g4[i * 32 + lane], bb = b4[i * 32 + lane];
          y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
          y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
          y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
          y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
        }
        store_hilo4(hi + i * 128, lo + i * 128, y);
      }
  }
  // rows that still hold data of an earlier, taller stage must read as zero
  if (rows_dirty > T) {
    const int n16 = (rows_dirty - T) * xstride / 8;   // uint4 = 8 halfs; xstride % 8 == 0
    uint4* zh = reinterpret_cast<uint4*>(xhi + (size_t)T * xstride);
    uint4* zl = reinterpret_cast<uint4*>(xlo + (size_t)T * xstride);
    for (int i = tid; i < n16; i += WM_DEC_THREADS) { zh[i] = make_uint4(0, 0, 0, 0); zl[i] = make_uint4(0, 0, 0, 0); }
  }
  rows_dirty = T;
}

// ---------------------------------------------------------------------------------------------
// GEMM stage fed from the ring (compute warps)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_gemm_ring(RingCtx& rc, const DecModel* m, const GemmDesc& g, __half* xhi, __half* xlo, float* partial,
                                int& rows_dirty, unsigned long long* pr) {
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int d = rc.d;
  const int xstride = d + WM_XPAD;
  const GemmWork wk = gemm_work(g.N, g.K, d, rc.cta, rc.ncta);
  if (wk.n_rows == 0) return;   // the chunk table has no entry for such stages either
  const int n_begin = wk.n_begin, n_rows = wk.n_rows;
  const int units = (n_rows + 15) >> 4;
  int nks = 8;               // k-slices per chunk: one warp each, both n8 tiles of the chunk
  while ((d / nks) % 32 != 0) nks >>= 1;
  const int KS = d / nks;
  const int T = g.x_rows;
  const int rs_h = rc.row_stride / 2;   // ring row stride in halfs
  const bool ksplit = wk.segs > 1;

  ring_stage_x(g, wk.seg, d, xhi, xlo, xstride, rows_dirty);
  // output element owned by this thread in the fold below
  const int e_ = tid >> 5, ln_ = tid & 31;
  const int j_ = (e_ >> 2) & 1, i_ = e_ & 3;
  const int token = (ln_ >> 2) + ((i_ >= 2) ? 8 : 0);
  const int rloc = j_ * 8 + 2 * (ln_ & 3) + (i_ & 1);
  // residual epilogue: fetch the old value while the MMAs run (single-unit stages only)
  float old = 0.f;
  const bool mine = tid < 256 && token < T;
  const bool pre_old = (g.epi == EPI_RESID) && !ksplit && units == 1 && mine && rloc < n_rows;
  if (pre_old) old = g.out[(size_t)token * g.ldo + n_begin + rloc];
  cta_sync();
  if (pr) pr[3] = global_timer_ns();
  for (int u = 0; u < units; ++u) {
    const unsigned int c = rc.consumed;
    const int slot = c % WM_RING_G;
    const int nvalid = min(16, n_rows - u * 16);
    // this thread's bias for the unit: in flight while the MMAs run
    float bias_v = 0.f;
    if (mine && rloc < nvalid && g.bias && !ksplit) bias_v = g.bias[n_begin + u * 16 + rloc];
    while (!mbar_try_wait(rc.full + slot, (c / WM_RING_G) & 1)) { }
    if (pr && u == 0) pr[4] = global_timer_ns();
    if (warp < nks) {
      const __half* sl = reinterpret_cast<const __half*>(rc.ring + (size_t)slot * rc.slot_bytes);
      const bool v0 = gq < nvalid, v1 = (gq + 8) < nvalid;
      const __half* w0p = sl + (size_t)gq * rs_h + warp * KS + 8 * tq;
      const __half* w1p = sl + (size_t)(gq + 8) * rs_h + warp * KS + 8 * tq;
      const __half* xh0 = xhi + (size_t)gq * xstride + warp * KS + 8 * tq;
      const __half* xh1 = xh0 + 8 * xstride;
      const __half* xl0 = xlo + (size_t)gq * xstride + warp * KS + 8 * tq;
      const __half* xl1 = xl0 + 8 * xstride;
      const bool t1 = (gq + 8) < T;   // token rows 8..15 are zero when T <= 8 + gq
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 5
      for (int kk = 0; kk < KS; kk += 32) {
        const uint4 wa = v0 ? *reinterpret_cast<const uint4*>(w0p + kk) : z;
        const uint4 wb = v1 ? *reinterpret_cast<const uint4*>(w1p + kk) : z;
        const uint4 ah0 = *reinterpret_cast<const uint4*>(xh0 + kk);
        const uint4 al0 = *reinterpret_cast<const uint4*>(xl0 + kk);
        const uint4 ah1 = t1 ? *reinterpret_cast<const uint4*>(xh1 + kk) : z;
        const uint4 al1 = t1 ? *reinterpret_cast<const uint4*>(xl1 + kk) : z;
        mma_16816(c0, ah0.x, ah1.x, ah0.y, ah1.y, wa.x, wa.y);
        mma_16816(c0, ah0.z, ah1.z, ah0.w, ah1.w, wa.z, wa.w);
        mma_16816(c0, al0.x, al1.x, al0.y, al1.y, wa.x, wa.y);
        mma_16816(c0, al0.z, al1.z, al0.w, al1.w, wa.z, wa.w);
        mma_16816(c1, ah0.x, ah1.x, ah0.y, ah1.y, wb.x, wb.y);
        mma_16816(c1, ah0.z, ah1.z, ah0.w, ah1.w, wb.z, wb.w);
        mma_16816(c1, al0.x, al1.x, al0.y, al1.y, wb.x, wb.y);
        mma_16816(c1, al0.z, al1.z, al0.w, al1.w, wb.z, wb.w);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        partial[(size_t)warp * 256 + e * 32 + lane] = c0[e];
        partial[(size_t)warp * 256 + (4 + e) * 32 + lane] = c1[e];
      }
    }
    cta_sync();   // partials visible; every read of the slot (and of X for this chunk) is done
    if (pr && u == 0) pr[5] = global_timer_ns();
    rc.consumed = c + 1;
    if (tid == 0) mbar_arrive(rc.empty + slot);   // hand the slot back to the producer
    if (mine && rloc < nvalid) {
      float s = 0.f;
      for (int ks = 0; ks < nks; ++ks) s += partial[(size_t)ks * 256 + tid];
      const int row = n_begin + u * 16 + rloc;
      // common epilogues inline (bias was fetched while the MMAs ran); the Medusa-head ones are rare
      if (ksplit) {
        m->gemm_part[((size_t)wk.seg * 16 + token) * g.N + row] = s;
      } else if (g.epi == EPI_RESID) {
        float* o = g.out + (size_t)token * g.ldo + row;
        *o = (pre_old ? old : *o) + (s + bias_v);
      } else if (g.epi == EPI_STORE || g.epi == EPI_LOGITS) {
        g.out[(size_t)token * g.ldo + row] = s + bias_v;
      } else if (g.epi == EPI_GELU) {
        g.out[(size_t)token * g.ldo + row] = gelu_erf(s + bias_v);
      } else if (g.epi == EPI_QKV) {
        const float v = s + bias_v;
        if (row < d) g.out[(size_t)token * g.ldo + row] = v;
        else if (row < 2 * d) g.kc[(size_t)(g.base + token) * d + (row - d)] = __float2half_rn(v);
        else g.vc[(size_t)(g.base + token) * d + (row - 2 * d)] = __float2half_rn(v);
      } else {
        gemm_epilogue_heads(g.epi, g.out, g.ldo, g.out_row0, d, bias_v, token, row, s, xhi, xlo, xstride);
      }
    }
    cta_sync();   // partial buffer reusable
    if (pr && u == 0) pr[6] = global_timer_ns();
  }
  if (ksplit) {
    // the last of the `segs` CTAs of this row block folds the segment partials, always in segment order
    __threadfence();
    cta_sync();
    if (tid == 0) {
      const unsigned int prev = atomicAdd(&m->gemm_cnt[wk.block], 1u);
      s_last = (prev == (unsigned int)(wk.segs - 1)) ? 1 : 0;
      if (s_last) m->gemm_cnt[wk.block] = 0u;
    }
    cta_sync();
    if (s_last) {
      __threadfence();
      for (int idx = tid; idx < T * n_rows; idx += WM_DEC_THREADS) {
        const int t = idx / n_rows, row = n_begin + (idx - t * n_rows);
        float s = 0.f;
        for (int sg = 0; sg < wk.segs; ++sg) s += __ldcg(m->gemm_part + ((size_t)sg * 16 + t) * g.N + row);
        // K-split stages are residual GEMMs (FC2)
        g.out[(size_t)t * g.ldo + row] += s + (g.bias ? g.bias[row] : 0.f);
      }
    }
  }
}

__host__ __device__ inline size_t ring_scratch_bytes(int d) {
  size_t scratch = (size_t)2 * 16 * (d + WM_XPAD) * sizeof(__half);
  if (scratch < cross_attn_smem_bytes()) scratch = cross_attn_smem_bytes();
  if (scratch < self_attn_smem_bytes()) scratch = self_attn_smem_bytes();
  return (scratch + 127) / 128 * 128;
}
__host__ __device__ inline size_t ring_model_bytes() { return (sizeof(DecModel) + 127) / 128 * 128; }
__host__ __device__ inline size_t ring_smem_bytes(int d) {
  return (size_t)WM_RING_G * 16 * (d * 2 + 64) + ring_scratch_bytes(d) + (size_t)8 * 256 * sizeof(float) + ring_model_bytes() + 128;
}

__global__ void __launch_bounds__(WM_RING_THREADS, 1)
dec_iteration_ring_kernel(const DecModel* __restrict__ gm) {
  extern __shared__ __align__(128) unsigned char smem[];
  const DecState* st = gm->st;
  if (st->done) return;
  const int need_a = st->need_a;
  const int L0 = st->L, kv0 = st->kv_len;
  const int d = gm->d;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int warp = threadIdx.x >> 5;
  RingCtx rc;
  rc.d = d;
  rc.row_stride = d * 2 + 64;
  rc.slot_bytes = 16 * rc.row_stride;
  rc.ring = smem;
  unsigned char* scratch_p = smem + (size_t)WM_RING_G * rc.slot_bytes;
  __half* xhi = reinterpret_cast<__half*>(scratch_p);
  __half* xlo = xhi + 16 * (d + WM_XPAD);
  float* partial = reinterpret_cast<float*>(scratch_p + ring_scratch_bytes(d));
  DecModel* sm = reinterpret_cast<DecModel*>(partial + 8 * 256);   // shared-memory copy of the model description
  rc.full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sm) + ring_model_bytes());
  rc.empty = rc.full + WM_RING_G;
  rc.cta = cta; rc.ncta = ncta;
  rc.consumed = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WM_RING_G; ++i) { mbar_init(rc.full + i, 1); mbar_init(rc.empty + i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int4* src = reinterpret_cast<const int4*>(gm);
    int4* dst = reinterpret_cast<int4*>(sm);
    // (the device copy of DecModel is allocated in 256-byte granules: reading up to the next 16 B is safe)
    for (int i = threadIdx.x; i < (int)((sizeof(DecModel) + 15) / 16); i += WM_RING_THREADS) dst[i] = src[i];
  }
  __syncthreads();   // the only full-CTA barrier: after it the producer warp goes its own way
  const DecModel* m = sm;

  if (warp == WM_DEC_THREADS / 32) {
    // ===== producer warp =====
    const int* off = m->chunk_off + cta * 4;
    ring_producer(rc, m->chunk_tab, need_a ? off[0] : off[1], off[3]);
    return;
  }

  // ===== compute warps =====
  {
    // the activation slice rows must read as zero beyond the rows a stage writes
    uint4* z = reinterpret_cast<uint4*>(scratch_p);
    const int n16 = (int)((size_t)2 * 16 * (d + WM_XPAD) * sizeof(__half) / 16);
    for (int i = threadIdx.x; i < n16; i += WM_DEC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  cta_sync();
  int rows_dirty = 0;
  unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&m->bar[2]);
  unsigned int* bar = m->bar;
  // pass geometry of this launch (the loop state only changes in the very last stage)
  PassGeom geom[3];
  geom[MODE_A].T = L0 - kv0;  geom[MODE_A].base = kv0;
  geom[MODE_B].T = m->K + 1;  geom[MODE_B].base = L0;
  geom[MODE_TAIL].T = 1;      geom[MODE_TAIL].base = L0 - 1;

  const int ip_first = need_a ? m->prog_off[0] : m->prog_off[1];
  const int ip_last = m->prog_off[3];
  StageInstr in = m->prog[ip_first];
  for (int ip = ip_first; ip < ip_last; ++ip) {
    // the next instruction is fetched before the barrier of this one: its L2 latency hides there
    StageInstr nxt = in;
    if (ip + 1 < ip_last) nxt = m->prog[ip + 1];
    // optional per-stage timeline (CTA 0 and the last CTA): begin / end of body / end of barrier
    const bool prof = m->prof != nullptr && threadIdx.x == 0 && (cta == 0 || cta == ncta - 1);
    unsigned long long* pr = prof ? m->prof + ((size_t)(cta == 0 ? 0 : 1) * m->prog_off[3] + ip) * 8 : nullptr;
    if (prof) pr[0] = global_timer_ns();
    PassGeom pgv;
    if (in.mode == MODE_A) pgv = geom[MODE_A];
    else if (in.mode == MODE_B) pgv = geom[MODE_B];
    else pgv = geom[MODE_TAIL];
    const PassGeom* pg = &pgv;
    if (is_gemm_stage(in.stage)) {
      GemmDesc g = make_gemm_desc(m, in.stage, in.mode, in.layer, pg);
      stage_gemm_ring(rc, m, g, xhi, xlo, partial, rows_dirty, pr);
    } else {
      run_stage<false>(m, in.stage, in.mode, in.layer, cta, ncta, scratch_p, pg);
      // attention / scan stages overlay the activation slice: everything there is dirty now
      if (in.stage == ST_SELF_ATTN || in.stage == ST_CROSS_ATTN || in.stage == ST_SELECT1 || in.stage == ST_SELECT2)
        rows_dirty = 16;
    }
    if (prof) pr[1] = global_timer_ns();
    grid_barrier(bar, epoch, ncta);
    if (prof) pr[2] = global_timer_ns();
    in = nxt;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) m->bar[2] = epoch;
}
