// Persistent decode kernel with a weight ring (included from decode.cu, inside namespace wm).
//
// The stage chain of one speculative iteration is latency-bound if every stage first waits for
// the grid barrier and only then starts pulling its weights from HBM: a d x d GEMV stage moves
// 3.3 MB -- far less than the bandwidth-delay product of the chip.  But WHICH weight bytes a CTA
// needs is static: for every GEMM stage it owns a fixed, contiguous range of rows of W.  So each
// CTA streams its rows for the upcoming stages through a shared-memory ring with bulk async
// copies (cp.async.bulk -> mbarrier complete_tx; SASS UBLKCP), independently of the activation
// dependency chain: the copies for stages s+1, s+2, ... are in flight while the CTA sits in the
// grid barrier of stage s.  When a stage's activations finally arrive, its weights are already
// on-chip and the stage costs: X staging (L2) + a few MMAs out of shared memory + epilogue.
//
// Ring geometry: a "chunk" = up to 16 weight rows x d columns (fp16), one row per bulk copy, row
// stride d*2 + 64 B (bank-conflict-free LDS.128 of the B fragments).  WM_RING_G chunks are
// resident.  Chunks are consumed in program order; the producer (thread 0) re-fills a slot as
// soon as the chunk that lived there has been consumed.
#pragma once

#define WM_RING_G 3

struct StageInstr { int stage, mode, layer; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

__device__ __forceinline__ bool is_gemm_stage(int st) {
  return st == ST_QKV || st == ST_OPROJ || st == ST_CROSS_Q || st == ST_CROSS_O || st == ST_FC1 || st == ST_FC2 ||
         st == ST_HEADS || st == ST_VOCAB;
}

// static part of a GEMM stage (what the producer needs): weights and shape
struct WDesc { const __half* W; int N, K; };
__device__ __forceinline__ WDesc stage_weights(const DecModel* m, int stage, int mode, int layer) {
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

__device__ __forceinline__ void cta_rows(int N, int cta, int ncta, int& n_begin, int& n_rows) {
  const int rows_per = N / ncta, rem = N % ncta;
  n_begin = cta * rows_per + min(cta, rem);
  n_rows = rows_per + (cta < rem ? 1 : 0);
}

// Producer-side iterator over the chunks of one iteration, in consumption order.
struct ChunkIter {
  const StageInstr* prog;
  int ip, ip_end;         // current instruction / end of the current list
  int list;               // 0 = sweep A, 1 = tail, 2 = verify
  const __half* W;        // current GEMM stage
  int K, n_begin, n_rows, units, segs;
  int sg, u;
  bool valid;
};

struct RingCtx {
  const DecModel* m;
  unsigned char* ring;     // WM_RING_G slots
  uint64_t* full;          // WM_RING_G mbarriers
  int row_stride;          // bytes
  int slot_bytes;
  int cta, ncta;
  unsigned int issued;     // chunks issued (thread 0 only)
  unsigned int consumed;   // chunks consumed (uniform)
  ChunkIter it;            // thread 0 only
};

__device__ void chunk_iter_seek(RingCtx& rc) {
  // position `it` on the first chunk of the next GEMM stage with rows for this CTA
  ChunkIter& it = rc.it;
  const DecModel* m = rc.m;
  while (true) {
    while (it.ip >= it.ip_end) {
      it.list += 1;
      if (it.list > 2) { it.valid = false; return; }
      it.ip = m->prog_off[it.list];
      it.ip_end = m->prog_off[it.list + 1];
    }
    const StageInstr in = it.prog[it.ip];
    if (is_gemm_stage(in.stage)) {
      WDesc w = stage_weights(m, in.stage, in.mode, in.layer);
      int nb, nr;
      cta_rows(w.N, rc.cta, rc.ncta, nb, nr);
      if (nr > 0) {
        it.W = w.W; it.K = w.K; it.n_begin = nb; it.n_rows = nr;
        it.units = (nr + 15) >> 4; it.segs = w.K / m->d; it.sg = 0; it.u = 0; it.valid = true;
        return;
      }
    }
    it.ip += 1;
  }
}
__device__ __forceinline__ void chunk_iter_next(RingCtx& rc) {
  ChunkIter& it = rc.it;
  it.u += 1;
  if (it.u >= it.units) { it.u = 0; it.sg += 1; }
  if (it.sg >= it.segs) { it.ip += 1; chunk_iter_seek(rc); }
}

// thread 0: issue bulk copies for as many future chunks as there are free slots
__device__ void ring_prefetch(RingCtx& rc) {
  const int d = rc.m->d;
  while (rc.it.valid && rc.issued - rc.consumed < WM_RING_G) {
    const ChunkIter& it = rc.it;
    const int slot = rc.issued % WM_RING_G;
    const int r0 = it.u * 16;
    const int nr = min(16, it.n_rows - r0);
    uint64_t* bar = rc.full + slot;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the slot vs async writes
    mbar_expect_tx(bar, (uint32_t)(nr * d * 2));
    const __half* src = it.W + (size_t)(it.n_begin + r0) * it.K + (size_t)it.sg * d;
    unsigned char* dst = rc.ring + (size_t)slot * rc.slot_bytes;
    for (int r = 0; r < nr; ++r) bulk_g2s(dst + (size_t)r * rc.row_stride, src + (size_t)r * it.K, (uint32_t)(d * 2), bar);
    rc.issued += 1;
    chunk_iter_next(rc);
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

__device__ void ring_stage_x(const GemmDesc& g, int seg, int d, __half* xhi, __half* xlo, int xstride, int& rows_dirty) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int T = g.x_rows;
  // rows that still hold data of an earlier, taller stage must read as zero
  if (rows_dirty > T) {
    const int n16 = (rows_dirty - T) * xstride / 8;   // uint4 = 8 halfs; xstride % 8 == 0
    uint4* zh = reinterpret_cast<uint4*>(xhi + (size_t)T * xstride);
    uint4* zl = reinterpret_cast<uint4*>(xlo + (size_t)T * xstride);
    for (int i = tid; i < n16; i += blockDim.x) { zh[i] = make_uint4(0, 0, 0, 0); zl[i] = make_uint4(0, 0, 0, 0); }
  }
  rows_dirty = T;
  if (g.xsrc == XS_LN) {
    const int nv = d >> 7;   // float4 per lane
    for (int r = warp; r < T; r += nwarps) {
      const float4* x4 = reinterpret_cast<const float4*>(g.X + (size_t)(g.x_row0 + r) * d);
      float4 v[WM_LN_MAXV];
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) v[i] = x4[i * 32 + lane];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
          q += (a * a + b * b) + (c * c + e * e);
        }
      const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
      const float4* g4 = reinterpret_cast<const float4*>(g.ln_g);
      const float4* b4 = reinterpret_cast<const float4*>(g.ln_b);
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) {
          const float4 gg = g4[i * 32 + lane], bb = b4[i * 32 + lane];
          float4 y;
          y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
          y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
          y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
          y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
          const int col = (i * 32 + lane) * 4;
          store_hilo4(xhi + (size_t)r * xstride + col, xlo + (size_t)r * xstride + col, y);
        }
    }
  } else {
    const int d4 = d >> 2;
    const int total = T * d4;
    for (int base = 0; base < total; base += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = base + i * blockDim.x + tid;
        if (idx < total) {
          const int r = idx / d4, c4 = idx - r * d4;
          v[i] = *reinterpret_cast<const float4*>(g.X + (size_t)(g.x_row0 + r) * g.K + (size_t)seg * d + c4 * 4);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = base + i * blockDim.x + tid;
        if (idx < total) {
          const int r = idx / d4, c4 = idx - r * d4;
          store_hilo4(xhi + (size_t)r * xstride + c4 * 4, xlo + (size_t)r * xstride + c4 * 4, v[i]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GEMM stage fed from the ring
// ---------------------------------------------------------------------------------------------
__device__ void stage_gemm_ring(RingCtx& rc, const GemmDesc& g, __half* xhi, __half* xlo, float* partial, int& rows_dirty) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int d = rc.m->d;
  const int xstride = d + WM_XPAD;
  int n_begin, n_rows;
  cta_rows(g.N, rc.cta, rc.ncta, n_begin, n_rows);
  if (n_rows == 0) return;   // the producer skips such stages as well
  const int units = (n_rows + 15) >> 4;
  const int segs = g.K / d;
  int nks = 8;               // k-slices per chunk: one warp each, both n8 tiles of the chunk
  while ((d / nks) % 32 != 0) nks >>= 1;
  const int KS = d / nks;
  const int T = g.x_rows;
  const int rs_h = rc.row_stride / 2;   // ring row stride in halfs
  float racc = 0.f;                      // multi-segment accumulator of output element `tid` (units == 1 when segs > 1)

  for (int sg = 0; sg < segs; ++sg) {
    if (sg > 0) __syncthreads();
    ring_stage_x(g, sg, d, xhi, xlo, xstride, rows_dirty);
    __syncthreads();
    for (int u = 0; u < units; ++u) {
      const unsigned int c = rc.consumed;
      const int slot = c % WM_RING_G;
      const int nvalid = min(16, n_rows - u * 16);
      while (!mbar_try_wait(rc.full + slot, (c / WM_RING_G) & 1)) { }
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
      __syncthreads();   // partials visible; every read of the slot and of X (for this chunk) is done
      // the slot is free: let the producer refill it while the epilogue runs
      rc.consumed = c + 1;
      if (tid == 0) ring_prefetch(rc);
      if (tid < 256) {
        const int e = tid >> 5, ln = tid & 31;
        const int j = e >> 2, i = e & 3;
        const int token = (ln >> 2) + ((i >= 2) ? 8 : 0);
        const int rloc = j * 8 + 2 * (ln & 3) + (i & 1);
        if (token < T && rloc < nvalid) {
          float s = 0.f;
          for (int ks = 0; ks < nks; ++ks) s += partial[(size_t)ks * 256 + tid];
          if (segs > 1) { racc += s; s = racc; }
          if (sg == segs - 1) gemm_epilogue(g, token, n_begin + u * 16 + rloc, s, xhi, xlo, xstride);
        }
      }
      __syncthreads();   // partial buffer reusable
    }
  }
}

__host__ __device__ inline size_t ring_smem_bytes(int d) {
  const size_t ring = (size_t)WM_RING_G * 16 * (d * 2 + 64);
  size_t scratch = (size_t)2 * 16 * (d + WM_XPAD) * sizeof(__half);
  if (scratch < cross_attn_smem_bytes()) scratch = cross_attn_smem_bytes();
  if (scratch < self_attn_smem_bytes()) scratch = self_attn_smem_bytes();
  return ring + scratch + (size_t)8 * 256 * sizeof(float) + 64;
}

__global__ void __launch_bounds__(WM_DEC_THREADS, 1)
dec_iteration_ring_kernel(const DecModel* __restrict__ m) {
  extern __shared__ __align__(128) unsigned char smem[];
  const DecState* st = m->st;
  if (st->done) return;
  const int need_a = st->need_a;
  const int d = m->d;
  RingCtx rc;
  rc.m = m;
  rc.row_stride = d * 2 + 64;
  rc.slot_bytes = 16 * rc.row_stride;
  rc.ring = smem;
  size_t scratch = (size_t)2 * 16 * (d + WM_XPAD) * sizeof(__half);
  if (scratch < cross_attn_smem_bytes()) scratch = cross_attn_smem_bytes();
  if (scratch < self_attn_smem_bytes()) scratch = self_attn_smem_bytes();
  unsigned char* scratch_p = smem + (size_t)WM_RING_G * rc.slot_bytes;
  __half* xhi = reinterpret_cast<__half*>(scratch_p);
  __half* xlo = xhi + 16 * (d + WM_XPAD);
  float* partial = reinterpret_cast<float*>(scratch_p + scratch);
  rc.full = reinterpret_cast<uint64_t*>(partial + 8 * 256);
  rc.cta = blockIdx.x; rc.ncta = gridDim.x;
  rc.issued = 0; rc.consumed = 0;
  rc.it.prog = m->prog;
  rc.it.list = need_a ? 0 : 1;
  rc.it.ip = m->prog_off[rc.it.list];
  rc.it.ip_end = m->prog_off[rc.it.list + 1];
  rc.it.valid = false;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WM_RING_G; ++i) mbar_init(rc.full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the activation slice rows must read as zero beyond the rows a stage writes
  {
    uint4* z = reinterpret_cast<uint4*>(scratch_p);
    const int n16 = (int)((size_t)2 * 16 * (d + WM_XPAD) * sizeof(__half) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) { chunk_iter_seek(rc); ring_prefetch(rc); }
  int rows_dirty = 0;
  unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&m->bar[2]);
  unsigned int* bar = m->bar;
  const int cta = blockIdx.x, ncta = gridDim.x;

  for (int list = need_a ? 0 : 1; list <= 2; ++list) {
    const int i0 = m->prog_off[list], i1 = m->prog_off[list + 1];
    for (int ip = i0; ip < i1; ++ip) {
      const StageInstr in = m->prog[ip];
      // optional per-stage timeline (CTA 0 and the last CTA): begin / end of body / end of barrier
      const bool prof = m->prof != nullptr && threadIdx.x == 0 && (cta == 0 || cta == ncta - 1);
      unsigned long long* pr = prof ? m->prof + ((size_t)(cta == 0 ? 0 : 1) * m->prog_off[3] + ip) * 3 : nullptr;
      if (prof) pr[0] = global_timer_ns();
      if (is_gemm_stage(in.stage)) {
        GemmDesc g = make_gemm_desc(m, in.stage, in.mode, in.layer);
        stage_gemm_ring(rc, g, xhi, xlo, partial, rows_dirty);
      } else {
        run_stage(m, in.stage, in.mode, in.layer, cta, ncta, scratch_p);
        // attention stages overlay the activation slice: everything there is dirty now
        if (in.stage == ST_SELF_ATTN || in.stage == ST_CROSS_ATTN || in.stage == ST_SELECT) rows_dirty = 16;
      }
      if (prof) pr[1] = global_timer_ns();
      grid_barrier(bar, epoch, ncta);
      if (prof) pr[2] = global_timer_ns();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) m->bar[2] = epoch;
}
