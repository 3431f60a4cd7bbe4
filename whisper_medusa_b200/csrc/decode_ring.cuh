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
  uint64_t* xbar;          // mbarrier of the activation-row bulk copies
  unsigned int xphase;     // its parity
};

// ---------------------------------------------------------------------------------------------
// producer warp: stream the chunk table through the ring
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void ring_producer(unsigned char* ring, uint64_t* full, uint64_t* empty, int row_stride, int slot_bytes,
                                           int d, const ChunkDesc* __restrict__ tab, int first, int last) {
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
      while (!mbar_try_wait(empty + slot, (round & 1) ^ 1)) { }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the slot vs async writes
      mbar_expect_tx(full + slot, (uint32_t)(dsc.nrows * d * 2));
    }
    __syncwarp();
    if (lane < dsc.nrows) {
      bulk_g2s(ring + (size_t)slot * slot_bytes + (size_t)lane * row_stride,
               reinterpret_cast<const unsigned char*>(dsc.src) + (size_t)lane * dsc.row_bytes, (uint32_t)(d * 2),
               full + slot);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// activation staging
//
// The T activation rows of a GEMM stage (fp32, written by other CTAs before the grid barrier)
// are pulled into shared memory by bulk async copies -- one per row, issued by the lanes of
// warp 0 the moment the barrier opens -- instead of 10 dependent ld.global per lane.  They land
// as fp32 rows of stride d*4 + 16 B and are split IN PLACE into the fp16 hi/lo operand format:
// every 8-byte pair of floats (x[k], x[k+1]) becomes { half2 hi(k,k+1), half2 lo(k,k+1) }, so one
// LDS.128 of the MMA loop fetches the hi AND lo A-fragment registers of two k-pairs.
// LayerNorm stages: warp-per-row statistics out of shared memory, then thread-per-column
// normalisation with that column's gamma/beta held in registers (loaded before the wait on X).
// Rows >= T keep stale bits: MMA rows are independent and rows >= T are never stored.
// ---------------------------------------------------------------------------------------------
#define WM_XS_PADB 16   // bytes of row padding: stride = 16 (mod 128) => conflict-free LDS.128 / STS.128

__device__ __forceinline__ uint4 split_hilo4(float4 y) {
  // packed conversions: hi = rn(x), lo = rn(x - hi)
  const __half2 a = __floats2half2_rn(y.x, y.y), b = __floats2half2_rn(y.z, y.w);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  const __half2 c = __floats2half2_rn(y.x - fa.x, y.y - fa.y);
  const __half2 e = __floats2half2_rn(y.z - fb.x, y.w - fb.y);
  uint4 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&c);
  o.z = *reinterpret_cast<const uint32_t*>(&b); o.w = *reinterpret_cast<const uint32_t*>(&e);
  return o;
}
// value of activation (row r, column n) back from the split buffer (hi + lo is exact in fp32)
__device__ __forceinline__ float xbuf_value(const unsigned char* xb, int xs, int r, int n) {
  const __half* p = reinterpret_cast<const __half*>(xb + (size_t)r * xs + (size_t)(n >> 1) * 8) + (n & 1);
  return __half2float(p[0]) + __half2float(p[2]);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#define WM_LN_MAXV 10   // float4 per lane: d <= 1280 (every Whisper size)

// ---------------------------------------------------------------------------------------------
// GEMM stage fed from the ring (compute warps).  `sd` = this CTA's resolved record (shared memory).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_gemm_ring(RingCtx& rc, const DecModel* m, const CtaStage* sd, int Tpass, int base,
                                                unsigned char* xb, float* partial, unsigned long long* pr) {
  __shared__ int s_last;
  __shared__ float2 s_stat[WM_MAX_T];
  const int n_rows = sd->n_rows;
  if (n_rows == 0) return;   // the chunk table has no entry for such stages either
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = WM_DEC_THREADS >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int d = rc.d;
  const int XS = d * 4 + WM_XS_PADB;
  const int T = sd->x_rows_fixed ? sd->x_rows_fixed : Tpass;
  const bool ln = sd->ln_g != nullptr;
  const int nv4 = d >> 2;

  if (pr) pr[7] = global_timer_ns();
  // ---- X rows: global (L2) -> shared, one bulk copy per row ----
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the buffer vs async writes
      mbar_expect_tx(rc.xbar, (uint32_t)(T * d * 4));
    }
    __syncwarp();
    if (lane < T) bulk_g2s(xb + (size_t)lane * XS, sd->X + (size_t)lane * sd->x_ld, (uint32_t)(d * 4), rc.xbar);
  }
  // this thread's LayerNorm column (in flight while X arrives)
  float4 gg = make_float4(1.f, 1.f, 1.f, 1.f), bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ln && tid < nv4) {
    gg = reinterpret_cast<const float4*>(sd->ln_g)[tid];
    bb = reinterpret_cast<const float4*>(sd->ln_b)[tid];
  }
  const int units = (n_rows + 15) >> 4;
  const bool ksplit = sd->segs > 1;
  const int epi = sd->epi;
  const int n_begin = sd->n_begin;
  // output element owned by this thread in the fold below
  const int e_ = tid >> 5, ln_ = tid & 31;
  const int j_ = (e_ >> 2) & 1, i_ = e_ & 3;
  const int token = (ln_ >> 2) + ((i_ >= 2) ? 8 : 0);
  const int rloc = j_ * 8 + 2 * (ln_ & 3) + (i_ & 1);
  const bool mine = tid < 256 && token < T;
  // residual epilogue: fetch the old value while X arrives / the MMAs run (single-unit stages only)
  float old = 0.f;
  const bool pre_old = (epi == EPI_RESID) && !ksplit && units == 1 && mine && rloc < n_rows;
  if (pre_old) old = ldcg_f(&sd->out[(size_t)token * sd->ldo + n_begin + rloc]);
  float bias_v = 0.f;        // bias of this thread's row in unit 0 (later units reload)
  {
    const float* bias = sd->bias;
    if (mine && rloc < n_rows && bias && !ksplit) bias_v = bias[n_begin + rloc];
  }
  // pull the next LayerNorm stage's gamma / beta into L2 (they are cold: 2 GB of weights pass through L2 per iteration)
  if (sd->pf[0] && tid >= 256 && tid < 256 + 2 * ((d * 4 + 127) >> 7)) {
    const int i = tid - 256, half = (d * 4 + 127) >> 7;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(i < half ? sd->pf[0] : sd->pf[1]);
    prefetch_l2(p + (size_t)(i < half ? i : i - half) * 128);
  }

  while (!mbar_try_wait(rc.xbar, rc.xphase)) { }
  rc.xphase ^= 1u;
  if (pr) pr[8] = global_timer_ns();
  if (ln) {
    // statistics: one warp per row, lane l sums float4 columns l, l+32, ... (two passes over shared memory)
    const int nv = d >> 7;
    for (int r = warp; r < T; r += nwarps) {
      const float4* x4 = reinterpret_cast<const float4*>(xb + (size_t)r * XS) + lane;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) { const float4 v = x4[i * 32]; s += (v.x + v.y) + (v.z + v.w); }
      const float mean = warp_sum(s) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < WM_LN_MAXV; ++i)
        if (i < nv) {
          const float4 v = x4[i * 32];
          const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
          q += (a * a + b * b) + (c * c + e * e);
        }
      const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
      if (lane == 0) s_stat[r] = make_float2(mean, rstd);
    }
    cta_sync();
    if (pr) pr[9] = global_timer_ns();
    if (tid < nv4) {
      for (int r = 0; r < T; ++r) {
        uint4* p = reinterpret_cast<uint4*>(xb + (size_t)r * XS) + tid;
        const float4 v = *reinterpret_cast<const float4*>(p);
        const float2 st = s_stat[r];
        float4 y;
        y.x = (v.x - st.x) * st.y * gg.x + bb.x;
        y.y = (v.y - st.x) * st.y * gg.y + bb.y;
        y.z = (v.z - st.x) * st.y * gg.z + bb.z;
        y.w = (v.w - st.x) * st.y * gg.w + bb.w;
        *p = split_hilo4(y);
      }
    }
  } else {
    // flat over the buffer (the 16-byte row pad is converted along: no index arithmetic)
    uint4* p = reinterpret_cast<uint4*>(xb);
    const int n16 = T * (XS >> 4);
    for (int idx = tid; idx < n16; idx += WM_DEC_THREADS) p[idx] = split_hilo4(*reinterpret_cast<const float4*>(p + idx));
  }
  if (pr) pr[10] = global_timer_ns();
  cta_sync();
  if (pr) pr[3] = global_timer_ns();
  int nks = 8;               // k-slices per chunk: one warp each, both n8 tiles of the chunk
  while ((d / nks) % 32 != 0) nks >>= 1;
  const int KS = d / nks;
  const int rs_h = rc.row_stride / 2;   // ring row stride in halfs
  for (int u = 0; u < units; ++u) {
    const unsigned int c = rc.consumed;
    const int slot = c % WM_RING_G;
    const int nvalid = min(16, n_rows - u * 16);
    // this thread's bias for the unit: in flight while the MMAs run
    if (u > 0) {
      const float* bias = sd->bias;
      bias_v = 0.f;
      if (mine && rloc < nvalid && bias && !ksplit) bias_v = bias[n_begin + u * 16 + rloc];
    }
    while (!mbar_try_wait(rc.full + slot, (c / WM_RING_G) & 1)) { }
    if (pr && u == 0) pr[4] = global_timer_ns();
    if (warp < nks) {
      const __half* sl = reinterpret_cast<const __half*>(rc.ring + (size_t)slot * rc.slot_bytes);
      const bool v0 = gq < nvalid, v1 = (gq + 8) < nvalid;
      const __half* w0p = sl + (size_t)gq * rs_h + warp * KS + 8 * tq;
      const __half* w1p = sl + (size_t)(gq + 8) * rs_h + warp * KS + 8 * tq;
      const unsigned char* x0 = xb + (size_t)gq * XS + (size_t)(warp * KS + 8 * tq) * 4;
      const unsigned char* x1 = x0 + (size_t)8 * XS;
      const bool t1 = (gq + 8) < T;   // token rows 8..15 contribute nothing when T <= 8 + gq
      // four independent accumulator chains: (n8 tile 0 / 1) x (hi / lo part of X)
      float c0h[4] = {0.f, 0.f, 0.f, 0.f}, c0l[4] = {0.f, 0.f, 0.f, 0.f};
      float c1h[4] = {0.f, 0.f, 0.f, 0.f}, c1l[4] = {0.f, 0.f, 0.f, 0.f};
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 5
      for (int kk = 0; kk < KS; kk += 32) {
        const uint4 wa = v0 ? *reinterpret_cast<const uint4*>(w0p + kk) : z;
        const uint4 wb = v1 ? *reinterpret_cast<const uint4*>(w1p + kk) : z;
        // {hi(k,k+1), lo(k,k+1), hi(k+2,k+3), lo(k+2,k+3)} for k = kk + 8 tq and k + 4
        const uint4 p0 = *reinterpret_cast<const uint4*>(x0 + kk * 4);
        const uint4 p1 = *reinterpret_cast<const uint4*>(x0 + kk * 4 + 16);
        const uint4 q0 = t1 ? *reinterpret_cast<const uint4*>(x1 + kk * 4) : z;
        const uint4 q1 = t1 ? *reinterpret_cast<const uint4*>(x1 + kk * 4 + 16) : z;
        mma_16816(c0h, p0.x, q0.x, p0.z, q0.z, wa.x, wa.y);
        mma_16816(c0l, p0.y, q0.y, p0.w, q0.w, wa.x, wa.y);
        mma_16816(c1h, p0.x, q0.x, p0.z, q0.z, wb.x, wb.y);
        mma_16816(c1l, p0.y, q0.y, p0.w, q0.w, wb.x, wb.y);
        mma_16816(c0h, p1.x, q1.x, p1.z, q1.z, wa.z, wa.w);
        mma_16816(c0l, p1.y, q1.y, p1.w, q1.w, wa.z, wa.w);
        mma_16816(c1h, p1.x, q1.x, p1.z, q1.z, wb.z, wb.w);
        mma_16816(c1l, p1.y, q1.y, p1.w, q1.w, wb.z, wb.w);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        partial[(size_t)warp * 256 + e * 32 + lane] = c0h[e] + c0l[e];
        partial[(size_t)warp * 256 + (4 + e) * 32 + lane] = c1h[e] + c1l[e];
      }
    }
    cta_sync();   // partials visible; every read of the slot is done
    if (pr && u == 0) pr[5] = global_timer_ns();
    rc.consumed = c + 1;
    if (tid == 0) mbar_arrive(rc.empty + slot);   // hand the slot back to the producer
    if (mine && rloc < nvalid) {
      float s = 0.f;
      for (int ks = 0; ks < nks; ++ks) s += partial[(size_t)ks * 256 + tid];
      const int row = n_begin + u * 16 + rloc;
      float* out = sd->out;
      const int ldo = sd->ldo;
      // common epilogues inline (bias was fetched while the MMAs ran); the Medusa-head ones are rare
      if (ksplit) {
        m->gemm_part[((size_t)sd->seg * 16 + token) * sd->N + row] = s;
      } else if (epi == EPI_RESID) {
        float* o = out + (size_t)token * ldo + row;
        *o = (pre_old ? old : ldcg_f(o)) + (s + bias_v);
      } else if (epi == EPI_STORE || epi == EPI_LOGITS) {
        out[(size_t)token * ldo + row] = s + bias_v;
      } else if (epi == EPI_GELU) {
        out[(size_t)token * ldo + row] = gelu_erf(s + bias_v);
      } else if (epi == EPI_QKV) {
        const float v = s + bias_v;
        if (row < d) out[(size_t)token * ldo + row] = v;
        else if (row < 2 * d) sd->kc[(size_t)(base + token) * d + (row - d)] = __float2half_rn(v);
        else sd->vc[(size_t)(base + token) * d + (row - 2 * d)] = __float2half_rn(v);
      } else if (epi == EPI_HEADS_A) {
        // head `row / d` on the newest token's hidden state: x + SiLU(W x + b)  (medusa ResBlock)
        const int head = row / d, n = row - head * d;
        out[(size_t)(sd->out_row0 + head) * ldo + n] = xbuf_value(xb, XS, 0, n) + silu(s + bias_v);
      } else {   // EPI_HEAD_B
        out[(size_t)token * ldo + row] = xbuf_value(xb, XS, token, row) + silu(s + bias_v);
      }
    }
    cta_sync();   // partial buffer reusable
    if (pr && u == 0) pr[6] = global_timer_ns();
  }
  if (ksplit) {
    // the last of the `segs` CTAs of this row block folds the segment partials, always in segment order
    const int segs = sd->segs, block = sd->block, N = sd->N;
    cta_sync();
    if (tid == 0) {
      const unsigned int prev = atom_add_release(&m->gemm_cnt[block], 1u);
      s_last = (prev == (unsigned int)(segs - 1)) ? 1 : 0;
      if (s_last) m->gemm_cnt[block] = 0u;
    }
    cta_sync();
    if (s_last) {
      const float* bias = sd->bias;
      float* out = sd->out;
      const int ldo = sd->ldo;
      for (int idx = tid; idx < T * n_rows; idx += WM_DEC_THREADS) {
        const int t = idx / n_rows, row = n_begin + (idx - t * n_rows);
        float s = 0.f;
        for (int sg = 0; sg < segs; ++sg) s += __ldcg(m->gemm_part + ((size_t)sg * 16 + t) * N + row);
        // K-split stages are residual GEMMs (FC2)
        float* o = out + (size_t)t * ldo + row;
        *o = ldcg_f(o) + (s + (bias ? bias[row] : 0.f));
      }
    }
  }
}

__host__ __device__ inline size_t ring_scratch_bytes(int d) {
  size_t scratch = (size_t)16 * (d * 4 + WM_XS_PADB);
  if (scratch < cross_attn_smem_bytes()) scratch = cross_attn_smem_bytes();
  if (scratch < self_attn_smem_bytes()) scratch = self_attn_smem_bytes();
  return (scratch + 127) / 128 * 128;
}
__host__ __device__ inline size_t ring_model_bytes() { return (sizeof(DecModel) + 127) / 128 * 128; }
__host__ __device__ inline size_t ring_smem_bytes(int d) {
  return (size_t)WM_RING_G * 16 * (d * 2 + 64) + ring_scratch_bytes(d) + (size_t)8 * 256 * sizeof(float) + ring_model_bytes() + 128;
}

template <bool PROF>
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
  float* partial = reinterpret_cast<float*>(scratch_p + ring_scratch_bytes(d));
  DecModel* sm = reinterpret_cast<DecModel*>(partial + 8 * 256);   // shared-memory copy of the model description
  rc.full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sm) + ring_model_bytes());
  rc.empty = rc.full + WM_RING_G;
  rc.xbar = rc.empty + WM_RING_G;
  rc.xphase = 0u;
  rc.cta = cta; rc.ncta = ncta;
  rc.consumed = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WM_RING_G; ++i) { mbar_init(rc.full + i, 1); mbar_init(rc.empty + i, 1); }
    mbar_init(rc.xbar, 1);
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
    ring_producer(rc.ring, rc.full, rc.empty, rc.row_stride, rc.slot_bytes, d, m->chunk_tab, need_a ? off[0] : off[1], off[3]);
    return;
  }

  // ===== compute warps =====
  __shared__ CtaStage s_desc[2];   // resolved record of the running stage / the next one
  unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(&m->bar[2]);
  unsigned int* bar = m->bar;
  const int Kh = m->K;   // the pass geometry derives from (L0, kv0, K): the loop state only changes in the very last stage
  const int lane = threadIdx.x & 31;

  const int ip_first = need_a ? m->prog_off[0] : m->prog_off[1];
  const int ip_last = m->prog_off[3];
  const CtaStage* tab = m->stage_tab + cta;   // record of instruction ip: tab[ip * ncta]
  if (warp == 0)
    reinterpret_cast<uint32_t*>(&s_desc[ip_first & 1])[lane] = reinterpret_cast<const uint32_t*>(tab + (size_t)ip_first * ncta)[lane];
  cta_sync();
  for (int ip = ip_first; ip < ip_last; ++ip) {
    const CtaStage* sd = &s_desc[ip & 1];
    // the next record is fetched while this stage runs (and the one after it pulled into L2)
    uint32_t nxt_w = 0u;
    const bool fetch = (warp == (WM_DEC_THREADS / 32 - 1)) && (ip + 1 < ip_last);
    if (fetch) {
      nxt_w = reinterpret_cast<const uint32_t*>(tab + (size_t)(ip + 1) * ncta)[lane];
      if (lane == 0 && ip + 2 < ip_last) prefetch_l2(tab + (size_t)(ip + 2) * ncta);
    }
    // optional per-stage timeline (CTA 0 and the last CTA): begin / end of body / end of barrier
    const bool prof = PROF && m->prof != nullptr && threadIdx.x == 0 && (cta == 0 || cta == ncta - 1);
    unsigned long long* pr = prof ? m->prof + ((size_t)(cta == 0 ? 0 : 1) * m->prog_off[3] + ip) * 16 : nullptr;
    if (prof) pr[0] = global_timer_ns();
    const int stage = sd->stage, mode = sd->mode;
    PassGeom pgv;   // (registers: every callee that takes it is inlined)
    if (mode == MODE_A) { pgv.T = L0 - kv0; pgv.base = kv0; }
    else if (mode == MODE_B) { pgv.T = Kh + 1; pgv.base = L0; }
    else { pgv.T = 1; pgv.base = L0 - 1; }
    if (is_gemm_stage(stage)) {
      stage_gemm_ring(rc, m, sd, pgv.T, pgv.base, scratch_p, partial, pr);
    } else {
      run_stage<false>(m, stage, mode, sd->layer, cta, ncta, scratch_p, &pgv);
    }
    if (fetch) reinterpret_cast<uint32_t*>(&s_desc[(ip + 1) & 1])[lane] = nxt_w;
    if (prof) pr[1] = global_timer_ns();
    epoch = grid_barrier_step<false>(bar, epoch, ncta);
    if (prof) pr[2] = global_timer_ns();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) m->bar[2] = epoch;
}
