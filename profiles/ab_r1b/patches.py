import sys, os
def rd(d, f): return open(os.path.join(d, 'csrc', f)).read()
def wr(d, f, s): open(os.path.join(d, 'csrc', f), 'w').write(s)
def rep(s, old, new, cnt=1):
    assert s.count(old) >= 1, "pattern not found: " + old[:80]
    return s.replace(old, new) if cnt == 0 else s.replace(old, new, cnt)

def presplit(d):
    c = rd(d, 'common.cuh')
    c = rep(c, "  const float* pf_bias;     // this CTA's bias slice of the next GEMM stage (L2 prefetch)\n  int pad_[4];",
      "  const float* pf_bias;     // this CTA's bias slice of the next GEMM stage (L2 prefetch)\n  int presplit;             // 1: X was written by its producer in the fp16 hi/lo operand format (no split pass)\n  int out_split;            // 1: the epilogue writes `out` in that format (the consumer is a presplit stage)\n  int pad_[2];")
    c = rep(c, "__device__ __forceinline__ void red_add_release(", '''// Activation element (row base `row`, column n) in the fp16 hi/lo operand format of the ring kernel's GEMM stages:
// every float pair (k, k+1) occupies its 8 bytes as { half2 hi(k,k+1), half2 lo(k,k+1) } (decode_ring.cuh).
__device__ __forceinline__ void store_split(float* row, int n, float v) {
  __half* p = reinterpret_cast<__half*>(row) + (size_t)(n >> 1) * 4 + (n & 1);
  const __half h = __float2half_rn(v);
  p[0] = h;
  p[2] = __float2half_rn(v - __half2float(h));
}
__device__ __forceinline__ void red_add_release(''')
    wr(d, 'common.cuh', c)
    t = rd(d, 'decode.cu')
    t = rep(t, "__device__ __forceinline__ void stage_self_attn(const DecModel* m, int mode,", "template <bool SPLIT_OUT = false>\n__device__ __forceinline__ void stage_self_attn(const DecModel* m, int mode,")
    t = rep(t, "        m->attn[(size_t)(t0 + rr) * d + h * 64 + c] = o / s_st[rr];",
      "        if (SPLIT_OUT) store_split(m->attn + (size_t)(t0 + rr) * d, h * 64 + c, o / s_st[rr]);   // (ring kernel: operand format of the O-projection)\n        else m->attn[(size_t)(t0 + rr) * d + h * 64 + c] = o / s_st[rr];")
    t = rep(t, "__device__ __forceinline__ void cross_attn_fold(const DecModel* m, int T, int h, int nch) {", "template <bool SPLIT_OUT>\n__device__ __forceinline__ void cross_attn_fold(const DecModel* m, int T, int h, int nch) {")
    t = rep(t, '''    float* o = m->attn + (size_t)rr * d + h * 64;
    o[lane] = num0 / den;
    o[32 + lane] = num1 / den;''', '''    if (SPLIT_OUT) {   // ring kernel: operand format of the cross-O projection (common.cuh: store_split)
      store_split(m->attn + (size_t)rr * d, h * 64 + lane, num0 / den);
      store_split(m->attn + (size_t)rr * d, h * 64 + 32 + lane, num1 / den);
    } else {
      float* o = m->attn + (size_t)rr * d + h * 64;
      o[lane] = num0 / den;
      o[32 + lane] = num1 / den;
    }''')
    t = rep(t, "template <class AfterQK, class AfterPV>\n__device__ __forceinline__ void cross_attn_core(", "template <bool SPLIT_OUT, class AfterQK, class AfterPV>\n__device__ __forceinline__ void cross_attn_core(")
    t = rep(t, "    cross_attn_fold(m, T, h, nch);", "    cross_attn_fold<SPLIT_OUT>(m, T, h, nch);")
    t = rep(t, "    cross_attn_core(m, g.T, h, c, nch, nk, nk_pad, sK, sV, cs, [] {}, [] {});", "    cross_attn_core<false>(m, g.T, h, c, nch, nk, nk_pad, sK, sV, cs, [] {}, [] {});")
    t = rep(t, '''      c.segs = wk.segs; c.seg = wk.seg; c.block = wk.block;
    }''', '''      c.segs = wk.segs; c.seg = wk.seg; c.block = wk.block;
      // activations that only ever feed one GEMM stage travel in the MMA operand format: the attention stages and
      // the GELU epilogue of FC1 write it, O-proj / cross-O / FC2 skip their split pass
      c.presplit = (stage == ST_OPROJ || stage == ST_CROSS_O || stage == ST_FC2) ? 1 : 0;
      c.out_split = (stage == ST_FC1) ? 1 : 0;
    }''')
    wr(d, 'decode.cu', t)
    s = rd(d, 'decode_ring.cuh')
    s = rep(s, "cross_attn_core(\n        m, T, h, c, nch, nk, nk_pad, sK, sV, cs,", "cross_attn_core<true>(\n        m, T, h, c, nch, nk, nk_pad, sK, sV, cs,")
    s = rep(s, '''  } else {
    // flat over the buffer (the 16-byte row pad is converted along: no index arithmetic)''', '''  } else if (!sd->presplit) {
    // flat over the buffer (the 16-byte row pad is converted along: no index arithmetic)''')
    s = rep(s, '''        } else if (epi == EPI_GELU) {
          out[(size_t)token * ldo + row] = gelu_erf(s + bias_v[k]);''', '''        } else if (epi == EPI_GELU) {
          const float v = gelu_erf(s + bias_v[k]);
          if (sd->out_split) store_split(out + (size_t)token * ldo, row, v);
          else out[(size_t)token * ldo + row] = v;''')
    s = rep(s, '''    } else if (stage == ST_CROSS_ATTN) {
      stage_cross_attn_ring<D>(rs, smem, m, pgv.T, cta, ncta, pr);''', '''    } else if (stage == ST_CROSS_ATTN) {
      stage_cross_attn_ring<D>(rs, smem, m, pgv.T, cta, ncta, pr);
    } else if (stage == ST_SELF_ATTN) {
      stage_self_attn<true>(m, mode, sd->layer, cta, ncta, smem + G::SCRATCH_OFF, &pgv, pr);''')
    wr(d, 'decode_ring.cuh', s)


def ln_direct(d):
    s = rd(d, 'decode_ring.cuh')
    s = rep(s, "  if (warp == 0) {\n    if (lane == 0) {\n      asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");   // earlier generic accesses of the buffer vs async writes\n      mbar_expect_tx(xbar, (uint32_t)(T * D * 4));",
               "  if (warp == 0 && !ln) {\n    if (lane == 0) {\n      asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");   // earlier generic accesses of the buffer vs async writes\n      mbar_expect_tx(xbar, (uint32_t)(T * D * 4));")
    s = rep(s, "  } else if (warp == 1) {\n    // this CTA's bias slice of the NEXT GEMM stage", "  } else if (warp == nwarps - 1) {\n    // this CTA's bias slice of the NEXT GEMM stage")
    a = s.index("  while (!mbar_try_wait(xbar, rs.xpar)) { }\n  rs.xpar ^= 1u;\n  if (pr) pr[8] = global_timer_ns();\n  if (ln) {")
    b = s.index("  } else if (!sd->presplit) {\n    // flat over the buffer") if "  } else if (!sd->presplit) {\n    // flat over the buffer" in s else s.index("  } else {\n    // flat over the buffer")
    presplit = "  } else if (!sd->presplit) {\n    // flat over the buffer" in s
    new = '''  if (ln) {
    // LayerNorm stages: each warp pulls its row straight from L2 into registers (all loads in flight at once; the raw
    // row never visits shared memory): lane l holds float4 columns l, l+32, ...; statistics (two passes over the
    // registers), normalisation and the hi/lo split without a CTA barrier in between
    float4 v[G::NV];
#pragma unroll
    for (int i = 0; i < G::NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp < T) {
      const float4* src = reinterpret_cast<const float4*>(sd->X + (size_t)warp * sd->x_ld) + lane;
#pragma unroll
      for (int i = 0; i < G::NV; ++i)
        if (i * 32 + lane < G::NV4) v[i] = __ldcg(src + i * 32);
    }
    // gamma / beta were bulk-copied into the (idle) partial buffer during the preceding barrier
    while (!mbar_try_wait(pbar, rs.ppar)) { }
    rs.ppar ^= 1u;
    if (pr) pr[8] = global_timer_ns();
    for (int r = warp; r < T; r += nwarps) {
      if (r != warp) {   // T > 11 rows: second round
        const float4* src = reinterpret_cast<const float4*>(sd->X + (size_t)r * sd->x_ld) + lane;
#pragma unroll
        for (int i = 0; i < G::NV; ++i)
          if (i * 32 + lane < G::NV4) v[i] = __ldcg(src + i * 32);
      }
      uint4* const row = reinterpret_cast<uint4*>(xb + (size_t)r * G::XS) + lane;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < G::NV; ++i)
        if (i * 32 + lane < G::NV4) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) / (float)D;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < G::NV; ++i)
        if (i * 32 + lane < G::NV4) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
          q += (a * a + b * b) + (c * c + e * e);
        }
      const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
#pragma unroll
      for (int i = 0; i < G::NV; ++i)
        if (i * 32 + lane < G::NV4) {
          const float4 gg = reinterpret_cast<const float4*>(partial)[i * 32 + lane];
          const float4 bb = reinterpret_cast<const float4*>(partial)[G::NV4 + i * 32 + lane];
          float4 y;
          y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
          y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
          y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
          y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
          row[i * 32] = split_hilo4(y);
        }
    }
  } else {
    while (!mbar_try_wait(xbar, rs.xpar)) { }
    rs.xpar ^= 1u;
    if (pr) pr[8] = global_timer_ns();
  }
'''
    if presplit:
        tail = "  if (!ln && !sd->presplit) {\n    // flat over the buffer"
        s = s[:a] + new + tail + s[b + len("  } else if (!sd->presplit) {\n    // flat over the buffer"):]
    else:
        tail = "  if (!ln) {\n    // flat over the buffer"
        s = s[:a] + new + tail + s[b + len("  } else {\n    // flat over the buffer"):]
    s = rep(s, "  __shared__ int s_last;\n  __shared__ float2 s_stat[WM_MAX_T];", "  __shared__ int s_last;")
    wr(d, 'decode_ring.cuh', s)

def ln_noinline(d):
    """after ln_direct: move the row processing into an out-of-line function"""
    s = rd(d, 'decode_ring.cuh')
    a = s.index("    float4 v[G::NV];\n#pragma unroll\n    for (int i = 0; i < G::NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);\n    if (warp < T) {")
    b = s.index("  } else {\n    while (!mbar_try_wait(xbar, rs.xpar)) { }")
    s = s[:a] + '''    // gamma / beta were bulk-copied into the (idle) partial buffer during the preceding barrier
    while (!mbar_try_wait(pbar, rs.ppar)) { }
    rs.ppar ^= 1u;
    if (pr) pr[8] = global_timer_ns();
    ln_rows_ring<D>(xb, partial, sd->X, sd->x_ld, T);
''' + s[b:]
    fn = '''template <int D>
__device__ __noinline__ void ln_rows_ring(unsigned char* xb, const float* gb, const float* X, int x_ld, int T) {
  using G = RingGeom<D>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = WM_DEC_THREADS >> 5;
  for (int r = warp; r < T; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(X + (size_t)r * x_ld) + lane;
    float4 v[G::NV];
#pragma unroll
    for (int i = 0; i < G::NV; ++i) {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i * 32 + lane < G::NV4) v[i] = __ldcg(src + i * 32);
    }
    uint4* const row = reinterpret_cast<uint4*>(xb + (size_t)r * G::XS) + lane;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G::NV; ++i)
      if (i * 32 + lane < G::NV4) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < G::NV; ++i)
      if (i * 32 + lane < G::NV4) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
#pragma unroll
    for (int i = 0; i < G::NV; ++i)
      if (i * 32 + lane < G::NV4) {
        const float4 gg = reinterpret_cast<const float4*>(gb)[i * 32 + lane];
        const float4 bb = reinterpret_cast<const float4*>(gb)[G::NV4 + i * 32 + lane];
        float4 y;
        y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
        y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
        y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
        y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
        row[i * 32] = split_hilo4(y);
      }
  }
}

'''
    marker = "// ---------------------------------------------------------------------------------------------\n// GEMM stage fed from the ring"
    s = rep(s, marker, fn + marker)
    wr(d, 'decode_ring.cuh', s)

def tail_stages(d):
    """parallel accept + read-only LN vectors in final_ln (from the exp-ring branch)"""
    import subprocess
    t = rd(d, 'decode.cu')
    e = subprocess.run("git show exp-ring:whisper_medusa_b200/csrc/decode.cu", shell=True, capture_output=True, text=True).stdout
    # accept from exp-ring is newer than the commit? (edited after) -> take from working copy saved in /tmp/var/accept_new.txt
    new_accept = open('/tmp/var/accept_new.txt').read()
    a = t.index("__device__ __noinline__ void stage_accept(const DecModel* m, int ncta) {")
    b = t.index("// -----------------------------------------------------------------------------------------\n// GEMM descriptors of the stages")
    t = t[:a] + new_accept + t[b:]
    t = rep(t, "        const float4 gg = g4[i * 32 + lane], bb = b4[i * 32 + lane];", "        const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);   // (read-only path: may run ahead of the stores below)")
    wr(d, 'decode.cu', t)

def mma_unroll1(d):
    s = rd(d, 'decode_ring.cuh')
    s = rep(s, "#pragma unroll\n      for (int kk = 0; kk < G::KS; kk += 32) {", "#pragma unroll 1\n      for (int kk = 0; kk < G::KS; kk += 32) {")
    wr(d, 'decode_ring.cuh', s)

def xpre(d):
    """issue the X-row bulk copies of the NEXT GEMM stage from warp 0 the moment the grid barrier opens"""
    s = rd(d, 'decode_ring.cuh')
    # 1. gemm stage: no issue, only wait
    a = s.index("  // ---- X rows: global (L2) -> shared, one bulk copy per row ----\n  if (warp == 0) {")
    b = s.index("  } else if (warp == 1) {\n    // this CTA's bias slice of the NEXT GEMM stage")
    s = s[:a] + "  // ---- X rows: global (L2) -> shared, one bulk copy per row: issued by warp 0 the moment the preceding grid barrier\n  // opened (ring_issue_x_rows, called from the kernel's main loop) ----\n  if (false) {\n" + s[b:]
    # 2. helper
    helper = '''// X rows of a GEMM stage: T bulk copies (one per row, lanes of warp 0) landing on `xbar`.  Called by warp 0 right
// after the grid barrier that precedes the stage opened (and once before the first stage of a launch).
template <int D>
__device__ __forceinline__ void ring_issue_x_rows(unsigned char* smem, const CtaStage* nd, int Tpass) {
  using G = RingGeom<D>;
  const int lane = threadIdx.x & 31;
  if (!is_gemm_stage(nd->stage) || nd->n_rows == 0) return;
  const int T = nd->x_rows_fixed ? nd->x_rows_fixed : Tpass;
  unsigned char* const xb = smem + G::SCRATCH_OFF;
  uint64_t* const xbar = reinterpret_cast<uint64_t*>(smem + G::BAR_OFF) + 2 * WM_RING_G;
  if (lane == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of the buffer vs async writes
    mbar_expect_tx(xbar, (uint32_t)(T * D * 4));
  }
  __syncwarp();
  if (lane < T) bulk_g2s(xb + (size_t)lane * G::XS, nd->X + (size_t)lane * nd->x_ld, (uint32_t)(D * 4), xbar);
}

'''
    marker = "template <int D, bool PROF>\n__global__ void __launch_bounds__(WM_RING_THREADS, 1)\ndec_iteration_ring_kernel("
    s = rep(s, marker, helper + marker)
    # 3. main loop: first stage + barrier with prefetch
    s = rep(s, "  cta_sync();\n  for (int ip = ip_first; ip < ip_last; ++ip) {\n    const CtaStage* sd = &s_desc[ip & 1];",
      '''  cta_sync();
  auto pass_rows = [&](int mode) { return mode == MODE_A ? L0 - kv0 : (mode == MODE_B ? m->K + 1 : 1); };
  if (warp == 0) ring_issue_x_rows<D>(smem, &s_desc[ip_first & 1], pass_rows(s_desc[ip_first & 1].mode));
  for (int ip = ip_first; ip < ip_last; ++ip) {
    const CtaStage* sd = &s_desc[ip & 1];''')
    s = rep(s, "    epoch = grid_barrier_step<false>(m->bar, epoch, ncta);",
      '''    {
      // grid barrier (see grid_barrier_step), with warp 0 pulling the next stage's X rows the moment it opens
      asm volatile("fence.proxy.async.global;" ::: "memory");
      cta_sync();
      if (warp == 0) {
        if (lane == 0) {
          const unsigned int target = (unsigned int)ncta * (epoch + 1u);
          red_add_release(&m->bar[0], 1u);
          while (ld_relaxed_u32(&m->bar[0]) < target) { }
        }
        __syncwarp();
        if (ip + 1 < ip_last) ring_issue_x_rows<D>(smem, &s_desc[(ip + 1) & 1], pass_rows(s_desc[(ip + 1) & 1].mode));
      }
      cta_sync();
      epoch += 1u;
    }''')
    wr(d, 'decode_ring.cuh', s)

def crossfuse(d):
    t = rd(d, 'decode.cu')
    a = t.index("  cta_sync();\n  if (pr) pr[4] = wm_timer_ns();\n  // ---- S = Q K^T * head_dim^-0.5 : warp w takes key tiles")
    b = t.index("  // ---- O = P V : warp w < 8 owns output dims")
    e = open('/tmp/exp_decode.cu').read()
    ea = e.index("  cta_sync();\n  if (pr) pr[4] = wm_timer_ns();\n  // ---- S = Q K^T * head_dim^-0.5 : warp w takes key tiles")
    eb = e.index("  // ---- O = P V : warp w < 8 owns output dims")
    t = t[:a] + e[ea:eb] + t[b:]
    t = rep(t, '''    out[64] = sM[tid];
    out[65] = sM[WM_MAX_T + tid];''', '''    float M = -INFINITY, sum = 0.f;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) { M = fmaxf(M, sWm[w * 16 + tid]); sum += sWs[w * 16 + tid]; }
    out[64] = M;
    out[65] = sum;''')
    t = rep(t, '''  return (size_t)WM_MAX_T * WM_SS_STRIDE * sizeof(float) + (size_t)2 * 16 * 72 * sizeof(__half) + (size_t)2 * WM_MAX_T * sizeof(float);
}''', '''  return (size_t)WM_MAX_T * WM_SS_STRIDE * sizeof(float) + (size_t)2 * 16 * 72 * sizeof(__half) +
         (size_t)2 * (WM_DEC_THREADS / 32) * 16 * sizeof(float);
}''')
    t = rep(t, "(size_t)2 * 16 * 72 * sizeof(__half) + (size_t)2 * WM_MAX_T * sizeof(float);   // K, V chunk + cross_scratch_bytes()",
               "(size_t)2 * 16 * 72 * sizeof(__half) + (size_t)2 * (WM_DEC_THREADS / 32) * 16 * sizeof(float);   // K, V chunk + cross_scratch_bytes()")
    wr(d, 'decode.cu', t)

def selfring(d):
    s = rd(d, 'decode_ring.cuh')
    e = open('/tmp/exp_ring.cuh').read()
    # defines + smem size
    a = e.index("// on-chip staging of the ring kernel's self-attention stage (stage_self_attn_ring)")
    b = e.index("#define WM_XS_PADB 16")
    s = rep(s, "#define WM_XS_PADB 16", e[a:b] + "#define WM_XS_PADB 16")
    s = rep(s, "static constexpr size_t SCRATCH = round128(cmax(cmax((size_t)16 * XS, cross_scratch_bytes()), self_attn_smem_bytes()));",
               "static constexpr size_t SCRATCH = round128(cmax(cmax((size_t)16 * XS, cross_scratch_bytes()), cmax(self_attn_smem_bytes(), self_attn_ring_smem_bytes())));")
    a = e.index("// ---------------------------------------------------------------------------------------------\n// causal self-attention of the ring kernel")
    b = e.index("// ---------------------------------------------------------------------------------------------\n// cross-attention fed from the ring")
    marker = "// ---------------------------------------------------------------------------------------------\n// cross-attention fed from the ring"
    s = rep(s, marker, e[a:b] + marker)
    s = rep(s, "      stage_self_attn<true>(m, mode, sd->layer, cta, ncta, smem + G::SCRATCH_OFF, &pgv, pr);",
               "      stage_self_attn_ring<D>(smem, m, sd->layer, pgv.T, pgv.base, cta, ncta, pr);")
    wr(d, 'decode_ring.cuh', s)
    c = rd(d, 'common.cuh')
    c = rep(c, "__device__ __forceinline__ uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }",
'''__device__ __forceinline__ uint4 ldcg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
// same, but issued where it is written (the compiler may not sink it towards its first use)
__device__ __forceinline__ uint4 ldcg_u4_now(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}''')
    wr(d, 'common.cuh', c)

def grouped(d):
    """stages of 2..3 units: ONE pass over the activations for all units (X fragments read once), epilogue by all threads"""
    s = rd(d, 'decode_ring.cuh')
    a = s.index("  // ---- unit loop, warp-specialised: warps 0..NKS-1 run the MMAs of unit u while the remaining warps finish unit")
    b = s.index("  cta_sync();   // every store of the stage issued; partial buffer and X buffer free")
    old_body = s[a:b]
    new = r'''  // ---- stages of 2 .. 3 units (QKV, FC1, FC2) with at most 11 token rows: ONE pass over the activations for all units.
  // The A fragments of X (2/3 of the shared-memory traffic of a unit) are read once per k-step and reused for the
  // weight rows of every unit; the k-slice partials of units 1, 2 go to the idle rows 11 .. 15 of the X buffer; the
  // epilogue is shared by all threads.  Needs all chunks of the stage in the ring at once (WM_RING_G = 3 slots).
  constexpr int GU = 3;   // units of a grouped stage (<= ring slots)
  constexpr bool CAN_GROUP = (size_t)5 * G::XS >= (size_t)(GU - 1) * G::NKS * 256 * sizeof(float) && WM_RING_G >= GU;
  if (CAN_GROUP && units >= 2 && units <= GU && T <= 11 && epi != EPI_HEADS_A && epi != EPI_HEAD_B) {
    float* const part2 = reinterpret_cast<float*>(xb + (size_t)11 * G::XS);
    // this thread's outputs (o = tid + k * threads over units * 256): bias / residual loads in flight during the MMAs
    constexpr int GOUT = (GU * 256 + WM_DEC_THREADS - 1) / WM_DEC_THREADS;
    float bias_v[GOUT], old[GOUT];
#pragma unroll
    for (int k = 0; k < GOUT; ++k) {
      const int o = tid + k * WM_DEC_THREADS, u = o >> 8, token = (o >> 4) & 15, rloc = o & 15;
      bias_v[k] = 0.f; old[k] = 0.f;
      if (u < units && token < T && u * 16 + rloc < n_rows && !ksplit) {
        const float* bias = sd->bias;
        if (bias) bias_v[k] = bias[n_begin + u * 16 + rloc];
        if (epi == EPI_RESID) old[k] = ldcg_f(&sd->out[(size_t)token * sd->ldo + n_begin + u * 16 + rloc]);
      }
    }
    if (warp < G::NKS) {
      int slot[GU];
#pragma unroll
      for (int u = 0; u < GU; ++u) {
        slot[u] = rs.slot;
        if (u < units) {
          while (!mbar_try_wait(full + rs.slot, rs.par)) { }
          if (++rs.slot == WM_RING_G) { rs.slot = 0; rs.par ^= 1u; }
        }
      }
      if (pr) pr[4] = global_timer_ns();
      const unsigned char* x0 = xb + (size_t)gq * G::XS + (size_t)(warp * G::KS + 8 * tq) * 4;
      const unsigned char* x1 = x0 + (size_t)8 * G::XS;
      const bool t1 = (gq + 8) < T;
      const uint4 z = make_uint4(0, 0, 0, 0);
      float ch[GU][8];   // (hi and lo parts of X accumulate into the same chain: 2 x GU independent chains are enough)
#pragma unroll
      for (int u = 0; u < GU; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[u][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < G::KS; kk += 32) {
        const uint4 p0 = *reinterpret_cast<const uint4*>(x0 + kk * 4);
        const uint4 p1 = *reinterpret_cast<const uint4*>(x0 + kk * 4 + 16);
        const uint4 q0 = t1 ? *reinterpret_cast<const uint4*>(x1 + kk * 4) : z;
        const uint4 q1 = t1 ? *reinterpret_cast<const uint4*>(x1 + kk * 4 + 16) : z;
#pragma unroll
        for (int u = 0; u < GU; ++u)
          if (u < units) {
            const int nvalid = min(16, n_rows - u * 16);
            const __half* w0p = reinterpret_cast<const __half*>(smem + (size_t)slot[u] * G::SLOT_BYTES) +
                                (size_t)gq * (G::ROW_STRIDE / 2) + warp * G::KS + 8 * tq + kk;
            const uint4 wa = (gq < nvalid) ? *reinterpret_cast<const uint4*>(w0p) : z;
            const uint4 wb = (gq + 8 < nvalid) ? *reinterpret_cast<const uint4*>(w0p + (size_t)8 * (G::ROW_STRIDE / 2)) : z;
            mma_16816(&ch[u][0], p0.x, q0.x, p0.z, q0.z, wa.x, wa.y);
            mma_16816(&ch[u][0], p0.y, q0.y, p0.w, q0.w, wa.x, wa.y);
            mma_16816(&ch[u][4], p0.x, q0.x, p0.z, q0.z, wb.x, wb.y);
            mma_16816(&ch[u][4], p0.y, q0.y, p0.w, q0.w, wb.x, wb.y);
            mma_16816(&ch[u][0], p1.x, q1.x, p1.z, q1.z, wa.z, wa.w);
            mma_16816(&ch[u][0], p1.y, q1.y, p1.w, q1.w, wa.z, wa.w);
            mma_16816(&ch[u][4], p1.x, q1.x, p1.z, q1.z, wb.z, wb.w);
            mma_16816(&ch[u][4], p1.y, q1.y, p1.w, q1.w, wb.z, wb.w);
          }
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < GU; ++u)
          if (u < units) mbar_arrive(empty + slot[u]);   // this warp is done with the slots
      }
#pragma unroll
      for (int u = 0; u < GU; ++u)
        if (u < units) {
          float* dst = (u == 0) ? partial : part2 + (size_t)(u - 1) * G::NKS * 256;
#pragma unroll
          for (int e = 0; e < 8; ++e) dst[warp * 256 + e * 32 + lane] = ch[u][e];
        }
      if (pr) pr[5] = global_timer_ns();
    } else {
      for (int u = 0; u < units; ++u)
        if (++rs.slot == WM_RING_G) { rs.slot = 0; rs.par ^= 1u; }   // keep the (uniform) ring state in step with the MMA warps
    }
    cta_sync();   // partials of all units written
    if (pr) pr[6] = global_timer_ns();
    float* out = sd->out;
    const int ldo = sd->ldo;
#pragma unroll
    for (int k = 0; k < GOUT; ++k) {
      const int o = tid + k * WM_DEC_THREADS, u = o >> 8, token = (o >> 4) & 15, rloc = o & 15;
      if (!(u < units && token < T && u * 16 + rloc < n_rows)) continue;
      const float* src = (u == 0) ? partial : part2 + (size_t)(u - 1) * G::NKS * 256;
      const int idx = (((rloc >> 3) * 4) + ((token >= 8) ? 2 : 0) + (rloc & 1)) * 32 + (token & 7) * 4 + ((rloc & 7) >> 1);
      float s = 0.f;
#pragma unroll
      for (int ks = 0; ks < G::NKS; ++ks) s += src[ks * 256 + idx];
      const int row = n_begin + u * 16 + rloc;
      if (ksplit) {
        m->gemm_part[((size_t)sd->seg * 16 + token) * sd->N + row] = s;
      } else if (epi == EPI_RESID) {
        out[(size_t)token * ldo + row] = old[k] + (s + bias_v[k]);
      } else if (epi == EPI_GELU) {
        const float v = gelu_erf(s + bias_v[k]);
        if (sd->out_split) store_split(out + (size_t)token * ldo, row, v);
        else out[(size_t)token * ldo + row] = v;
      } else if (epi == EPI_QKV) {
        const float v = s + bias_v[k];
        const DecLayer& L = m->layers[sd->layer];
        if (row < D) out[(size_t)token * ldo + row] = v;
        else if (row < 2 * D) L.self_k[(size_t)(base + token) * D + (row - D)] = __float2half_rn(v);
        else L.self_v[(size_t)(base + token) * D + (row - 2 * D)] = __float2half_rn(v);
      } else {   // EPI_STORE / EPI_LOGITS (the Medusa-head epilogues never have 2 .. 3 units per CTA on a full grid; see below)
        out[(size_t)token * ldo + row] = s + bias_v[k];
      }
    }
  } else {
'''
    # indent old body? keep as is inside else { }
    s = s[:a] + new + old_body + "  }\n" + s[b:]
    wr(d, 'decode_ring.cuh', s)

def bar4(d):
    """grid barrier with 4 arrival counters in different L2 lines (the REDs of 148 CTAs on one word serialise)"""
    s = rd(d, 'decode_ring.cuh')
    helper = '''// Grid barrier of the ring kernel: like grid_barrier_step<false> (decode.cu) but with 4 arrival counters in
// different L2 lines (CTA c arrives on counter c & 3): the release REDs of all CTAs on ONE word serialise in its L2 slice.
__device__ __forceinline__ unsigned int ring_barrier(unsigned int* bar, unsigned int epoch, int cta, int ncta) {
  asm volatile("fence.proxy.async.global;" ::: "memory");
  cta_sync();
  if (threadIdx.x == 0) {
    red_add_release(&bar[32 * (1 + (cta & 3))], 1u);
    const unsigned int e1 = epoch + 1u;
    const unsigned int t0 = (unsigned int)((ncta + 3) >> 2) * e1, t1 = (unsigned int)((ncta + 2) >> 2) * e1;
    const unsigned int t2 = (unsigned int)((ncta + 1) >> 2) * e1, t3 = (unsigned int)(ncta >> 2) * e1;
    for (;;) {
      const unsigned int v0 = ld_relaxed_u32(&bar[32]), v1 = ld_relaxed_u32(&bar[64]);
      const unsigned int v2 = ld_relaxed_u32(&bar[96]), v3 = ld_relaxed_u32(&bar[128]);
      if (v0 >= t0 && v1 >= t1 && v2 >= t2 && v3 >= t3) break;
    }
  }
  cta_sync();
  return epoch + 1u;
}

'''
    marker = "template <int D, bool PROF>\n__global__ void __launch_bounds__(WM_RING_THREADS, 1)\ndec_iteration_ring_kernel("
    s = rep(s, marker, helper + marker)
    s = rep(s, "    epoch = grid_barrier_step<false>(m->bar, epoch, ncta);", "    epoch = ring_barrier(m->bar, epoch, cta, ncta);")
    wr(d, 'decode_ring.cuh', s)
    e = rd(d, 'engine.cu')
    e = rep(e, "  CK(dalloc(&h->bar, 8));", "  CK(dalloc(&h->bar, 256));")
    e = rep(e, "    CK(cudaMemsetAsync(h->bar, 0, 8 * sizeof(unsigned int), s));", "    CK(cudaMemsetAsync(h->bar, 0, 256 * sizeof(unsigned int), s));")
    wr(d, 'engine.cu', e)

def mcast(d):
    """2-CTA clusters: the X rows of a GEMM stage are fetched once per CTA pair (each CTA issues every other row as a
    multicast bulk copy that lands in both CTAs' shared memory).  EXPERIMENT: assumes every CTA has rows in every stage."""
    s = rd(d, 'decode_ring.cuh')
    s = rep(s, "__device__ __forceinline__ unsigned long long global_timer_ns() {", '''__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {''')
    s = rep(s, "    if (lane < T) bulk_g2s(xb + (size_t)lane * G::XS, sd->X + (size_t)lane * sd->x_ld, (uint32_t)(D * 4), xbar);",
               "    if (lane < T && (uint32_t)(lane & 1) == cluster_rank())\n      bulk_g2s_mc(xb + (size_t)lane * G::XS, sd->X + (size_t)lane * sd->x_ld, (uint32_t)(D * 4), xbar, (uint16_t)3);")
    s = rep(s, "template <int D, bool PROF>\n__global__ void __launch_bounds__(WM_RING_THREADS, 1)\ndec_iteration_ring_kernel(",
               "template <int D, bool PROF>\n__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WM_RING_THREADS, 1)\ndec_iteration_ring_kernel(")
    s = rep(s, "  __syncthreads();   // the only full-CTA barrier: after it the producer warp goes its own way",
               "  __syncthreads();   // the only full-CTA barrier: after it the producer warp goes its own way\n  // the peer CTA may multicast into this CTA's buffers from its first stage on: barriers initialised cluster-wide first\n  asm volatile(\"barrier.cluster.arrive.release.aligned;\\nbarrier.cluster.wait.acquire.aligned;\" ::: \"memory\");")
    wr(d, 'decode_ring.cuh', s)
    t = rd(d, 'decode.cu')
    t = rep(t, "  if (!ACQUIRE) asm volatile(\"fence.proxy.async.global;\" ::: \"memory\");", "  if (!ACQUIRE) asm volatile(\"fence.proxy.async;\" ::: \"memory\");")
    wr(d, 'decode.cu', t)

def poll4(d):
    """grid barrier: the polling thread keeps 4 loads of the counter in flight (detection delay ~ RTT/4 instead of RTT/2 + RTT)"""
    t = rd(d, 'decode.cu')
    t = rep(t, "    while (ld_relaxed_u32(&bar[0]) < target) { }\n    if (ACQUIRE) __threadfence();",
'''    if (ACQUIRE) {
      while (ld_relaxed_u32(&bar[0]) < target) { }
      __threadfence();
    } else {
      // pipelined poll: 4 loads of the counter in flight, re-issued one by one as they come back
      unsigned int v0 = ld_relaxed_u32(&bar[0]);
      unsigned int v1 = ld_relaxed_u32(&bar[0]);
      unsigned int v2 = ld_relaxed_u32(&bar[0]);
      unsigned int v3 = ld_relaxed_u32(&bar[0]);
      for (;;) {
        if (v0 >= target) break;
        v0 = ld_relaxed_u32(&bar[0]);
        if (v1 >= target) break;
        v1 = ld_relaxed_u32(&bar[0]);
        if (v2 >= target) break;
        v2 = ld_relaxed_u32(&bar[0]);
        if (v3 >= target) break;
        v3 = ld_relaxed_u32(&bar[0]);
      }
    }''')
    wr(d, 'decode.cu', t)

def pollwarp(d):
    """grid barrier: the 32 lanes of warp 0 poll, lane i starting i * 16 ns after lane 0 (staggered by a dependent delay)"""
    t = rd(d, 'decode.cu')
    t = rep(t, "  cta_sync();\n  if (threadIdx.x == 0) {\n    const unsigned int target = (unsigned int)ncta * (epoch + 1u);\n    red_add_release(&bar[0], 1u);\n    while (ld_relaxed_u32(&bar[0]) < target) { }\n    if (ACQUIRE) __threadfence();\n  }\n  cta_sync();",
'''  cta_sync();
  if (!ACQUIRE) {
    if (threadIdx.x < 32) {
      const unsigned int target = (unsigned int)ncta * (epoch + 1u);
      if (threadIdx.x == 0) red_add_release(&bar[0], 1u);
      __syncwarp();
      // 8 lanes poll, de-phased: any lane that sees the target ends the wait for the warp
      bool done = false;
      if ((threadIdx.x & 3) == 0) __nanosleep(20u * (threadIdx.x >> 2));
      for (;;) {
        if ((threadIdx.x & 3) == 0) done = ld_relaxed_u32(&bar[0]) >= target;
        if (__any_sync(0xffffffffu, done)) break;
      }
    }
  } else if (threadIdx.x == 0) {
    const unsigned int target = (unsigned int)ncta * (epoch + 1u);
    red_add_release(&bar[0], 1u);
    while (ld_relaxed_u32(&bar[0]) < target) { }
    __threadfence();
  }
  cta_sync();''')
    wr(d, 'decode.cu', t)

def epifence(d):
    """epilogue warps fence their own stores (in parallel) before the final CTA barrier of a GEMM stage"""
    s = rd(d, 'decode_ring.cuh')
    s = rep(s, "    for (int u = 0; u < units; ++u)\n      if (++rs.slot == WM_RING_G) { rs.slot = 0; rs.par ^= 1u; }   // keep the (uniform) ring state in step with the MMA warps\n  }\n  cta_sync();   // every store of the stage issued",
               "    for (int u = 0; u < units; ++u)\n      if (++rs.slot == WM_RING_G) { rs.slot = 0; rs.par ^= 1u; }   // keep the (uniform) ring state in step with the MMA warps\n    __threadfence();   // this thread's stores are in L2 before the CTA's barrier arrival (whose fence then finds nothing to wait for)\n  }\n  cta_sync();   // every store of the stage issued")
    wr(d, 'decode_ring.cuh', s)

def _bar_body(d, body):
    t = rd(d, 'decode.cu')
    t = rep(t, "    while (ld_relaxed_u32(&bar[0]) < target) { }\n    if (ACQUIRE) __threadfence();", body)
    wr(d, 'decode.cu', t)

def ns100(d):
    _bar_body(d, "    if (ACQUIRE) { while (ld_relaxed_u32(&bar[0]) < target) { } __threadfence(); }\n    else { while (ld_relaxed_u32(&bar[0]) < target) { __nanosleep(100); } }")

def ns300(d):
    _bar_body(d, "    if (ACQUIRE) { while (ld_relaxed_u32(&bar[0]) < target) { } __threadfence(); }\n    else { while (ld_relaxed_u32(&bar[0]) < target) { __nanosleep(300); } }")

def first400(d):
    _bar_body(d, "    if (ACQUIRE) { while (ld_relaxed_u32(&bar[0]) < target) { } __threadfence(); }\n    else { __nanosleep(400); while (ld_relaxed_u32(&bar[0]) < target) { } }")

def flagbc(d):
    """arrivals on one line (atom with return), the last arriver publishes the epoch on another line that everybody polls"""
    t = rd(d, 'decode.cu')
    t = rep(t, "    red_add_release(&bar[0], 1u);\n    while (ld_relaxed_u32(&bar[0]) < target) { }\n    if (ACQUIRE) __threadfence();",
'''    if (ACQUIRE) {
      red_add_release(&bar[0], 1u);
      while (ld_relaxed_u32(&bar[0]) < target) { }
      __threadfence();
    } else {
      // arrivals and polls on different L2 lines: the release atomics of 148 CTAs and their polling loads on ONE line
      // serialise in its slice.  The last arriver (its atomic returns target - 1) publishes the epoch on the poll line.
      const unsigned int prev = atom_add_release(&bar[0], 1u);
      if (prev + 1u == target) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&bar[64]), "r"(epoch + 1u) : "memory");
      else while (ld_relaxed_u32(&bar[64]) < epoch + 1u) { }
    }''')
    wr(d, 'decode.cu', t)
    e = rd(d, 'engine.cu')
    e = rep(e, "  CK(dalloc(&h->bar, 8));", "  CK(dalloc(&h->bar, 256));")
    e = rep(e, "    CK(cudaMemsetAsync(h->bar, 0, 8 * sizeof(unsigned int), s));", "    CK(cudaMemsetAsync(h->bar, 0, 256 * sizeof(unsigned int), s));")
    wr(d, 'engine.cu', e)

def enc2cta(d):
    """encoder GEMM: 3-stage ring (97 KB) and two CTAs per SM: one CTA's epilogue overlaps the other's main loop"""
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "#define TC_STAGES 6", "#define TC_STAGES 3")
    s = rep(s, "__global__ void __launch_bounds__(TC_THREADS, 1)\nenc_gemm_tc_kernel(", "__global__ void __launch_bounds__(TC_THREADS, 2)\nenc_gemm_tc_kernel(")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_bn64(d):
    """encoder GEMM: 128 x 64 output tiles when N <= 1536 (120 tiles of 128 x 128 leave SMs idle)"""
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "#define TC_STAGE_BYTES ((TC_BM + TC_BN) * TC_BK * 2)", "#define TC_STAGE_BYTES ((TC_BM + TC_BN) * TC_BK * 2)   /* layout of a stage for every tile width: A at 0, W at 16 KB */")
    s = rep(s, "__device__ __forceinline__ constexpr uint32_t tc_instr_desc() {\n  return (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);\n}",
               "__device__ __forceinline__ constexpr uint32_t tc_instr_desc(int bn) {\n  return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);\n}")
    s = rep(s, "template <int EPI>\n__global__ void __launch_bounds__(TC_THREADS,", "template <int EPI, int BN>\n__global__ void __launch_bounds__(TC_THREADS,")
    s = rep(s, "  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;", "  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;")
    s = rep(s, "        tc_mbar_expect_tx(full, TC_STAGE_BYTES);", "        tc_mbar_expect_tx(full, (TC_BM + BN) * TC_BK * 2);")
    s = rep(s, "      const uint32_t idesc = tc_instr_desc();", "      const uint32_t idesc = tc_instr_desc(BN);")
    s = rep(s, "    for (int cb = 0; cb < TC_BN / 32; ++cb) {", "    for (int cb = 0; cb < BN / 32; ++cb) {")
    # tensor map for W with BN rows per box
    s = rep(s, "static bool make_map(CUtensorMap* map, const __half* base, uint64_t rows, uint64_t K, uint64_t ld) {", "static bool make_map(CUtensorMap* map, const __half* base, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows) {")
    s = rep(s, "  cuuint32_t box[2] = {TC_BK, TC_BM};", "  cuuint32_t box[2] = {TC_BK, box_rows};")
    s = rep(s, "  e = cudaFuncSetAttribute(enc_gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem); \\\n  if (e != cudaSuccess) return e;",
               "  e = cudaFuncSetAttribute(enc_gemm_tc_kernel<EPI, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem); \\\n  if (e != cudaSuccess) return e;                                                                                 \\\n  e = cudaFuncSetAttribute(enc_gemm_tc_kernel<EPI, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);  \\\n  if (e != cudaSuccess) return e;")
    s = rep(s, "  typedef std::tuple<const void*, uint64_t, uint64_t, uint64_t> Key;", "  typedef std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t> Key;")
    s = rep(s, "  auto get = [&](const __half* base, uint64_t rows, uint64_t K, uint64_t ld, CUtensorMap* out) -> bool {\n    Key k(base, rows, K, ld);",
               "  auto get = [&](const __half* base, uint64_t rows, uint64_t K, uint64_t ld, uint32_t box_rows, CUtensorMap* out) -> bool {\n    Key k(base, rows, K, ld, box_rows);")
    s = rep(s, "      if (!make_map(&m, base, rows, K, ld)) return false;", "      if (!make_map(&m, base, rows, K, ld, box_rows)) return false;")
    s = rep(s, "  if (!get(g.A, (uint64_t)a_rows, (uint64_t)g.K, (uint64_t)g.lda, &ma)) return cudaErrorInvalidValue;\n  if (!get(g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.K, &mw)) return cudaErrorInvalidValue;",
               "  const int bn = (g.N <= 1536 && g.N % 64 == 0) ? 64 : TC_BN;   // narrow outputs: twice the tiles, all SMs busy\n  if (!get(g.A, (uint64_t)a_rows, (uint64_t)g.K, (uint64_t)g.lda, TC_BM, &ma)) return cudaErrorInvalidValue;\n  if (!get(g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.K, (uint32_t)bn, &mw)) return cudaErrorInvalidValue;")
    s = rep(s, "  dim3 grid(g.N / TC_BN, (g.M + TC_BM - 1) / TC_BM);", "  dim3 grid(g.N / bn, (g.M + TC_BM - 1) / TC_BM);")
    for e in ["ENC_EPI_BIAS_F16", "ENC_EPI_BIAS_GELU_F16", "ENC_EPI_BIAS_RES_F32", "ENC_EPI_BIAS_GELU_POS_F32"]:
        s = rep(s, f"    case {e}: enc_gemm_tc_kernel<{e}><<<grid, TC_THREADS, kTcSmem, s>>>(ma, mw, a); break;",
                   f"    case {e}:\n      if (bn == 64) enc_gemm_tc_kernel<{e}, 64><<<grid, TC_THREADS, kTcSmem, s>>>(ma, mw, a);\n      else enc_gemm_tc_kernel<{e}, 128><<<grid, TC_THREADS, kTcSmem, s>>>(ma, mw, a);\n      break;")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_thr2560(d):
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "(g.N <= 1536 && g.N % 64 == 0) ? 64 : TC_BN;", "(g.N <= 2560 && g.N % 64 == 0) ? 64 : TC_BN;")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_all64(d):
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "(g.N <= 1536 && g.N % 64 == 0) ? 64 : TC_BN;", "(g.N % 64 == 0) ? 64 : TC_BN;")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_epi2(d):
    """epilogue: two TMEM loads in flight before the wait"""
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, '      : "r"(taddr)\n      : "memory");\n  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");\n}', '      : "r"(taddr)\n      : "memory");\n}\n__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }')
    s = rep(s, "#pragma unroll 1\n    for (int cb = 0; cb < BN / 32; ++cb) {\n      uint32_t v[32];\n      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);\n      if (!row_ok) continue;",
'''    uint32_t vbuf[2][32];
    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16), vbuf[0]);
#pragma unroll
    for (int cb = 0; cb < BN / 32; ++cb) {
      tc_ld_wait();
      if (cb + 1 < BN / 32) tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((cb + 1) * 32), vbuf[(cb + 1) & 1]);
      uint32_t (&v)[32] = vbuf[cb & 1];
      if (!row_ok) continue;''')
    wr(d, 'enc_gemm_tc.cu', s)

def enc3cta(d):
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "#define TC_STAGES 3", "#define TC_STAGES 2")
    s = rep(s, "__global__ void __launch_bounds__(TC_THREADS, 2)\nenc_gemm_tc_kernel(", "__global__ void __launch_bounds__(TC_THREADS, 3)\nenc_gemm_tc_kernel(")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_thr0(d):
    """128x128 tiles everywhere (2 CTAs/SM)"""
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "(g.N <= 1536 && g.N % 64 == 0) ? 64 : TC_BN;", "TC_BN;")
    wr(d, 'enc_gemm_tc.cu', s)

def enc_epi8(d):
    """8 epilogue warps: two per TMEM lane quarter, alternating 32-column blocks"""
    s = rd(d, 'enc_gemm_tc.cu')
    s = rep(s, "#define TC_THREADS 192", "#define TC_THREADS 320")
    s = rep(s, "#pragma unroll 1\n    for (int cb = 0; cb < BN / 32; ++cb) {", "#pragma unroll 1\n    for (int cb = (warp - 2) >> 2; cb < BN / 32; cb += 2) {")
    wr(d, 'enc_gemm_tc.cu', s)
