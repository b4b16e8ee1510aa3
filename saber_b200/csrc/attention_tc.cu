// saber_b200 — single-head, head_dim-256 flash attention on tcgen05 / TMEM / TMA: SAM2's memory attention
// (sam2/modeling/memory_attention.py MemoryAttentionLayer: RoPE self-attention over the 4096 frame tokens and RoPE
// cross-attention to the <= 8256-token memory bank, 1 head x 256; upstream calls F.scaled_dot_product_attention).
// The mma.sync kernel (attention.cu, HDP = 256: Q in shared memory, 32-key tiles) runs these at ~200 TFLOP/s and is
// 46 % of a propagated frame; here
//   S = Q K^T : UMMA M128 N64 K256 (Q tile resident, K tile K-major from the TMA ring), accumulator in TMEM
//   O += P V  : UMMA M128 N256 K64 (P written by the softmax warps as a K-major 128B-swizzled tile, V tile read through
//               an MN-major descriptor straight from its row-major TMA staging), accumulator in TMEM (fp32, 256 columns)
// CTA = (128-query tile, batch entry). Warp 0 TMA producer (K and V rings of two 64-key tiles each, released
// separately: K after QK^T, V after PV), warp 1 MMA issuer (converged warp, elected lane), warp 2 TMEM allocator,
// warps 4-11 online softmax: thread = (query row, 32-key half), row maxima exchanged through shared memory, lazy
// rescaling of O (only when a row maximum grows by more than 2^8), final 1/l scaling and a coalesced bf16 store.
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int AT_KT = 64;
constexpr int AT_THREADS = 384;
constexpr int AT_OFF_Q = 0;                          // 4 K-blocks x [128 rows x 128 B]
constexpr int AT_OFF_K = 65536;                      // 2 stages x 4 K-blocks x [64 keys x 128 B]
constexpr int AT_OFF_V = AT_OFF_K + 2 * 32768;       // 2 stages x 4 blocks x [64 keys x 128 B]
constexpr int AT_OFF_P = AT_OFF_V + 2 * 32768;       // 2 x [128 rows x 128 B]; output staging after the last PV
constexpr int AT_OFF_BAR = AT_OFF_P + 2 * 16384;
constexpr int AT_OFF_XCH = AT_OFF_BAR + 256;         // [2 parities][2 halves][128 rows] fp32
constexpr int AT_SMEM = AT_OFF_XCH + 2048;           // 231,680 B

struct AttnTCParams {
  bf16* out;          // [B*nq, 256] (row pitch o_ld)
  long long o_ld;
  int nq, nk;
  int q_bstride, kv_bstride;  // rows between batch entries (0 = shared)
  float scale_log2;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void pair_barrier(int q) {
  asm volatile("bar.sync %0, 64;" ::"r"(q + 2) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t swz64(int row, int chunk) {
  return static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_d256_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTCParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 2
  uint64_t* p_full = bars + 11;     // 2
  uint64_t* pv_done = bars + 13;    // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y;
  const int T = (p.nk + AT_KT - 1) / AT_KT;

  if ((sb::smem_u32(smem) & 1023u) != 0u) __trap();
  if (warp == 0 && lane == 0) {
    sb::tma_prefetch_desc(&tmQ);
    sb::tma_prefetch_desc(&tmK);
    sb::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    sb::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      sb::mbar_init(&k_full[i], 1);
      sb::mbar_init(&k_empty[i], 1);
      sb::mbar_init(&v_full[i], 1);
      sb::mbar_init(&v_empty[i], 1);
      sb::mbar_init(&s_full[i], 1);
      sb::mbar_init(&p_full[i], 8);
      sb::mbar_init(&pv_done[i], 1);
    }
    sb::fence_barrier_init();
  }
  if (warp == 2) {
    sb::tmem_alloc(tmem_ptr, 512);
    sb::tmem_relinquish();
  }
  sb::tc_fence_before();
  __syncthreads();
  sb::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // TMEM columns: S buffers at [0,64) and [64,128); O at [128,384)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int qrow0 = b * p.q_bstride + qt * 128;
      sb::mbar_arrive_expect_tx(q_full, 65536);
      for (int kb = 0; kb < 4; ++kb) sb::tma_load_2d(smem + AT_OFF_Q + kb * 16384, &tmQ, q_full, kb * 64, qrow0);
      const int krow0 = b * p.kv_bstride;
      for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        if (t >= 2) sb::mbar_wait(&k_empty[s], static_cast<uint32_t>(((t >> 1) - 1) & 1));
        sb::mbar_arrive_expect_tx(&k_full[s], 32768);
        for (int kb = 0; kb < 4; ++kb)
          sb::tma_load_2d(smem + AT_OFF_K + s * 32768 + kb * 8192, &tmK, &k_full[s], kb * 64, krow0 + t * AT_KT);
        if (t >= 2) sb::mbar_wait(&v_empty[s], static_cast<uint32_t>(((t >> 1) - 1) & 1));
        sb::mbar_arrive_expect_tx(&v_full[s], 32768);
        for (int kb = 0; kb < 4; ++kb)
          sb::tma_load_2d(smem + AT_OFF_V + s * 32768 + kb * 8192, &tmV, &v_full[s], kb * 64, krow0 + t * AT_KT);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane per tcgen05 instruction) =====================
    constexpr uint32_t idesc_qk = sb::umma_idesc_bf16(128, 64);
    constexpr uint32_t idesc_pv = sb::umma_idesc_bf16(128, 256) | (1u << 16);  // B operand MN-major
    const uint32_t sbase = sb::smem_u32(smem);
    auto issue_qk = [&](int t) {
      const int s = t & 1;
      sb::mbar_wait(&k_full[s], static_cast<uint32_t>((t >> 1) & 1));
      sb::tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(s * 64);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t da = sb::umma_desc_k_sw128(sbase + AT_OFF_Q + kb * 16384);
        const uint64_t db = sb::umma_desc_k_sw128(sbase + AT_OFF_K + s * 32768 + kb * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (sb::elect_one())
            sb::umma_bf16(d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc_qk,
                          static_cast<uint32_t>((kb | k) != 0));
      }
      if (sb::elect_one()) {
        sb::umma_commit(&s_full[s]);
        sb::umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    sb::mbar_wait(q_full, 0);
    issue_qk(0);
    if (T > 1) issue_qk(1);
    for (int t = 0; t < T; ++t) {
      const int s = t & 1;
      sb::mbar_wait(&p_full[s], static_cast<uint32_t>((t >> 1) & 1));
      sb::mbar_wait(&v_full[s], static_cast<uint32_t>((t >> 1) & 1));
      sb::tc_fence_after();
      const uint32_t d = tmem_base + 128u;
      const uint64_t da = sb::umma_desc_k_sw128(sbase + AT_OFF_P + s * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = umma_desc_mn_sw128(sbase + AT_OFF_V + s * 32768 + k * 2048, 8192, 1024);
        if (sb::elect_one())
          sb::umma_bf16(d, da + static_cast<uint64_t>(2 * k), db, idesc_pv, static_cast<uint32_t>((t | k) != 0));
      }
      if (sb::elect_one()) {
        sb::umma_commit(&pv_done[s]);
        sb::umma_commit(&v_empty[s]);
      }
      __syncwarp();
      if (t + 2 < T) issue_qk(t + 2);
    }
  } else if (warp >= 4) {
    // ===================== online softmax: thread = (query row, key half) =====================
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sbase = sb::smem_u32(smem);
    float* xch = reinterpret_cast<float*>(smem + AT_OFF_XCH);
    const float c = p.scale_log2;
    float m = -INFINITY, l = 0.f;  // m in the scaled (log2) domain
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      sb::mbar_wait(&s_full[t & 1], static_cast<uint32_t>((t >> 1) & 1));
      sb::tc_fence_after();
      uint32_t v[32];
      const uint32_t ta = tmem_base + tlane + static_cast<uint32_t>((t & 1) * 64 + hf * 32);
      sb::tmem_ld_32x16(ta, v);
      sb::tmem_ld_32x16(ta + 16, v + 16);
      sb::tmem_ld_wait();
      const int kbase = t * AT_KT + hf * 32;
      if (kbase + 32 > p.nk) {  // ragged last tile: keys beyond nk are whatever the next batch entry holds
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (kbase + j >= p.nk) v[j] = __float_as_uint(-INFINITY);
      }
      float mx4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) mx4[j] = __uint_as_float(v[j]);
#pragma unroll
      for (int j = 4; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c;
      xch[((t & 1) * 2 + hf) * 128 + r] = mx;
      pair_barrier(q);
      mx = fmaxf(mx, xch[((t & 1) * 2 + (hf ^ 1)) * 128 + r]);
      if (__any_sync(0xffffffffu, mx > m + 8.f)) {
        const float mn = fmaxf(m, mx);
        const float alpha = sb::fast_exp2(m - mn);  // m = -inf on the first tile -> 0
        m = mn;
        l *= alpha;
        if (t > 0) {
          sb::mbar_wait(&pv_done[(t - 1) & 1], static_cast<uint32_t>(((t - 1) >> 1) & 1));
          sb::tc_fence_after();
          const uint32_t to = tmem_base + tlane + 128u + static_cast<uint32_t>(hf * 128);
#pragma unroll 1
          for (int cc = 0; cc < 8; ++cc) {
            uint32_t o[16];
            sb::tmem_ld_32x16(to + cc * 16, o);
            sb::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
            sb::tmem_st_32x16(to + cc * 16, o);
          }
          sb::tmem_st_wait();
        }
      }
      if (t >= 2) sb::mbar_wait(&pv_done[t & 1], static_cast<uint32_t>(((t >> 1) - 1) & 1));  // P(t-2) consumed
      const uint32_t prow = sbase + AT_OFF_P + (t & 1) * 16384 + r * 128;
      const float2 c2 = sb::splat2(c), nm2 = sb::splat2(-m);
      float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float2 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 d = sb::fma2(make_float2(__uint_as_float(v[ch * 8 + 2 * j]), __uint_as_float(v[ch * 8 + 2 * j + 1])), c2, nm2);
          e[j] = make_float2(sb::fast_exp2(d.x), sb::fast_exp2(d.y));
          l2[j & 1] = sb::add2(l2[j & 1], e[j]);
        }
        sts128(prow + (((hf * 4 + ch) ^ (r & 7)) << 4),
               make_uint4(sb::pack_bf16x2(e[0].x, e[0].y), sb::pack_bf16x2(e[1].x, e[1].y), sb::pack_bf16x2(e[2].x, e[2].y),
                          sb::pack_bf16x2(e[3].x, e[3].y)));
      }
      l += (l2[0].x + l2[0].y) + (l2[1].x + l2[1].y);
      sb::tc_fence_before();
      sb::fence_proxy_async();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&p_full[t & 1]);
    }
    // ---- O / l -> bf16, staged through the (now idle) P buffers for 64-byte-per-row coalesced stores
    pair_barrier(q);  // the partner has read this row's last maximum before the slot is reused for the row sums
    xch[hf * 128 + r] = l;
    pair_barrier(q);
    l += xch[(hf ^ 1) * 128 + r];
    const float inv = 1.f / l;
    sb::mbar_wait(&pv_done[(T - 1) & 1], static_cast<uint32_t>(((T - 1) >> 1) & 1));
    if (T > 1) sb::mbar_wait(&pv_done[(T - 2) & 1], static_cast<uint32_t>(((T - 2) >> 1) & 1));
    sb::tc_fence_after();
    const uint32_t stg = sbase + AT_OFF_P + (warp - 4) * 2048;
    const uint32_t to = tmem_base + tlane + 128u + static_cast<uint32_t>(hf * 128);
    uint8_t* orow = reinterpret_cast<uint8_t*>(p.out + (static_cast<long long>(b) * p.nq + qt * 128 + q * 32) * p.o_ld + hf * 128);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t o[32];
      sb::tmem_ld_32x16(to + cc * 32, o);
      sb::tmem_ld_32x16(to + cc * 32 + 16, o + 16);
      sb::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(stg + swz64(lane, j),
               make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv)));
      __syncwarp();
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const int id = tt * 32 + lane;
        const int row = id >> 2, ch = id & 3;
        *reinterpret_cast<uint4*>(orow + static_cast<long long>(row) * p.o_ld * 2 + cc * 64 + ch * 16) = lds128(stg + swz64(row, ch));
      }
      __syncwarp();
    }
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Called by sb_attention (attention.cu) for heads == 1, hd == 256, nq % 128 == 0, nk >= 64. Returns SB_OK, or
// SB_ERR_UNSUPPORTED when the operands do not fit the TMA path (the caller then uses the mma.sync kernel).
int sb_internal_attention_d256_tc(const void* q, long long q_ld, const void* k, long long k_ld, const void* v,
                                  long long v_ld, void* o, long long o_ld, int batch, int nq, int nk, float scale,
                                  int q_shared, int kv_shared, cudaStream_t stream) {
  if ((nq % 128) != 0 || nk < AT_KT || (q_ld % 8) != 0 || (k_ld % 8) != 0 || (v_ld % 8) != 0 || (o_ld % 8) != 0 ||
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(o)) & 15) != 0)
    return SB_ERR_UNSUPPORTED;
  CUtensorMap tmQ, tmK, tmV;
  int rc = sb_make_tmap_2d_bf16(&tmQ, q, static_cast<uint64_t>(q_shared ? 1 : batch) * nq, 256, static_cast<uint64_t>(q_ld), 128, 64);
  if (rc != SB_OK) return rc;
  rc = sb_make_tmap_2d_bf16(&tmK, k, static_cast<uint64_t>(kv_shared ? 1 : batch) * nk, 256, static_cast<uint64_t>(k_ld), AT_KT, 64);
  if (rc != SB_OK) return rc;
  rc = sb_make_tmap_2d_bf16(&tmV, v, static_cast<uint64_t>(kv_shared ? 1 : batch) * nk, 256, static_cast<uint64_t>(v_ld), AT_KT, 64);
  if (rc != SB_OK) return rc;
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attn_d256_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_once.mark();
  }
  AttnTCParams p;
  p.out = static_cast<bf16*>(o);
  p.o_ld = o_ld;
  p.nq = nq;
  p.nk = nk;
  p.q_bstride = q_shared ? 0 : nq;
  p.kv_bstride = kv_shared ? 0 : nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  attn_d256_tc_kernel<<<dim3(nq / 128, batch), AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
