// saber_b200 — Hiera multi-head attention (head_dim 72) on tcgen05 / TMEM / TMA: the 16 x 16-token windows of
// hiera-L's stage 3 (256 queries x 256 keys per window and head) and its three global blocks (4096 x 4096 per crop and
// head). Upstream: sam2/modeling/backbones/hieradet.py MultiScaleBlock.forward (window_partition -> MultiScaleAttention
// -> window_unpartition; F.scaled_dot_product_attention); reached from REF saber/adapters/sam2/automask.py:62 and
// predictor.py:24-26 through build_sam2. Replaces flash_attn_kernel<80, *, 64> (mma.sync), which ran
// the windows at 24 % tensor-pipe activity and the global blocks at 153 TFLOP/s (profiles/r01y..., r01zc...).
//
// One persistent CTA per SM walks a list of work items (crop, window, head, 128-query tile); every pipeline below keeps
// running across items, so the loads and the Q K^T of the next item overlap the softmax and epilogue of the current one.
//   operands : the fused qkv buffer [B*H*W, 3*heads*72] is addressed through 4-D tensor maps (d, head, x, y). head_dim
//              72 = 64 + 8: every operand tile is a 64-column box (SWIZZLE_128B) plus a 16-column tail box starting
//              at d = 64 (SWIZZLE_32B; the tensor map's inner extent is 72, so TMA zero-fills d = 72..79).
//   S = Q K^T: UMMA M128 N128, 4 k-steps on the main blocks + 1 on the tails; S double-buffered in TMEM
//   softmax  : 8 warps, thread = (query row, 64-key half); online with lazy rescaling of O (threshold 2^8);
//              P (bf16, unnormalised) is written back into TENSOR MEMORY over the first 64 columns of its S buffer
//              (two keys per 32-bit column) — the first version kept P in shared memory and was bound by the
//              shared-memory bandwidth (196 KB of operand traffic per 128 x 128 tile = 1 530 of its 2 190 clocks;
//              profiles/r02b_attn_probe.log)
//   O += P V : A operand from TMEM; per 16 keys one UMMA M128 N64 (V main block through an MN-major SWIZZLE_128B
//              descriptor) and one M128 N16 (V tail, MN-major SWIZZLE_32B); O (64 + 16 columns) double-buffered in
//              TMEM across items. The tensor pipe executes in issue order, so QK^T(g + 2) — issued after PV(g) —
//              overwrites the S / P buffer only after PV(g) has consumed it.
//   epilogue : four dedicated warps: O / l -> bf16 -> [128 x 144 B] staging tile -> one TMA tensor store per item.
// Warp roles (512 threads): 0 TMA producer, 1 MMA issuer (converged warp, elected lane), 2 TMEM allocator,
// 4-11 softmax, 12-15 epilogue.
#include "common.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int HA_THREADS = 512;
constexpr int HA_QM = 16384, HA_QT = 4096;          // main / tail bytes of a 128-row operand tile
constexpr int HA_TILE = HA_QM + HA_QT;              // 20480
constexpr int HA_KS = 3, HA_VS = 3;                 // K / V ring depths
constexpr int HA_OFF_Q = 0;                                  // 2 x tile
constexpr int HA_OFF_K = HA_OFF_Q + 2 * HA_TILE;             // KS x tile
constexpr int HA_OFF_V = HA_OFF_K + HA_KS * HA_TILE;         // VS x tile
constexpr int HA_OFF_OUT = HA_OFF_V + HA_VS * HA_TILE;       // 2 x [128 rows x 144 B]
constexpr int HA_OFF_BAR = HA_OFF_OUT + 2 * 18432;
constexpr int HA_OFF_XCH = HA_OFF_BAR + 512;                 // [2 parities][2 halves][128] fp32 row maxima
constexpr int HA_OFF_LSUM = HA_OFF_XCH + 2048;               // [2 O buffers][2 halves][128] fp32 row sums
constexpr int HA_SMEM = HA_OFF_LSUM + 2048;                  // 205,312 B
static_assert(HA_SMEM <= 227 * 1024, "hiera attention: shared memory plan does not fit");

struct HieraAttnParams {
  int heads, nwx, nwin;   // windows per row / per crop
  int H;                  // token rows per crop
  int wsx;                // window width in tokens (= box x extent)
  int by;                 // token rows per 128-row tile (128 / wsx)
  int wrows;              // token rows per window
  int qtiles, ktiles;     // tiles per (crop, window, head)
  int nitems;
  float scale_log2;
  long long* prof;        // PROF builds: [gridDim.x][16] accumulated clock64 deltas (see sb_hiera_attention_tc_prof)
};

__device__ __forceinline__ void pair_barrier(int q) {
  asm volatile("bar.sync %0, 64;" ::"r"(q + 2) : "memory");
}
__device__ __forceinline__ void epilogue_barrier() {  // the 4 epilogue warps
  asm volatile("bar.sync 6, 128;" ::: "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct ItemCoord {
  int head, x0, yq, yk0;
};
__device__ __forceinline__ ItemCoord item_coord(const HieraAttnParams& p, int item) {
  const int qt = item % p.qtiles;
  int g = item / p.qtiles;
  ItemCoord c;
  c.head = g % p.heads;
  g /= p.heads;
  const int win = g % p.nwin, b = g / p.nwin;
  const int wy = win / p.nwx, wx = win % p.nwx;
  c.x0 = wx * p.wsx;
  c.yk0 = b * p.H + wy * p.wrows;
  c.yq = c.yk0 + qt * p.by;
  return c;
}

#define HA_CLK(var) \
  if (PROF) var = clock64()

template <bool PROF>
__global__ void __launch_bounds__(HA_THREADS, 1)
hiera_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQt,
                     const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmKt,
                     const __grid_constant__ CUtensorMap tmVm, const __grid_constant__ CUtensorMap tmVt,
                     const __grid_constant__ CUtensorMap tmO, const HieraAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HA_OFF_BAR);
  uint64_t* q_full = bars;           // [2]
  uint64_t* q_empty = bars + 2;      // [2]
  uint64_t* k_full = bars + 4;       // [KS <= 4]
  uint64_t* k_empty = bars + 8;      // [KS]
  uint64_t* v_full = bars + 12;      // [VS <= 4]
  uint64_t* v_empty = bars + 16;     // [VS]
  uint64_t* s_full = bars + 20;      // [2]
  uint64_t* p_full = bars + 22;      // [2] 8 warp arrivals
  uint64_t* pv_done = bars + 24;     // [2]
  uint64_t* o_full = bars + 26;      // [2]
  uint64_t* o_empty = bars + 28;     // [2] 4 warp arrivals (epilogue)
  uint64_t* l_full = bars + 30;      // [2] 8 warp arrivals (softmax: row sums published)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = sb::smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();
  const int G = gridDim.x;
  const int my_items = (p.nitems - static_cast<int>(blockIdx.x) + G - 1) / G;
  const int KT = p.ktiles;

  if (warp == 0 && lane == 0) {
    sb::tma_prefetch_desc(&tmQm);
    sb::tma_prefetch_desc(&tmQt);
    sb::tma_prefetch_desc(&tmKm);
    sb::tma_prefetch_desc(&tmKt);
    sb::tma_prefetch_desc(&tmVm);
    sb::tma_prefetch_desc(&tmVt);
    sb::tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      sb::mbar_init(&q_full[i], 1);
      sb::mbar_init(&q_empty[i], 1);
      sb::mbar_init(&s_full[i], 1);
      sb::mbar_init(&p_full[i], 8);
      sb::mbar_init(&pv_done[i], 1);
      sb::mbar_init(&o_full[i], 1);
      sb::mbar_init(&o_empty[i], 4);
      sb::mbar_init(&l_full[i], 8);
    }
    for (int i = 0; i < HA_KS; ++i) {
      sb::mbar_init(&k_full[i], 1);
      sb::mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < HA_VS; ++i) {
      sb::mbar_init(&v_full[i], 1);
      sb::mbar_init(&v_empty[i], 1);
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
  // TMEM columns: S buffers [0,128) and [128,256) (P overwrites the first 64 columns of its S buffer);
  // O buffers at 256 and 384 (64 main + 16 tail columns each)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      long long t0 = 0, t1 = 0, w_k = 0, w_v = 0;
      int ks = 0, vs = 0;
      uint32_t kph = 1, vph = 1;  // "empty" parity of the current ring pass
      for (int i = 0; i < my_items; ++i) {
        const ItemCoord c = item_coord(p, static_cast<int>(blockIdx.x) + i * G);
        const int qs = i & 1;
        sb::mbar_wait(&q_empty[qs], static_cast<uint32_t>(((i >> 1) & 1) ^ 1));
        sb::mbar_arrive_expect_tx(&q_full[qs], HA_TILE);
        sb::tma_load_4d(smem + HA_OFF_Q + qs * HA_TILE, &tmQm, &q_full[qs], 0, c.head, c.x0, c.yq);
        sb::tma_load_4d(smem + HA_OFF_Q + qs * HA_TILE + HA_QM, &tmQt, &q_full[qs], 64, c.head, c.x0, c.yq);
        for (int t = 0; t < KT; ++t) {
          const int yk = c.yk0 + t * p.by;
          HA_CLK(t0);
          sb::mbar_wait(&k_empty[ks], kph);
          HA_CLK(t1);
          if (PROF) w_k += t1 - t0;
          sb::mbar_arrive_expect_tx(&k_full[ks], HA_TILE);
          sb::tma_load_4d(smem + HA_OFF_K + ks * HA_TILE, &tmKm, &k_full[ks], 0, c.head, c.x0, yk);
          sb::tma_load_4d(smem + HA_OFF_K + ks * HA_TILE + HA_QM, &tmKt, &k_full[ks], 64, c.head, c.x0, yk);
          if (++ks == HA_KS) {
            ks = 0;
            kph ^= 1u;
          }
          HA_CLK(t0);
          sb::mbar_wait(&v_empty[vs], vph);
          HA_CLK(t1);
          if (PROF) w_v += t1 - t0;
          sb::mbar_arrive_expect_tx(&v_full[vs], HA_TILE);
          sb::tma_load_4d(smem + HA_OFF_V + vs * HA_TILE, &tmVm, &v_full[vs], 0, c.head, c.x0, yk);
          sb::tma_load_4d(smem + HA_OFF_V + vs * HA_TILE + HA_QM, &tmVt, &v_full[vs], 64, c.head, c.x0, yk);
          if (++vs == HA_VS) {
            vs = 0;
            vph ^= 1u;
          }
        }
      }
      if (PROF) {
        p.prof[blockIdx.x * 16 + 0] = w_k;
        p.prof[blockIdx.x * 16 + 1] = w_v;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane per tcgen05 instruction) =====================
    constexpr uint32_t idesc_qk = sb::umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pvm = sb::umma_idesc_bf16(128, 64) | (1u << 16);  // B operand MN-major
    constexpr uint32_t idesc_pvt = sb::umma_idesc_bf16(128, 16) | (1u << 16);
    const int total = my_items * KT;
    long long t0 = 0, t1 = 0, w_q = 0, w_k = 0, w_p = 0, w_v = 0, w_o = 0;
    // QK cursor (runs two tiles ahead of the PV cursor)
    int qk_i = 0, qk_t = 0, qk_ks = 0;
    uint32_t qk_kph = 0;
    auto issue_qk = [&](int g) {
      const int s = g & 1;
      if (qk_t == 0) {
        HA_CLK(t0);
        sb::mbar_wait(&q_full[qk_i & 1], static_cast<uint32_t>((qk_i >> 1) & 1));
        HA_CLK(t1);
        if (PROF) w_q += t1 - t0;
      }
      HA_CLK(t0);
      sb::mbar_wait(&k_full[qk_ks], qk_kph);
      HA_CLK(t1);
      if (PROF) w_k += t1 - t0;
      sb::tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(s * 128);
      const uint32_t qb = sbase + HA_OFF_Q + (qk_i & 1) * HA_TILE, kb = sbase + HA_OFF_K + qk_ks * HA_TILE;
      const uint64_t da = sb::umma_desc_k_sw128(qb), db = sb::umma_desc_k_sw128(kb);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (sb::elect_one())
          sb::umma_bf16(d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc_qk,
                        static_cast<uint32_t>(k != 0));
      if (sb::elect_one()) {
        sb::umma_bf16(d, sb::umma_desc_k_sw32(qb + HA_QM), sb::umma_desc_k_sw32(kb + HA_QM), idesc_qk, 1u);
        sb::umma_commit(&s_full[s]);
        sb::umma_commit(&k_empty[qk_ks]);
        if (qk_t == KT - 1) sb::umma_commit(&q_empty[qk_i & 1]);
      }
      __syncwarp();
      if (++qk_ks == HA_KS) {
        qk_ks = 0;
        qk_kph ^= 1u;
      }
      if (++qk_t == KT) {
        qk_t = 0;
        ++qk_i;
      }
    };
    if (total > 0) issue_qk(0);
    if (total > 1) issue_qk(1);
    int i = 0, t = 0, vs = 0;
    uint32_t vph = 0;
    for (int g = 0; g < total; ++g) {
      const int s = g & 1, ob = i & 1;
      HA_CLK(t0);
      sb::mbar_wait(&p_full[s], static_cast<uint32_t>((g >> 1) & 1));
      HA_CLK(t1);
      if (PROF) w_p += t1 - t0;
      sb::mbar_wait(&v_full[vs], vph);
      HA_CLK(t0);
      if (PROF) w_v += t0 - t1;
      if (t == 0) {
        sb::mbar_wait(&o_empty[ob], static_cast<uint32_t>(((i >> 1) & 1) ^ 1));
        HA_CLK(t1);
        if (PROF) w_o += t1 - t0;
      }
      sb::tc_fence_after();
      const uint32_t dO = tmem_base + 256u + static_cast<uint32_t>(ob * 128);
      const uint32_t aP = tmem_base + static_cast<uint32_t>(s * 128);  // P: 16 keys = 8 columns per k-step
      const uint32_t vb = sbase + HA_OFF_V + vs * HA_TILE;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t dbm = sb::umma_desc_mn(vb + k * 2048, 16384, 1024, 2);
        const uint64_t dbt = sb::umma_desc_mn(vb + HA_QM + k * 512, 4096, 256, 6);
        if (sb::elect_one()) {
          sb::umma_bf16_ts(dO, aP + static_cast<uint32_t>(k * 8), dbm, idesc_pvm, static_cast<uint32_t>((t | k) != 0));
          sb::umma_bf16_ts(dO + 64u, aP + static_cast<uint32_t>(k * 8), dbt, idesc_pvt, static_cast<uint32_t>((t | k) != 0));
        }
      }
      if (sb::elect_one()) {
        sb::umma_commit(&pv_done[s]);
        sb::umma_commit(&v_empty[vs]);
        if (t == KT - 1) sb::umma_commit(&o_full[ob]);
      }
      __syncwarp();
      if (++vs == HA_VS) {
        vs = 0;
        vph ^= 1u;
      }
      if (++t == KT) {
        t = 0;
        ++i;
      }
      if (g + 2 < total) issue_qk(g + 2);
    }
    if (PROF && lane == 0) {
      p.prof[blockIdx.x * 16 + 2] = w_q;
      p.prof[blockIdx.x * 16 + 3] = w_k;
      p.prof[blockIdx.x * 16 + 4] = w_p;
      p.prof[blockIdx.x * 16 + 5] = w_v;
      p.prof[blockIdx.x * 16 + 6] = w_o;
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== online softmax: thread = (query row, 64-key half) =====================
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + HA_OFF_XCH);
    float* lsum = reinterpret_cast<float*>(smem + HA_OFF_LSUM);
    const float c = p.scale_log2;
    const float2 c2 = sb::splat2(c);
    long long t0 = 0, t1 = 0, w_s = 0, c_ld = 0, c_xch = 0, c_exp = 0, c_end = 0, c_all = 0, t_begin = 0;
    HA_CLK(t_begin);
    int g = 0;
#pragma unroll 1
    for (int i = 0; i < my_items; ++i) {
      const int ob = i & 1;
      const uint32_t tO = tmem_base + tlane + 256u + static_cast<uint32_t>(ob * 128);
      float m = -INFINITY, l = 0.f;  // m in the scaled (log2) domain
#pragma unroll 1
      for (int t = 0; t < KT; ++t, ++g) {
        const int s = g & 1;
        HA_CLK(t0);
        sb::mbar_wait(&s_full[s], static_cast<uint32_t>((g >> 1) & 1));
        HA_CLK(t1);
        if (PROF) w_s += t1 - t0;
        sb::tc_fence_after();
        uint32_t v[64];
        const uint32_t tS = tmem_base + tlane + static_cast<uint32_t>(s * 128);
        sb::tmem_ld_32x32(tS + hf * 64, v);
        sb::tmem_ld_32x32(tS + hf * 64 + 32, v + 32);
        sb::tmem_ld_wait();
        HA_CLK(t0);
        if (PROF) c_ld += t0 - t1;
        float mx4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mx4[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 4; j < 64; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c;
        xch[(s * 2 + hf) * 128 + r] = mx;
        pair_barrier(q);  // also: both halves of the row hold their S values in registers before P overwrites them
        mx = fmaxf(mx, xch[(s * 2 + (hf ^ 1)) * 128 + r]);
        if (__any_sync(0xffffffffu, mx > m + 8.f)) {
          const float mn = fmaxf(m, mx);
          const float alpha = sb::fast_exp2(m - mn);  // m = -inf on the item's first tile -> 0
          m = mn;
          l *= alpha;
          if (t > 0) {  // rescale this thread's share of O: hf 0 -> columns [0,48), hf 1 -> [48,80)
            sb::mbar_wait(&pv_done[(g - 1) & 1], static_cast<uint32_t>(((g - 1) >> 1) & 1));
            sb::tc_fence_after();
            const uint32_t to = tO + static_cast<uint32_t>(hf * 48);
            const int nch = hf == 0 ? 3 : 2;
#pragma unroll 1
            for (int cc = 0; cc < nch; ++cc) {
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
        HA_CLK(t1);
        if (PROF) c_xch += t1 - t0;
        // exponentials -> P (bf16 pairs) straight back into tensor memory: keys hf*64 + 2j, 2j+1 -> column hf*32 + j
        const float2 nm2 = sb::splat2(-m);
        float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 d = sb::fma2(make_float2(__uint_as_float(v[hh * 32 + 2 * j]), __uint_as_float(v[hh * 32 + 2 * j + 1])), c2, nm2);
            const float2 e = make_float2(sb::fast_exp2(d.x), sb::fast_exp2(d.y));
            l2[j & 1] = sb::add2(l2[j & 1], e);
            pw[j] = sb::pack_bf16x2(e.x, e.y);
          }
          sb::tmem_st_32x16(tS + hf * 32 + hh * 16, pw);
        }
        l += (l2[0].x + l2[0].y) + (l2[1].x + l2[1].y);
        sb::tmem_st_wait();
        sb::tc_fence_before();
        __syncwarp();
        if (lane == 0) sb::mbar_arrive(&p_full[s]);
        HA_CLK(t0);
        if (PROF) c_exp += t0 - t1;
      }
      // ---- publish the row sums of the item for the epilogue warps (the slot was last read by the epilogue of item i - 2)
      if (i >= 2) sb::mbar_wait(&o_empty[ob], static_cast<uint32_t>(((i >> 1) & 1) ^ 1));
      lsum[(ob * 2 + hf) * 128 + r] = l;
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&l_full[ob]);
      HA_CLK(t1);
      if (PROF) c_end += t1 - t0;
    }
    if (PROF && warp == 4 && lane == 0) {
      c_all = clock64() - t_begin;
      long long* o = p.prof + blockIdx.x * 16;
      o[7] = w_s;
      o[8] = c_ld;
      o[9] = c_xch;
      o[10] = c_end;
      o[11] = c_exp;
      o[13] = c_all;
      o[14] = static_cast<long long>(my_items) * KT;
    }
  } else if (warp >= 12) {
    // ===================== epilogue: O / l -> bf16 staging tile -> TMA store (thread = query row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    const float* lsum = reinterpret_cast<const float*>(smem + HA_OFF_LSUM);
    long long t0 = 0, t1 = 0, c_epi = 0;
#pragma unroll 1
    for (int i = 0; i < my_items; ++i) {
      const int ob = i & 1;
      const uint32_t ph = static_cast<uint32_t>((i >> 1) & 1);
      sb::mbar_wait(&l_full[ob], ph);
      const float inv = 1.f / (lsum[(ob * 2 + 0) * 128 + r] + lsum[(ob * 2 + 1) * 128 + r]);
      sb::mbar_wait(&o_full[ob], ph);
      HA_CLK(t0);
      sb::tc_fence_after();
      if (warp == 12 && lane == 0) sb::bulk_wait_read<1>();  // the store of item i - 2 has read this staging tile
      epilogue_barrier();
      const uint32_t tO = tmem_base + tlane + 256u + static_cast<uint32_t>(ob * 128);
      const uint32_t srow = sbase + HA_OFF_OUT + ob * 18432 + r * 144;
      uint32_t o[80];
#pragma unroll
      for (int cc = 0; cc < 5; ++cc) sb::tmem_ld_32x16(tO + cc * 16, o + cc * 16);
      sb::tmem_ld_wait();
      sb::tc_fence_before();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&o_empty[ob]);  // O and the row sums are in registers
#pragma unroll
      for (int j = 0; j < 9; ++j)  // columns 72..79 are padding
        sts128(srow + j * 16,
               make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                          sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv)));
      sb::fence_proxy_async();
      epilogue_barrier();
      if (warp == 12 && lane == 0) {
        const ItemCoord ic = item_coord(p, static_cast<int>(blockIdx.x) + i * G);
        sb::tma_store_4d(&tmO, smem + HA_OFF_OUT + ob * 18432, 0, ic.head, ic.x0, ic.yq);
        sb::bulk_commit();
      }
      HA_CLK(t1);
      if (PROF) c_epi += t1 - t0;
    }
    if (warp == 12 && lane == 0) {
      sb::bulk_wait<0>();
      if (PROF) p.prof[blockIdx.x * 16 + 12] = c_epi;
    }
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, 512);
  }
}

int g_ha_sms = 0;

template <bool PROF>
int ha_launch(const CUtensorMap* tm, const HieraAttnParams& p, int grid, cudaStream_t stream) {
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(hiera_attn_tc_kernel<PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, HA_SMEM));
    attr_once.mark();
  }
  hiera_attn_tc_kernel<PROF><<<grid, HA_THREADS, HA_SMEM, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

int ha_run(const void* qkv, void* out, int batch, int H, int W, int heads, int ws, float scale, long long* prof,
           cudaStream_t stream) {
  SB_REQUIRE(qkv && out && batch > 0 && H > 0 && W > 0 && heads > 0, "sb_hiera_attention_tc: bad arguments");
  const bool global = ws <= 0 || ws >= (H > W ? H : W);
  int wsx, wrows;
  if (global) {
    wsx = W;
    wrows = H;
  } else {
    wsx = ws;
    wrows = ws;
  }
  if (wsx > 128 || (128 % wsx) != 0 || (W % wsx) != 0 || (H % wrows) != 0) return SB_ERR_UNSUPPORTED;
  const int by = 128 / wsx;
  if ((wrows % by) != 0 || by > 256) return SB_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return SB_ERR_UNSUPPORTED;
  const int C = heads * 72;
  const int tiles = wrows / by;  // 128-row tiles per window
  HieraAttnParams p;
  p.heads = heads;
  p.nwx = W / wsx;
  p.nwin = p.nwx * (H / wrows);
  p.H = H;
  p.wsx = wsx;
  p.by = by;
  p.wrows = wrows;
  p.qtiles = tiles;
  p.ktiles = tiles;
  const long long nitems = static_cast<long long>(batch) * p.nwin * heads * tiles;
  SB_REQUIRE(nitems < (1ll << 30), "sb_hiera_attention_tc: too many work items");
  p.nitems = static_cast<int>(nitems);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.prof = prof;

  const bf16* base = static_cast<const bf16*>(qkv);
  const uint64_t dims[4] = {72, static_cast<uint64_t>(heads), static_cast<uint64_t>(W), static_cast<uint64_t>(batch) * H};
  const uint64_t str_in[3] = {144, static_cast<uint64_t>(3 * C) * 2, static_cast<uint64_t>(W) * 3 * C * 2};
  const uint64_t str_out[3] = {144, static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2};
  const uint32_t box_m[4] = {64, 1, static_cast<uint32_t>(wsx), static_cast<uint32_t>(by)};
  const uint32_t box_t[4] = {16, 1, static_cast<uint32_t>(wsx), static_cast<uint32_t>(by)};
  const uint32_t box_o[4] = {72, 1, static_cast<uint32_t>(wsx), static_cast<uint32_t>(by)};
  CUtensorMap tm[7];
  int rc;
  for (int o = 0; o < 3; ++o) {
    if ((rc = sb_make_tmap_nd_bf16(&tm[2 * o], base + o * C, 4, dims, str_in, box_m, 128)) != SB_OK) return rc;
    if ((rc = sb_make_tmap_nd_bf16(&tm[2 * o + 1], base + o * C, 4, dims, str_in, box_t, 32)) != SB_OK) return rc;
  }
  if ((rc = sb_make_tmap_nd_bf16(&tm[6], out, 4, dims, str_out, box_o, 0)) != SB_OK) return rc;
  if (g_ha_sms == 0) {
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    SB_CHECK_CUDA(cudaDeviceGetAttribute(&g_ha_sms, cudaDevAttrMultiProcessorCount, dev));
    if (const char* e = getenv("SB_GEMM_SMS")) {  // same SM budget as the persistent GEMMs (see gemm_tcgen05.cu)
      const int lim = atoi(e);
      if (lim >= 2 && lim < g_ha_sms) g_ha_sms = lim;
    }
  }
  const int grid = p.nitems < g_ha_sms ? p.nitems : g_ha_sms;
  if (prof) return ha_launch<true>(tm, p, grid, stream);
  return ha_launch<false>(tm, p, grid, stream);
}

}  // namespace

// Hiera attention for head_dim 72 over a fused qkv buffer [B*H*W, 3*heads*72] (q | k | v, head-major), no q-pooling:
// windows of ws x ws tokens with 128 % ws == 0 and ws*ws a multiple of 128 (ws = 16: 256-token windows), or global
// attention over the H x W grid (ws <= 0 or ws >= max(H, W); W <= 128, 128 % W == 0, H*W a multiple of 128).
// out [B*H*W, heads*72] bf16. Returns SB_ERR_UNSUPPORTED for other shapes (sb_window_attention then uses the
// mma.sync kernel).
extern "C" int sb_hiera_attention_tc(const void* qkv, void* out, int batch, int H, int W, int heads, int ws, float scale,
                                     void* stream_) {
  return ha_run(qkv, out, batch, H, W, heads, ws, scale, nullptr, reinterpret_cast<cudaStream_t>(stream_));
}

// Instrumented build of the same kernel (tools/attn_probe.py): prof [num_sms][16] int64 receives per-CTA clock64 sums —
// producer: 0 wait k_empty, 1 wait v_empty; MMA warp: 2 wait q_full, 3 wait k_full, 4 wait p_full, 5 wait v_full,
// 6 wait o_empty; softmax warp 4: 7 wait s_full, 8 TMEM load of S, 9 row max + exchange (+ rescale), 10 end-of-item
// publication of the row sums, 11 exponentials + P store to TMEM, 13 whole role, 14 tiles processed; epilogue warp 12:
// 12 O read + staging + TMA store.
extern "C" int sb_hiera_attention_tc_prof(const void* qkv, void* out, int batch, int H, int W, int heads, int ws,
                                          float scale, long long* prof, void* stream_) {
  SB_REQUIRE(prof, "sb_hiera_attention_tc_prof: prof buffer required");
  return ha_run(qkv, out, batch, H, W, heads, ws, scale, prof, reinterpret_cast<cudaStream_t>(stream_));
}
