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
//              P (bf16, unnormalised) -> two K-major 128B-swizzled blocks in shared memory
//   O += P V : per 16 keys one UMMA M128 N64 (V main block through an MN-major SWIZZLE_128B descriptor) and one
//              M128 N16 (V tail, MN-major SWIZZLE_32B); O (64 + 16 columns) double-buffered in TMEM across items
//   epilogue : O / l -> bf16 -> [128 x 144 B] staging tile -> one TMA tensor store per item.
// Warp roles (384 threads): 0 TMA producer, 1 MMA issuer (converged warp, elected lane), 2 TMEM allocator,
// 4-11 softmax / epilogue.
#include "common.cuh"
#include <stdlib.h>

namespace {

using bf16 = __nv_bfloat16;

constexpr int HA_THREADS = 384;
constexpr int HA_QM = 16384, HA_QT = 4096;          // main / tail bytes of a 128-row operand tile
constexpr int HA_TILE = HA_QM + HA_QT;              // 20480
constexpr int HA_OFF_Q = 0;                         // 2 x tile
constexpr int HA_OFF_K = HA_OFF_Q + 2 * HA_TILE;    // 2 x tile
constexpr int HA_OFF_V = HA_OFF_K + 2 * HA_TILE;    // 2 x tile
constexpr int HA_OFF_P = HA_OFF_V + 2 * HA_TILE;    // 2 x [2 blocks x 128 rows x 128 B]   (122880)
constexpr int HA_OFF_OUT = HA_OFF_P + 2 * 32768;    // [128 rows x 144 B]                 (188416)
constexpr int HA_OFF_BAR = HA_OFF_OUT + 18432;      // (206848)
constexpr int HA_OFF_XCH = HA_OFF_BAR + 256;        // [2 parities][2 halves][128] fp32 maxima; [2][128] sums
constexpr int HA_SMEM = HA_OFF_XCH + 3072;          // 210,176 B

struct HieraAttnParams {
  int heads, nwx, nwin;   // windows per row / per crop
  int H;                  // token rows per crop
  int wsx;                // window width in tokens (= box x extent)
  int by;                 // token rows per 128-row tile (128 / wsx)
  int wrows;              // token rows per window
  int qtiles, ktiles;     // tiles per (crop, window, head)
  int nitems;
  float scale_log2;
};

__device__ __forceinline__ void pair_barrier(int q) {
  asm volatile("bar.sync %0, 64;" ::"r"(q + 2) : "memory");
}
__device__ __forceinline__ void softmax_barrier() {  // the 8 softmax warps
  asm volatile("bar.sync 1, 256;" ::: "memory");
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

__global__ void __launch_bounds__(HA_THREADS, 1)
hiera_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQm, const __grid_constant__ CUtensorMap tmQt,
                     const __grid_constant__ CUtensorMap tmKm, const __grid_constant__ CUtensorMap tmKt,
                     const __grid_constant__ CUtensorMap tmVm, const __grid_constant__ CUtensorMap tmVt,
                     const __grid_constant__ CUtensorMap tmO, const HieraAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HA_OFF_BAR);
  uint64_t* q_full = bars;           // [2]
  uint64_t* q_empty = bars + 2;      // [2]
  uint64_t* k_full = bars + 4;       // [2]
  uint64_t* k_empty = bars + 6;      // [2]
  uint64_t* v_full = bars + 8;       // [2]
  uint64_t* v_empty = bars + 10;     // [2]
  uint64_t* s_full = bars + 12;      // [2]
  uint64_t* p_full = bars + 14;      // [2] 8 warp arrivals
  uint64_t* pv_done = bars + 16;     // [2]
  uint64_t* o_full = bars + 18;      // [2]
  uint64_t* o_empty = bars + 20;     // [2] 8 warp arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 22);

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
      sb::mbar_init(&k_full[i], 1);
      sb::mbar_init(&k_empty[i], 1);
      sb::mbar_init(&v_full[i], 1);
      sb::mbar_init(&v_empty[i], 1);
      sb::mbar_init(&s_full[i], 1);
      sb::mbar_init(&p_full[i], 8);
      sb::mbar_init(&pv_done[i], 1);
      sb::mbar_init(&o_full[i], 1);
      sb::mbar_init(&o_empty[i], 8);
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
  // TMEM columns: S buffers [0,128) and [128,256); O buffers at 256 and 384 (64 main + 16 tail columns each)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0;
      for (int i = 0; i < my_items; ++i) {
        const ItemCoord c = item_coord(p, static_cast<int>(blockIdx.x) + i * G);
        const int qs = i & 1;
        sb::mbar_wait(&q_empty[qs], static_cast<uint32_t>(((i >> 1) & 1) ^ 1));
        sb::mbar_arrive_expect_tx(&q_full[qs], HA_TILE);
        sb::tma_load_4d(smem + HA_OFF_Q + qs * HA_TILE, &tmQm, &q_full[qs], 0, c.head, c.x0, c.yq);
        sb::tma_load_4d(smem + HA_OFF_Q + qs * HA_TILE + HA_QM, &tmQt, &q_full[qs], 64, c.head, c.x0, c.yq);
        for (int t = 0; t < KT; ++t, ++g) {
          const int s = g & 1;
          const uint32_t ph = static_cast<uint32_t>(((g >> 1) & 1) ^ 1);
          const int yk = c.yk0 + t * p.by;
          sb::mbar_wait(&k_empty[s], ph);
          sb::mbar_arrive_expect_tx(&k_full[s], HA_TILE);
          sb::tma_load_4d(smem + HA_OFF_K + s * HA_TILE, &tmKm, &k_full[s], 0, c.head, c.x0, yk);
          sb::tma_load_4d(smem + HA_OFF_K + s * HA_TILE + HA_QM, &tmKt, &k_full[s], 64, c.head, c.x0, yk);
          sb::mbar_wait(&v_empty[s], ph);
          sb::mbar_arrive_expect_tx(&v_full[s], HA_TILE);
          sb::tma_load_4d(smem + HA_OFF_V + s * HA_TILE, &tmVm, &v_full[s], 0, c.head, c.x0, yk);
          sb::tma_load_4d(smem + HA_OFF_V + s * HA_TILE + HA_QM, &tmVt, &v_full[s], 64, c.head, c.x0, yk);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane per tcgen05 instruction) =====================
    constexpr uint32_t idesc_qk = sb::umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_pvm = sb::umma_idesc_bf16(128, 64) | (1u << 16);  // B operand MN-major
    constexpr uint32_t idesc_pvt = sb::umma_idesc_bf16(128, 16) | (1u << 16);
    const int total = my_items * KT;
    auto issue_qk = [&](int g) {
      const int i = g / KT, t = g - i * KT;
      const int s = g & 1;
      if (t == 0) sb::mbar_wait(&q_full[i & 1], static_cast<uint32_t>((i >> 1) & 1));
      sb::mbar_wait(&k_full[s], static_cast<uint32_t>((g >> 1) & 1));
      sb::tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(s * 128);
      const uint32_t qb = sbase + HA_OFF_Q + (i & 1) * HA_TILE, kb = sbase + HA_OFF_K + s * HA_TILE;
      const uint64_t da = sb::umma_desc_k_sw128(qb), db = sb::umma_desc_k_sw128(kb);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (sb::elect_one())
          sb::umma_bf16(d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc_qk,
                        static_cast<uint32_t>(k != 0));
      if (sb::elect_one()) {
        sb::umma_bf16(d, sb::umma_desc_k_sw32(qb + HA_QM), sb::umma_desc_k_sw32(kb + HA_QM), idesc_qk, 1u);
        sb::umma_commit(&s_full[s]);
        sb::umma_commit(&k_empty[s]);
        if (t == KT - 1) sb::umma_commit(&q_empty[i & 1]);
      }
      __syncwarp();
    };
    if (total > 0) issue_qk(0);
    if (total > 1) issue_qk(1);
    for (int g = 0; g < total; ++g) {
      const int i = g / KT, t = g - i * KT;
      const int s = g & 1, ob = i & 1;
      sb::mbar_wait(&p_full[s], static_cast<uint32_t>((g >> 1) & 1));
      sb::mbar_wait(&v_full[s], static_cast<uint32_t>((g >> 1) & 1));
      if (t == 0) sb::mbar_wait(&o_empty[ob], static_cast<uint32_t>(((i >> 1) & 1) ^ 1));
      sb::tc_fence_after();
      const uint32_t dO = tmem_base + 256u + static_cast<uint32_t>(ob * 128);
      const uint32_t pb = sbase + HA_OFF_P + s * 32768, vb = sbase + HA_OFF_V + s * HA_TILE;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t da = sb::umma_desc_k_sw128(pb + (k >> 2) * 16384) + static_cast<uint64_t>(2 * (k & 3));
        const uint64_t dbm = sb::umma_desc_mn(vb + k * 2048, 16384, 1024, 2);
        const uint64_t dbt = sb::umma_desc_mn(vb + HA_QM + k * 512, 4096, 256, 6);
        if (sb::elect_one()) {
          sb::umma_bf16(dO, da, dbm, idesc_pvm, static_cast<uint32_t>((t | k) != 0));
          sb::umma_bf16(dO + 64u, da, dbt, idesc_pvt, static_cast<uint32_t>((t | k) != 0));
        }
      }
      if (sb::elect_one()) {
        sb::umma_commit(&pv_done[s]);
        sb::umma_commit(&v_empty[s]);
        if (t == KT - 1) sb::umma_commit(&o_full[ob]);
      }
      __syncwarp();
      if (g + 2 < total) issue_qk(g + 2);
    }
  } else if (warp >= 4) {
    // ===================== online softmax + epilogue: thread = (query row, 64-key half) =====================
    const int q = warp & 3;
    const int hf = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + HA_OFF_XCH);
    const float c = p.scale_log2;
    const float2 c2 = sb::splat2(c);
    int g = 0;
#pragma unroll 1
    for (int i = 0; i < my_items; ++i) {
      const int ob = i & 1;
      const uint32_t tO = tmem_base + tlane + 256u + static_cast<uint32_t>(ob * 128);
      float m = -INFINITY, l = 0.f;  // m in the scaled (log2) domain
#pragma unroll 1
      for (int t = 0; t < KT; ++t, ++g) {
        const int s = g & 1;
        sb::mbar_wait(&s_full[s], static_cast<uint32_t>((g >> 1) & 1));
        sb::tc_fence_after();
        uint32_t v[64];
        const uint32_t ta = tmem_base + tlane + static_cast<uint32_t>(s * 128 + hf * 64);
        sb::tmem_ld_32x32(ta, v);
        sb::tmem_ld_32x32(ta + 32, v + 32);
        sb::tmem_ld_wait();
        float mx4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mx4[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 4; j < 64; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * c;
        xch[(s * 2 + hf) * 128 + r] = mx;
        pair_barrier(q);
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
        if (g >= 2) sb::mbar_wait(&pv_done[s], static_cast<uint32_t>(((g >> 1) - 1) & 1));  // P(g-2) consumed
        const uint32_t prow = sbase + HA_OFF_P + s * 32768 + hf * 16384 + r * 128;
        const float2 nm2 = sb::splat2(-m);
        float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float2 e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 d = sb::fma2(make_float2(__uint_as_float(v[ch * 8 + 2 * j]), __uint_as_float(v[ch * 8 + 2 * j + 1])), c2, nm2);
            e[j] = make_float2(sb::fast_exp2(d.x), sb::fast_exp2(d.y));
            l2[j & 1] = sb::add2(l2[j & 1], e[j]);
          }
          sts128(prow + ((ch ^ (r & 7)) << 4),
                 make_uint4(sb::pack_bf16x2(e[0].x, e[0].y), sb::pack_bf16x2(e[1].x, e[1].y), sb::pack_bf16x2(e[2].x, e[2].y),
                            sb::pack_bf16x2(e[3].x, e[3].y)));
        }
        l += (l2[0].x + l2[0].y) + (l2[1].x + l2[1].y);
        sb::tc_fence_before();
        sb::fence_proxy_async();
        __syncwarp();
        if (lane == 0) sb::mbar_arrive(&p_full[s]);
      }
      // ---- epilogue of the item: O / l -> bf16 staging tile -> TMA store
      xch[512 + hf * 128 + r] = l;
      if (warp == 4 && lane == 0) sb::bulk_wait_read<0>();  // the previous item's store has read the staging tile
      softmax_barrier();
      l += xch[512 + (hf ^ 1) * 128 + r];
      const float inv = 1.f / l;
      sb::mbar_wait(&o_full[ob], static_cast<uint32_t>((i >> 1) & 1));
      sb::tc_fence_after();
      const uint32_t srow = sbase + HA_OFF_OUT + r * 144;
      if (hf == 0) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          uint32_t o[16];
          sb::tmem_ld_32x16(tO + cc * 16, o);
          sb::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            sts128(srow + (cc * 2 + j) * 16,
                   make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv)));
        }
      } else {
#pragma unroll
        for (int cc = 3; cc < 5; ++cc) {
          uint32_t o[16];
          sb::tmem_ld_32x16(tO + cc * 16, o);
          sb::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (cc * 16 + j * 8 >= 72) break;  // columns 72..79 are padding
            sts128(srow + (cc * 2 + j) * 16,
                   make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                              sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv)));
          }
        }
      }
      sb::tc_fence_before();
      sb::fence_proxy_async();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&o_empty[ob]);
      softmax_barrier();
      if (warp == 4 && lane == 0) {
        const ItemCoord ic = item_coord(p, static_cast<int>(blockIdx.x) + i * G);
        sb::tma_store_4d(&tmO, smem + HA_OFF_OUT, 0, ic.head, ic.x0, ic.yq);
        sb::bulk_commit();
      }
    }
    if (warp == 4 && lane == 0) sb::bulk_wait<0>();
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, 512);
  }
}

int g_ha_sms = 0;

}  // namespace

// Hiera attention for head_dim 72 over a fused qkv buffer [B*H*W, 3*heads*72] (q | k | v, head-major), no q-pooling,
// windows of ws x ws tokens with 128 % ws == 0 ... (ws = 16: 256-token windows; ws >= W = 64: global attention over
// the H x W grid, H*W a multiple of 128). out [B*H*W, heads*72] bf16. Returns SB_ERR_UNSUPPORTED for other shapes
// (the caller falls back to the mma.sync kernel).
extern "C" int sb_hiera_attention_tc(const void* qkv, void* out, int batch, int H, int W, int heads, int ws, float scale,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
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

  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(hiera_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HA_SMEM));
    attr_once.mark();
  }
  if (g_ha_sms == 0) {
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    SB_CHECK_CUDA(cudaDeviceGetAttribute(&g_ha_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.nitems < g_ha_sms ? p.nitems : g_ha_sms;
  hiera_attn_tc_kernel<<<grid, HA_THREADS, HA_SMEM, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
