// saber_b200 — fused multi-head attention (flash-style online softmax) for the SAM2 path.
//
// One kernel covers: Hiera windowed attention with optional 2x2 max-pooled queries and upstream's
// zero-padding of non-divisible windows (padded tokens carry the QKV bias), Hiera global attention
// (a single window), and the plain batched attention of the mask decoder / memory attention.
// Q/K/V are read in place from the projection output (token-major rows, head-major columns), so no
// window-partition / head-transpose copies are materialised; the output is written token-major.
// Restates sam2/modeling/backbones/hieradet.py MultiScaleAttention + window_partition/unpartition
// and sam2/modeling/sam/transformer.py Attention (HF modeling_sam2.py:282-345).
//
// v1 math runs on mma.sync.m16n8k16 (bf16 in, fp32 accumulate); softmax statistics are fp32.
#include "common.cuh"
#include <stdlib.h>

extern "C" int sb_hiera_attention_tc(const void* qkv, void* out, int batch, int H, int W, int heads, int ws, float scale,
                                     void* stream);
int sb_internal_attention_d256_tc(const void* q, long long q_ld, const void* k, long long k_ld, const void* v,
                                  long long v_ld, void* o, long long o_ld, int batch, int nq, int nk, float scale,
                                  int q_shared, int kv_shared, cudaStream_t stream);

namespace {

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  long long q_ld, k_ld, v_ld;  // row pitch (elements)
  __nv_bfloat16* o;
  long long o_ld;
  const float* q_bias;  // per-column bias that padded tokens carry (windowed mode), may be null
  const float* k_bias;
  const float* v_bias;
  int heads, hd;
  float scale_log2;  // softmax scale * log2(e)
  int mode;          // 0 plain batched, 1 windowed
  // plain: q row = b * q_bstride + i, k/v row = b * kv_bstride + j (stride 0 = shared over batch)
  int nq, nk;
  long long q_bstride, kv_bstride;
  // windowed: input grid H x W (unpadded), window ws (in key space), q pooling stride (1 or 2),
  // output grid Ho x Wo (= floor(H/pool), floor(W/pool)), nwx/nwy windows per padded grid.
  int H, W, ws, pool, Ho, Wo, nwx, nwy;
  int qtiles;  // q tiles per (batch, window)
  // plain mode only: optional additive key term shared by all batch entries, k_eff[j] = k[b, j] + k_add[j]
  // ([nk, heads*hd] bf16). Applied as a second MMA (S = Q K^T + Q k_add^T), so the producer of k needs no
  // broadcast-residual epilogue (mask decoder: k_add = image_pe @ Wk^T + bk).
  const __nv_bfloat16* k_add;
  long long k_add_ld;
  // K / V staging depth: 2 (double-buffered tiles) or 1 when the whole key set is one tile — Hiera's 8x8 / 4x4 windows.
  // With two stages those CTAs held 14-56 KB of shared memory for 16-64 keys and only 4-16 of them fit an SM: the
  // kernels ran at 16-30 % of the HBM peak on latency alone (profiles/r02zo_win_probe.log).
  int kv_stages;
  // work items (q tile x window x batch x head, head fastest): CTAs walk them with a grid stride — Hiera's 4x4 windows are
  // 32 768 one-warp items per 8 crops, and as one CTA each they were bound by the CTA launch rate (19 us per CTA)
  long long ntasks;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb::smem_u32(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                              uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Row index (into the token-major projection buffer) of key j of window `win` in batch b, or -1 for a
// padded position.
__device__ __forceinline__ long long key_row(const AttnParams& p, int b, int win, int j) {
  if (p.mode == 0) return static_cast<long long>(b) * p.kv_bstride + j;
  const int wy = win / p.nwx, wx = win % p.nwx;
  const int y = wy * p.ws + j / p.ws, x = wx * p.ws + j % p.ws;
  if (y >= p.H || x >= p.W) return -1;
  return (static_cast<long long>(b) * p.H + y) * p.W + x;
}

// KT = keys per tile. HDP > 128 (memory attention: one head of 256) keeps the Q fragments in shared memory instead of
// registers (the fp32 output accumulator alone is 128 registers per thread) and uses 32-key tiles.
template <int HDP, int NWARPS, int KT>
__global__ void __launch_bounds__(NWARPS * 32, NWARPS == 8 ? 2 : 1)
flash_attn_kernel(const AttnParams p) {
  constexpr bool Q_IN_REGS = HDP <= 128;
  constexpr int PITCH = HDP * 2 + 16;  // bytes; odd multiple of 16 -> conflict-free ldmatrix
  constexpr int QROWS = 16 * NWARPS;
  constexpr int NT = NWARPS * 32;
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + QROWS * PITCH;
  uint8_t* sV = sK + p.kv_stages * KT * PITCH;
  uint8_t* sR = sV + p.kv_stages * KT * PITCH;  // only allocated / touched when p.k_add != nullptr
  const bool has_kadd = p.k_add != nullptr;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hd = p.hd;
  const int chunks = hd >> 3;  // 16-byte chunks per row
  // One-warp CTAs (Hiera's 4x4 windows and 16-query pooled windows) walk the work items with a grid stride, head fastest:
  // as one CTA per item they were bound by the CTA launch rate. Multi-warp CTAs keep one item per CTA (compile-time: the
  // loop form cost the 4-warp instantiation occupancy — stage 1 349 -> 419-484 us — so it is not even a runtime branch).
  constexpr bool PERSIST = NWARPS == 1;
  long long task = blockIdx.x;
  do {
  const long long nbx_all = p.ntasks / p.heads;
  const int head = PERSIST ? static_cast<int>(task % p.heads) : static_cast<int>(task / nbx_all);
  const long long bxid = PERSIST ? task / p.heads : task % nbx_all;
  const int qtile = static_cast<int>(bxid % p.qtiles);
  const int bw = static_cast<int>(bxid / p.qtiles);
  int b, win, nq, nk, wq;
  if (p.mode == 0) {
    b = bw;
    win = 0;
    nq = p.nq;
    nk = p.nk;
    wq = 0;
  } else {
    const int nwin = p.nwx * p.nwy;
    b = bw / nwin;
    win = bw % nwin;
    wq = p.ws / p.pool;
    nq = wq * wq;
    nk = p.ws * p.ws;
  }
  const int col0 = head * hd;

  // The loaders below write columns [0, hd) of every staged row (data, bias or zeros); only the pad columns
  // [hd, HDP) — never written afterwards — are zeroed here so that they contribute exact zeros to the dot products.
  {
    constexpr int PADMAX = HDP / 8;
    const int npad = PADMAX - chunks;
    if (npad > 0) {
      const int nrows = QROWS + (has_kadd ? 3 : 2) * p.kv_stages * KT;
      for (int i = tid; i < nrows * npad; i += NT)
        *reinterpret_cast<uint4*>(smem + (i / npad) * PITCH + (chunks + i % npad) * 16) = make_uint4(0, 0, 0, 0);
    }
  }

  // ---- load the Q tile (with optional 2x2 max pooling; padded tokens = bias) ----
  // Rows are split over threads so that the (division-heavy) row addressing is evaluated once per row, not once per
  // 16-byte chunk: TPR threads share a row and stride over its chunks.
  const int q0 = qtile * QROWS;
  constexpr int TPRQ = NT >= QROWS ? NT / QROWS : 1;
  const int win_y = p.mode ? win / p.nwx : 0, win_x = p.mode ? win % p.nwx : 0;
  for (int r = tid / TPRQ; r < QROWS; r += NT / TPRQ) {
    const int qi = q0 + r;
    if (qi >= nq) continue;
    const int qy = p.mode ? win_y * wq + qi / wq : 0, qx = p.mode ? win_x * wq + qi % wq : 0;  // pooled padded grid coords
    const __nv_bfloat16* qrow0 = p.q + (static_cast<long long>(b) * p.q_bstride + qi) * p.q_ld + col0;
   for (int c = tid % TPRQ; c < chunks; c += TPRQ) {
    uint4 val;
    if (p.mode == 0) {
      val = *reinterpret_cast<const uint4*>(qrow0 + c * 8);
    } else {
      __nv_bfloat162 acc[4];
      bool first = true;
      for (int dy = 0; dy < p.pool; ++dy)
        for (int dx = 0; dx < p.pool; ++dx) {
          const int y = qy * p.pool + dy, x = qx * p.pool + dx;
          uint4 t;
          if (y < p.H && x < p.W) {
            t = *reinterpret_cast<const uint4*>(
                p.q + ((static_cast<long long>(b) * p.H + y) * p.W + x) * p.q_ld + col0 + c * 8);
          } else {
            const float* bq = p.q_bias ? p.q_bias + col0 + c * 8 : nullptr;
            t.x = bq ? sb::pack_bf16x2(bq[0], bq[1]) : 0u;
            t.y = bq ? sb::pack_bf16x2(bq[2], bq[3]) : 0u;
            t.z = bq ? sb::pack_bf16x2(bq[4], bq[5]) : 0u;
            t.w = bq ? sb::pack_bf16x2(bq[6], bq[7]) : 0u;
          }
          const __nv_bfloat162* tv = reinterpret_cast<const __nv_bfloat162*>(&t);
          if (first) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = tv[e];
            first = false;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = __hmax2(acc[e], tv[e]);
          }
        }
      val = *reinterpret_cast<uint4*>(acc);
    }
    *reinterpret_cast<uint4*>(sQ + r * PITCH + c * 16) = val;
   }
  }

  // ---- K/V tile loader ----
  auto load_kv = [&](int t, int stage) {
    uint8_t* dK = sK + stage * KT * PITCH;
    uint8_t* dV = sV + stage * KT * PITCH;
    constexpr int TPRK = NT >= KT ? NT / KT : 1;
    for (int r = tid / TPRK; r < KT; r += NT / TPRK) {
      const int j = t * KT + r;
      const long long row = (j < nk) ? key_row(p, b, win, j) : -2;
     for (int c = tid % TPRK; c < chunks; c += TPRK) {
      if (row >= 0) {
        cp_async16(dK + r * PITCH + c * 16, p.k + row * p.k_ld + col0 + c * 8);
        cp_async16(dV + r * PITCH + c * 16, p.v + row * p.v_ld + col0 + c * 8);
        if (has_kadd)
          cp_async16(sR + stage * KT * PITCH + r * PITCH + c * 16, p.k_add + static_cast<long long>(j) * p.k_add_ld + col0 + c * 8);
      } else {
        uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (row == -1) {  // window padding: token is exactly the projection bias
          if (p.k_bias) {
            const float* bk = p.k_bias + col0 + c * 8;
            kk.x = sb::pack_bf16x2(bk[0], bk[1]);
            kk.y = sb::pack_bf16x2(bk[2], bk[3]);
            kk.z = sb::pack_bf16x2(bk[4], bk[5]);
            kk.w = sb::pack_bf16x2(bk[6], bk[7]);
          }
          if (p.v_bias) {
            const float* bv = p.v_bias + col0 + c * 8;
            vv.x = sb::pack_bf16x2(bv[0], bv[1]);
            vv.y = sb::pack_bf16x2(bv[2], bv[3]);
            vv.z = sb::pack_bf16x2(bv[4], bv[5]);
            vv.w = sb::pack_bf16x2(bv[6], bv[7]);
          }
        }
        *reinterpret_cast<uint4*>(dK + r * PITCH + c * 16) = kk;
        *reinterpret_cast<uint4*>(dV + r * PITCH + c * 16) = vv;
      }
     }
    }
  };

  const int ntiles = (nk + KT - 1) / KT;
  load_kv(0, 0);
  cp_async_commit();
  __syncthreads();  // Q tile visible

  // Q fragments stay in registers for the whole KV loop
  uint32_t qf[Q_IN_REGS ? HDP / 16 : 1][4];
  const uint32_t qbase = sb::smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + (lane >> 4) * 16);
  if constexpr (Q_IN_REGS) {
#pragma unroll
    for (int ks = 0; ks < HDP / 16; ++ks)
      ldsm_x4(qbase + ks * 32, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  float o[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint8_t* tK = sK + stage * KT * PITCH;
    const uint8_t* tV = sV + stage * KT * PITCH;

    // S = Q K^T  (16 x 64 per warp)
    float s[KT / 8][4];
#pragma unroll
    for (int i = 0; i < KT / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      const int id = lane >> 3;
      const uint32_t kbase =
          sb::smem_u32(tK + ((lane & 7) + (id >> 1) * 8) * PITCH + (id & 1) * 16);
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ++ks) {
        uint32_t qs[4];
        if constexpr (!Q_IN_REGS) ldsm_x4(qbase + ks * 32, qs[0], qs[1], qs[2], qs[3]);
        const uint32_t* qa = Q_IN_REGS ? qf[Q_IN_REGS ? ks : 0] : qs;
#pragma unroll
        for (int np = 0; np < KT / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(kbase + np * 16 * PITCH + ks * 32, b0, b1, b2, b3);
          mma_bf16_16816(s[2 * np], qa, b0, b1);
          mma_bf16_16816(s[2 * np + 1], qa, b2, b3);
        }
      }
      if (has_kadd) {  // S += Q k_add^T
        const uint32_t rbase = kbase + static_cast<uint32_t>(sR - sK);  // same stage / lane offsets, R region
#pragma unroll
        for (int ks = 0; ks < HDP / 16; ++ks) {
          uint32_t qs[4];
          if constexpr (!Q_IN_REGS) ldsm_x4(qbase + ks * 32, qs[0], qs[1], qs[2], qs[3]);
          const uint32_t* qa = Q_IN_REGS ? qf[Q_IN_REGS ? ks : 0] : qs;
#pragma unroll
          for (int np = 0; np < KT / 16; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(rbase + np * 16 * PITCH + ks * 32, b0, b1, b2, b3);
            mma_bf16_16816(s[2 * np], qa, b0, b1);
            mma_bf16_16816(s[2 * np + 1], qa, b2, b3);
          }
        }
      }
    }
    // mask keys beyond nk (only possible in the last tile)
    const int kbase_idx = t * KT;
    if (kbase_idx + KT > nk) {
#pragma unroll
      for (int i = 0; i < KT / 8; ++i) {
        const int kk = kbase_idx + i * 8 + (lane & 3) * 2;
        if (kk >= nk) s[i][0] = s[i][2] = -INFINITY;
        if (kk + 1 >= nk) s[i][1] = s[i][3] = -INFINITY;
      }
    }
    // online softmax (rows g and g+8 of this warp's 16)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < KT / 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = sb::fast_exp2((m0 - mn0) * p.scale_log2), c1 = sb::fast_exp2((m1 - mn1) * p.scale_log2);
    m0 = mn0;
    m1 = mn1;
    const float ms0 = mn0 * p.scale_log2, ms1 = mn1 * p.scale_log2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[KT / 16][4];
#pragma unroll
    for (int i = 0; i < KT / 8; ++i) {
      const float p0 = sb::fast_exp2(s[i][0] * p.scale_log2 - ms0);
      const float p1 = sb::fast_exp2(s[i][1] * p.scale_log2 - ms0);
      const float p2 = sb::fast_exp2(s[i][2] * p.scale_log2 - ms1);
      const float p3 = sb::fast_exp2(s[i][3] * p.scale_log2 - ms1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[i >> 1][(i & 1) * 2 + 0] = sb::pack_bf16x2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = sb::pack_bf16x2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      o[i][0] *= c0;
      o[i][1] *= c0;
      o[i][2] *= c1;
      o[i][3] *= c1;
    }
    // O += P V
    {
      const int id = lane >> 3;
      const uint32_t vbase =
          sb::smem_u32(tV + ((lane & 7) + (id & 1) * 8) * PITCH + (id >> 1) * 16);
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk) {
#pragma unroll
        for (int np = 0; np < HDP / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(vbase + kk * 16 * PITCH + np * 32, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * np], pf[kk], b0, b1);
          mma_bf16_16816(o[2 * np + 1], pf[kk], b2, b3);
        }
      }
    }
    __syncthreads();  // all warps done with this stage before it is refilled
  }

  // finalize: row sums across the quad, normalise, store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int qi = q0 + warp * 16 + g + half * 8;
    if (qi >= nq) continue;
    long long orow;
    if (p.mode == 0) {
      orow = static_cast<long long>(b) * nq + qi;
    } else {
      const int wy = win / p.nwx, wx = win % p.nwx;
      const int qy = wy * wq + qi / wq, qx = wx * wq + qi % wq;
      if (qy >= p.Ho || qx >= p.Wo) continue;  // cropped by window_unpartition
      orow = (static_cast<long long>(b) * p.Ho + qy) * p.Wo + qx;
    }
    __nv_bfloat16* dst = p.o + orow * p.o_ld + col0;
    const float inv = half ? inv1 : inv0;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      const int d = i * 8 + tq * 2;
      if (d < hd) {
        *reinterpret_cast<uint32_t*>(dst + d) =
            sb::pack_bf16x2(o[i][half * 2] * inv, o[i][half * 2 + 1] * inv);
      }
    }
  }
  if (PERSIST) __syncthreads();  // the staged tiles are free before the next item overwrites them
  } while (PERSIST && (task += gridDim.x) < p.ntasks);
}

// ------------------------------------------------------------------------------------------------
// Attention with a handful of keys (mask decoder "image attends to tokens": 4096 image queries per prompt, 7-9 token
// keys, 8 heads x 16). The flash kernel above pads the keys to a 64-wide MMA tile and spends a CTA-wide pipeline on
// 8 keys; here the op is what it is — a stream over the query matrix (32 B in, 32 B out per (row, head)) with the
// prompt's K / V held in shared memory as fp32. One thread per (query row, head); HBM-bound.
// ------------------------------------------------------------------------------------------------
constexpr int FK_MAX_KEYS = 16;
constexpr int FK_R = 2;            // query rows per thread: every K / V shared-memory load is reused FK_R times
constexpr int FK_ROWS = 32 * FK_R;  // query rows per block iteration (256 threads = 32 row groups x 8 heads)

__global__ void __launch_bounds__(256)
fewkeys_attn_kernel(const __nv_bfloat16* __restrict__ q, long long q_ld, long long q_bstride,
                    const float* __restrict__ q_add /* [nq, 128] fp32 added to every batch entry's q, or null */,
                    const __nv_bfloat16* __restrict__ k, long long k_ld, const __nv_bfloat16* __restrict__ v,
                    long long v_ld, __nv_bfloat16* __restrict__ o, long long o_ld, int nq, int nk, float scale_log2,
                    int rows_per_block) {
  constexpr int C = 128;
  // per key: 8 heads x (16 values + 4 pad floats). The 8 lanes of a quarter-warp read the 8 heads with LDS.128; a head
  // pitch of 20 floats spreads them over all 32 banks (16 would be a 4-way conflict).
  constexpr int HP = 20, KP = 8 * HP;
  __shared__ __align__(16) float sK[FK_MAX_KEYS * KP];
  __shared__ __align__(16) float sV[FK_MAX_KEYS * KP];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < nk * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    sK[j * KP + (c >> 4) * HP + (c & 15)] = __bfloat162float(k[(static_cast<long long>(b) * nk + j) * k_ld + c]);
    sV[j * KP + (c >> 4) * HP + (c & 15)] = __bfloat162float(v[(static_cast<long long>(b) * nk + j) * v_ld + c]);
  }
  __syncthreads();
  const int head = threadIdx.x & 7, rsub = threadIdx.x >> 3;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(nq, r_begin + rows_per_block);
  for (int r0 = r_begin; r0 < r_end; r0 += FK_ROWS) {
    // this thread's rows: r0 + rsub + 32 * u (a warp still touches 4 consecutive rows x 256 B per load instruction)
    float qf[FK_R][16];
#pragma unroll
    for (int u = 0; u < FK_R; ++u) {
      const int r = min(r0 + rsub + 32 * u, r_end - 1);
      const __nv_bfloat16* qp = q + (static_cast<long long>(b) * q_bstride + r) * q_ld + head * 16;
      const uint4 qa = *reinterpret_cast<const uint4*>(qp);
      const uint4 qb = *reinterpret_cast<const uint4*>(qp + 8);
      const uint32_t w8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        qf[u][2 * e] = sb::bf16_lo(w8[e]);
        qf[u][2 * e + 1] = sb::bf16_hi(w8[e]);
      }
      if (q_add != nullptr) {  // positional term of the query projection (shared by all prompts, L2-resident)
        const float4* ap = reinterpret_cast<const float4*>(q_add + static_cast<long long>(r) * C + head * 16);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 a4 = __ldg(ap + t);
          // the GEMM this replaces rounded (acc + residual) to bf16 once; keep that rounding point
          qf[u][4 * t + 0] = __bfloat162float(__float2bfloat16(qf[u][4 * t + 0] + a4.x));
          qf[u][4 * t + 1] = __bfloat162float(__float2bfloat16(qf[u][4 * t + 1] + a4.y));
          qf[u][4 * t + 2] = __bfloat162float(__float2bfloat16(qf[u][4 * t + 2] + a4.z));
          qf[u][4 * t + 3] = __bfloat162float(__float2bfloat16(qf[u][4 * t + 3] + a4.w));
        }
      }
    }
    float sc[FK_R][FK_MAX_KEYS];
    float mx[FK_R];
#pragma unroll
    for (int u = 0; u < FK_R; ++u) mx[u] = -INFINITY;
#pragma unroll
    for (int j = 0; j < FK_MAX_KEYS; ++j) {
      if (j < nk) {
        const float4* kp = reinterpret_cast<const float4*>(sK + j * KP + head * HP);
        float acc[FK_R];
#pragma unroll
        for (int u = 0; u < FK_R; ++u) acc[u] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 kk = kp[t];
#pragma unroll
          for (int u = 0; u < FK_R; ++u) {
            acc[u] = fmaf(qf[u][4 * t + 0], kk.x, acc[u]);
            acc[u] = fmaf(qf[u][4 * t + 1], kk.y, acc[u]);
            acc[u] = fmaf(qf[u][4 * t + 2], kk.z, acc[u]);
            acc[u] = fmaf(qf[u][4 * t + 3], kk.w, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < FK_R; ++u) {
          sc[u][j] = acc[u] * scale_log2;
          mx[u] = fmaxf(mx[u], sc[u][j]);
        }
      }
    }
    float l[FK_R];
    float out[FK_R][16];
#pragma unroll
    for (int u = 0; u < FK_R; ++u) {
      l[u] = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) out[u][d] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < FK_MAX_KEYS; ++j) {
      if (j < nk) {
        float pj[FK_R];
#pragma unroll
        for (int u = 0; u < FK_R; ++u) {
          pj[u] = sb::fast_exp2(sc[u][j] - mx[u]);
          l[u] += pj[u];
        }
        const float4* vp = reinterpret_cast<const float4*>(sV + j * KP + head * HP);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 vv = vp[t];
#pragma unroll
          for (int u = 0; u < FK_R; ++u) {
            out[u][4 * t + 0] = fmaf(pj[u], vv.x, out[u][4 * t + 0]);
            out[u][4 * t + 1] = fmaf(pj[u], vv.y, out[u][4 * t + 1]);
            out[u][4 * t + 2] = fmaf(pj[u], vv.z, out[u][4 * t + 2]);
            out[u][4 * t + 3] = fmaf(pj[u], vv.w, out[u][4 * t + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < FK_R; ++u) {
      const int r = r0 + rsub + 32 * u;
      if (r >= r_end) continue;
      const float inv = __fdividef(1.f, l[u]);
      __nv_bfloat16* op = o + (static_cast<long long>(b) * nq + r) * o_ld + head * 16;
      *reinterpret_cast<uint4*>(op) = make_uint4(
          sb::pack_bf16x2(out[u][0] * inv, out[u][1] * inv), sb::pack_bf16x2(out[u][2] * inv, out[u][3] * inv),
          sb::pack_bf16x2(out[u][4] * inv, out[u][5] * inv), sb::pack_bf16x2(out[u][6] * inv, out[u][7] * inv));
      *reinterpret_cast<uint4*>(op + 8) = make_uint4(
          sb::pack_bf16x2(out[u][8] * inv, out[u][9] * inv), sb::pack_bf16x2(out[u][10] * inv, out[u][11] * inv),
          sb::pack_bf16x2(out[u][12] * inv, out[u][13] * inv), sb::pack_bf16x2(out[u][14] * inv, out[u][15] * inv));
    }
  }
}

template <int HDP, int NWARPS, int KT>
int launch_attn_kt(const AttnParams& p, long long nblocks_x, cudaStream_t stream);

template <int HDP, int NWARPS>
int launch_attn(const AttnParams& p, long long nblocks_x, cudaStream_t stream) {
  // tiny key sets (Hiera 4x4 windows: 16 keys): a 16-key tile quarters the staging area, so 4x more CTAs fit per SM
  const int nk_eff = p.mode == 0 ? p.nk : p.ws * p.ws;
  if (NWARPS == 1 && HDP <= 128 && nk_eff <= 16 && p.k_add == nullptr) return launch_attn_kt<HDP, NWARPS, 16>(p, nblocks_x, stream);
  return launch_attn_kt<HDP, NWARPS, (HDP > 128 ? 32 : 64)>(p, nblocks_x, stream);
}

template <int HDP, int NWARPS, int KT>
int launch_attn_kt(const AttnParams& p_in, long long nblocks_x, cudaStream_t stream) {
  constexpr int PITCH = HDP * 2 + 16;
  constexpr int SMEM_MAX_ = (16 * NWARPS + 6 * KT) * PITCH;
  AttnParams p = p_in;
  const int nk_all = p.mode == 0 ? p.nk : p.ws * p.ws;
  p.kv_stages = nk_all <= KT ? 1 : 2;
  const int SMEM = (16 * NWARPS + (p.k_add ? 3 : 2) * p.kv_stages * KT) * PITCH;
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HDP, NWARPS, KT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX_ > 227 * 1024 ? 227 * 1024 : SMEM_MAX_));
    attr_once.mark();
  }
  SB_REQUIRE(SMEM <= 227 * 1024, "sb_attention: k_add does not fit in shared memory at head_dim %d", p.hd);
  p.ntasks = nblocks_x * p.heads;
  int sms = 0, dev = 0;
  SB_CHECK_CUDA(cudaGetDevice(&dev));
  SB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int per_sm = (220 * 1024) / (SMEM + 1024);
  if (per_sm > 2048 / (NWARPS * 32)) per_sm = 2048 / (NWARPS * 32);
  if (per_sm > 32) per_sm = 32;
  if (per_sm < 1) per_sm = 1;
  // grid-stride only for the one-warp items (launch-rate bound as single CTAs: 154 -> 135 us for the 4x4 windows); with
  // 4 / 8-warp CTAs the hardware scheduler overlaps items better than a serial loop (stage 1: 349 vs 414 us)
  long long nblk = NWARPS == 1 ? static_cast<long long>(sms) * per_sm : p.ntasks;
  if (nblk > p.ntasks) nblk = p.ntasks;
  SB_REQUIRE(nblk < (1ll << 31), "sb_attention: grid too large");
  flash_attn_kernel<HDP, NWARPS, KT><<<static_cast<unsigned>(nblk), NWARPS * 32, SMEM, stream>>>(p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

template <int NWARPS>
int dispatch_hd(const AttnParams& p, long long nbx, cudaStream_t stream) {
  const int hd = p.hd;
  if (hd <= 16) return launch_attn<16, NWARPS>(p, nbx, stream);
  if (hd <= 32) return launch_attn<32, NWARPS>(p, nbx, stream);
  if (hd <= 64) return launch_attn<64, NWARPS>(p, nbx, stream);
  if (hd <= 80) return launch_attn<80, NWARPS>(p, nbx, stream);
  if (hd <= 96) return launch_attn<96, NWARPS>(p, nbx, stream);
  if (hd <= 128) return launch_attn<128, NWARPS>(p, nbx, stream);
  if (hd <= 256) return launch_attn<256, NWARPS>(p, nbx, stream);
  sb_set_error("sb_attention: head_dim %d not supported (max 256)", hd);
  return SB_ERR_UNSUPPORTED;
}

int run_attn(AttnParams& p, int batch, int nq_per_window, int nwin, cudaStream_t stream) {
  SB_REQUIRE((p.hd % 8) == 0, "sb_attention: head_dim must be a multiple of 8 (got %d)", p.hd);
  SB_REQUIRE((p.q_ld % 8) == 0 && (p.k_ld % 8) == 0 && (p.v_ld % 8) == 0 && (p.o_ld % 2) == 0,
             "sb_attention: row pitches must be multiples of 8 elements");
  // 128-query CTAs (8 warps) for Hiera's 256-token windows and global blocks: the K / V tiles (and the loaders' per-row
  // addressing) are shared by twice as many queries as with 64-query CTAs
  static int wide = -1;
  if (wide < 0) {
    const char* e = getenv("SB_ATTN_WIDE");
    wide = (e && e[0] == '0') ? 0 : 1;
  }
  const bool use8 = wide && p.mode == 1 && p.hd > 64 && p.hd <= 80 && nq_per_window > 16 && (nq_per_window % 128) == 0;
  const int nwarps = nq_per_window <= 16 ? 1 : (use8 ? 8 : 4);  // (256-query CTAs measured slower: 38.3 vs 37.7 ms / batch)
  p.qtiles = (nq_per_window + 16 * nwarps - 1) / (16 * nwarps);
  const long long nbx = static_cast<long long>(batch) * nwin * p.qtiles;
  SB_REQUIRE(nbx > 0 && nbx < (1ll << 31), "sb_attention: grid too large");
  if (nwarps == 1) return dispatch_hd<1>(p, nbx, stream);
  if (nwarps == 8) return launch_attn<80, 8>(p, nbx, stream);
  return dispatch_hd<4>(p, nbx, stream);
}

}  // namespace

// Plain batched attention: q [B*nq, heads*hd] (pitch q_ld), k/v [B*nk, heads*hd], o [B*nq, heads*hd].
// q_shared / kv_shared: the operand has a single batch entry that every batch element reads.
extern "C" int sb_attention(const void* q, long long q_ld, const void* k, long long k_ld,
                            const void* v, long long v_ld, void* o, long long o_ld, int batch,
                            int heads, int hd, int nq, int nk, float scale, int q_shared,
                            int kv_shared, void* stream) {
  SB_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0, "sb_attention: empty problem");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.o = static_cast<__nv_bfloat16*>(o);
  p.q_ld = q_ld;
  p.k_ld = k_ld;
  p.v_ld = v_ld;
  p.o_ld = o_ld;
  p.heads = heads;
  p.hd = hd;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mode = 0;
  p.nq = nq;
  p.nk = nk;
  p.q_bstride = q_shared ? 0 : nq;
  p.kv_bstride = kv_shared ? 0 : nk;
  // few keys, 8 heads x 16, many queries: the streaming kernel (see fewkeys_attn_kernel)
  if (nk <= FK_MAX_KEYS && heads == 8 && hd == 16 && !kv_shared && nq >= 1024 && (q_ld % 8) == 0 && (o_ld % 8) == 0 &&
      ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
    const int rows_per_block = 256;
    dim3 grid((nq + rows_per_block - 1) / rows_per_block, batch);
    fewkeys_attn_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p.q, q_ld, p.q_bstride, nullptr, p.k, k_ld, p.v, v_ld, p.o, o_ld, nq, nk, p.scale_log2, rows_per_block);
    SB_CHECK_LAUNCH();
    return SB_OK;
  }
  // single head x 256 (SAM2 memory attention): tcgen05 / TMA flash kernel (attention_tc.cu); SB_ATTN_TC=0 keeps mma.sync
  if (heads == 1 && hd == 256 && (nq % 128) == 0 && nk >= 64) {
    static int use_tc = -1;
    if (use_tc < 0) {
      const char* e = getenv("SB_ATTN_TC");
      use_tc = (e && e[0] == '0') ? 0 : 1;
    }
    if (use_tc) {
      const int rc = sb_internal_attention_d256_tc(q, q_ld, k, k_ld, v, v_ld, o, o_ld, batch, nq, nk, scale, q_shared,
                                                   kv_shared, reinterpret_cast<cudaStream_t>(stream));
      if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
  }
  return run_attn(p, batch, nq, 1, reinterpret_cast<cudaStream_t>(stream));
}

// Mask-decoder image -> token attention with the query's positional term folded in: q_eff = bf16(q + q_add[row]),
// 8 heads x 16, nk <= 16 keys per prompt. q [batch*nq (or nq when q_shared), 128] bf16, q_add [nq, 128] fp32 or null.
extern "C" int sb_attention_few_keys(const void* q, long long q_ld, const float* q_add, const void* k, long long k_ld,
                                     const void* v, long long v_ld, void* o, long long o_ld, int batch, int nq, int nk,
                                     float scale, int q_shared, void* stream) {
  SB_REQUIRE(batch > 0 && nq > 0 && nk > 0 && nk <= FK_MAX_KEYS, "sb_attention_few_keys: nk must be in 1..%d", FK_MAX_KEYS);
  SB_REQUIRE((q_ld % 8) == 0 && (o_ld % 8) == 0 &&
                 ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(q_add)) & 15) == 0,
             "sb_attention_few_keys: q / o / q_add must be 16-byte aligned with pitches that are multiples of 8");
  const int rows_per_block = 256;
  dim3 grid((nq + rows_per_block - 1) / rows_per_block, batch);
  fewkeys_attn_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(q), q_ld, q_shared ? 0 : nq, q_add, static_cast<const __nv_bfloat16*>(k), k_ld,
      static_cast<const __nv_bfloat16*>(v), v_ld, static_cast<__nv_bfloat16*>(o), o_ld, nq, nk,
      scale * 1.4426950408889634f, rows_per_block);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// sb_attention with an additive key term shared by all batch entries: scores = q (k[b] + k_add)^T (k_add [nk, heads*hd]
// bf16, pitch k_add_ld). Mask decoder token -> image attention: the image positional term of the key projection.
extern "C" int sb_attention_kadd(const void* q, long long q_ld, const void* k, long long k_ld, const void* k_add,
                                 long long k_add_ld, const void* v, long long v_ld, void* o, long long o_ld, int batch,
                                 int heads, int hd, int nq, int nk, float scale, int kv_shared, void* stream) {
  SB_REQUIRE(batch > 0 && heads > 0 && nq > 0 && nk > 0 && k_add, "sb_attention_kadd: bad arguments");
  SB_REQUIRE((k_add_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(k_add) & 15) == 0, "sb_attention_kadd: k_add alignment");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.o = static_cast<__nv_bfloat16*>(o);
  p.q_ld = q_ld;
  p.k_ld = k_ld;
  p.v_ld = v_ld;
  p.o_ld = o_ld;
  p.heads = heads;
  p.hd = hd;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mode = 0;
  p.nq = nq;
  p.nk = nk;
  p.q_bstride = nq;
  p.kv_bstride = kv_shared ? 0 : nk;
  p.k_add = static_cast<const __nv_bfloat16*>(k_add);
  p.k_add_ld = k_add_ld;
  return run_attn(p, batch, nq, 1, reinterpret_cast<cudaStream_t>(stream));
}

// Hiera attention over a fused QKV buffer [B*H*W, 3*heads*hd] (columns: q | k | v, head-major).
// ws = window size in the (unpadded) H x W grid (ws >= max(H,W) means global attention), pool = 1 or
// 2 (queries 2x2 max-pooled after projection). qkv_bias (fp32, 3*heads*hd) is what padded tokens
// carry when ws does not divide H/W. Output o [B*Ho*Wo, heads*hd], Ho = H/pool, Wo = W/pool.
extern "C" int sb_window_attention(const void* qkv, const float* qkv_bias, void* o, int batch,
                                   int H, int W, int heads, int hd, int ws, int pool, float scale,
                                   void* stream) {
  SB_REQUIRE(batch > 0 && H > 0 && W > 0 && heads > 0, "sb_window_attention: empty problem");
  SB_REQUIRE(pool == 1 || pool == 2, "sb_window_attention: pool must be 1 or 2");
  SB_REQUIRE(ws > 0 && (ws % pool) == 0, "sb_window_attention: ws %% pool != 0");
  const int C = heads * hd;
  // hiera-L (head_dim 72), no q-pooling, 16 x 16 windows or global: tcgen05 / TMA kernel (hiera_attn_tc.cu);
  // SB_HIERA_TC=0 keeps the mma.sync kernel for A/B timing
  if (pool == 1 && hd == 72) {
    static int hiera_tc = -1;
    if (hiera_tc < 0) {
      const char* e = getenv("SB_HIERA_TC");
      hiera_tc = (e && e[0] == '0') ? 0 : 1;
    }
    if (hiera_tc) {
      const int rc = sb_hiera_attention_tc(qkv, o, batch, H, W, heads, ws, scale, stream);
      if (rc != SB_ERR_UNSUPPORTED) return rc;
    }
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  p.q = base;
  p.k = base + C;
  p.v = base + 2 * C;
  p.q_ld = p.k_ld = p.v_ld = 3ll * C;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.o_ld = C;
  if (qkv_bias) {
    p.q_bias = qkv_bias;
    p.k_bias = qkv_bias + C;
    p.v_bias = qkv_bias + 2 * C;
  }
  p.heads = heads;
  p.hd = hd;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mode = 1;
  p.H = H;
  p.W = W;
  p.ws = ws;
  p.pool = pool;
  p.Ho = H / pool;
  p.Wo = W / pool;
  p.nwy = (H + ws - 1) / ws;
  p.nwx = (W + ws - 1) / ws;
  const int wq = ws / pool;
  return run_attn(p, batch, wq * wq, p.nwx * p.nwy, reinterpret_cast<cudaStream_t>(stream));
}
