// saber_b200 — bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA with 128-byte swizzle), persistent, warp-specialised, with fused epilogues:
//
//   STD : out[m,n] = act(alpha * sum_k A[m,k] W[n,k] + bias[n]) + residual[m (mod res_mod), n]
//   LN  : out[m,:] = LayerNorm_N(acc + bias + residual) * gamma + beta            (N <= tile width)
//   UP1 : mask-decoder output_upscaling[0..2]: ConvTranspose2d(256->64,k2,s2) as a GEMM whose epilogue does the
//         pixel shuffle, adds the high-res skip (feat_s1), LayerNorm2d(64) and GELU            (N = 4*64)
//   UP2 : output_upscaling[3..4] + hyper-network product: ConvTranspose2d(64->32,k2,s2), + feat_s0, GELU, then the
//         dot product with the 4 hyper-network vectors of the prompt -> 4 mask logits per output pixel (N = 4*32)
//
// This one kernel carries every Linear / 1x1-conv / im2col'd conv / transposed-conv of the SAM2 path (Hiera
// QKV / proj / MLP, FPN laterals, mask-decoder projections and MLPs). Replaces the cuBLASLt / cuDNN calls made by
// torch inside upstream sam2, reached from REF saber/adapters/sam2/predictor.py:24-26 and automask.py:62
// (SURVEY §8a U1/U3). The fused epilogues remove the fp32 round trips of sam2/modeling/sam/transformer.py
// (norm4 after cross_attn_image_to_token) and mask_decoder.py (output_upscaling, hyper_in @ upscaled_embedding).
//
// Roles (384 threads, 1 CTA / SM): warp0 = TMA producer, warp1 = MMA issuer (one elected lane), warp2 = TMEM
// allocator, warps 4-7 = epilogue set 0, warps 8-11 = epilogue set 1. TMEM holds two accumulator stages; set s
// drains stage s, so two tile epilogues (the bottleneck of the small-K, HBM-bound GEMMs of this model) run
// concurrently and overlap the main loop of the following tiles. Inside an epilogue the TMEM load and the residual
// loads of column chunk c+1 are issued before chunk c is processed.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int CH = 16;  // epilogue column chunk
constexpr int NTHREADS = 384;

enum { EPI_STD = 0, EPI_LN = 1, EPI_UP1 = 2, EPI_UP2 = 3 };

struct GemmParams {
  int M, N, K;
  void* out;
  long long ldo;
  int out_f32;
  const float* bias;
  int act;  // 0 none, 1 gelu(erf), 2 relu, 3 sigmoid
  const void* res;
  long long ldr;
  int res_f32;
  int res_mod;  // 0: residual row = m ; >0: residual row = m % res_mod (broadcast over batch)
  float alpha;
  // fused epilogues
  const float* gamma;
  const float* beta;
  float eps;
  const float* skip;          // UP1: feat_s1 [.., (2h)(2w), 64] fp32 ; UP2: feat_s0 [.., (2h)(2w), 32] fp32
  long long skip_bstride;     // elements between batch entries of skip (0 = shared by all prompts)
  const float* hyper;         // UP2: [B, 4, 32] fp32
  int gh, gw;                 // input token grid of the transposed conv (rows m = (b, y, x), y < gh, x < gw)
};

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return sb::gelu_erf(x);
  if (act == 2) return fmaxf(x, 0.0f);
  if (act == 3) return 1.0f / (1.0f + expf(-x));
  return x;
}

// residual chunk (16 columns) -> up to 4 x uint4 registers
__device__ __forceinline__ void load_res(const GemmParams& p, long long rrow, int n0, uint4* r) {
  if (p.res_f32) {
    const uint4* g = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = __ldg(g + j);
  } else {
    const uint4* g =
        reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow * p.ldr + n0);
    r[0] = __ldg(g);
    r[1] = __ldg(g + 1);
  }
}
__device__ __forceinline__ void add_res(const GemmParams& p, const uint4* r, float* f) {
  if (p.res_f32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[4 * j + 0] += __uint_as_float(r[j].x);
      f[4 * j + 1] += __uint_as_float(r[j].y);
      f[4 * j + 2] += __uint_as_float(r[j].z);
      f[4 * j + 3] += __uint_as_float(r[j].w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      f[8 * j + 0] += sb::bf16_lo(r[j].x);
      f[8 * j + 1] += sb::bf16_hi(r[j].x);
      f[8 * j + 2] += sb::bf16_lo(r[j].y);
      f[8 * j + 3] += sb::bf16_hi(r[j].y);
      f[8 * j + 4] += sb::bf16_lo(r[j].z);
      f[8 * j + 5] += sb::bf16_hi(r[j].z);
      f[8 * j + 6] += sb::bf16_lo(r[j].w);
      f[8 * j + 7] += sb::bf16_hi(r[j].w);
    }
  }
}
__device__ __forceinline__ void store16(void* out, int out_f32, long long off, const float* f) {
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + off);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint4 t;
      t.x = sb::pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
      t.y = sb::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
      t.z = sb::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
      t.w = sb::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
      o[j] = t;
    }
  }
}

// ---- STD epilogue of one 128 x BN tile (this warp: 32 rows) -------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_std(const GemmParams& p, uint32_t tmem_acc, int m_idx, int n_idx, int q,
                                             int lane) {
  constexpr int NCH = BN / CH;
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const long long rrow = p.res_mod > 0 ? (row % p.res_mod) : row;
  const bool vec_ok = ((p.ldo & 7) == 0) && (!p.res || (p.ldr & 7) == 0) && ((p.N & 15) == 0);
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const bool use_res = p.res != nullptr && row_ok && vec_ok;
  uint32_t v[2][CH];
  uint4 rb[2][4];
  sb::tmem_ld_32x16(taddr, v[0]);
  if (use_res && n_idx < p.N) load_res(p, rrow, n_idx, rb[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int n0 = n_idx + c * CH;
    if (n0 >= p.N) break;  // warp-uniform
    sb::tmem_ld_wait();
    if (c + 1 < NCH && n0 + CH < p.N) {
      sb::tmem_ld_32x16(taddr + static_cast<uint32_t>((c + 1) * CH), v[(c + 1) & 1]);
      if (use_res) load_res(p, rrow, n0 + CH, rb[(c + 1) & 1]);
    }
    float f[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[c & 1][j]) * p.alpha;
    if (vec_ok) {
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = __ldg(b4 + j);
          f[4 * j + 0] += t.x;
          f[4 * j + 1] += t.y;
          f[4 * j + 2] += t.z;
          f[4 * j + 3] += t.w;
        }
      }
      if (p.act) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = apply_act(f[j], p.act);
      }
      if (row_ok) {
        if (p.res) add_res(p, rb[c & 1], f);
        store16(p.out, p.out_f32, static_cast<long long>(row) * p.ldo + n0, f);
      }
    } else {
      const int ncols = min(CH, p.N - n0);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (j < ncols && row_ok) {
          float x = f[j] + (p.bias ? __ldg(p.bias + n0 + j) : 0.f);
          x = apply_act(x, p.act);
          if (p.res) {
            x += p.res_f32 ? reinterpret_cast<const float*>(p.res)[rrow * p.ldr + n0 + j]
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res)[rrow * p.ldr + n0 + j]);
          }
          if (p.out_f32)
            reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + n0 + j] = x;
          else
            reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(row) * p.ldo + n0 + j] = __float2bfloat16(x);
        }
      }
    }
  }
}

// ---- LN epilogue: the tile spans the whole row (N <= BN, N % 16 == 0) ----------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_ln(const GemmParams& p, uint32_t tmem_acc, int m_idx, int q, int lane) {
  constexpr int NCH = BN / CH;
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const long long rrow = p.res_mod > 0 ? (row % p.res_mod) : row;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const bool use_res = p.res != nullptr && row_ok;
  float sum = 0.f, sumsq = 0.f;
  float mean = 0.f, rstd = 0.f;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t v[2][CH];
    uint4 rb[2][4];
    sb::tmem_ld_32x16(taddr, v[0]);
    if (use_res) load_res(p, rrow, 0, rb[0]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int n0 = c * CH;
      if (n0 >= p.N) break;
      sb::tmem_ld_wait();
      if (c + 1 < NCH && n0 + CH < p.N) {
        sb::tmem_ld_32x16(taddr + static_cast<uint32_t>((c + 1) * CH), v[(c + 1) & 1]);
        if (use_res) load_res(p, rrow, n0 + CH, rb[(c + 1) & 1]);
      }
      float f[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[c & 1][j]);
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = __ldg(b4 + j);
          f[4 * j + 0] += t.x;
          f[4 * j + 1] += t.y;
          f[4 * j + 2] += t.z;
          f[4 * j + 3] += t.w;
        }
      }
      if (use_res) add_res(p, rb[c & 1], f);
      if (pass == 0) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          sum += f[j];
          sumsq += f[j] * f[j];
        }
      } else if (row_ok) {
        const float4* g4 = reinterpret_cast<const float4*>(p.gamma + n0);
        const float4* be4 = reinterpret_cast<const float4*>(p.beta + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 g = __ldg(g4 + j), b = __ldg(be4 + j);
          f[4 * j + 0] = (f[4 * j + 0] - mean) * rstd * g.x + b.x;
          f[4 * j + 1] = (f[4 * j + 1] - mean) * rstd * g.y + b.y;
          f[4 * j + 2] = (f[4 * j + 2] - mean) * rstd * g.z + b.z;
          f[4 * j + 3] = (f[4 * j + 3] - mean) * rstd * g.w + b.w;
        }
        store16(p.out, p.out_f32, static_cast<long long>(row) * p.ldo + n0, f);
      }
    }
    if (pass == 0) {
      mean = sum / static_cast<float>(p.N);
      const float var = fmaxf(sumsq / static_cast<float>(p.N) - mean * mean, 0.f);
      rstd = rsqrtf(var + p.eps);
    }
  }
}

// ---- UP1 epilogue: N = 4 groups x 64 channels; row m = (b, y, x) of the gh x gw token grid ---------------
__device__ __forceinline__ void epilogue_up1(const GemmParams& p, uint32_t tmem_acc, int m_idx, int q, int lane) {
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int x = row % p.gw, y = (row / p.gw) % p.gh;
  const long long b = row / (p.gw * p.gh);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll 1
  for (int d = 0; d < 4; ++d) {
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    const long long pix = static_cast<long long>(oy) * (2 * p.gw) + ox;
    const float4* s4 = reinterpret_cast<const float4*>(p.skip + (row_ok ? b * p.skip_bstride + pix * 64 : 0));
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + d * 64);
    float f[64];
    float sum = 0.f;
    uint32_t v[2][CH];
    float4 sk[2][4];
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 64), v[0]);
#pragma unroll
    for (int j = 0; j < 4; ++j) sk[0][j] = __ldg(s4 + j);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sb::tmem_ld_wait();
      if (c + 1 < 4) {
        sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 64 + (c + 1) * CH), v[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) sk[(c + 1) & 1][j] = __ldg(s4 + (c + 1) * 4 + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = __ldg(b4 + c * 4 + j);
        const float4 kk = sk[c & 1][j];
        const int e = c * CH + 4 * j;
        f[e + 0] = __uint_as_float(v[c & 1][4 * j + 0]) + bb.x + kk.x;
        f[e + 1] = __uint_as_float(v[c & 1][4 * j + 1]) + bb.y + kk.y;
        f[e + 2] = __uint_as_float(v[c & 1][4 * j + 2]) + bb.z + kk.z;
        f[e + 3] = __uint_as_float(v[c & 1][4 * j + 3]) + bb.w + kk.w;
        sum += f[e + 0] + f[e + 1] + f[e + 2] + f[e + 3];
      }
    }
    if (!row_ok) continue;
    const float mean = sum * (1.f / 64.f);
    float vs = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const float dd = f[j] - mean;
      vs += dd * dd;
    }
    const float rstd = rsqrtf(vs * (1.f / 64.f) + p.eps);
    uint4* o = reinterpret_cast<uint4*>(out + ((b * (2 * p.gh) + oy) * (2 * p.gw) + ox) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        g[e] = sb::gelu_erf((f[8 * j + e] - mean) * rstd * __ldg(p.gamma + 8 * j + e) + __ldg(p.beta + 8 * j + e));
      uint4 t;
      t.x = sb::pack_bf16x2(g[0], g[1]);
      t.y = sb::pack_bf16x2(g[2], g[3]);
      t.z = sb::pack_bf16x2(g[4], g[5]);
      t.w = sb::pack_bf16x2(g[6], g[7]);
      o[j] = t;
    }
  }
}

// ---- UP2 epilogue: N = 4 groups x 32 channels -> 4 mask logits per output pixel --------------------------
__device__ __forceinline__ void epilogue_up2(const GemmParams& p, uint32_t tmem_acc, int m_idx, int q, int lane) {
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int x = row % p.gw, y = (row / p.gw) % p.gh;
  const long long b = row / (p.gw * p.gh);
  const int H2 = 2 * p.gh, W2 = 2 * p.gw;
  float* masks = reinterpret_cast<float*>(p.out);
  const float4* hy = reinterpret_cast<const float4*>(p.hyper + b * 128);
#pragma unroll 1
  for (int d = 0; d < 4; ++d) {
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    uint32_t v[32];
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 32), v);
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 32 + CH), v + CH);
    float4 sk[8];
    if (row_ok) {
      const float4* s4 = reinterpret_cast<const float4*>(p.skip + b * p.skip_bstride +
                                                          (static_cast<long long>(oy) * W2 + ox) * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) sk[j] = __ldg(s4 + j);
    }
    sb::tmem_ld_wait();
    if (!row_ok) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + d * 32) + j);
      const float g0 = sb::gelu_erf(__uint_as_float(v[4 * j + 0]) + bb.x + sk[j].x);
      const float g1 = sb::gelu_erf(__uint_as_float(v[4 * j + 1]) + bb.y + sk[j].y);
      const float g2 = sb::gelu_erf(__uint_as_float(v[4 * j + 2]) + bb.z + sk[j].z);
      const float g3 = sb::gelu_erf(__uint_as_float(v[4 * j + 3]) + bb.w + sk[j].w);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 h = __ldg(hy + m * 8 + j);  // warp-uniform address: one broadcast transaction
        acc[m] += g0 * h.x + g1 * h.y + g2 * h.z + g3 * h.w;
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) masks[((b * 4 + m) * H2 + oy) * W2 + ox] = acc[m];
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tfull_bar = bars + 2 * C::STAGES;
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;
  const int last_ksteps = ((p.K - (num_kb - 1) * BK) + 15) / 16;

  if (warp == 0 && lane == 0) {
    sb::tma_prefetch_desc(&tmA);
    sb::tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      sb::mbar_init(&full_bar[i], 1);
      sb::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sb::mbar_init(&tfull_bar[i], 1);
      sb::mbar_init(&tempty_bar[i], 4);  // one arrive per warp of the epilogue set that drains this stage
    }
    sb::fence_barrier_init();
  }
  if (warp == 2) {
    sb::tmem_alloc(tmem_ptr, C::TMEM_COLS);
    sb::tmem_relinquish();
  }
  sb::tc_fence_before();
  __syncthreads();
  sb::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = (tile / n_tiles) * BM;
        const int n_idx = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          sb::mbar_wait(&empty_bar[stage], phase ^ 1);
          sb::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          sb::tma_load_2d(sA + stage * C::A_BYTES, &tmA, &full_bar[stage], kb * BK, m_idx);
          sb::tma_load_2d(sB + stage * C::B_BYTES, &tmB, &full_bar[stage], kb * BK, n_idx);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = sb::umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        sb::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        sb::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          sb::mbar_wait(&full_bar[stage], phase);
          sb::tc_fence_after();
          const uint64_t da = sb::umma_desc_k_sw128(sb::smem_u32(sA + stage * C::A_BYTES));
          const uint64_t db = sb::umma_desc_k_sw128(sb::smem_u32(sB + stage * C::B_BYTES));
          const int ksteps = (kb == num_kb - 1) ? last_ksteps : (BK / 16);
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in 16-B units
            sb::umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                          static_cast<uint32_t>((kb | k) != 0));
          }
          sb::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        sb::umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: set s = (warp - 4) / 4 drains accumulator stage s =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int set = (warp - 4) >> 2;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != set) continue;
      const int m_idx = (tile / n_tiles) * BM;
      const int n_idx = (tile % n_tiles) * BN;
      sb::mbar_wait(&tfull_bar[set], acc_phase);
      sb::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(set * BN);
      if (EPI == EPI_STD)
        epilogue_std<BN>(p, tmem_acc, m_idx, n_idx, q, lane);
      else if (EPI == EPI_LN)
        epilogue_ln<BN>(p, tmem_acc, m_idx, q, lane);
      else if (EPI == EPI_UP1)
        epilogue_up1(p, tmem_acc, m_idx, q, lane);
      else
        epilogue_up2(p, tmem_acc, m_idx, q, lane);
      sb::tc_fence_before();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&tempty_bar[set]);
      acc_phase ^= 1;
    }
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_tcgen05_kernel<BN, EPI><<<grid, NTHREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

int g_num_sms = 0;

int ensure_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    SB_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  return SB_OK;
}

int make_maps(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int bn,
              CUtensorMap* tmA, CUtensorMap* tmB) {
  SB_REQUIRE(M > 0 && N > 0 && K > 0, "sb_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  SB_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "sb_gemm: lda/ldw must be multiples of 8");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
             "sb_gemm: A/W must be 16-byte aligned");
  int rc = sb_make_tmap_2d_bf16(tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K),
                                static_cast<uint64_t>(lda), BM, BK);
  if (rc != SB_OK) return rc;
  return sb_make_tmap_2d_bf16(tmB, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                              static_cast<uint64_t>(ldw), static_cast<uint32_t>(bn), BK);
}

}  // namespace

extern "C" int sb_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out,
                            long long ldo, int M, int N, int K, const float* bias, int act,
                            const void* residual, long long ldr, int res_mod, int flags,
                            float alpha, int force_bn, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(act >= 0 && act <= 3, "sb_gemm_bf16: bad act %d", act);
  if (ensure_sms() != SB_OK) return SB_ERR_CUDA;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.out = out;
  p.ldo = ldo;
  p.out_f32 = (flags & 1) ? 1 : 0;
  p.bias = bias;
  p.act = act;
  p.res = residual;
  p.ldr = ldr;
  p.res_f32 = (flags & 2) ? 1 : 0;
  p.res_mod = res_mod;
  p.alpha = alpha;

  // Tile-N choice: minimise waves x tile cost (BN as proxy for per-tile time).
  int bn = force_bn;
  if (bn == 0) {
    if (N <= 64) {
      bn = 64;
    } else {
      const long long m_tiles = (M + BM - 1) / BM;
      long long best = -1;
      const int cands[3] = {256, 128, 64};
      for (int i = 0; i < 3; ++i) {
        const int c = cands[i];
        const long long tiles = m_tiles * ((N + c - 1) / c);
        const long long waves = (tiles + g_num_sms - 1) / g_num_sms;
        const long long cost = waves * (c + 24);  // +24: fixed per-tile overhead proxy
        if (best < 0 || cost < best) {
          best = cost;
          bn = c;
        }
      }
    }
  }
  SB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "sb_gemm_bf16: bad tile N %d", bn);
  CUtensorMap tmA, tmB;
  int rc = make_maps(A, lda, W, ldw, M, N, K, bn, &tmA, &tmB);
  if (rc != SB_OK) return rc;
  if (bn == 256) return launch_gemm<256, EPI_STD>(tmA, tmB, p, g_num_sms, stream);
  if (bn == 128) return launch_gemm<128, EPI_STD>(tmA, tmB, p, g_num_sms, stream);
  return launch_gemm<64, EPI_STD>(tmA, tmB, p, g_num_sms, stream);
}

// out[M,N] = LayerNorm_N(A @ W^T + bias + residual[m % res_mod or m]) * gamma + beta, N in {64,128,256} x (N%16==0).
// flags: bit0 out fp32 (else bf16), bit1 residual fp32 (else bf16).
extern "C" int sb_gemm_ln(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                          int M, int N, int K, const float* bias, const void* residual, long long ldr,
                          int res_mod, int flags, const float* gamma, const float* beta, float eps,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(N <= 256 && (N % 16) == 0, "sb_gemm_ln: N must be a multiple of 16 and <= 256 (got %d)", N);
  SB_REQUIRE((ldo % 8) == 0 && (!residual || (ldr % 8) == 0), "sb_gemm_ln: ldo/ldr must be multiples of 8");
  SB_REQUIRE(gamma && beta, "sb_gemm_ln: gamma/beta required");
  if (ensure_sms() != SB_OK) return SB_ERR_CUDA;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.out = out;
  p.ldo = ldo;
  p.out_f32 = (flags & 1) ? 1 : 0;
  p.bias = bias;
  p.res = residual;
  p.ldr = ldr;
  p.res_f32 = (flags & 2) ? 1 : 0;
  p.res_mod = res_mod;
  p.alpha = 1.f;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  CUtensorMap tmA, tmB;
  int rc = make_maps(A, lda, W, ldw, M, N, K, bn, &tmA, &tmB);
  if (rc != SB_OK) return rc;
  if (bn == 256) return launch_gemm<256, EPI_LN>(tmA, tmB, p, g_num_sms, stream);
  if (bn == 128) return launch_gemm<128, EPI_LN>(tmA, tmB, p, g_num_sms, stream);
  return launch_gemm<64, EPI_LN>(tmA, tmB, p, g_num_sms, stream);
}

// Mask-decoder output_upscaling stage 1: keys [B*gh*gw, 256] bf16 @ W1 [4*64, 256] (+bias[256]) -> pixel shuffle,
// + feat_s1 (fp32 [.., 2gh*2gw, 64], batch stride skip_bstride), LayerNorm2d(64, eps), GELU -> u1 [B*2gh*2gw, 64] bf16.
extern "C" int sb_gemm_upscale1(const void* A, long long lda, const void* W, long long ldw, int B, int gh, int gw,
                                const float* bias, const float* feat_s1, long long skip_bstride,
                                const float* gamma, const float* beta, float eps, void* u1, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && gh > 0 && gw > 0 && bias && feat_s1 && gamma && beta, "sb_gemm_upscale1: bad arguments");
  if (ensure_sms() != SB_OK) return SB_ERR_CUDA;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * gh * gw;
  p.N = 256;
  p.K = 256;
  p.out = u1;
  p.bias = bias;
  p.alpha = 1.f;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.skip = feat_s1;
  p.skip_bstride = skip_bstride;
  p.gh = gh;
  p.gw = gw;
  CUtensorMap tmA, tmB;
  int rc = make_maps(A, lda, W, ldw, p.M, 256, 256, 256, &tmA, &tmB);
  if (rc != SB_OK) return rc;
  return launch_gemm<256, EPI_UP1>(tmA, tmB, p, g_num_sms, stream);
}

// Stage 2: u1 [B*gh*gw, 64] bf16 @ W2 [4*32, 64] (+bias[128]) -> pixel shuffle, + feat_s0 (fp32 [.., 2gh*2gw, 32]), GELU,
// dot with hyper [B,4,32] -> masks [B, 4, 2gh, 2gw] fp32. gh*gw must be a multiple of 128 (a tile never spans prompts).
extern "C" int sb_gemm_upscale2(const void* A, long long lda, const void* W, long long ldw, int B, int gh, int gw,
                                const float* bias, const float* feat_s0, long long skip_bstride, const float* hyper,
                                float* masks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && gh > 0 && gw > 0 && bias && feat_s0 && hyper, "sb_gemm_upscale2: bad arguments");
  if (ensure_sms() != SB_OK) return SB_ERR_CUDA;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * gh * gw;
  p.N = 128;
  p.K = 64;
  p.out = masks;
  p.bias = bias;
  p.alpha = 1.f;
  p.skip = feat_s0;
  p.skip_bstride = skip_bstride;
  p.hyper = hyper;
  p.gh = gh;
  p.gw = gw;
  CUtensorMap tmA, tmB;
  int rc = make_maps(A, lda, W, ldw, p.M, 128, 64, 128, &tmA, &tmB);
  if (rc != SB_OK) return rc;
  return launch_gemm<128, EPI_UP2>(tmA, tmB, p, g_num_sms, stream);
}
