// saber_b200 — bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA with 128-byte swizzle), persistent, warp-specialised, with a fused epilogue:
//
//     out[m, n] = act( sum_k A[m,k] * W[n,k] + bias[n] ) + residual[m (mod res_mod), n]
//
// This one kernel carries every Linear / 1x1-conv / im2col'd conv / transposed-conv of the SAM2
// path (Hiera QKV / proj / MLP, FPN laterals, mask-decoder projections and MLPs, memory attention
// projections). Replaces the cuBLASLt calls made by torch.nn.Linear inside upstream sam2, reached
// from REF saber/adapters/sam2/predictor.py:24-26 and automask.py:62.
//
// Roles (256 threads, 1 CTA / SM):  warp0 = TMA producer, warp1 = MMA issuer (one elected lane),
// warp2 = TMEM allocator, warps4-7 = epilogue (TMEM -> registers -> global). Two accumulator
// stages in TMEM let the epilogue of tile i overlap the main loop of tile i+1.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row

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
  float alpha;  // scales the accumulator before bias (1.0 normally)
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

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
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
      sb::mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
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
            sb::umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                          idesc, static_cast<uint32_t>((kb | k) != 0));
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
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((p.ldo & 7) == 0) && (!p.res || (p.ldr & 7) == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = (tile / n_tiles) * BM;
      const int n_idx = (tile % n_tiles) * BN;
      sb::mbar_wait(&tfull_bar[acc], acc_phase);
      sb::tc_fence_after();
      const int row = m_idx + q * 32 + lane;
      const bool row_ok = row < p.M;
      const long long rrow = p.res_mod > 0 ? (row % p.res_mod) : row;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_idx + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        sb::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(acc * BN + c * 32),
                          v);
        sb::tmem_ld_wait();
        const int ncols = min(32, p.N - n0);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) f[j] += __ldg(p.bias + n0 + j);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = sb::gelu_erf(f[j]);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
        } else if (p.act == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 1.0f / (1.0f + expf(-f[j]));
        }
        if (row_ok) {
          if (ncols == 32 && vec_ok) {
            if (p.res) {
              if (p.res_f32) {
                const float4* r = reinterpret_cast<const float4*>(
                    reinterpret_cast<const float*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float4 t = __ldg(r + j);
                  f[4 * j + 0] += t.x;
                  f[4 * j + 1] += t.y;
                  f[4 * j + 2] += t.z;
                  f[4 * j + 3] += t.w;
                }
              } else {
                const uint4* r = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow * p.ldr + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 t = __ldg(r + j);
                  f[8 * j + 0] += sb::bf16_lo(t.x);
                  f[8 * j + 1] += sb::bf16_hi(t.x);
                  f[8 * j + 2] += sb::bf16_lo(t.y);
                  f[8 * j + 3] += sb::bf16_hi(t.y);
                  f[8 * j + 4] += sb::bf16_lo(t.z);
                  f[8 * j + 5] += sb::bf16_hi(t.z);
                  f[8 * j + 6] += sb::bf16_lo(t.w);
                  f[8 * j + 7] += sb::bf16_hi(t.w);
                }
              }
            }
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                                    static_cast<long long>(row) * p.ldo + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                  static_cast<long long>(row) * p.ldo + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 t;
                t.x = sb::pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
                t.y = sb::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                t.z = sb::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                t.w = sb::pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                o[j] = t;
              }
            }
          } else {
            // ragged / unaligned tail: scalar path
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < ncols) {
                float x = f[j];
                if (p.res) {
                  x += p.res_f32 ? reinterpret_cast<const float*>(p.res)[rrow * p.ldr + n0 + j]
                                 : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(
                                       p.res)[rrow * p.ldr + n0 + j]);
                }
                if (p.out_f32)
                  reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + n0 + j] = x;
                else
                  reinterpret_cast<__nv_bfloat16*>(
                      p.out)[static_cast<long long>(row) * p.ldo + n0 + j] = __float2bfloat16(x);
              }
            }
          }
        }
      }
      sb::tc_fence_before();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_tcgen05_kernel<BN><<<grid, 256, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

int g_num_sms = 0;

}  // namespace

extern "C" int sb_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out,
                            long long ldo, int M, int N, int K, const float* bias, int act,
                            const void* residual, long long ldr, int res_mod, int flags,
                            float alpha, int force_bn, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(M > 0 && N > 0 && K > 0, "sb_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  SB_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "sb_gemm_bf16: lda/ldw must be multiples of 8");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
             "sb_gemm_bf16: A/W must be 16-byte aligned");
  SB_REQUIRE(act >= 0 && act <= 3, "sb_gemm_bf16: bad act %d", act);
  if (g_num_sms == 0) {
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    SB_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  GemmParams p;
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
  int rc = sb_make_tmap_2d_bf16(&tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K),
                                static_cast<uint64_t>(lda), BM, BK);
  if (rc != SB_OK) return rc;
  rc = sb_make_tmap_2d_bf16(&tmB, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                            static_cast<uint64_t>(ldw), static_cast<uint32_t>(bn), BK);
  if (rc != SB_OK) return rc;
  if (bn == 256) return launch_gemm<256>(tmA, tmB, p, g_num_sms, stream);
  if (bn == 128) return launch_gemm<128>(tmA, tmB, p, g_num_sms, stream);
  return launch_gemm<64>(tmA, tmB, p, g_num_sms, stream);
}
