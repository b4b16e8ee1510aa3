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
// An optional variant of the two up-scaling epilogues (UP1 / UP2) runs SIXTEEN epilogue warps (640 threads): every tile
// is drained by four sets at once, set q taking the q-th of the four d-groups of the transposed convolution (64 resp.
// 32 accumulator columns). It was built on the hypothesis that 8 warps cannot fill the issue slots; measured, it is 10 %
// slower (same instruction count, more barriers), so it is opt-in (SB_UP_WARPS=16) and kept for A/B measurements.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int CH = 16;  // epilogue column chunk
constexpr int NTHREADS = 384;   // 8 epilogue warps; the EW = 16 variants run 640 threads

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
  static constexpr int MAX_STAGES = 8;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = BN == 192 ? 256 : BN;  // TMEM column distance between the two accumulator stages
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;         // 128 / 256 / 512: powers of two
};
constexpr int SMEM_MAX = 227 * 1024;
constexpr int BAR_BYTES = 1024;           // barriers + tmem pointer, at the (1024-aligned) start of dynamic smem
constexpr int STG_BYTES = 32 * 128;       // one staging tile: 32 rows x 128 B
constexpr int NEPI_WARPS = 8;
constexpr int VEC_FLOATS = 3 * 256;        // per epilogue set: bias | gamma (or hyper) | beta of the current tile's columns
constexpr int VEC_BYTES = 2 * VEC_FLOATS * 4;

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == 1) return sb::gelu_erf(x);
  if (ACT == 2) return fmaxf(x, 0.0f);
  if (ACT == 3) return __fdividef(1.0f, 1.0f + __expf(-x));
  return x;
}
// named barrier among the 4 warps of one epilogue set (ids 1 and 2; 0 is __syncthreads)
__device__ __forceinline__ void set_barrier(int set) {
  asm volatile("bar.sync %0, 128;" ::"r"(set + 1) : "memory");
}
// Cooperative (128 threads of a set) staging of n floats src[0..n) -> dst (zero beyond `valid` / null src); n % 4 == 0.
__device__ __forceinline__ void stage_vec(float* dst, const float* src, int n, int valid, int tid128) {
  for (int c = tid128 * 4; c < n; c += 512) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src != nullptr && c < valid) v = __ldg(reinterpret_cast<const float4*>(src + c));
    *reinterpret_cast<float4*>(dst + c) = v;
  }
}

// ---- per-warp staging tiles (32 rows x RB bytes, RB = 128 or 64), XOR-swizzled in 16-byte chunks so that both the
// "one thread = one row" accesses of the TMEM epilogue and the "8 (4) lanes = one row" coalesced global accesses are
// bank-conflict free. All global traffic of the epilogue goes through them: per-thread-row LDG/STG.128 would cost
// one L1 wavefront per 16 bytes, the cooperative form moves 64-128 bytes per wavefront. -----------------------------
template <int RB>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return RB == 128 ? static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4))
                   : static_cast<uint32_t>(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// staged per-column vectors: float index -> float4 / float2 (warp-uniform address: one broadcast wavefront)
__device__ __forceinline__ float4 ldsf4(uint32_t base, int fidx) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + fidx * 4));
  return v;
}
__device__ __forceinline__ float2 ldsf2(uint32_t base, int fidx) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + fidx * 4));
  return v;
}

// Asynchronous cooperative gather of 32 rows x `valid` bytes (valid <= RB, multiple of 16) into a staging tile.
// Lane r owns row r: its source is gbase + off16 * 16 bytes (off16 < 0 = row not loaded).
template <int RB>
__device__ __forceinline__ void gather_async(uint32_t stg, const uint8_t* gbase, int off16, int valid, int lane) {
  constexpr int PPR = RB / 16;
#pragma unroll
  for (int t = 0; t < PPR; ++t) {
    const int id = t * 32 + lane;
    const int row = id / PPR, ch = id % PPR;
    const int o = __shfl_sync(0xffffffffu, off16, row);
    if (o >= 0 && ch * 16 < valid) cp_async16(stg + swz<RB>(row, ch), gbase + (static_cast<long long>(o) + ch) * 16);
  }
  cp_async_commit();
}
// Cooperative coalesced store of a staging tile: row r goes to gbase + off16 * 16 bytes (off16 < 0 = skip).
template <int RB>
__device__ __forceinline__ void scatter_store(uint32_t stg, uint8_t* gbase, int off16, int valid, int lane) {
  constexpr int PPR = RB / 16;
#pragma unroll
  for (int t = 0; t < PPR; ++t) {
    const int id = t * 32 + lane;
    const int row = id / PPR, ch = id % PPR;
    const int o = __shfl_sync(0xffffffffu, off16, row);
    if (o >= 0 && ch * 16 < valid) {
      const uint4 v = lds128(stg + swz<RB>(row, ch));
      *reinterpret_cast<uint4*>(gbase + (static_cast<long long>(o) + ch) * 16) = v;
    }
  }
}

struct EpiSmem {
  uint32_t out_stg;     // shared-space address of this warp's output staging tile
  uint32_t res_stg[2];  // residual / skip staging tiles (res_stg[1] == res_stg[0] when single-buffered)
  int nres;             // number of distinct residual staging tiles (0, 1 or 2)
  uint32_t vec;         // shared-space address of this set's staged vectors: [0,256) bias, [256,512) gamma / hyper,
                        // [512,768) beta. Read with ld.shared (a generic-pointer dereference compiles to LD.E: the
                        // generic-address path, reported by ncu as long-scoreboard stalls in every epilogue)
};

__device__ __forceinline__ void unpack_bf16x8(const uint4 r, float* f) {
  f[0] = sb::bf16_lo(r.x);
  f[1] = sb::bf16_hi(r.x);
  f[2] = sb::bf16_lo(r.y);
  f[3] = sb::bf16_hi(r.y);
  f[4] = sb::bf16_lo(r.z);
  f[5] = sb::bf16_hi(r.z);
  f[6] = sb::bf16_lo(r.w);
  f[7] = sb::bf16_hi(r.w);
}

// Residual of 16 columns (half `h` of the current 32-column group) of this thread's row: staging -> f[16] += residual
template <bool RES_F32>
__device__ __forceinline__ void add_res_from_stg(uint32_t stg, int lane, int h, float* f) {
  if (RES_F32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 r = lds128(stg + swz<128>(lane, h * 4 + j));
      f[4 * j + 0] += __uint_as_float(r.x);
      f[4 * j + 1] += __uint_as_float(r.y);
      f[4 * j + 2] += __uint_as_float(r.z);
      f[4 * j + 3] += __uint_as_float(r.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float t[8];
      unpack_bf16x8(lds128(stg + swz<64>(lane, h * 2 + j)), t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[8 * j + e] += t[e];
    }
  }
}

// this thread's 16 results (half `h` of the group) -> output staging (fp32: 128-byte rows, bf16: 64-byte rows)
__device__ __forceinline__ void stage_out(uint32_t stg, int lane, int h, int out_f32, const float* f) {
  if (out_f32) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stg + swz<128>(lane, h * 4 + j), make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                                          __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3])));
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      sts128(stg + swz<64>(lane, h * 2 + j),
             make_uint4(sb::pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), sb::pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                        sb::pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), sb::pack_bf16x2(f[8 * j + 6], f[8 * j + 7])));
  }
}

__device__ __forceinline__ void issue_res_gather(const GemmParams& p, uint32_t stg, int res_off16_row, int n0,
                                                 int ncols, int lane) {
  // res_off16_row: 16-byte offset of this lane's residual row start (or < 0); group starts n0 columns in
  if (p.res_f32)
    gather_async<128>(stg, reinterpret_cast<const uint8_t*>(p.res), res_off16_row < 0 ? -1 : res_off16_row + n0 / 4,
                      ncols * 4, lane);
  else
    gather_async<64>(stg, reinterpret_cast<const uint8_t*>(p.res), res_off16_row < 0 ? -1 : res_off16_row + n0 / 8,
                     ncols * 2, lane);
}

// ---- STD / LN epilogue of one 128 x BN tile (this warp: 32 rows). Arithmetic runs on 16-column chunks (TMEM loads
// one chunk ahead), global IO on 32-column groups through the staging tiles.
// LN: two passes over the row (statistics, then normalise); the tile spans the whole row (N <= BN).
template <int BN, bool LN, int ACT>
__device__ __forceinline__ void epilogue_rows(const GemmParams& p, const EpiSmem& es, uint32_t tmem_acc, int m_idx,
                                              int n_idx, int q, int lane) {
  constexpr int NG = BN / 32;
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const long long rrow = p.res_mod > 0 ? (row % p.res_mod) : row;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const bool has_res = p.res != nullptr;
  // 16-byte offsets of this lane's rows (all pitches are multiples of 16 bytes on this path)
  const int res_off16 = (has_res && row_ok) ? static_cast<int>(rrow * p.ldr / (p.res_f32 ? 4 : 8)) : -1;
  const int out_off16 = row_ok ? static_cast<int>(static_cast<long long>(row) * p.ldo / (p.out_f32 ? 4 : 8)) : -1;
  float sum = 0.f, sumsq = 0.f, mean = 0.f, rstd = 0.f;
#pragma unroll 1
  for (int pass = 0; pass < (LN ? 2 : 1); ++pass) {
    if (n_idx >= p.N) break;
    // LN: pass 0 computes acc + bias + residual, accumulates the row statistics and parks the sums back in TMEM
    // (tcgen05.st); pass 1 only re-reads TMEM — the residual stream is gathered once.
    const bool use_res = has_res && !(LN && pass == 1);
    if (use_res) issue_res_gather(p, es.res_stg[0], res_off16, n_idx, min(32, p.N - n_idx), lane);
    uint32_t v[2][CH];
    sb::tmem_ld_32x16(taddr, v[0]);
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      const int n0 = n_idx + g * 32;
      if (n0 >= p.N) break;  // warp-uniform
      const int ncols = min(32, p.N - n0);  // 16 or 32
      if (use_res) {
        const bool more = (n0 + 32 < p.N) && (g + 1 < NG);
        if (es.nres == 2 && more) {  // double-buffered: the next group's gather is in flight while this one is consumed
          issue_res_gather(p, es.res_stg[(g + 1) & 1], res_off16, n0 + 32, min(32, p.N - n0 - 32), lane);
          cp_async_wait_1();
        } else {
          cp_async_wait_all();
        }
        __syncwarp();
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h * CH >= ncols) break;  // warp-uniform
        const int c0 = n0 + h * CH;
        sb::tmem_ld_wait();
        if (c0 + CH < p.N && (g * 2 + h + 1) < 2 * NG)  // next chunk of this tile row
          sb::tmem_ld_32x16(taddr + static_cast<uint32_t>((g * 2 + h + 1) * CH), v[(h + 1) & 1]);
        float f[CH];
        const float2 al2 = sb::splat2(p.alpha);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = ldsf4(es.vec, (c0 - n_idx) + 4 * j);  // staged bias (LDS broadcast)
          const float2 v01 = make_float2(__uint_as_float(v[h][4 * j + 0]), __uint_as_float(v[h][4 * j + 1]));
          const float2 v23 = make_float2(__uint_as_float(v[h][4 * j + 2]), __uint_as_float(v[h][4 * j + 3]));
          float2 r01, r23;
          if (!LN) {  // act(alpha * acc + bias); packed fp32 pairs (GELU: 6 packed + 2 MUFU per pair)
            r01 = sb::fma2(v01, al2, make_float2(b.x, b.y));
            r23 = sb::fma2(v23, al2, make_float2(b.z, b.w));
            if (ACT == 1) {
              r01 = sb::gelu_erf2(r01);
              r23 = sb::gelu_erf2(r23);
            } else if (ACT != 0) {
              r01 = make_float2(apply_act<ACT>(r01.x), apply_act<ACT>(r01.y));
              r23 = make_float2(apply_act<ACT>(r23.x), apply_act<ACT>(r23.y));
            }
          } else if (pass == 0) {
            r01 = sb::add2(v01, make_float2(b.x, b.y));
            r23 = sb::add2(v23, make_float2(b.z, b.w));
          } else {
            r01 = v01;
            r23 = v23;
          }
          f[4 * j + 0] = r01.x;
          f[4 * j + 1] = r01.y;
          f[4 * j + 2] = r23.x;
          f[4 * j + 3] = r23.y;
        }
        if (use_res) {
          if (p.res_f32)
            add_res_from_stg<true>(es.res_stg[g & 1], lane, h, f);
          else
            add_res_from_stg<false>(es.res_stg[g & 1], lane, h, f);
        }
        if (LN) {
          if (pass == 0) {
            float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float2 ff = make_float2(f[2 * j], f[2 * j + 1]);
              s2 = sb::add2(s2, ff);
              q2 = sb::fma2(ff, ff, q2);
            }
            sum += s2.x + s2.y;
            sumsq += q2.x + q2.y;
            sb::tmem_st_32x16(taddr + static_cast<uint32_t>((g * 2 + h) * CH), reinterpret_cast<const uint32_t*>(f));
            continue;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 ga = ldsf4(es.vec, 256 + c0 + 4 * j);
            const float4 be = ldsf4(es.vec, 512 + c0 + 4 * j);
            const float2 nm2 = sb::splat2(-mean), rs2 = sb::splat2(rstd);
            const float2 o01 = sb::fma2(sb::add2(make_float2(f[4 * j + 0], f[4 * j + 1]), nm2),
                                        sb::mul2(rs2, make_float2(ga.x, ga.y)), make_float2(be.x, be.y));
            const float2 o23 = sb::fma2(sb::add2(make_float2(f[4 * j + 2], f[4 * j + 3]), nm2),
                                        sb::mul2(rs2, make_float2(ga.z, ga.w)), make_float2(be.z, be.w));
            f[4 * j + 0] = o01.x;
            f[4 * j + 1] = o01.y;
            f[4 * j + 2] = o23.x;
            f[4 * j + 3] = o23.y;
          }
        }
        stage_out(es.out_stg, lane, h, p.out_f32, f);
      }
      __syncwarp();  // residual tile fully consumed, output tile fully written
      if (use_res && es.nres != 2 && (n0 + 32 < p.N) && (g + 1 < NG))
        issue_res_gather(p, es.res_stg[(g + 1) & 1], res_off16, n0 + 32, min(32, p.N - n0 - 32), lane);
      if (!(LN && pass == 0)) {
        if (p.out_f32)
          scatter_store<128>(es.out_stg, reinterpret_cast<uint8_t*>(p.out), out_off16 < 0 ? -1 : out_off16 + n0 / 4,
                             ncols * 4, lane);
        else
          scatter_store<64>(es.out_stg, reinterpret_cast<uint8_t*>(p.out), out_off16 < 0 ? -1 : out_off16 + n0 / 8,
                            ncols * 2, lane);
        __syncwarp();
      }
    }
    if (LN && pass == 0) {
      sb::tmem_st_wait();
      mean = sum / static_cast<float>(p.N);
      const float var = fmaxf(sumsq / static_cast<float>(p.N) - mean * mean, 0.f);
      rstd = rsqrtf(var + p.eps);
    }
  }
}

// ---- scalar fallback for shapes the staged path cannot take (N % 16 != 0 or pitches not multiples of 16 bytes) -----
template <int BN, int ACT>
__device__ __forceinline__ void epilogue_scalar(const GemmParams& p, uint32_t tmem_acc, int m_idx, int n_idx, int q,
                                                int lane) {
  constexpr int NCH = BN / CH;
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const long long rrow = p.res_mod > 0 ? (row % p.res_mod) : row;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int n0 = n_idx + c * CH;
    if (n0 >= p.N) break;  // warp-uniform
    uint32_t v[CH];
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(c * CH), v);
    sb::tmem_ld_wait();
    const int ncols = min(CH, p.N - n0);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (j < ncols && row_ok) {
        float x = __uint_as_float(v[j]) * p.alpha + (p.bias ? __ldg(p.bias + n0 + j) : 0.f);
        x = apply_act<ACT>(x);
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

// ---- UP1 epilogue: N = 4 groups x 64 channels; row m = (b, y, x) of the gh x gw token grid ---------------
__device__ __forceinline__ void epilogue_up1(const GemmParams& p, const EpiSmem& es, uint32_t tmem_acc, int m_idx,
                                             int q, int lane, int d_begin, int d_end) {
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int x = row % p.gw, y = (row / p.gw) % p.gh;
  const long long b = row / (p.gw * p.gh);
  const uint8_t* skip = reinterpret_cast<const uint8_t*>(p.skip);
#pragma unroll 1
  for (int d = d_begin; d < d_end; ++d) {
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    const long long pix = static_cast<long long>(oy) * (2 * p.gw) + ox;
    // skip row: 64 fp32 = 256 bytes = 16 x 16 B ; output row: 64 bf16 = 128 bytes = 8 x 16 B
    const int skip_off16 = row_ok ? static_cast<int>((b * p.skip_bstride + pix * 64) / 4) : -1;
    const int out_off16 = row_ok ? static_cast<int>(((b * (2 * p.gh) + oy) * (2 * p.gw) + ox) * 8) : -1;
    float f[64];
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      gather_async<128>(es.res_stg[0], skip, skip_off16 < 0 ? -1 : skip_off16 + h * 8, 128, lane);
      uint32_t v[32];
      sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 64 + h * 32), v);
      sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 64 + h * 32 + 16), v + 16);
      cp_async_wait_all();
      __syncwarp();
      sb::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 sk = lds128(es.res_stg[0] + swz<128>(lane, j));
        const float4 bb = ldsf4(es.vec, d * 64 + h * 32 + 4 * j);
        const int e = h * 32 + 4 * j;
        const float2 f01 = sb::add2(
            sb::add2(make_float2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), make_float2(bb.x, bb.y)),
            make_float2(__uint_as_float(sk.x), __uint_as_float(sk.y)));
        const float2 f23 = sb::add2(
            sb::add2(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(bb.z, bb.w)),
            make_float2(__uint_as_float(sk.z), __uint_as_float(sk.w)));
        f[e + 0] = f01.x;
        f[e + 1] = f01.y;
        f[e + 2] = f23.x;
        f[e + 3] = f23.y;
        sum2 = sb::add2(sum2, sb::add2(f01, f23));
      }
      __syncwarp();
    }
    const float mean = (sum2.x + sum2.y) * (1.f / 64.f);
    const float2 nmean = sb::splat2(-mean);
    float2 vs2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float2 dd = sb::add2(make_float2(f[2 * j], f[2 * j + 1]), nmean);
      vs2 = sb::fma2(dd, dd, vs2);
    }
    const float rstd = rsqrtf((vs2.x + vs2.y) * (1.f / 64.f) + p.eps);
    const float2 rs2 = sb::splat2(rstd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 ga = ldsf2(es.vec, 256 + 8 * j + 2 * e);
        const float2 be = ldsf2(es.vec, 512 + 8 * j + 2 * e);
        const float2 dd = sb::add2(make_float2(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]), nmean);
        g[e] = sb::gelu_erf2(sb::fma2(dd, sb::mul2(rs2, ga), be));
      }
      sts128(es.out_stg + swz<128>(lane, j), make_uint4(sb::pack_bf16x2(g[0].x, g[0].y), sb::pack_bf16x2(g[1].x, g[1].y),
                                                         sb::pack_bf16x2(g[2].x, g[2].y), sb::pack_bf16x2(g[3].x, g[3].y)));
    }
    __syncwarp();
    scatter_store<128>(es.out_stg, reinterpret_cast<uint8_t*>(p.out), out_off16, 128, lane);
    __syncwarp();
  }
}

// ---- UP2 epilogue: N = 4 groups x 32 channels -> 4 mask logits per output pixel --------------------------
__device__ __forceinline__ void epilogue_up2(const GemmParams& p, const EpiSmem& es, uint32_t tmem_acc, int m_idx,
                                             int q, int lane, int d_begin, int d_end, int hy_off = 256) {
  const int row = m_idx + q * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int x = row % p.gw, y = (row / p.gw) % p.gh;
  const long long b = row / (p.gw * p.gh);
  const int H2 = 2 * p.gh, W2 = 2 * p.gw;
  float* masks = reinterpret_cast<float*>(p.out);
  const uint32_t hy = es.vec + hy_off * 4;  // this tile's prompt: staged hyper-network vectors
  const uint8_t* skip = reinterpret_cast<const uint8_t*>(p.skip);
  // skip row of (d): 32 fp32 = 128 bytes
  auto skip_off = [&](int d) -> int {
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    return row_ok ? static_cast<int>((b * p.skip_bstride + (static_cast<long long>(oy) * W2 + ox) * 32) / 4) : -1;
  };
  gather_async<128>(es.res_stg[d_begin & 1], skip, skip_off(d_begin), 128, lane);
  float held[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int d = d_begin; d < d_end; ++d) {
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    uint32_t v[32];
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 32), v);
    sb::tmem_ld_32x16(taddr + static_cast<uint32_t>(d * 32 + CH), v + CH);
    cp_async_wait_all();
    __syncwarp();
    if (d + 1 < d_end) gather_async<128>(es.res_stg[(d + 1) & 1], skip, skip_off(d + 1), 128, lane);
    sb::tmem_ld_wait();
    // packed fp32 pairs throughout: the epilogue is issue-bound (128 GELUs + 512 MACs per row on 8 warps per SM)
    float2 acc2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bb = ldsf4(es.vec, d * 32 + 4 * j);
      const uint4 sk = lds128(es.res_stg[d & 1] + swz<128>(lane, j));  // skip tile of d (the tile of d + 1 is the other one)
      const float2 g01 = sb::gelu_erf2(sb::add2(
          sb::add2(make_float2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), make_float2(bb.x, bb.y)),
          make_float2(__uint_as_float(sk.x), __uint_as_float(sk.y))));
      const float2 g23 = sb::gelu_erf2(sb::add2(
          sb::add2(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(bb.z, bb.w)),
          make_float2(__uint_as_float(sk.z), __uint_as_float(sk.w))));
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float4 h = ldsf4(hy, (m * 8 + j) * 4);  // warp-uniform shared-memory address: broadcast
        acc2[m] = sb::fma2(g01, make_float2(h.x, h.y), acc2[m]);
        acc2[m] = sb::fma2(g23, make_float2(h.z, h.w), acc2[m]);
      }
    }
    const float acc[4] = {acc2[0].x + acc2[0].y, acc2[1].x + acc2[1].y, acc2[2].x + acc2[2].y, acc2[3].x + acc2[3].y};
    // d = 2 dy + dx: the two dx of an output row are adjacent pixels -> one 8-byte store per mask (a warp then writes 256
    // contiguous bytes per instruction instead of two half-used 256-byte spans: the LSU data pipe was 59 % busy)
    if ((d & 1) == 0 && d + 1 < d_end) {
#pragma unroll
      for (int m = 0; m < 4; ++m) held[m] = acc[m];
    } else if ((d & 1) == 1 && d - 1 >= d_begin) {
      if (row_ok) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          *reinterpret_cast<float2*>(masks + ((b * 4 + m) * H2 + oy) * W2 + ox - 1) = make_float2(held[m], acc[m]);
      }
    } else if (row_ok) {
#pragma unroll
      for (int m = 0; m < 4; ++m) masks[((b * 4 + m) * H2 + oy) * W2 + ox] = acc[m];
    }
  }
}

template <int BN, int EPI, int ACT, int EW>
__global__ void __launch_bounds__((4 + EW) * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p, const int stages, const int res_bufs, const int staged) {
  static_assert(EW == 8 || (EW == 16 && (EPI == EPI_UP1 || EPI == EPI_UP2)), "16 epilogue warps: up-scaling epilogues only");
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * C::MAX_STAGES;
  uint64_t* tempty_bar = bars + 2 * C::MAX_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::MAX_STAGES + 4);
  uint8_t* sA = smem + BAR_BYTES;
  uint8_t* sB = sA + stages * C::A_BYTES;
  float* sVec = reinterpret_cast<float*>(sA + stages * C::STAGE_BYTES);
  uint8_t* sEpi = reinterpret_cast<uint8_t*>(sVec) + VEC_BYTES;

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
    for (int i = 0; i < stages; ++i) {
      sb::mbar_init(&full_bar[i], 1);
      sb::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      sb::mbar_init(&tfull_bar[i], 1);
      sb::mbar_init(&tempty_bar[i], EW == 16 ? 16 : 4);  // one arrive per warp that drains this stage
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
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-converged; one elected lane issues each tcgen05 instruction (inside a divergent `if (lane == 0)` every UMMA is
    // wrapped in an ELECT / BRA.U.ANY loop: ~13 instructions per UMMA, which bounds the BN = 64 / 128 tiles)
    {
      constexpr uint32_t idesc = sb::umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        sb::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        sb::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * C::ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          sb::mbar_wait(&full_bar[stage], phase);
          sb::tc_fence_after();
          const uint64_t da = sb::umma_desc_k_sw128(sb::smem_u32(sA + stage * C::A_BYTES));
          const uint64_t db = sb::umma_desc_k_sw128(sb::smem_u32(sB + stage * C::B_BYTES));
          const int ksteps = (kb == num_kb - 1) ? last_ksteps : (BK / 16);
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in 16-B units
            if (sb::elect_one())
              sb::umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                            static_cast<uint32_t>((kb | k) != 0));
          }
          if (sb::elect_one()) sb::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (sb::elect_one()) sb::umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        __syncwarp();
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
    if (EW == 16) {
      // ---- four sets drain every tile together: set `set` takes d-group `set` of the transposed convolution
      EpiSmem es;
      {
        // UP1: one 4 KB tile per warp (the output staging re-uses the skip staging: the skip values are in registers
        // before the normalised row is written); UP2: two skip tiles (this tile's and, prefetched, the next tile's are
        // not needed: one d-group per warp and tile -> a single tile)
        uint8_t* mine = sEpi + (warp - 4) * STG_BYTES;
        es.out_stg = sb::smem_u32(mine);
        es.res_stg[0] = es.res_stg[1] = sb::smem_u32(mine);
        es.nres = 1;
      }
      const int tid512 = (warp - 4) * 32 + lane;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t acc_phase = static_cast<uint32_t>((it >> 1) & 1);
        const int m_idx = (tile / n_tiles) * BM;
        float* vec = sVec + st * VEC_FLOATS;
        es.vec = sb::smem_u32(vec);
        asm volatile("bar.sync 1, 512;" ::: "memory");  // every warp is done with the previous tiles' vectors
        if (EPI == EPI_UP1) {
          for (int c = tid512 * 4; c < 256; c += 2048) *reinterpret_cast<float4*>(vec + c) = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          if (tid512 < 16) *reinterpret_cast<float4*>(vec + 256 + tid512 * 4) = __ldg(reinterpret_cast<const float4*>(p.gamma + tid512 * 4));
          else if (tid512 < 32) *reinterpret_cast<float4*>(vec + 512 + (tid512 - 16) * 4) = __ldg(reinterpret_cast<const float4*>(p.beta + (tid512 - 16) * 4));
        } else {
          if (tid512 < 32) *reinterpret_cast<float4*>(vec + tid512 * 4) = __ldg(reinterpret_cast<const float4*>(p.bias + tid512 * 4));
          else if (tid512 < 64)
            *reinterpret_cast<float4*>(vec + 256 + (tid512 - 32) * 4) =
                __ldg(reinterpret_cast<const float4*>(p.hyper + static_cast<long long>(m_idx / (p.gh * p.gw)) * 128 + (tid512 - 32) * 4));
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        sb::mbar_wait(&tfull_bar[st], acc_phase);
        sb::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(st * C::ACC_STRIDE);
        if (EPI == EPI_UP1)
          epilogue_up1(p, es, tmem_acc, m_idx, q, lane, set, set + 1);
        else
          epilogue_up2(p, es, tmem_acc, m_idx, q, lane, set, set + 1);
        sb::tc_fence_before();
        __syncwarp();
        if (lane == 0) sb::mbar_arrive(&tempty_bar[st]);
      }
    } else {
    EpiSmem es;
    {
      uint8_t* mine = sEpi + (warp - 4) * (1 + res_bufs) * STG_BYTES;
      es.out_stg = sb::smem_u32(mine);
      es.res_stg[0] = sb::smem_u32(mine + (res_bufs > 0 ? STG_BYTES : 0));
      es.res_stg[1] = sb::smem_u32(mine + (res_bufs > 1 ? 2 * STG_BYTES : (res_bufs > 0 ? STG_BYTES : 0)));
      es.nres = res_bufs;
      es.vec = sb::smem_u32(sVec + set * VEC_FLOATS);
    }
    float* myvec = sVec + set * VEC_FLOATS;
    const int tid128 = (warp & 3) * 32 + lane;
    uint32_t acc_phase = 0;
    int it = 0;
    // Per-column vectors that do not change from tile to tile are staged ONCE: the two 128-thread barriers per tile
    // that guarded the re-staging were 1.8 stalled warps per issued instruction in the up-scaling epilogues
    // (profiles/r02f_upscale_ncu_summary.txt). UP2's per-prompt hyper-network vectors are staged per warp (512 B,
    // one LDG.128 + STS.128 per lane, __syncwarp only).
    const bool const_vec = EPI == EPI_UP1 || EPI == EPI_UP2 || EPI == EPI_LN || (EPI == EPI_STD && n_tiles == 1);
    if (const_vec) {
      if (EPI == EPI_STD) {
        stage_vec(myvec, p.bias, BN, (p.N + 3) & ~3, tid128);
      } else if (EPI == EPI_LN) {
        stage_vec(myvec, p.bias, BN, p.N, tid128);
        stage_vec(myvec + 256, p.gamma, BN, p.N, tid128);
        stage_vec(myvec + 512, p.beta, BN, p.N, tid128);
      } else if (EPI == EPI_UP1) {
        stage_vec(myvec, p.bias, 256, 256, tid128);
        stage_vec(myvec + 256, p.gamma, 64, 64, tid128);
        stage_vec(myvec + 512, p.beta, 64, 64, tid128);
      } else {
        stage_vec(myvec, p.bias, 128, 128, tid128);
      }
      set_barrier(set);
    }
    const int hy_off = 256 + (warp & 3) * 128;  // UP2: this warp's private copy of the prompt's 4 x 32 hyper vector
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != set) continue;
      const int m_idx = (tile / n_tiles) * BM;
      const int n_idx = (tile % n_tiles) * BN;
      if (!const_vec) {
        // stage this tile's per-column vectors once per set: every read in the epilogues is a shared-memory broadcast
        // (per-chunk __ldg of bias / gamma / hyper was the epilogue's critical path: ~4x the main loop at K = 576)
        set_barrier(set);  // the previous tile's readers are done
        stage_vec(myvec, p.bias ? p.bias + n_idx : nullptr, BN, ((p.N - n_idx) + 3) & ~3, tid128);
        set_barrier(set);
      } else if (EPI == EPI_UP2) {
        __syncwarp();  // the previous tile's reads of this warp's hyper copy are done
        *reinterpret_cast<float4*>(myvec + hy_off + lane * 4) =
            __ldg(reinterpret_cast<const float4*>(p.hyper + static_cast<long long>(m_idx / (p.gh * p.gw)) * 128 + lane * 4));
        __syncwarp();
      }
      sb::mbar_wait(&tfull_bar[set], acc_phase);
      sb::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(set * C::ACC_STRIDE);
      if (EPI == EPI_STD) {
        if (staged)
          epilogue_rows<BN, false, ACT>(p, es, tmem_acc, m_idx, n_idx, q, lane);
        else
          epilogue_scalar<BN, ACT>(p, tmem_acc, m_idx, n_idx, q, lane);
      } else if (EPI == EPI_LN) {
        epilogue_rows<BN, true, 0>(p, es, tmem_acc, m_idx, 0, q, lane);
      } else if (EPI == EPI_UP1) {
        epilogue_up1(p, es, tmem_acc, m_idx, q, lane, 0, 4);
      } else {
        epilogue_up2(p, es, tmem_acc, m_idx, q, lane, 0, 4, hy_off);
      }
      sb::tc_fence_before();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(&tempty_bar[set]);
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

template <int BN, int EPI, int ACT = 0, int EW = 8>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms,
                cudaStream_t stream) {
  using C = Cfg<BN>;
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, ACT, EW>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    attr_once.mark();
  }
  // staged (coalesced) epilogue needs 16-byte-aligned pitches and N % 16 == 0
  int staged = 1;
  if (EPI == EPI_STD) {
    const long long ob = p.ldo * (p.out_f32 ? 4 : 2), rb = p.ldr * (p.res_f32 ? 4 : 2);
    staged = ((p.N & 15) == 0) && (ob % 16 == 0) && (!p.res || rb % 16 == 0) &&
             ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0) &&
             (p.res_mod == 0 || (p.res_mod % 32) == 0) &&
             (static_cast<long long>(p.M) * ob < (1ll << 35)) && (!p.res || static_cast<long long>(p.M) * rb < (1ll << 35));
  }
  const bool needs_res = (EPI == EPI_UP1 || EPI == EPI_UP2) || (p.res != nullptr);
  int res_bufs = EW == 16 ? 0 : (needs_res ? 2 : 0);  // 16 epilogue warps: one 4 KB staging tile per warp
  const int num_kb = (p.K + BK - 1) / BK;
  auto stages_for = [&](int rbufs) {
    const int epi = EW * (1 + rbufs) * STG_BYTES + VEC_BYTES;
    return (SMEM_MAX - 1024 - BAR_BYTES - epi) / C::STAGE_BYTES;
  };
  int stages = stages_for(res_bufs);
  if (EW == 8 && needs_res && stages < 3 && EPI != EPI_UP2) {  // BN = 256: trade the second residual buffer for a pipeline stage
    res_bufs = 1;
    stages = stages_for(res_bufs);
  }
  if (stages > C::MAX_STAGES) stages = C::MAX_STAGES;
  if (stages > num_kb + 2) stages = num_kb + 2 > 2 ? num_kb + 2 : 2;
  const int smem_bytes = 1024 + BAR_BYTES + stages * C::STAGE_BYTES + VEC_BYTES + EW * (1 + res_bufs) * STG_BYTES;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_tcgen05_kernel<BN, EPI, ACT, EW><<<grid, (4 + EW) * 32, smem_bytes, stream>>>(tmA, tmB, p, stages, res_bufs, staged);
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
      // 192 divides Hiera-L's 576 / 1152-wide MLP outputs exactly (256 wastes a quarter of the last tile). Measured
      // (profiles/r02d_encoder_probe.log vs r01zc): a win for the main-loop-bound K >= 1152 shapes (fc2: 605 -> 754
      // TFLOP/s), a loss for the epilogue-bound K <= 576 ones (qkv: 935 -> 876), so it is only a candidate for large K.
      const int cands[4] = {256, 192, 128, 64};
      for (int i = 0; i < 4; ++i) {
        const int c = cands[i];
        if (c == 192 && K < 1152) continue;
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
  SB_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, "sb_gemm_bf16: bad tile N %d", bn);
  CUtensorMap tmA, tmB;
  int rc = make_maps(A, lda, W, ldw, M, N, K, bn, &tmA, &tmB);
  if (rc != SB_OK) return rc;
#define SB_DISPATCH_ACT(BN_)                                                        \
  switch (act) {                                                                   \
    case 1: return launch_gemm<BN_, EPI_STD, 1>(tmA, tmB, p, g_num_sms, stream);   \
    case 2: return launch_gemm<BN_, EPI_STD, 2>(tmA, tmB, p, g_num_sms, stream);   \
    case 3: return launch_gemm<BN_, EPI_STD, 3>(tmA, tmB, p, g_num_sms, stream);   \
    default: return launch_gemm<BN_, EPI_STD, 0>(tmA, tmB, p, g_num_sms, stream);  \
  }
  if (bn == 256) { SB_DISPATCH_ACT(256) }
  if (bn == 192) { SB_DISPATCH_ACT(192) }
  if (bn == 128) { SB_DISPATCH_ACT(128) }
  SB_DISPATCH_ACT(64)
#undef SB_DISPATCH_ACT
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
  // SB_UP_WARPS=16 selects the 16-warp epilogue. Measured (profiles/r02e_decoder_probe*.log): 336 vs 289 us for the
  // 192-prompt batch — the epilogue is bound by the number of instructions the SM has to issue (ncu: issue slots),
  // not by per-warp latency, so more warps only add barrier and staging overhead; 8 warps stay the default.
  static int wide = -1;
  if (wide < 0) {
    const char* e = getenv("SB_UP_WARPS");
    wide = (e && atoi(e) == 16) ? 1 : 0;
  }
  if (wide) return launch_gemm<256, EPI_UP1, 0, 16>(tmA, tmB, p, g_num_sms, stream);
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
  static int wide = -1;  // see sb_gemm_upscale1: 592 vs 539 us per 192-prompt batch with 16 epilogue warps
  if (wide < 0) {
    const char* e = getenv("SB_UP_WARPS");
    wide = (e && atoi(e) == 16) ? 1 : 0;
  }
  if (wide) return launch_gemm<128, EPI_UP2, 0, 16>(tmA, tmB, p, g_num_sms, stream);
  return launch_gemm<128, EPI_UP2>(tmA, tmB, p, g_num_sms, stream);
}
