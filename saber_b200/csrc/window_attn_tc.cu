// saber_b200 — EXPERIMENTAL (opt-in with SB_WINDOW_TC=1): Hiera windowed attention for 16 x 16 windows (stage 3 of
// hiera-L: 256 tokens per window, head_dim 72, no q-pooling) on tcgen05 / TMEM. Parity-green on a B200 against the fp32
// reference (tests/test_gpu_kernels.py::test_window_attention with SB_WINDOW_TC=1), not yet timed or run end to end
// (the round's GPU budget ended there), hence not the default.
// Upstream: sam2/modeling/backbones/hieradet.py MultiScaleBlock.forward (window_partition -> MultiScaleAttention ->
// window_unpartition). See DESIGN.md section 3, "design note for (0)".
//
// CTA = (crop, window, head). head_dim 72 is not a multiple of the 64-element K-block of the 128B-swizzled UMMA
// layouts and the fused qkv rows interleave q | k | v and the heads, so the operand tiles are filled by the CTA's
// threads (16-byte cp.async into the canonical swizzled addresses, then fence.proxy.async) instead of TMA:
//   K-block 0 = dims 0-63, K-block 1 = dims 64-79 with dims 72-79 zeroed (only the first 32 bytes of its rows are used).
//   S = Q K^T : per 128-query half one UMMA group M128 N256 (4 + 1 k-steps), S in TMEM columns [0, 256)
//   softmax   : all 256 keys at once (no online rescaling); 8 warps, thread = (query row, 128-key half), two passes
//               over TMEM (maximum, then exponentials); P (bf16, unnormalised) -> four K-major K-blocks in shared memory
//   O = P V   : 16 UMMAs M128 N80 K16, V read through MN-major descriptors (64-dim atom + 16-dim atom), O in TMEM
//               columns [256, 336); 1 / l applied in the epilogue, 144-byte rows written straight from registers.
#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int WA_THREADS = 288;          // warps 0-7 workers, warp 8 MMA issuer + TMEM allocator
constexpr int WA_OFF_Q0 = 0;             // [128 x 128 B]
constexpr int WA_OFF_Q1 = 16384;         // [128 x 128 B] (32 B per row used)
constexpr int WA_OFF_K0 = 32768;         // [256 x 128 B]
constexpr int WA_OFF_K1 = 65536;
constexpr int WA_OFF_V0 = 98304;
constexpr int WA_OFF_V1 = 131072;        // = V0 + 32768: second MN atom
constexpr int WA_OFF_P = 163840;         // 4 x [128 x 128 B]
constexpr int WA_OFF_BAR = 229376;
constexpr int WA_OFF_XCH = WA_OFF_BAR + 256;  // [2][128] fp32 maxima, [2][128] fp32 sums
constexpr int WA_SMEM = WA_OFF_XCH + 2048;

struct WinAttnParams {
  const bf16* qkv;   // [B*H*W, 3*C]
  bf16* out;         // [B*H*W, C]
  int H, W, heads, nwx, nwy;
  float scale_log2;
};

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
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

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attn_tc_kernel(const WinAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WA_OFF_BAR);
  uint64_t* fill_done = bars;      // 256 arrivals: K, V and Q(half 0) are in shared memory
  uint64_t* q1_done = bars + 1;    // 128 arrivals: Q(half 1)
  uint64_t* s_full = bars + 2;     // QK^T committed (once per half)
  uint64_t* p_full = bars + 3;     // 8 warp arrivals: P written, S consumed
  uint64_t* o_full = bars + 4;     // PV committed
  uint64_t* o_empty = bars + 5;    // 8 warp arrivals: O consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);
  float* xch = reinterpret_cast<float*>(smem + WA_OFF_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int nwin = p.nwx * p.nwy;
  const int b = blockIdx.x / nwin, win = blockIdx.x % nwin;
  const int wy = win / p.nwx, wx = win % p.nwx;
  const int C = p.heads * 72;
  const uint32_t sbase = sb::smem_u32(smem);
  if ((sbase & 1023u) != 0u) __trap();

  if (warp == 8) {
    if (lane == 0) {
      sb::mbar_init(fill_done, 256);
      sb::mbar_init(q1_done, 128);
      sb::mbar_init(s_full, 1);
      sb::mbar_init(p_full, 8);
      sb::mbar_init(o_full, 1);
      sb::mbar_init(o_empty, 8);
      sb::fence_barrier_init();
    }
    __syncwarp();
    sb::tmem_alloc(tmem_ptr, 512);
    sb::tmem_relinquish();
  }
  sb::tc_fence_before();
  __syncthreads();
  sb::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // token (row of qkv / out) of window-local index j in [0, 256)
  auto token = [&](int j) -> long long {
    return (static_cast<long long>(b) * p.H + wy * 16 + (j >> 4)) * p.W + wx * 16 + (j & 15);
  };
  // one 72-dim row -> K-block 0 (8 chunks) + K-block 1 (chunk 0 = dims 64-71, chunk 1 = zeros)
  auto fill_row = [&](uint32_t blk0, uint32_t blk1, int row, const bf16* src) {
    const uint32_t r0 = blk0 + row * 128, r1 = blk1 + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 8; ++c) cp_async16(r0 + ((c ^ sw) << 4), src + c * 8);
    cp_async16(r1 + ((0 ^ sw) << 4), src + 64);
    sts128(r1 + ((1 ^ sw) << 4), make_uint4(0u, 0u, 0u, 0u));
  };

  if (warp == 8) {
    // ===================== MMA issuer (converged warp, elected lane) =====================
    constexpr uint32_t idesc_qk = sb::umma_idesc_bf16(128, 256);
    constexpr uint32_t idesc_pv = sb::umma_idesc_bf16(128, 80) | (1u << 16);  // B operand MN-major
    for (int half = 0; half < 2; ++half) {
      if (half == 0)
        sb::mbar_wait(fill_done, 0);
      else
        sb::mbar_wait(q1_done, 0);
      sb::tc_fence_after();
      // S = Q K^T
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (sb::elect_one())
          sb::umma_bf16(tmem_base, sb::umma_desc_k_sw128(sbase + WA_OFF_Q0) + static_cast<uint64_t>(2 * k),
                        sb::umma_desc_k_sw128(sbase + WA_OFF_K0) + static_cast<uint64_t>(2 * k), idesc_qk,
                        static_cast<uint32_t>(k != 0));
      if (sb::elect_one()) {
        sb::umma_bf16(tmem_base, sb::umma_desc_k_sw128(sbase + WA_OFF_Q1), sb::umma_desc_k_sw128(sbase + WA_OFF_K1), idesc_qk, 1u);
        sb::umma_commit(s_full);
      }
      __syncwarp();
      // O = P V
      sb::mbar_wait(p_full, static_cast<uint32_t>(half));
      if (half == 1) sb::mbar_wait(o_empty, 0);
      sb::tc_fence_after();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t da = sb::umma_desc_k_sw128(sbase + WA_OFF_P + (k >> 2) * 16384) + static_cast<uint64_t>(2 * (k & 3));
        const uint64_t db = umma_desc_mn_sw128(sbase + WA_OFF_V0 + k * 2048, 32768, 1024);
        if (sb::elect_one()) sb::umma_bf16(tmem_base + 256u, da, db, idesc_pv, static_cast<uint32_t>(k != 0));
      }
      if (sb::elect_one()) sb::umma_commit(o_full);
      __syncwarp();
    }
  } else {
    // ===================== workers: fills, softmax, epilogue =====================
    const int t = threadIdx.x;  // 0..255
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    {
      const bf16* row = p.qkv + token(t) * (3ll * C) + head * 72;
      fill_row(sbase + WA_OFF_K0, sbase + WA_OFF_K1, t, row + C);
      fill_row(sbase + WA_OFF_V0, sbase + WA_OFF_V1, t, row + 2 * C);
      if (t < 128) fill_row(sbase + WA_OFF_Q0, sbase + WA_OFF_Q1, t, row);
      cp_async_wait_all();
      sb::fence_proxy_async();
      sb::mbar_arrive(fill_done);
    }
    const float c = p.scale_log2;
    for (int half = 0; half < 2; ++half) {
      sb::mbar_wait(s_full, static_cast<uint32_t>(half));
      sb::tc_fence_after();
      if (half == 0 && t < 128) {  // Q(half 0) has been consumed by the tensor core: stage the second half's queries
        const bf16* row = p.qkv + token(128 + t) * (3ll * C) + head * 72;
        fill_row(sbase + WA_OFF_Q0, sbase + WA_OFF_Q1, t, row);
      }
      // ---- pass 1: row maximum over this thread's 128 keys
      const uint32_t ts = tmem_base + tlane + static_cast<uint32_t>(hf * 128);
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        sb::tmem_ld_32x16(ts + cc * 32, v);
        sb::tmem_ld_32x16(ts + cc * 32 + 16, v + 16);
        sb::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      xch[hf * 128 + r] = mx;
      pair_barrier(q);
      mx = fmaxf(mx, xch[(hf ^ 1) * 128 + r]);
      const float2 c2 = sb::splat2(c), nm2 = sb::splat2(-mx * c);
      if (half == 1) {  // P of the first half must have been consumed by its PV before it is overwritten
        sb::mbar_wait(o_full, 0);
        sb::tc_fence_after();
      }
      // ---- pass 2: exponentials, row sum, P (bf16, unnormalised) as K-major 128B-swizzled K-blocks
      float2 l2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        sb::tmem_ld_32x16(ts + cc * 32, v);
        sb::tmem_ld_32x16(ts + cc * 32 + 16, v + 16);
        sb::tmem_ld_wait();
        const uint32_t prow = sbase + WA_OFF_P + (hf * 2 + (cc >> 1)) * 16384 + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float2 e[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 d = sb::fma2(make_float2(__uint_as_float(v[ch * 8 + 2 * j]), __uint_as_float(v[ch * 8 + 2 * j + 1])), c2, nm2);
            e[j] = make_float2(sb::fast_exp2(d.x), sb::fast_exp2(d.y));
            l2 = sb::add2(l2, e[j]);
          }
          sts128(prow + ((((cc & 1) * 4 + ch) ^ (r & 7)) << 4),
                 make_uint4(sb::pack_bf16x2(e[0].x, e[0].y), sb::pack_bf16x2(e[1].x, e[1].y), sb::pack_bf16x2(e[2].x, e[2].y),
                            sb::pack_bf16x2(e[3].x, e[3].y)));
        }
      }
      float l = l2.x + l2.y;
      xch[256 + hf * 128 + r] = l;
      sb::tc_fence_before();
      sb::fence_proxy_async();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(p_full);
      if (half == 0 && t < 128) {  // second half's queries: visible to the tensor core before its QK^T
        cp_async_wait_all();
        sb::fence_proxy_async();
        sb::mbar_arrive(q1_done);
      }
      pair_barrier(q);
      l += xch[256 + (hf ^ 1) * 128 + r];
      const float inv = 1.f / l;
      // ---- epilogue: O / l -> bf16, columns [0, 48) by the first key-half's warps, [48, 72) by the second's
      sb::mbar_wait(o_full, static_cast<uint32_t>(half));
      sb::tc_fence_after();
      bf16* orow = p.out + token(half * 128 + r) * C + head * 72;
      const uint32_t to = tmem_base + tlane + 256u;
      if (hf == 0) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          uint32_t o[16];
          sb::tmem_ld_32x16(to + cc * 16, o);
          sb::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(orow + cc * 16 + j * 8) =
                make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
        }
      } else {
#pragma unroll
        for (int cc = 3; cc < 5; ++cc) {
          uint32_t o[16];
          sb::tmem_ld_32x16(to + cc * 16, o);
          sb::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (cc * 16 + j * 8 >= 72) break;  // columns 72..79 are padding
            *reinterpret_cast<uint4*>(orow + cc * 16 + j * 8) =
                make_uint4(sb::pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                           sb::pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv));
          }
        }
      }
      sb::tc_fence_before();
      __syncwarp();
      if (lane == 0) sb::mbar_arrive(o_empty);
      pair_barrier(q);  // the exchange slots are reused by the second half
    }
  }

  sb::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    sb::tc_fence_after();
    sb::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Called by sb_window_attention when SB_WINDOW_TC=1 for ws == 16, pool == 1, head_dim 72, H and W multiples of 16.
int sb_internal_window_attn_tc(const void* qkv, void* o, int batch, int H, int W, int heads, float scale,
                               cudaStream_t stream) {
  if ((H % 16) != 0 || (W % 16) != 0 || ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o)) & 15) != 0)
    return SB_ERR_UNSUPPORTED;
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
    attr_once.mark();
  }
  WinAttnParams p;
  p.qkv = static_cast<const bf16*>(qkv);
  p.out = static_cast<bf16*>(o);
  p.H = H;
  p.W = W;
  p.heads = heads;
  p.nwx = W / 16;
  p.nwy = H / 16;
  p.scale_log2 = scale * 1.4426950408889634f;
  window_attn_tc_kernel<<<dim3(batch * p.nwx * p.nwy, heads), WA_THREADS, WA_SMEM, stream>>>(p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
