// saber_b200 — bandwidth-bound token-major kernels of the SAM2 encoder / decoder:
// LayerNorm (one warp per token row, vectorised, fp32 statistics), im2col for the 7x7/s4 patch
// embedding, 2x2 max-pool, nearest-2x top-down add of the FPN neck, NHWC<->NCHW relayouts.
// Restates sam2/modeling/backbones/hieradet.py (PatchEmbed, do_pool), image_encoder.py (FpnNeck)
// and torch.nn.LayerNorm / sam2_utils.LayerNorm2d as used on the path.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim C of a row-major [M, C] matrix. in: fp32 or bf16; out: bf16 or fp32.
// act: 0 none, 1 GELU(erf) applied after the affine transform.
// ------------------------------------------------------------------------------------------
template <bool IN_F32>
__device__ __forceinline__ float ld_elem(const void* p, long long i) {
  if (IN_F32) return reinterpret_cast<const float*>(p)[i];
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

template <bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ in, long long ld_in, void* __restrict__ out,
                 long long ld_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int M, int C, float eps, int act) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
       row < M; row += static_cast<long long>(gridDim.x) * warps_per_block) {
    const long long ib = row * ld_in;
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += ld_elem<IN_F32>(in, ib + c);
    const float mean = sb::warp_sum(sum) / C;
    float vs = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = ld_elem<IN_F32>(in, ib + c) - mean;
      vs += d * d;
    }
    const float rstd = rsqrtf(sb::warp_sum(vs) / C + eps);
    const long long ob = row * ld_out;
    for (int c = lane; c < C; c += 32) {
      float y = (ld_elem<IN_F32>(in, ib + c) - mean) * rstd * gamma[c] + beta[c];
      if (act == 1) y = sb::gelu_erf(y);
      if (OUT_F32)
        reinterpret_cast<float*>(out)[ob + c] = y;
      else
        reinterpret_cast<__nv_bfloat16*>(out)[ob + c] = __float2bfloat16(y);
    }
  }
}

// Fast path for fp32 rows of up to 1536 channels (every Hiera LayerNorm): one 16-byte load per lane and 128 channels,
// the row stays in registers for the two-pass statistics and the normalisation, 8- / 16-byte stores. NV = float4 per
// lane (compile time: C <= 128 NV). Two things took the stage-3 streams ([32768, 576] fp32 -> bf16) from 2.5 TB/s up
// (profiles/r02k_encoder_batch_launches_summary.txt: 46 us per launch, 14 % of the encoder): the loads of the warp's
// NEXT row are issued before the current row is reduced (a warp always has a row in flight; before, load -> two
// shuffle reductions -> store were serial per warp), and gamma / beta stay in registers across rows for NV <= 5
// (re-reading them per row doubled the LSU wavefronts of the kernel).
template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(256)
layernorm_rows_f32_kernel(const float* __restrict__ in, long long ld_in, void* __restrict__ out, long long ld_out,
                          const float* __restrict__ gamma, const float* __restrict__ beta, int M, int C, float eps,
                          int act) {
  constexpr bool GB_REGS = NV <= 5;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nv = C >> 2;  // float4 per row
  const float inv_c = 1.f / static_cast<float>(C);
  const long long stride = static_cast<long long>(gridDim.x) * warps_per_block;
  long long row = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  if (row >= M) return;
  float4 g[GB_REGS ? NV : 1], b[GB_REGS ? NV : 1];
  if (GB_REGS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      g[i] = c4 < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      b[i] = c4 < nv ? __ldg(reinterpret_cast<const float4*>(beta) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float4 nxt[NV];
  {
    const float4* src = reinterpret_cast<const float4*>(in + row * ld_in);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      nxt[i] = c4 < nv ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (; row < M; row += stride) {
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = nxt[i];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    if (row + stride < M) {  // warp-uniform
      const float4* src = reinterpret_cast<const float4*>(in + (row + stride) * ld_in);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = lane + 32 * i;
        nxt[i] = c4 < nv ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = sb::warp_sum(sum) * inv_c;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nv) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        vs += (a * a + bb * bb) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(sb::warp_sum(vs) * inv_c + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < nv) {
        const float4 gg = GB_REGS ? g[i] : __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 be = GB_REGS ? b[i] : __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float y0 = (v[i].x - mean) * rstd * gg.x + be.x, y1 = (v[i].y - mean) * rstd * gg.y + be.y;
        float y2 = (v[i].z - mean) * rstd * gg.z + be.z, y3 = (v[i].w - mean) * rstd * gg.w + be.w;
        if (act == 1) {
          y0 = sb::gelu_erf(y0);
          y1 = sb::gelu_erf(y1);
          y2 = sb::gelu_erf(y2);
          y3 = sb::gelu_erf(y3);
        }
        if (OUT_F32)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * ld_out + 4 * c4) = make_float4(y0, y1, y2, y3);
        else
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * ld_out + 4 * c4) =
              make_uint2(sb::pack_bf16x2(y0, y1), sb::pack_bf16x2(y2, y3));
      }
    }
  }
}

template <int NV>
void launch_ln_rows(int g, cudaStream_t stream, const float* in, long long ld_in, void* out, long long ld_out, int out_f32,
                    const float* gamma, const float* beta, int M, int C, float eps, int act) {
  if (out_f32)
    layernorm_rows_f32_kernel<NV, true><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
  else
    layernorm_rows_f32_kernel<NV, false><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
}

// ------------------------------------------------------------------------------------------
// im2col for Conv2d(k=7, stride=4, pad=3): img [B, Cin, S, S] fp32 -> cols [B*(S/4)^2, Kp] bf16,
// column index = c*49 + ky*7 + kx, zero beyond Cin*49 (Kp is the padded pitch).
// ------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256)
im2col_k7s4_kernel(const float* __restrict__ img, TO* __restrict__ cols, int B, int Cin,
                   int S, int Kp) {
  const int T = S / 4;
  const long long total = static_cast<long long>(B) * T * T * Kp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kcol = static_cast<int>(i % Kp);
    const long long tok = i / Kp;
    float v = 0.f;
    if (kcol < Cin * 49) {
      const int c = kcol / 49, r = kcol % 49, ky = r / 7, kx = r % 7;
      const int tx = static_cast<int>(tok % T);
      const int ty = static_cast<int>((tok / T) % T);
      const int b = static_cast<int>(tok / (static_cast<long long>(T) * T));
      const int y = ty * 4 - 3 + ky, x = tx * 4 - 3 + kx;
      if (y >= 0 && y < S && x >= 0 && x < S)
        v = img[((static_cast<long long>(b) * Cin + c) * S + y) * S + x];
    }
    cols[i] = static_cast<TO>(v);
  }
}

// 2x2/stride-2 max pool over token-major [B, H, W, C] (floor mode), fp32 or bf16.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int x = static_cast<int>(t % Wo);
    t /= Wo;
    const int y = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const T* p = in + ((static_cast<long long>(b) * H + 2 * y) * W + 2 * x) * C + c;
    float a = static_cast<float>(p[0]), bb = static_cast<float>(p[C]);
    float cc = static_cast<float>(p[static_cast<long long>(W) * C]);
    float d = static_cast<float>(p[static_cast<long long>(W) * C + C]);
    out[i] = static_cast<T>(fmaxf(fmaxf(a, bb), fmaxf(cc, d)));
  }
}

// dst[b, y, x, c] += src[b, y/2, x/2, c]   (FPN top-down, nearest 2x), fp32 token-major.
__global__ void __launch_bounds__(256)
add_upsample2x_kernel(float* __restrict__ dst, const float* __restrict__ src, int B, int H, int W,
                      int C) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int x = static_cast<int>(t % W);
    t /= W;
    const int y = static_cast<int>(t % H);
    const int b = static_cast<int>(t / H);
    dst[i] += src[((static_cast<long long>(b) * (H / 2) + y / 2) * (W / 2) + x / 2) * C + c];
  }
}

// [B, HW, C] (token-major) -> [B, C, HW] via a 32x32 shared-memory transpose; optional per-channel
// add (e.g. no_mem_embed). TIN/TOUT in {float, bf16}.
template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const TIN* __restrict__ in, TOUT* __restrict__ out, int HW, int C,
                    const float* __restrict__ chan_add) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    if (t < HW && c < C)
      tile[r][tx] = static_cast<float>(in[(static_cast<long long>(b) * HW + t) * C + c]);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (t < HW && c < C) {
      float v = tile[tx][r];
      if (chan_add) v += chan_add[c];
      out[(static_cast<long long>(b) * C + c) * HW + t] = static_cast<TOUT>(v);
    }
  }
}

// [B, C, HW] -> [B, HW, C]
template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const TIN* __restrict__ in, TOUT* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (t < HW && c < C)
      tile[r][tx] = static_cast<float>(in[(static_cast<long long>(b) * C + c) * HW + t]);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    if (t < HW && c < C)
      out[(static_cast<long long>(b) * HW + t) * C + c] = static_cast<TOUT>(tile[tx][r]);
  }
}

// out = a (+ b) with dtype conversion; n elements. b may be null. Row-broadcast of b via b_mod
// (b index = i % b_mod when b_mod > 0).
template <typename TA, typename TO>
__global__ void __launch_bounds__(256)
add_cast_kernel(const TA* __restrict__ a, const float* __restrict__ b, TO* __restrict__ out,
                long long n, long long b_mod) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = static_cast<float>(a[i]);
    if (b) v += b[b_mod > 0 ? (i % b_mod) : i];
    out[i] = static_cast<TO>(v);
  }
}

inline int grid_for(long long total, int block = 256, int max_blocks = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

extern "C" int sb_layernorm(const void* in, long long ld_in, int in_f32, void* out,
                            long long ld_out, int out_f32, const float* gamma, const float* beta,
                            int M, int C, float eps, int act, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(M > 0 && C > 0, "sb_layernorm: empty problem");
  const int warps = 8;
  long long blocks = (static_cast<long long>(M) + warps - 1) / warps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int g = static_cast<int>(blocks);
  const bool fast = in_f32 && (C % 4) == 0 && C <= 1536 && (ld_in % 4) == 0 && (ld_out % 4) == 0 &&
                    ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (fast) {
    const float* inf = static_cast<const float*>(in);
    const int nvl = (C / 4 + 31) / 32;  // float4 per lane
    if (nvl <= 1) launch_ln_rows<1>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    else if (nvl <= 2) launch_ln_rows<2>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    else if (nvl <= 3) launch_ln_rows<3>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    else if (nvl <= 5) launch_ln_rows<5>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    else if (nvl <= 9) launch_ln_rows<9>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    else launch_ln_rows<12>(g, stream, inf, ld_in, out, ld_out, out_f32, gamma, beta, M, C, eps, act);
    SB_CHECK_LAUNCH();
    return SB_OK;
  }
  if (in_f32 && out_f32)
    layernorm_kernel<true, true><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
  else if (in_f32)
    layernorm_kernel<true, false><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
  else if (out_f32)
    layernorm_kernel<false, true><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
  else
    layernorm_kernel<false, false><<<g, 256, 0, stream>>>(in, ld_in, out, ld_out, gamma, beta, M, C, eps, act);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_im2col_k7s4(const float* img, void* cols, int B, int Cin, int S, int Kp,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(S % 4 == 0 && Kp >= Cin * 49 && (Kp % 8) == 0, "sb_im2col_k7s4: bad S=%d Kp=%d", S, Kp);
  const long long total = static_cast<long long>(B) * (S / 4) * (S / 4) * Kp;
  im2col_k7s4_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(
      img, static_cast<__nv_bfloat16*>(cols), B, Cin, S, Kp);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// fp32 patches (validation mode)
extern "C" int sb_im2col_k7s4_f32(const float* img, float* cols, int B, int Cin, int S, int Kp, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(S % 4 == 0 && Kp >= Cin * 49 && (Kp % 8) == 0, "sb_im2col_k7s4_f32: bad S=%d Kp=%d", S, Kp);
  const long long total = static_cast<long long>(B) * (S / 4) * (S / 4) * Kp;
  im2col_k7s4_kernel<float><<<grid_for(total), 256, 0, stream>>>(img, cols, B, Cin, S, Kp);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_maxpool2x2(const void* in, void* out, int is_f32, int B, int H, int W, int C,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * C;
  SB_REQUIRE(total > 0, "sb_maxpool2x2: empty problem");
  if (is_f32)
    maxpool2x2_kernel<float><<<grid_for(total), 256, 0, stream>>>(
        static_cast<const float*>(in), static_cast<float*>(out), B, H, W, C);
  else
    maxpool2x2_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), B, H, W, C);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_add_upsample2x(float* dst, const float* src, int B, int H, int W, int C,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE((H % 2) == 0 && (W % 2) == 0, "sb_add_upsample2x: odd grid");
  const long long total = static_cast<long long>(B) * H * W * C;
  add_upsample2x_kernel<<<grid_for(total), 256, 0, stream>>>(dst, src, B, H, W, C);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// dtype codes: 0 = bf16, 1 = fp32
extern "C" int sb_nhwc_to_nchw(const void* in, int in_f32, void* out, int out_f32, int B, int HW,
                               int C, const float* chan_add, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  if (in_f32 && out_f32)
    nhwc_to_nchw_kernel<float, float><<<grid, 256, 0, stream>>>(
        static_cast<const float*>(in), static_cast<float*>(out), HW, C, chan_add);
  else if (in_f32)
    nhwc_to_nchw_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), HW, C, chan_add);
  else if (out_f32)
    nhwc_to_nchw_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), HW, C, chan_add);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), HW, C, chan_add);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_nchw_to_nhwc(const void* in, int in_f32, void* out, int out_f32, int B, int HW,
                               int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  if (in_f32 && out_f32)
    nchw_to_nhwc_kernel<float, float><<<grid, 256, 0, stream>>>(
        static_cast<const float*>(in), static_cast<float*>(out), HW, C);
  else if (in_f32)
    nchw_to_nhwc_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out), HW, C);
  else if (out_f32)
    nchw_to_nhwc_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<float*>(out), HW, C);
  else
    nchw_to_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), HW, C);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// out[i] = a[i] + b[i % b_mod]  with conversion (a: bf16|fp32, out: bf16|fp32, b: fp32 or null)
extern "C" int sb_add_cast(const void* a, int a_f32, const float* b, long long b_mod, void* out,
                           int out_f32, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0, "sb_add_cast: empty");
  const int g = grid_for(n);
  if (a_f32 && out_f32)
    add_cast_kernel<float, float><<<g, 256, 0, stream>>>(static_cast<const float*>(a), b,
                                                         static_cast<float*>(out), n, b_mod);
  else if (a_f32)
    add_cast_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>(
        static_cast<const float*>(a), b, static_cast<__nv_bfloat16*>(out), n, b_mod);
  else if (out_f32)
    add_cast_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(a), b, static_cast<float*>(out), n, b_mod);
  else
    add_cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(a), b, static_cast<__nv_bfloat16*>(out), n, b_mod);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
