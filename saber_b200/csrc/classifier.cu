// saber_b200 — kernels of SABER's expert classifier path that are not GEMM / LayerNorm / max-pool:
//   * monai NormalizeIntensity of the slice (global mean / population std),           REF classifier/models/predictor.py:60,142
//   * per-mask bounding boxes, adaptive crop + bilinear (image) / nearest (mask) resize to 320^2 with the cropped mask's
//     area,                                                                            REF classifier/datasets/RandMaskCrop.py:44-203
//   * ROI / RONI masking of the SAM2 image embedding (nearest-resized mask),           REF classifier/models/SAM2.py:164-197
//   * PReLU, 3x3 stride-1 im2col, token mean (adaptive_avg_pool2d), row softmax        REF classifier/models/SAM2.py:59-88,152-161
#include "common.cuh"
#include <float.h>

namespace {

inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

__global__ void __launch_bounds__(256)
sum_sumsq_partial_kernel(const float* __restrict__ in, long long n, double* __restrict__ partials) {
  double s = 0.0, q = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double v = in[i];
    s += v;
    q += v * v;
  }
  __shared__ double ss[256], sq[256];
  ss[threadIdx.x] = s;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sq[threadIdx.x] += sq[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = ss[0];
    partials[2 * blockIdx.x + 1] = sq[0];
  }
}

__global__ void mean_std_final_kernel(const double* __restrict__ partials, int nb, long long n, float* __restrict__ out) {
  double s = 0.0, q = 0.0;
  for (int i = 0; i < nb; ++i) {
    s += partials[2 * i];
    q += partials[2 * i + 1];
  }
  const double mean = s / static_cast<double>(n);
  const double var = fmax(q / static_cast<double>(n) - mean * mean, 0.0);
  out[0] = static_cast<float>(mean);
  out[1] = static_cast<float>(sqrt(var));
}

// out = (in - ms[0]) / ms[1]   (division skipped when std == 0, as monai does)
__global__ void __launch_bounds__(256)
standardize_kernel(const float* __restrict__ in, long long n, const float* __restrict__ ms, float* __restrict__ out) {
  const float m = ms[0], s = ms[1];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = __fsub_rn(in[i], m);
    out[i] = s != 0.f ? __fdiv_rn(d, s) : d;
  }
}

// bbox[n] = (y_min, y_max, x_min, x_max) of the non-zero pixels of masks[n] ([H, W] uint8), or (-1,-1,-1,-1)
__global__ void __launch_bounds__(256)
mask_bbox_kernel(const unsigned char* __restrict__ masks, int H, int W, int* __restrict__ bbox) {
  const unsigned char* m = masks + static_cast<long long>(blockIdx.x) * H * W;
  int y0 = INT_MAX, y1 = -1, x0 = INT_MAX, x1 = -1;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    if (m[i]) {
      const int y = i / W, x = i % W;
      y0 = min(y0, y);
      y1 = max(y1, y);
      x0 = min(x0, x);
      x1 = max(x1, x);
    }
  }
  __shared__ int s[4][256];
  s[0][threadIdx.x] = y0;
  s[1][threadIdx.x] = y1;
  s[2][threadIdx.x] = x0;
  s[3][threadIdx.x] = x1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] = min(s[0][threadIdx.x], s[0][threadIdx.x + o]);
      s[1][threadIdx.x] = max(s[1][threadIdx.x], s[1][threadIdx.x + o]);
      s[2][threadIdx.x] = min(s[2][threadIdx.x], s[2][threadIdx.x + o]);
      s[3][threadIdx.x] = max(s[3][threadIdx.x], s[3][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const bool empty = s[1][0] < 0;
    bbox[4 * blockIdx.x + 0] = empty ? -1 : s[0][0];
    bbox[4 * blockIdx.x + 1] = empty ? -1 : s[1][0];
    bbox[4 * blockIdx.x + 2] = empty ? -1 : s[2][0];
    bbox[4 * blockIdx.x + 3] = empty ? -1 : s[3][0];
  }
}

// ATen area_pixel_compute_source_index (align_corners = False) + bilinear blend, same expression tree as torch CPU
__device__ __forceinline__ void bl_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fmaf_rn(scale, __fadd_rn(static_cast<float>(dst), 0.5f), -0.5f);
  if (s < 0.f) s = 0.f;
  i0 = min(static_cast<int>(floorf(s)), in_size - 1);
  l1 = fminf(fmaxf(__fsub_rn(s, static_cast<float>(i0)), 0.f), 1.f);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l0 = __fsub_rn(1.f, l1);
}

// For mask n: crop (top, left, ch, cw) = geom[n] of img [H, W] fp32 and masks[n] [H, W] u8 -> bilinear / nearest resize
// to S x S; area[n] = number of set pixels of the resized mask.
__global__ void __launch_bounds__(256)
crop_resize_kernel(const float* __restrict__ img, const unsigned char* __restrict__ masks, const int* __restrict__ geom,
                   int H, int W, int S, float* __restrict__ out_img, unsigned char* __restrict__ out_mask,
                   int* __restrict__ area) {
  const int n = blockIdx.y;
  const int top = geom[4 * n], left = geom[4 * n + 1], ch = geom[4 * n + 2], cw = geom[4 * n + 3];
  const float sh = __fdiv_rn(static_cast<float>(ch), static_cast<float>(S));
  const float sw = __fdiv_rn(static_cast<float>(cw), static_cast<float>(S));
  const unsigned char* m = masks + static_cast<long long>(n) * H * W;
  int cnt = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < S * S; t += gridDim.x * blockDim.x) {
    const int ox = t % S, oy = t / S;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bl_index(sh, oy, ch, y0, y1, ly0, ly1);
    bl_index(sw, ox, cw, x0, x1, lx0, lx1);
    const float* p = img + static_cast<long long>(top) * W + left;
    const float v00 = p[y0 * W + x0], v01 = p[y0 * W + x1], v10 = p[y1 * W + x0], v11 = p[y1 * W + x1];
    const float t0 = __fmaf_rn(v00, lx0, __fmul_rn(v01, lx1));
    const float t1 = __fmaf_rn(v10, lx0, __fmul_rn(v11, lx1));
    out_img[(static_cast<long long>(n) * S + oy) * S + ox] = __fmaf_rn(t0, ly0, __fmul_rn(t1, ly1));
    // F.interpolate(mode='nearest'): src = min(floor(dst * scale), in - 1)
    const int sy = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(oy), sh))), ch - 1);
    const int sx = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(ox), sw))), cw - 1);
    const unsigned char mv = m[static_cast<long long>(top + sy) * W + left + sx] != 0;
    out_mask[(static_cast<long long>(n) * S + oy) * S + ox] = mv;
    cnt += mv;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&area[n], cnt);
}

// out[b*T + t, 0:C] = feat[b*T + t, :] * m ; out[.., C:2C] = feat * (1 - m), m = mask[b, floor(y * S/G), floor(x * S/G)]
__global__ void __launch_bounds__(256)
mask_features_kernel(const float* __restrict__ feat, const unsigned char* __restrict__ mask, int B, int G, int S, int C,
                     __nv_bfloat16* __restrict__ out) {
  const float sc = __fdiv_rn(static_cast<float>(S), static_cast<float>(G));
  const int c4 = C / 4;
  const long long total = static_cast<long long>(B) * G * G * c4;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(t % c4);
    const long long tok = t / c4;
    const int x = static_cast<int>(tok % G), y = static_cast<int>((tok / G) % G);
    const long long b = tok / (static_cast<long long>(G) * G);
    const int sy = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(y), sc))), S - 1);
    const int sx = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(x), sc))), S - 1);
    const float m = mask[(b * S + sy) * S + sx] ? 1.f : 0.f;
    const float4 f = *reinterpret_cast<const float4*>(feat + tok * C + cc * 4);
    const float im = 1.f - m;
    __nv_bfloat16* o = out + tok * 2 * C + cc * 4;
    *reinterpret_cast<uint2*>(o) = make_uint2(sb::pack_bf16x2(f.x * m, f.y * m), sb::pack_bf16x2(f.z * m, f.w * m));
    *reinterpret_cast<uint2*>(o + C) = make_uint2(sb::pack_bf16x2(f.x * im, f.y * im), sb::pack_bf16x2(f.z * im, f.w * im));
  }
}

template <typename TIN>
__global__ void __launch_bounds__(256)
prelu_kernel(const TIN* __restrict__ in, long long n, float slope, __nv_bfloat16* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v;
    if constexpr (sizeof(TIN) == 4) v = in[i];
    else v = __bfloat162float(in[i]);
    out[i] = __float2bfloat16(v >= 0.f ? v : v * slope);
  }
}

// im2col for Conv2d(k3, s1, p1) on NHWC bf16: in [B, H, W, C] -> cols [B*H*W, 9*C], column = (ky*3+kx)*C + c
__global__ void __launch_bounds__(256)
im2col_3x3s1_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C, __nv_bfloat16* __restrict__ cols) {
  const int c8 = C / 8;
  const long long total = static_cast<long long>(B) * H * W * 9 * c8;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(t % c8);
    const int k = static_cast<int>((t / c8) % 9);
    const long long pix = t / (9ll * c8);
    const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
    const long long b = pix / (static_cast<long long>(W) * H);
    const int iy = y - 1 + k / 3, ix = x - 1 + k % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(in + ((b * H + iy) * W + ix) * C + cc * 8);
    *reinterpret_cast<uint4*>(cols + pix * 9 * C + k * C + cc * 8) = v;
  }
}

// out[b, c] = mean_t in[b, t, c]   (bf16 in, fp32 out)
__global__ void __launch_bounds__(256)
mean_tokens_kernel(const __nv_bfloat16* __restrict__ in, int T, int C, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += __bfloat162float(in[(static_cast<long long>(b) * T + t) * C + c]);
    out[b * C + c] = acc / static_cast<float>(T);
  }
}

__global__ void softmax_rows_kernel(const float* __restrict__ in, int B, int C, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float mx = -FLT_MAX;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, in[b * C + c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(in[b * C + c] - mx);
  for (int c = 0; c < C; ++c) out[b * C + c] = expf(in[b * C + c] - mx) / s;
}

}  // namespace

// ms[0] = mean, ms[1] = population std of n floats. ws: 2 * 1024 doubles.
extern "C" int sb_mean_std(const float* in, long long n, float* ms, double* ws, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && in && ms && ws, "sb_mean_std: bad arguments");
  const int nb = grid_for(n, 256, 1024);
  sum_sumsq_partial_kernel<<<nb, 256, 0, stream>>>(in, n, ws);
  SB_CHECK_LAUNCH();
  mean_std_final_kernel<<<1, 1, 0, stream>>>(ws, nb, n, ms);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_standardize(const float* in, long long n, const float* ms, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && in && ms && out, "sb_standardize: bad arguments");
  standardize_kernel<<<grid_for(n), 256, 0, stream>>>(in, n, ms, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mask_bbox(const unsigned char* masks, int N, int H, int W, int* bbox, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(N > 0 && H > 0 && W > 0 && masks && bbox, "sb_mask_bbox: bad arguments");
  mask_bbox_kernel<<<N, 256, 0, stream>>>(masks, H, W, bbox);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// geom: device int32 [N,4] (top, left, crop_h, crop_w); area: device int32 [N], zeroed by the caller.
extern "C" int sb_crop_resize(const float* img, const unsigned char* masks, const int* geom, int N, int H, int W, int S,
                              float* out_img, unsigned char* out_mask, int* area, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(N > 0 && H > 0 && W > 0 && S > 0 && img && masks && geom && out_img && out_mask && area,
             "sb_crop_resize: bad arguments");
  dim3 grid((S * S + 255) / 256 > 64 ? 64 : (S * S + 255) / 256, N);
  crop_resize_kernel<<<grid, 256, 0, stream>>>(img, masks, geom, H, W, S, out_img, out_mask, area);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mask_features(const float* feat, const unsigned char* mask, int B, int G, int S, int C, void* out,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && G > 0 && S > 0 && C > 0 && (C % 4) == 0, "sb_mask_features: bad arguments");
  mask_features_kernel<<<grid_for(static_cast<long long>(B) * G * G * (C / 4)), 256, 0, stream>>>(
      feat, mask, B, G, S, C, static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_prelu(const void* in, int in_f32, long long n, float slope, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && in && out, "sb_prelu: bad arguments");
  if (in_f32)
    prelu_kernel<float><<<grid_for(n), 256, 0, stream>>>(static_cast<const float*>(in), n, slope,
                                                         static_cast<__nv_bfloat16*>(out));
  else
    prelu_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), n, slope,
                                                                 static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_im2col_3x3s1(const void* in, int B, int H, int W, int C, void* cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0, "sb_im2col_3x3s1: bad arguments");
  im2col_3x3s1_kernel<<<grid_for(static_cast<long long>(B) * H * W * 9 * (C / 8)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in), B, H, W, C, static_cast<__nv_bfloat16*>(cols));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mean_tokens(const void* in, int B, int T, int C, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && T > 0 && C > 0, "sb_mean_tokens: bad arguments");
  mean_tokens_kernel<<<B, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), T, C, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_softmax_rows(const float* in, int B, int C, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && C > 0, "sb_softmax_rows: bad arguments");
  softmax_rows_kernel<<<(B + 127) / 128, 128, 0, stream>>>(in, B, C, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
