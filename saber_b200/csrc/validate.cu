// saber_b200 — fp32 VALIDATION MODE (BASELINE north_star: "1e-4 in the fp32 validation mode"). The production path
// computes in bf16 (2e-2 tolerance); this mode separates rounding from bugs by running the SAME host orchestration
// (layouts, window addressing, pooling, residual wiring) and the SAME tcgen05 GEMM kernel at fp32 accuracy:
//
//  * sb_split3_bf16: x (fp32) = h + m + l with h = bf16(x), m = bf16(x - h), l = bf16(x - h - m) (24 mantissa bits in
//    three bf16 values). A product A W^T is then the sum of the six partial products with i + j <= 4,
//        a_h w_h + a_h w_m + a_h w_l + a_m w_h + a_m w_m + a_l w_h      (dropped terms <= 2^-25 relative),
//    which is ONE bf16 GEMM with K' = 6 K on K-concatenated operands: activations [h h h m m l], weights [h m l h m h].
//    bf16 x bf16 products are exact in fp32 and the tensor core accumulates in fp32 (TMEM), so the kernel under test
//    — tiles, TMA, epilogues, residuals — is the production one.
//  * sb_window_attention_f32: Hiera's windowed / global / q-pooled attention in plain fp32 on CUDA cores (one block per
//    query and head; scores in shared memory), addressing the fused qkv buffer exactly like the bf16 kernels.
// Replaces nothing on the product path; used by saber_b200.ops when SB_VALIDATE_FP32 is on
// (upstream arithmetic: sam2/modeling/backbones/hieradet.py, torch fp32).
#include "common.cuh"
#include <float.h>

namespace {

__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, long long ldx, int M, int K, int role, __nv_bfloat16* __restrict__ out,
              long long ldo) {
  const long long total = static_cast<long long>(M) * K;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const long long m = i / K;
    const float v = x[m * ldx + k];
    const __nv_bfloat16 h = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 mid = __float2bfloat16(r1);
    const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
    __nv_bfloat16* o = out + m * ldo + k;
    if (role == 0) {  // activations: [h h h m m l]
      o[0] = h;
      o[K] = h;
      o[2 * K] = h;
      o[3 * K] = mid;
      o[4 * K] = mid;
      o[5 * K] = lo;
    } else {  // weights: [h m l h m h]
      o[0] = h;
      o[K] = mid;
      o[2 * K] = lo;
      o[3 * K] = h;
      o[4 * K] = mid;
      o[5 * K] = h;
    }
  }
}

// One block per (batch, output token, head). Keys: the ws x ws window of the (unpooled) grid that contains the query;
// positions beyond H / W (ragged windows) carry qkv_bias, as the zero-padded tokens do upstream. pool = 2: the query
// is the element-wise max over its 2 x 2 tokens.
__global__ void __launch_bounds__(128)
window_attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias, float* __restrict__ out,
                            int H, int W, int heads, int hd, int ws, int pool, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;           // [hd]
  float* ss = sm + hd;      // [ws * ws]
  __shared__ float red[4];
  const int C = heads * hd;
  const int Ho = H / pool, Wo = W / pool;
  long long id = blockIdx.x;
  const int h = static_cast<int>(id % heads);
  id /= heads;
  const int ox = static_cast<int>(id % Wo);
  id /= Wo;
  const int oy = static_cast<int>(id % Ho);
  const int b = static_cast<int>(id / Ho);
  const int wso = ws / pool;
  const int wy = oy / wso, wx = ox / wso;
  const float* base = qkv + static_cast<long long>(b) * H * W * 3 * C;
  auto tok = [&](int y, int x, int part, int d) -> float {  // part 0 q, 1 k, 2 v
    if (y < H && x < W) return base[(static_cast<long long>(y) * W + x) * 3 * C + part * C + h * hd + d];
    return qkv_bias ? qkv_bias[part * C + h * hd + d] : 0.f;
  };
  for (int d = threadIdx.x; d < hd; d += blockDim.x) {
    float q = -FLT_MAX;
    for (int dy = 0; dy < pool; ++dy)
      for (int dx = 0; dx < pool; ++dx) q = fmaxf(q, tok(oy * pool + dy, ox * pool + dx, 0, d));
    sq[d] = q;
  }
  __syncthreads();
  const int nk = ws * ws;
  float mx = -FLT_MAX;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const int ky = wy * ws + k / ws, kx = wx * ws + k % ws;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(sq[d], tok(ky, kx, 1, d), acc);
    acc *= scale;
    ss[k] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = sb::warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const float e = expf(ss[k] - mx);
    ss[k] = e;
    sum += e;
  }
  sum = sb::warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = (red[0] + red[1]) + (red[2] + red[3]);
  const float inv = 1.f / sum;
  for (int d = threadIdx.x; d < hd; d += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < nk; ++k) {
      const int ky = wy * ws + k / ws, kx = wx * ws + k % ws;
      acc = fmaf(ss[k], tok(ky, kx, 2, d), acc);
    }
    out[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + h * hd + d] = acc * inv;
  }
}

}  // namespace

// out [M, 6K] bf16 (pitch ldo): role 0 = activation operand, 1 = weight operand of the 6-term split product.
extern "C" int sb_split3_bf16(const float* x, long long ldx, int M, int K, int role, void* out, long long ldo,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(M > 0 && K > 0 && x && out && (role == 0 || role == 1) && ldo >= 6ll * K, "sb_split3_bf16: bad arguments");
  long long g = (static_cast<long long>(M) * K + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  split3_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, ldx, M, K, role, static_cast<__nv_bfloat16*>(out), ldo);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// fp32 twin of sb_window_attention: qkv [B*H*W, 3*heads*hd] fp32 -> o [B*(H/pool)*(W/pool), heads*hd] fp32.
extern "C" int sb_window_attention_f32(const float* qkv, const float* qkv_bias, float* o, int batch, int H, int W,
                                       int heads, int hd, int ws, int pool, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(batch > 0 && H > 0 && W > 0 && heads > 0 && hd > 0, "sb_window_attention_f32: empty problem");
  SB_REQUIRE((pool == 1 || pool == 2) && ws > 0 && (ws % pool) == 0 && (H % pool) == 0 && (W % pool) == 0,
             "sb_window_attention_f32: bad window / pool");
  const size_t smem = (static_cast<size_t>(hd) + static_cast<size_t>(ws) * ws) * sizeof(float);
  SB_REQUIRE(smem <= 200 * 1024, "sb_window_attention_f32: window of %d x %d keys does not fit shared memory", ws, ws);
  static SbPerDeviceOnce once;
  if (once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(window_attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    once.mark();
  }
  const long long blocks = static_cast<long long>(batch) * (H / pool) * (W / pool) * heads;
  SB_REQUIRE(blocks < (1ll << 31), "sb_window_attention_f32: too many queries");
  window_attention_f32_kernel<<<static_cast<unsigned>(blocks), 128, smem, stream>>>(qkv, qkv_bias, o, H, W, heads, hd, ws,
                                                                                   pool, scale);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

namespace {
__global__ void __launch_bounds__(256) gelu_exact_kernel(float* __restrict__ x, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    x[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  }
}
}  // namespace

// exact (erff) GELU in place: the fused epilogue's MUFU.TANH form is accurate to 4e-4 only, above the validation bar
extern "C" int sb_gelu_exact_f32(float* x, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(x && n > 0, "sb_gelu_exact_f32: bad arguments");
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  gelu_exact_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, n);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

namespace {
// Generic batched multi-head attention in fp32 (CUDA cores): one block per (batch, head, query); optional additive key
// term shared by all batch entries (scores = q (k[b] + k_add)^T). q / k / v may hold one batch entry read by all.
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ q, long long q_ld, const float* __restrict__ k, long long k_ld,
                     const float* __restrict__ k_add, long long ka_ld, const float* __restrict__ v, long long v_ld,
                     float* __restrict__ o, long long o_ld, int heads, int hd, int nq, int nk, float scale, int q_shared,
                     int kv_shared) {
  extern __shared__ float sm[];
  float* sq = sm;       // [hd]
  float* ss = sm + hd;  // [nk]
  __shared__ float red[4];
  long long id = blockIdx.x;
  const int qi = static_cast<int>(id % nq);
  id /= nq;
  const int h = static_cast<int>(id % heads);
  const int b = static_cast<int>(id / heads);
  const float* qrow = q + (static_cast<long long>(q_shared ? 0 : b) * nq + qi) * q_ld + h * hd;
  const float* kb = k + static_cast<long long>(kv_shared ? 0 : b) * nk * k_ld + h * hd;
  const float* vb = v + static_cast<long long>(kv_shared ? 0 : b) * nk * v_ld + h * hd;
  for (int d = threadIdx.x; d < hd; d += blockDim.x) sq[d] = qrow[d];
  __syncthreads();
  float mx = -FLT_MAX;
  for (int j = threadIdx.x; j < nk; j += blockDim.x) {
    const float* kr = kb + static_cast<long long>(j) * k_ld;
    const float* ka = k_add ? k_add + static_cast<long long>(j) * ka_ld + h * hd : nullptr;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(sq[d], ka ? kr[d] + ka[d] : kr[d], acc);
    acc *= scale;
    ss[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = sb::warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += blockDim.x) {
    const float e = expf(ss[j] - mx);
    ss[j] = e;
    sum += e;
  }
  sum = sb::warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  const float inv = 1.f / ((red[0] + red[1]) + (red[2] + red[3]));
  for (int d = threadIdx.x; d < hd; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < nk; ++j) acc = fmaf(ss[j], vb[static_cast<long long>(j) * v_ld + d], acc);
    o[(static_cast<long long>(b) * nq + qi) * o_ld + h * hd + d] = acc * inv;
  }
}
}  // namespace

// fp32 twin of sb_attention / sb_attention_kadd (k_add may be NULL): F.scaled_dot_product_attention in torch fp32.
extern "C" int sb_attention_f32(const float* q, long long q_ld, const float* k, long long k_ld, const float* k_add,
                                long long ka_ld, const float* v, long long v_ld, float* o, long long o_ld, int batch,
                                int heads, int hd, int nq, int nk, float scale, int q_shared, int kv_shared,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(batch > 0 && heads > 0 && hd > 0 && nq > 0 && nk > 0 && q && k && v && o, "sb_attention_f32: bad arguments");
  const size_t smem = (static_cast<size_t>(hd) + static_cast<size_t>(nk)) * sizeof(float);
  SB_REQUIRE(smem <= 200 * 1024, "sb_attention_f32: %d keys do not fit shared memory", nk);
  static SbPerDeviceOnce once;
  if (once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    once.mark();
  }
  const long long blocks = static_cast<long long>(batch) * heads * nq;
  SB_REQUIRE(blocks < (1ll << 31), "sb_attention_f32: too many queries");
  attention_f32_kernel<<<static_cast<unsigned>(blocks), 128, smem, stream>>>(q, q_ld, k, k_ld, k_add, ka_ld, v, v_ld, o, o_ld,
                                                                            heads, hd, nq, nk, scale, q_shared, kv_shared);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

namespace {
__global__ void __launch_bounds__(256)
rope_f32_kernel(const float* __restrict__ x, long long ld_in, float* __restrict__ out, long long ld_out, long long rows,
                int C, int rows_per_batch, int n_rope, int ntok, const float2* __restrict__ cs) {
  const int half = C / 2;
  const long long total = rows * half;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t % half);
    const long long r = t / half;
    const int rb = static_cast<int>(r % rows_per_batch);
    float a = x[r * ld_in + 2 * i], b = x[r * ld_in + 2 * i + 1];
    if (rb < n_rope) {
      const float2 f = cs[static_cast<long long>(rb % ntok) * half + i];
      const float ra = a * f.x - b * f.y;
      const float rbv = a * f.y + b * f.x;
      a = ra;
      b = rbv;
    }
    out[r * ld_out + 2 * i] = a;
    out[r * ld_out + 2 * i + 1] = b;
  }
}
}  // namespace

// fp32-in / fp32-out twin of sb_rope_apply (memory attention in the validation mode)
extern "C" int sb_rope_apply_f32(const float* x, long long ld_in, float* out, long long ld_out, long long rows, int C,
                                 int rows_per_batch, int n_rope, int ntok, const float* cos_sin, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(rows > 0 && C > 0 && (C % 2) == 0 && rows_per_batch > 0 && ntok > 0 && n_rope >= 0 &&
                 n_rope <= rows_per_batch && (rows % rows_per_batch) == 0,
             "sb_rope_apply_f32: bad arguments");
  long long g = (rows * (C / 2) + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  rope_f32_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, ld_in, out, ld_out, rows, C, rows_per_batch, n_rope, ntok,
                                                           reinterpret_cast<const float2*>(cos_sin));
  SB_CHECK_LAUNCH();
  return SB_OK;
}
