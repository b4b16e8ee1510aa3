// saber_b200 — SAM2 prompt-encoder / mask-decoder glue kernels (everything that is not a GEMM,
// LayerNorm or attention): point-prompt token assembly with random-Fourier positional encoding,
// the fused mask-prompt down-scaling convs, the two transposed-conv post stages (pixel shuffle +
// skip + LayerNorm2d + GELU, and pixel shuffle + skip + GELU + hyper-network mask product), and the
// dynamic-multimask-via-stability selection.
// Restates sam2/modeling/sam/prompt_encoder.py and mask_decoder.py (HF
// modeling_sam2.py:712-856, 859-1296) as called from REF saber/adapters/sam2/automask.py:66-78.
#include "common.cuh"

namespace {

constexpr float kTwoPi = 6.283185307179586f;

// tokens[b, 0..5] = (obj_score_token, iou_token, mask_tokens[0..3]); tokens[b, 6 + p] = point p's
// embedding; last = padding point (label -1) when pad != 0. coords are in model-input pixels.
__global__ void __launch_bounds__(256)
prompt_tokens_kernel(const float* __restrict__ coords, const int* __restrict__ labels, int B, int Np,
                     int pad, const float* __restrict__ gauss /*[2,128]*/,
                     const float* __restrict__ point_emb /*[4,256]*/,
                     const float* __restrict__ not_a_point /*[256]*/,
                     const float* __restrict__ out_tokens /*[6,256]*/, float inv_size,
                     float* __restrict__ tokens, int Nt) {
  const int b = blockIdx.x;
  const int c = threadIdx.x;  // 0..255
  float* tb = tokens + static_cast<long long>(b) * Nt * 256;
  for (int t = 0; t < 6; ++t) tb[t * 256 + c] = out_tokens[t * 256 + c];
  const int f = c & 127;
  for (int p = 0; p < Np + (pad ? 1 : 0); ++p) {
    float x = 0.f, y = 0.f;
    int label = -1;
    if (p < Np) {
      x = coords[(static_cast<long long>(b) * Np + p) * 2 + 0];
      y = coords[(static_cast<long long>(b) * Np + p) * 2 + 1];
      label = labels[static_cast<long long>(b) * Np + p];
    }
    // (coord + 0.5) / size -> [0,1] -> [-1,1] -> @ gaussian -> * 2pi -> [sin | cos]
    const float nx = 2.f * ((x + 0.5f) * inv_size) - 1.f;
    const float ny = 2.f * ((y + 0.5f) * inv_size) - 1.f;
    const float proj = kTwoPi * (nx * gauss[f] + ny * gauss[128 + f]);
    float v = (c < 128) ? sinf(proj) : cosf(proj);
    if (label == -1)
      v = not_a_point[c];
    else if (label >= 0 && label < 4)
      v += point_emb[label * 256 + c];
    tb[(6 + p) * 256 + c] = v;
  }
}

// Fused mask_downscaling[0..5]: Conv2d(1->4,k2,s2) + LN2d(4) + GELU + Conv2d(4->16,k2,s2) + LN2d(16)
// + GELU. in: [B, S, S] fp32 (S = 256) -> out: [B*(S/4)^2, 16] bf16 token-major.
__global__ void __launch_bounds__(256)
mask_downscale_kernel(const float* __restrict__ in, int B, int S, int cpp, float clampv,
                      const float* __restrict__ w1,
                      const float* __restrict__ b1, const float* __restrict__ g1,
                      const float* __restrict__ be1, const float* __restrict__ w2,
                      const float* __restrict__ b2, const float* __restrict__ g2,
                      const float* __restrict__ be2, __nv_bfloat16* __restrict__ out) {
  __shared__ float sw1[16], sb1[4], sg1[4], sbe1[4], sw2[256], sb2[16], sg2[16], sbe2[16];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sw2[i] = w2[i];
  if (threadIdx.x < 16) {
    sw1[threadIdx.x] = w1[threadIdx.x];
    sb2[threadIdx.x] = b2[threadIdx.x];
    sg2[threadIdx.x] = g2[threadIdx.x];
    sbe2[threadIdx.x] = be2[threadIdx.x];
  }
  if (threadIdx.x < 4) {
    sb1[threadIdx.x] = b1[threadIdx.x];
    sg1[threadIdx.x] = g1[threadIdx.x];
    sbe1[threadIdx.x] = be1[threadIdx.x];
  }
  __syncthreads();
  const int T = S / 4;
  const long long total = static_cast<long long>(B) * T * T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tx = static_cast<int>(i % T), ty = static_cast<int>((i / T) % T);
    const int b = static_cast<int>(i / (static_cast<long long>(T) * T));
    // cpp == 3: mask prompt b is token 1 + b % 3 of prompt b / 3 in a [*, 4, S, S] decoder output (m2m pass)
    const long long pl = cpp == 3 ? static_cast<long long>(b / 3) * 4 + 1 + b % 3 : b;
    const float* src = in + (pl * S + ty * 4) * S + tx * 4;
    float h1[2][2][4];  // [py][px][channel] after conv1 + LN + GELU
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        float a00 = src[(py * 2 + 0) * S + px * 2 + 0], a01 = src[(py * 2 + 0) * S + px * 2 + 1];
        float a10 = src[(py * 2 + 1) * S + px * 2 + 0], a11 = src[(py * 2 + 1) * S + px * 2 + 1];
        if (clampv > 0.f) {  // upstream clamps the low-res logits fed back as mask prompts to +-32
          a00 = fminf(fmaxf(a00, -clampv), clampv);
          a01 = fminf(fmaxf(a01, -clampv), clampv);
          a10 = fminf(fmaxf(a10, -clampv), clampv);
          a11 = fminf(fmaxf(a11, -clampv), clampv);
        }
        float c[4], mean = 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          c[o] = sb1[o] + sw1[o * 4 + 0] * a00 + sw1[o * 4 + 1] * a01 + sw1[o * 4 + 2] * a10 +
                 sw1[o * 4 + 3] * a11;
          mean += c[o];
        }
        mean *= 0.25f;
        float var = 0.f;
#pragma unroll
        for (int o = 0; o < 4; ++o) var += (c[o] - mean) * (c[o] - mean);
        const float rstd = rsqrtf(var * 0.25f + 1e-6f);
#pragma unroll
        for (int o = 0; o < 4; ++o)
          h1[py][px][o] = sb::gelu_erf((c[o] - mean) * rstd * sg1[o] + sbe1[o]);
      }
    float c2[16], mean = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float acc = sb2[o];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px) acc += sw2[((o * 4 + ci) * 2 + py) * 2 + px] * h1[py][px][ci];
      c2[o] = acc;
      mean += acc;
    }
    mean *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o) var += (c2[o] - mean) * (c2[o] - mean);
    const float rstd = rsqrtf(var * (1.f / 16.f) + 1e-6f);
    uint32_t packed[8];
#pragma unroll
    for (int o = 0; o < 16; o += 2) {
      const float v0 = sb::gelu_erf((c2[o] - mean) * rstd * sg2[o] + sbe2[o]);
      const float v1 = sb::gelu_erf((c2[o + 1] - mean) * rstd * sg2[o + 1] + sbe2[o + 1]);
      packed[o >> 1] = sb::pack_bf16x2(v0, v1);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
    dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
  }
}


// keys[b*T + t, n] = bf16(image_embed[t, n] + bias[n] + sum_c ds[b*T + t, c] * w[n, c])   (n < 256, c < 16):
// mask_downscaling[6] (1x1 conv 16 -> 256) plus the dense-prompt add of the mask decoder (src = image_embeddings +
// dense_prompt_embeddings), i.e. the per-prompt image stream of the m2m pass, written once. A K = 16 contraction is
// not tensor-core work: one warp per token, 8 channels per lane with their 8 x 16 weights in registers; the kernel
// is bound by the 512 B it stores per token.
__global__ void __launch_bounds__(128)
mask_embed_keys_kernel(const __nv_bfloat16* __restrict__ ds, const float* __restrict__ w /*[256,16]*/,
                       const float* __restrict__ bias, const float* __restrict__ image_embed /*[T,256]*/, int T,
                       long long ntok, __nv_bfloat16* __restrict__ keys) {
  // two warps per token (4 channels per lane: 64 weight registers), two tokens per iteration with all loads issued
  // before the math, so enough bytes are in flight per SM to cover the latency of the streaming store / L2 loads
  const int lane = threadIdx.x & 31;
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int c0 = static_cast<int>(gw & 1) * 128 + lane * 4;
  float wr[4][16];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int c = 0; c < 16; ++c) wr[e][c] = w[(c0 + e) * 16 + c];
  const float4 br = *reinterpret_cast<const float4*>(bias + c0);
  // 32-bit token arithmetic (ntok < 2^31 is checked by the launcher): a 64-bit modulo per token costs more
  // instructions than the 64 FMAs of the contraction
  const unsigned npairs = (gridDim.x * blockDim.x) >> 6;  // warp pairs in the grid
  const unsigned n = static_cast<unsigned>(ntok), uT = static_cast<unsigned>(T);
  for (unsigned t0 = static_cast<unsigned>(gw >> 1) * 2; t0 < n; t0 += npairs * 2) {
    uint4 d[2][2];
    float4 ie[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned tok = t0 + u < n ? t0 + u : t0;
      d[u][0] = *reinterpret_cast<const uint4*>(ds + static_cast<size_t>(tok) * 16);
      d[u][1] = *reinterpret_cast<const uint4*>(ds + static_cast<size_t>(tok) * 16 + 8);
      ie[u] = __ldg(reinterpret_cast<const float4*>(image_embed + static_cast<size_t>(tok % uT) * 256 + c0));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t0 + u >= n) break;
      float dv[16];
      dv[0] = sb::bf16_lo(d[u][0].x); dv[1] = sb::bf16_hi(d[u][0].x); dv[2] = sb::bf16_lo(d[u][0].y); dv[3] = sb::bf16_hi(d[u][0].y);
      dv[4] = sb::bf16_lo(d[u][0].z); dv[5] = sb::bf16_hi(d[u][0].z); dv[6] = sb::bf16_lo(d[u][0].w); dv[7] = sb::bf16_hi(d[u][0].w);
      dv[8] = sb::bf16_lo(d[u][1].x); dv[9] = sb::bf16_hi(d[u][1].x); dv[10] = sb::bf16_lo(d[u][1].y); dv[11] = sb::bf16_hi(d[u][1].y);
      dv[12] = sb::bf16_lo(d[u][1].z); dv[13] = sb::bf16_hi(d[u][1].z); dv[14] = sb::bf16_lo(d[u][1].w); dv[15] = sb::bf16_hi(d[u][1].w);
      float acc[4] = {ie[u].x + br.x, ie[u].y + br.y, ie[u].z + br.z, ie[u].w + br.w};
#pragma unroll
      for (int c = 0; c < 16; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(dv[c], wr[e][c], acc[e]);
      *reinterpret_cast<uint2*>(keys + static_cast<size_t>(t0 + u) * 256 + c0) =
          make_uint2(sb::pack_bf16x2(acc[0], acc[1]), sb::pack_bf16x2(acc[2], acc[3]));
    }
  }
}

// Stage 1 of output_upscaling after the ConvTranspose2d(256->64,k2,s2) GEMM:
// g1 [B*h*w, 4*64] (col = (dy*2+dx)*64 + co, bias included) -> pixel shuffle -> + feat_s1 ->
// LayerNorm2d(64) -> GELU -> u1 [B*(2h)*(2w), 64] bf16. One warp per (token, dydx).
__global__ void __launch_bounds__(256)
upscale1_post_kernel(const __nv_bfloat16* __restrict__ g1, const float* __restrict__ feat_s1,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int B, int h,
                     int w, long long s1_batch_stride, __nv_bfloat16* __restrict__ u1) {
  const int lane = threadIdx.x & 31;
  const long long nwarp_total = static_cast<long long>(B) * h * w * 4;
  const float g0 = gamma[2 * lane], g1v = gamma[2 * lane + 1];
  const float b0 = beta[2 * lane], b1v = beta[2 * lane + 1];
  for (long long wi = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
       wi < nwarp_total; wi += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int d = static_cast<int>(wi & 3);
    const long long tok = wi >> 2;
    const int x = static_cast<int>(tok % w), y = static_cast<int>((tok / w) % h);
    const long long b = tok / (static_cast<long long>(w) * h);
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    const uint32_t gv = *reinterpret_cast<const uint32_t*>(g1 + tok * 256 + d * 64 + 2 * lane);
    const float2 sv = *reinterpret_cast<const float2*>(
        feat_s1 + b * s1_batch_stride + (static_cast<long long>(oy) * (2 * w) + ox) * 64 + 2 * lane);
    const float v0 = sb::bf16_lo(gv) + sv.x, v1 = sb::bf16_hi(gv) + sv.y;
    const float mean = sb::warp_sum(v0 + v1) * (1.f / 64.f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float rstd = rsqrtf(sb::warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-6f);
    const float o0 = sb::gelu_erf(d0 * rstd * g0 + b0), o1 = sb::gelu_erf(d1 * rstd * g1v + b1v);
    *reinterpret_cast<uint32_t*>(u1 + ((b * (2 * h) + oy) * (2 * w) + ox) * 64 + 2 * lane) =
        sb::pack_bf16x2(o0, o1);
  }
}

// Stage 2: g2 [B*H1*W1, 4*32] (col = (dy*2+dx)*32 + co, bias included) -> pixel shuffle -> + feat_s0
// -> GELU -> dot with hyper_in[b, m, 0..31] -> masks [B, 4, 2*H1, 2*W1] fp32. One warp per token:
// lane = dydx*8 + cg handles 4 channels of one of the 4 output pixels.
__global__ void __launch_bounds__(256)
upscale2_mask_kernel(const __nv_bfloat16* __restrict__ g2, const float* __restrict__ feat_s0,
                     const float* __restrict__ hyper, int B, int H1, int W1,
                     long long s0_batch_stride, float* __restrict__ masks) {
  const int lane = threadIdx.x & 31;
  const int d = lane >> 3, cg = lane & 7;
  const long long ntok = static_cast<long long>(B) * H1 * W1;
  const int H2 = 2 * H1, W2 = 2 * W1;
  for (long long tok = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; tok < ntok;
       tok += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int x = static_cast<int>(tok % W1), y = static_cast<int>((tok / W1) % H1);
    const long long b = tok / (static_cast<long long>(W1) * H1);
    const int oy = 2 * y + (d >> 1), ox = 2 * x + (d & 1);
    const uint2 gv = *reinterpret_cast<const uint2*>(g2 + tok * 128 + d * 32 + cg * 4);
    const float4 sv = *reinterpret_cast<const float4*>(
        feat_s0 + b * s0_batch_stride + (static_cast<long long>(oy) * W2 + ox) * 32 + cg * 4);
    const float v0 = sb::gelu_erf(sb::bf16_lo(gv.x) + sv.x);
    const float v1 = sb::gelu_erf(sb::bf16_hi(gv.x) + sv.y);
    const float v2 = sb::gelu_erf(sb::bf16_lo(gv.y) + sv.z);
    const float v3 = sb::gelu_erf(sb::bf16_hi(gv.y) + sv.w);
    float acc[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float4 hv = *reinterpret_cast<const float4*>(hyper + (b * 4 + m) * 32 + cg * 4);
      acc[m] = v0 * hv.x + v1 * hv.y + v2 * hv.z + v3 * hv.w;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], 1);
      acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], 2);
      acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], 4);
    }
    if (cg < 4) {
      // lanes cg=0..3 of each pixel group store mask m = cg
      const float val = cg == 0 ? acc[0] : (cg == 1 ? acc[1] : (cg == 2 ? acc[2] : acc[3]));
      masks[((b * 4 + cg) * H2 + oy) * W2 + ox] = val;
    }
  }
}

// dynamic_multimask_via_stability (single-mask output): per prompt, stability of mask token 0 =
// count(logit > delta) / count(logit > -delta) (1 when the union is empty); if >= thresh keep token 0
// else the best-IoU token among 1..3 (first max). Writes the chosen token index and IoU.
__global__ void __launch_bounds__(256)
select_mask_kernel(const float* __restrict__ masks /*[B,4,HW]*/, const float* __restrict__ ious /*[B,4]*/,
                   int HW, float delta, float thresh, int* __restrict__ sel_idx,
                   float* __restrict__ sel_iou) {
  const int b = blockIdx.x;
  const float* m0 = masks + static_cast<long long>(b) * 4 * HW;
  int ci = 0, cu = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = m0[i];
    ci += v > delta;
    cu += v > -delta;
  }
  __shared__ int si[8], su[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ci += __shfl_xor_sync(0xffffffffu, ci, o);
    cu += __shfl_xor_sync(0xffffffffu, cu, o);
  }
  if ((threadIdx.x & 31) == 0) {
    si[threadIdx.x >> 5] = ci;
    su[threadIdx.x >> 5] = cu;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ti = 0, tu = 0;
    for (int k = 0; k < 8; ++k) {
      ti += si[k];
      tu += su[k];
    }
    const float stab = tu > 0 ? static_cast<float>(ti) / static_cast<float>(tu) : 1.0f;
    int idx = 0;
    if (!(stab >= thresh)) {
      idx = 1;
      float best = ious[b * 4 + 1];
      for (int k = 2; k < 4; ++k)
        if (ious[b * 4 + k] > best) {
          best = ious[b * 4 + k];
          idx = k;
        }
    }
    sel_idx[b] = idx;
    sel_iou[b] = ious[b * 4 + idx];
  }
}

inline int blocks_for(long long threads_needed, int block = 256, int cap = 148 * 16) {
  long long g = (threads_needed + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

extern "C" int sb_prompt_tokens(const float* coords, const int* labels, int B, int Np, int pad,
                                const float* gauss, const float* point_emb, const float* not_a_point,
                                const float* out_tokens, int image_size, float* tokens, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && Np >= 0, "sb_prompt_tokens: bad sizes");
  const int Nt = 6 + Np + (pad ? 1 : 0);
  prompt_tokens_kernel<<<B, 256, 0, stream>>>(coords, labels, B, Np, pad, gauss, point_emb,
                                              not_a_point, out_tokens, 1.0f / image_size, tokens, Nt);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mask_downscale(const float* in, int B, int S, int cpp, float clampv, const float* w1,
                                 const float* b1,
                                 const float* g1, const float* be1, const float* w2, const float* b2,
                                 const float* g2, const float* be2, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && S % 4 == 0 && (cpp == 1 || cpp == 3), "sb_mask_downscale: bad sizes");
  const long long total = static_cast<long long>(B) * (S / 4) * (S / 4);
  mask_downscale_kernel<<<blocks_for(total), 256, 0, stream>>>(
      in, B, S, cpp, clampv, w1, b1, g1, be1, w2, b2, g2, be2, static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mask_embed_keys(const void* ds, const float* w, const float* bias, const float* image_embed, int T,
                                  long long ntok, void* keys, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(ds && w && bias && image_embed && keys && T > 0 && ntok > 0 && ntok < (1ll << 31),
             "sb_mask_embed_keys: bad arguments");
  long long blocks = (ntok + 3) / 4;  // 4 warps = 2 warp pairs = 4 tokens per block iteration
  if (blocks > 148 * 16) blocks = 148 * 16;
  mask_embed_keys_kernel<<<static_cast<int>(blocks), 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(ds), w, bias, image_embed, T, ntok, static_cast<__nv_bfloat16*>(keys));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_upscale1_post(const void* g1, const float* feat_s1, long long s1_batch_stride,
                                const float* gamma, const float* beta, int B, int h, int w, void* u1,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && h > 0 && w > 0, "sb_upscale1_post: bad sizes");
  const long long threads = static_cast<long long>(B) * h * w * 4 * 32;
  upscale1_post_kernel<<<blocks_for(threads, 256, 148 * 32), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(g1), feat_s1, gamma, beta, B, h, w, s1_batch_stride,
      static_cast<__nv_bfloat16*>(u1));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_upscale2_mask(const void* g2, const float* feat_s0, long long s0_batch_stride,
                                const float* hyper, int B, int H1, int W1, float* masks,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && H1 > 0 && W1 > 0, "sb_upscale2_mask: bad sizes");
  const long long threads = static_cast<long long>(B) * H1 * W1 * 32;
  upscale2_mask_kernel<<<blocks_for(threads, 256, 148 * 32), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(g2), feat_s0, hyper, B, H1, W1, s0_batch_stride, masks);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_select_mask(const float* masks, const float* ious, int B, int HW, float delta,
                              float thresh, int* sel_idx, float* sel_iou, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && HW > 0, "sb_select_mask: bad sizes");
  select_mask_kernel<<<B, 256, 0, stream>>>(masks, ious, HW, delta, thresh, sel_idx, sel_iou);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
