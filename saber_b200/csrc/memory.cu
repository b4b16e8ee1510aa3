// saber_b200 — z-axis propagation (SAM2 video predictor) glue kernels: everything of the memory-attention /
// memory-encoder / tracking step that is not a GEMM, LayerNorm or attention:
//   * axial RoPE on projected q / k (sam2/modeling/position_encoding.py apply_rotary_enc, rope_k_repeat, object-pointer
//     tokens excluded),
//   * the mask down-sampler's 3x3 stride-2 convs + LayerNorm2d + GELU (memory_encoder.py MaskDownSampler) with the
//     sigmoid*20-10 / binarise input transform of SAM2Base._encode_new_memory, and its im2col for the last (GEMM) stage,
//   * CXBlock's 7x7 depth-wise conv fused with its LayerNorm (memory_encoder.py CXBlock),
//   * occlusion embedding + bf16 cast of the memory features (SAM2Base._encode_new_memory, sam2_video_predictor.py
//     `maskmem_features.to(torch.bfloat16)`),
//   * best-IoU mask / token selection with the object-score gate (SAM2Base._forward_sam_heads), object-pointer mixing,
//   * hole filling of low-res mask scores (sam2/utils/misc.py fill_holes_in_mask_scores; 8-connected, area <= 8),
//   * mask prompt helpers (binarise, SAM2Base.mask_downsample 4x4/s4 conv),
//   * the per-frame label stitch of REF saber/adapters/sam2/predictor.py:289-297 (threshold > 0, nearest resize to the
//     tomogram's (H, W) as skimage.transform.resize(order=0), higher object id wins).
// Reached from REF saber/adapters/sam2/predictor.py:164-169,196-202,232-348.
#include "common.cuh"

namespace {

inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------
// RoPE. x: [rows, C] (fp32 or bf16, pitch ld_in), rows = batch * rows_per_batch. Row r of a batch entry is rotated with
// the frequencies of token r % ntok when r < n_rope, copied otherwise. cs: [ntok, C/2, 2] fp32 (cos, sin).
// Pair (2i, 2i+1) is one complex number: (a + ib)(cos + i sin).
// ------------------------------------------------------------------------------------------------
template <typename TIN>
__global__ void __launch_bounds__(256)
rope_kernel(const TIN* __restrict__ x, long long ld_in, __nv_bfloat16* __restrict__ out, long long ld_out,
            long long rows, int C, int rows_per_batch, int n_rope, int ntok, const float2* __restrict__ cs) {
  const int half = C / 2;
  const long long total = rows * half;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t % half);
    const long long r = t / half;
    const int rb = static_cast<int>(r % rows_per_batch);
    float a, b;
    if constexpr (sizeof(TIN) == 4) {
      const float2 v = *reinterpret_cast<const float2*>(x + r * ld_in + 2 * i);
      a = v.x;
      b = v.y;
    } else {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(x + r * ld_in + 2 * i);
      a = sb::bf16_lo(u);
      b = sb::bf16_hi(u);
    }
    if (rb < n_rope) {
      const float2 f = cs[static_cast<long long>(rb % ntok) * half + i];
      const float ra = a * f.x - b * f.y;
      const float rbv = a * f.y + b * f.x;
      a = ra;
      b = rbv;
    }
    *reinterpret_cast<uint32_t*>(out + r * ld_out + 2 * i) = sb::pack_bf16x2(a, b);
  }
}

// ------------------------------------------------------------------------------------------------
// Conv2d(CIN -> COUT, k3, s2, p1) + LayerNorm2d(COUT, eps) + GELU on NHWC activations; one thread per output pixel.
// in: [B, Hi, Wi, CIN] (fp32, or bf16 when IN_BF16); w: [COUT, CIN, 3, 3] fp32; out: [B, Ho, Wo, COUT] bf16.
// in_xf (CIN == 1 only): 0 none, 1 v -> sigmoid(v)*20-10, 2 v -> (v > 0 ? 1 : 0)*20-10.
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool IN_BF16>
__global__ void __launch_bounds__(128)
conv3x3s2_ln_gelu_kernel(const void* __restrict__ in_, int B, int Hi, int Wi, const float* __restrict__ w,
                         const float* __restrict__ bias, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, int in_xf, float xf_scale, float xf_bias,
                         __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sw[];  // [9][CIN][COUT] + bias/gamma/beta
  float* sbias = sw + 9 * CIN * COUT;
  float* sg = sbias + COUT;
  float* sbe = sg + COUT;
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % CIN, k = i / (COUT * CIN);
    sw[i] = w[(co * CIN + ci) * 9 + k];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    sbias[i] = bias[i];
    sg[i] = gamma[i];
    sbe[i] = beta[i];
  }
  __syncthreads();
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(t % Wo), oy = static_cast<int>((t / Wo) % Ho);
    const long long b = t / (static_cast<long long>(Wo) * Ho);
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = sbias[co];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = 2 * ox - 1 + kx;
        if (ix < 0 || ix >= Wi) continue;
        const long long pix = (b * Hi + iy) * Wi + ix;
        const float* wk = sw + (ky * 3 + kx) * CIN * COUT;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          float v;
          if (IN_BF16)
            v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in_)[pix * CIN + ci]);
          else
            v = reinterpret_cast<const float*>(in_)[pix * CIN + ci];
          if (CIN == 1) {
            if (in_xf == 1) v = (1.f / (1.f + expf(-v))) * xf_scale + xf_bias;
            if (in_xf == 2) v = (v > 0.f ? 1.f : 0.f) * xf_scale + xf_bias;
          }
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[co] = fmaf(wk[ci * COUT + co], v, acc[co]);
        }
      }
    }
    float mean = 0.f;
#pragma unroll
    for (int co = 0; co < COUT; ++co) mean += acc[co];
    mean *= (1.f / COUT);
    float var = 0.f;
#pragma unroll
    for (int co = 0; co < COUT; ++co) var += (acc[co] - mean) * (acc[co] - mean);
    const float rstd = rsqrtf(var * (1.f / COUT) + eps);
    __nv_bfloat16* dst = out + t * COUT;
#pragma unroll
    for (int co = 0; co < COUT; co += 2) {
      const float a = sb::gelu_erf((acc[co] - mean) * rstd * sg[co] + sbe[co]);
      const float c = sb::gelu_erf((acc[co + 1] - mean) * rstd * sg[co + 1] + sbe[co + 1]);
      *reinterpret_cast<uint32_t*>(dst + co) = sb::pack_bf16x2(a, c);
    }
  }
}

// im2col for Conv2d(k3, s2, p1) on NHWC bf16: in [B, Hi, Wi, C] -> cols [B*Ho*Wo, 9*C], column = (ky*3+kx)*C + c.
__global__ void __launch_bounds__(256)
im2col_3x3s2_kernel(const __nv_bfloat16* __restrict__ in, int B, int Hi, int Wi, int C,
                    __nv_bfloat16* __restrict__ cols) {
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  const int c8 = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * c8;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(t % c8);
    const int k = static_cast<int>((t / c8) % 9);
    const long long pix = t / (9ll * c8);
    const int ox = static_cast<int>(pix % Wo), oy = static_cast<int>((pix / Wo) % Ho);
    const long long b = pix / (static_cast<long long>(Wo) * Ho);
    const int iy = 2 * oy - 1 + k / 3, ix = 2 * ox - 1 + k % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi)
      v = *reinterpret_cast<const uint4*>(in + ((b * Hi + iy) * Wi + ix) * C + cc * 8);
    *reinterpret_cast<uint4*>(cols + pix * 9 * C + k * C + cc * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// CXBlock front half: depth-wise Conv2d(C, C, k7, p3, groups=C) + LayerNorm(C, eps) on NHWC fp32 [B, H, W, C];
// one warp per pixel, C = 256 (8 channels per lane). w: [C, 49] fp32. out: bf16 [B*H*W, C].
// ------------------------------------------------------------------------------------------------
// One warp per group of 4 consecutive pixels of a row (8 channels per lane): the 7 x 7 taps are staged transposed in
// shared memory ([49][256] fp32, two 16-byte loads per tap and lane instead of eight stride-49 scalar loads), every tap
// row is held in registers while the 10 input columns of the group slide past it (70 input loads per 4 pixels instead
// of 196). The per-pixel version with global scalar weight loads took 256 us per [8, 64, 64, 256] batch.
__global__ void __launch_bounds__(256)
dwconv7_ln_kernel(const float* __restrict__ in, int B, int H, int W, const float* __restrict__ w,
                  const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, __nv_bfloat16* __restrict__ out) {
  constexpr int C = 256, P = 4;
  extern __shared__ __align__(16) float dw_wt[];  // [49][256]
  for (int i = threadIdx.x; i < 49 * C; i += blockDim.x) {
    const int c = i / 49, k = i % 49;  // coalesced read of w[c][k], transposed store
    dw_wt[k * C + c] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int gpr = (W + P - 1) / P;  // pixel groups per row
  const long long total = static_cast<long long>(B) * H * gpr;
  const int c0 = lane * 8;
  float bs[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    bs[e] = bias[c0 + e];
    ga[e] = gamma[c0 + e];
    be[e] = beta[c0 + e];
  }
  for (long long g = warp; g < total; g += nwarps) {
    const int x0 = static_cast<int>(g % gpr) * P, y = static_cast<int>((g / gpr) % H);
    const long long b = g / (static_cast<long long>(gpr) * H);
    float acc[P][8];
#pragma unroll
    for (int pp = 0; pp < P; ++pp)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[pp][e] = bs[e];
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = y - 3 + ky;
      if (iy < 0 || iy >= H) continue;
      float wk[7][8];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4 a = *reinterpret_cast<const float4*>(dw_wt + (ky * 7 + kx) * C + c0);
        const float4 bq = *reinterpret_cast<const float4*>(dw_wt + (ky * 7 + kx) * C + c0 + 4);
        wk[kx][0] = a.x; wk[kx][1] = a.y; wk[kx][2] = a.z; wk[kx][3] = a.w;
        wk[kx][4] = bq.x; wk[kx][5] = bq.y; wk[kx][6] = bq.z; wk[kx][7] = bq.w;
      }
      const float* rowp = in + ((b * H + iy) * W) * C + c0;
#pragma unroll
      for (int j = 0; j < P + 6; ++j) {  // input column x0 - 3 + j feeds pixel pp with tap kx = j - pp
        const int ix = x0 - 3 + j;
        if (ix < 0 || ix >= W) continue;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(rowp + static_cast<long long>(ix) * C));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(rowp + static_cast<long long>(ix) * C + 4));
        const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int pp = 0; pp < P; ++pp) {
          const int kx = j - pp;
          if (kx >= 0 && kx < 7) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[pp][e] = fmaf(wk[kx][e], vv[e], acc[pp][e]);
          }
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < P; ++pp) {
      if (x0 + pp >= W) break;
      float s_ = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s_ += acc[pp][e];
      const float mean = sb::warp_sum(s_) * (1.f / C);
      float vs = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) vs += (acc[pp][e] - mean) * (acc[pp][e] - mean);
      const float rstd = rsqrtf(sb::warp_sum(vs) * (1.f / C) + eps);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (acc[pp][e] - mean) * rstd * ga[e] + be[e];
      const long long pix = (b * H + y) * W + x0 + pp;
      *reinterpret_cast<uint4*>(out + pix * C + c0) =
          make_uint4(sb::pack_bf16x2(o[0], o[1]), sb::pack_bf16x2(o[2], o[3]), sb::pack_bf16x2(o[4], o[5]),
                     sb::pack_bf16x2(o[6], o[7]));
    }
  }
}

// out[b, r, c] = bf16(x[b, r, c] + (score[b] > 0 ? 0 : vec[c])): occlusion embedding + the bf16 storage cast.
__global__ void __launch_bounds__(256)
add_vec_cond_kernel(const float* __restrict__ x, const float* __restrict__ score, const float* __restrict__ vec,
                    long long rows_per_batch, int C, long long total, __nv_bfloat16* __restrict__ out) {
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(t % C);
    const long long b = t / (rows_per_batch * C);
    const float add = score[b] > 0.f ? 0.f : vec[c];
    out[t] = __float2bfloat16(x[t] + add);
  }
}

// ------------------------------------------------------------------------------------------------
// SAM heads of a tracking step. masks [B,4,S,S], ious [B,4], obj [B], hs [B,Nt,256].
// multimask != 0: best = argmax(ious[b,1:4]) (first maximum), plane 1+best, token hs[b, 3+best];
// else sel (nullable) picks the plane (dynamic multimask via stability) and the token is hs[b, 2].
// low_res[b] = obj[b] > 0 ? plane : -1024 ; token[b] = chosen token ; best_out[b] = plane index.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
track_select_kernel(const float* __restrict__ masks, const float* __restrict__ ious, const float* __restrict__ obj,
                    const float* __restrict__ hs, const int* __restrict__ sel, int multimask, int Nt, int SS,
                    float* __restrict__ low_res, float* __restrict__ token, int* __restrict__ best_out) {
  const int b = blockIdx.y;
  int plane, tok;
  if (multimask) {
    const float i1 = ious[b * 4 + 1], i2 = ious[b * 4 + 2], i3 = ious[b * 4 + 3];
    int best = 0;
    float m = i1;
    if (i2 > m) { m = i2; best = 1; }
    if (i3 > m) { m = i3; best = 2; }
    plane = 1 + best;
    tok = 3 + best;
  } else {
    plane = sel ? sel[b] : 0;
    tok = 2;
  }
  const bool appear = obj[b] > 0.f;
  const float* src = masks + (static_cast<long long>(b) * 4 + plane) * SS;
  float* dst = low_res + static_cast<long long>(b) * SS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SS; i += gridDim.x * blockDim.x)
    dst[i] = appear ? src[i] : -1024.f;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < 256; c += blockDim.x)
      token[b * 256 + c] = hs[(static_cast<long long>(b) * Nt + tok) * 256 + c];
    if (threadIdx.x == 0) best_out[b] = plane;
  }
}

// ptr[b, c] = cond[b] > 0 ? ptr[b, c] : no_obj_ptr[c]   (lambda * ptr + (1 - lambda) * no_obj_ptr with lambda in {0,1})
__global__ void __launch_bounds__(256)
objptr_mix_kernel(float* __restrict__ ptr, const float* __restrict__ cond, const float* __restrict__ no_obj_ptr, int B,
                  int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  if (!(cond[b] > 0.f)) ptr[i] = no_obj_ptr[c];
}

// ------------------------------------------------------------------------------------------------
// fill_holes_in_mask_scores: background (score <= 0) 8-connected components with area <= max_area get +0.1.
// One CTA per S x S mask (S <= 256): union-find over the background pixels in shared memory (uint16 parents with
// S*S <= 65536 ... stored as int for atomics), iterated hooking until stable, then per-root areas.
// ------------------------------------------------------------------------------------------------
// parent[] is updated with atomics (resolved in L2) by other threads of the CTA: every read bypasses the
// (non-coherent) L1 with ld.global.cg, otherwise a stale "root" splits a component and its area is under-counted.
__device__ __forceinline__ int uf_find(const int* parent, int i) {
  while (true) {
    const int p = __ldcg(parent + i);
    if (p == i) return i;
    i = p;
  }
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }  // a > b : hook a under b
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(1024)
fill_holes_kernel(const float* __restrict__ in, float* __restrict__ out, int S, int max_area, int* __restrict__ ws) {
  // ws: per-mask workspace of 2*S*S ints (parents, areas) in global memory (L2-resident; 512 KB per 256^2 mask)
  const int n = S * S;
  const float* src = in + static_cast<long long>(blockIdx.x) * n;
  float* dst = out + static_cast<long long>(blockIdx.x) * n;
  int* parent = ws + static_cast<long long>(blockIdx.x) * 2 * n;
  int* area = parent + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    parent[i] = (src[i] <= 0.f) ? i : -1;
    area[i] = 0;
  }
  __syncthreads();
  // hook every background pixel to its 4 raster-preceding 8-neighbours (W, NW, N, NE)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!(src[i] <= 0.f)) continue;
    const int y = i / S, x = i % S;
    if (x > 0 && src[i - 1] <= 0.f) uf_union(parent, i, i - 1);
    if (y > 0) {
      if (src[i - S] <= 0.f) uf_union(parent, i, i - S);
      if (x > 0 && src[i - S - 1] <= 0.f) uf_union(parent, i, i - S - 1);
      if (x < S - 1 && src[i - S + 1] <= 0.f) uf_union(parent, i, i - S + 1);
    }
  }
  __syncthreads();
  // roots are final after the barrier: count areas per root, then classify every background pixel by its root's area
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!(src[i] <= 0.f)) continue;
    const int r = uf_find(parent, i);
    atomicAdd(&area[r], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = src[i];
    bool hole = false;
    if (v <= 0.f) hole = __ldcg(area + uf_find(parent, i)) <= max_area;
    dst[i] = hole ? 0.1f : v;
  }
}

// ---- shared-memory version for S % 32 == 0, S * S <= 65536 (SAM2's 256 x 256 low-res masks): 16-bit parents in shared
// memory (128 KB), background bitmap (8 KB). Horizontal runs are labelled with one ballot per 32 pixels (no unions
// inside a run), vertical / diagonal unions only where they can change connectivity (run starts / ends), areas counted
// with warp-aggregated atomics. The global-memory union-find above spends 6 ms per launch on noise-like masks (every
// hop is an L2 round trip, four unions per pixel).
__device__ __forceinline__ unsigned int uf16_find(const volatile unsigned short* parent, unsigned int i) {
  while (true) {
    const unsigned int p = parent[i];
    if (p == i) return i;
    i = p;
  }
}
__device__ __forceinline__ unsigned int atomic_min_u16(unsigned short* addr, unsigned int val) {
  unsigned int* w = reinterpret_cast<unsigned int*>(reinterpret_cast<uintptr_t>(addr) & ~static_cast<uintptr_t>(3));
  const int sh = static_cast<int>(reinterpret_cast<uintptr_t>(addr) & 2) * 8;
  unsigned int old = *reinterpret_cast<volatile unsigned int*>(w);
  while (true) {
    const unsigned int cur = (old >> sh) & 0xffffu;
    if (cur <= val) return cur;
    const unsigned int nw = (old & ~(0xffffu << sh)) | (val << sh);
    const unsigned int prev = atomicCAS(w, old, nw);
    if (prev == old) return cur;
    old = prev;
  }
}
__device__ __forceinline__ void uf16_union(unsigned short* parent, unsigned int a, unsigned int b) {
  while (true) {
    a = uf16_find(parent, a);
    b = uf16_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const unsigned int t = a;
      a = b;
      b = t;
    }  // a > b : hook a under b
    const unsigned int old = atomic_min_u16(parent + a, b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(1024)
fill_holes_smem_kernel(const float* __restrict__ in, float* __restrict__ out, int S, int max_area, int* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char fh_smem[];
  const int n = S * S;
  unsigned short* parent = reinterpret_cast<unsigned short*>(fh_smem);
  unsigned int* bits = reinterpret_cast<unsigned int*>(fh_smem + static_cast<size_t>(n) * 2);
  const float* src = in + static_cast<long long>(blockIdx.x) * n;
  float* dst = out + static_cast<long long>(blockIdx.x) * n;
  int* area = ws + static_cast<long long>(blockIdx.x) * 2 * n;
  const int lane = threadIdx.x & 31;
  auto bg = [&](int i) -> bool { return (bits[i >> 5] >> (i & 31)) & 1u; };
  // 1. background bitmap + run labels inside every 32-pixel segment (segments never straddle rows: S % 32 == 0)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const bool b = src[i] <= 0.f;
    const unsigned int m = __ballot_sync(0xffffffffu, b);
    if (lane == 0) bits[i >> 5] = m;
    const unsigned int below = (1u << lane) - 1u;
    const unsigned int zeros = ~m & below;                       // non-background pixels left of this lane
    const int start = zeros ? 32 - __clz(zeros) : 0;             // first lane of this pixel's run
    parent[i] = static_cast<unsigned short>(b ? (i - lane + start) : i);
    area[i] = 0;
  }
  __syncthreads();
  // 2. runs that continue across a 32-pixel boundary of the same row
  for (int i = threadIdx.x * 32; i < n; i += blockDim.x * 32) {
    if ((i % S) != 0 && bg(i) && bg(i - 1)) uf16_union(parent, i, i - 1);
  }
  __syncthreads();
  // 3. unions with the row above, only where they can add connectivity
  for (int i = S + threadIdx.x; i < n; i += blockDim.x) {
    if (!bg(i)) continue;
    const int x = i % S, up = i - S;
    const bool Wb = x > 0 && bg(i - 1), Eb = x < S - 1 && bg(i + 1);
    const bool Nb = bg(up), NWb = x > 0 && bg(up - 1), NEb = x < S - 1 && bg(up + 1);
    if (Nb) {
      if (!(Wb && NWb)) uf16_union(parent, i, up);  // otherwise W joined NW (its N) and NW-N are one run
    } else {
      if (NWb && !Wb) uf16_union(parent, i, up - 1);  // with W in the run, W-NW (W's N) already links
      if (NEb && !Eb) uf16_union(parent, i, up + 1);  // with E in the run, E-NE (E's N) already links
    }
  }
  __syncthreads();
  // 4. flatten + areas (one atomic per distinct root per warp)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const bool b = bg(i);
    unsigned int r = 0xffffffffu;
    if (b) {
      r = uf16_find(parent, i);
      parent[i] = static_cast<unsigned short>(r);  // still an ancestor for concurrent finds
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, r);
    if (b && lane == __ffs(peers) - 1) atomicAdd(&area[r], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = src[i];
    bool hole = false;
    if (bg(i)) hole = __ldcg(area + uf16_find(parent, i)) <= max_area;
    dst[i] = hole ? 0.1f : v;
  }
}

// out = (in >= thr ? 1 : 0) * scale + bias
__global__ void __launch_bounds__(256)
threshold_affine_kernel(const float* __restrict__ in, float thr, float scale, float bias, long long n,
                        float* __restrict__ out) {
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    out[t] = (in[t] >= thr ? 1.f : 0.f) * scale + bias;
}

// SAM2Base.mask_downsample: Conv2d(1, 1, k4, s4) on [B, S, S] fp32 -> [B, S/4, S/4]
__global__ void __launch_bounds__(256)
conv4x4s4_kernel(const float* __restrict__ in, int B, int S, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out) {
  const int T = S / 4;
  const long long total = static_cast<long long>(B) * T * T;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(t % T), oy = static_cast<int>((t / T) % T);
    const long long b = t / (static_cast<long long>(T) * T);
    const float* src = in + (b * S + oy * 4) * S + ox * 4;
    float acc = bias[0];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float4 v = *reinterpret_cast<const float4*>(src + ky * S);
      acc = fmaf(w[ky * 4 + 0], v.x, acc);
      acc = fmaf(w[ky * 4 + 1], v.y, acc);
      acc = fmaf(w[ky * 4 + 2], v.z, acc);
      acc = fmaf(w[ky * 4 + 3], v.w, acc);
    }
    out[t] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-frame label stitch (REF saber/adapters/sam2/predictor.py:289-297): for objects i = 0..N-1 in order,
// labels[y, x] = ids[i] where logits_i[sy, sx] > 0, with (sy, sx) = floor((y + 0.5) * Sv / H) (skimage order-0 resize
// of the Sv x Sv mask to (H, W); identity when H == W == Sv). logits: [N, Sv, Sv] fp32. labels: uint16 [H, W],
// updated in place (pixels no object claims keep their value).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stitch_objects_kernel(const float* __restrict__ logits, const int* __restrict__ ids, int N, int Sv, int H, int W,
                      unsigned short* __restrict__ labels) {
  const long long total = static_cast<long long>(H) * W;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(t % W), y = static_cast<int>(t / W);
    // scipy.ndimage.zoom(order=0, grid_mode=True) evaluates c = (i + 0.5) * (in / out) - 0.5 in double and rounds to
    // nearest (floor(c + 0.5)); exact ties land either side of the integer by rounding error, so the same double
    // expression (no FMA contraction) is evaluated here to stay bit-identical with the reference's skimage resize.
    const double zy = __ddiv_rn(static_cast<double>(Sv), static_cast<double>(H));
    const double zx = __ddiv_rn(static_cast<double>(Sv), static_cast<double>(W));
    const double cy = __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(y), 0.5), zy), -0.5);
    const double cx = __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(x), 0.5), zx), -0.5);
    const int sy = max(0, min(static_cast<int>(floor(__dadd_rn(cy, 0.5))), Sv - 1));
    const int sx = max(0, min(static_cast<int>(floor(__dadd_rn(cx, 0.5))), Sv - 1));
    int lab = -1;
    for (int i = 0; i < N; ++i)
      if (logits[(static_cast<long long>(i) * Sv + sy) * Sv + sx] > 0.f) lab = ids[i];
    if (lab >= 0) labels[t] = static_cast<unsigned short>(lab);
  }
}

// any[z] = 1 if slice z of a uint16 [Z, n] volume has a non-zero voxel. grid = (chunks, Z): every block ORs a chunk of
// the slice with 16-byte loads and the blocks that found something store 1 (any[] is zeroed by the launcher). One block
// of 256 threads per slice walking 2-byte loads took 3.4 ms per 928 x 960 slice stack.
__global__ void __launch_bounds__(256)
slice_any_kernel(const unsigned short* __restrict__ vol, long long n, unsigned char* __restrict__ any) {
  const unsigned short* s = vol + static_cast<long long>(blockIdx.y) * n;
  const long long per_block = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = static_cast<long long>(blockIdx.x) * per_block;
  const long long i1 = i0 + per_block < n ? i0 + per_block : n;
  unsigned int acc = 0;
  // head up to a 16-byte boundary, vector body, scalar tail
  long long i = i0 + threadIdx.x;
  const long long a0 = i0 + ((8 - ((reinterpret_cast<uintptr_t>(s + i0) >> 1) & 7)) & 7);
  for (long long j = i; j < (a0 < i1 ? a0 : i1); j += blockDim.x) acc |= s[j];
  const long long nvec = a0 < i1 ? (i1 - a0) / 8 : 0;
  const uint4* v = reinterpret_cast<const uint4*>(s + a0);
  for (long long j = threadIdx.x; j < nvec; j += blockDim.x) {
    const uint4 w = __ldg(v + j);
    acc |= w.x | w.y | w.z | w.w;
  }
  for (long long j = a0 + nvec * 8 + threadIdx.x; j < i1; j += blockDim.x) acc |= s[j];
  if (__syncthreads_or(acc != 0) && threadIdx.x == 0) any[blockIdx.y] = 1;
}

// labels[i] = 0 where labels[i] == id (presence-score filtering of one object in one slice)
__global__ void __launch_bounds__(256)
erase_label_kernel(unsigned short* __restrict__ labels, long long n, int id) {
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x)
    if (labels[t] == id) labels[t] = 0;
}

template <int CIN, int COUT, bool IN_BF16>
int launch_conv(const void* in, int B, int Hi, int Wi, const float* w, const float* bias, const float* gamma,
                const float* beta, float eps, int in_xf, float xf_scale, float xf_bias, void* out,
                cudaStream_t stream) {
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  const int smem = (9 * CIN * COUT + 3 * COUT) * static_cast<int>(sizeof(float));
  static SbPerDeviceOnce attr_once;
  if (smem > 48 * 1024 && attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(conv3x3s2_ln_gelu_kernel<CIN, COUT, IN_BF16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_once.mark();
  }
  conv3x3s2_ln_gelu_kernel<CIN, COUT, IN_BF16>
      <<<grid_for(static_cast<long long>(B) * Ho * Wo, 128, 148 * 32), 128, smem, stream>>>(
          in, B, Hi, Wi, w, bias, gamma, beta, eps, in_xf, xf_scale, xf_bias, static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

}  // namespace

extern "C" int sb_rope_apply(const void* x, long long ld_in, int in_f32, void* out, long long ld_out, long long rows,
                             int C, int rows_per_batch, int n_rope, int ntok, const float* cos_sin, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(rows > 0 && C > 0 && (C % 2) == 0 && rows_per_batch > 0 && ntok > 0 && n_rope >= 0 &&
                 n_rope <= rows_per_batch && (rows % rows_per_batch) == 0,
             "sb_rope_apply: bad arguments");
  SB_REQUIRE((ld_in % 2) == 0 && (ld_out % 2) == 0, "sb_rope_apply: pitches must be even");
  const long long total = rows * (C / 2);
  const float2* cs = reinterpret_cast<const float2*>(cos_sin);
  if (in_f32)
    rope_kernel<float><<<grid_for(total), 256, 0, stream>>>(static_cast<const float*>(x), ld_in,
                                                            static_cast<__nv_bfloat16*>(out), ld_out, rows, C,
                                                            rows_per_batch, n_rope, ntok, cs);
  else
    rope_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld_in,
                                                                    static_cast<__nv_bfloat16*>(out), ld_out, rows, C,
                                                                    rows_per_batch, n_rope, ntok, cs);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// stage: 0 = 1->4 (fp32 in, input transform in_xf), 1 = 4->16, 2 = 16->64 (bf16 in). Output bf16 NHWC.
extern "C" int sb_conv3x3s2_ln_gelu(const void* in, int stage, int B, int Hi, int Wi, const float* w, const float* bias,
                                    const float* gamma, const float* beta, float eps, int in_xf, float xf_scale,
                                    float xf_bias, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && Hi > 0 && Wi > 0 && in && w && bias && gamma && beta && out, "sb_conv3x3s2_ln_gelu: bad arguments");
  if (stage == 0) return launch_conv<1, 4, false>(in, B, Hi, Wi, w, bias, gamma, beta, eps, in_xf, xf_scale, xf_bias, out, stream);
  if (stage == 1) return launch_conv<4, 16, true>(in, B, Hi, Wi, w, bias, gamma, beta, eps, 0, 0.f, 0.f, out, stream);
  if (stage == 2) return launch_conv<16, 64, true>(in, B, Hi, Wi, w, bias, gamma, beta, eps, 0, 0.f, 0.f, out, stream);
  sb_set_error("sb_conv3x3s2_ln_gelu: stage %d not supported", stage);
  return SB_ERR_ARG;
}

extern "C" int sb_im2col_3x3s2(const void* in, int B, int Hi, int Wi, int C, void* cols, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && Hi > 0 && Wi > 0 && C > 0 && (C % 8) == 0, "sb_im2col_3x3s2: bad arguments");
  const long long total = static_cast<long long>(B) * ((Hi + 1) / 2) * ((Wi + 1) / 2) * 9 * (C / 8);
  im2col_3x3s2_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), B, Hi, Wi, C,
                                                           static_cast<__nv_bfloat16*>(cols));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_dwconv7_ln(const float* in, int B, int H, int W, int C, const float* w, const float* bias,
                             const float* gamma, const float* beta, float eps, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && H > 0 && W > 0 && C == 256, "sb_dwconv7_ln: C must be 256 (got %d)", C);
  const long long total = static_cast<long long>(B) * H * ((W + 3) / 4);  // one warp per 4-pixel group
  const int smem = 49 * 256 * 4;
  static SbPerDeviceOnce attr_once;
  if (attr_once.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(dwconv7_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_once.mark();
  }
  long long blocks = (total + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  dwconv7_ln_kernel<<<static_cast<int>(blocks), 256, smem, stream>>>(in, B, H, W, w, bias, gamma, beta, eps,
                                                                     static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_add_vec_cond(const float* x, const float* score, const float* vec, int B, long long rows_per_batch,
                               int C, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && rows_per_batch > 0 && C > 0, "sb_add_vec_cond: bad arguments");
  const long long total = static_cast<long long>(B) * rows_per_batch * C;
  add_vec_cond_kernel<<<grid_for(total), 256, 0, stream>>>(x, score, vec, rows_per_batch, C, total,
                                                           static_cast<__nv_bfloat16*>(out));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_track_select(const float* masks, const float* ious, const float* obj, const float* hs, const int* sel,
                               int multimask, int B, int Nt, int S, float* low_res, float* token, int* best,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && Nt >= 6 && S > 0, "sb_track_select: bad arguments");
  dim3 grid(16, B);
  track_select_kernel<<<grid, 256, 0, stream>>>(masks, ious, obj, hs, sel, multimask, Nt, S * S, low_res, token, best);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_objptr_mix(float* ptr, const float* cond, const float* no_obj_ptr, int B, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && C > 0, "sb_objptr_mix: bad arguments");
  objptr_mix_kernel<<<(B * C + 255) / 256, 256, 0, stream>>>(ptr, cond, no_obj_ptr, B, C);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// ws: B * 2 * S * S int32 workspace
extern "C" int sb_fill_holes(const float* in, float* out, int B, int S, int max_area, int* ws, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && S > 0 && max_area > 0 && ws, "sb_fill_holes: bad arguments");
  if ((S % 32) == 0 && S * S <= 65536) {
    const int smem = S * S * 2 + (S * S / 32) * 4;
    static SbPerDeviceOnce attr_once;
    if (attr_once.need()) {
      SB_CHECK_CUDA(cudaFuncSetAttribute(fill_holes_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 * 2 + 8192));
      attr_once.mark();
    }
    fill_holes_smem_kernel<<<B, 1024, smem, stream>>>(in, out, S, max_area, ws);
  } else {
    fill_holes_kernel<<<B, 1024, 0, stream>>>(in, out, S, max_area, ws);
  }
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_threshold_affine(const float* in, float thr, float scale, float bias, long long n, float* out,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0, "sb_threshold_affine: empty");
  threshold_affine_kernel<<<grid_for(n), 256, 0, stream>>>(in, thr, scale, bias, n, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_conv4x4s4(const float* in, int B, int S, const float* w, const float* bias, float* out,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(B > 0 && S > 0 && (S % 4) == 0, "sb_conv4x4s4: bad arguments");
  conv4x4s4_kernel<<<grid_for(static_cast<long long>(B) * (S / 4) * (S / 4)), 256, 0, stream>>>(in, B, S, w, bias, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_stitch_objects(const float* logits, const int* ids, int N, int Sv, int H, int W, void* labels,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(N > 0 && Sv > 0 && H > 0 && W > 0, "sb_stitch_objects: bad arguments");
  stitch_objects_kernel<<<grid_for(static_cast<long long>(H) * W), 256, 0, stream>>>(
      logits, ids, N, Sv, H, W, static_cast<unsigned short*>(labels));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_slice_any(const void* vol, int Z, long long n, unsigned char* any, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && n > 0, "sb_slice_any: bad arguments");
  SB_CHECK_CUDA(cudaMemsetAsync(any, 0, static_cast<size_t>(Z), stream));
  int chunks = static_cast<int>((n + 65535) / 65536);  // >= 64 K elements (128 KB) per block
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  slice_any_kernel<<<dim3(chunks, Z), 256, 0, stream>>>(static_cast<const unsigned short*>(vol), n, any);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_erase_label(void* labels, long long n, int id, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0, "sb_erase_label: empty");
  erase_label_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<unsigned short*>(labels), n, id);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
