// saber_b200 — whole-tomogram bandwidth kernels on the 3-D propagation path:
//   * global min / max and the affine min-max normalisations of REF saber/adapters/preprocessing.py:72-76
//     (normalize_tomogram) and REF saber/utils/preprocessing.py:20-37 (normalize), device-side scalars (no host sync);
//   * skimage.transform.resize(order=1, mode='reflect', anti_aliasing=True) of every z-slice to the model input size
//     (REF saber/adapters/preprocessing.py:16-28), i.e. scipy.ndimage.zoom(grid_mode=True, mode='mirror') preceded by a
//     mirror-mode Gaussian when down-sampling, fused with the `2x - 1` of :59;
//   * the 15-tap z-axis Gaussian of REF saber/filters/gaussian.py:17-74 (zero padding);
//   * the z-slab mean of REF saber/utils/preprocessing.py:39-66 (project_tomogram).
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
minmax_partial_kernel(const float* __restrict__ in, long long n, float* __restrict__ partials) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  const long long n4 = n / 4;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = in4[i];
    mn = fminf(fminf(mn, v.x), fminf(fminf(v.y, v.z), v.w));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      mn = fminf(mn, in[i]);
      mx = fmaxf(mx, in[i]);
    }
  __shared__ float smn[8], smx[8];
  mn = -sb::warp_max(-mn);
  mx = sb::warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      mn = fminf(mn, smn[w]);
      mx = fmaxf(mx, smx[w]);
    }
    partials[2 * blockIdx.x] = mn;
    partials[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void minmax_final_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ out) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    mn = fminf(mn, partials[2 * i]);
    mx = fmaxf(mx, partials[2 * i + 1]);
  }
  mn = -sb::warp_max(-mn);
  mx = sb::warp_max(mx);
  if (threadIdx.x == 0) {
    out[0] = mn;
    out[1] = mx;
  }
}

// y = ((x - mn) / ((mx - mn) + eps)) * a + b, every step rounded to fp32 as numpy evaluates it (no FMA contraction)
__device__ __forceinline__ float mm_affine1(float x, float mn, float den, float a, float b) {
  return __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mn), den), a), b);
}
// VEC: 16-byte loads / stores, two per thread and iteration in flight (the scalar form ran at 3.5 TB/s of 6.5)
template <bool VEC>
__global__ void __launch_bounds__(256)
minmax_affine_kernel(const float* __restrict__ in, long long n, const float* __restrict__ mm, float eps, float a,
                     float b, float* __restrict__ out) {
  const float mn = mm[0];
  const float den = __fadd_rn(__fsub_rn(mm[1], mn), eps);
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  if (VEC) {
    const long long n4 = n >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    long long i = tid;
    for (; i + nthr < n4; i += 2 * nthr) {
      const float4 u = __ldcs(in4 + i), v = __ldcs(in4 + i + nthr);
      out4[i] = make_float4(mm_affine1(u.x, mn, den, a, b), mm_affine1(u.y, mn, den, a, b), mm_affine1(u.z, mn, den, a, b),
                            mm_affine1(u.w, mn, den, a, b));
      out4[i + nthr] = make_float4(mm_affine1(v.x, mn, den, a, b), mm_affine1(v.y, mn, den, a, b),
                                   mm_affine1(v.z, mn, den, a, b), mm_affine1(v.w, mn, den, a, b));
    }
    if (i < n4) {
      const float4 u = __ldcs(in4 + i);
      out4[i] = make_float4(mm_affine1(u.x, mn, den, a, b), mm_affine1(u.y, mn, den, a, b), mm_affine1(u.z, mn, den, a, b),
                            mm_affine1(u.w, mn, den, a, b));
    }
    for (long long j = (n4 << 2) + tid; j < n; j += nthr) out[j] = mm_affine1(in[j], mn, den, a, b);
  } else {
    for (long long i = tid; i < n; i += nthr) out[i] = mm_affine1(in[i], mn, den, a, b);
  }
}

__device__ __forceinline__ int mirror_idx(int i, int n) {
  // scipy 'mirror' (whole-sample symmetric): d c b | a b c d | c b a
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  int j = i % period;
  if (j < 0) j += period;
  return j < n ? j : period - j;
}

// scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True) of each [Hi, Wi] slice to [Ho, Wo]; coordinates and the
// linear blend are evaluated in double (scipy's spline path), the result is rounded to fp32, then out = a * v + b.
// One thread owns an output position (oy, ox) and walks the slices: the double-precision coordinate, floor, mirror
// (integer modulo) and weight terms are computed once per position instead of once per voxel — the per-voxel form was
// bound by them (0.8 TB/s) — and a warp's four loads per slice are near-contiguous runs of the two source rows.
__global__ void __launch_bounds__(256)
zoom_linear_mirror_kernel(const float* __restrict__ in, int Z, int Hi, int Wi, int Ho, int Wo, int zchunk, float a,
                          float b, float* __restrict__ out) {
  const double zy = static_cast<double>(Hi) / static_cast<double>(Ho);
  const double zx = static_cast<double>(Wi) / static_cast<double>(Wo);
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (ox >= Wo) return;
  const double cy = (static_cast<double>(oy) + 0.5) * zy - 0.5;
  const double cx = (static_cast<double>(ox) + 0.5) * zx - 0.5;
  const double fy = floor(cy), fx = floor(cx);
  const double wy1 = cy - fy, wx1 = cx - fx;
  const double wy0 = 1.0 - wy1, wx0 = 1.0 - wx1;
  const int y0 = mirror_idx(static_cast<int>(fy), Hi), y1 = mirror_idx(static_cast<int>(fy) + 1, Hi);
  const int x0 = mirror_idx(static_cast<int>(fx), Wi), x1 = mirror_idx(static_cast<int>(fx) + 1, Wi);
  const int o00 = y0 * Wi + x0, o01 = y0 * Wi + x1, o10 = y1 * Wi + x0, o11 = y1 * Wi + x1;
  const long long islice = static_cast<long long>(Hi) * Wi, oslice = static_cast<long long>(Ho) * Wo;
  const int z0 = blockIdx.z * zchunk, z1 = min(Z, z0 + zchunk);
  const float* p = in + z0 * islice;
  float* q = out + z0 * oslice + static_cast<long long>(oy) * Wo + ox;
#pragma unroll 4
  for (int z = z0; z < z1; ++z, p += islice, q += oslice) {
    const double v00 = __ldg(p + o00), v01 = __ldg(p + o01), v10 = __ldg(p + o10), v11 = __ldg(p + o11);
    const double v = wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11);
    __stcs(q, __fadd_rn(__fmul_rn(static_cast<float>(v), a), b));
  }
}

// scipy.ndimage.correlate1d(mode='mirror') with a Gaussian kernel (weights [2r+1], double) along axis 1 (rows, y) or
// 2 (columns, x) of a [Z, H, W] fp32 stack; double accumulation, fp32 storage between passes as scipy does.
__global__ void __launch_bounds__(256)
gauss1d_mirror_kernel(const float* __restrict__ in, int Z, int H, int W, int axis, const double* __restrict__ wts,
                      int r, float* __restrict__ out) {
  const long long total = static_cast<long long>(Z) * H * W;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(t % W), y = static_cast<int>((t / W) % H);
    const long long z = t / (static_cast<long long>(W) * H);
    const float* p = in + z * H * W;
    double acc = 0.0;
    for (int k = -r; k <= r; ++k) {
      const float v = axis == 1 ? p[mirror_idx(y + k, H) * W + x] : p[y * W + mirror_idx(x + k, W)];
      acc += wts[k + r] * static_cast<double>(v);
    }
    out[t] = static_cast<float>(acc);
  }
}

// out[z, i] = sum_k w[k] * in[z + k - r, i] with zero padding (F.conv1d(padding = ks/2) along z)
__global__ void __launch_bounds__(256)
gaussian_z_kernel(const float* __restrict__ in, int Z, long long plane, const float* __restrict__ w, int ks,
                  float* __restrict__ out) {
  const int r = ks / 2;
  const long long total = static_cast<long long>(Z) * plane;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = t % plane;
    const int z = static_cast<int>(t / plane);
    float acc = 0.f;
    for (int k = 0; k < ks; ++k) {
      const int zz = z + k - r;
      if (zz >= 0 && zz < Z) acc = fmaf(w[k], in[static_cast<long long>(zz) * plane + i], acc);
    }
    out[t] = acc;
  }
}
// The same as a sliding window: a thread owns four adjacent columns (one float4) of a z-chunk and keeps the last KS
// planes' values in registers, so every input voxel is read once per chunk (+ 2r halo planes) instead of KS times;
// the FMA order (k ascending, out-of-range taps skipped) is that of the kernel above, so results are bit-identical.
template <int KS>
__global__ void __launch_bounds__(128)
gaussian_z_window_kernel(const float* __restrict__ in, int Z, long long plane4, int zchunk, const float* __restrict__ w,
                         float* __restrict__ out) {
  constexpr int R = KS / 2;
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 >= plane4) return;
  const int z0 = blockIdx.y * zchunk, z1 = min(Z, z0 + zchunk);
  float wk[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) wk[k] = __ldg(w + k);
  const float4* src = reinterpret_cast<const float4*>(in) + i4;
  float4* dst = reinterpret_cast<float4*>(out) + i4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 win[KS];  // win[k] = in[z + k - R]
#pragma unroll
  for (int k = 0; k < KS - 1; ++k) {
    const int zz = z0 + k - R;
    win[k + 1] = (zz >= 0 && zz < Z) ? __ldg(src + zz * plane4) : zero;
  }
  for (int z = z0; z < z1; ++z) {
#pragma unroll
    for (int k = 0; k < KS - 1; ++k) win[k] = win[k + 1];
    const int zn = z + R;
    win[KS - 1] = zn < Z ? __ldg(src + zn * plane4) : zero;
    float4 acc = zero;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int zz = z + k - R;
      if (zz >= 0 && zz < Z) {  // warp-uniform; skipping keeps the rounding sequence of the reference kernel
        acc.x = fmaf(wk[k], win[k].x, acc.x);
        acc.y = fmaf(wk[k], win[k].y, acc.y);
        acc.z = fmaf(wk[k], win[k].z, acc.z);
        acc.w = fmaf(wk[k], win[k].w, acc.w);
      }
    }
    __stcs(dst + z * plane4, acc);
  }
}
template <int KS>
void launch_gaussian_z_window(const float* in, int Z, long long plane, const float* w, float* out, cudaStream_t stream) {
  const long long plane4 = plane / 4;
  const int bx = static_cast<int>((plane4 + 127) / 128);
  int zsplit = 1;
  while (static_cast<long long>(bx) * zsplit < 148 * 8 && Z / (zsplit * 2) >= 4 * KS) zsplit *= 2;
  const int zchunk = (Z + zsplit - 1) / zsplit;
  dim3 grid(bx, (Z + zchunk - 1) / zchunk);
  gaussian_z_window_kernel<KS><<<grid, 128, 0, stream>>>(in, Z, plane4, zchunk, w, out);
}

// out[i] = mean_{z0 <= z < z1} in[z, i]
__global__ void __launch_bounds__(256)
mean_z_kernel(const float* __restrict__ in, long long plane, int z0, int z1, float* __restrict__ out) {
  const float inv = 1.f / static_cast<float>(z1 - z0);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int z = z0; z < z1; ++z) acc += in[static_cast<long long>(z) * plane + i];
    out[i] = acc * inv;
  }
}


// ---- REF saber/filters/masks.py:230-309 + gaussian.py:76-138 (fast_3d_gaussian_smoothing, R14) and
// ---- REF saber/analysis/refine_membranes.py:100-117,274-333 (ball erosion / dilation / opening, R17) ----------------

// out[i] = (vol[i] == label) ? 1 : 0 (fp32); *count += number of matches
template <typename T>
__global__ void __launch_bounds__(256)
label_equals_kernel(const T* __restrict__ vol, long long n, unsigned int label, float* __restrict__ out,
                    unsigned long long* __restrict__ count) {
  unsigned int c = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool m = static_cast<unsigned int>(vol[i]) == label;
    out[i] = m ? 1.f : 0.f;
    c += m;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, static_cast<unsigned long long>(c));
}

// 1-D correlation with zero padding along `axis` (0 = z, 1 = y, 2 = x) of a [Z, Y, X] fp32 volume
__global__ void __launch_bounds__(256)
corr1d_zero_kernel(const float* __restrict__ in, int Z, int Y, int X, int axis, const float* __restrict__ w, int ks,
                   float* __restrict__ out) {
  const int r = ks / 2;
  const long long n = static_cast<long long>(Z) * Y * X;
  const long long stride = axis == 0 ? static_cast<long long>(Y) * X : (axis == 1 ? X : 1);
  const int len = axis == 0 ? Z : (axis == 1 ? Y : X);
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(t % X), y = static_cast<int>((t / X) % Y), z = static_cast<int>(t / (static_cast<long long>(X) * Y));
    const int p = axis == 0 ? z : (axis == 1 ? y : x);
    float acc = 0.f;
    for (int k = 0; k < ks; ++k) {
      const int q = p + k - r;
      if (q >= 0 && q < len) acc = fmaf(w[k], in[t + static_cast<long long>(k - r) * stride], acc);
    }
    out[t] = acc;
  }
}

// result[i] = label where smoothed[i] > thr (later labels overwrite earlier ones); result is uint8
__global__ void __launch_bounds__(256)
threshold_label_kernel(const float* __restrict__ sm, long long n, float thr, unsigned char label,
                       unsigned char* __restrict__ result) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (sm[i] > thr) result[i] = label;
}

// Binary erosion (op 0: every ball voxel set; outside the volume counts as 0) / dilation (op 1: any ball voxel set) with
// the radius-r ball {dz^2 + dy^2 + dx^2 <= r^2}: the exact outcome of the reference's conv3d >= sum / > 0 tests.
__global__ void __launch_bounds__(256)
morph_ball_kernel(const unsigned char* __restrict__ in, int Z, int Y, int X, int r, int op, int cube,
                  unsigned char* __restrict__ out) {
  const long long n = static_cast<long long>(Z) * Y * X;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(t % X), y = static_cast<int>((t / X) % Y), z = static_cast<int>(t / (static_cast<long long>(X) * Y));
    bool res = (op == 0);
    for (int dz = -r; dz <= r && res == (op == 0); ++dz)
      for (int dy = -r; dy <= r && res == (op == 0); ++dy)
        for (int dx = -r; dx <= r; ++dx) {
          if (!cube && dz * dz + dy * dy + dx * dx > r * r) continue;
          const int zz = z + dz, yy = y + dy, xx = x + dx;
          const bool inb = zz >= 0 && zz < Z && yy >= 0 && yy < Y && xx >= 0 && xx < X;
          const bool v = inb && in[(static_cast<long long>(zz) * Y + yy) * X + xx] != 0;
          if (op == 0 && !v) { res = false; break; }
          if (op == 1 && v) { res = true; break; }
        }
    out[t] = res ? 1 : 0;
  }
}

}  // namespace

// mm[0] = min, mm[1] = max of n floats; partials: 2 * 1024 float workspace. No host synchronisation.
extern "C" int sb_minmax(const float* in, long long n, float* mm, float* partials, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && in && mm && partials, "sb_minmax: bad arguments");
  SB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "sb_minmax: input must be 16-byte aligned");
  const int nb = grid_for(n / 4 + 1, 256, 1024);
  minmax_partial_kernel<<<nb, 256, 0, stream>>>(in, n, partials);
  SB_CHECK_LAUNCH();
  minmax_final_kernel<<<1, 32, 0, stream>>>(partials, nb, mm);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// out = ((in - mm[0]) / ((mm[1] - mm[0]) + eps)) * a + b   (mm on the device)
extern "C" int sb_minmax_affine(const float* in, long long n, const float* mm, float eps, float a, float b, float* out,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && in && mm && out, "sb_minmax_affine: bad arguments");
  if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    minmax_affine_kernel<true><<<grid_for(n / 8 + 1, 256, 148 * 8), 256, 0, stream>>>(in, n, mm, eps, a, b, out);
  else
    minmax_affine_kernel<false><<<grid_for(n), 256, 0, stream>>>(in, n, mm, eps, a, b, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_zoom_linear_mirror(const float* in, int Z, int Hi, int Wi, int Ho, int Wo, float a, float b, float* out,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "sb_zoom_linear_mirror: bad arguments");
  SB_REQUIRE(Ho <= 65535, "sb_zoom_linear_mirror: output height %d too large", Ho);
  // enough blocks for a few waves of the 148 SMs: split the slices when the plane alone is small
  const int bx = (Wo + 255) / 256;
  int zsplit = 1;
  while (static_cast<long long>(bx) * Ho * zsplit < 148 * 16 && zsplit < Z) zsplit *= 2;
  const int zchunk = (Z + zsplit - 1) / zsplit;
  dim3 grid(bx, Ho, (Z + zchunk - 1) / zchunk);
  zoom_linear_mirror_kernel<<<grid, 256, 0, stream>>>(in, Z, Hi, Wi, Ho, Wo, zchunk, a, b, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// axis: 1 = along y, 2 = along x. weights: device double [2r+1].
extern "C" int sb_gauss1d_mirror(const float* in, int Z, int H, int W, int axis, const double* weights, int r, float* out,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && H > 0 && W > 0 && (axis == 1 || axis == 2) && r >= 0 && weights, "sb_gauss1d_mirror: bad arguments");
  gauss1d_mirror_kernel<<<grid_for(static_cast<long long>(Z) * H * W), 256, 0, stream>>>(in, Z, H, W, axis, weights, r, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_gaussian_z(const float* in, int Z, long long plane, const float* w, int ks, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && plane > 0 && ks > 0 && (ks & 1) && w, "sb_gaussian_z: bad arguments");
  const bool vec = (plane % 4) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                   (plane / 4 + 127) / 128 < 2147483647LL;
  switch (vec ? ks : 0) {
#define SB_GZ_CASE(K_) case K_: launch_gaussian_z_window<K_>(in, Z, plane, w, out, stream); break;
    SB_GZ_CASE(3) SB_GZ_CASE(5) SB_GZ_CASE(7) SB_GZ_CASE(9) SB_GZ_CASE(11) SB_GZ_CASE(13) SB_GZ_CASE(15) SB_GZ_CASE(17)
    SB_GZ_CASE(19) SB_GZ_CASE(21) SB_GZ_CASE(23) SB_GZ_CASE(25)
#undef SB_GZ_CASE
    default:
      gaussian_z_kernel<<<grid_for(static_cast<long long>(Z) * plane), 256, 0, stream>>>(in, Z, plane, w, ks, out);
  }
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mean_z(const float* in, long long plane, int z0, int z1, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(plane > 0 && z1 > z0 && z0 >= 0, "sb_mean_z: bad arguments");
  mean_z_kernel<<<grid_for(plane), 256, 0, stream>>>(in, plane, z0, z1, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// out = (vol == label) as fp32; count (device u64, zeroed by the caller) += matches. elem_bytes in {1, 2, 4}.
extern "C" int sb_label_equals(const void* vol, int elem_bytes, long long n, unsigned int label, float* out,
                               unsigned long long* count, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && vol && out && count, "sb_label_equals: bad arguments");
  const int g = grid_for(n);
  if (elem_bytes == 1) label_equals_kernel<unsigned char><<<g, 256, 0, stream>>>(static_cast<const unsigned char*>(vol), n, label, out, count);
  else if (elem_bytes == 2) label_equals_kernel<unsigned short><<<g, 256, 0, stream>>>(static_cast<const unsigned short*>(vol), n, label, out, count);
  else if (elem_bytes == 4) label_equals_kernel<unsigned int><<<g, 256, 0, stream>>>(static_cast<const unsigned int*>(vol), n, label, out, count);
  else { sb_set_error("sb_label_equals: elem_bytes %d", elem_bytes); return SB_ERR_ARG; }
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_corr1d_zero(const float* in, int Z, int Y, int X, int axis, const float* w, int ks, float* out,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && Y > 0 && X > 0 && axis >= 0 && axis <= 2 && ks > 0 && (ks & 1) && w, "sb_corr1d_zero: bad arguments");
  corr1d_zero_kernel<<<grid_for(static_cast<long long>(Z) * Y * X), 256, 0, stream>>>(in, Z, Y, X, axis, w, ks, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_threshold_label(const float* sm, long long n, float thr, int label, unsigned char* result, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && sm && result, "sb_threshold_label: bad arguments");
  threshold_label_kernel<<<grid_for(n), 256, 0, stream>>>(sm, n, thr, static_cast<unsigned char>(label), result);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// op: 0 erosion, 1 dilation; in / out uint8 {0,1} [Z, Y, X]
extern "C" int sb_morph_ball(const unsigned char* in, int Z, int Y, int X, int r, int op, unsigned char* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && Y > 0 && X > 0 && r >= 0 && (op == 0 || op == 1) && in && out, "sb_morph_ball: bad arguments");
  morph_ball_kernel<<<grid_for(static_cast<long long>(Z) * Y * X), 256, 0, stream>>>(in, Z, Y, X, r, op, 0, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// The same with the full (2r+1)^3 cube as the structuring element: scipy.ndimage.binary_erosion(structure=np.ones((3,3,3)))
// of REF saber/analysis/refine_membranes.py:172 (border_value 0 = zero padding).
extern "C" int sb_morph_cube(const unsigned char* in, int Z, int Y, int X, int r, int op, unsigned char* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(Z > 0 && Y > 0 && X > 0 && r >= 0 && (op == 0 || op == 1) && in && out, "sb_morph_cube: bad arguments");
  morph_ball_kernel<<<grid_for(static_cast<long long>(Z) * Y * X), 256, 0, stream>>>(in, Z, Y, X, r, op, 1, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
