// saber_b200 — Fourier-space resampling and band-pass for the import step before the path:
// FourierRescale3D / FourierRescale2D (REF saber/filters/downsample.py:67-129,153-204: fftn -> fftshift -> centre crop ->
// ifftshift -> ifftn) and Filter3D.apply (REF saber/filters/tomograms.py:67-184: fftn -> radial cosine band-pass -> ifftn).
//
// The reference calls torch.fft (cuFFT / pocketfft) and materialises the shifted spectrum, the crop, the un-shifted crop
// and (for the band-pass) a full D x H x W filter volume. Here a 3-D transform is three line passes of one kernel. A CTA
// stages T lines of length n in shared memory (T adjacent columns for the strided axes, so global accesses stay
// coalesced), runs a mixed-radix Stockham FFT there (radix 4 / 2 butterflies, one-output-per-thread DFT stages for the
// odd prime factors - 928 = 2^5 * 29, 200 = 2^3 * 5^2 - with roots of unity from a table computed in double precision)
// and writes the line back with the shift + crop folded into the store index, the band-pass evaluated on the fly from the
// signed frequency coordinates, and the final real part / modulus + normalisation fused into the last pass. Cropping
// passes only ever store the kept frequencies, so later passes run on the smaller array. HBM-bound: 8 B read + 8 B
// written per complex voxel and pass (4 B for the real input / output pass).
#include "common.cuh"

namespace {

constexpr int FFT_THREADS = 512;
constexpr int MAX_FACTORS = 24;

struct FftPass {
  const void* in;
  void* out;
  const float2* tw;  // tw[k] = exp(-2 pi i k / n)
  int n, m;          // transform length; values stored per line (m <= n)
  int rows_mode;     // 1: lines are contiguous rows of [lines][n]; 0: columns of [batch][n][lines]
  long long lines;
  int batch;
  int T, logT;       // lines per CTA (power of two)
  int in_real;       // input is float (imaginary part 0)
  int out_mode;      // 0 complex, 1 real part, 2 modulus
  int inverse;
  float scale;
  int crop, start;   // stored q <- spectrum[(start + (q + m / 2) % m - n / 2) mod n]
  int nfac;
  int fac[MAX_FACTORS];
  int filt;          // band-pass on store (column pass along z of a D x H x W spectrum)
  int D, H, W;
  float lp, lp_lo, lp_hi, lpd, hp, hp_lo, hp_hi, hpd;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// one side of the cosine band-pass of REF saber/filters/tomograms.py:94-137 (construct_filter), fp32 like the reference
__device__ __forceinline__ float cosine_edge(float r, float freq, float lo, float hi, float decay, bool highpass) {
  if (freq == 0.f && decay == 0.f) return 1.f;  // "skip filter": ones for both modes
  float v = r < freq ? 1.f : 0.f;
  if (decay != 0.f && r > lo && r < hi)
    v = __fadd_rn(0.5f, __fmul_rn(0.5f, cosf(__fdiv_rn(__fmul_rn(3.14159274101257324f, __fsub_rn(r, lo)), decay))));
  return highpass ? __fsub_rn(1.f, v) : v;
}

__device__ __forceinline__ float bandpass_at(int fz, int fy, int fx, const FftPass& p) {
  const float r = __fsqrt_rn(static_cast<float>(fx * fx + fy * fy + fz * fz));
  return __fmul_rn(cosine_edge(r, p.lp, p.lp_lo, p.lp_hi, p.lpd, false), cosine_edge(r, p.hp, p.hp_lo, p.hp_hi, p.hpd, true));
}

__device__ __forceinline__ int signed_freq(int k, int n) {  // coordinate of unshifted bin k on the fftshift-ed grid
  return (k + n / 2) % n - n / 2;
}

__global__ void __launch_bounds__(FFT_THREADS)
fft_lines_kernel(const __grid_constant__ FftPass p) {
  extern __shared__ float2 fft_smem[];
  const int n = p.n, T = p.T, logT = p.logT;
  float2* cur = fft_smem;
  float2* nxt = fft_smem + static_cast<size_t>(T) * n;
  const int tid = threadIdx.x;

  long long l0;
  int b = 0;
  if (p.rows_mode) {
    l0 = static_cast<long long>(blockIdx.x) * T;
  } else {
    const long long tiles = (p.lines + T - 1) / T;
    b = static_cast<int>(blockIdx.x / tiles);
    l0 = (blockIdx.x % tiles) * T;
  }
  const int nl = static_cast<int>(min(static_cast<long long>(T), p.lines - l0));
  const int total = T * n;

  // ---- load: shared layout [i][line] (line fastest: conflict-free butterflies for every stride)
  if (p.rows_mode) {
    for (int idx = tid; idx < total; idx += FFT_THREADS) {
      const int line = idx / n, i = idx - line * n;
      float2 v = make_float2(0.f, 0.f);
      if (line < nl) {
        const long long g = (l0 + line) * n + i;
        if (p.in_real) v.x = static_cast<const float*>(p.in)[g];
        else v = static_cast<const float2*>(p.in)[g];
      }
      cur[(i << logT) + line] = v;
    }
  } else {
    for (int idx = tid; idx < total; idx += FFT_THREADS) {
      const int i = idx >> logT, c = idx & (T - 1);
      float2 v = make_float2(0.f, 0.f);
      if (c < nl) {
        const long long g = (static_cast<long long>(b) * n + i) * p.lines + l0 + c;
        if (p.in_real) v.x = static_cast<const float*>(p.in)[g];
        else v = static_cast<const float2*>(p.in)[g];
      }
      cur[idx] = v;
    }
  }
  __syncthreads();

  // ---- Stockham stages
  const float sgn = p.inverse ? -1.f : 1.f;  // conjugate roots for the inverse transform
  int Ns = 1;
  for (int f = 0; f < p.nfac; ++f) {
    const int R = p.fac[f];
    if (R == 4) {
      const int q = n >> 2, tstep = n / (Ns * 4);
      for (int idx = tid; idx < (q << logT); idx += FFT_THREADS) {
        const int j = idx >> logT, line = idx & (T - 1);
        const int k = j % Ns;
        float2 v0 = cur[(j << logT) + line];
        float2 v1 = cur[((j + q) << logT) + line];
        float2 v2 = cur[((j + 2 * q) << logT) + line];
        float2 v3 = cur[((j + 3 * q) << logT) + line];
        if (k) {
          float2 w1 = __ldg(p.tw + k * tstep), w2 = __ldg(p.tw + 2 * k * tstep), w3 = __ldg(p.tw + 3 * k * tstep);
          w1.y *= sgn; w2.y *= sgn; w3.y *= sgn;
          v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
        }
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
        const float2 a3 = make_float2(sgn * d.y, -sgn * d.x);  // (v1 - v3) * (-i) forward, * (+i) inverse
        const int j0 = (j - k) * 4 + k;
        nxt[(j0 << logT) + line] = make_float2(a0.x + a2.x, a0.y + a2.y);
        nxt[((j0 + Ns) << logT) + line] = make_float2(a1.x + a3.x, a1.y + a3.y);
        nxt[((j0 + 2 * Ns) << logT) + line] = make_float2(a0.x - a2.x, a0.y - a2.y);
        nxt[((j0 + 3 * Ns) << logT) + line] = make_float2(a1.x - a3.x, a1.y - a3.y);
      }
    } else if (R == 2) {
      const int q = n >> 1, tstep = n / (Ns * 2);
      for (int idx = tid; idx < (q << logT); idx += FFT_THREADS) {
        const int j = idx >> logT, line = idx & (T - 1);
        const int k = j % Ns;
        const float2 v0 = cur[(j << logT) + line];
        float2 v1 = cur[((j + q) << logT) + line];
        if (k) {
          float2 w = __ldg(p.tw + k * tstep);
          w.y *= sgn;
          v1 = cmul(v1, w);
        }
        const int j0 = (j - k) * 2 + k;
        nxt[(j0 << logT) + line] = make_float2(v0.x + v1.x, v0.y + v1.y);
        nxt[((j0 + Ns) << logT) + line] = make_float2(v0.x - v1.x, v0.y - v1.y);
      }
    } else {
      // odd prime factor R: every thread forms one output as an R-term DFT sum; the stage twiddle and the DFT_R root
      // share one table index, advanced by a constant step per term
      const int q = n / R, span = Ns * R, tstep = n / span;
      for (int idx = tid; idx < total; idx += FFT_THREADS) {
        const int o = idx >> logT, line = idx & (T - 1);
        const int blk = o / span, rem = o - blk * span;
        const int t = rem / Ns, k = rem - t * Ns;
        const int j = blk * Ns + k;
        const int step = (k * tstep + t * q) % n;
        float2 acc = cur[(j << logT) + line];
        int ti = 0;
        for (int r = 1; r < R; ++r) {
          ti += step;
          if (ti >= n) ti -= n;
          float2 w = __ldg(p.tw + ti);
          w.y *= sgn;
          const float2 v = cmul(cur[((j + r * q) << logT) + line], w);
          acc.x += v.x;
          acc.y += v.y;
        }
        nxt[idx] = acc;
      }
    }
    __syncthreads();
    float2* tmp = cur;
    cur = nxt;
    nxt = tmp;
    Ns *= R;
  }

  // ---- store: shift + crop in the index, band-pass, normalisation, real part / modulus
  const int m = p.m;
  const int stotal = T * m;
  for (int idx = tid; idx < stotal; idx += FFT_THREADS) {
    int qo, line;
    if (p.rows_mode) {
      line = idx / m;
      qo = idx - line * m;
    } else {
      qo = idx >> logT;
      line = idx & (T - 1);
    }
    if (line >= nl) continue;
    int src = qo;
    if (p.crop) {
      src = p.start + (qo + m / 2) % m - n / 2;
      src %= n;
      if (src < 0) src += n;
    }
    float2 v = cur[(src << logT) + line];
    float s = p.scale;
    if (p.filt) {
      const long long col = l0 + line;  // = y * W + x
      const int y = static_cast<int>(col / p.W), x = static_cast<int>(col - static_cast<long long>(y) * p.W);
      s *= bandpass_at(signed_freq(qo, p.D), signed_freq(y, p.H), signed_freq(x, p.W), p);
    }
    v.x *= s;
    v.y *= s;
    const long long g = p.rows_mode ? (l0 + line) * m + qo : (static_cast<long long>(b) * m + qo) * p.lines + l0 + line;
    if (p.out_mode == 0) static_cast<float2*>(p.out)[g] = v;
    else if (p.out_mode == 1) static_cast<float*>(p.out)[g] = v.x;
    else static_cast<float*>(p.out)[g] = __fsqrt_rn(v.x * v.x + v.y * v.y);
  }
}

__global__ void __launch_bounds__(256)
fft_twiddle_kernel(int n, float2* __restrict__ tw) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double s, c;
    sincospi(-2.0 * static_cast<double>(k) / static_cast<double>(n), &s, &c);
    tw[k] = make_float2(static_cast<float>(c), static_cast<float>(s));
  }
}

// the fftshift-ed D x H x W filter volume the reference keeps as Filter3D.filter
__global__ void __launch_bounds__(256)
bandpass_volume_kernel(const FftPass p, float* __restrict__ out) {
  const long long n = static_cast<long long>(p.D) * p.H * p.W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % p.W), y = static_cast<int>((i / p.W) % p.H), z = static_cast<int>(i / (static_cast<long long>(p.W) * p.H));
    out[i] = bandpass_at(z - p.D / 2, y - p.H / 2, x - p.W / 2, p);
  }
}

int factorize(int n, int* fac) {
  int k = 0;
  while (n % 4 == 0 && k < MAX_FACTORS) { fac[k++] = 4; n /= 4; }
  while (n % 2 == 0 && k < MAX_FACTORS) { fac[k++] = 2; n /= 2; }
  for (int p = 3; n > 1 && k < MAX_FACTORS; p += 2)
    while (n % p == 0 && k < MAX_FACTORS) { fac[k++] = p; n /= p; }
  return n == 1 ? k : -1;
}

void set_bandpass(FftPass& p, int D, int H, int W, const float* bp) {
  p.D = D; p.H = H; p.W = W;
  p.lp = bp[0]; p.lp_lo = bp[1]; p.lp_hi = bp[2]; p.lpd = bp[3];
  p.hp = bp[4]; p.hp_lo = bp[5]; p.hp_hi = bp[6]; p.hpd = bp[7];
}

SbPerDeviceOnce g_fft_attr;

}  // namespace

// tw[k] = exp(-2 pi i k / n), k < n, evaluated in double precision
extern "C" int sb_fft_twiddles(int n, void* tw, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(n > 0 && tw, "sb_fft_twiddles: bad arguments");
  fft_twiddle_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, static_cast<float2*>(tw));
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// One line pass of a multi-dimensional complex FFT (see the file header).
//  rows_mode 1: `lines` contiguous rows of length n;  rows_mode 0: columns of a [batch][n][lines] array.
//  m values are stored per line: all n (crop 0) or the centred crop of the fftshift-ed spectrum starting at `start`
//  (crop 1; REF downsample.py:78-88). in_real: float input. out_mode 0 complex64, 1 real part, 2 modulus (float).
//  bandpass (host pointer, nullable; columns along z only, batch 1, lines = H * W): {lp, lp - lpd/2, lp + lpd/2, lpd,
//  hp, hp - hpd/2, hp + hpd/2, hpd} in pixels (REF tomograms.py:94-137), multiplied into the stored spectrum.
extern "C" int sb_fft_lines(const void* in, void* out, const void* tw, int n, int m, int rows_mode, long long lines,
                            int batch, int in_real, int out_mode, int inverse, float scale, int crop, int start,
                            const float* bandpass, int D, int H, int W, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(in && out && tw && n > 0 && m > 0 && m <= n && lines > 0 && batch > 0, "sb_fft_lines: bad arguments");
  SB_REQUIRE(in != out, "sb_fft_lines: the pass is out of place");
  SB_REQUIRE(crop || m == n, "sb_fft_lines: m < n needs crop");
  SB_REQUIRE(!crop || (start >= 0 && start + m <= n), "sb_fft_lines: crop window outside the spectrum");
  SB_REQUIRE(!bandpass || (!rows_mode && batch == 1 && n == D && lines == static_cast<long long>(H) * W && m == n),
             "sb_fft_lines: the band-pass is applied on the z pass of a D x H x W spectrum");
  FftPass p{};
  p.in = in; p.out = out; p.tw = static_cast<const float2*>(tw);
  p.n = n; p.m = m; p.rows_mode = rows_mode; p.lines = lines; p.batch = rows_mode ? 1 : batch;
  p.in_real = in_real; p.out_mode = out_mode; p.inverse = inverse; p.scale = scale; p.crop = crop; p.start = start;
  p.nfac = factorize(n, p.fac);
  SB_REQUIRE(p.nfac >= 0, "sb_fft_lines: n = %d has too many prime factors", n);
  if (bandpass) {
    p.filt = 1;
    set_bandpass(p, D, H, W, bandpass);
  }
  // lines per CTA: two ping-pong buffers of T * n complex values within ~192 KB of shared memory
  const size_t budget = 192 * 1024;
  int T = 32, logT = 5;
  while (T > 1 && static_cast<size_t>(2) * T * n * sizeof(float2) > budget) { T >>= 1; --logT; }
  while (T > 1 && (T >> 1) >= lines) { T >>= 1; --logT; }
  const size_t smem = static_cast<size_t>(2) * T * n * sizeof(float2);
  SB_REQUIRE(smem <= 220 * 1024, "sb_fft_lines: lines of %d values do not fit in shared memory", n);
  p.T = T; p.logT = logT;
  if (g_fft_attr.need()) {
    SB_CHECK_CUDA(cudaFuncSetAttribute(fft_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    g_fft_attr.mark();
  }
  const long long tiles = (lines + T - 1) / T * p.batch;
  SB_REQUIRE(tiles < (1ll << 31), "sb_fft_lines: too many line tiles");
  fft_lines_kernel<<<static_cast<unsigned>(tiles), FFT_THREADS, smem, stream>>>(p);
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// out[D,H,W] = the fftshift-ed band-pass volume (REF tomograms.py:67-92, Filter3D.filter)
extern "C" int sb_bandpass_volume(int D, int H, int W, const float* bandpass, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SB_REQUIRE(D > 0 && H > 0 && W > 0 && bandpass && out, "sb_bandpass_volume: bad arguments");
  FftPass p{};
  set_bandpass(p, D, H, W, bandpass);
  const long long n = static_cast<long long>(D) * H * W;
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  bandpass_volume_kernel<<<static_cast<int>(g), 256, 0, stream>>>(p, out);
  SB_CHECK_LAUNCH();
  return SB_OK;
}
