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
  unsigned mag_ns[MAX_FACTORS];    // exact division by Ns / by Ns * R of the odd-radix stages (see fast_div)
  unsigned mag_q[MAX_FACTORS];     // exact division by n / R
  unsigned mag_n, mag_m;
  int filt;          // band-pass on store (column pass along z of a D x H x W spectrum)
  int D, H, W;
  float lp, lp_lo, lp_hi, lpd, hp, hp_lo, hp_hi, hpd;
};

// x / d for 0 <= x < 65536, 1 <= d < 65536 with magic = 2^32 / d + 1 (0 encodes d == 1): exact because x * d < 2^32
__device__ __forceinline__ unsigned fast_div(unsigned x, unsigned magic) {
  return magic ? __umulhi(x, magic) : x;
}

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

// ---- radix-R butterflies on registers (forward: sgn = +1, inverse: sgn = -1) ----------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * (-i) forward, a * (+i) inverse
__device__ __forceinline__ float2 crot(float2 a, float sgn) { return make_float2(sgn * a.y, -sgn * a.x); }

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R], float sgn);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&v)[2], float) {
  const float2 a = v[0];
  v[0] = cadd(a, v[1]);
  v[1] = csub(a, v[1]);
}

template <>
__device__ __forceinline__ void butterfly<3>(float2 (&v)[3], float sgn) {
  const float2 t = cadd(v[1], v[2]);
  const float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
  const float2 d = csub(v[1], v[2]);
  const float2 r = crot(make_float2(0.86602540378443865f * d.x, 0.86602540378443865f * d.y), sgn);
  v[0] = cadd(v[0], t);
  v[1] = cadd(m, r);
  v[2] = csub(m, r);
}

template <>
__device__ __forceinline__ void butterfly<4>(float2 (&v)[4], float sgn) {
  const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]), a2 = cadd(v[1], v[3]);
  const float2 a3 = crot(csub(v[1], v[3]), sgn);
  v[0] = cadd(a0, a2);
  v[1] = cadd(a1, a3);
  v[2] = csub(a0, a2);
  v[3] = csub(a1, a3);
}

template <>
__device__ __forceinline__ void butterfly<5>(float2 (&v)[5], float sgn) {
  constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
  const float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
  const float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
  const float2 n1 = crot(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y), sgn);
  const float2 n2 = crot(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y), sgn);
  v[0] = cadd(v[0], cadd(a1, a2));
  v[1] = cadd(m1, n1);
  v[4] = csub(m1, n1);
  v[2] = cadd(m2, n2);
  v[3] = csub(m2, n2);
}

// One Stockham stage of radix R: thread (line, j) loads its R inputs, applies the stage twiddles, runs the butterfly in
// registers and scatters the outputs. `line` is fixed per thread (folded into the base pointers), j strides by
// FFT_THREADS / T. LOGT and "Ns is a power of two" are compile-time so that every shared-memory address is base + shift.
template <int R, int LOGT, bool POW2>
__device__ __forceinline__ void radix_stage(const float2* __restrict__ cur, float2* __restrict__ nxt,
                                            const float2* __restrict__ tw, unsigned n, unsigned Ns, unsigned mag_ns,
                                            unsigned jbase, float sgn) {
  constexpr unsigned jstride = FFT_THREADS >> LOGT;
  const unsigned q = n / R, tstep = q / Ns;
  if (Ns == 1) {  // first stage: no twiddles, outputs at j * R + r
#pragma unroll 2
    for (unsigned j = jbase; j < q; j += jstride) {
      float2 v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = cur[(j + r * q) << LOGT];
      butterfly<R>(v, sgn);
#pragma unroll
      for (int r = 0; r < R; ++r) nxt[(j * R + r) << LOGT] = v[r];
    }
    return;
  }
#pragma unroll 2
  for (unsigned j = jbase; j < q; j += jstride) {
    const unsigned k = POW2 ? (j & (Ns - 1)) : j - fast_div(j, mag_ns) * Ns;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cur[(j + r * q) << LOGT];
    const unsigned kt = k * tstep;  // tw[0] = 1: no branch for k == 0
#pragma unroll
    for (int r = 1; r < R; ++r) {
      float2 w = tw[r * kt];
      w.y *= sgn;
      v[r] = cmul(v[r], w);
    }
    butterfly<R>(v, sgn);
    const unsigned j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) nxt[(j0 + r * Ns) << LOGT] = v[r];
  }
}

template <int R, int LOGT>
__device__ __forceinline__ void radix_stage_any(const float2* cur, float2* nxt, const float2* tw, unsigned n, unsigned Ns,
                                                unsigned mag_ns, bool pow2, unsigned jbase, float sgn) {
  if (pow2) radix_stage<R, LOGT, true>(cur, nxt, tw, n, Ns, mag_ns, jbase, sgn);
  else radix_stage<R, LOGT, false>(cur, nxt, tw, n, Ns, mag_ns, jbase, sgn);
}

// BIG: the length has a prime factor other than 2, 3, 5 (the general odd-prime stage needs more registers; lengths
// without one get the leaner kernel and three resident CTAs per SM)
template <int LOGT, bool BIG>
__global__ void __launch_bounds__(FFT_THREADS, BIG ? 2 : 3)
fft_lines_kernel(const __grid_constant__ FftPass p) {
  extern __shared__ float2 fft_smem[];
  constexpr unsigned T = 1u << LOGT;
  constexpr int logT = LOGT;
  const unsigned n = p.n;
  float2* cur = fft_smem;
  float2* nxt = fft_smem + static_cast<size_t>(T) * n;
  float2* tw = fft_smem + static_cast<size_t>(2) * T * n;  // the n roots of unity, staged once per CTA
  const unsigned tid = threadIdx.x;

  long long l0;
  unsigned b = 0;
  if (p.rows_mode) {
    l0 = static_cast<long long>(blockIdx.x) * T;
  } else {
    const unsigned tiles = static_cast<unsigned>((p.lines + T - 1) / T);
    b = blockIdx.x / tiles;
    l0 = static_cast<long long>(blockIdx.x - b * tiles) * T;
  }
  const unsigned nl = static_cast<unsigned>(min(static_cast<long long>(T), p.lines - l0));
  const unsigned total = T * n;

  for (unsigned i = tid; i < n; i += FFT_THREADS) tw[i] = __ldg(p.tw + i);
  // ---- load: shared layout [i][line] (line fastest: conflict-free butterfly reads for every stride)
  if (p.rows_mode) {
    // the tile's rows are contiguous in memory: element idx of the tile is element (line = idx / n, i = idx % n)
    const unsigned valid = nl * n;
    if (p.in_real) {
      const float* src = static_cast<const float*>(p.in) + l0 * n;
#pragma unroll 4
      for (unsigned idx = tid; idx < total; idx += FFT_THREADS) {
        const unsigned line = fast_div(idx, p.mag_n), i = idx - line * n;
        cur[(i << logT) + line] = make_float2(idx < valid ? src[idx] : 0.f, 0.f);
      }
    } else {
      const float2* src = static_cast<const float2*>(p.in) + l0 * n;
#pragma unroll 4
      for (unsigned idx = tid; idx < total; idx += FFT_THREADS) {
        const unsigned line = fast_div(idx, p.mag_n), i = idx - line * n;
        cur[(i << logT) + line] = idx < valid ? src[idx] : make_float2(0.f, 0.f);
      }
    }
  } else {
    const size_t base = static_cast<size_t>(b) * n * p.lines + l0;
    const size_t pitch = static_cast<size_t>(p.lines);
    if (p.in_real) {
      const float* src = static_cast<const float*>(p.in) + base;
#pragma unroll 4
      for (unsigned idx = tid; idx < total; idx += FFT_THREADS) {
        const unsigned i = idx >> logT, c = idx & (T - 1);
        cur[idx] = make_float2(c < nl ? src[i * pitch + c] : 0.f, 0.f);
      }
    } else {
      const float2* src = static_cast<const float2*>(p.in) + base;
#pragma unroll 4
      for (unsigned idx = tid; idx < total; idx += FFT_THREADS) {
        const unsigned i = idx >> logT, c = idx & (T - 1);
        cur[idx] = c < nl ? src[i * pitch + c] : make_float2(0.f, 0.f);
      }
    }
  }
  __syncthreads();

  // ---- Stockham stages
  const float sgn = p.inverse ? -1.f : 1.f;  // conjugate roots for the inverse transform
  const unsigned line = tid & (T - 1), jbase = tid >> logT, jstride = FFT_THREADS >> logT;
  unsigned Ns = 1;
  for (int f = 0; f < p.nfac; ++f) {
    const unsigned R = p.fac[f];
    const bool pow2 = (Ns & (Ns - 1)) == 0;
    if (R == 4) {
      radix_stage<4, LOGT, true>(cur + line, nxt + line, tw, n, Ns, 0u, jbase, sgn);  // radix 4 / 2 come first: Ns = 2^x
    } else if (R == 2) {
      radix_stage<2, LOGT, true>(cur + line, nxt + line, tw, n, Ns, 0u, jbase, sgn);
    } else if (R == 3) {
      radix_stage_any<3, LOGT>(cur + line, nxt + line, tw, n, Ns, p.mag_ns[f], pow2, jbase, sgn);
    } else if (R == 5) {
      radix_stage_any<5, LOGT>(cur + line, nxt + line, tw, n, Ns, p.mag_ns[f], pow2, jbase, sgn);
    } else if (BIG) {
      // Any other odd prime R (928 = 2^5 x 29): stage twiddles applied in place first, then the DFT_R of every j as
      // output PAIRS (t, R - t): with a_r = v_r + v_{R-r}, b_r = v_r - v_{R-r} (r = 1 .. h = (R-1)/2)
      //   out_t = v_0 + sum_r a_r cos(2 pi r t / R) -+ i sum_r b_r sin(2 pi r t / R),   out_{R-t} = its mirror,
      // so each product is used for two outputs. cos / sin come from the staged table: W_R^x = tw[x * n / R].
      const unsigned q = n / R, tstep = q / Ns, h = (R - 1) / 2, mag_ns = p.mag_ns[f];
      for (unsigned idx = tid; idx < total; idx += FFT_THREADS) {
        const unsigned i = idx >> logT;
        const unsigned r = fast_div(i, p.mag_q[f]), j = i - r * q;
        const unsigned k = pow2 ? (j & (Ns - 1)) : j - static_cast<unsigned>(fast_div(j, mag_ns)) * Ns;
        if (r && k) {
          float2 w = tw[r * k * tstep];
          w.y *= sgn;
          cur[idx] = cmul(cur[idx], w);
        }
      }
      __syncthreads();
      // t = 0: the plain sum
      for (unsigned j = jbase; j < q; j += jstride) {
        const unsigned k = pow2 ? (j & (Ns - 1)) : j - fast_div(j, mag_ns) * Ns;
        float2 A = cur[(j << logT) + line];
        for (unsigned r = 1; r < R; ++r) A = cadd(A, cur[((j + r * q) << logT) + line]);
        nxt[(((j - k) * R + k) << logT) + line] = A;
      }
      // t = 1 .. h in groups of four output pairs per thread: the two loads and the sum / difference of a term are
      // shared by the group, each pair adds one root lookup and four FMAs
      constexpr unsigned G = 4;
      const unsigned groups = (h + G - 1) / G, items = q * groups;
      for (unsigned it = jbase; it < items; it += jstride) {
        const unsigned tg = fast_div(it, p.mag_q[f]), j = it - tg * q;
        const unsigned k = pow2 ? (j & (Ns - 1)) : j - fast_div(j, mag_ns) * Ns;
        const unsigned j0 = (j - k) * R + k;
        unsigned t[G], x[G];
        float2 A[G], B[G];
        const float2 v0 = cur[(j << logT) + line];
#pragma unroll
        for (unsigned g = 0; g < G; ++g) {
          t[g] = 1 + tg * G + g;
          if (t[g] > h) t[g] = 0;  // padding lane of the last group: multiplies by tw[0] = 1, never stored
          x[g] = 0;
          A[g] = v0;
          B[g] = make_float2(0.f, 0.f);
        }
        for (unsigned r = 1; r <= h; ++r) {
          const float2 u = cur[((j + r * q) << logT) + line], z = cur[((j + (R - r) * q) << logT) + line];
          const float2 a = cadd(u, z), d = csub(u, z);
#pragma unroll
          for (unsigned g = 0; g < G; ++g) {
            x[g] += t[g];  // r * t mod R
            if (x[g] >= R) x[g] -= R;
            const float2 w = tw[x[g] * q];  // (cos, -sin) of 2 pi x / R
            A[g].x += a.x * w.x;
            A[g].y += a.y * w.x;
            B[g].x += d.x * w.y;  // accumulates -sin * b
            B[g].y += d.y * w.y;
          }
        }
#pragma unroll
        for (unsigned g = 0; g < G; ++g) {
          if (t[g] == 0) continue;
          // forward: out_t = A - i S, S = sum b sin = -B  ->  A + i B;  inverse: A - i B
          const float2 iB = make_float2(-sgn * B[g].y, sgn * B[g].x);
          nxt[((j0 + t[g] * Ns) << logT) + line] = cadd(A[g], iB);
          nxt[((j0 + (R - t[g]) * Ns) << logT) + line] = csub(A[g], iB);
        }
      }
    }
    __syncthreads();
    float2* tmp = cur;
    cur = nxt;
    nxt = tmp;
    Ns *= R;
  }

  // ---- store: shift + crop in the index, band-pass, normalisation, real part / modulus
  const unsigned m = p.m;
  const unsigned stotal = T * m;
  const size_t out_base = p.rows_mode ? static_cast<size_t>(l0) * m : static_cast<size_t>(b) * m * p.lines + l0;
  const size_t pitch = static_cast<size_t>(p.lines);
#pragma unroll 2
  for (unsigned idx = tid; idx < stotal; idx += FFT_THREADS) {
    unsigned qo, ln;
    if (p.rows_mode) {
      ln = fast_div(idx, p.mag_m);
      qo = idx - ln * m;
    } else {
      qo = idx >> logT;
      ln = idx & (T - 1);
    }
    if (ln >= nl) continue;
    unsigned src = qo;
    if (p.crop) {  // spectrum bin (start + (qo + m/2) mod m - n/2) mod n, without divisions
      unsigned sh = qo + m / 2;
      if (sh >= m) sh -= m;
      int sb = static_cast<int>(p.start + sh) - static_cast<int>(n / 2);
      if (sb < 0) sb += n;
      src = static_cast<unsigned>(sb);
    }
    float2 v = cur[(src << logT) + ln];
    float s = p.scale;
    if (p.filt) {
      const long long col = l0 + ln;  // = y * W + x
      const int y = static_cast<int>(col / p.W), x = static_cast<int>(col - static_cast<long long>(y) * p.W);
      s *= bandpass_at(signed_freq(qo, p.D), signed_freq(y, p.H), signed_freq(x, p.W), p);
    }
    v.x *= s;
    v.y *= s;
    const size_t g = out_base + (p.rows_mode ? static_cast<size_t>(idx) : qo * pitch + ln);
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
  auto magic = [](int d) -> unsigned { return d <= 1 ? 0u : static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(d) + 1ull); };
  {
    int Ns = 1;
    for (int f = 0; f < p.nfac; ++f) {
      p.mag_ns[f] = magic(Ns);
      p.mag_q[f] = magic(n / p.fac[f]);
      Ns *= p.fac[f];
    }
    p.mag_n = magic(n);
    p.mag_m = magic(m);
  }
  // lines per CTA: two ping-pong buffers of T * n complex values. ~72 KB per CTA keeps three CTAs (48 warps) on an SM so
  // the loads of one overlap the butterflies of another; at least 4 lines (32 B sectors on the strided axes) as long as
  // they fit at all.
  int T = 32, logT = 5;
  auto bytes = [&](int t) { return (static_cast<size_t>(2) * t + 1) * n * sizeof(float2); };  // + the root table
  while (T > 4 && bytes(T) > 74 * 1024) { T >>= 1; --logT; }
  while (T > 1 && bytes(T) > 220 * 1024) { T >>= 1; --logT; }
  while (T > 1 && (T >> 1) >= lines) { T >>= 1; --logT; }
  const size_t smem = bytes(T);
  SB_REQUIRE(smem <= 220 * 1024 && T * n < 65536, "sb_fft_lines: lines of %d values do not fit in shared memory", n);
  p.T = T; p.logT = logT;
  bool big = false;
  for (int f = 0; f < p.nfac; ++f) big = big || p.fac[f] > 5;
  if (g_fft_attr.need()) {
    auto prep = [](const void* fn) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      return e != cudaSuccess ? e : cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    };
    const void* fns[] = {(const void*)fft_lines_kernel<0, false>, (const void*)fft_lines_kernel<1, false>,
                         (const void*)fft_lines_kernel<2, false>, (const void*)fft_lines_kernel<3, false>,
                         (const void*)fft_lines_kernel<4, false>, (const void*)fft_lines_kernel<5, false>,
                         (const void*)fft_lines_kernel<0, true>,  (const void*)fft_lines_kernel<1, true>,
                         (const void*)fft_lines_kernel<2, true>,  (const void*)fft_lines_kernel<3, true>,
                         (const void*)fft_lines_kernel<4, true>,  (const void*)fft_lines_kernel<5, true>};
    for (const void* fn : fns) SB_CHECK_CUDA(prep(fn));
    g_fft_attr.mark();
  }
  const long long tiles = (lines + T - 1) / T * p.batch;
  SB_REQUIRE(tiles < (1ll << 31), "sb_fft_lines: too many line tiles");
  const unsigned grid = static_cast<unsigned>(tiles);
#define SB_FFT_LAUNCH(L)                                                                  \
  case L:                                                                                 \
    if (big) fft_lines_kernel<L, true><<<grid, FFT_THREADS, smem, stream>>>(p);           \
    else fft_lines_kernel<L, false><<<grid, FFT_THREADS, smem, stream>>>(p);              \
    break;
  switch (logT) {
    SB_FFT_LAUNCH(0) SB_FFT_LAUNCH(1) SB_FFT_LAUNCH(2) SB_FFT_LAUNCH(3) SB_FFT_LAUNCH(4)
    default:
      if (big) fft_lines_kernel<5, true><<<grid, FFT_THREADS, smem, stream>>>(p);
      else fft_lines_kernel<5, false><<<grid, FFT_THREADS, smem, stream>>>(p);
      break;
  }
#undef SB_FFT_LAUNCH
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
