// Fused log-mel front-end: zero centre-padding, framing (n_fft 2048, hop 512), periodic Hann window,
// real FFT, squared magnitude, sparse slaney mel filterbank, 10*log10(max(., 1e-10)) -- one kernel, one
// pass over the waveform (torchaudio MelSpectrogram + AmplitudeToDB as configured at
// adaptive_attack_eval.py:83-85).  HBM traffic is 4*L bytes in + 4*n_mels*frames bytes out per clip; the
// filterbank (<= 2 non-zeros per bin) and twiddles stay in shared memory / L2.
//
// One CTA handles one pair of adjacent frames of one clip: the two real frames are packed as the real and
// imaginary parts of ONE 2048-point complex FFT (radix-2 DIT in shared memory, fp32, host-computed
// double-precision twiddles) and separated afterwards with the conjugate-symmetry identities.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ap {

constexpr int kNfft = 2048;
constexpr int kHop = 512;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kMelThreads = 256;
constexpr int kMaxFbNnz = 2304;  // filterbank weights staged in shared memory
constexpr int kFftPad = kNfft + kNfft / 32;  // one float2 of skew per 32 elements: de-conflicts the bit-reversed scatter

__device__ __forceinline__ int fpad(int i) { return i + (i >> 5); }

// exp(-i * 2*pi*k/2048) for the butterflies.  The shared twiddle table is only read with unit stride (window);
// the strided butterfly reads (up to 32-way bank conflicts) are replaced by two MUFU ops (abs. error 2^-21).
__device__ __forceinline__ float2 twiddle(int k) {
  float sn, cs;
  __sincosf(-3.14159265358979323846f * static_cast<float>(k) * (1.0f / 1024.0f), &sn, &cs);
  return make_float2(cs, sn);
}

struct MelArgs {
  const float* x;        // [B][L]
  float* out;            // [B][n_mels][n_frames]
  const float2* tw;      // [1024] exp(-2*pi*i*k/2048)
  const int* fb_start;   // [n_mels] first bin with non-zero weight
  const int* fb_len;     // [n_mels] number of consecutive non-zero bins
  const int* fb_off;     // [n_mels] offset of that run in fb_w
  const float* fb_w;     // concatenated non-zero weights
  int B, L, n_frames, n_mels, fb_nnz;
};

// In-place 2048-point radix-2 DIT FFT on bit-reversed input held in (padded) shared memory.
// sign = +1: forward (e^{-i...}); sign = -1: unnormalised inverse (conjugated twiddles).
__device__ __forceinline__ void fft2048_inplace(float2* zs, int tid, float sign) {
#pragma unroll 1
  for (int s = 0; s < 11; ++s) {
    const int half = 1 << s;
    for (int j = tid; j < kNfft / 2; j += kMelThreads) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos;
      const int i1 = i0 + half;
      float2 t = twiddle(pos << (10 - s));
      t.y *= sign;
      const float2 u = zs[fpad(i0)], v = zs[fpad(i1)];
      const float2 vt = make_float2(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x);
      zs[fpad(i0)] = make_float2(u.x + vt.x, u.y + vt.y);
      zs[fpad(i1)] = make_float2(u.x - vt.x, u.y - vt.y);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Forward kernel.  2048 = 16 x 16 x 8: every one of 128 threads keeps 16 complex points in registers, so the
// transform is three register-resident passes (16-point, 16-point, 2 x 8-point DFTs) with two shared-memory
// exchanges, instead of eleven __syncthreads-separated radix-2 stages (r01: 41.8 us for 64 clips, latency-bound).
//   n = 128 n1 + n2,  n2 = 8 m1 + m2;   k = k1 + 16 j1 + 256 j2
//   pass 1 (thread n2):       A[k1]   = sum_n1 z[128 n1 + n2] W16^(n1 k1);   B = A * W2048^(n2 k1)
//   pass 2 (thread k1, m2):   C[j1]   = sum_m1 B[k1][8 m1 + m2] W16^(m1 j1); D = C * W128^(m2 j1)
//   pass 3 (thread k1 + 16 j1, two of them per thread):  Z[k] = sum_m2 D[k1][j1][m2] W8^(m2 j2)
// The CTA is persistent over (clip, frame pair) items; the filterbank weights are staged in shared memory once per
// CTA with cp.async while the first transform runs.
// ---------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 128;
constexpr int kRowPad = 136;  // float2 stride of a k1 row in the exchange buffer: 136 = 8 (mod 16) keeps both sides of
                              // the exchanges at the minimum number of shared-memory wavefronts

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// exp(-2 pi i m / M), 0 <= m < M, M a power of two: two MUFU ops on an angle in (-pi, 0] (abs. error 2^-21)
template <int M>
__device__ __forceinline__ float2 unit_root(int m) {
  const bool flip = m >= M / 2;
  if (flip) m -= M / 2;
  float sn, cs;
  __sincosf(-6.28318530717958647692f / M * static_cast<float>(m), &sn, &cs);
  return flip ? make_float2(-cs, -sn) : make_float2(cs, sn);
}

// forward 4-point DFT in place
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = cmul_mi(csub(b, d));
  a = cadd(s0, s2);
  b = cadd(s1, s3);
  c = csub(s0, s2);
  d = csub(s1, s3);
}

// forward 16-point DFT: v[n] -> v[k], natural order in and out (n = 4a + b, k = c + 4d)
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float kC1 = 0.92387953251128675613f, kS1 = 0.38268343236508977173f, kR = 0.70710678118654752440f;
  // W16^m = exp(-2 pi i m / 16)
  const float2 w[10] = {{1.f, 0.f}, {kC1, -kS1}, {kR, -kR}, {kS1, -kC1}, {0.f, -1.f}, {-kS1, -kC1}, {-kR, -kR},
                        {-kC1, -kS1}, {-1.f, 0.f}, {-kC1, kS1}};
#pragma unroll
  for (int b = 0; b < 4; ++b) {  // over a (stride 4): y[b][c]
    dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
#pragma unroll
    for (int c = 1; c < 4; ++c)
      if (b) v[4 * c + b] = cmul(v[4 * c + b], w[b * c]);  // W16^(b c), b c <= 9
  }
  // now v[4c + b] = y[b][c] * W16^(bc); X[c + 4d] = sum_b (.) W4^(bd)
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  // v[4c + d] = X[c + 4d]  ->  natural order: transpose the 4 x 4
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) {
      const float2 t = v[4 * c + d];
      v[4 * c + d] = v[4 * d + c];
      v[4 * d + c] = t;
    }
}

// forward 8-point DFT, natural order in and out (n = 2a + b, k = c + 4d)
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  constexpr float kR = 0.70710678118654752440f;
  dft4(v[0], v[2], v[4], v[6]);  // even samples: E[c]
  dft4(v[1], v[3], v[5], v[7]);  // odd samples:  O[c]
  v[3] = cmul(v[3], make_float2(kR, -kR));   // O[1] * W8^1
  v[5] = cmul_mi(v[5]);                      // O[2] * W8^2
  v[7] = cmul(v[7], make_float2(-kR, -kR));  // O[3] * W8^3
  float2 o[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    o[c] = cadd(v[2 * c], v[2 * c + 1]);
    o[c + 4] = csub(v[2 * c], v[2 * c + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = o[k];
}

__global__ void __launch_bounds__(kFwdThreads, 7) logmel_kernel(const MelArgs a) {
  __shared__ __align__(16) float2 zs[16 * kRowPad];  // exchange buffer, then Z[k] in natural order, then the two power spectra
  __shared__ __align__(16) float fbw[kMaxFbNnz];

  const int tid = threadIdx.x;
  // filterbank weights -> shared memory, asynchronously (consumed after the first transform)
  for (int i = tid * 4; i < a.fb_nnz; i += kFwdThreads * 4) {
    if (i + 4 <= a.fb_nnz && (reinterpret_cast<uintptr_t>(a.fb_w) & 15) == 0) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(fbw + i))),
                   "l"(a.fb_w + i)
                   : "memory");
    } else {
      for (int j = i; j < a.fb_nnz && j < i + 4; ++j) fbw[j] = a.fb_w[j];
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // per-thread constants: window phase and the two twiddle families
  float wsn, wcs;
  sincospif(static_cast<float>(tid) * (1.0f / 1024.0f), &wsn, &wcs);  // angle 2 pi n2 / 2048
  const int k1b = tid >> 3, m2b = tid & 7;

  const int pairs = (a.n_frames + 1) >> 1;
  const int items = a.B * pairs;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / pairs;
    const int f0 = (item - b * pairs) * 2;
    const bool has2 = (f0 + 1) < a.n_frames;
    const float* x = a.x + static_cast<size_t>(b) * a.L;

    // ---- load: frame f0 covers samples [f0*hop - 1024, +2048), frame f0+1 the same shifted by hop = 4 * 128, so
    // thread n2 needs x[base + 128 q + n2] for q = 0..19 (coalesced); outside [0, L) is the zero padding ----
    float xs[20];
    const int first = f0 * kHop - kNfft / 2, base = first + tid;
    if (first >= 0 && first + 20 * 128 <= a.L && has2) {  // interior pair (most of them): no bounds checks
#pragma unroll
      for (int q = 0; q < 20; ++q) xs[q] = __ldg(x + base + 128 * q);
    } else {
#pragma unroll
      for (int q = 0; q < 20; ++q) {
        const int i = base + 128 * q;
        xs[q] = (i >= 0 && i < a.L && (q < 16 || has2)) ? __ldg(x + i) : 0.f;
      }
    }
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      // periodic Hann: 0.5 - 0.5 cos(2 pi (128 n1 + n2) / 2048), cos(a + b) with a = 2 pi n1 / 16 a compile-time constant
      constexpr float kC1 = 0.92387953251128675613f, kS1 = 0.38268343236508977173f, kR = 0.70710678118654752440f;
      constexpr float kCos[16] = {1.f, kC1, kR, kS1, 0.f, -kS1, -kR, -kC1, -1.f, -kC1, -kR, -kS1, 0.f, kS1, kR, kC1};
      constexpr float kSin[16] = {0.f, kS1, kR, kC1, 1.f, kC1, kR, kS1, 0.f, -kS1, -kR, -kC1, -1.f, -kC1, -kR, -kS1};
      const float w = 0.5f - 0.5f * (kCos[n1] * wcs - kSin[n1] * wsn);
      v[n1] = make_float2(xs[n1] * w, xs[n1 + 4] * w);
    }
    // ---- pass 1 ----
    dft16(v);
    __syncthreads();  // the previous item's mel reduction has finished reading zs
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 t = k1 ? cmul(v[k1], unit_root<2048>((tid * k1) & 2047)) : v[0];
      zs[k1 * kRowPad + tid] = t;
    }
    __syncthreads();
    // ---- pass 2: thread (k1b, m2b) ----
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) v[m1] = zs[k1b * kRowPad + 8 * m1 + m2b];
    dft16(v);
    __syncthreads();
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) {
      const float2 t = (j1 && m2b) ? cmul(v[j1], unit_root<128>(m2b * j1)) : v[j1];
      zs[k1b * kRowPad + 8 * j1 + m2b] = t;
    }
    __syncthreads();
    // ---- pass 3: rows r = tid and tid + 128 with r = k1 + 16 j1 (k1 fastest: natural-order stores are contiguous) ----
    float2 u[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = tid + 128 * h, k1 = r & 15, j1 = r >> 4;
      const float4* src = reinterpret_cast<const float4*>(zs + k1 * kRowPad + 8 * j1);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = src[q];
        u[h][2 * q] = make_float2(t.x, t.y);
        u[h][2 * q + 1] = make_float2(t.z, t.w);
      }
      dft8(u[h]);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) zs[tid + 128 * h + 256 * j2] = u[h][j2];  // Z[k], k = k1 + 16 j1 + 256 j2
    __syncthreads();
    // ---- Z = F1 + i F2  ->  F1[k] = (Z[k] + conj(Z[N-k]))/2,  F2[k] = (Z[k] - conj(Z[N-k]))/(2i); power spectra of
    // both frames kept as one float2 per bin, in place (bin k only reads entries k and N-k >= 1024) ----
    float2 pw[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int k = tid + 128 * q;
      if (k < kBins) {
        const float2 zk = zs[k], zn = zs[(kNfft - k) & (kNfft - 1)];
        const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
        const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
        pw[q] = make_float2(ar * ar + ai * ai, br * br + bi * bi);
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int k = tid + 128 * q;
      if (k < kBins) zs[k] = pw[q];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- sparse mel filterbank + dB: four lanes per filter (32 filters x 4 = one pass of the CTA), both frames at
    // once; fixed summation order, so the result does not depend on scheduling ----
    for (int m = tid >> 2; m < a.n_mels; m += kFwdThreads / 4) {
      const int start = __ldg(a.fb_start + m), len = __ldg(a.fb_len + m), off = __ldg(a.fb_off + m);
      float acc0 = 0.f, acc1 = 0.f;
      for (int i = tid & 3; i < len; i += 4) {
        const float wgt = fbw[off + i];
        const float2 p = zs[start + i];
        acc0 = fmaf(wgt, p.x, acc0);
        acc1 = fmaf(wgt, p.y, acc1);
      }
#pragma unroll
      for (int d = 2; d > 0; d >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, d);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, d);
      }
      if ((tid & 3) == 0) {
        float* o = a.out + (static_cast<size_t>(b) * a.n_mels + m) * a.n_frames + f0;
        o[0] = 10.f * log10f(fmaxf(acc0, 1e-10f));
        if (has2) o[1] = 10.f * log10f(fmaxf(acc1, 1e-10f));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Backward of the log-mel front-end (for gradient-based attacks through AcousticSystem, white_box_attack.py:
// 437-439): grad_x = d<grad_out, logmel(x)>/dx.  One CTA per clip; the clip's gradient is accumulated in shared
// memory (frames overlap 4x) and written once.  Per frame pair: recompute the forward FFT, then
//   dM = g * 10/(ln10 * M) (0 where M was clamped),  dP[k] = sum_m fb[k][m] dM[m],  Z[k] = 2 dP[k] X[k],
//   dy[n] = Re sum_{k=0}^{1024} Z[k] e^{+2 pi i k n / N}  -- evaluated as the inverse FFT of the Hermitian
//   extension of Z so that, again, two frames share one complex transform --  and grad_x += window * dy.
// ---------------------------------------------------------------------------------------------------
struct MelBwdArgs {
  const float* x;         // [B][L]
  const float* grad_out;  // [B][n_mels][n_frames]
  float* grad_x;          // [B][L]
  const float2* tw;
  const int* fb_start;
  const int* fb_len;
  const int* fb_off;
  const float* fb_w;
  int B, L, n_frames, n_mels;
};

__global__ void __launch_bounds__(kMelThreads) logmel_backward_kernel(const MelBwdArgs a) {
  extern __shared__ float gacc[];  // [L] gradient of this clip
  __shared__ float2 zs[kFftPad];
  __shared__ float2 tws[kNfft / 2];
  __shared__ float2 spec[2][kBins];  // X of the two frames, then Z = 2 dP X
  __shared__ float dmel[2][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const float* x = a.x + static_cast<size_t>(b) * a.L;
  for (int k = tid; k < kNfft / 2; k += kMelThreads) tws[k] = a.tw[k];
  for (int i = tid; i < a.L; i += kMelThreads) gacc[i] = 0.f;
  __syncthreads();

  for (int f0 = 0; f0 < a.n_frames; f0 += 2) {
    const bool has2 = (f0 + 1) < a.n_frames;
    // ---- forward: windowed frames -> X1, X2 (as in logmel_kernel) ----
    for (int n = tid; n < kNfft; n += kMelThreads) {
      const float c = (n < kNfft / 2) ? tws[n].x : -tws[n - kNfft / 2].x;
      const float w = 0.5f - 0.5f * c;
      const int i1 = f0 * kHop - kNfft / 2 + n, i2 = i1 + kHop;
      const float v1 = (i1 >= 0 && i1 < a.L) ? __ldg(x + i1) : 0.f;
      const float v2 = (has2 && i2 >= 0 && i2 < a.L) ? __ldg(x + i2) : 0.f;
      zs[fpad(__brev(static_cast<unsigned>(n)) >> 21)] = make_float2(v1 * w, v2 * w);
    }
    __syncthreads();
    fft2048_inplace(zs, tid, 1.f);
    for (int k = tid; k < kBins; k += kMelThreads) {
      const float2 zk = zs[fpad(k)], zn = zs[fpad((kNfft - k) & (kNfft - 1))];
      spec[0][k] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      spec[1][k] = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
    }
    __syncthreads();
    // ---- mel energies -> dL/dM ----
    for (int o = warp; o < 2 * a.n_mels; o += kMelThreads / 32) {
      const int fr = o / a.n_mels, m = o - fr * a.n_mels;
      const int start = a.fb_start[m], len = a.fb_len[m], off = a.fb_off[m];
      float acc = 0.f;
      for (int i = lane; i < len; i += 32) {
        const float2 v = spec[fr][start + i];
        acc = fmaf(a.fb_w[off + i], v.x * v.x + v.y * v.y, acc);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (lane == 0) {
        float g = 0.f;
        if ((fr == 0 || has2) && acc > 1e-10f)
          g = a.grad_out[(static_cast<size_t>(b) * a.n_mels + m) * a.n_frames + f0 + fr] * (4.342944819032518f / acc);
        dmel[fr][m] = g;  // 10 / ln(10) = 4.3429...
      }
    }
    __syncthreads();
    // ---- dL/dP[k] = sum_m fb[k][m] dM[m]; Z = 2 dP X (in place) ----
    for (int k = tid; k < kBins; k += kMelThreads) {
      float dp0 = 0.f, dp1 = 0.f;
      for (int m = 0; m < a.n_mels; ++m) {
        const int i = k - a.fb_start[m];
        if (i >= 0 && i < a.fb_len[m]) {
          const float wgt = a.fb_w[a.fb_off[m] + i];
          dp0 = fmaf(wgt, dmel[0][m], dp0);
          dp1 = fmaf(wgt, dmel[1][m], dp1);
        }
      }
      spec[0][k] = make_float2(2.f * dp0 * spec[0][k].x, 2.f * dp0 * spec[0][k].y);
      spec[1][k] = make_float2(2.f * dp1 * spec[1][k].x, 2.f * dp1 * spec[1][k].y);
    }
    __syncthreads();
    // ---- Hermitian extension H (H_0 = Re Z_0, H_1024 = Re Z_1024, H_k = Z_k/2, H_{N-k} = conj(Z_k)/2), packed
    //      as H1 + i H2, bit-reversed for the inverse transform ----
    for (int k = tid; k < kNfft; k += kMelThreads) {
      float2 h1, h2;
      if (k == 0 || k == kNfft / 2) {
        h1 = make_float2(spec[0][k].x, 0.f);
        h2 = make_float2(spec[1][k].x, 0.f);
      } else if (k < kNfft / 2) {
        h1 = make_float2(0.5f * spec[0][k].x, 0.5f * spec[0][k].y);
        h2 = make_float2(0.5f * spec[1][k].x, 0.5f * spec[1][k].y);
      } else {
        const int kk = kNfft - k;
        h1 = make_float2(0.5f * spec[0][kk].x, -0.5f * spec[0][kk].y);
        h2 = make_float2(0.5f * spec[1][kk].x, -0.5f * spec[1][kk].y);
      }
      zs[fpad(__brev(static_cast<unsigned>(k)) >> 21)] = make_float2(h1.x - h2.y, h1.y + h2.x);  // H1 + i*H2
    }
    __syncthreads();
    fft2048_inplace(zs, tid, -1.f);
    // ---- window and overlap-add (frames f0 and f0+1 touch disjoint phases of this loop: one sync between) ----
    for (int n = tid; n < kNfft; n += kMelThreads) {
      const float c = (n < kNfft / 2) ? tws[n].x : -tws[n - kNfft / 2].x;
      const float w = 0.5f - 0.5f * c;
      const int i1 = f0 * kHop - kNfft / 2 + n;
      if (i1 >= 0 && i1 < a.L) gacc[i1] += w * zs[fpad(n)].x;
    }
    __syncthreads();
    if (has2) {
      for (int n = tid; n < kNfft; n += kMelThreads) {
        const float c = (n < kNfft / 2) ? tws[n].x : -tws[n - kNfft / 2].x;
        const float w = 0.5f - 0.5f * c;
        const int i2 = (f0 + 1) * kHop - kNfft / 2 + n;
        if (i2 >= 0 && i2 < a.L) gacc[i2] += w * zs[fpad(n)].y;
      }
    }
    __syncthreads();
  }
  float* gx = a.grad_x + static_cast<size_t>(b) * a.L;
  for (int i = tid; i < a.L; i += kMelThreads) gx[i] = gacc[i];
}

}  // namespace ap
