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

__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const MelArgs a) {
  __shared__ float2 zs[kFftPad];
  __shared__ float2 tws[kNfft / 2];
  __shared__ float pw[2][kBins + 3];
  __shared__ float fbw[kMaxFbNnz];  // a global read per filter tap made the mel reduction latency-bound (29 of 42 us)
  __shared__ int fbd[3][128];

  const int tid = threadIdx.x;
  for (int i = tid; i < a.fb_nnz; i += kMelThreads) fbw[i] = a.fb_w[i];
  for (int i = tid; i < a.n_mels; i += kMelThreads) {
    fbd[0][i] = a.fb_start[i];
    fbd[1][i] = a.fb_len[i];
    fbd[2][i] = a.fb_off[i];
  }
  const int pairs = (a.n_frames + 1) >> 1;
  const int b = blockIdx.x / pairs;
  const int f0 = (blockIdx.x - b * pairs) * 2;
  const bool has2 = (f0 + 1) < a.n_frames;
  const float* x = a.x + static_cast<size_t>(b) * a.L;

  for (int k = tid; k < kNfft / 2; k += kMelThreads) tws[k] = a.tw[k];
  __syncthreads();

  // frame f covers original samples [f*hop - n_fft/2, f*hop + n_fft/2); outside [0, L) is the zero padding
  for (int n = tid; n < kNfft; n += kMelThreads) {
    const float c = (n < kNfft / 2) ? tws[n].x : -tws[n - kNfft / 2].x;  // cos(2*pi*n/N)
    const float w = 0.5f - 0.5f * c;
    const int i1 = f0 * kHop - kNfft / 2 + n;
    const int i2 = i1 + kHop;
    const float v1 = (i1 >= 0 && i1 < a.L) ? __ldg(x + i1) : 0.f;
    const float v2 = (has2 && i2 >= 0 && i2 < a.L) ? __ldg(x + i2) : 0.f;
    zs[fpad(__brev(static_cast<unsigned>(n)) >> 21)] = make_float2(v1 * w, v2 * w);
  }
  __syncthreads();
  fft2048_inplace(zs, tid, 1.f);

  // Z = F1 + i*F2  ->  F1[k] = (Z[k] + conj(Z[N-k]))/2,  F2[k] = (Z[k] - conj(Z[N-k]))/(2i)
  for (int k = tid; k < kBins; k += kMelThreads) {
    const float2 zk = zs[fpad(k)], zn = zs[fpad((kNfft - k) & (kNfft - 1))];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    pw[0][k] = ar * ar + ai * ai;
    pw[1][k] = br * br + bi * bi;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  for (int o = warp; o < 2 * a.n_mels; o += kMelThreads / 32) {
    const int fr = o / a.n_mels, m = o - fr * a.n_mels;
    const int start = fbd[0][m], len = fbd[1][m], off = fbd[2][m];
    float acc = 0.f;
    for (int i = lane; i < len; i += 32) acc = fmaf(fbw[off + i], pw[fr][start + i], acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0 && (fr == 0 || has2))
      a.out[(static_cast<size_t>(b) * a.n_mels + m) * a.n_frames + f0 + fr] = 10.f * log10f(fmaxf(acc, 1e-10f));
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
