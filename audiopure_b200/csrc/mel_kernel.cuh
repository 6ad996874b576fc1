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

struct MelArgs {
  const float* x;        // [B][L]
  float* out;            // [B][n_mels][n_frames]
  const float2* tw;      // [1024] exp(-2*pi*i*k/2048)
  const int* fb_start;   // [n_mels] first bin with non-zero weight
  const int* fb_len;     // [n_mels] number of consecutive non-zero bins
  const int* fb_off;     // [n_mels] offset of that run in fb_w
  const float* fb_w;     // concatenated non-zero weights
  int B, L, n_frames, n_mels;
};

__global__ void __launch_bounds__(kMelThreads) logmel_kernel(const MelArgs a) {
  __shared__ float2 zs[kNfft];
  __shared__ float2 tws[kNfft / 2];
  __shared__ float pw[2][kBins + 3];

  const int tid = threadIdx.x;
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
    zs[__brev(static_cast<unsigned>(n)) >> 21] = make_float2(v1 * w, v2 * w);
  }
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < 11; ++s) {
    const int half = 1 << s;
    for (int j = tid; j < kNfft / 2; j += kMelThreads) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos;
      const int i1 = i0 + half;
      const float2 t = tws[pos << (10 - s)];
      const float2 u = zs[i0], v = zs[i1];
      const float2 vt = make_float2(v.x * t.x - v.y * t.y, v.x * t.y + v.y * t.x);
      zs[i0] = make_float2(u.x + vt.x, u.y + vt.y);
      zs[i1] = make_float2(u.x - vt.x, u.y - vt.y);
    }
    __syncthreads();
  }

  // Z = F1 + i*F2  ->  F1[k] = (Z[k] + conj(Z[N-k]))/2,  F2[k] = (Z[k] - conj(Z[N-k]))/(2i)
  for (int k = tid; k < kBins; k += kMelThreads) {
    const float2 zk = zs[k], zn = zs[(kNfft - k) & (kNfft - 1)];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    pw[0][k] = ar * ar + ai * ai;
    pw[1][k] = br * br + bi * bi;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  for (int o = warp; o < 2 * a.n_mels; o += kMelThreads / 32) {
    const int fr = o / a.n_mels, m = o - fr * a.n_mels;
    const int start = a.fb_start[m], len = a.fb_len[m], off = a.fb_off[m];
    float acc = 0.f;
    for (int i = lane; i < len; i += 32) acc = fmaf(a.fb_w[off + i], pw[fr][start + i], acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0 && (fr == 0 || has2))
      a.out[(static_cast<size_t>(b) * a.n_mels + m) * a.n_frames + f0 + fr] = 10.f * log10f(fmaxf(acc, 1e-10f));
  }
}

}  // namespace ap
