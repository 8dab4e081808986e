// Waveform front-end of ReasoningCodec_film's tokenize direction ("move the CPU log-mel on-device"):
//   tools/tokenizer/ReasoningCodec_film/reason_tokenizer.py
//     :37      transfer16k = torchaudio.transforms.Resample(24000, 16000)         polyphase windowed-sinc FIR, conv1d stride = orig
//     :67-72   get_whisper_features: transfer16k -> .cpu().numpy() -> WhisperFeatureExtractor (torch.stft n_fft 400 hop 160, hann,
//              reflect-centred; |.|^2; Slaney mel 201 x 80; log10(clamp 1e-10); max(., clip max - 8); (. + 4) / 4) -> back to the device
//   tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
//     :363-365 wavlm_transfer(wav_24k) + 160 zero samples (the WavLM encoder's input)
// The reference leaves the device for this step (one D2H of 6 x 30 s of audio, numpy / torch-CPU STFT, one H2D of the features); here
// both stay on the device: one launch for the resampler, two for the features.
//
//   fe_resample_kernel     y[b, new * m + p] = sum_k kern[p][k] * x[b, orig * m + k - width]   (zero outside), fp32 FMA; positions
//                          past the valid output length are written as zeros (the 30 s padding / the 160 appended zeros)
//   fe_logmel_kernel       CTA = 8 frames of one clip: windowed frames in shared memory as doubles (reflect padding resolved on the fly), a
//                          direct 400-point DFT with thread = frequency bin: the twiddle advances by a complex rotation in double
//                          registers and all 8 frames share it; double accumulation (the whole 6 x 30 s batch is 2.9 G FMA - noise
//                          next to the encoder's 7 TFLOP - and the result carries no DFT rounding at all: what differs from torch.stft
//                          is torch's own fp32 FFT error), power -> mel projection (double accumulation) -> log10 -> out (B, n_mels, n_frames)
//   fe_logmel_norm_kernel  one CTA per clip: max over the clip, then (max(v, mx - 8) + 4) / 4 in the reference's fp32 operation order
#include <algorithm>
#include <string>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

constexpr int FE_FRAMES = 8;      // frames per CTA of fe_logmel_kernel
constexpr int FE_MAX_FFT = 512;   // n_fft bound (shared-memory tables)
constexpr int FE_MAX_TAPS = 2048; // new * (2 * width + orig) bound of the resampler's filter table

__global__ void __launch_bounds__(256) fe_resample_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ kern,
                                                          float* __restrict__ y, long long ldy, int L, int n_valid, int n_store, int orig,
                                                          int newf, int width, int K) {
  __shared__ float ks[FE_MAX_TAPS];
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < newf * K; i += blockDim.x) ks[i] = kern[i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (j >= n_store) return;
  float acc = 0.f;
  if (j < n_valid) {
    const int m = j / newf, p = j - m * newf;
    const long long base = (long long)orig * m - width;
    const float* xb = x + (size_t)b * ldx;
    const float* kp = ks + p * K;
    for (int k = 0; k < K; ++k) {
      const long long i = base + k;
      const float v = (i >= 0 && i < L) ? xb[i] : 0.f;
      acc = fmaf(kp[k], v, acc);
    }
  }
  y[(size_t)b * ldy + j] = acc;
}

// thread = one frequency bin for all FE_FRAMES frames of the CTA.  The twiddle e^{-2 pi i k n / n_fft} advances by a complex rotation in
// double registers (error ~ n_fft * 2^-53), the windowed samples sit in shared memory as doubles and every load is a broadcast: per
// (frame, bin, sample) 2 DFMA + 1/2 rotation instruction and no table gather.  (First version: cos / sin tables in shared memory
// gathered at index k n mod n_fft - ncu: 330 M shared-memory bank conflicts, L1 pipe 93 % busy, fp64 pipe 9 %, 2.07 ms for 6 x 30 s.)
__global__ void __launch_bounds__(256) fe_logmel_kernel(const float* __restrict__ wav, long long ld, const float* __restrict__ window,
                                                        const float* __restrict__ filt, float* __restrict__ out, int L, int n_fft, int hop,
                                                        int n_mels, int n_frames) {
  __shared__ double xw[FE_FRAMES][FE_MAX_FFT];
  __shared__ float pw[FE_FRAMES][FE_MAX_FFT / 2 + 1];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y, f0 = blockIdx.x * FE_FRAMES;
  const int n_bins = n_fft / 2 + 1, half = n_fft / 2;
  const float* xb = wav + (size_t)b * ld;
  for (int i = tid; i < FE_FRAMES * n_fft; i += blockDim.x) {
    const int f = i / n_fft, n = i - f * n_fft;
    float v = 0.f;
    if (f0 + f < n_frames) {
      long long s = (long long)(f0 + f) * hop - half + n;  // torch.stft(center=True, pad_mode='reflect')
      if (s < 0) s = -s;
      if (s >= L) s = 2LL * (L - 1) - s;
      v = __fmul_rn(xb[s], window[n]);
    }
    xw[f][n] = (double)v;
  }
  __syncthreads();
  for (int k = tid; k < n_bins; k += blockDim.x) {
    const double a = 6.283185307179586476925286766559 * (double)k / (double)n_fft;
    const double ca = cos(a), sa = sin(a);
    double c = 1.0, s = 0.0;  // cos / sin(2 pi k n / n_fft)
    double re[FE_FRAMES], im[FE_FRAMES];
#pragma unroll
    for (int f = 0; f < FE_FRAMES; ++f) re[f] = im[f] = 0.0;
    for (int n = 0; n < n_fft; ++n) {
#pragma unroll
      for (int f = 0; f < FE_FRAMES; ++f) {
        const double v = xw[f][n];
        re[f] = fma(v, c, re[f]);
        im[f] = fma(v, s, im[f]);
      }
      const double cn = fma(c, ca, -(s * sa));
      s = fma(s, ca, c * sa);
      c = cn;
    }
#pragma unroll
    for (int f = 0; f < FE_FRAMES; ++f) pw[f][k] = (float)(re[f] * re[f] + im[f] * im[f]);
  }
  __syncthreads();
  for (int o = tid; o < FE_FRAMES * n_mels; o += blockDim.x) {
    const int f = o / n_mels, m = o - f * n_mels;
    if (f0 + f >= n_frames) continue;
    double acc = 0.0;
    for (int k = 0; k < n_bins; ++k) acc = fma((double)filt[(size_t)k * n_mels + m], (double)pw[f][k], acc);
    const float mel = fmaxf((float)acc, 1e-10f);
    out[((size_t)b * n_mels + m) * n_frames + f0 + f] = (float)log10((double)mel);
  }
}

__global__ void __launch_bounds__(1024) fe_logmel_norm_kernel(float* __restrict__ x, long long n) {
  __shared__ float red[32];
  __shared__ float mx_s;
  pdl_launch_dependents();
  pdl_wait();
  float* xb = x + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (long long i = tid; i < n; i += blockDim.x) mx = fmaxf(mx, xb[i]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_max(lane < nw ? red[lane] : -INFINITY);
    if (lane == 0) mx_s = t;
  }
  __syncthreads();
  const float floor_v = __fsub_rn(mx_s, 8.0f);
  for (long long i = tid; i < n; i += blockDim.x) xb[i] = __fdiv_rn(__fadd_rn(fmaxf(xb[i], floor_v), 4.0f), 4.0f);
}

cudaError_t launch_fe_resample(const LaunchCtx& lc, const float* x, long long ldx, const float* kern, float* y, long long ldy, int B, int L,
                               int n_valid, int n_store, int orig, int newf, int width) {
  const int K = 2 * width + orig;
  return launch(lc, fe_resample_kernel, dim3((unsigned)((n_store + 255) / 256), (unsigned)B), dim3(256), 0, x, ldx, kern, y, ldy, L, n_valid,
                n_store, orig, newf, width, K);
}

cudaError_t launch_fe_logmel(const LaunchCtx& lc, const float* wav, long long ld, const float* window, const float* filt, float* out, int B, int L,
                             int n_fft, int hop, int n_mels, int n_frames) {
  cudaError_t e = launch(lc, fe_logmel_kernel, dim3((unsigned)((n_frames + FE_FRAMES - 1) / FE_FRAMES), (unsigned)B), dim3(256), 0, wav, ld,
                         window, filt, out, L, n_fft, hop, n_mels, n_frames);
  if (e != cudaSuccess) return e;
  return launch(lc, fe_logmel_norm_kernel, dim3((unsigned)B), dim3(1024), 0, out, (long long)n_mels * n_frames);
}

}  // namespace
}  // namespace ua2

using namespace ua2;

extern "C" {

int ua2_resample_f32(const float* x, long long ldx, const float* kernel, float* y, long long ldy, int B, int L, int n_valid, int n_store, int orig,
                     int newf, int width, void* stream) {
  UA2_REQUIRE(x && kernel && y, "null argument");
  UA2_REQUIRE(B >= 1 && B <= 65535 && L >= 1 && n_store >= 1 && n_valid >= 0 && n_valid <= n_store, "bad sizes");
  UA2_REQUIRE(orig >= 1 && newf >= 1 && width >= 0 && (long long)newf * (2 * width + orig) <= FE_MAX_TAPS, "filter table too large");
  UA2_REQUIRE(ldx >= L && ldy >= n_store, "row strides shorter than the rows");
  UA2_REQUIRE((long long)(n_valid + newf - 1) / newf * orig <= (long long)L + orig, "n_valid exceeds ceil(new * L / orig)");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_fe_resample(lc, x, ldx, kernel, y, ldy, B, L, n_valid, n_store, orig, newf, width));
  return UA2_OK;
}

int ua2_whisper_logmel_f32(const float* wav16, long long ld, const float* window, const float* mel_filters, float* out, int B, int L, int n_fft,
                           int hop, int n_mels, int n_frames, void* stream) {
  UA2_REQUIRE(wav16 && window && mel_filters && out, "null argument");
  UA2_REQUIRE(B >= 1 && B <= 65535, "batch out of range");
  UA2_REQUIRE(n_fft >= 2 && n_fft <= FE_MAX_FFT && (n_fft % 2) == 0 && hop >= 1 && n_mels >= 1 && n_mels <= 1024, "bad transform sizes");
  UA2_REQUIRE(L > n_fft / 2 && ld >= L, "clip shorter than half a window (reflect padding), or row stride shorter than the row");
  UA2_REQUIRE(n_frames >= 1 && (long long)(n_frames - 1) * hop <= L, "n_frames exceeds 1 + L / hop");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_fe_logmel(lc, wav16, ld, window, mel_filters, out, B, L, n_fft, hop, n_mels, n_frames));
  return UA2_OK;
}

}  // extern "C"
