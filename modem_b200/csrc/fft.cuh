// modem_b200/csrc/fft.cuh — CTA-cooperative mixed-radix Stockham FFT in shared memory (N = 1280 = 4^4*5, 640 = 4^3*2*5).
// Unnormalised, forward sign (exp(-2 pi j n k / N)) like DSP::FastFourierTransform<N,cmplx,-1> at
// /root/reference/decode.cc:43,191; the backward transform (decode.cc:44) is conj(fwd(conj(x))).
// A radix-5 pass is unavoidable for these lengths; twiddles come from a W_N^k table (global/L1).
#pragma once
#include "common.cuh"

namespace ofdmrx {

template <int R> __device__ __forceinline__ void bfly(cfx *v);
template <> __device__ __forceinline__ void bfly<2>(cfx *v)
{
	const cfx a = v[0], b = v[1];
	v[0] = cadd(a, b);
	v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void bfly<4>(cfx *v)
{
	const cfx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
	const cfx t3 = make_float2(d.y, -d.x); // d * (-j)
	v[0] = cadd(t0, t2);
	v[1] = cadd(t1, t3);
	v[2] = csub(t0, t2);
	v[3] = csub(t1, t3);
}
template <> __device__ __forceinline__ void bfly<5>(cfx *v)
{
	const float c1 = 0.30901699437494742f, s1 = 0.95105651629515357f, c2 = -0.80901699437494742f, s2 = 0.58778525229247313f;
	const cfx a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
	const cfx r1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
	const cfx r2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
	const cfx i1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
	const cfx i2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
	v[0] = cadd(v[0], cadd(a1, a2));
	// X1 = r1 - j i1, X4 = r1 + j i1, X2 = r2 - j i2, X3 = r2 + j i2   (-j (x+jy) = y - jx)
	v[1] = make_float2(r1.x + i1.y, r1.y - i1.x);
	v[4] = make_float2(r1.x - i1.y, r1.y + i1.x);
	v[2] = make_float2(r2.x + i2.y, r2.y - i2.x);
	v[3] = make_float2(r2.x - i2.y, r2.y + i2.x);
}

// one decimation-in-time Stockham pass: combines R sub-transforms of length m into length m*R
template <int N, int R>
__device__ __forceinline__ void fft_pass(const cfx *src, cfx *dst, int m, const cfx *tw, int tid, int nthr)
{
	const int l = N / (m * R);
	for (int b = tid; b < N / R; b += nthr) {
		const int j = b / m, k = b - j * m;
		cfx v[R];
#pragma unroll
		for (int q = 0; q < R; ++q) v[q] = src[k + m * (j + l * q)];
#pragma unroll
		for (int q = 1; q < R; ++q) v[q] = cmul(v[q], tw[q * k * l]);
		bfly<R>(v);
#pragma unroll
		for (int p = 0; p < R; ++p) dst[k + m * (p + R * j)] = v[p];
	}
}

// in: buf0 (N values), scratch buf1; returns the buffer holding the result (buf1 for 640 / 1280, buf0 for 2560, which takes
// an even number of passes).  All threads of the CTA must call.
template <int N>
__device__ __forceinline__ cfx *fft_fwd(cfx *buf0, cfx *buf1, const cfx *tw, int tid, int nthr)
{
	static_assert(N == 2560 || N == 1280 || N == 640, "lengths of the 8 / 16 kHz receive paths");
	fft_pass<N, 4>(buf0, buf1, 1, tw, tid, nthr); __syncthreads();
	fft_pass<N, 4>(buf1, buf0, 4, tw, tid, nthr); __syncthreads();
	fft_pass<N, 4>(buf0, buf1, 16, tw, tid, nthr); __syncthreads();
	if constexpr (N == 2560) {
		fft_pass<N, 4>(buf1, buf0, 64, tw, tid, nthr); __syncthreads();
		fft_pass<N, 2>(buf0, buf1, 256, tw, tid, nthr); __syncthreads();
		fft_pass<N, 5>(buf1, buf0, 512, tw, tid, nthr); __syncthreads();
		return buf0;
	} else if constexpr (N == 1280) {
		fft_pass<N, 4>(buf1, buf0, 64, tw, tid, nthr); __syncthreads();
		fft_pass<N, 5>(buf0, buf1, 256, tw, tid, nthr); __syncthreads();
		return buf1;
	} else {
		fft_pass<N, 2>(buf1, buf0, 64, tw, tid, nthr); __syncthreads();
		fft_pass<N, 5>(buf0, buf1, 128, tw, tid, nthr); __syncthreads();
		return buf1;
	}
}

} // namespace ofdmrx
