// modem_b200/csrc/fft.cuh — CTA-cooperative mixed-radix Stockham FFT in shared memory (radices 2, 3, 4, 5, 7:
// N = 640, 1280, 2560, 3528, 3840, 7056, 7680 — the symbol and half-symbol lengths of the four sample rates).
// Unnormalised, forward sign (exp(-2 pi j n k / N)) like DSP::FastFourierTransform<N,cmplx,-1> at
// /root/reference/decode.cc:43,191; the backward transform (decode.cc:44) is conj(fwd(conj(x))).
// A radix-5 pass is unavoidable for these lengths; twiddles come from a W_N^k table (global/L1).
#pragma once
#include "common.cuh"

namespace ofdmrx {

#ifdef __CUDACC__
#define OFDMRX_HD __host__ __device__ __forceinline__
#else
#define OFDMRX_HD inline
#endif
// CTA-wide barrier of the cooperative routines: __syncthreads() on the device; nothing in the one-"thread" host builds of the
// CPU tests; a real barrier over host threads when a test harness defines OFDMRX_HOST_CTA (tests/stimulus_cta_tsan.cu runs the
// routines on several host threads under ThreadSanitizer to prove that no barrier is missing).
#if defined(__CUDA_ARCH__)
#define OFDMRX_CTA_SYNC() __syncthreads()
#elif defined(OFDMRX_HOST_CTA)
void ofdmrx_host_cta_sync();
#define OFDMRX_CTA_SYNC() ofdmrx_host_cta_sync()
#else
#define OFDMRX_CTA_SYNC() ((void)0)
#endif

template <int R> OFDMRX_HD void bfly(cfx *v);
template <> OFDMRX_HD void bfly<2>(cfx *v)
{
	const cfx a = v[0], b = v[1];
	v[0] = cadd(a, b);
	v[1] = csub(a, b);
}
template <> OFDMRX_HD void bfly<4>(cfx *v)
{
	const cfx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
	const cfx t3 = make_float2(d.y, -d.x); // d * (-j)
	v[0] = cadd(t0, t2);
	v[1] = cadd(t1, t3);
	v[2] = csub(t0, t2);
	v[3] = csub(t1, t3);
}
template <> OFDMRX_HD void bfly<5>(cfx *v)
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

template <> OFDMRX_HD void bfly<3>(cfx *v)
{
	const float h = 0.86602540378443864676f; // sin(2 pi / 3)
	const cfx t1 = cadd(v[1], v[2]), d = csub(v[1], v[2]);
	const cfx t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
	const cfx sd = make_float2(h * d.y, -h * d.x); // -j h d
	v[0] = cadd(v[0], t1);
	v[1] = cadd(t2, sd);
	v[2] = csub(t2, sd);
}
template <> OFDMRX_HD void bfly<7>(cfx *v)
{
	// cos / sin of 2 pi k / 7, k = 1, 2, 3
	const float c1 = 0.62348980185873353053f, c2 = -0.22252093395631440429f, c3 = -0.90096886790241912624f;
	const float s1 = 0.78183148246802980871f, s2 = 0.97492791218182360702f, s3 = 0.43388373911755812048f;
	const cfx a1 = cadd(v[1], v[6]), a2 = cadd(v[2], v[5]), a3 = cadd(v[3], v[4]);
	const cfx b1 = csub(v[1], v[6]), b2 = csub(v[2], v[5]), b3 = csub(v[3], v[4]);
	// X_p = R_p - j I_p, X_{7-p} = R_p + j I_p with R_p = v0 + sum_k cos(p k t) a_k, I_p = sum_k sin(p k t) b_k
	const cfx r1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x + c3 * a3.x, v[0].y + c1 * a1.y + c2 * a2.y + c3 * a3.y);
	const cfx r2 = make_float2(v[0].x + c2 * a1.x + c3 * a2.x + c1 * a3.x, v[0].y + c2 * a1.y + c3 * a2.y + c1 * a3.y);
	const cfx r3 = make_float2(v[0].x + c3 * a1.x + c1 * a2.x + c2 * a3.x, v[0].y + c3 * a1.y + c1 * a2.y + c2 * a3.y);
	const cfx i1 = make_float2(s1 * b1.x + s2 * b2.x + s3 * b3.x, s1 * b1.y + s2 * b2.y + s3 * b3.y);
	const cfx i2 = make_float2(s2 * b1.x - s3 * b2.x - s1 * b3.x, s2 * b1.y - s3 * b2.y - s1 * b3.y);
	const cfx i3 = make_float2(s3 * b1.x - s1 * b2.x + s2 * b3.x, s3 * b1.y - s1 * b2.y + s2 * b3.y);
	v[0] = cadd(v[0], cadd(a1, cadd(a2, a3)));
	v[1] = make_float2(r1.x + i1.y, r1.y - i1.x); v[6] = make_float2(r1.x - i1.y, r1.y + i1.x);
	v[2] = make_float2(r2.x + i2.y, r2.y - i2.x); v[5] = make_float2(r2.x - i2.y, r2.y + i2.x);
	v[3] = make_float2(r3.x + i3.y, r3.y - i3.x); v[4] = make_float2(r3.x - i3.y, r3.y + i3.x);
}

// one decimation-in-time Stockham pass: combines R sub-transforms of length m into length m*R
template <int N, int R>
OFDMRX_HD void fft_pass(const cfx *src, cfx *dst, int m, const cfx *tw, int tid, int nthr)
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

// radices of the passes, in order (symbol lengths and half lengths of the four sample rates)
template <int N> struct FftPlan;
template <> struct FftPlan<640>  { static constexpr int n = 5; static constexpr int r[7] = {4, 4, 4, 2, 5, 1, 1}; };
template <> struct FftPlan<1280> { static constexpr int n = 5; static constexpr int r[7] = {4, 4, 4, 4, 5, 1, 1}; };
template <> struct FftPlan<2560> { static constexpr int n = 6; static constexpr int r[7] = {4, 4, 4, 4, 2, 5, 1}; };
template <> struct FftPlan<3528> { static constexpr int n = 6; static constexpr int r[7] = {4, 2, 3, 3, 7, 7, 1}; };
template <> struct FftPlan<7056> { static constexpr int n = 6; static constexpr int r[7] = {4, 4, 3, 3, 7, 7, 1}; };
template <> struct FftPlan<3840> { static constexpr int n = 6; static constexpr int r[7] = {4, 4, 4, 4, 3, 5, 1}; };
template <> struct FftPlan<7680> { static constexpr int n = 7; static constexpr int r[7] = {4, 4, 4, 4, 2, 3, 5}; };

template <int N, int P, int M>
OFDMRX_HD cfx *fft_run(cfx *src, cfx *dst, const cfx *tw, int tid, int nthr)
{
	if constexpr (P == FftPlan<N>::n) {
		return src;
	} else {
		constexpr int R = FftPlan<N>::r[P];
		fft_pass<N, R>(src, dst, M, tw, tid, nthr);
		OFDMRX_CTA_SYNC();
		return fft_run<N, P + 1, M * R>(dst, src, tw, tid, nthr);
	}
}

// in: buf0 (N values), scratch buf1; returns the buffer holding the result (buf1 after an odd number of passes, else buf0).
// All threads of the CTA must call.
template <int N>
OFDMRX_HD cfx *fft_fwd(cfx *buf0, cfx *buf1, const cfx *tw, int tid, int nthr)
{
	return fft_run<N, 0, 1>(buf0, buf1, tw, tid, nthr);
}

} // namespace ofdmrx
