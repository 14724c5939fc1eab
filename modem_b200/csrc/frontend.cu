// modem_b200/csrc/frontend.cu — sample ingest and Schmidl-Cox timing metric / detection.
//
// Replaces, for a batch of independent sample windows:
//   K0  next_sample(): ReadWAV int16 -> float, BlockDC, Hilbert<21>          (/root/reference/decode.cc:294-301,386)
//   K1a SchmidlCox::operator() metric part: P, R, timing = box161(|P|^2/R^2)  (decode.cc:86-91)
//   K1b Schmitt trigger + falling edge + running arg-max -> detection list    (decode.cc:93-115)
// The reference is a per-sample streaming state machine; here every sliding sum is a tile-parallel prefix-sum
// difference, the IIR DC blocker is a linear-recurrence scan, and the hysteresis/arg-max logic is a scan over
// {state -> state} maps, so a 95 200-sample window is processed by whole CTAs instead of one thread.
// HBM-bound: K1a reads each IQ sample once from DRAM (8 B/sample; the 1439-sample tile halo comes from L2).
#include "common.cuh"
#include "frontend.cuh"

namespace ofdmrx {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------ K0 mono
// y[t] = b (x[t] - x[t-1]) + a y[t-1]  (BlockDC, a = 2879/2880), then the 21-tap Hilbert FIR.
// One CTA per window, tiles of 2048 samples, 8 consecutive samples per thread; the recurrence is carried
// across threads by a scan over (alpha, beta) pairs of the affine map y_out = alpha * y_in + beta.
constexpr int kFeThreads = 256, kFePer = 8, kFeTile = kFeThreads * kFePer;

// ReadWAV scaling of a 16-bit sample (decode.cc:576: v / (2^15 - 1)); float input is what ReadWAV already delivered
__device__ __forceinline__ float fe_sample(int16_t v) { return (float)v / 32767.f; }
__device__ __forceinline__ float fe_sample(float v) { return v; }

template <int S, typename SampleT>
__global__ void __launch_bounds__(kFeThreads) k_frontend_mono(const SampleT *pcm, int64_t pcm_stride, const int32_t *n_samples,
	int n_default, cfx *iq, int64_t iq_stride, int iq_len, FrontendConsts fc)
{
	// Hilbert<T>: the output at step t is formed before x[t] is pushed: centre tap y[t-1-mid], odd offsets up to mid-1,
	// i.e. it uses y[t-2 mid .. t-2] (T = 21: y[t-20 .. t-2]); ybuf[j] = y[t0 - 2 mid + j]
	constexpr int T = Geo<S>::kFilterLen, kMid = (T - 1) / 2, kHist = 2 * kMid, kCo = (T - 1) / 4;
	__shared__ float ybuf[kHist + kFeTile];
	__shared__ float2 wsum[kFeThreads / 32];
	__shared__ float carry_y, carry_x;
	const float dc_a = fc.dc_a, dc_b = fc.dc_b;
	const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int n = n_samples ? n_samples[f] : n_default;
	const SampleT *src = pcm + (size_t)f * pcm_stride;
	cfx *dst = iq + (size_t)f * iq_stride;
	if (tid < kHist) ybuf[tid] = 0.f;
	if (tid == 0) { carry_y = 0.f; carry_x = 0.f; }
	__syncthreads();
	float a8 = dc_a;
	a8 = a8 * a8; a8 = a8 * a8; a8 = a8 * a8; // a^8
	for (int t0 = 0; t0 < iq_len; t0 += kFeTile) {
		float x[kFePer], u[kFePer];
		const int tb = t0 + tid * kFePer;
#pragma unroll
		for (int k = 0; k < kFePer; ++k) {
			const int t = tb + k;
			x[k] = t < n ? fe_sample(src[t]) : 0.f;
		}
		float xprev = tid == 0 ? carry_x : (tb - 1 < n ? fe_sample(src[tb - 1]) : 0.f);
		// local response with zero carry-in
		float y = 0.f;
#pragma unroll
		for (int k = 0; k < kFePer; ++k) {
			u[k] = dc_b * (x[k] - xprev);
			xprev = x[k];
			y = u[k] + dc_a * y;
		}
		// inclusive scan of (alpha, beta) over threads: compose earlier (A,B) then mine (a8,y) -> (A*a8, a8*B + y)
		float al = a8, be = y;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const float pa = __shfl_up_sync(FULL, al, d), pb = __shfl_up_sync(FULL, be, d);
			if (lane >= d) { be = al * pb + be; al = al * pa; }
		}
		if (lane == 31) wsum[wid] = make_float2(al, be);
		__syncthreads();
		// carry into this thread = state after all previous threads of the tile, starting from carry_y
		float cy = carry_y;
		for (int w = 0; w < wid; ++w) cy = wsum[w].x * cy + wsum[w].y;
		{
			const float pa = __shfl_up_sync(FULL, al, 1), pb = __shfl_up_sync(FULL, be, 1);
			if (lane > 0) cy = pa * cy + pb;
		}
		float yy = cy;
#pragma unroll
		for (int k = 0; k < kFePer; ++k) {
			yy = u[k] + dc_a * yy;
			ybuf[kHist + tid * kFePer + k] = yy;
		}
		__syncthreads();
		if (tid == kFeThreads - 1) { carry_y = yy; carry_x = x[kFePer - 1]; }
#pragma unroll
		for (int k = 0; k < kFePer; ++k) {
			const int j = k * kFeThreads + tid; // coalesced store order
			const int t = t0 + j;
			if (t < iq_len) {
				const float *c = &ybuf[j + kMid - 1]; // y[t-1-mid]
				const float re = fc.reco * c[0];
				float im = fc.imco[0] * (c[-1] - c[1]);
#pragma unroll
				for (int i = 1; i < kCo; ++i) im += fc.imco[i] * (c[-(2 * i + 1)] - c[2 * i + 1]);
				dst[t] = make_float2(re, im);
			}
		}
		__syncthreads();
		if (tid < kHist) ybuf[tid] = ybuf[kFeTile + tid];
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------ K0 IQ int16 / float2
__global__ void k_frontend_iq16(const int16_t *pcm, int64_t pcm_stride, const int32_t *n_samples, int n_default,
	cfx *iq, int64_t iq_stride, int iq_len)
{
	const int f = blockIdx.y;
	const int n = n_samples ? n_samples[f] : n_default;
	const short2 *src = reinterpret_cast<const short2 *>(pcm + (size_t)f * pcm_stride * 2);
	cfx *dst = iq + (size_t)f * iq_stride;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < iq_len; t += gridDim.x * blockDim.x) {
		cfx v = make_float2(0.f, 0.f);
		if (t < n) { const short2 s = src[t]; v = make_float2((float)s.x / 32767.f, (float)s.y / 32767.f); }
		dst[t] = v;
	}
}
__global__ void k_frontend_f32(const cfx *in, int64_t in_stride, const int32_t *n_samples, int n_default,
	cfx *iq, int64_t iq_stride, int iq_len)
{
	const int f = blockIdx.y;
	const int n = n_samples ? n_samples[f] : n_default;
	const cfx *src = in + (size_t)f * in_stride;
	cfx *dst = iq + (size_t)f * iq_stride;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < iq_len; t += gridDim.x * blockDim.x)
		dst[t] = t < n ? src[t] : make_float2(0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------ K1a timing metric
// For stream index t (the sample just pushed): c[t] = a[t-5119] conj(a[t-4479]), e[t] = |a[t-4479]|^2,
// P[t] = sum_{k<640} c[t-k], R[t] = max(0.5 sum_{k<1280} e[t-k], 0.064), m[t] = |P|^2/R^2,
// timing[t] = sum_{k<161} m[t-k]  (decode.cc:86-90 with search_pos = 2880, buffer_len = 8640).
// Tile of kMtTile outputs; extended index j = t - t0 + kMtHalo addresses a[t0 - 5918 + j].
#ifndef OFDMRX_MT_TILE
#define OFDMRX_MT_TILE 2048
#endif
constexpr int kMtTile = OFDMRX_MT_TILE; // outputs per CTA; the halo of 2 half-symbols + 161 samples is recomputed per tile
template <int S>
struct Mt {
	using G = Geo<S>;
	static constexpr int kThreads = G::kSymLen > 4096 ? 1024 : (kMtTile > 2048 ? 512 : 256);    // per-thread chunk stays ~11-20 samples
	static constexpr int kHalo = 2 * G::kHalf + G::kMatchLen - 2, kExt = kMtTile + kHalo;      // 1439, 3487 at 8 kHz
	static constexpr int kPer = (kExt + kThreads - 1) / kThreads;                               // 14
	static constexpr int kPad = kThreads * kPer;                                                // 3584
	// newest sample is buffer tap kBufferLen - 1: the correlator's taps search_pos + half and search_pos + symbol_len
	// (decode.cc:86) are the stream samples t - kOffOld and t - kOffCur
	// (the other tap, t - kOffOld with kOffOld = kBufferLen - 1 - (kSearchPos + kHalf) = 5119, is kOffCur + kHalf: the same stream, lagged)
	static constexpr int kOffCur = G::kBufferLen - 1 - (G::kSearchPos + G::kSymLen);            // 4479
};

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const T o = __shfl_up_sync(FULL, v, d);
		if (lane >= d) v += o;
	}
	return v;
}

// One bulk-asynchronous copy (cp.async.bulk, the 1-D form of TMA: SASS UBLKCP) stages the tile's span of the analytic stream
// in shared memory; it serves both correlator taps (the current stream and the one lagging by half a symbol are the same
// samples 640 apart).  Its base t0 - (kOffCur + kHalo + kLag) is an even sample index at all four rates, i.e. 16-byte aligned.
// OFDMRX_SYNC_TMA=0 keeps the round-1 staging (two coalesced register-load streams) for A/B runs.
#ifndef OFDMRX_SYNC_TMA
#define OFDMRX_SYNC_TMA 0 // measured per 10 000 windows at 8 kHz: 6.90 ms with the bulk copy, 6.27 ms without
#endif
#if OFDMRX_SYNC_TMA
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// Outputs: the Schmitt trigger's two comparisons as bit masks (hi: v > high, lo: v < low; one word per 32 stream steps — all
// k_sync_detect needs to list the edges) and the timing values themselves only for tiles that hold a value >= low (every
// sample between a rise and its fall does) or when the caller keeps the stage taps (write_all).
template <int S>
__global__ void __launch_bounds__(Mt<S>::kThreads) k_sync_metric(const cfx *iq, int64_t iq_stride, int iq_len, const int32_t *n_samples,
	int n_default, float *timing, int64_t timing_stride, uint32_t *masks, int mask_words, int write_all)
{
	constexpr int kMtHalo = Mt<S>::kHalo, kMtExt = Mt<S>::kExt, kMtPer = Mt<S>::kPer, kMtPad = Mt<S>::kPad, kMtThreads = Mt<S>::kThreads;
	constexpr int kLag = Geo<S>::kHalf, kLen2 = Geo<S>::kSymLen, kBox = Geo<S>::kMatchLen;
	extern __shared__ __align__(16) float sm[];
	float *sre = sm;                                    // c.re, then its prefix, later the prefix of m   [kMtPad]
	float *sim = sre + kMtPad;                          // c.im, then its prefix
	float *se = sim + kMtPad;                           // e, then its prefix
	__shared__ float wtot[3][Mt<S>::kThreads / 32];
#if OFDMRX_SYNC_TMA
	__shared__ __align__(8) uint64_t bar;
#endif
	const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int n = n_samples ? n_samples[f] : n_default;
	const int t0 = blockIdx.x * kMtTile;
	if (t0 > n) return; // stream has n+1 steps: t = 0..n
	const cfx *a = iq + (size_t)f * iq_stride;
	const int base = t0 - (Mt<S>::kOffCur + kMtHalo); // extended index j <-> a[t - kOffCur], t = t0 + j - kMtHalo
	{ // c[j] = a[j - lag] conj(a[j]) and e[j] = |a[j]|^2
		cfx cur[kMtPer], old[kMtPer];
#if OFDMRX_SYNC_TMA
		static_assert(3 * kMtPad * sizeof(float) >= (size_t)(kMtExt + kLag + 1) * sizeof(cfx), "the staged span fits the three arrays it is unpacked into");
		static_assert(((Mt<S>::kOffCur + kMtHalo + kLag) & 1) == 0 && (kMtTile & 1) == 0, "16-byte aligned span base");
		cfx *stage = reinterpret_cast<cfx *>(sm);       // stage[e] = a[lo + e], e in [0, kMtExt + kLag]
		const int lo = base - kLag;
		const int c0 = max(lo, 0), c1 = min((lo + kMtExt + kLag + 1) & ~1, iq_len); // copied samples [c0, c1): even bounds
		const int ncopy = max(c1 - c0, 0);
		if (tid == 0) mbar_init(&bar, 1);
		for (int e = tid; e < kMtExt + kLag + 1; e += kMtThreads) { // the parts of the span outside the stream read as zero
			const int idx = lo + e;
			if (idx < c0 || idx >= c1) stage[e] = make_float2(0.f, 0.f);
		}
		__syncthreads();
		if (tid == 0 && ncopy > 0) {
			mbar_expect_tx(&bar, (uint32_t)ncopy * (uint32_t)sizeof(cfx));
			bulk_g2s(stage + (c0 - lo), a + c0, (uint32_t)ncopy * (uint32_t)sizeof(cfx), &bar);
		}
		if (ncopy > 0) mbar_wait(&bar, 0);
#pragma unroll
		for (int k = 0; k < kMtPer; ++k) {
			const int j = tid + k * kMtThreads;
			cur[k] = j < kMtExt ? stage[j + kLag] : make_float2(0.f, 0.f);
			old[k] = j < kMtExt ? stage[j] : make_float2(0.f, 0.f);
		}
		__syncthreads(); // every thread holds its samples: the span is overwritten by the three arrays below
#else
		// straight from coalesced global loads (the lagged stream comes from L1/L2); all loads of a thread are issued before
		// the first shared-memory store
#pragma unroll
		for (int k = 0; k < kMtPer; ++k) {
			const int j = tid + k * kMtThreads, idx = base + j;
			cur[k] = (j < kMtExt && idx >= 0 && idx < iq_len) ? __ldg(&a[idx]) : make_float2(0.f, 0.f);
			const int io = idx - kLag;
			old[k] = (j >= kLag && j < kMtExt && io >= 0 && io < iq_len) ? __ldg(&a[io]) : make_float2(0.f, 0.f);
		}
#endif
#pragma unroll
		for (int k = 0; k < kMtPer; ++k) {
			const int j = tid + k * kMtThreads;
			const cfx c = j >= kLag ? cmulc(old[k], cur[k]) : make_float2(0.f, 0.f);
			sre[j] = c.x; sim[j] = c.y; se[j] = cnorm(cur[k]);
		}
	}
	__syncthreads();
	// per-thread chunk [j0, j0+kMtPer): local sums of c and e, then block scan
	const int j0 = tid * kMtPer;
	float cre[kMtPer], cim[kMtPer], ce[kMtPer];
	float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
	for (int k = 0; k < kMtPer; ++k) {
		const int j = j0 + k;
		s0 += sre[j]; s1 += sim[j]; s2 += se[j];
		cre[k] = s0; cim[k] = s1; ce[k] = s2;
	}
	float w0 = warp_incl_scan(s0, lane), w1 = warp_incl_scan(s1, lane), w2 = warp_incl_scan(s2, lane);
	if (lane == 31) { wtot[0][wid] = w0; wtot[1][wid] = w1; wtot[2][wid] = w2; }
	__syncthreads();
	float o0 = w0 - s0, o1 = w1 - s1, o2 = w2 - s2; // exclusive within warp
	for (int w = 0; w < wid; ++w) { o0 += wtot[0][w]; o1 += wtot[1][w]; o2 += wtot[2][w]; }
#pragma unroll
	for (int k = 0; k < kMtPer; ++k) {
		sre[j0 + k] = o0 + cre[k];
		sim[j0 + k] = o1 + cim[k];
		se[j0 + k] = o2 + ce[k];
	}
	__syncthreads();
	// m[j] for j in [1279, kMtExt): needs prefix[j] - prefix[j-640] (c) and prefix[j] - prefix[j-1280] (e)
	float mloc[kMtPer];
	float ms = 0.f;
#pragma unroll
	for (int k = 0; k < kMtPer; ++k) {
		const int j = j0 + k;
		float m = 0.f;
		if (j >= kLen2 - 1 && j < kMtExt) {
			const float pr = sre[j] - sre[j - kLag], pi = sim[j] - sim[j - kLag];
			float r = 0.5f * (se[j] - (j >= kLen2 ? se[j - kLen2] : 0.f));
			r = fmaxf(r, 0.0001f * (float)kLag); // decode.cc:87-89
			m = __fdiv_rn(pr * pr + pi * pi, r * r);
			// windows that start before the stream did (t < 0 contributions) are zero because a[<0] = 0
		}
		ms += m;
		mloc[k] = ms;
	}
	__syncthreads(); // everyone is done reading sre before it is overwritten with the prefix of m
	float wm = warp_incl_scan(ms, lane);
	if (lane == 31) wtot[0][wid] = wm;
	__syncthreads();
	float om = wm - ms;
	for (int w = 0; w < wid; ++w) om += wtot[0][w];
#pragma unroll
	for (int k = 0; k < kMtPer; ++k) sre[j0 + k] = om + mloc[k];
	__syncthreads();
	// trigger masks (decode.cc:76,93: low 0.17 x 161, high 0.19 x 161): one ballot pair per 32 stream steps
	const float low = (float)(0.17 * kBox), high = (float)(0.19 * kBox);
	uint32_t *mh = masks + (size_t)f * 2 * mask_words, *ml = mh + mask_words;
	bool keep = false;
	for (int i = tid; i < kMtTile; i += kMtThreads) {
		const int t = t0 + i, j = i + kMtHalo;
		const bool valid = t <= n;
		const float v = sre[j] - sre[j - kBox];
		const unsigned h = __ballot_sync(FULL, valid && v > high), l = __ballot_sync(FULL, valid && v < low);
		if (lane == 0) { mh[t >> 5] = h; ml[t >> 5] = l; }
		keep |= valid && v >= low;
	}
	if (!write_all && !__syncthreads_or(keep)) return; // nothing in this tile can lie between a rise and its fall
	float *out = timing + (size_t)f * timing_stride;
	for (int i = tid; i < kMtTile; i += kMtThreads) {
		const int t = t0 + i;
		if (t > n) break;
		const int j = i + kMtHalo;
		out[t] = sre[j] - sre[j - kBox];
	}
}

// ------------------------------------------------------------------------------------------------ K1b detection
// Schmitt trigger (low 0.17*161, high 0.19*161) + falling edge + first strict maximum inside each
// [rise, fall] segment (decode.cc:93-108).  One CTA per window.  The trigger only ever reacts to "v > high" while low
// and to "v < low" while high: k_sync_metric hands over those two comparisons as bit masks (one word per 32 samples);
// one warp walks the words (almost all of them are skipped with one test) and lists the edges in stream order; a warp
// per (rise, fall) segment finds the maximum among the timing values of the segment (the only ones ever read).
constexpr int kDtThreads = 256;
constexpr int kDtTile = 65536;                 // samples per pass (state and edge list carry over)
constexpr int kDtWords = kDtTile / 32;

template <int S>
__global__ void __launch_bounds__(kDtThreads) k_sync_detect(const float *timing, int64_t timing_stride, const uint32_t *masks, int mask_words,
	const int32_t *n_samples, int n_default, Detection *det, int32_t *det_count, int det_cap, int32_t *edges)
{
	constexpr int kMatchDel = Geo<S>::kMatchDel, kHalf = Geo<S>::kHalf, kGuardLen = Geo<S>::kGuardLen;
	__shared__ uint32_t hi_m[kDtWords], lo_m[kDtWords];
	__shared__ int n_ev_s, state_s;
	int32_t *ev_t = edges + (size_t)blockIdx.x * (2 * det_cap + 2); // rise / fall stream indices of this window, in order
	const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int n = (n_samples ? n_samples[f] : n_default) + 1; // steps t = 0..n_samples
	const float *tm = timing + (size_t)f * timing_stride;
	const uint32_t *gh = masks + (size_t)f * 2 * mask_words, *gl = gh + mask_words;
	if (tid == 0) { n_ev_s = 0; state_s = 0; } // the trigger starts low, so the first edge is always a rise
	for (int t0 = 0; t0 < n; t0 += kDtTile) {
		const int words = min(kDtWords, (n - t0 + 31) >> 5);
		for (int w = tid; w < words; w += kDtThreads) { hi_m[w] = gh[(t0 >> 5) + w]; lo_m[w] = gl[(t0 >> 5) + w]; }
		__syncthreads();
		if (wid == 0) { // all lanes run the same walk (uniform); 32 words are skipped per step while nothing can happen
			int s = state_s, ne = n_ev_s;
			for (int w0 = 0; w0 < words; w0 += 32) {
				const int wl = w0 + lane;
				const uint32_t mine = wl < words ? (s ? lo_m[wl] : hi_m[wl]) : 0u;
				if (__ballot_sync(FULL, mine != 0u) == 0u) continue;
				for (int w = w0; w < min(w0 + 32, words); ++w) {
					uint32_t m = s ? lo_m[w] : hi_m[w];
					while (m) {
						const int b = __ffs(m) - 1;
						if (ne < 2 * det_cap && lane == 0) ev_t[ne] = t0 + 32 * w + b;
						++ne;
						s ^= 1;
						m = b == 31 ? 0u : (s ? lo_m[w] : hi_m[w]) & (0xfffffffeu << b);
					}
				}
			}
			if (lane == 0) { state_s = s; n_ev_s = ne; }
		}
		__syncthreads();
	}
	__syncthreads(); // the edge list (global memory, written by warp 0) is read by every warp below
	const int n_ev = min(n_ev_s, 2 * det_cap);
	const int n_seg = n_ev / 2; // (rise, fall) pairs; an unfinished segment never fires
	// segments: edges alternate rise, fall, rise, ...  One warp per segment finds the first strict maximum.
	for (int sgi = tid >> 5; sgi < n_seg; sgi += kDtThreads / 32) {
		const int rise = ev_t[2 * sgi], fall = ev_t[2 * sgi + 1];
		float best = 0.f; // timing_max starts at 0 and only a strictly larger value replaces it
		int bi = -1;
		for (int t = rise + lane; t < fall; t += 32) { // (the fall step itself is below `low`: it never is the maximum)
			const float v = tm[t];
			if (v > best) { best = v; bi = t; }
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) {
			const float ob = __shfl_xor_sync(FULL, best, d);
			const int oi = __shfl_xor_sync(FULL, bi, d);
			if (oi >= 0 && (ob > best || (ob == best && (bi < 0 || oi < bi)))) { best = ob; bi = oi; }
		}
		if (lane == 0) {
			Detection d;
			d.t_fall = fall;
			d.t_max = bi;
			d.timing_max = best;
			// index_max: match_del at the maximum, +1 per later collect/process step, capped (decode.cc:99-105)
			d.index_max = bi < 0 ? 0 : min(kMatchDel + (fall - bi), kHalf + kGuardLen + kMatchDel);
			det[(size_t)f * det_cap + sgi] = d;
		}
	}
	if (tid == 0) det_count[f] = n_seg | (n_ev_s > 2 * det_cap ? kDetOverflowBit : 0);
}

} // namespace

cudaError_t launch_frontend(int rate, int format, const void *samples, int64_t stride, const int32_t *n_samples, int n_default, int n_frames,
	cfx *iq, int64_t iq_stride, int iq_len, const FrontendConsts &fc, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
	if (format == 0) {
#define OFDMRX_CALL(R) k_frontend_mono<R, int16_t><<<n_frames, kFeThreads, 0, s>>>((const int16_t *)samples, stride, n_samples, n_default, iq, iq_stride, iq_len, fc)
		OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	} else if (format == 3) {
#define OFDMRX_CALL(R) k_frontend_mono<R, float><<<n_frames, kFeThreads, 0, s>>>((const float *)samples, stride, n_samples, n_default, iq, iq_stride, iq_len, fc)
		OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	} else if (format == 1) {
		dim3 g((iq_len + 1023) / 1024, n_frames);
		k_frontend_iq16<<<g, 256, 0, s>>>((const int16_t *)samples, stride, n_samples, n_default, iq, iq_stride, iq_len);
	} else {
		dim3 g((iq_len + 1023) / 1024, n_frames);
		k_frontend_f32<<<g, 256, 0, s>>>((const cfx *)samples, stride, n_samples, n_default, iq, iq_stride, iq_len);
	}
	return cudaGetLastError();
}

template <int S>
static cudaError_t launch_sync_metric_t(const cfx *iq, int64_t iq_stride, int iq_len, const int32_t *n_samples, int n_default, int n_max, int n_frames,
	float *timing, int64_t timing_stride, uint32_t *masks, int mask_words, int write_all, cudaStream_t s)
{
	static DeviceOnce once;
	const size_t smem = (size_t)3 * Mt<S>::kPad * sizeof(float);
	if (cudaError_t e = set_dynamic_smem_once(once, k_sync_metric<S>, (int)smem)) return e;
	dim3 g((n_max + 1 + kMtTile - 1) / kMtTile, n_frames);
	k_sync_metric<S><<<g, Mt<S>::kThreads, smem, s>>>(iq, iq_stride, iq_len, n_samples, n_default, timing, timing_stride, masks, mask_words, write_all);
	return cudaGetLastError();
}

cudaError_t launch_sync_metric(int rate, const cfx *iq, int64_t iq_stride, int iq_len, const int32_t *n_samples, int n_default, int n_max, int n_frames,
	float *timing, int64_t timing_stride, uint32_t *masks, int mask_words, int write_all, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
	cudaError_t e = cudaSuccess;
#define OFDMRX_CALL(R) e = launch_sync_metric_t<R>(iq, iq_stride, iq_len, n_samples, n_default, n_max, n_frames, timing, timing_stride, masks, mask_words, write_all, s)
	OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	return e;
}

cudaError_t launch_sync_detect(int rate, const float *timing, int64_t timing_stride, const uint32_t *masks, int mask_words, const int32_t *n_samples,
	int n_default, int n_frames, Detection *det, int32_t *det_count, int det_cap, int32_t *edges, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
#define OFDMRX_CALL(R) k_sync_detect<R><<<n_frames, kDtThreads, 0, s>>>(timing, timing_stride, masks, mask_words, n_samples, n_default, det, det_count, det_cap, edges)
	OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	return cudaGetLastError();
}

} // namespace ofdmrx
