// modem_b200/csrc/demod.cu — payload symbols: FFT, differential demodulation, Theil–Sen phase line, soft demapping.
//
// Replaces (mode 6: 432 carriers x 50 rows, 8PSK), as three kernels over a batch of windows:
//   k_demod_fft   one CTA per window: 1 pilot + 50 x (mix by the frame phasor, FFT-1280, cons = X_j / X_{j-1} with
//                 erasure), 8PSK hard/map and the decision-directed phase error of every carrier
//                                                         (/root/reference/decode.cc:456-477,483-486, psk.hh:118-139)
//   k_theil_sen   one WARP per row (50 x windows independent rows): DSP::TheilSenEstimator — exact upper median of
//                 all 93 096 pairwise slopes, then of the 432 intercepts                          (decode.cc:488-492)
//   k_soft_demap  one CTA per window: derotation, cumulative Es/N0 -> precision, PhaseShiftKeying<8>::soft ->
//                 code[3*(432 j + i) + b], lengthen()                 (decode.cc:493-495,505-529, psk.hh:125-130)
//
// Theil–Sen without sorting 93 096 quotients: a robust pilot line (Huber M-estimate) and its residual scale give a bracket
// [blo, bhi) ~2e-4 sigma wide that almost always holds rank 46 548.  With u_k = y_k - blo x_k and w = bhi - blo a pair
// (i < j) lies below the bracket iff u_j < u_i and inside it iff additionally u_j - u_i < w (x_j - x_i), so ONE sweep over
// all pairs costs compares and no division; only pairs inside the bracket (or within a rounding margin of its edges) are
// evaluated as IEEE quotients, exactly as the reference forms them, and the answer is selected among those — the result
// is the exact order statistic std::nth_element returns.  If the rank falls outside, the counts are exact with respect
// to the bracket edges, so the enclosure tightens and the next bracket is extrapolated; rows that defeat the bracket search
// altogether (long runs of tied quotients) take a bit-wise binary search with exact counting inside that enclosure.
#include "common.cuh"
#include "frontend.cuh"
#include "fft.cuh"
#include <algorithm>
#include <cstdlib>

namespace ofdmrx {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kDmThreads = 320; // one radix-4 butterfly of the FFT-1280 per thread

// per-row geometry of the Theil-Sen estimator: n carriers at x = i - n/2 (decode.cc:452,484), ranks as std::nth_element
// is asked for them (element count/2 of the n(n-1)/2 slopes and of the n intercepts; mode 6: 432 -> 93 096 / 46 548 / 216)
struct TsDims {
	int n, half, nblk, pairs, rank_slope, rank_yint;
	__device__ explicit TsDims(int cols) : n(cols), half(cols / 2), nblk((cols + 31) >> 5), pairs(cols * (cols - 1) / 2),
		rank_slope(cols * (cols - 1) / 4), rank_yint(cols / 2) {}
};

__device__ __forceinline__ int f2ord(float v) { const int i = __float_as_int(v); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// PhaseShiftKeying<4>::hard + map (psk.hh:72-87)
__device__ __forceinline__ void psk4_hard_map(cfx c, cfx &m)
{
	const float r = 0.70710678118654752440f;
	m = make_float2(c.x < 0.f ? -r : r, c.y < 0.f ? -r : r);
}
// PhaseShiftKeying<8>::hard + map (psk.hh:118-139)
__device__ __forceinline__ void psk8_hard_map(cfx c, cfx &m)
{
	const float cos_pi_8 = 0.92387953251128675613f, sin_pi_8 = 0.38268343236508977173f;
	const bool swap = fabsf(c.x) < fabsf(c.y);
	const float re = swap ? sin_pi_8 : cos_pi_8, im = swap ? cos_pi_8 : sin_pi_8;
	m = make_float2(c.x < 0.f ? -re : re, c.y < 0.f ? -im : im);
}
__device__ __forceinline__ void psk_hard_map(cfx c, cfx &m, bool qpsk)
{
	if (qpsk) psk4_hard_map(c, m);
	else psk8_hard_map(c, m);
}

// ================================================================================================ k_demod_fft
template <int S>
struct FftShared {
	cfx buf0[Geo<S>::kSymLen];
	cfx buf1[Geo<S>::kSymLen];
	cfx rot[Geo<S>::kSymLen];   // exp(-j cfo i), i < symbol_len: the frame phasor inside one symbol
	cfx base[kMaxRows + 2];     // exp(-j cfo n0(sym)): the frame phasor at the start of every symbol
	cfx prev[kMaxCols];
};

template <int S>
__global__ void __launch_bounds__(kDmThreads) k_demod_fft(const cfx *iq, int64_t iq_stride, int iq_len, const FrameState *stv,
	const cfx *tw, cfx *cons_raw, float *yph)
{
	constexpr int kSymLen = Geo<S>::kSymLen, kPitch = Geo<S>::kPitch; // shadow the 8 kHz constants
	extern __shared__ __align__(16) unsigned char smraw_fft[];
	FftShared<S> &s = *reinterpret_cast<FftShared<S> *>(smraw_fft);
	const int f = blockIdx.x, tid = threadIdx.x;
	const FrameState &st = stv[f];
	if (st.status != ST_OK) return;
	const cfx *a = iq + (size_t)f * iq_stride;
	const ModeInfo mi = mode_info(st.mode);
	const int cols = mi.cols;
	const bool qpsk = mi.mod_bits == 2;
	const int p0 = st.sc_pos + 2 * kPitch; // pilot body (decode.cc:456-459)
	const double turns = -(double)st.cfo_rad / 6.283185307179586476925286766559;
	// the frame phasor osc (decode.cc:403,459-461,468-470) at sample i of symbol sym is exp(-j cfo (n0 + i)): the factor
	// of i comes from a per-window table, the factor of n0 is one evaluation per symbol (both from a double phase)
	for (int i = tid; i < kSymLen; i += kDmThreads) s.rot[i] = phasor_turns(turns * (double)i);
	for (int sym = tid; sym <= mi.rows; sym += kDmThreads) s.base[sym] = phasor_turns(turns * (double)(kSymLen + kPitch * sym)); // steps since the header symbol
	__syncthreads();
	// the samples of symbol sym + 1 are loaded into registers while symbol sym goes through its FFT passes
	constexpr int kPerThread = (kSymLen + kDmThreads - 1) / kDmThreads; // 4 (8 / 23 / 24 at 16 / 44.1 / 48 kHz)
	cfx nxt[kPerThread];
#pragma unroll
	for (int k = 0; k < kPerThread; ++k) {
		const int i = tid + k * kDmThreads, idx = p0 + i;
		nxt[k] = (i < kSymLen && idx >= 0 && idx < iq_len) ? a[idx] : make_float2(0.f, 0.f);
	}
	for (int sym = 0; sym <= mi.rows; ++sym) {
		const cfx base = s.base[sym];
#pragma unroll
		for (int k = 0; k < kPerThread; ++k) {
			const int i = tid + k * kDmThreads;
			if (i < kSymLen) s.buf0[i] = cmul(nxt[k], cmul(base, s.rot[i]));
		}
		if (sym < mi.rows) {
			const int w1 = p0 + kPitch * (sym + 1);
#pragma unroll
			for (int k = 0; k < kPerThread; ++k) {
				const int i = tid + k * kDmThreads, idx = w1 + i;
				nxt[k] = (i < kSymLen && idx >= 0 && idx < iq_len) ? a[idx] : make_float2(0.f, 0.f);
			}
		}
		__syncthreads();
		const cfx *const X = fft_fwd<kSymLen>(s.buf0, s.buf1, tw, tid, kDmThreads);
		for (int k = tid; k < cols; k += kDmThreads) {
			const cfx cur = X[(k - cols / 2 + kSymLen) % kSymLen];
			if (sym > 0) {
				const int row = sym - 1;
				const cfx c = demod_or_erase(cur, s.prev[k]);
				const size_t o = (size_t)f * kMaxCons + row * cols + k;
				cons_raw[o] = c;
				cfx m;
				psk_hard_map(c, m, qpsk);
				const cfx e = cmulc(c, m);
				yph[o] = atan2f(e.y, e.x); // decode.cc:483-486
			}
			s.prev[k] = cur;
		}
		__syncthreads();
	}
}

// ================================================================================================ k_theil_sen
constexpr int kTsWarps = 4;            // rows in flight per CTA (one per warp)
constexpr int kTsCap = 2048;           // in-bracket pairs a warp can queue ...
constexpr int kTsLaneCap = kTsCap / 32; // ... as 32 private sub-queues
constexpr int kTsPad = kMaxCols;       // 16 x 32 columns, the tail beyond the mode's carrier count holds +inf
[[maybe_unused]] constexpr int kTsRows = kTsPad / 32;   // up to 16 row blocks: lane L owns carriers i = 32 R + L

constexpr int kTsCandCap = 1280;       // in-bracket quotients kept for the final select
// The row's phase values: a copy in shared memory (1) or read where they lie through the read-only path (0).  Without the
// copy a row needs 9 KB instead of 11 KB and six 4-warp CTAs fit an SM instead of five (A/B switch).
#ifndef OFDMRX_TS_Y_SMEM
#define OFDMRX_TS_Y_SMEM 0
#endif
constexpr int kTsCtasPerSm = OFDMRX_TS_Y_SMEM ? 5 : 6;
#ifndef OFDMRX_TS_HUBER_ITS
#define OFDMRX_TS_HUBER_ITS 4 // re-weighted least-squares steps of the pilot at most (A/B switch)
#endif
#ifndef OFDMRX_TS_MIN_ITS
#define OFDMRX_TS_MIN_ITS 1     // ... but at least this many
#endif
#ifndef OFDMRX_TS_CONV
#define OFDMRX_TS_CONV 0.0f  // ... fewer when a step moved the slope by less than this fraction of the bracket's half-width (0: never).
// Measured with 0.4: 0.8 ms faster per 10 000 clean windows, but one README-chain row in 500 000 then went through the exact
// bisection (60 rows' worth of work at the tail of the launch, +4 ms): one lane's sub-queue ran over while the bracket was far
// from full, and the overflow zoom of the time did not zoom (ts_slope below).  That is fixed; the early
// stop stays off until the fixed build has been timed.
#endif

constexpr int kTsTaskCap = 16;         // scan continuations a lane can park per sweep (beyond that they run on the spot)
struct TsSweep {                       // per sweep: every 32-column chunk sorted by u
	float su[kTsPad];                  // u of the chunk's columns in ascending order
	uint16_t sj[kTsPad];               // original column of each sorted entry
	uint16_t task[32 * kTsTaskCap];    // parked scans (row block << 9 | chunk << 5 | sorted position), one private list per lane
};
struct TsShared {
#if OFDMRX_TS_Y_SMEM
	float y[kTsPad];
#endif
	union {
		TsSweep sw;
		int cand[kTsCandCap];          // after the sweep: ordered-int images of the exact in-bracket quotients
	};
	union {
		uint16_t q[kTsCap];            // queued pairs (row block << 9 | column), one private sub-queue per lane
		int hist[256];                 // after their evaluation: scratch of the radix selects
	};
};
static_assert(sizeof(TsSweep) <= kTsCandCap * sizeof(int) && kTsCandCap >= kTsPad, "the sweep scratch sits under the candidate / intercept scratch");
static_assert(sizeof(TsShared) == (OFDMRX_TS_Y_SMEM ? 11264 : 9216), "11 KB per row: five 4-warp CTAs (20 rows in flight) per SM; 9 KB: six");
static_assert(kTsCtasPerSm * (kTsWarps * sizeof(TsShared) + 1024) <= 228 * 1024, "shared memory of the resident CTAs");

// k-th smallest (0-based) of the n values v[] (shared memory of this warp; order-preserving integer images of
// floats, see f2ord), exact: radix select, 8 bits per level inside the [min,max] range; hist = 256 ints of warp scratch.
__device__ __noinline__ int warp_select_radix(const int *v, int n, int k, int *hist, int lane)
{
	int mn = 0x7fffffff, mx = (int)0x80000000;
#pragma unroll 1
	for (int i = lane; i < n; i += 32) { const int o = v[i]; mn = min(mn, o); mx = max(mx, o); }
	mn = __reduce_min_sync(FULL, mn);
	mx = __reduce_max_sync(FULL, mx);
	int lo = mn;
	int sh = max(0, 32 - __clz(((unsigned)mx - (unsigned)mn) | 1u) - 8);
	for (int level = 0; level < 5; ++level) {
#pragma unroll
		for (int b = 0; b < 8; ++b) hist[lane + 32 * b] = 0;
		__syncwarp();
#pragma unroll 1
		for (int i = lane; i < n; i += 32) {
			const unsigned b = ((unsigned)v[i] - (unsigned)lo) >> sh; // values below lo wrap to huge bins
			if (b < 256u) atomicAdd(&hist[b], 1);
		}
		__syncwarp();
		int h[8], sum = 0;
#pragma unroll
		for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; sum += h[b]; }
		int incl = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
		const int excl = incl - sum;
		const bool mine = k >= excl && k < incl;
		int bin = 0, kk = 0;
		if (mine) {
			kk = k - excl;
#pragma unroll
			for (int b = 0; b < 8; ++b) { if (bin == b && kk >= h[b]) { kk -= h[b]; bin = b + 1; } }
			bin += lane * 8;
		}
		const unsigned bal = __ballot_sync(FULL, mine);
		if (bal == 0u) return mx; // k out of range (callers never ask for it): keep the function total
		const int src = __ffs(bal) - 1;
		bin = __shfl_sync(FULL, bin, src);
		k = __shfl_sync(FULL, kk, src);
		lo += (int)((unsigned)bin << sh);
		__syncwarp();
		if (sh == 0) break;
		sh = max(0, sh - 8);
	}
	return lo;
}

// The same order statistic in two passes for values that spread like measurements do: 256 buckets linear in the float
// value between min and max (a monotone map, so the k-th smallest lies in the bucket where the running count passes k),
// then the few members of that bucket are ranked against each other.  Crowded buckets (ties, outliers that stretch the
// range) go to the radix select, on the bucket's members only.  v[] is overwritten from its start with those members.
// (lo <= every value < hi when the caller knows such bounds — the slope candidates lie inside their bracket —, else lo > hi)
__device__ __noinline__ int warp_select_kth(int *v, int n, int k, int *hist, int lane, float lo, float hi)
{
	float mn = lo, mx = hi;
	if (!(lo < hi)) {
		mn = __int_as_float(0x7f800000); mx = -mn;
#pragma unroll 2
		for (int i = lane; i < n; i += 32) { const float x = ord2f(v[i]); mn = fminf(mn, x); mx = fmaxf(mx, x); }
#pragma unroll
		for (int d = 16; d; d >>= 1) { mn = fminf(mn, __shfl_xor_sync(FULL, mn, d)); mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, d)); }
	}
	if (!(mx - mn < 3.0e38f) || !(mx > mn)) return warp_select_radix(v, n, k, hist, lane); // all equal, or not finite
	const float sc = 256.f / (mx - mn);
#pragma unroll
	for (int b = 0; b < 8; ++b) hist[lane + 32 * b] = 0;
	__syncwarp();
#pragma unroll 2
	for (int i = lane; i < n; i += 32) atomicAdd(&hist[min(255, __float2int_rz((ord2f(v[i]) - mn) * sc))], 1);
	__syncwarp();
	int h[8], sum = 0;
#pragma unroll
	for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; sum += h[b]; }
	int incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
	const int excl = incl - sum;
	const bool mine = k >= excl && k < incl;
	int bin = 0, kk = 0;
	if (mine) {
		kk = k - excl;
#pragma unroll
		for (int b = 0; b < 8; ++b) { if (bin == b && kk >= h[b]) { kk -= h[b]; bin = b + 1; } }
		bin += lane * 8;
	}
	const unsigned bal = __ballot_sync(FULL, mine);
	if (bal == 0u) return f2ord(mx); // k out of range (callers never ask for it): keep the function total
	const int src = __ffs(bal) - 1;
	bin = __shfl_sync(FULL, bin, src);
	kk = __shfl_sync(FULL, kk, src);
	// members of that bucket, compacted to the front of v[] (reads run ahead of the writes: slot <= i)
	int m = 0;
#pragma unroll 1
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + lane;
		const int o = i < n ? v[i] : 0;
		const bool in = i < n && min(255, __float2int_rz((ord2f(o) - mn) * sc)) == bin;
		const unsigned bl = __ballot_sync(FULL, in);
		__syncwarp();
		if (in) v[m + __popc(bl & ((1u << lane) - 1u))] = o;
		m += __popc(bl);
	}
	__syncwarp();
	if (m > 32) return warp_select_radix(v, m, kk, hist, lane);
	const int mine_v = lane < m ? v[lane] : 0x7fffffff;
	int rank = 0;
#pragma unroll 2
	for (int e = 0; e < m; ++e) { const int o = v[e]; rank += (o < mine_v) || (o == mine_v && e < lane); }
	const unsigned hit = __ballot_sync(FULL, lane < m && rank == kk);
	return __shfl_sync(FULL, mine_v, __ffs(hit) - 1);
}

// the row's phase values (valid indices only: 0 <= i < n)
#if OFDMRX_TS_Y_SMEM
#define TS_Y(i) s.y[i]
#else
#define TS_Y(i) __ldg(&yrow[i])
#endif

// ---- the pair sweep ---------------------------------------------------------------------------------------------
// With u_k = y_k - blo x_k and w = bhi - blo >= 0 a pair (i < j) lies below the bracket iff u_j < a_i = u_i - eps and inside
// it (or within the rounding margin eps of an edge) iff additionally u_j - u_i < w (x_j - x_i) + eps — compares only, no
// division; eps covers the roundings of u and of the reference's own fl(fl(y_j - y_i) / d) with a factor > 2 to spare.
// The columns are sorted by u inside every chunk of 32 (one warp bitonic sort per chunk), so for a row i and a chunk the
// count below is a 6-step binary search instead of 32 compares, and the in-bracket columns are among the few sorted entries
// that follow: u_j < t = u_i + w (x_max(chunk) - x_i) + 2 eps bounds the scan.  The chunk holding i itself only counts
// columns j > i: a prefix bit-set over the sorted order (one warp scan per row block, kept in registers) turns that into a
// shuffle and a popcount.
// Lane L owns the rows i = 32 R + L; in-bracket pairs go to the lane's private sub-queue s.q[32 n + L].
// A third of the (row, chunk) visits find an entry inside the scan bound, one in fourteen a second one — but a loop over
// "my entries" runs as long as the slowest lane's for all 32 (3.7 trips on average: it was half of this function).  So a
// visit examines the first entry in line, and a second entry inside the bound parks the rest of the scan in a per-lane task
// list; the parked scans run afterwards in one flat loop where every lane advances its own list at its own pace.
__device__ __forceinline__ void ts_sort_chunks(TsShared &s, const float *yrow, const TsDims &d, int lane, float blo)
{
	const float inf = __int_as_float(0x7f800000);
#pragma unroll 1
	for (int K = 0; K < d.nblk; ++K) {
		const int j = 32 * K + lane;
		float key = j < d.n ? fmaf(-blo, (float)(j - d.half), TS_Y(j)) : inf;
		int idx = lane;
		// bitonic sort of (key, idx) across the warp, ascending
#pragma unroll
		for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
			for (int dd = k >> 1; dd > 0; dd >>= 1) {
				const float ok = __shfl_xor_sync(FULL, key, dd);
				const int oi = __shfl_xor_sync(FULL, idx, dd);
				const bool keep_min = ((lane & dd) == 0) == ((lane & k) == 0);
				const bool other_less = ok < key || (ok == key && oi < idx);
				if (keep_min == other_less) { key = ok; idx = oi; }
			}
		}
		s.sw.su[j] = key;
		s.sw.sj[j] = (uint16_t)(32 * K + idx);
	}
	__syncwarp();
}

// number of entries of the sorted chunk with u < a (0..32), for two chunks at once (two independent chains of six dependent
// shared-memory loads); one word per entry: 32 lanes hit 32 banks or the same word
__device__ __forceinline__ void ts_lower_bound2(const float *c0, const float *c1, float a, int &p0, int &p1)
{
	p0 = 0; p1 = 0;
#pragma unroll
	for (int st = 16; st > 0; st >>= 1) {
		const float e0 = c0[p0 + st - 1], e1 = c1[p1 + st - 1];
		p0 += e0 < a ? st : 0;
		p1 += e1 < a ? st : 0;
	}
	const float e0 = c0[p0], e1 = c1[p1]; // p <= 31 here
	p0 += e0 < a ? 1 : 0;
	p1 += e1 < a ? 1 : 0;
}

struct TsLane { int cb, nq, ntask; };

// row i = 32 R + lane against chunk K, p = entries of the chunk below a: count, examine the first entry inside the scan
// bound, park the scan if a second one follows
__device__ __forceinline__ void ts_visit(TsShared &s, TsLane &ln, int lane, int R, int K, int p, int i, float ui, float w, float eps, uint32_t first_p_set)
{
	const float *chunk = s.sw.su + 32 * K;
	ln.cb += K == R ? __popc(first_p_set) : p;
	// in-bracket candidates: sorted entries from p on while u < t (rows beyond the carriers: t = -inf)
	const float t = ui + fmaf(w, (float)(32 * K + 31 - i), eps + eps);
	const float u0 = chunk[min(p, 31)];
	if (p < 32 && u0 < t) {
		const int j = s.sw.sj[32 * K + p];
		if (u0 - ui < fmaf(w, (float)(j - i), eps) && (K != R || j > i)) {
			if (ln.nq < kTsLaneCap) s.q[32 * ln.nq + lane] = (uint16_t)((R << 9) | j);
			++ln.nq;
		}
		++p;
		if (p < 32 && chunk[p] < t) {
			if (ln.ntask < kTsTaskCap) {
				s.sw.task[32 * ln.ntask + lane] = (uint16_t)((R << 9) | (K << 5) | p);
				++ln.ntask;
			} else { // (no room to park it: finish the scan here)
				do {
					const int jj = s.sw.sj[32 * K + p];
					if (chunk[p] - ui < fmaf(w, (float)(jj - i), eps) && (K != R || jj > i)) {
						if (ln.nq < kTsLaneCap) s.q[32 * ln.nq + lane] = (uint16_t)((R << 9) | jj);
						++ln.nq;
					}
					++p;
				} while (p < 32 && chunk[p] < t);
			}
		}
	}
}

// all pairs: returns this lane's count of pairs definitely below the bracket; nq = pairs this lane queued
__device__ __forceinline__ int sweep_pairs(TsShared &s, const float *yrow, const TsDims &d, int lane, float blo, float bhi, float eps, int &nq)
{
	const float ninf = __int_as_float(0xff800000);
	const float w = bhi - blo, eps2 = eps + eps;
	const uint32_t above = lane == 31 ? 0u : 0xfffffffeu << lane; // column offsets beyond the lane's own
	TsLane ln;
	ln.cb = 0; ln.nq = nq; ln.ntask = 0;
#pragma unroll 1
	for (int R = 0; R < d.nblk; ++R) {
		const int i = 32 * R + lane;
		const float ui = i < d.n ? fmaf(-blo, (float)(i - d.half), TS_Y(i)) : ninf; // (the value the sort stored for column i)
		const float a = ui - eps;
		// own chunk: set of in-chunk column offsets among the first p sorted entries, p = lane + 1 (inclusive scan)
		uint32_t pm = 1u << (s.sw.sj[32 * R + lane] & 31);
#pragma unroll
		for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(FULL, pm, dd); if (lane >= dd) pm |= o; }
#pragma unroll 1
		for (int K = R; K < d.nblk; K += 2) {
			const bool two = K + 1 < d.nblk; // (warp-uniform)
			int p0, p1;
			ts_lower_bound2(s.sw.su + 32 * K, s.sw.su + 32 * (two ? K + 1 : K), a, p0, p1);
			uint32_t own = 0u;
			if (K == R) {
				const uint32_t first_p = __shfl_sync(FULL, pm, max(p0 - 1, 0)); // every lane takes part; p = 0 -> empty set
				own = (p0 > 0 ? first_p : 0u) & above;
			}
			ts_visit(s, ln, lane, R, K, p0, i, ui, w, eps, own);
			if (two) ts_visit(s, ln, lane, R, K + 1, p1, i, ui, w, eps, 0u);
		}
	}
	int ntask = ln.ntask;
	nq = ln.nq;
	int cb = ln.cb;
	// the parked scans: entry p of chunk K is inside row i's bound
	{
		int tn = 0, R = 0, K = 0, p = 0, i = 0;
		float ui = 0.f, t = 0.f;
		bool have = false;
		for (;;) {
			if (!have && tn < ntask) {
				const uint32_t code = s.sw.task[32 * tn + lane];
				++tn;
				R = (int)(code >> 9); K = (int)((code >> 5) & 15u); p = (int)(code & 31u);
				i = 32 * R + lane;
				ui = fmaf(-blo, (float)(i - d.half), TS_Y(i));
				t = ui + fmaf(w, (float)(32 * K + 31 - i), eps2);
				have = true;
			}
			if (!__any_sync(FULL, have)) break;
			if (have) {
				const float uj = s.sw.su[32 * K + p];
				const int j = s.sw.sj[32 * K + p];
				if (uj - ui < fmaf(w, (float)(j - i), eps) && (K != R || j > i)) {
					if (nq < kTsLaneCap) s.q[32 * nq + lane] = (uint16_t)((R << 9) | j);
					++nq;
				}
				++p;
				have = p < 32 && s.sw.su[32 * K + p] < t;
			}
		}
	}
	return cb;
}

// fallback for rows that defeat the bracket search: smallest value T with #(quotient <= T) > rank, by a binary
// search over the ordered-int image of the floats with an exact count (one IEEE division per pair) per step.
// L <= answer < U when the caller's bracket search has established such bounds (exact counts on both sides), else -inf / +inf:
// the search then starts inside them (two ordered-int steps of slack for the -0 / +0 pair).
__device__ __noinline__ float ts_slope_bisect(const float *y, int n, int rank, int lane, float L, float U)
{
	int lo = (int)0x80000000, hi = 0x7fffffff; // answer in [lo, hi]
	if (L > __int_as_float(0xff800000)) lo = f2ord(L) - 2;
	if (U < __int_as_float(0x7f800000)) hi = f2ord(U) + 1;
	while (lo < hi) {
		const int mid = (int)(((long long)lo + (long long)hi) >> 1);
		const float t = ord2f(mid);
		int cnt = 0;
		for (int i = 0; i < n - 1; ++i) {
			const float yi = y[i];
			for (int j = i + 1 + lane; j < n; j += 32) cnt += __fdiv_rn(y[j] - yi, (float)(j - i)) <= t;
		}
		cnt = __reduce_add_sync(FULL, cnt);
		if (cnt > rank) hi = mid; else lo = mid + 1;
	}
	return ord2f(lo);
}

// exact upper median of the pairwise slopes of (x = i - 216, y[i]); one warp, yrow = the row's n values
// hint: expected (Theil-Sen - OLS) of this row, in slope units (0 = none); pilot_out / half_out: the OLS slope and the
// bracket half-width used, for the caller's running estimate of that gap
__device__ float ts_slope(TsShared &s, const float *yrow, const TsDims &d, int lane, int &sweeps, float hint, float half_k, float &pilot_out, float &half_out)
{
	pilot_out = 0.f; half_out = 0.f;
	// ---- pilot: least-squares line and residual spread (only steers the bracket, so fp32 sums are good enough).
	// The passes over the row are rolled loops that re-read the values (L1 / shared memory) instead of holding sixteen of them
	// in registers under sixteen-fold unrolled code: this kernel runs out of instruction cache before anything else
	// (7 200 instructions, 24 warps at different places of them).
	// (the pilot's passes read a copy of the row in the candidate scratch, which is free until the first sweep sorts into it;
	// every lane reads back only the columns it stored.  Re-reading the row from L1 / L2 up to ten times made the kernel's time
	// depend on what the L2 happened to hold: 16.4 ms or 20 - 24 ms for the same README-chain batch.)
	float *ys = reinterpret_cast<float *>(s.cand);
	float a0 = 0.f, a1 = 0.f;
	float ymin = __int_as_float(0x7f800000), ymax = -ymin;
	const float x0 = (float)(lane - d.half) + 0.5f; // centred abscissa of column `lane`; column 32 r + lane: x0 + 32 r
#pragma unroll 2
	for (int r = 0; r < d.nblk; ++r) {
		const int i = 32 * r + lane;
		if (i < d.n) {
			const float yv = TS_Y(i);
			ys[i] = yv;
			a0 += yv;
			a1 = fmaf(x0 + (float)(32 * r), yv, a1);
			ymin = fminf(ymin, yv);
			ymax = fmaxf(ymax, yv);
		}
	}
#pragma unroll
	for (int dd = 16; dd; dd >>= 1) {
		a0 += __shfl_xor_sync(FULL, a0, dd);
		a1 += __shfl_xor_sync(FULL, a1, dd);
		ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, dd));
		ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, dd));
	}
	if (ymin == ymax) return 0.f; // erased row: every difference is 0, every quotient +0
	if (!(ymax - ymin < 3.0e38f)) return ts_slope_bisect(yrow, d.n, d.rank_slope, lane, __int_as_float(0xff800000), __int_as_float(0x7f800000)); // non-finite input: stay total
	const float sxx = (float)d.n * ((float)d.n * (float)d.n - 1.f) / 12.f;
	float c0 = a1 / sxx, icpt = a0 / (float)d.n; // least-squares line through the centred abscissa x + 0.5
	float r2 = 0.f;
#pragma unroll 2
	for (int r = 0; r < d.nblk; ++r) {
		const int i = 32 * r + lane;
		if (i < d.n) {
			const float e = ys[i] - icpt - c0 * (x0 + (float)(32 * r));
			r2 = fmaf(e, e, r2);
		}
	}
#pragma unroll
	for (int dd = 16; dd; dd >>= 1) r2 += __shfl_xor_sync(FULL, r2, dd);
	float srob = sqrtf(r2 / (float)(d.n - 2));
	// Huber M-estimate of the line (k = 1.345, scale by Huber's proposal 2, four re-weighted least-squares steps).  Least
	// squares is a poor stand-in for Theil-Sen once a few carriers misbehave (multipath notch at a band edge, wrapped phase
	// decisions): on the README impairment chain the two differ by ~5 bracket widths (3 sweeps per row); the Huber line
	// stays within ~0.4 bracket widths of the Theil-Sen slope on clean, AWGN and multipath rows alike.
	{
		const float kh = 1.345f, beta = 0.71016f; // beta = E[min(z^2, k^2)], z ~ N(0,1)
		const float conv_scale = (432.f / (float)d.n) * sqrtf(432.f / (float)d.n) * (d.n > 432 ? (432.f / (float)d.n) * sqrtf(432.f / (float)d.n) : 1.f); // sigma and shrink factors of the bracket below
#pragma unroll 1
		for (int it = 0; it < OFDMRX_TS_HUBER_ITS; ++it) {
			const float cap = kh * srob;
			float acc = 0.f;
#pragma unroll 2
			for (int r = 0; r < d.nblk; ++r) {
				const int i = 32 * r + lane;
				if (i < d.n) {
					const float e = ys[i] - icpt - c0 * (x0 + (float)(32 * r));
					acc += fminf(e * e, cap * cap);
				}
			}
#pragma unroll
			for (int dd = 16; dd; dd >>= 1) acc += __shfl_xor_sync(FULL, acc, dd);
			srob = sqrtf(acc / ((float)d.n * beta));
			const float lim = kh * srob;
			float sw = 0.f, swx = 0.f, swy = 0.f, swxx = 0.f, swxy = 0.f;
#pragma unroll 2
			for (int r = 0; r < d.nblk; ++r) {
				const int i = 32 * r + lane;
				if (i < d.n) {
					const float x = x0 + (float)(32 * r), yv = ys[i];
					const float e = fabsf(yv - icpt - c0 * x);
					const float w = e <= lim ? 1.f : __fdividef(lim, e);
					sw += w; swx = fmaf(w, x, swx); swy = fmaf(w, yv, swy);
					swxx = fmaf(w * x, x, swxx); swxy = fmaf(w * x, yv, swxy);
				}
			}
#pragma unroll
			for (int dd = 16; dd; dd >>= 1) {
				sw += __shfl_xor_sync(FULL, sw, dd); swx += __shfl_xor_sync(FULL, swx, dd); swy += __shfl_xor_sync(FULL, swy, dd);
				swxx += __shfl_xor_sync(FULL, swxx, dd); swxy += __shfl_xor_sync(FULL, swxy, dd);
			}
			const float den = sw * swxx - swx * swx;
			if (!(den > 0.f) || !(srob > 0.f)) break; // degenerate weights: keep the previous line
			const float c1 = (sw * swxy - swx * swy) / den;
			const float moved = fabsf(c1 - c0);
			c0 = c1;
			icpt = (swy - c0 * swx) / sw;
			// (optional early stop: the step was under OFDMRX_TS_CONV of the bracket's half-width)
			if (it + 1 >= OFDMRX_TS_MIN_ITS && moved < OFDMRX_TS_CONV * half_k * srob * conv_scale) break;
		}
	}
	// robust residual scale, rescaled so that the bracket below (sized for 432 carriers) keeps its width in units of the
	// standard deviation of (Theil-Sen - pilot), which goes like sigma / n^1.5
	const float scale = 432.f / (float)d.n;
	const float sigma = srob * scale * sqrtf(scale);
	const float yabs = fmaxf(fabsf(ymin), fabsf(ymax));
	// +-1.35e-4 sigma around the pilot holds rank 46 548 in ~99 % of the rows and ~1100 of the 93 096 quotients
	// (~35 per lane's sub-queue).
	// (more carriers than 432: the quotient count inside a bracket of given relative width grows like n^1.5; shrink it to
	// keep ~1100 candidates, which is what the per-row scratch is sized for)
	const float shrink = d.n > 432 ? scale * sqrtf(scale) : 1.f;
	float half = fmaxf(half_k * sigma * shrink, fmaxf(fabsf(c0) * 4e-6f, 1e-12f));
	pilot_out = c0; half_out = half;
	float blo = c0 + hint - half, bhi = c0 + hint + half;
	// enclosure of the answer established so far: #(q < L) = cL <= rank < cU = #(q < U)  (exact counts; +-inf = unknown)
	const float inf = __int_as_float(0x7f800000);
	float L = -inf, U = inf;
	int cL = 0, cU = d.pairs;
	for (int attempt = 0; attempt < 16; ++attempt) {
		++sweeps;
		const float bmax = fmaxf(fabsf(blo), fabsf(bhi));
		// |u_k| rounding (fma, <= 2^-24 |u|) on both sides plus the two roundings of the reference's quotient
		const float eps = 6e-7f * (yabs + 432.f * bmax) + 1e-30f;
		__syncwarp();
		ts_sort_chunks(s, yrow, d, lane, blo);
		int nql = 0; // pairs this lane queued
		int cb = sweep_pairs(s, yrow, d, lane, blo, bhi, eps, nql);
		__syncwarp();
		const int nq = __reduce_add_sync(FULL, nql), nqmax = __reduce_max_sync(FULL, nql);
		const float width = bhi - blo;
		if (nqmax > kTsLaneCap) {
			// more in-bracket pairs than a sub-queue holds.  The definite counts are still complete: cd pairs lie below blo
			// for certain and at most cd + nq lie below bhi, which tells where inside the bracket the rank sits.
			const int cd = __reduce_add_sync(FULL, cb);
			const int kd = d.rank_slope - cd;
			if (kd < 0) { U = blo; cU = cd; }
			else if (kd >= nq) { L = bhi; cL = cd + nq; }
			else {
				// zoom in around where the rank should sit: to ~cap/3 queued pairs in all, and — when it is ONE lane's sub-queue
				// that ran over while the bracket as a whole is far from full (a carrier lying on the median line pairs up with
				// many others) — far enough for that lane to fit.  (Without the second bound such a bracket was "zoomed" to more
				// than its own width, i.e. not at all, sixteen times over, and the row went through the bisection below.)
				const float centre = blo + width * (((float)kd + 0.5f) / (float)nq);
				const float hw = 0.5f * width * fminf((float)kTsCap / (3.f * (float)nq), (0.75f * (float)kTsLaneCap) / (float)nqmax);
				blo = fmaxf(centre - hw, blo);
				bhi = fminf(centre + hw, bhi);
				if (!(blo < bhi)) break;
				continue;
			}
		} else {
			// exact quotients of the queued pairs, as the reference forms them; the in-bracket ones are compacted into s.cand
			int nin = 0;
			for (int n = 0; n < nqmax; ++n) {
				bool in = false;
				float q = 0.f;
				if (n < nql) {
					const uint32_t code = s.q[32 * n + lane];
					const int i = 32 * (int)(code >> 9) + lane, j = (int)(code & 511u);
					q = __fdiv_rn(TS_Y(j) - TS_Y(i), (float)(j - i));
					if (q < blo) ++cb;
					else in = q < bhi;
				}
				const unsigned bal = __ballot_sync(FULL, in);
				__syncwarp();
				const int slot = nin + __popc(bal & ((1u << lane) - 1u));
				if (in && slot < kTsCandCap) s.cand[slot] = f2ord(q);
				nin += __popc(bal);
			}
			cb = __reduce_add_sync(FULL, cb);
			__syncwarp();
			const int kk = d.rank_slope - cb;
			if (kk >= 0 && kk < nin) {
				if (nin <= kTsCandCap) return ord2f(warp_select_kth(s.cand, nin, kk, s.hist, lane, blo, bhi));
				// the rank is inside but the bracket holds more quotients than the select scratch: zoom in (counts are exact)
				L = blo; cL = cb; U = bhi; cU = cb + nin;
				const float centre = blo + width * (((float)kk + 0.5f) / (float)nin);
				const float hw = width * ((float)kTsCandCap / (4.f * (float)nin));
				blo = fmaxf(centre - hw, L);
				bhi = fminf(centre + hw, U);
				if (!(blo < bhi)) break;
				continue;
			}
			// the counts are exact with respect to blo/bhi: tighten the enclosure, then extrapolate with the local density
			if (kk < 0) { U = blo; cU = cb; }
			else { L = bhi; cL = cb + nin; }
			if (nin >= 64 && !(L > -inf && U < inf)) {
				const float per = width / (float)nin; // slope distance per quotient around here
				const float centre = kk < 0 ? blo + ((float)kk - 0.5f) * per : bhi + ((float)(kk - nin) + 0.5f) * per;
				const float hw = 0.5f * width * fminf(1.5f, (float)(kTsCap / 3) / (float)nin);
				blo = fmaxf(centre - hw, L);
				bhi = fminf(centre + hw, U);
				if (!(blo < bhi)) break;
				continue;
			}
		}
		// next bracket: interpolate inside a two-sided enclosure (aiming at ~cap/4 quotients), else step outwards
		if (L > -inf && U < inf) {
			const float span = U - L;
			const float dens = (float)max(cU - cL, 1);
			const float centre = L + span * (((float)(d.rank_slope - cL) + 0.5f) / dens);
			const float hw = fmaxf(span * ((float)kTsCap / (8.f * dens)), 0.25f * half);
			blo = fmaxf(centre - hw, L);
			bhi = fminf(centre + hw, U);
		} else if (U < inf) { bhi = U; blo = U - 2.f * width; }
		else { blo = L; bhi = L + 2.f * width; }
		if (!(blo < bhi)) break;
	}
	sweeps += 100;
	return ts_slope_bisect(yrow, d.n, d.rank_slope, lane, L, U);
}

// Work items: with a status array, item = (window f, chain c of n_chains): the chain walks a contiguous block of the window's
// rows (row r: cols(mode) phase values at yph + f * kMaxCons + r * cols) one after the other.  Without a status array (test
// hook) every item is one dense row of fixed_cols values.
__global__ void __launch_bounds__(kTsWarps * 32, kTsCtasPerSm) k_theil_sen(const float *yph, FrameState *stv, int n_items, int n_chains, int fixed_cols, float half_k, float *ts_out)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	TsShared &s = reinterpret_cast<TsShared *>(smraw)[wid];
	for (int item = blockIdx.x * kTsWarps + wid; item < n_items; item += gridDim.x * kTsWarps) {
		int cols = fixed_cols, r0 = 0, r1 = 1, total_sweeps = 0, frame = -1;
		const float *ybase = yph + (size_t)item * fixed_cols;
		float *tbase = ts_out + (size_t)item * 3;
		if (stv) {
			const int f = item / n_chains, c = item - f * n_chains;
			if (stv[f].status != ST_OK) continue;
			frame = f;
			const ModeInfo mi = mode_info(stv[f].mode);
			const int per = (mi.rows + n_chains - 1) / n_chains;
			r0 = c * per;
			r1 = min(mi.rows, r0 + per);
			cols = mi.cols;
			ybase = yph + (size_t)f * kMaxCons;
			tbase = ts_out + (size_t)f * kMaxRows * 3;
		}
		const TsDims d(cols);
		const float gap = 0.f; // no re-centring: the Huber pilot leaves no systematic gap to the Theil-Sen slope
		for (int r = r0; r < r1; ++r) {
			const float *yrow = ybase + (size_t)r * cols;
			__syncwarp();
#if OFDMRX_TS_Y_SMEM
#pragma unroll
			for (int k = 0; k < kTsRows; ++k) {
				const int i = 32 * k + lane;
				s.y[i] = i < d.n ? yrow[i] : __int_as_float(0x7f800000);
			}
			__syncwarp();
#endif
			int sweeps = 0;
			float pilot, half;
			const float slope = ts_slope(s, yrow, d, lane, sweeps, gap, half_k, pilot, half);
			total_sweeps += sweeps;
			// intercept: upper median of y_i - slope * x_i
			__syncwarp();
			int *z = s.cand;
#pragma unroll 2
			for (int k = 0; k < d.nblk; ++k) {
				const int i = 32 * k + lane;
				if (i < d.n) z[i] = f2ord(__fsub_rn(TS_Y(i), __fmul_rn(slope, (float)(i - d.half))));
			}
			__syncwarp();
			const float yint = ord2f(warp_select_kth(z, d.n, d.rank_yint, s.hist, lane, 1.f, 0.f));
			if (lane == 0) {
				float *t = tbase + (size_t)(stv ? r : 0) * 3;
				t[0] = slope;
				t[1] = yint;
				t[2] = (float)sweeps; // diagnostic; k_soft_demap stores the row's precision here
			}
		}
		if (lane == 0 && frame >= 0) atomicAdd(&stv[frame].ts_sweeps, total_sweeps);
	}
}

// ================================================================================================ k_soft_demap
constexpr int kSdThreads = 512, kSdWarps = kSdThreads / 32;

__device__ __forceinline__ cfx derotate(cfx c, float slope, float yint, int x)
{
	const float th = -__fadd_rn(yint, __fmul_rn(slope, (float)x)); // decode.cc:493-494, x = i + code_off
	float sn, cs;
	sincosf(th, &sn, &cs);
	return cmul(c, make_float2(cs, sn));
}

__global__ void __launch_bounds__(kSdThreads) k_soft_demap(const cfx *cons_raw, const FrameState *stv, float *ts, cfx *cons_out, float *llr)
{
	__shared__ float rsp[kMaxRows], rnp[kMaxRows], prec[kMaxRows], rsl[kMaxRows], ryi[kMaxRows];
	const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (stv[f].status != ST_OK) return;
	const ModeInfo mi = mode_info(stv[f].mode);
	const int cols = mi.cols, rows = mi.rows, half = cols / 2;
	const bool qpsk = mi.mod_bits == 2;
	const cfx *cr = cons_raw + (size_t)f * kMaxCons;
	float *code = llr + (size_t)f * kCodeLen;
	float *tsf = ts + (size_t)f * kMaxRows * 3;
	if (tid < rows) { rsl[tid] = tsf[3 * tid]; ryi[tid] = tsf[3 * tid + 1]; }
	__syncthreads();
	// pass 1: signal and noise power of every row after derotation (decode.cc:507-516)
	for (int row = wid; row < rows; row += kSdWarps) {
		const float slope = rsl[row], yint = ryi[row];
		float lsp = 0.f, lnp = 0.f;
		for (int i = lane; i < cols; i += 32) {
			const cfx c = derotate(cr[row * cols + i], slope, yint, i - half);
			if (cons_out) cons_out[(size_t)f * kMaxCons + row * cols + i] = c;
			cfx m;
			psk_hard_map(c, m, qpsk);
			lsp += cnorm(m);
			lnp += cnorm(csub(c, m));
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) { lsp += __shfl_xor_sync(FULL, lsp, d); lnp += __shfl_xor_sync(FULL, lnp, d); }
		if (lane == 0) { rsp[row] = lsp; rnp[row] = lnp; }
	}
	__syncthreads();
	if (tid == 0) { // cumulative over rows, never reset (decode.cc:507)
		float sp = 0.f, np = 0.f;
		for (int row = 0; row < rows; ++row) {
			sp += rsp[row]; np += rnp[row];
			prec[row] = sp / np;
			tsf[3 * row + 2] = prec[row];
		}
	}
	__syncthreads();
	// pass 2: PhaseShiftKeying<8|4>::soft with the row's precision (psk.hh:125-130, 78-82)
	const float rcp_sqrt_2 = 0.70710678118654752440f;
	const float dist = qpsk ? 2.f * rcp_sqrt_2 : 2.f * 0.38268343236508977173f;
	for (int idx = tid; idx < rows * cols; idx += kSdThreads) {
		const int row = idx / cols, i = idx - row * cols;
		const cfx c = derotate(cr[idx], rsl[row], ryi[row], i - half);
		const float g = dist * prec[row];
		if (qpsk) {
			float *o = code + 2 * idx;
			o[0] = c.x * g;
			o[1] = c.y * g;
		} else {
			float *o = code + 3 * idx;
			o[0] = (rcp_sqrt_2 * (fabsf(c.x) - fabsf(c.y))) * g;
			o[1] = c.x * g;
			o[2] = c.y * g;
		}
	}
	// lengthen(): the trailing indices are non-frozen positions carrying a known +1 (decode.cc:245-253,529)
	for (int i = mi.cons_bits + tid; i < kCodeLen; i += kSdThreads) code[i] = 9000.f;
}

} // namespace

static int theil_sen_grid(int rows, int n_sm, int *smem)
{
	static DeviceOnce once;
	*smem = (int)(kTsWarps * sizeof(TsShared));
	if (set_dynamic_smem_once(once, k_theil_sen, *smem, true) != cudaSuccess) return -1;
	int grid = (rows + kTsWarps - 1) / kTsWarps;
	const int resident = n_sm * (int)((227 * 1024) / (*smem + 1024));
	if (grid > 4 * resident) grid = 4 * resident; // a few waves of persistent CTAs: rows differ little in cost
	return grid;
}

// first bracket of the slope search: +- half_k robust sigmas around the pilot (OFDMRX_TS_HALF overrides for A/B runs).
// Measured per 10 000 windows (demod stage, clean / README chain): 1.35e-4 26.1 / 29.4 ms, 0.9e-4 24.4 / 27.5, 0.6e-4 23.8 / 28.8:
// a narrower bracket queues fewer pairs but misses the rank more often (a second, extrapolated sweep).
static float ts_half_k()
{
	static const float k = [] { const char *e = std::getenv("OFDMRX_TS_HALF"); return e ? (float)std::atof(e) : 0.9e-4f; }();
	return k;
}

// chains per window (OFDMRX_TS_CHAINS overrides for A/B runs; 0 = by batch size)
static int ts_chains_override()
{
	static const int k = [] { const char *e = std::getenv("OFDMRX_TS_CHAINS"); return e ? std::atoi(e) : 0; }();
	return k;
}

// test hook: n_rows dense rows of `cols` phase values -> (slope, yint, sweeps) per row
cudaError_t launch_theil_sen_rows(const float *yph, int n_rows, int cols, float *ts, int n_sm, cudaStream_t s)
{
	if (n_rows <= 0) return cudaSuccess;
	int smem;
	const int grid = theil_sen_grid(n_rows, n_sm, &smem);
	if (grid < 0) return cudaErrorInvalidValue; // the shared-memory attribute could not be set on this device
	k_theil_sen<<<grid, kTsWarps * 32, smem, s>>>(yph, nullptr, n_rows, 1, cols, ts_half_k(), ts);
	return cudaGetLastError();
}

template <int S>
static cudaError_t launch_demod_fft_t(const cfx *iq, int64_t iq_stride, int iq_len, const FrameState *st, int n_frames, const cfx *tw,
	cfx *cons_raw, float *yph, cudaStream_t s)
{
	static DeviceOnce once;
	if (cudaError_t e = set_dynamic_smem_once(once, k_demod_fft<S>, (int)sizeof(FftShared<S>))) return e;
	k_demod_fft<S><<<n_frames, kDmThreads, sizeof(FftShared<S>), s>>>(iq, iq_stride, iq_len, st, tw, cons_raw, yph);
	return cudaGetLastError();
}

cudaError_t launch_demod(int rate, const cfx *iq, int64_t iq_stride, int iq_len, FrameState *st, int n_frames, const cfx *tw1280,
	cfx *cons_raw, float *yph, cfx *cons, float *ts, float *llr, int n_sm, cudaStream_t s, cudaEvent_t ev_fft_done, cudaEvent_t ev_ts_done)
{
	if (n_frames <= 0) return cudaSuccess;
	cudaError_t e = cudaSuccess;
#define OFDMRX_CALL(R) e = launch_demod_fft_t<R>(iq, iq_stride, iq_len, st, n_frames, tw1280, cons_raw, yph, s)
	OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	if (e != cudaSuccess) return e;
	if (ev_fft_done) cudaEventRecord(ev_fft_done, s);
	// chains per window: one when there are windows enough to fill the GPU twice over, more (shorter) ones for small batches
	int ts_smem;
	theil_sen_grid(1, n_sm, &ts_smem);
	const int resident_warps = n_sm * (int)((227 * 1024) / (ts_smem + 1024)) * kTsWarps;
	int n_chains = std::max(5, std::min(9, (2 * resident_warps + n_frames - 1) / n_frames)); // >= 5: short items keep the tail short
	if (ts_chains_override() > 0) n_chains = ts_chains_override();
	const int grid = theil_sen_grid(n_frames * n_chains, n_sm, &ts_smem);
	if (grid < 0) return cudaErrorInvalidValue;
	k_theil_sen<<<grid, kTsWarps * 32, ts_smem, s>>>(yph, st, n_frames * n_chains, n_chains, 0, ts_half_k(), ts);
	if (ev_ts_done) cudaEventRecord(ev_ts_done, s);
	k_soft_demap<<<n_frames, kSdThreads, 0, s>>>(cons_raw, st, ts, cons, llr);
	return cudaGetLastError();
}

} // namespace ofdmrx
