// modem_b200/csrc/demod.cu — payload symbols: FFT, differential demodulation, Theil–Sen phase line, soft demapping.
//
// Replaces, for one window per CTA (mode 6: 432 carriers x 50 rows, 8PSK):
//   data-symbol loop: 1 pilot + 50 x (mix by the frame phasor, FFT-1280, cons = X_j / X_{j-1} with erasure)  (/root/reference/decode.cc:456-477)
//   per row: 8PSK hard/map, phase error, DSP::TheilSenEstimator (exact upper median of all 93 096 pairwise
//     slopes, then of the 432 intercepts), derotation                                                      (decode.cc:479-495, psk.hh:118-139)
//   cumulative Es/N0 -> precision, PhaseShiftKeying<8>::soft -> code[3*(432 j + i) + b], lengthen()          (decode.cc:505-529, psk.hh:125-130)
// The pairwise-slope median is found without materialising or sorting the 93 096 slopes: a 256-bin histogram over
// a bracket (seeded by the 216 longest-baseline pairs) locates the bin holding rank 46 548, a second sweep counts
// what lies below and collects the bin's members as exactly rounded quotients, and the answer is selected among
// those — so the result is the exact order statistic the reference computes with std::nth_element.
#include "common.cuh"
#include "frontend.cuh"
#include "fft.cuh"

namespace ofdmrx {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kDmThreads = 448, kDmWarps = kDmThreads / 32;
constexpr int kCols = kConsCols;           // 432
constexpr int kPairs = kCols * (kCols - 1) / 2; // 93096
constexpr int kRankSlope = kPairs / 2;      // element count/2 after nth_element
constexpr int kRankYint = kCols / 2;
constexpr int kBins = 256, kCandCap = 4096, kBinCap = 2048, kMineCap = 40;

struct DmShared {
	cfx buf0[kSymLen];
	cfx buf1[kSymLen];
	cfx prev[kCols];
	cfx cons[kCols];
	float y[kCols + 8];  // 8 entries of +inf padding: overshoot of the unrolled pair loops never counts
	float rcp[kCols];    // 1/d for d = 1..431 (index 0 unused)
	float rcpfar[kCols]; // 1/d for d >= 217, NaN below: pairs closer than 217 belong to the near loop of the other thread
	float z[kCols];
	float cand[kCandCap];
	int hist[kBins];
	float red[kDmWarps][2];
	int wcnt[kDmWarps];
	double redd[kDmWarps][2];
	int sel_lo, sel_k, sel_cnt, ncand2;
	int small[32];
	unsigned short mine[kMineCap][kCols]; // per-thread list of pairs whose approximate slope falls in the bracket
	int below, ncand, sel_bin, state;
	int cmin, cmax; // ordered-int images of the smallest / largest collected quotient
	float lo, hi, blo, bhi, result;
	float q1, q2, q3;
};

__device__ __forceinline__ int f2ord(float v) { const int i = __float_as_int(v); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ void psk8_hard_map(cfx c, cfx &m)
{
	const float cos_pi_8 = 0.92387953251128675613f, sin_pi_8 = 0.38268343236508977173f;
	const bool swap = fabsf(c.x) < fabsf(c.y);
	const float re = swap ? sin_pi_8 : cos_pi_8, im = swap ? cos_pi_8 : sin_pi_8;
	m = make_float2(c.x < 0.f ? -re : re, c.y < 0.f ? -im : im);
}

// pair {i, (i + dx) mod 432} ordered by index as the reference forms it: diff = y_hi - y_lo, dist = x_hi - x_lo;
// rd/rw = 1/dx and 1/(432-dx) (uniform per iteration)
__device__ __forceinline__ void pair_terms(const float *y, float yi, int i, int dx, float rd, float rw, float &diff, int &dist, float &rcp)
{
	int j = i + dx;
	if (j >= kCols) { j -= kCols; diff = yi - y[j]; dist = kCols - dx; rcp = rw; }
	else { diff = y[j] - yi; dist = dx; rcp = rd; }
}

// warp 0: locate rank k inside a 256-bin histogram without a serial scan: 8 bins per lane, shuffle prefix.
// Writes s.sel_bin (bin holding rank k, or -1 if k >= total), s.sel_k (rank inside that bin), s.sel_cnt (its count).
__device__ __forceinline__ void find_bin(DmShared &s, int k, int tid)
{
	if (tid >= 32) return;
	const int lane = tid;
	int h[8], sum = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) { h[i] = s.hist[lane * 8 + i]; sum += h[i]; }
	int incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
	const int excl = incl - sum;
	const bool mine = k >= excl && k < incl;
	const unsigned bal = __ballot_sync(FULL, mine);
	if (bal == 0u) { if (lane == 0) { s.sel_bin = -1; s.sel_k = k - __shfl_sync(FULL, incl, 31); s.sel_cnt = 0; } return; }
	if (mine) {
		int kk = k - excl, b = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i) { if (b == i && kk >= h[i]) { kk -= h[i]; b = i + 1; } }
		s.sel_bin = lane * 8 + b;
		s.sel_k = kk;
		s.sel_cnt = h[b < 8 ? b : 7];
	}
}

// k-th smallest (0-based) of v[0..n) in shared memory, exact: radix select on the order-preserving integer image of
// the floats, 8 bits per level inside the [min,max] range, 256-bin shared histogram; as soon as the selected bin
// holds <= 32 values they are gathered and ranked by one warp.  All threads of the CTA call.
__device__ float select_kth(DmShared &s, const float *v, int n, int k, int tid)
{
	const int lane = tid & 31;
	int mn = 0x7fffffff, mx = (int)0x80000000;
	for (int i = tid; i < n; i += kDmThreads) { const int o = f2ord(v[i]); mn = min(mn, o); mx = max(mx, o); }
#pragma unroll
	for (int d = 16; d; d >>= 1) { mn = min(mn, __shfl_xor_sync(FULL, mn, d)); mx = max(mx, __shfl_xor_sync(FULL, mx, d)); }
	if (tid == 0) { s.cmin = 0x7fffffff; s.cmax = (int)0x80000000; }
	__syncthreads();
	if (lane == 0) { atomicMin(&s.cmin, mn); atomicMax(&s.cmax, mx); }
	__syncthreads();
	int lo = s.cmin, kk = k;
	int sh = max(0, 32 - __clz(((unsigned)s.cmax - (unsigned)s.cmin) | 1u) - 8);
	for (int level = 0; level < 5; ++level) {
		for (int b = tid; b < kBins; b += kDmThreads) s.hist[b] = 0;
		if (tid == 0) s.ncand2 = 0;
		__syncthreads();
		for (int i = tid; i < n; i += kDmThreads) {
			const unsigned b = ((unsigned)f2ord(v[i]) - (unsigned)lo) >> sh; // values below lo wrap to huge bins
			if (b < (unsigned)kBins) atomicAdd(&s.hist[b], 1);
		}
		__syncthreads();
		find_bin(s, kk, tid);
		__syncthreads();
		const int b = s.sel_bin, cnt = s.sel_cnt;
		kk = s.sel_k;
		lo += (int)((unsigned)b << sh);
		if (sh == 0) break;               // bins are single values: lo is the answer
		if (cnt <= 32) {                  // finish: gather the bin's members, rank them in one warp
			for (int i = tid; i < n; i += kDmThreads) {
				const int o = f2ord(v[i]);
				if ((((unsigned)o - (unsigned)lo) >> sh) == 0u) s.small[atomicAdd(&s.ncand2, 1) & 31] = o;
			}
			__syncthreads();
			if (tid < 32) {
				const int m = s.ncand2;
				const int mine = lane < m ? s.small[lane] : 0x7fffffff;
				int r = 0;
				for (int j = 0; j < m; ++j) { const int o = __shfl_sync(FULL, mine, j); r += (o < mine) || (o == mine && j < lane); }
				if (lane < m && r == kk) s.sel_lo = mine;
			}
			__syncthreads();
			return ord2f(s.sel_lo);
		}
		sh = max(0, sh - 8);
	}
	return ord2f(lo);
}

// pilot bracket: ordinary least squares slope c of y on x = i - 216 and the residual standard deviation.  The exact
// Theil–Sen median lies within a few 1e-4 * sigma of c for Gaussian-like phase noise (its efficiency relative to OLS is
// 0.955), so a sweep over [c - d, c + d) with d = 3.5e-4 sigma usually contains rank 46 548 and <= ~2500 quotients.
__device__ void ols_pilot(DmShared &s, int tid, float &c, float &sigma)
{
	const int lane = tid & 31, wid = tid >> 5;
	const bool act = tid < kCols;
	const double x = (double)(tid - kCols / 2) + 0.5; // centred abscissa
	const double yv = act ? (double)s.y[tid] : 0.0;
	double a0 = yv, a1 = act ? x * yv : 0.0;
#pragma unroll
	for (int d = 16; d; d >>= 1) { a0 += __shfl_xor_sync(FULL, a0, d); a1 += __shfl_xor_sync(FULL, a1, d); }
	if (lane == 0) { s.redd[wid][0] = a0; s.redd[wid][1] = a1; }
	__syncthreads();
	double sy = 0.0, sxy = 0.0;
	for (int w = 0; w < kDmWarps; ++w) { sy += s.redd[w][0]; sxy += s.redd[w][1]; }
	const double sxx = (double)kCols * ((double)kCols * kCols - 1.0) / 12.0;
	const double slope = sxy / sxx, mean = sy / kCols;
	__syncthreads();
	double r = act ? yv - mean - slope * x : 0.0;
	double r2 = r * r;
#pragma unroll
	for (int d = 16; d; d >>= 1) r2 += __shfl_xor_sync(FULL, r2, d);
	if (lane == 0) s.redd[wid][0] = r2;
	__syncthreads();
	double ss = 0.0;
	for (int w = 0; w < kDmWarps; ++w) ss += s.redd[w][0];
	__syncthreads();
	c = (float)slope;
	sigma = (float)sqrt(ss / (kCols - 2));
}

// exact upper median of the pairwise slopes of (x = i - 216, y[i]); all threads of the CTA call
__device__ float theil_sen_slope(DmShared &s, int tid)
{
	const int lane = tid & 31;
	const bool act = tid < kCols;
	const float yi = act ? s.y[tid] : 0.f;
	// ---- fast path: exact sweeps over the OLS pilot bracket; if the rank falls just outside, slide the bracket
	float c, sigma;
	ols_pilot(s, tid, c, sigma);
	const float dlt = fmaxf(3.5e-4f * sigma, fmaxf(fabsf(c) * 4e-6f, 1e-10f));
	float blo = c - dlt, bhi = c + dlt;
	for (int attempt = 0; attempt < 5; ++attempt) {
		// approximate slopes (diff * 1/d) are within 2 ulp of the exact quotient: anything within mg of the bracket is
		// re-evaluated exactly, the rest is classified by the approximation
		const float mg = 2e-6f * fmaxf(fabsf(blo), fabsf(bhi)) + 1e-30f;
		const float blo_m = blo - mg, bhi_m = bhi + mg;
		if (tid == 0) { s.below = 0; s.ncand = 0; s.state = 0; }
		__syncthreads();
		int cb = 0, cnt = 0, nin = 0;
		const float *yp = s.y + tid;
		// exact quotient of one remembered pair: code = d (near pair (tid, tid+d)) or 0x8000|j (far pair (j, tid))
		auto exact_q = [&](int code) {
			float diff; int dist;
			if (code & 0x8000) { const int j = code & 0x3fff; diff = yi - s.y[j]; dist = tid - j; }
			else { diff = yp[code & 0x3fff] - yi; dist = code & 0x3fff; }
			return __fdiv_rn(diff, (float)dist);
		};
		if (act) {
			// phase 1: classify by the approximate slope only; remember the few pairs near/inside the bracket.
			// Thread i owns the pairs (i, i+d), d <= min(216, 431-i), and the far pairs (j, i), i-j >= 217: every unordered
			// pair exactly once, 215 or 216 per thread, and no wrap-around logic inside the loops.
			const int nA = min(216, kCols - 1 - tid);
			for (int d0 = 1; d0 <= nA; d0 += 8) { // 216 = 27 * 8; shorter rows run into the +inf padding
				unsigned hits = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const float sl = (yp[d0 + k] - yi) * s.rcp[d0 + k];
					cb += sl < blo_m;
					hits |= (unsigned)((sl >= blo_m) & (sl < bhi_m)) << k;
				}
				while (hits) {
					const int k = __ffs(hits) - 1;
					hits &= hits - 1;
					if (cnt < kMineCap) s.mine[cnt][tid] = (unsigned short)(d0 + k);
					++cnt;
				}
			}
			const float *rp = s.rcpfar + tid;
			for (int j0 = 0; j0 <= tid - 217; j0 += 8) { // overshoot hits the NaN part of rcpfar: never counted
				unsigned hits = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const float sl = (yi - s.y[j0 + k]) * rp[-(j0 + k)];
					cb += sl < blo_m;
					hits |= (unsigned)((sl >= blo_m) & (sl < bhi_m)) << k;
				}
				while (hits) {
					const int k = __ffs(hits) - 1;
					hits &= hits - 1;
					if (cnt < kMineCap) s.mine[cnt][tid] = (unsigned short)(0x8000 | (j0 + k));
					++cnt;
				}
			}
			if (cnt > kMineCap) s.state = 1; // list overflow (not seen in practice): take the general path
			// phase 2a: exact quotients of the remembered pairs; mark those inside the bracket (bit 14)
			const int m = min(cnt, kMineCap);
			for (int k = 0; k < m; ++k) {
				const int code = s.mine[k][tid];
				const float q = exact_q(code);
				if (q < blo) ++cb;
				else if (q < bhi) { s.mine[k][tid] = (unsigned short)(code | 0x4000); ++nin; }
			}
		}
		// slots for the survivors: block-wide exclusive scan of the per-thread counts (no atomics on a single counter)
		int incl = nin;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
		if (lane == 31) s.wcnt[tid >> 5] = incl;
#pragma unroll
		for (int d = 16; d; d >>= 1) cb += __shfl_xor_sync(FULL, cb, d);
		if (lane == 0 && cb) atomicAdd(&s.below, cb);
		__syncthreads();
		int off = incl - nin, total = 0;
		for (int w = 0; w < kDmWarps; ++w) { if (w < (tid >> 5)) off += s.wcnt[w]; total += s.wcnt[w]; }
		if (act && total <= kCandCap) {
			const int m = min(cnt, kMineCap);
			for (int k = 0; k < m; ++k) {
				const int code = s.mine[k][tid];
				if (code & 0x4000) s.cand[off++] = exact_q(code);
			}
		}
		if (tid == 0) s.ncand = total;
		__syncthreads();
		const int ovf = s.state;
		const int kk = kRankSlope - s.below, nc = s.ncand;
		__syncthreads();
		if (ovf) break;
		if (kk >= 0 && kk < nc && nc <= kCandCap) return select_kth(s, s.cand, nc, kk, tid);
		if (nc > kCandCap) break;
		// the counts are exact with respect to blo/bhi, so the neighbouring bracket is the next place to look
		const float w = (bhi - blo) * (float)(2 << attempt);
		if (kk < 0) { bhi = blo; blo = blo - w; }
		else { blo = bhi; bhi = bhi + w; }
	}
	// ---- general path (pilot missed: outliers, erased rows, very low SNR)
	// seed bracket: quartiles of the 216 slopes with baseline 216
	if (tid < 216) s.z[tid] = (s.y[tid + 216] - s.y[tid]) / 216.f;
	__syncthreads();
	if (tid < 216) {
		const float v = s.z[tid];
		int r = 0;
		for (int j = 0; j < 216; ++j) { const float o = s.z[j]; r += (o < v) || (o == v && j < tid); }
		if (r == 54) s.q1 = v;
		if (r == 108) s.q2 = v;
		if (r == 162) s.q3 = v;
	}
	__syncthreads();
	if (tid == 0) {
		float hw = 0.5f * (s.q3 - s.q1);
		hw = fmaxf(hw, fmaxf(fabsf(s.q2) * 1e-5f, 1e-9f));
		s.lo = s.q2 - hw;
		s.hi = s.q2 + hw;
		s.state = 0;
	}
	__syncthreads();
	for (int iter = 0; iter < 64; ++iter) {
		// ---- histogram sweep over [lo, hi) with approximate slopes
		for (int b = tid; b < kBins; b += kDmThreads) s.hist[b] = 0;
		if (tid == 0) s.below = 0;
		__syncthreads();
		const float lo = s.lo, hi = s.hi;
		const float inv_w = (float)kBins / (hi - lo);
		int below = 0;
		if (act) {
			for (int dx = 1; dx <= 216; ++dx) {
				if (dx == 216 && tid >= 216) break;
				float diff, rc; int dist;
				pair_terms(s.y, yi, tid, dx, s.rcp[dx], s.rcp[kCols - dx], diff, dist, rc);
				const float sl = diff * rc;
				if (sl < lo) ++below;
				else if (sl < hi) {
					int b = (int)((sl - lo) * inv_w);
					b = min(max(b, 0), kBins - 1);
					atomicAdd(&s.hist[b], 1);
				}
			}
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) below += __shfl_xor_sync(FULL, below, d);
		if (lane == 0 && below) atomicAdd(&s.below, below);
		__syncthreads();
		find_bin(s, kRankSlope - s.below, tid);
		__syncthreads();
		if (tid == 0) {
			const int bsel = s.sel_bin;
			const float w = (hi - lo) / (float)kBins;
			if (kRankSlope < s.below) { // rank lies below the bracket: slide down and widen
				s.hi = lo; s.lo = lo - 16.f * (hi - lo); s.state = 0;
			} else if (bsel < 0) { // above the bracket
				s.lo = hi; s.hi = hi + 16.f * (hi - lo); s.state = 0;
			} else {
				const float blo = lo + (float)bsel * w, bhi = bsel == kBins - 1 ? hi : lo + (float)(bsel + 1) * w;
				if (s.sel_cnt > kBinCap && bhi > blo && (bhi - blo) > 1e-30f) { s.lo = blo; s.hi = bhi; s.state = 0; }
				else { s.blo = blo; s.bhi = bhi; s.state = 1; }
			}
		}
		__syncthreads();
		if (s.state == 0) continue;
		// ---- exact sweep: count quotients below blo, collect those in [blo, bhi)
		for (int widen = 0; widen < 4; ++widen) {
			if (tid == 0) { s.below = 0; s.ncand = 0; s.cmin = 0x7fffffff; s.cmax = (int)0x80000000; }
			__syncthreads();
			const float blo = s.blo, bhi = s.bhi;
			const float mg = 2e-6f * fmaxf(fabsf(blo), fabsf(bhi)) + 1e-30f;
			const float blo_m = blo - mg, bhi_m = bhi + mg;
			int cb = 0;
			if (act) {
				for (int dx = 1; dx <= 216; ++dx) {
					if (dx == 216 && tid >= 216) break;
					float diff, rc; int dist;
					pair_terms(s.y, yi, tid, dx, s.rcp[dx], s.rcp[kCols - dx], diff, dist, rc);
					const float sl = diff * rc;
					if (sl < blo_m) ++cb;
					else if (sl < bhi_m) {
						const float q = __fdiv_rn(diff, (float)dist);
						if (q < blo) ++cb;
						else if (q < bhi) {
							const int p = atomicAdd(&s.ncand, 1);
							if (p < kCandCap) s.cand[p] = q;
							else { atomicMin(&s.cmin, f2ord(q)); atomicMax(&s.cmax, f2ord(q)); }
						}
					}
				}
			}
#pragma unroll
			for (int d = 16; d; d >>= 1) cb += __shfl_xor_sync(FULL, cb, d);
			if (lane == 0 && cb) atomicAdd(&s.below, cb);
			__syncthreads();
			const int kk = kRankSlope - s.below, nc = s.ncand;
			if (kk >= 0 && kk < nc && nc <= kCandCap) { __syncthreads(); return select_kth(s, s.cand, nc, kk, tid); }
			if (kk >= 0 && kk < nc && nc > kCandCap) {
				// overflow: if every quotient of the bin is the same value (erased rows: all phases equal) that value is the answer
				for (int t = tid; t < kCandCap; t += kDmThreads) { atomicMin(&s.cmin, f2ord(s.cand[t])); atomicMax(&s.cmax, f2ord(s.cand[t])); }
				__syncthreads();
				if (s.cmin == s.cmax) return ord2f(s.cmin);
			}
			__syncthreads();
			if (nc > kCandCap) break; // too crowded: go back to the histogram loop with this bin as the bracket
			if (tid == 0) { // the approximate bin edges missed the rank by a few elements: take one more bin either side
				const float w = s.bhi - s.blo;
				s.blo -= w; s.bhi += w;
			}
			__syncthreads();
		}
		if (tid == 0) { s.lo = s.blo; s.hi = s.bhi; s.state = 0; }
		__syncthreads();
	}
	return s.q2; // not reached for finite inputs; keeps the kernel total
}

__global__ void __launch_bounds__(kDmThreads) k_demod(const cfx *iq, int64_t iq_stride, int iq_len, const FrameState *stv,
	const cfx *tw1280, cfx *cons_raw, cfx *cons_out, float *ts_out, float *llr)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	DmShared &s = *reinterpret_cast<DmShared *>(smraw);
	const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const FrameState &st = stv[f];
	if (st.status != ST_OK) return;
	const cfx *a = iq + (size_t)f * iq_stride;
	float *code = llr + (size_t)f * kCodeLen;
	const int p0 = st.sc_pos + 2 * kPitch; // pilot body (decode.cc:456-459)
	const double turns = -(double)st.cfo_rad / 6.283185307179586476925286766559;
	if (tid < kCols) {
		s.rcp[tid] = tid ? __frcp_rn((float)tid) : 0.f;
		s.rcpfar[tid] = tid >= 217 ? __frcp_rn((float)tid) : __int_as_float(0x7fc00000);
	}
	if (tid < 8) s.y[kCols + tid] = __int_as_float(0x7f800000);
	float sp = 0.f, np = 0.f; // cumulative, never reset (decode.cc:507)
	for (int sym = 0; sym <= kConsRows; ++sym) {
		const int w0 = p0 + kPitch * sym;
		const int n0 = kSymLen + kPitch * sym; // phasor steps since the header symbol (decode.cc:404-405,459-461,468-470)
		for (int i = tid; i < kSymLen; i += kDmThreads) {
			const int idx = w0 + i;
			const cfx v = (idx >= 0 && idx < iq_len) ? a[idx] : make_float2(0.f, 0.f);
			s.buf0[i] = cmul(v, phasor_turns(turns * (double)(n0 + i)));
		}
		__syncthreads();
		fft_fwd<kSymLen>(s.buf0, s.buf1, tw1280, tid, kDmThreads);
		if (sym == 0) {
			if (tid < kCols) s.prev[tid] = s.buf1[(tid - kCols / 2 + kSymLen) % kSymLen];
			__syncthreads();
			continue;
		}
		const int row = sym - 1;
		cfx c = make_float2(0.f, 0.f);
		if (tid < kCols) {
			const cfx cur = s.buf1[(tid - kCols / 2 + kSymLen) % kSymLen];
			c = demod_or_erase(cur, s.prev[tid]);
			s.prev[tid] = cur;
			if (cons_raw) cons_raw[((size_t)f * kConsRows + row) * kCols + tid] = c;
			cfx m;
			psk8_hard_map(c, m);
			const cfx e = cmulc(c, m);
			s.y[tid] = atan2f(e.y, e.x);
		}
		__syncthreads();
		const float slope = theil_sen_slope(s, tid);
		// intercept: upper median of y_i - slope * x_i (theil_sen.hh, recalled)
		if (tid < kCols) s.z[tid] = __fsub_rn(s.y[tid], __fmul_rn(slope, (float)(tid - kCols / 2)));
		__syncthreads();
		const float yint = select_kth(s, s.z, kCols, kRankYint, tid);
		float lsp = 0.f, lnp = 0.f;
		if (tid < kCols) {
			const float th = -__fadd_rn(yint, __fmul_rn(slope, (float)(tid - kCols / 2)));
			float sn, cs;
			sincosf(th, &sn, &cs);
			c = cmul(c, make_float2(cs, sn));
			s.cons[tid] = c;
			if (cons_out) cons_out[((size_t)f * kConsRows + row) * kCols + tid] = c;
			cfx m;
			psk8_hard_map(c, m);
			lsp = cnorm(m);
			lnp = cnorm(csub(c, m));
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) { lsp += __shfl_xor_sync(FULL, lsp, d); lnp += __shfl_xor_sync(FULL, lnp, d); }
		if (lane == 0) { s.red[wid][0] = lsp; s.red[wid][1] = lnp; }
		__syncthreads();
		for (int w2 = 0; w2 < kDmWarps; ++w2) { sp += s.red[w2][0]; np += s.red[w2][1]; }
		const float precision = sp / np;
		if (tid == 0 && ts_out) {
			float *t = ts_out + ((size_t)f * kConsRows + row) * 3;
			t[0] = slope; t[1] = yint; t[2] = precision;
		}
		if (tid < kCols) {
			const float rcp_sqrt_2 = 0.70710678118654752440f, DIST = 2.f * 0.38268343236508977173f;
			const float g = DIST * precision;
			float *o = code + 3 * (kCols * row + tid);
			o[0] = (rcp_sqrt_2 * (fabsf(c.x) - fabsf(c.y))) * g;
			o[1] = c.x * g;
			o[2] = c.y * g;
		}
		__syncthreads();
	}
	// lengthen(): the 736 trailing indices are non-frozen positions carrying a known +1 (decode.cc:245-253,529)
	for (int i = kConsBits + tid; i < kCodeLen; i += kDmThreads) code[i] = 9000.f;
}

} // namespace

cudaError_t launch_demod(const cfx *iq, int64_t iq_stride, int iq_len, const FrameState *st, int n_frames, const cfx *tw1280,
	cfx *cons_raw, cfx *cons, float *ts_out, float *llr, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
	static bool attr = false;
	if (!attr) {
		cudaFuncSetAttribute(k_demod, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DmShared));
		attr = true;
	}
	k_demod<<<n_frames, kDmThreads, sizeof(DmShared), s>>>(iq, iq_stride, iq_len, st, tw1280, cons_raw, cons, ts_out, llr);
	return cudaGetLastError();
}

} // namespace ofdmrx
