// modem_b200/csrc/acquire.cu — per-window acquisition: fine Schmidl-Cox stage and frame header.
//
// Replaces, per accepted detection:
//   SchmidlCox::operator() fine part: derotate, FFT-640, bin-differential demod, MLS0 cross-correlation via
//     FFT * kern -> IFFT, peak/runner-up test, pos_err, cfo_rad                       (/root/reference/decode.cc:110-151)
//   header: mix by -cfo_rad, FFT-1280, MLS1 descramble, int8 soft bits               (decode.cc:403-416)
//   CODE::OrderedStatisticsDecoder<255,71,4> on the BCH(255,71) generator            (decode.cc:199,378-384,417)
//   CRC-16 / mode / call-sign checks and the SKIP loop                                (decode.cc:390-448)
// One CTA per window.  The OSD keeps the reference's result exactly (winner + `unique`) but replaces the literal
// 1 031 347-candidate sweep by a branch and bound over bit-packed rows (see oracle/ref_code.hh decode_pruned for
// the argument): on clean or moderately noisy headers only the order-0 candidate is evaluated.
#include "common.cuh"
#include "frontend.cuh"
#include "fft.cuh"

namespace ofdmrx {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kAcqThreads = 320, kAcqWarps = kAcqThreads / 32;
constexpr int K = kHdrK, NB = kHdrBits;

struct AcqShared { // followed in dynamic shared memory by the two FFT buffers of symbol_len values each
	uint32_t rows[K][8];
	short lut[32][256];
	int soft[256];
	int perm[256];
	int w[256];
	unsigned char hbits[256];
	unsigned char rowflag[K + 1];
	float redf[kAcqWarps][4];
	int redi[kAcqWarps][2];
	uint32_t c0[8];
	unsigned long long best;
	int ties, piv, col, bestD, m0, wneg;
	float bc_f[4];
	int bc_i[4];
};

__device__ __forceinline__ int lut_dist(const AcqShared &s, const uint32_t *e)
{
	int d = 0;
#pragma unroll
	for (int p = 0; p < 32; ++p) d += s.lut[p][(e[p >> 2] >> (8 * (p & 3))) & 255u];
	return d;
}
__device__ __forceinline__ unsigned long long osd_key(int D, int a, int b, int c, int d)
{
	return ((unsigned long long)(D + 32768) << 28) | ((unsigned long long)a << 21) | ((unsigned long long)b << 14) |
		((unsigned long long)c << 7) | (unsigned long long)d;
}

// Enumerate the <= 4-flip candidates under the bound  sum(w of flipped basis bits) + wneg <= best.
// COUNT = false: minimise (D, code) into s.best.  COUNT = true: count candidates with D == s.bestD into s.ties.
template <bool COUNT>
__device__ void osd_search(AcqShared &s, int tid)
{
	const int wneg = s.wneg;
	volatile unsigned long long *vbest = &s.best;
	for (int id = tid; id < K * K; id += kAcqThreads) {
		const int a = id / K, b = id - a * K;
		if (b < a) continue;
		int bestD = COUNT ? s.bestD : (int)(*vbest >> 28) - 32768;
		const int wa = s.w[a];
		if (wa + wneg > bestD) continue;
		uint32_t e[8];
		if (b == a) {
#pragma unroll
			for (int i = 0; i < 8; ++i) e[i] = s.rows[a][i];
			const int D = lut_dist(s, e);
			if (COUNT) { if (D == bestD) atomicAdd(&s.ties, 1); }
			else if (D <= bestD) atomicMin(&s.best, osd_key(D, a, 127, 127, 127));
			continue;
		}
		const int wab = wa + s.w[b];
		if (wab + wneg > bestD) continue;
#pragma unroll
		for (int i = 0; i < 8; ++i) e[i] = s.rows[a][i] ^ s.rows[b][i];
		{
			const int D = lut_dist(s, e);
			if (COUNT) { if (D == bestD) atomicAdd(&s.ties, 1); }
			else if (D <= bestD) atomicMin(&s.best, osd_key(D, a, b, 127, 127));
		}
		for (int c = b + 1; c < K; ++c) {
			if (!COUNT) bestD = (int)(*vbest >> 28) - 32768;
			const int wabc = wab + s.w[c];
			if (wabc + wneg > bestD) continue;
			uint32_t e3[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) e3[i] = e[i] ^ s.rows[c][i];
			{
				const int D = lut_dist(s, e3);
				if (COUNT) { if (D == bestD) atomicAdd(&s.ties, 1); }
				else if (D <= bestD) atomicMin(&s.best, osd_key(D, a, b, c, 127));
			}
			for (int d = c + 1; d < K; ++d) {
				if (wabc + s.w[d] + wneg > bestD) continue;
				uint32_t e4[8];
#pragma unroll
				for (int i = 0; i < 8; ++i) e4[i] = e3[i] ^ s.rows[d][i];
				const int D = lut_dist(s, e4);
				if (COUNT) { if (D == bestD) atomicAdd(&s.ties, 1); }
				else if (D <= bestD) { atomicMin(&s.best, osd_key(D, a, b, c, d)); bestD = min(bestD, D); }
			}
		}
	}
}

// OSD on s.soft[0..254]; leaves hard bits in s.hbits[0..254]; returns `unique` (decode.cc:417).
// Every thread of the CTA must call; control flow is CTA-uniform.
__device__ bool osd_decode(AcqShared &s, const uint32_t *gen_rows, int tid)
{
	const int lane = tid & 31;
	// reliability order: stable, descending |max(soft,-127)|
	if (tid < NB) {
		const int ri = abs(max(s.soft[tid], -127));
		int r = 0;
		for (int j = 0; j < NB; ++j) {
			const int rj = abs(max(s.soft[j], -127));
			r += (rj > ri) || (rj == ri && j < tid);
		}
		s.perm[r] = tid;
	}
	__syncthreads();
	for (int idx = tid; idx < K * 8; idx += kAcqThreads) {
		const int row = idx >> 3, wd = idx & 7;
		uint32_t word = 0;
		for (int b = 0; b < 32; ++b) {
			const int col = wd * 32 + b;
			if (col < NB) {
				const int pc = s.perm[col];
				word |= ((__ldg(&gen_rows[row * 8 + (pc >> 5)]) >> (pc & 31)) & 1u) << b;
			}
		}
		s.rows[row][wd] = word;
	}
	__syncthreads();
	// Gauss-Jordan with the reference's pivoting: first row >= k with a one in column k; if there is none, the
	// first later column with a one in some row >= k is swapped in (osd.hh row_echelon()/systematic(), recalled).
	bool singular = false;
	for (int k = 0; k < K; ++k) {
		for (int attempt = 0; attempt < 2; ++attempt) {
			if (tid < 32) {
				int found = 1 << 30;
				for (int j = k + lane; j < K; j += 32)
					if ((s.rows[j][k >> 5] >> (k & 31)) & 1u) { found = j; break; }
#pragma unroll
				for (int d = 16; d; d >>= 1) found = min(found, __shfl_xor_sync(FULL, found, d));
				if (lane == 0) { s.piv = found < K ? found : -1; s.col = 1 << 30; }
			}
			__syncthreads();
			if (s.piv >= 0 || attempt == 1) break;
			if (tid > k && tid < NB) {
				bool has = false;
				for (int h = k; h < K && !has; ++h) has = (s.rows[h][tid >> 5] >> (tid & 31)) & 1u;
				if (has) atomicMin(&s.col, tid);
			}
			__syncthreads();
			const int cj = s.col;
			if (cj < NB) {
				if (tid < K) {
					const uint32_t bk = (s.rows[tid][k >> 5] >> (k & 31)) & 1u, bj = (s.rows[tid][cj >> 5] >> (cj & 31)) & 1u;
					if (bk != bj) { s.rows[tid][k >> 5] ^= 1u << (k & 31); s.rows[tid][cj >> 5] ^= 1u << (cj & 31); }
				}
				if (tid == 0) { const int t = s.perm[k]; s.perm[k] = s.perm[cj]; s.perm[cj] = t; }
			}
			__syncthreads();
		}
		const int piv = s.piv;
		if (piv < 0) { singular = true; break; } // cannot happen for a rank-71 generator
		if (piv != k && tid < 8) { const uint32_t t = s.rows[k][tid]; s.rows[k][tid] = s.rows[piv][tid]; s.rows[piv][tid] = t; }
		__syncthreads();
		if (tid < K) s.rowflag[tid] = tid != k && ((s.rows[tid][k >> 5] >> (k & 31)) & 1u);
		__syncthreads();
		for (int idx = tid; idx < K * 8; idx += kAcqThreads) {
			const int row = idx >> 3, wd = idx & 7;
			if (s.rowflag[row]) s.rows[row][wd] ^= s.rows[k][wd];
		}
		__syncthreads();
	}
	if (singular) return false;
	// order-0 codeword, weights w_i = (1-2 c0_i) softperm_i, M0, sum of negative parity weights
	if (tid < 8) {
		uint32_t acc = 0;
		for (int j = 0; j < K; ++j) {
			const int sp = max(s.soft[s.perm[j]], -127);
			if (sp < 0) acc ^= s.rows[j][tid];
		}
		s.c0[tid] = acc;
	}
	__syncthreads();
	if (tid < 256) {
		int wv = 0;
		if (tid < NB) {
			const int sp = max(s.soft[s.perm[tid]], -127);
			const int cb = (s.c0[tid >> 5] >> (tid & 31)) & 1u;
			wv = (1 - 2 * cb) * sp;
		}
		s.w[tid] = wv;
	}
	__syncthreads();
	if (tid < 32) {
		int m0 = 0, wn = 0;
		for (int i = lane; i < 256; i += 32) { const int wv = s.w[i]; m0 += wv; if (i >= K && wv < 0) wn += wv; }
#pragma unroll
		for (int d = 16; d; d >>= 1) { m0 += __shfl_xor_sync(FULL, m0, d); wn += __shfl_xor_sync(FULL, wn, d); }
		if (lane == 0) { s.m0 = m0; s.wneg = wn; s.best = osd_key(0, 127, 127, 127, 127); s.ties = 0; }
	}
	for (int idx = tid; idx < 32 * 256; idx += kAcqThreads) {
		const int p = idx >> 8, v = idx & 255;
		int acc = 0;
#pragma unroll
		for (int b = 0; b < 8; ++b) if ((v >> b) & 1) acc += s.w[8 * p + b];
		s.lut[p][v] = (short)acc;
	}
	__syncthreads();
	osd_search<false>(s, tid);
	__syncthreads();
	if (tid == 0) { s.bestD = (int)(s.best >> 28) - 32768; s.ties = s.bestD == 0 ? 1 : 0; }
	__syncthreads();
	osd_search<true>(s, tid);
	__syncthreads();
	// winner -> hard bits at their original positions
	const unsigned long long bk = s.best;
	const int sel[4] = {(int)(bk >> 21) & 127, (int)(bk >> 14) & 127, (int)(bk >> 7) & 127, (int)bk & 127};
	if (tid < NB) {
		uint32_t bit = (s.c0[tid >> 5] >> (tid & 31)) & 1u;
#pragma unroll
		for (int q = 0; q < 4; ++q)
			if (sel[q] < K) bit ^= (s.rows[sel[q]][tid >> 5] >> (tid & 31)) & 1u;
		s.hbits[s.perm[tid]] = (unsigned char)bit;
	}
	__syncthreads();
	const int bestM = s.m0 - 2 * s.bestD;
	return s.ties == 1 && bestM != -1;
}

__device__ __forceinline__ cfx load_iq(const cfx *a, int idx, int iq_len)
{
	return (idx >= 0 && idx < iq_len) ? a[idx] : make_float2(0.f, 0.f);
}

constexpr size_t kAcqBufOff = (sizeof(AcqShared) + 15) & ~(size_t)15;

template <int S>
__global__ void __launch_bounds__(kAcqThreads, 2) k_acquire(const cfx *iq, int64_t iq_stride, int iq_len, const Detection *det,
	const int32_t *det_count, int det_cap, int skip, FrameState *stv, int8_t *soft_out, AcquireConsts ac)
{
	// geometry of this sample rate (shadows the 8 kHz constants of host_tables.h)
	constexpr int kSymLen = Geo<S>::kSymLen, kHalf = Geo<S>::kHalf, kPitch = Geo<S>::kPitch, kGuardLen = Geo<S>::kGuardLen;
	constexpr int kBufferLen = Geo<S>::kBufferLen, kSearchPos = Geo<S>::kSearchPos, kMatchDel = Geo<S>::kMatchDel;
	constexpr int kOffOld = kBufferLen - 1 - (kSearchPos + kHalf), kOffCur = kBufferLen - 1 - (kSearchPos + kSymLen); // 5119, 4479
	extern __shared__ __align__(16) unsigned char smraw[];
	AcqShared &s = *reinterpret_cast<AcqShared *>(smraw);
	cfx *const buf0 = reinterpret_cast<cfx *>(smraw + kAcqBufOff), *const buf1 = buf0 + kSymLen;
	const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const cfx *a = iq + (size_t)f * iq_stride;
	FrameState &st = stv[f];
	const int nd = min(det_count[f] & (kDetOverflowBit - 1), det_cap);
	const int det_overflow = (det_count[f] & kDetOverflowBit) ? 1 : 0;
	int status = ST_NO_SYNC, skip_left = skip, accepted = 0;
	bool okay = false;
	// results of the last accepted detection (CTA-uniform copies)
	int r_tfire = -1, r_sympos = 0, r_scpos = 0, r_imax = 0, r_shift = 0, r_poserr = 0, r_unique = 0, r_mode = 0;
	float r_tmax = 0.f, r_frac = 0.f, r_cfo = 0.f;
	unsigned long long r_md = 0;

	for (int di = 0; di < nd; ++di) {
		const Detection D = det[(size_t)f * det_cap + di];
		if (D.t_max < 0) continue;
		// phase_max = arg(P[t_max - 80]),  P[t] = sum_{k<640} a[t-k-5119] conj(a[t-k-4479])   (decode.cc:86,91,101)
		{
			const int tau = D.t_max - kMatchDel;
			float pr = 0.f, pi = 0.f;
			for (int k = tid; k < kHalf; k += kAcqThreads) {
				const cfx c = cmulc(load_iq(a, tau - k - kOffOld, iq_len), load_iq(a, tau - k - kOffCur, iq_len));
				pr += c.x; pi += c.y;
			}
#pragma unroll
			for (int d = 16; d; d >>= 1) { pr += __shfl_xor_sync(FULL, pr, d); pi += __shfl_xor_sync(FULL, pi, d); }
			if (lane == 0) { s.redf[wid][0] = pr; s.redf[wid][1] = pi; }
			__syncthreads();
			if (tid == 0) {
				float x = 0.f, y = 0.f;
				for (int w2 = 0; w2 < kAcqWarps; ++w2) { x += s.redf[w2][0]; y += s.redf[w2][1]; }
				s.bc_f[0] = tau >= 0 ? atan2f(y, x) : 0.f;
			}
			__syncthreads();
		}
		const float phase_max = s.bc_f[0];
		const float frac_cfo = phase_max / (float)kHalf;
		int symbol_pos = kSearchPos - D.index_max;
		const int win0 = D.t_fall - (kBufferLen - 1);
		for (int i = tid; i < kHalf; i += kAcqThreads) {
			float sn, cs;
			sincosf(frac_cfo * (float)i, &sn, &cs);
			buf0[i] = cmul(load_iq(a, win0 + symbol_pos + kHalf + i, iq_len), make_float2(cs, sn));
		}
		__syncthreads();
		// three half-length transforms; each result lands in one of the two buffers (depends on the pass count of the
		// length), the next input is written into the other one
		cfx *const sp = fft_fwd<kHalf>(buf0, buf1, ac.tw640, tid, kAcqThreads); // spectrum
		cfx *const in2 = sp == buf0 ? buf1 : buf0;
		for (int i = tid; i < kHalf; i += kAcqThreads) in2[i] = demod_or_erase(sp[i], sp[(i + kHalf - 1) % kHalf]);
		__syncthreads();
		cfx *const x2 = fft_fwd<kHalf>(in2, sp, ac.tw640, tid, kAcqThreads);
		cfx *const in3 = x2 == in2 ? sp : in2;
		// * kern, then backward transform as conj(fwd(conj(.)))
		for (int i = tid; i < kHalf; i += kAcqThreads) {
			const cfx v = cmul(x2[i], ac.kern640[i]);
			in3[i] = make_float2(v.x, -v.y);
		}
		__syncthreads();
		const cfx *const xc = fft_fwd<kHalf>(in3, x2, ac.tw640, tid, kAcqThreads); // conj(result)
		// peak / runner-up (decode.cc:127-139): first index of the maximum, second largest of the multiset
		float pk = -1.f;
		int pki = 1 << 30;
		for (int i = tid; i < kHalf; i += kAcqThreads) {
			const float p = cnorm(xc[i]);
			if (p > pk) { pk = p; pki = i; }
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) {
			const float op = __shfl_xor_sync(FULL, pk, d);
			const int oi = __shfl_xor_sync(FULL, pki, d);
			if (op > pk || (op == pk && oi < pki)) { pk = op; pki = oi; }
		}
		if (lane == 0) { s.redf[wid][0] = pk; s.redi[wid][0] = pki; }
		__syncthreads();
		if (tid == 0) {
			float bp = -1.f;
			int bi = 1 << 30;
			for (int w2 = 0; w2 < kAcqWarps; ++w2)
				if (s.redf[w2][0] > bp || (s.redf[w2][0] == bp && s.redi[w2][0] < bi)) { bp = s.redf[w2][0]; bi = s.redi[w2][0]; }
			s.bc_f[1] = bp;
			s.bc_i[0] = bi;
		}
		__syncthreads();
		const float peak = s.bc_f[1];
		const int shift = peak > 0.f ? s.bc_i[0] : 0;
		float nx = 0.f;
		for (int i = tid; i < kHalf; i += kAcqThreads)
			if (i != shift) nx = fmaxf(nx, cnorm(xc[i]));
#pragma unroll
		for (int d = 16; d; d >>= 1) nx = fmaxf(nx, __shfl_xor_sync(FULL, nx, d));
		if (lane == 0) s.redf[wid][1] = nx;
		__syncthreads();
		if (tid == 0) {
			float v = 0.f;
			for (int w2 = 0; w2 < kAcqWarps; ++w2) v = fmaxf(v, s.redf[w2][1]);
			s.bc_f[2] = v;
		}
		__syncthreads();
		const float next = s.bc_f[2];
		const float pkv = fmaxf(peak, 0.f);
		if (pkv <= next * 4.f) { __syncthreads(); continue; }
		const cfx top = make_float2(xc[shift].x, -xc[shift].y); // undo the conj of the backward transform
		const int pos_err = (int)rintf(__fdiv_rn(atan2f(top.y, top.x) * (float)kHalf, 6.28318530717958647692f));
		if (abs(pos_err) > kGuardLen / 2) { __syncthreads(); continue; }
		symbol_pos -= pos_err;
		float cfo_rad = (float)shift * (6.28318530717958647692f / (float)kHalf) - frac_cfo;
		if (cfo_rad >= 3.14159265358979323846f) cfo_rad -= 6.28318530717958647692f;
		__syncthreads();

		// ---------------- accepted detection: header symbol (decode.cc:398-447)
		++accepted;
		r_tfire = D.t_fall; r_sympos = symbol_pos; r_scpos = win0 + symbol_pos; r_imax = D.index_max; r_shift = shift;
		r_poserr = pos_err; r_tmax = D.timing_max; r_frac = frac_cfo; r_cfo = cfo_rad;
		const double turns = -(double)cfo_rad / 6.283185307179586476925286766559;
		for (int i = tid; i < kSymLen; i += kAcqThreads)
			buf0[i] = cmul(load_iq(a, r_scpos + kPitch + i, iq_len), phasor_turns(turns * (double)i));
		__syncthreads();
		const cfx *const X = fft_fwd<kSymLen>(buf0, buf1, ac.tw1280, tid, kAcqThreads);
		if (tid < 256) s.soft[tid] = 0;
		__syncthreads();
		if (tid < NB) {
			const int kc = (tid - 127 + kSymLen) % kSymLen, kp = (tid - 128 + kSymLen) % kSymLen;
			cfx cur = X[kc], prev = X[kp];
			if (ac.mls1[tid]) { cur.x = -cur.x; cur.y = -cur.y; }
			if (tid > 0 && ac.mls1[tid - 1]) { prev.x = -prev.x; prev.y = -prev.y; }
			float v = rintf(127.f * demod_or_erase(cur, prev).x);
			v = fminf(fmaxf(v, -128.f), 127.f);
			s.soft[tid] = (int)v;
			if (soft_out) soft_out[(size_t)f * 256 + tid] = (int8_t)v;
		}
		__syncthreads();
		const bool unique = osd_decode(s, ac.bch_rows, tid);
		r_unique = unique;
		bool good = false;
		if (!unique) status = ST_OSD_FAIL;
		else {
			if (tid == 0) {
				unsigned long long md = 0;
				for (int i = 0; i < 55; ++i) md |= (unsigned long long)s.hbits[i] << i;
				unsigned cs = 0;
				for (int i = 0; i < 16; ++i) cs |= (unsigned)s.hbits[55 + i] << i;
				const unsigned long long v = md << 9;
				unsigned crc = 0;
				for (int b = 0; b < 64; ++b) {
					const unsigned bit = (unsigned)(v >> b) & 1u;
					crc = (crc >> 1) ^ (((crc ^ bit) & 1u) ? 0xA8F4u : 0u);
				}
				s.best = md;
				s.bc_i[1] = crc == cs;
			}
			__syncthreads();
			r_md = s.best;
			const bool crc_ok = s.bc_i[1];
			r_mode = (int)(r_md & 255ull);
			if (!crc_ok) status = ST_HDR_CRC;
			else if (r_mode < 6 || r_mode > 13) status = ST_BAD_MODE;
			else if ((r_md >> 8) == 0ull || (long long)(r_md >> 8) >= kCallSignLimit) status = ST_BAD_CALL;
			else good = true;
		}
		okay = good;
		__syncthreads();
		if (skip_left == 0) break;
		--skip_left;
		okay = false; // more detections are to be consumed; the decision is taken on the last one only
	}
	if (tid == 0) {
		st.status = okay ? ST_OK : status;
		st.detections = accepted;
		st.t_fire = r_tfire; st.symbol_pos = r_sympos; st.sc_pos = r_scpos; st.index_max = r_imax; st.shift = r_shift; st.pos_err = r_poserr;
		st.timing_max = r_tmax; st.frac_cfo = r_frac; st.cfo_rad = r_cfo;
		st.osd_unique = r_unique; st.mode = r_mode;
		st.md_lo = (uint32_t)r_md; st.md_hi = (uint32_t)(r_md >> 32);
		st.best_lane = -1; st.flips = -1;
		for (int k = 0; k < 8; ++k) st.metrics[k] = 0.f;
		st.osd_visited = 0;
		st.ts_sweeps = 0; st.det_overflow = det_overflow;
	}
}

__global__ void k_compact(const FrameState *st, int n_frames, int *cw_list, int *n_cw)
{
	// single CTA: ordered compaction of the header-ok frames, code table 0 (modes 6..9) first, then — from the next
	// multiple of four on — code table 1 (modes 10..13): the list decoder takes four codewords of ONE table per warp
	__shared__ int base;
	__shared__ int wsum[32];
	for (int tb = 0; tb < 2; ++tb) {
		__syncthreads();
		if (threadIdx.x == 0) base = tb ? (n_cw[0] + 3) & ~3 : 0;
		__syncthreads();
		const int first = base;
		for (int f0 = 0; f0 < n_frames; f0 += blockDim.x) {
			const int f = f0 + threadIdx.x;
			const int ok = f < n_frames && st[f].status == ST_OK && mode_info(st[f].mode).table == tb;
			const unsigned bal = __ballot_sync(FULL, ok);
			const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
			if (lane == 0) wsum[wid] = __popc(bal);
			__syncthreads();
			int off = base;
			for (int w = 0; w < wid; ++w) off += wsum[w];
			if (ok) cw_list[off + __popc(bal & ((1u << lane) - 1u))] = f;
			__syncthreads();
			if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w]; base += t; }
			__syncthreads();
		}
		if (threadIdx.x == 0) n_cw[tb] = base - first;
	}
}


} // namespace

template <int S>
static cudaError_t launch_acquire_t(const cfx *iq, int64_t iq_stride, int iq_len, const Detection *det, const int32_t *det_count, int det_cap, int skip,
	int n_frames, FrameState *st, int8_t *soft_out, const AcquireConsts &ac, cudaStream_t s)
{
	static DeviceOnce once;
	const size_t smem = kAcqBufOff + 2 * (size_t)Geo<S>::kSymLen * sizeof(cfx);
	if (cudaError_t e = set_dynamic_smem_once(once, k_acquire<S>, (int)smem)) return e;
	k_acquire<S><<<n_frames, kAcqThreads, smem, s>>>(iq, iq_stride, iq_len, det, det_count, det_cap, skip, st, soft_out, ac);
	return cudaGetLastError();
}

cudaError_t launch_acquire(int rate, const cfx *iq, int64_t iq_stride, int iq_len, const Detection *det, const int32_t *det_count, int det_cap, int skip,
	int n_frames, FrameState *st, int8_t *soft_out, const AcquireConsts &ac, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
	cudaError_t e = cudaSuccess;
#define OFDMRX_CALL(R) e = launch_acquire_t<R>(iq, iq_stride, iq_len, det, det_count, det_cap, skip, n_frames, st, soft_out, ac, s)
	OFDMRX_FOR_RATE(rate, OFDMRX_CALL)
#undef OFDMRX_CALL
	return e;
}

cudaError_t launch_compact(const FrameState *st, int n_frames, int *cw_list, int *n_cw, cudaStream_t s)
{
	k_compact<<<1, 1024, 0, s>>>(st, n_frames, cw_list, n_cw);
	return cudaGetLastError();
}

} // namespace ofdmrx
