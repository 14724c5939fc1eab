// modem_b200/csrc/stimulus.cuh — device-side stimulus: the transmitter of /root/reference/encode.cc:27-318 and the
// re-specified `disorders` impairment chain (README.md:49), batched — SURVEY.md §8(f1).
//
// B200 shape of the work (not the reference's: it runs one FFT-5120 pair per symbol through one heap object):
//   * payload -> code bits: one CTA per frame, bit-packed XOR butterflies in shared memory (k_tx_code);
//   * one CTA per OFDM symbol (k_tx_symbol).  The 4x-oversampled PAPR clip (encode.cc:80-100) never materialises the
//     4N-point grid: phase p of the oversampled signal is an N-point inverse transform of the spectrum twisted by
//     e^{+j 2 pi i p / 4N}, clipping is point-wise, and only the N occupied-band bins of the forward 4N transform are needed
//     (= sum over p of twisted N-point transforms).  Eight N-point Stockham transforms in 2N+1024 complex values of shared memory
//     instead of two 4N-point ones in 8N — which is what lets 44.1/48 kHz symbols (4N = 28224 / 30720) fit at all;
//   * the three frame-constant symbols (pilot, Schmidl-Cox, metadata) are built once per call, not per frame;
//   * guard cross-fade, multipath, CFO, SFO, AWGN and quantisation are a gather over the symbol store (k_tx_stream_a/_b).
//
// Every routine below is written against (tid, nthr) so that tests/stimulus_host.cu can run the very same code on the host
// with one "thread" (the CPU suite then checks it against the oracle's transmitter without a GPU).
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include <cmath>

namespace ofdmrx {

#define OFDMRX_TX_SYNC() OFDMRX_CTA_SYNC()

constexpr int kTxMaxCarriers = 512;                        // widest occupied band (mode 10), encode.cc:232
constexpr int kTxSymPilot = 0, kTxSymSc = 1, kTxSymMeta = 2; // frame-constant symbols
constexpr int kTxCodeWords = kCodeLen / 32;

struct TxSpec { int first, step, count; };                  // occupied carriers: signed index first + step * c, c < count

// impairments as the oracle's `Impair` (oracle/ref_modem.hh) defines them — OUR re-specification of aicodix/disorders
struct TxImpair {
	int multipath;   // fixed 4-tap sparse complex FIR
	float cfo_hz;    // complex mixer
	float sfo_ppm;   // 33-tap Kaiser-windowed-sinc resampling by (1 + ppm 1e-6)
	int awgn;        // complex Gaussian, total variance 10^(awgn_db/10)
	float awgn_db;
	unsigned long long seed;    // Philox key; window i of a call draws from the counter plane (window0 + i, sample)
	unsigned long long window0; // index of the kernel's window 0 inside the call (chunked calls), set by the launcher
};

struct TxParams {
	int rate, sym_len, guard_len, pitch;
	int cols, mod_bits, rows, cons_bits, table;
	int code_off;
	int frames_per_window, n_sym;        // symbols per window: 1 + frames_per_window * (3 + rows) + 1
	long long len;                       // sample frames per window before resampling: 2 rate + n_sym * pitch
	TxSpec spec[3];
	const cfx *common_fdom;              // [3][kTxMaxCarriers] occupied-carrier values of pilot / S-C / metadata
	const cfx *tw_sym, *tw_4n;           // exp(-2 pi j k / N), exp(-2 pi j k / 4N)
	const float *ramp;                   // raised-cosine cross-fade weights, guard_len values (encode.cc:111-112)
	cfx *tdom_common;                    // [3][N] time-domain bodies of the frame-constant symbols
	cfx *tdom;                           // [frames][rows][N] data symbol bodies
};

// ---- exact-rounding helpers: the prefix products and the cross-fade follow the reference's operation order without FMA
OFDMRX_HD float tx_mul(float a, float b)
{
#ifdef __CUDA_ARCH__
	return __fmul_rn(a, b);
#else
	return a * b;
#endif
}
OFDMRX_HD float tx_add(float a, float b)
{
#ifdef __CUDA_ARCH__
	return __fadd_rn(a, b);
#else
	return a + b;
#endif
}
OFDMRX_HD cfx tx_cmul(cfx a, cfx b)
{
	return make_float2(tx_add(tx_mul(a.x, b.x), -tx_mul(a.y, b.y)), tx_add(tx_mul(a.x, b.y), tx_mul(a.y, b.x)));
}
OFDMRX_HD cfx tx_conj(cfx a) { return make_float2(a.x, -a.y); }
OFDMRX_HD int tx_bin(int carrier, int n) { int b = carrier % n; return b < 0 ? b + n : b; }

// ---- PSK mapping (psk.hh:84-87 QPSK, :132-139 8PSK); bit t of `bits` is code bit t, a set bit is the value -1
OFDMRX_HD cfx tx_psk_map(int mod_bits, uint32_t bits)
{
	const float b0 = (bits & 1u) ? -1.f : 1.f, b1 = (bits & 2u) ? -1.f : 1.f;
	if (mod_bits == 2) return make_float2(0.70710678118654752440f * b0, 0.70710678118654752440f * b1);
	const float b2 = (bits & 4u) ? -1.f : 1.f;
	float re = 0.92387953251128675613f, im = 0.38268343236508977173f;
	if (b0 < 0.f) { const float t = re; re = im; im = t; }
	return make_float2(re * b1, im * b2);
}
OFDMRX_HD uint32_t tx_code_bits(const uint32_t *code, int idx, int count)
{
	const int w = idx >> 5, s = idx & 31;
	uint32_t v = code[w] >> s;
	if (s + count > 32) v |= code[w + 1] << (32 - s);
	return v & ((1u << count) - 1u);
}
// carrier c of data row `row`: pilot value times the constellation points of rows 0..row, in the reference's order
// (encode.cc:305-307: fdom *= mod_map(...) once per row — differential encoding along time)
OFDMRX_HD cfx tx_data_carrier(const uint32_t *code, int cols, int mod_bits, int row, int c, cfx pilot)
{
	cfx v = pilot;
	for (int r = 0; r <= row; ++r) v = tx_cmul(v, tx_psk_map(mod_bits, tx_code_bits(code, mod_bits * (cols * r + c), mod_bits)));
	return v;
}

// ---- payload -> transmitted code bits (encode.cc:293-303,180-186 + the scrambling of encode.cc:417-419) --------------------
// payload: 5380 plain bytes; scr: the Xorshift32 byte stream packed little-endian; lut: reflected CRC-32 table of
// 0xD419CC15; frozen / msg_off: the code table of the mode (bit set = frozen; msg_off[w] = free positions before word w).
// mesg: 1378 words of scratch, cw: 2048 words of scratch (shared memory on the device); out: 2048 words, bit i = code bit i
// (bits >= cons_bits are not transmitted: shorten() == truncation, tests/test_oracle_kat.py).
OFDMRX_HD void tx_polar_transform(uint32_t *cw, int tid, int nthr)
{
	for (int w = tid; w < kTxCodeWords; w += nthr) {
		uint32_t v = cw[w];
		v ^= (v >> 1) & 0x55555555u;
		v ^= (v >> 2) & 0x33333333u;
		v ^= (v >> 4) & 0x0f0f0f0fu;
		v ^= (v >> 8) & 0x00ff00ffu;
		v ^= (v >> 16) & 0x0000ffffu;
		cw[w] = v;
	}
	OFDMRX_TX_SYNC();
	for (int h = 1; h < kTxCodeWords; h <<= 1) {
		for (int q = tid; q < kTxCodeWords / 2; q += nthr) {
			const int i = ((q & ~(h - 1)) << 1) | (q & (h - 1)); // word index with bit h clear
			cw[i] ^= cw[i + h];
		}
		OFDMRX_TX_SYNC();
	}
}
constexpr int kTxMesgWords = 1378 + 2; // 44096 bits + one word of slack for the straddling read
OFDMRX_HD void tx_code_core(const uint8_t *payload, const uint32_t *scr, const uint32_t *lut, const uint32_t *frozen,
	const uint32_t *msg_off, uint32_t *mesg, uint32_t *cw, uint32_t *out, int tid, int nthr)
{
	for (int w = tid; w < kTxMesgWords; w += nthr) {
		uint32_t v = 0;
		if (w < kDataBytes / 4) {
			const uint8_t *p = payload + 4 * w;
			v = ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24) ^ scr[w];
		}
		mesg[w] = v;
	}
	OFDMRX_TX_SYNC();
	if (tid == 0) { // CRC-32 of the scrambled bytes, appended LSB first (encode.cc:295-299); bits above stay 0 (= +1)
		uint32_t crc = 0;
		for (int w = 0; w < kDataBytes / 4; ++w) {
			uint32_t v = mesg[w];
			for (int b = 0; b < 4; ++b, v >>= 8) crc = lut[(crc ^ v) & 255u] ^ (crc >> 8);
		}
		mesg[kDataBits / 32] = crc;
	}
	OFDMRX_TX_SYNC();
	for (int w = tid; w < kTxCodeWords; w += nthr) { // message bits into the free positions, frozen positions 0
		uint32_t free_mask = ~frozen[w], v = 0;
		const uint32_t o = msg_off[w];
		uint64_t src = ((uint64_t)mesg[(o >> 5) + 1] << 32 | mesg[o >> 5]) >> (o & 31);
		while (free_mask) {
			const uint32_t low = free_mask & (0u - free_mask);
			if (src & 1u) v |= low;
			src >>= 1;
			free_mask ^= low;
		}
		cw[w] = v;
	}
	OFDMRX_TX_SYNC();
	tx_polar_transform(cw, tid, nthr); // two-pass systematic encoding (PolarSysEnc, encode.cc:48,302)
	for (int w = tid; w < kTxCodeWords; w += nthr) cw[w] &= ~frozen[w];
	OFDMRX_TX_SYNC();
	tx_polar_transform(cw, tid, nthr);
	for (int w = tid; w < kTxCodeWords; w += nthr) out[w] = cw[w];
}

// ---- one OFDM symbol body (encode.cc:80-109): occupied carriers -> N time-domain samples -----------------------------------
// car: `count` occupied-carrier values (shared); acc: same size scratch; b0, b1: N complex values each (shared).
template <int N>
OFDMRX_HD void tx_symbol_core(const cfx *car, cfx *acc, TxSpec sp, bool papr, cfx *b0, cfx *b1, const cfx *tw, const cfx *tw4,
	cfx *tdom_out, int tid, int nthr)
{
	const float sc4 = sqrtf(float(4 * N)), sc8 = sqrtf(float(8 * N));
	if (papr) {
		for (int c = tid; c < sp.count; c += nthr) acc[c] = make_float2(0.f, 0.f);
		for (int p = 0; p < 4; ++p) {
			for (int i = tid; i < N; i += nthr) b0[i] = make_float2(0.f, 0.f);
			OFDMRX_TX_SYNC();
			for (int c = tid; c < sp.count; c += nthr) { // conj(X e^{+j 2 pi i p / 4N}): inverse transform through the forward one
				const int i = sp.first + sp.step * c;
				b0[tx_bin(i, N)] = cmul(tx_conj(car[c]), tw4[tx_bin(i * p, 4 * N)]);
			}
			OFDMRX_TX_SYNC();
			cfx *r = fft_fwd<N>(b0, b1, tw, tid, nthr);
			for (int n = tid; n < N; n += nthr) { // sample 4n+p of the oversampled symbol, clipped to the unit square (encode.cc:87-93)
				cfx x = make_float2(r[n].x / sc4, -r[n].y / sc4);
				const float amp = fmaxf(fabsf(x.x), fabsf(x.y));
				if (amp > 1.f) { x.x /= amp; x.y /= amp; }
				r[n] = x;
			}
			OFDMRX_TX_SYNC();
			const cfx *f = fft_fwd<N>(r, r == b0 ? b1 : b0, tw, tid, nthr);
			for (int c = tid; c < sp.count; c += nthr) {
				const int i = sp.first + sp.step * c;
				acc[c] = cadd(acc[c], cmul(f[tx_bin(i, N)], tw4[tx_bin(i * p, 4 * N)]));
			}
			OFDMRX_TX_SYNC();
		}
	}
	for (int i = tid; i < N; i += nthr) b0[i] = make_float2(0.f, 0.f);
	OFDMRX_TX_SYNC();
	for (int c = tid; c < sp.count; c += nthr) {
		const cfx v = papr ? make_float2(acc[c].x / sc4, acc[c].y / sc4) : car[c];
		b0[tx_bin(sp.first + sp.step * c, N)] = tx_conj(v);
	}
	OFDMRX_TX_SYNC();
	const cfx *r = fft_fwd<N>(b0, b1, tw, tid, nthr);
	for (int n = tid; n < N; n += nthr) tdom_out[n] = make_float2(r[n].x / sc8, -r[n].y / sc8);
}

// ---- the analytic stream of one window before impairments (encode.cc:101-131,288-313,423,441) -------------------------------
// symbol s of a window: 0 leading pilot; 1 + k (3 + rows) + {0 S-C, 1 metadata, 2 pilot, 3 + j data row j} for frame k; last = zero
OFDMRX_HD cfx tx_symbol_sample(const TxParams &p, long long window, int s, int o)
{
	if (s == p.n_sym - 1) return make_float2(0.f, 0.f);
	if (s == 0) return p.tdom_common[(size_t)kTxSymPilot * p.sym_len + o];
	const int q = s - 1, per = 3 + p.rows, k = q / per, r = q - k * per;
	if (r == 0) return p.tdom_common[(size_t)kTxSymSc * p.sym_len + o];
	if (r == 1) return p.tdom_common[(size_t)kTxSymMeta * p.sym_len + o];
	if (r == 2) return p.tdom_common[(size_t)kTxSymPilot * p.sym_len + o];
	return p.tdom[(((size_t)window * p.frames_per_window + k) * p.rows + (r - 3)) * p.sym_len + o];
}
OFDMRX_HD cfx tx_stream_sample(const TxParams &p, long long window, long long n)
{
	const cfx zero = make_float2(0.f, 0.f);
	if (n < p.rate) return zero; // one second of silence either side; n < 0 included
	const long long m = n - p.rate;
	const int s = (int)(m / p.pitch);
	if (s >= p.n_sym) return zero;
	const int o = (int)(m - (long long)s * p.pitch);
	if (o >= p.guard_len) return tx_symbol_sample(p, window, s, o - p.guard_len);
	const cfx a = s > 0 ? tx_symbol_sample(p, window, s - 1, o) : zero;
	const cfx b = tx_symbol_sample(p, window, s, o + p.sym_len - p.guard_len);
	const float x = p.ramp[o], y = 1.f - x; // DSP::lerp(a, b, x) = (1 - x) a + x b
	return make_float2(tx_add(tx_mul(y, a.x), tx_mul(x, b.x)), tx_add(tx_mul(y, a.y), tx_mul(x, b.y)));
}

// ---- impairments -----------------------------------------------------------------------------------------------------------
// multipath + CFO at output index n of the un-resampled stream
OFDMRX_HD cfx tx_channel_sample(const TxParams &p, const TxImpair &im, long long window, long long n)
{
	cfx v;
	if (im.multipath) {
		const int dly[4] = {0, 3, 7, 10};
		const cfx tap[4] = {make_float2(1.f, 0.f), make_float2(0.35f, -0.25f), make_float2(-0.2f, 0.15f), make_float2(0.1f, 0.1f)};
		v = make_float2(0.f, 0.f);
		for (int t = 0; t < 4; ++t)
			if (n >= dly[t]) {
				const cfx m = tx_cmul(tap[t], tx_stream_sample(p, window, n - dly[t]));
				v = make_float2(tx_add(v.x, m.x), tx_add(v.y, m.y));
			}
	} else {
		v = tx_stream_sample(p, window, n);
	}
	if (im.cfo_hz != 0.f) {
		const double ph = 2.0 * M_PI * fmod((double)im.cfo_hz * (double)n / (double)p.rate, 1.0);
		v = tx_cmul(v, make_float2((float)cos(ph), (float)sin(ph)));
	}
	return v;
}
// Kaiser window helper of the oracle's resampler: sum_{n<35} ((x/2)^n / n!)^2 in fp32.  The terms fall monotonically once
// n > x/2 and a term below half an ulp of the running sum cannot change it, so stopping there returns the very same float as
// the full 35 steps at a third of the work.
OFDMRX_HD float tx_bessel_i0(float x)
{
	float sum = 1, val = 1;
	for (int n = 1; n < 35; ++n) {
		val *= x / float(2 * n);
		const float t = val * val;
		sum += t;
		if (float(2 * n) > x && t < sum * 1.4e-8f) break; // 2^-26 = 1.49e-8: below half an ulp of sum
	}
	return sum;
}
constexpr int kTxSfoHalf = 16;
OFDMRX_HD long long tx_resampled_len(long long len, float sfo_ppm)
{
	if (sfo_ppm == 0.f) return len;
	return (long long)((double)len / (1.0 + (double)sfo_ppm * 1e-6));
}
// output sample n of the resampled stream; src: `len` samples of the window after multipath + CFO.
// weight(k) = sinc(k - frac) * kaiser((k - frac) / 17); sin(pi (k - frac)) = -(-1)^k sin(pi frac): one sine per sample.
OFDMRX_HD cfx tx_resample(const cfx *src, long long len, float sfo_ppm, long long n)
{
	const double ratio = 1.0 + (double)sfo_ppm * 1e-6, pos = (double)n * ratio;
	const long long base = (long long)floor(pos);
	const double frac = pos - (double)base;
	const double i0b = (double)tx_bessel_i0((float)(M_PI * 2.5));
	const double s0 = sin(M_PI * frac);
	double are = 0, aim = 0;
	for (int k = -kTxSfoHalf; k <= kTxSfoHalf; ++k) {
		const long long idx = base + k;
		if (idx < 0 || idx >= len) continue;
		const double x = (double)k - frac;
		const double sinc = fabs(x) < 1e-12 ? 1.0 : ((k & 1) ? s0 : -s0) / (M_PI * x);
		const double t = x / (double)(kTxSfoHalf + 1);
		const double win = fabs(t) >= 1.0 ? 0.0 : (double)tx_bessel_i0((float)(M_PI * 2.5 * sqrt(1.0 - t * t))) / i0b;
		are += sinc * win * (double)src[idx].x;
		aim += sinc * win * (double)src[idx].y;
	}
	return make_float2((float)are, (float)aim);
}

// Philox-4x32-10 keyed by the call's seed, counter = (sample index, window index of the call): the device's noise stream;
// streams of different (seed, window) pairs never coincide, whatever the chunking.  (The oracle draws from
// mt19937_64; the two streams are different realisations of the same distribution — tests compare statistics, and decode
// parity is checked on whatever windows this generator produced.)
OFDMRX_HD uint32_t tx_mulhi(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
OFDMRX_HD void tx_philox(unsigned long long seed, unsigned long long ctr, unsigned long long plane, uint32_t out[4])
{
	uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)plane, c3 = (uint32_t)(plane >> 32), k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
	for (int r = 0; r < 10; ++r) {
		const uint32_t h0 = tx_mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = tx_mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
		const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
		c0 = n0; c1 = l1; c2 = n2; c3 = l0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
OFDMRX_HD cfx tx_noise(const TxImpair &im, long long window, long long n)
{
	uint32_t r[4];
	tx_philox(im.seed, (unsigned long long)n, im.window0 + (unsigned long long)window, r);
	const double u1 = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) / 18446744073709551616.0;
	const float u2 = ((float)(r[2] >> 8) + 0.5f) / 16777216.f;
	const float sigma = sqrtf(powf(10.f, im.awgn_db / 10.f) / 2.f);
	const float rad = (float)sqrt(-2.0 * log(u1)) * sigma, a = 6.28318530717958647692f * u2;
	return make_float2(rad * cosf(a), rad * sinf(a));
}

// DSP::WritePCM<float> at 16 bits (recalled: clamp to [-1, 1], nearbyint(32767 x)); oracle/ref_dsp.hh quantize16
OFDMRX_HD int16_t tx_quantize16(float x)
{
	x = fminf(fmaxf(x, -1.f), 1.f);
	return (int16_t)(int)rintf(32767.f * x);
}
// out: window base pointer; format as OFDMRX_FMT_* (0 int16 real part, 1 int16 I/Q, 2 float2)
OFDMRX_HD void tx_store(void *out, int format, long long n, cfx v)
{
	if (format == 0) ((int16_t *)out)[n] = tx_quantize16(v.x);
	else if (format == 1) { ((int16_t *)out)[2 * n] = tx_quantize16(v.x); ((int16_t *)out)[2 * n + 1] = tx_quantize16(v.y); }
	else ((cfx *)out)[n] = v;
}

} // namespace ofdmrx
