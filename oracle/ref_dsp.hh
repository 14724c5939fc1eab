// oracle/ref_dsp.hh — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Restatement of the DSP primitives the reference pulls from the ABSENT, UNPINNED
// header library aicodix/dsp (`-I../dsp`, /root/reference/Makefile:2): complex, FFT,
// sliding sums, Hilbert, DC blocker, phasor, triggers, WAV.  Each block cites the
// reference call site it serves.  Semantics marked (recalled) come from the public
// aicodix/dsp sources as remembered, not from files on this box: PARITY UNPINNED.
//
// Plain scalar fp32, compile with -O2 -ffp-contract=off (no -ffast-math) so results are
// reproducible across hosts.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <vector>
#include <string>
#include <algorithm>

namespace ref {

// ---------------------------------------------------------------- complex (complex.hh, recalled)
struct cf {
	float re, im;
	cf() : re(0), im(0) {}
	cf(float r, float i = 0) : re(r), im(i) {}
};
static inline cf operator+(cf a, cf b) { return cf(a.re + b.re, a.im + b.im); }
static inline cf operator-(cf a, cf b) { return cf(a.re - b.re, a.im - b.im); }
static inline cf operator*(cf a, cf b) { return cf(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cf operator*(float a, cf b) { return cf(a * b.re, a * b.im); }
static inline cf operator*(cf b, float a) { return cf(a * b.re, a * b.im); }
static inline cf operator/(cf a, float b) { return cf(a.re / b, a.im / b); }
static inline cf conj(cf a) { return cf(a.re, -a.im); }
static inline float norm(cf a) { return a.re * a.re + a.im * a.im; }
static inline float arg(cf a) { return std::atan2(a.im, a.re); }
static inline float cabs(cf a) { return std::sqrt(norm(a)); }
// textbook division as in DSP::Complex: (a * conj(b)) / norm(b)
static inline cf operator/(cf a, cf b) { return (a * conj(b)) / norm(b); }
static inline cf polar(float r, float th) { return cf(r * std::cos(th), r * std::sin(th)); }

static const float kPi = 3.14159265358979323846f;
static const float kTwoPi = 6.28318530717958647692f;

// ---------------------------------------------------------------- FFT (fft.hh, recalled: unnormalised mixed radix)
// X[k] = sum_n x[n] * exp(sign * j*2*pi*n*k/N); sign=-1 forward (decode.cc:43,191), +1 backward (decode.cc:44).
// Generic Stockham autosort for N = 2^a 3^b 5^c 7^d; fp32 arithmetic, twiddles rounded from double.
class FFT {
	int n_, sign_;
	std::vector<int> radices_;
	std::vector<std::vector<cf>> tw_;   // per stage: twiddle w^(q*k) for q in [1,r), k in [0,m)
	std::vector<std::vector<cf>> wr_;   // per stage: r x r DFT matrix
public:
	FFT(int n, int sign) : n_(n), sign_(sign)
	{
		int rem = n;
		// stage order: large radices last so the radix-5 pass is a single stage
		const int cand[] = {4, 2, 3, 5, 7};
		for (int c : cand)
			while (rem % c == 0) { radices_.push_back(c); rem /= c; }
		if (rem != 1) { std::fprintf(stderr, "FFT: unsupported length %d\n", n); std::abort(); }
		int m = 1; // product of radices handled so far
		for (size_t s = 0; s < radices_.size(); ++s) {
			int r = radices_[s];
			std::vector<cf> tw((size_t)(r - 1) * m);
			for (int q = 1; q < r; ++q)
				for (int k = 0; k < m; ++k) {
					double a = sign * 2.0 * M_PI * (double)q * (double)k / (double)(m * r);
					tw[(size_t)(q - 1) * m + k] = cf((float)std::cos(a), (float)std::sin(a));
				}
			tw_.push_back(tw);
			std::vector<cf> wr((size_t)r * r);
			for (int p = 0; p < r; ++p)
				for (int q = 0; q < r; ++q) {
					double a = sign * 2.0 * M_PI * (double)((p * q) % r) / (double)r;
					wr[(size_t)p * r + q] = cf((float)std::cos(a), (float)std::sin(a));
				}
			wr_.push_back(wr);
			m *= r;
		}
	}
	int size() const { return n_; }
	// out-of-place, in may equal out
	void operator()(cf *out, const cf *in) const
	{
		// Decimation in time, Stockham: stage s combines r sub-transforms of length m into length m*r.
		// Data layout before stage s: x[(j*m + k) ... ] with stride structure handled via two buffers.
		std::vector<cf> a(n_), b(n_);
		// start: sub-transforms of length 1: element i of sub-transform j is in[j] with j = i (natural)
		// We use the classic formulation: y[k + m*(p + r*j')] ... implemented with index arithmetic below.
		const cf *src = in;
		cf *bufs[2] = {a.data(), b.data()};
		int cur = 0;
		int m = 1;
		int l = n_; // number of sub-transforms remaining = n / m
		for (size_t s = 0; s < radices_.size(); ++s) {
			int r = radices_[s];
			l /= r;
			cf *dst = bufs[cur];
			const std::vector<cf> &tw = tw_[s];
			const std::vector<cf> &wr = wr_[s];
			// src holds l*r sub-transforms of length m: sub-transform index g in [0, l*r), element k: src[g*m... ]?
			// Stockham DIT indexing: input  x[k + m*(j + l*q)]  (q in [0,r)), output y[k + m*(q' ) ...]
			for (int j = 0; j < l; ++j)
				for (int k = 0; k < m; ++k) {
					cf t[7];
					t[0] = src[k + m * (j + l * 0)];
					for (int q = 1; q < r; ++q)
						t[q] = src[k + m * (j + l * q)] * tw[(size_t)(q - 1) * m + k];
					for (int p = 0; p < r; ++p) {
						cf acc = t[0];
						for (int q = 1; q < r; ++q)
							acc = acc + t[q] * wr[(size_t)p * r + q];
						dst[k + m * (p + r * j)] = acc;
					}
				}
			src = dst;
			cur ^= 1;
			m *= r;
		}
		if (radices_.empty()) { out[0] = in[0]; return; }
		std::memcpy(out, src, sizeof(cf) * n_);
	}
};

// ---------------------------------------------------------------- sliding window aggregate (sma.hh SMA4 / swa.hh, recalled)
// Exact re-summation over a heap-shaped tree of the last NUM inputs (decode.cc:45-47).
template <typename T>
class SlidingSum {
	int num_, leaf_;
	std::vector<T> tree_;
public:
	explicit SlidingSum(int num) : num_(num), leaf_(num), tree_(2 * (size_t)num, T()) {}
	T operator()(T in)
	{
		tree_[leaf_] = in;
		for (int child = leaf_, parent = leaf_ / 2; parent; child = parent, parent /= 2)
			tree_[parent] = tree_[child & ~1] + tree_[child | 1];
		if (++leaf_ >= 2 * num_) leaf_ = num_;
		return tree_[1];
	}
};

// ---------------------------------------------------------------- Kaiser window + Hilbert (window.hh / hilbert.hh, recalled)
static inline float kaiser_i0(float x)
{
	float sum = 1, val = 1;
	for (int n = 1; n < 35; ++n) {
		val *= x / float(2 * n);
		sum += val * val;
	}
	return sum;
}
static inline float kaiser(float a, int n, int N)
{
	float t = float(2 * n) / float(N - 1) - 1.f;
	return kaiser_i0(kPi * a * std::sqrt(1.f - t * t)) / kaiser_i0(kPi * a);
}
// Hilbert<cmplx,TAPS> (decode.cc:172,193,299): re = centre tap * reco (delay (TAPS-1)/2),
// im = sum over odd offsets k of imco[(k-1)/2] * (x[n-c+k ... ]) — antisymmetric FIR 2/(pi k) * Kaiser(a=2).
struct HilbertCoeffs {
	int taps;
	float reco;
	std::vector<float> imco; // (taps-1)/4 coefficients for odd offsets 1,3,5,...
	explicit HilbertCoeffs(int taps_, float a = 2.f) : taps(taps_)
	{
		reco = kaiser(a, (taps - 1) / 2, taps);
		for (int i = 0; i < (taps - 1) / 4; ++i)
			imco.push_back(kaiser(a, (2 * i + 1) + (taps - 1) / 2, taps) * 2.f / (float(2 * i + 1) * kPi));
	}
};
class Hilbert {
	HilbertCoeffs c_;
	std::vector<float> hist_; // hist_[0] oldest … hist_[taps-1] newest
public:
	explicit Hilbert(int taps) : c_(taps), hist_(taps, 0.f) {}
	cf operator()(float in)
	{
		int T = c_.taps, mid = (T - 1) / 2;
		float re = c_.reco * hist_[mid];
		float im = c_.imco[0] * (hist_[mid - 1] - hist_[mid + 1]);
		for (int i = 1; i < (T - 1) / 4; ++i)
			im += c_.imco[i] * (hist_[mid - (2 * i + 1)] - hist_[mid + (2 * i + 1)]);
		for (int i = 0; i < T - 1; ++i) hist_[i] = hist_[i + 1];
		hist_[T - 1] = in;
		return cf(re, im);
	}
};

// ---------------------------------------------------------------- DC blocker (blockdc.hh, recalled; decode.cc:192,299,386)
class BlockDC {
	float x1_ = 0, y1_ = 0, a_ = 0, b_ = 0.5f;
public:
	void samples(int s) { a_ = float(s - 1) / float(s); b_ = (1.f + a_) / 2.f; }
	float a() const { return a_; }
	float b() const { return b_; }
	float operator()(float x0)
	{
		float y0 = b_ * (x0 - x1_) + a_ * y1_;
		x1_ = x0;
		y1_ = y0;
		return y0;
	}
};

// ---------------------------------------------------------------- Phasor (phasor.hh, recalled; decode.cc:112,387,403)
class Phasor {
	cf prev_{1, 0}, delta_{1, 0};
public:
	void omega(float v) { delta_ = cf(std::cos(v), std::sin(v)); }
	cf operator()()
	{
		cf tmp = prev_;
		prev_ = prev_ * delta_;
		prev_ = prev_ / cabs(prev_);
		return tmp;
	}
};

// ---------------------------------------------------------------- triggers / delay (trigger.hh, delay.hh, recalled)
struct SchmittTrigger {
	float low, high;
	bool state = false;
	SchmittTrigger(float l, float h) : low(l), high(h) {}
	bool operator()(float in)
	{
		if (state) { if (in < low) state = false; }
		else { if (in > high) state = true; }
		return state;
	}
};
struct FallingEdge {
	bool prev = false;
	bool operator()(bool in) { bool t = prev; prev = in; return t && !in; }
};
template <typename T>
class Delay {
	std::vector<T> buf_;
	int pos_ = 0;
public:
	explicit Delay(int n) : buf_(n, T()) {}
	T operator()(T in)
	{
		T t = buf_[pos_];
		buf_[pos_] = in;
		if (++pos_ >= (int)buf_.size()) pos_ = 0;
		return t;
	}
};

// ---------------------------------------------------------------- WAV / PCM (wav.hh, pcm.hh, recalled; decode.cc:576, encode.cc:422)
// int PCM <-> float: v / (2^(bits-1)-1), 8-bit has offset 128.  Canonical 44-byte RIFF header on write;
// the reader additionally skips unknown chunks before "data".
struct WavData {
	int rate = 0, bits = 0, channels = 0;
	std::vector<float> samples; // interleaved
	size_t frames() const { return channels ? samples.size() / channels : 0; }
};
static inline int pcm_factor(int bits) { return (1 << (bits - 1)) - 1; }
static inline bool wav_parse(const uint8_t *d, size_t len, WavData &w)
{
	auto rd32 = [&](size_t o) { return (uint32_t)d[o] | (uint32_t)d[o + 1] << 8 | (uint32_t)d[o + 2] << 16 | (uint32_t)d[o + 3] << 24; };
	auto rd16 = [&](size_t o) { return (uint32_t)d[o] | (uint32_t)d[o + 1] << 8; };
	if (len < 44 || std::memcmp(d, "RIFF", 4) || std::memcmp(d + 8, "WAVE", 4)) return false;
	size_t o = 12;
	bool have_fmt = false;
	while (o + 8 <= len) {
		uint32_t sz = rd32(o + 4);
		if (!std::memcmp(d + o, "fmt ", 4)) {
			if (rd16(o + 8) != 1) return false; // PCM only
			w.channels = rd16(o + 10);
			w.rate = rd32(o + 12);
			w.bits = rd16(o + 22);
			have_fmt = true;
		} else if (!std::memcmp(d + o, "data", 4)) {
			if (!have_fmt) return false;
			size_t avail = len - (o + 8);
			size_t n = std::min<size_t>(sz, avail);
			if (sz == 0xffffffffu || sz == 0) n = avail; // streamed WAV of unknown length
			int bytes = w.bits / 8;
			if (bytes < 1 || bytes > 4 || w.channels < 1) return false;
			size_t cnt = n / bytes / w.channels * w.channels;
			w.samples.resize(cnt);
			float fac = (float)pcm_factor(w.bits);
			int off = bytes == 1 ? 128 : 0;
			for (size_t i = 0; i < cnt; ++i) {
				const uint8_t *p = d + o + 8 + i * bytes;
				int32_t v = 0;
				for (int b = 0; b < bytes; ++b) v |= (int32_t)p[b] << (8 * b);
				if (bytes > 1 && bytes < 4 && (v & (1 << (8 * bytes - 1)))) v |= ~((1 << (8 * bytes)) - 1);
				w.samples[i] = float(v - off) / fac;
			}
			return true;
		}
		o += 8 + sz + (sz & 1);
	}
	return false;
}
static inline std::vector<uint8_t> wav_serialize(int rate, int bits, int channels, const std::vector<float> &interleaved)
{
	int bytes = bits / 8;
	size_t n = interleaved.size();
	std::vector<uint8_t> o(44 + n * bytes);
	auto wr32 = [&](size_t p, uint32_t v) { for (int b = 0; b < 4; ++b) o[p + b] = (v >> (8 * b)) & 255; };
	auto wr16 = [&](size_t p, uint32_t v) { for (int b = 0; b < 2; ++b) o[p + b] = (v >> (8 * b)) & 255; };
	std::memcpy(&o[0], "RIFF", 4); wr32(4, (uint32_t)(36 + n * bytes)); std::memcpy(&o[8], "WAVEfmt ", 8);
	wr32(16, 16); wr16(20, 1); wr16(22, channels); wr32(24, rate); wr32(28, rate * channels * bytes);
	wr16(32, channels * bytes); wr16(34, bits); std::memcpy(&o[36], "data", 4); wr32(40, (uint32_t)(n * bytes));
	float fac = (float)pcm_factor(bits);
	int off = bytes == 1 ? 128 : 0;
	for (size_t i = 0; i < n; ++i) {
		float x = std::min(std::max(interleaved[i], -1.f), 1.f);
		int32_t v = (int32_t)std::nearbyint(fac * x) + off;
		for (int b = 0; b < bytes; ++b) o[44 + i * bytes + b] = (v >> (8 * b)) & 255;
	}
	return o;
}
// float -> int16 exactly as the 16-bit WAV writer quantises (used for batch stimulus without file I/O)
static inline int16_t quantize16(float x)
{
	x = std::min(std::max(x, -1.f), 1.f);
	return (int16_t)std::nearbyint(32767.f * x);
}

} // namespace ref
