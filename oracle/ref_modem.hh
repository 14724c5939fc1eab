// oracle/ref_modem.hh — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// CPU restatement of the reference transmitter (/root/reference/encode.cc) and receiver
// (/root/reference/decode.cc) for all modes 6..13 at 8/16/44.1/48 kHz, plus a re-specified impairment
// chain standing in for the absent aicodix/disorders tools (README.md:49).  The receiver exposes STAGE
// TAPS so each CUDA kernel can be compared at its own output.  PARITY UNPINNED for the third-party arithmetic:
// a stock reference cannot be compiled here (aicodix/dsp + aicodix/code absent).  What pins this file is
//   (0) the reference's OWN decode.cc / encode.cc compiled against shim/ (the primitives of ref_dsp.hh / ref_code.hh behind
//       the absent headers' API): same WAV bytes, same payloads, same stderr as this restatement (tests/test_reference_tu.py),
//   (1) polar_tables.hh regenerated bit-exactly by ref_freezer.hh (the one real golden vector),
//   (2) SURVEY.md Appendix B known answers (tests/test_oracle_kat.py),
//   (3) encode -> decode loop-back returning the exact payload.
#pragma once
#include "ref_dsp.hh"
#include "ref_code.hh"
#include "ref_freezer.hh"
#include <random>
#include <memory>

namespace ref {

static const int kDataBits = 43040, kDataBytes = 5380, kCrcBits = 43072;
static const long long kCallSignLimit = 129961739795077LL; // 37^9 (decode.cc:439, encode.cc:358)

// ---------------------------------------------------------------- mode table (decode.cc:302-374 == encode.cc:197-270)
struct ModeParams {
	int cons_cols, mod_bits, code_order, cons_bits, mesg_bits, table; // table 0: frozen_64800_43072, 1: frozen_64512_43072
	int cons_cnt() const { return cons_bits / mod_bits; }
	int cons_rows() const { return cons_cnt() / cons_cols; }
};
static inline bool mode_params(int mode, ModeParams &m)
{
	static const int tab[8][3] = {{432, 3, 0}, {400, 3, 0}, {400, 2, 0}, {360, 2, 0}, {512, 3, 1}, {384, 3, 1}, {384, 2, 1}, {256, 2, 1}};
	if (mode < 6 || mode > 13) return false;
	const int *t = tab[mode - 6];
	m.cons_cols = t[0]; m.mod_bits = t[1]; m.table = t[2];
	m.code_order = 16;
	m.cons_bits = t[2] ? 64512 : 64800;
	m.mesg_bits = t[2] ? 44096 : 43808;
	return true;
}
static inline const uint32_t *frozen_table(int table)
{
	static const std::vector<uint32_t> t0 = make_frozen_table(16, 64800, kCrcBits);
	static const std::vector<uint32_t> t1 = make_frozen_table(16, 64512, kCrcBits);
	return table ? t1.data() : t0.data();
}
static inline int band_width(int mode) // encode.cc:363-387
{
	switch (mode) { case 6: return 2700; case 7: case 8: return 2500; case 9: return 2250; case 10: return 3200; case 11: case 12: return 2400; case 13: return 1600; }
	return -1;
}

// QPSK (reference psk.hh:49-88)
struct PSK4 {
	static constexpr float rcp_sqrt_2 = 0.70710678118654752440f;
	static constexpr float DIST = 2 * rcp_sqrt_2;
	static void hard(float *b, cf c) { b[0] = c.re < 0.f ? -1.f : 1.f; b[1] = c.im < 0.f ? -1.f : 1.f; }
	static cf map(const float *b) { return rcp_sqrt_2 * cf(b[0], b[1]); }
	static void soft(float *b, cf c, float precision) { b[0] = c.re * (DIST * precision); b[1] = c.im * (DIST * precision); }
};
static inline cf mod_map(int mod_bits, const float *b) { return mod_bits == 3 ? PSK8::map(b) : PSK4::map(b); }
static inline void mod_hard(int mod_bits, float *b, cf c) { if (mod_bits == 3) PSK8::hard(b, c); else PSK4::hard(b, c); }
static inline void mod_soft(int mod_bits, float *b, cf c, float p) { if (mod_bits == 3) PSK8::soft(b, c, p); else PSK4::soft(b, c, p); }

// ================================================================ transmitter (encode.cc:27-318)
class Transmitter {
	int rate_, symbol_len_, guard_len_;
	FFT bwd_, fwd4_, bwd4_;
	ModeParams mp_;
	int mode_ = 0, code_off_ = 0, mls0_off_ = 0, mls1_off_ = 0;
	std::vector<cf> fdom_, temp_, tdom_, guard_, fdom4_, tdom4_;
	std::vector<cf> *out_ = nullptr;
	BCH255_71 bch_;
	int bin(int c) const { return (c + symbol_len_) % symbol_len_; }
	int bin4(int c) const { return (c + 4 * symbol_len_) % (4 * symbol_len_); }
	static int nrz(bool b) { return 1 - 2 * (int)b; }

	void improve_papr() // encode.cc:80-100
	{
		int n4 = 4 * symbol_len_;
		for (int i = 0; i < n4; ++i) fdom4_[i] = cf();
		for (int i = -symbol_len_ / 2; i < symbol_len_ / 2; ++i) fdom4_[bin4(i)] = fdom_[bin(i)];
		bwd4_(tdom4_.data(), fdom4_.data());
		float sc = std::sqrt(float(n4));
		for (int i = 0; i < n4; ++i) tdom4_[i] = tdom4_[i] / sc;
		for (int i = 0; i < n4; ++i) {
			float amp = std::max(std::abs(tdom4_[i].re), std::abs(tdom4_[i].im));
			if (amp > 1.f) tdom4_[i] = tdom4_[i] / amp;
		}
		fwd4_(fdom4_.data(), tdom4_.data());
		for (int i = -symbol_len_ / 2; i < symbol_len_ / 2; ++i)
			temp_[bin(i)] = norm(temp_[bin(i)]) != 0.f ? fdom4_[bin4(i)] / sc : cf();
	}
	void symbol(bool papr = true) // encode.cc:101-131
	{
		temp_ = fdom_;
		if (papr) improve_papr();
		bwd_(tdom_.data(), temp_.data());
		float sc = std::sqrt(float(8 * symbol_len_));
		for (int i = 0; i < symbol_len_; ++i) tdom_[i] = tdom_[i] / sc;
		for (int i = 0; i < guard_len_; ++i) {
			float x = float(i) / float(guard_len_ - 1);
			x = 0.5f * (1.f - std::cos(kPi * x));
			cf a = guard_[i], b = tdom_[i + symbol_len_ - guard_len_];
			guard_[i] = (1.f - x) * a + x * b; // DSP::lerp
		}
		out_->insert(out_->end(), guard_.begin(), guard_.end());
		out_->insert(out_->end(), tdom_.begin(), tdom_.end());
		for (int i = 0; i < guard_len_; ++i) guard_[i] = tdom_[i];
	}
	void pilot_block() // encode.cc:132-141
	{
		MLS seq2(0b100101010001);
		float fac = std::sqrt(float(symbol_len_) / float(mp_.cons_cols));
		std::fill(fdom_.begin(), fdom_.end(), cf());
		for (int i = code_off_; i < code_off_ + mp_.cons_cols; ++i) fdom_[bin(i)] = cf(fac * nrz(seq2()));
		symbol();
	}
	void schmidl_cox() // encode.cc:142-154
	{
		MLS seq0(0b10001001);
		float fac = std::sqrt(float(2 * symbol_len_) / 127.f);
		std::fill(fdom_.begin(), fdom_.end(), cf());
		fdom_[bin(mls0_off_ - 2)] = cf(fac);
		for (int i = 0; i < 127; ++i) fdom_[bin(2 * i + mls0_off_)] = cf((float)nrz(seq0()));
		for (int i = 0; i < 127; ++i) fdom_[bin(2 * i + mls0_off_)] = fdom_[bin(2 * i + mls0_off_)] * fdom_[bin(2 * (i - 1) + mls0_off_)];
		symbol(false);
	}
	void meta_data(uint64_t md) // encode.cc:155-179
	{
		uint8_t bits[71], par[184];
		for (int i = 0; i < 55; ++i) bits[i] = (md >> i) & 1;
		CRC<uint16_t> crc0(0xA8F4);
		uint16_t cs = crc0.u64(md << 9);
		for (int i = 0; i < 16; ++i) bits[55 + i] = (cs >> i) & 1;
		bch_.encode_bits(bits, par);
		MLS seq1(0b100101011);
		float fac = std::sqrt(float(symbol_len_) / 255.f);
		std::fill(fdom_.begin(), fdom_.end(), cf());
		fdom_[bin(mls1_off_ - 1)] = cf(fac);
		for (int i = 0; i < 71; ++i) fdom_[bin(i + mls1_off_)] = cf((float)nrz(bits[i]));
		for (int i = 71; i < 255; ++i) fdom_[bin(i + mls1_off_)] = cf((float)nrz(par[i - 71]));
		for (int i = 0; i < 255; ++i) fdom_[bin(i + mls1_off_)] = fdom_[bin(i + mls1_off_)] * fdom_[bin(i - 1 + mls1_off_)];
		for (int i = 0; i < 255; ++i) fdom_[bin(i + mls1_off_)] = fdom_[bin(i + mls1_off_)] * (float)nrz(seq1());
		symbol();
	}
public:
	explicit Transmitter(int rate) : rate_(rate), symbol_len_(1280 * rate / 8000), guard_len_(symbol_len_ / 8),
		bwd_(symbol_len_, 1), fwd4_(4 * symbol_len_, -1), bwd4_(4 * symbol_len_, 1),
		fdom_(symbol_len_), temp_(symbol_len_), tdom_(symbol_len_), guard_(guard_len_),
		fdom4_(4 * symbol_len_), tdom4_(4 * symbol_len_) {}

	// scrambled payload -> code bits (encode.cc:293-303): data LE bits, CRC-32 of the scrambled bytes LSB first,
	// zero padding, systematic polar encode, shorten (== keep code[0..cons_bits))
	static void payload_to_code(const uint8_t *scrambled, const ModeParams &mp, std::vector<uint8_t> &code)
	{
		std::vector<uint8_t> mesg(mp.mesg_bits, 0);
		for (int i = 0; i < kDataBits; ++i) mesg[i] = get_le_bit(scrambled, i);
		CRC<uint32_t> crc1(0xD419CC15);
		for (int i = 0; i < kDataBytes; ++i) crc1.byte(scrambled[i]);
		for (int i = 0; i < 32; ++i) mesg[kDataBits + i] = (crc1() >> i) & 1;
		FrozenSet fs{frozen_table(mp.table)};
		code.assign(1 << mp.code_order, 0);
		polar_sys_encode(code.data(), mesg.data(), fs, mp.code_order);
		// shorten(): encode.cc:180-186 — literal restatement, asserted equal to truncation in the KAT tests
		int n = 1 << mp.code_order;
		std::vector<uint8_t> sh;
		sh.reserve(mp.cons_bits);
		for (int i = 0, k = 0; i < n; ++i)
			if (fs.frozen(i) || k++ < kCrcBits) sh.push_back(code[i]);
		sh.resize(mp.cons_bits);
		code = sh;
	}
	// payloads: count x 5380 plain bytes (scrambling applied here as encode.cc:417-419 does in main()).
	// Returns the analytic stream incl. 1 s of silence either side (encode.cc:423,441).
	bool encode(std::vector<cf> &out, const uint8_t *payloads, int count, int freq_off, long long call_sign, int mode)
	{
		if (!mode_params(mode, mp_)) return false;
		mode_ = mode;
		out_ = &out;
		out.clear();
		out.insert(out.end(), rate_, cf());
		std::fill(guard_.begin(), guard_.end(), cf());
		int offset = (freq_off * symbol_len_) / rate_;
		code_off_ = offset - mp_.cons_cols / 2;
		mls0_off_ = offset - 127 + 1;
		mls1_off_ = offset - 255 / 2;
		pilot_block();
		std::vector<uint8_t> code;
		for (int k = 0; k < count; ++k) {
			uint8_t scr[kDataBytes];
			Xorshift32 prng;
			for (int i = 0; i < kDataBytes; ++i) scr[i] = payloads[(size_t)k * kDataBytes + i] ^ (uint8_t)prng();
			schmidl_cox();
			meta_data(((uint64_t)call_sign << 8) | (uint64_t)mode);
			pilot_block();
			payload_to_code(scr, mp_, code);
			for (int j = 0; j < mp_.cons_rows(); ++j) {
				for (int i = 0; i < mp_.cons_cols; ++i) {
					float b[3];
					for (int t = 0; t < mp_.mod_bits; ++t) b[t] = (float)nrz(code[mp_.mod_bits * (mp_.cons_cols * j + i) + t]);
					fdom_[bin(i + code_off_)] = fdom_[bin(i + code_off_)] * mod_map(mp_.mod_bits, b);
				}
				symbol();
			}
		}
		std::fill(fdom_.begin(), fdom_.end(), cf());
		symbol();
		out.insert(out.end(), rate_, cf());
		return true;
	}
	static bool check_args(int rate, int chan, int freq_off, int mode, long long cs) // encode.cc:353-397
	{
		if (mode < 6 || mode > 13) return false;
		if (cs <= 0 || cs >= kCallSignLimit) return false;
		int bw = band_width(mode);
		if ((chan == 1 && freq_off < bw / 2) || freq_off < bw / 2 - rate / 2 || freq_off > rate / 2 - bw / 2) return false;
		if (freq_off % 50) return false;
		return true;
	}
};

// ================================================================ impairments (re-specified aicodix/disorders, README.md:49)
// The tools are absent; these are OUR definitions (documented in DESIGN.md), applied to the analytic stream:
//   multipath: fixed sparse complex FIR (taps below); cfo: mixer exp(j 2 pi f t / rate);
//   sfo: band-limited resampling by (1 + ppm 1e-6) (Kaiser-windowed sinc, 33 taps);
//   awgn: complex Gaussian, total variance 10^(level_db/10) (re and im each half), seeded per frame.
struct Impair {
	bool multipath = false;
	float cfo_hz = 0.f;
	float sfo_ppm = 0.f;
	bool awgn = false;
	float awgn_db = -30.f;
	uint64_t seed = 1;
};
static inline void apply_impairments(std::vector<cf> &s, int rate, const Impair &im)
{
	if (im.multipath) {
		static const int dly[4] = {0, 3, 7, 10};
		static const cf tap[4] = {cf(1.f, 0.f), cf(0.35f, -0.25f), cf(-0.2f, 0.15f), cf(0.1f, 0.1f)};
		std::vector<cf> o(s.size());
		for (size_t n = 0; n < s.size(); ++n) {
			cf acc;
			for (int t = 0; t < 4; ++t) if (n >= (size_t)dly[t]) acc = acc + tap[t] * s[n - dly[t]];
			o[n] = acc;
		}
		s.swap(o);
	}
	if (im.cfo_hz != 0.f) {
		for (size_t n = 0; n < s.size(); ++n) {
			double ph = 2.0 * M_PI * std::fmod((double)im.cfo_hz * (double)n / (double)rate, 1.0);
			s[n] = s[n] * cf((float)std::cos(ph), (float)std::sin(ph));
		}
	}
	if (im.sfo_ppm != 0.f) {
		const int H = 16;
		double ratio = 1.0 + (double)im.sfo_ppm * 1e-6;
		size_t nout = (size_t)((double)s.size() / ratio);
		std::vector<cf> o(nout);
		for (size_t n = 0; n < nout; ++n) {
			double pos = (double)n * ratio;
			long long base = (long long)std::floor(pos);
			double frac = pos - (double)base;
			double are = 0, aim = 0;
			for (int k = -H; k <= H; ++k) {
				long long idx = base + k;
				if (idx < 0 || idx >= (long long)s.size()) continue;
				double x = (double)k - frac;
				double sinc = std::abs(x) < 1e-12 ? 1.0 : std::sin(M_PI * x) / (M_PI * x);
				double t = x / (double)(H + 1);
				double win = std::abs(t) >= 1.0 ? 0.0 : (double)kaiser_i0((float)(M_PI * 2.5 * std::sqrt(1.0 - t * t))) / (double)kaiser_i0((float)(M_PI * 2.5));
				are += sinc * win * s[idx].re;
				aim += sinc * win * s[idx].im;
			}
			o[n] = cf((float)are, (float)aim);
		}
		s.swap(o);
	}
	if (im.awgn) {
		std::mt19937_64 gen(im.seed * 0x9E3779B97F4A7C15ull + 12345);
		// Box-Muller on explicit uniform draws so the stream is identical across libstdc++ versions
		double sigma = std::sqrt(std::pow(10.0, (double)im.awgn_db / 10.0) / 2.0);
		for (size_t n = 0; n < s.size(); ++n) {
			double u1 = ((double)(gen() >> 11) + 0.5) / 9007199254740992.0;
			double u2 = ((double)(gen() >> 11) + 0.5) / 9007199254740992.0;
			double r = std::sqrt(-2.0 * std::log(u1)) * sigma;
			s[n] = s[n] + cf((float)(r * std::cos(2.0 * M_PI * u2)), (float)(r * std::sin(2.0 * M_PI * u2)));
		}
	}
}
// analytic stream -> interleaved float PCM as the WAV writer would store it (1 ch: real part; 2 ch: re, im)
static inline void to_pcm16(const std::vector<cf> &s, int channels, std::vector<int16_t> &pcm)
{
	pcm.resize(s.size() * channels);
	for (size_t n = 0; n < s.size(); ++n) {
		pcm[n * channels] = quantize16(s[n].re);
		if (channels == 2) pcm[n * channels + 1] = quantize16(s[n].im);
	}
}

// ================================================================ receiver (decode.cc:37-153 SchmidlCox, :161-557 Decoder)
enum Status { // per-frame outcome (what the reference only prints on stderr)
	ST_OK = 0, ST_NO_SYNC = 1, ST_OSD_FAIL = 2, ST_HDR_CRC = 3, ST_BAD_MODE = 4, ST_BAD_CALL = 5, ST_PAYLOAD_CRC = 6
};
struct Taps {
	int status = ST_NO_SYNC;
	int detections = 0;
	int t_fire = -1;          // stream index (0-based) of the sample whose arrival fired the accepted detection
	int symbol_pos = 0;       // correlator.symbol_pos after the fine correction (decode.cc:146,398)
	int sc_pos = 0;           // absolute stream index of the S-C symbol body = t_fire - (buffer_len-1) + symbol_pos
	int index_max = 0, shift = 0, pos_err = 0;
	float timing_max = 0, frac_cfo = 0, cfo_rad = 0;
	int8_t soft[255] = {0};
	uint8_t hdr[32] = {0};
	int osd_unique = 0;
	uint64_t md = 0;
	int mode = 0;
	char call_sign[10] = {0};
	std::vector<cf> cons_raw; // after demod_or_erase (decode.cc:475)
	std::vector<cf> cons;     // after Theil–Sen derotation (decode.cc:494)
	std::vector<float> slope, yint, precision;
	std::vector<float> llr;   // code[] after lengthen() (decode.cc:529), 65536 values
	float metrics[8] = {0};
	int best_lane = -1, flips = -1;
	long long forks = 0, osd_visited = 0;
};
struct RxOptions {
	int list_size = 8;         // SIMD width of the reference build: 8 (AVX2) or 4
	int r0_max = 1 << 16;      // largest rate-0 node handled at node level (1 = leaf by leaf)
	bool osd_literal = false;  // true: enumerate all 1 031 347 candidates like the reference
};

class Receiver {
	int rate_, symbol_len_, guard_len_, filter_len_, buffer_len_, search_pos_, half_;
	int match_len_, match_del_;
	FFT fwd_, fwdh_, bwdh_;
	std::vector<cf> kern_;
	int8_t genmat_[255 * 71];
	// input stream
	const float *pcm_ = nullptr;
	size_t n_frames_ = 0, pos_ = 0;
	int channels_ = 1;
	bool good_ = true;
	BlockDC blockdc_;
	std::unique_ptr<Hilbert> hilbert_;
	std::vector<cf> ring_; // bi-partite history: window = last buffer_len_ samples, contiguous
	int ring_pos_ = 0;
	const cf *buf_ = nullptr;
	// correlator state (decode.cc:45-56)
	SlidingSum<cf> cor_;
	SlidingSum<float> pwr_, match_;
	Delay<float> delay_;
	SchmittTrigger threshold_;
	FallingEdge falling_;
	float timing_max_ = 0, phase_max_ = 0;
	int index_max_ = 0;
	int bin(int c) const { return (c + symbol_len_) % symbol_len_; }
	int binh(int c) const { return (c + half_) % half_; }
	static cf demod_or_erase(cf curr, cf prev) // decode.cc:62-70, 227-235
	{
		if (!(norm(prev) > 0.f)) return cf();
		cf cons = curr / prev;
		if (!(norm(cons) <= 4.f)) return cf();
		return cons;
	}
	const cf *next_sample() // decode.cc:294-301 (+ EOF: zeros; the reference leaves this undefined)
	{
		cf tmp;
		if (pos_ < n_frames_) {
			tmp.re = pcm_[pos_ * channels_];
			if (channels_ == 2) tmp.im = pcm_[pos_ * channels_ + 1];
			++pos_;
		} else {
			good_ = false;
		}
		if (channels_ == 1) tmp = (*hilbert_)(blockdc_(tmp.re));
		// bi-partite buffer: write twice so that the last buffer_len_ samples are contiguous
		ring_[ring_pos_] = tmp;
		ring_[ring_pos_ + buffer_len_] = tmp;
		ring_pos_ = (ring_pos_ + 1) % buffer_len_;
		buf_ = &ring_[ring_pos_];
		++stream_count_;
		return buf_;
	}
	long long stream_count_ = 0; // samples pushed so far
public:
	Taps taps;
	std::string header_log; // what decode.cc:400-446 prints on stderr, one block per consumed detection (SKIP walks several)
	explicit Receiver(int rate) : rate_(rate), symbol_len_(1280 * rate / 8000), guard_len_(symbol_len_ / 8),
		filter_len_((((21 * rate) / 8000) & ~3) | 1), buffer_len_(6 * (symbol_len_ + guard_len_)),
		search_pos_(buffer_len_ - 4 * (symbol_len_ + guard_len_)), half_(symbol_len_ / 2),
		match_len_(guard_len_ | 1), match_del_((match_len_ - 1) / 2),
		fwd_(symbol_len_, -1), fwdh_(half_, -1), bwdh_(half_, 1), kern_(half_),
		cor_(half_), pwr_(2 * half_), match_(match_len_), delay_(match_del_),
		threshold_(float(0.17 * match_len_), float(0.19 * match_len_))
	{
		// MLS0 template (decode.cc:236-244) and its conjugate spectrum / N (decode.cc:76-83)
		std::vector<cf> seq(half_);
		MLS seq0(0b10001001);
		const int mls0_len = 127, mls0_off = -mls0_len + 1;
		for (int i = 0; i < mls0_len; ++i) seq[(i + mls0_off / 2 + half_) % half_] = cf((float)(1 - 2 * (int)seq0()));
		fwdh_(kern_.data(), seq.data());
		for (int i = 0; i < half_; ++i) kern_[i] = conj(kern_[i]) / float(half_);
		BCH255_71 bch;
		bch.matrix(genmat_);
	}
	int symbol_len() const { return symbol_len_; }
	int guard_len() const { return guard_len_; }
	int buffer_len() const { return buffer_len_; }

	// one step of SchmidlCox::operator() (decode.cc:84-152); returns true on an accepted detection
	bool correlate(const cf *samples)
	{
		cf P = cor_(samples[search_pos_ + half_] * conj(samples[search_pos_ + 2 * half_]));
		float R = 0.5f * pwr_(norm(samples[search_pos_ + 2 * half_]));
		float min_R = float(0.0001 * half_);
		R = std::max(R, min_R);
		float timing = match_(norm(P) / (R * R));
		float phase = delay_(arg(P));
		bool collect = threshold_(timing);
		bool process = falling_(collect);
		if (!collect && !process) return false;
		if (timing_max_ < timing) {
			timing_max_ = timing;
			phase_max_ = phase;
			index_max_ = match_del_;
		} else if (index_max_ < half_ + guard_len_ + match_del_) {
			++index_max_;
		}
		if (!process) return false;
		++taps.detections;
		float frac_cfo = phase_max_ / float(half_);
		Phasor osc;
		osc.omega(frac_cfo);
		int symbol_pos = search_pos_ - index_max_;
		taps.index_max = index_max_;
		taps.timing_max = timing_max_;
		index_max_ = 0;
		timing_max_ = 0;
		std::vector<cf> t0(half_), t1(half_), t2(half_);
		for (int i = 0; i < half_; ++i) t1[i] = samples[i + symbol_pos + half_] * osc();
		fwdh_(t0.data(), t1.data());
		for (int i = 0; i < half_; ++i) t1[i] = demod_or_erase(t0[i], t0[binh(i - 1)]);
		fwdh_(t0.data(), t1.data());
		for (int i = 0; i < half_; ++i) t0[i] = t0[i] * kern_[i];
		bwdh_(t2.data(), t0.data());
		int shift = 0;
		float peak = 0, next = 0;
		for (int i = 0; i < half_; ++i) {
			float power = norm(t2[i]);
			if (power > peak) { next = peak; peak = power; shift = i; }
			else if (power > next) next = power;
		}
		if (peak <= next * 4.f) return false;
		int pos_err = (int)std::nearbyint(arg(t2[shift]) * float(half_) / kTwoPi);
		if (std::abs(pos_err) > guard_len_ / 2) return false;
		symbol_pos -= pos_err;
		float cfo_rad = float(shift) * (kTwoPi / float(half_)) - frac_cfo;
		if (cfo_rad >= kPi) cfo_rad -= kTwoPi;
		taps.symbol_pos = symbol_pos;
		taps.shift = shift;
		taps.pos_err = pos_err;
		taps.frac_cfo = frac_cfo;
		taps.cfo_rad = cfo_rad;
		return true;
	}

	// Stage taps of the streaming front end, one entry per stream step t = 0..n_frames (the reference takes one step past the
	// end of the file before it notices, decode.cc:391-396): iq[t] = the sample next_sample() pushed at step t (decode.cc:294-301:
	// BlockDC + Hilbert for one channel, I/Q as read for two), timing[t] = the matched-filter output of decode.cc:90 at that step.
	// The trigger logic is left out: these are the inputs it sees.
	void front_taps(const float *pcm, size_t n_frames, int channels, std::vector<cf> &iq, std::vector<float> &timing)
	{
		pcm_ = pcm; n_frames_ = n_frames; channels_ = channels; pos_ = 0; good_ = true; stream_count_ = 0;
		blockdc_ = BlockDC();
		blockdc_.samples(2 * (symbol_len_ + guard_len_));
		hilbert_.reset(new Hilbert(filter_len_));
		ring_.assign(2 * (size_t)buffer_len_, cf());
		ring_pos_ = 0;
		cor_ = SlidingSum<cf>(half_); pwr_ = SlidingSum<float>(2 * half_); match_ = SlidingSum<float>(match_len_);
		iq.clear(); timing.clear();
		while (good_) {
			const cf *samples = next_sample();
			iq.push_back(samples[buffer_len_ - 1]);
			cf P = cor_(samples[search_pos_ + half_] * conj(samples[search_pos_ + 2 * half_]));
			float R = 0.5f * pwr_(norm(samples[search_pos_ + 2 * half_]));
			R = std::max(R, float(0.0001 * half_));
			timing.push_back(match_(norm(P) / (R * R)));
		}
	}

	// Decoder::Decoder (decode.cc:375-556).  pcm: interleaved float frames as ReadWAV delivers them.
	// Returns Status; out (5380 B) is written only on ST_OK and is NOT yet de-scrambled (decode.cc:613-615 does that in main).
	int run(uint8_t *out, const float *pcm, size_t n_frames, int channels, int skip_count, const RxOptions &opt = RxOptions())
	{
		taps = Taps();
		header_log.clear();
		pcm_ = pcm; n_frames_ = n_frames; channels_ = channels; pos_ = 0; good_ = true; stream_count_ = 0;
		blockdc_ = BlockDC();
		blockdc_.samples(2 * (symbol_len_ + guard_len_));
		hilbert_.reset(new Hilbert(filter_len_));
		ring_.assign(2 * (size_t)buffer_len_, cf());
		ring_pos_ = 0;
		cor_ = SlidingSum<cf>(half_); pwr_ = SlidingSum<float>(2 * half_); match_ = SlidingSum<float>(match_len_);
		delay_ = Delay<float>(match_del_);
		threshold_.state = false; falling_.prev = false;
		timing_max_ = phase_max_ = 0; index_max_ = 0;

		Phasor osc;
		const cf *buf = nullptr;
		ModeParams mp{};
		std::vector<cf> fdom(symbol_len_), tdom(symbol_len_);
		OSD255_71 osd;
		bool okay;
		int symbol_pos = 0;
		float cfo_rad = 0;
		do {
			okay = false;
			do {
				if (!good_) { taps.status = taps.detections ? taps.status : ST_NO_SYNC; return taps.status; }
				buf = next_sample();
			} while (!correlate(buf));
			symbol_pos = taps.symbol_pos;
			cfo_rad = taps.cfo_rad;
			taps.t_fire = (int)(stream_count_ - 1);
			taps.sc_pos = taps.t_fire - (buffer_len_ - 1) + symbol_pos;
			{
				char line[96]; // operator<<(float) prints like %.6g
				std::snprintf(line, sizeof(line), "symbol pos: %d\ncoarse cfo: %.6g Hz \n", symbol_pos, (double)(cfo_rad * (rate_ / kTwoPi)));
				header_log += line;
			}
			osc.omega(-cfo_rad);
			for (int i = 0; i < symbol_len_; ++i) tdom[i] = buf[i + symbol_pos + (symbol_len_ + guard_len_)] * osc();
			fwd_(fdom.data(), tdom.data());
			MLS seq1(0b100101011);
			const int mls1_len = 255, mls1_off = -mls1_len / 2;
			for (int i = 0; i < mls1_len; ++i) fdom[bin(i + mls1_off)] = fdom[bin(i + mls1_off)] * (float)(1 - 2 * (int)seq1());
			for (int i = 0; i < mls1_len; ++i) {
				float v = std::nearbyint(127.f * demod_or_erase(fdom[bin(i + mls1_off)], fdom[bin(i - 1 + mls1_off)]).re);
				taps.soft[i] = (int8_t)std::min(std::max(v, -128.f), 127.f);
			}
			bool unique = opt.osd_literal ? osd.decode_full(taps.hdr, taps.soft, genmat_) : osd.decode_pruned(taps.hdr, taps.soft, genmat_);
			taps.osd_unique = unique;
			taps.osd_visited = osd.visited;
			if (!unique) { taps.status = ST_OSD_FAIL; header_log += "OSD error.\n"; continue; }
			uint64_t md = 0;
			for (int i = 0; i < 55; ++i) md |= (uint64_t)get_be_bit(taps.hdr, i) << i;
			uint16_t cs = 0;
			for (int i = 0; i < 16; ++i) cs |= (uint16_t)get_be_bit(taps.hdr, i + 55) << i;
			CRC<uint16_t> crc0(0xA8F4);
			taps.md = md;
			if (crc0.u64(md << 9) != cs) { taps.status = ST_HDR_CRC; header_log += "header CRC error.\n"; continue; }
			taps.mode = md & 255;
			if (!mode_params(taps.mode, mp)) { taps.status = ST_BAD_MODE; header_log += "operation mode " + std::to_string(taps.mode) + " unsupported.\n"; continue; }
			header_log += "oper mode: " + std::to_string(taps.mode) + "\n";
			if ((md >> 8) == 0 || (long long)(md >> 8) >= kCallSignLimit) { taps.status = ST_BAD_CALL; header_log += "call sign unsupported.\n"; continue; }
			base37_decode(taps.call_sign, md >> 8, 9);
			taps.call_sign[9] = 0;
			header_log += std::string("call sign: ") + taps.call_sign + "\n";
			okay = true;
		} while (skip_count--);
		if (!okay) return taps.status;

		int cons_rows = mp.cons_rows(), cons_cols = mp.cons_cols, code_off = -cons_cols / 2, mod_bits = mp.mod_bits;
		std::vector<cf> cons((size_t)mp.cons_cnt()), prev(cons_cols);
		for (int i = 0; i < symbol_pos + 2 * (symbol_len_ + guard_len_); ++i) buf = next_sample();
		for (int i = 0; i < symbol_len_; ++i) tdom[i] = buf[i] * osc();
		for (int i = 0; i < guard_len_; ++i) osc();
		fwd_(fdom.data(), tdom.data());
		for (int j = 0; j < cons_rows; ++j) {
			for (int i = 0; i < symbol_len_ + guard_len_; ++i) buf = next_sample();
			for (int i = 0; i < symbol_len_; ++i) tdom[i] = buf[i] * osc();
			for (int i = 0; i < guard_len_; ++i) osc();
			for (int i = 0; i < cons_cols; ++i) prev[i] = fdom[bin(i + code_off)];
			fwd_(fdom.data(), tdom.data());
			for (int i = 0; i < cons_cols; ++i) cons[cons_cols * j + i] = demod_or_erase(fdom[bin(i + code_off)], prev[i]);
		}
		taps.cons_raw = cons;
		// Theil–Sen phase line per row (decode.cc:479-504)
		TheilSen tse;
		std::vector<float> index(cons_cols), phase(cons_cols);
		taps.slope.resize(cons_rows); taps.yint.resize(cons_rows); taps.precision.resize(cons_rows);
		for (int j = 0; j < cons_rows; ++j) {
			for (int i = 0; i < cons_cols; ++i) {
				float b[3];
				mod_hard(mod_bits, b, cons[cons_cols * j + i]);
				index[i] = float(i + code_off);
				phase[i] = arg(cons[cons_cols * j + i] * conj(mod_map(mod_bits, b)));
			}
			tse.compute(index.data(), phase.data(), cons_cols);
			taps.slope[j] = tse.slope;
			taps.yint[j] = tse.yint;
			for (int i = 0; i < cons_cols; ++i) cons[cons_cols * j + i] = cons[cons_cols * j + i] * polar(1.f, -tse(float(i + code_off)));
		}
		taps.cons = cons;
		// cumulative Es/N0 and soft demapping (decode.cc:505-523)
		std::vector<float> code((size_t)1 << mp.code_order);
		float sp = 0, np = 0;
		for (int j = 0; j < cons_rows; ++j) {
			for (int i = 0; i < cons_cols; ++i) {
				float b[3];
				mod_hard(mod_bits, b, cons[cons_cols * j + i]);
				cf hard = mod_map(mod_bits, b);
				cf error = cons[cons_cols * j + i] - hard;
				sp += norm(hard);
				np += norm(error);
			}
			float precision = sp / np;
			taps.precision[j] = precision;
			for (int i = 0; i < cons_cols; ++i) mod_soft(mod_bits, &code[mod_bits * (cons_cols * j + i)], cons[cons_cols * j + i], precision);
		}
		// lengthen() (decode.cc:245-253)
		FrozenSet fs{frozen_table(mp.table)};
		int code_bits = 1 << mp.code_order;
		for (int i = code_bits - 1, j = mp.cons_bits - 1, k = mp.mesg_bits - 1; i >= 0; --i) {
			if (fs.frozen(i) || k-- < kCrcBits) code[i] = code[j--];
			else code[i] = 9000.f;
		}
		taps.llr = code;
		// list decoding + CRC-32 selection (decode.cc:530-545)
		std::vector<std::vector<uint8_t>> lanes;
		float metrics[8] = {0};
		int L = opt.list_size;
		if (L == 8) { PolarListDecoder<8> dec(mp.code_order, fs.bits); dec.r0_max = opt.r0_max; dec.decode(code.data(), lanes, metrics); taps.forks = dec.forks; }
		else { PolarListDecoder<4> dec(mp.code_order, fs.bits); dec.r0_max = opt.r0_max; dec.decode(code.data(), lanes, metrics); taps.forks = dec.forks; }
		for (int k = 0; k < L; ++k) taps.metrics[k] = metrics[k];
		int best = -1;
		std::vector<uint8_t> mesg(mp.mesg_bits);
		for (int k = 0; k < L && best < 0; ++k) {
			for (int i = 0, j = 0; i < code_bits && j < mp.mesg_bits; ++i) if (!fs.frozen(i)) mesg[j++] = lanes[k][i];
			CRC<uint32_t> crc1(0xD419CC15);
			for (int i = 0; i < kCrcBits; ++i) crc1.bit(mesg[i]);
			if (crc1() == 0) best = k;
		}
		taps.best_lane = best;
		if (best < 0) { taps.status = ST_PAYLOAD_CRC; return taps.status; }
		int flips = 0;
		for (int i = 0, j = 0; i < kDataBits; ++i, ++j) {
			while (fs.frozen(j)) ++j;
			bool received = code[j] < 0.f, decoded = mesg[i];
			flips += received != decoded;
			set_le_bit(out, i, decoded);
		}
		taps.flips = flips;
		taps.status = ST_OK;
		return ST_OK;
	}
};

static inline void descramble(uint8_t *data) // decode.cc:613-615
{
	Xorshift32 prng;
	for (int i = 0; i < kDataBytes; ++i) data[i] ^= (uint8_t)prng();
}

} // namespace ref
