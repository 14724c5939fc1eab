// oracle/oracle_capi.cc — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU oracle for ctypes
// (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg — nothing else may load this).
#include "ref_modem.hh"
#include <thread>
#include <atomic>

using namespace ref;

extern "C" {

struct ref_impair_c {
	int32_t multipath;
	float cfo_hz, sfo_ppm;
	int32_t awgn;
	float awgn_db;
	uint64_t seed;
};

struct ref_taps_c {
	int32_t status, detections, t_fire, symbol_pos, sc_pos, index_max, shift, pos_err, osd_unique, mode, best_lane, flips, rows, cols;
	float timing_max, frac_cfo, cfo_rad;
	int64_t md, forks, osd_visited;
	int8_t soft[256];
	uint8_t hdr[32];
	char call_sign[12];
	float metrics[8];
	float slope[128], yint[128], precision[128];
	float cons_raw[2 * 32400], cons[2 * 32400];
	float llr[65536];
};

// ---- stimulus ------------------------------------------------------------------------------------------
// returns number of sample frames written (<0 on error).  payloads: count x 5380 plain bytes.
int64_t ref_encode_pcm16(const uint8_t *payloads, int count, int rate, int channels, int freq_off, const char *call_sign,
	int mode, const ref_impair_c *imp, int16_t *out, int64_t max_frames)
{
	long long cs = base37_encode(call_sign);
	if (!Transmitter::check_args(rate, channels, freq_off, mode, cs)) return -1;
	Transmitter tx(rate);
	std::vector<cf> s;
	if (!tx.encode(s, payloads, count, freq_off, cs, mode)) return -1;
	if (imp) {
		Impair im;
		im.multipath = imp->multipath; im.cfo_hz = imp->cfo_hz; im.sfo_ppm = imp->sfo_ppm;
		im.awgn = imp->awgn; im.awgn_db = imp->awgn_db; im.seed = imp->seed;
		apply_impairments(s, rate, im);
	}
	if ((int64_t)s.size() > max_frames) return -2;
	std::vector<int16_t> pcm;
	to_pcm16(s, channels, pcm);
	std::memcpy(out, pcm.data(), pcm.size() * sizeof(int16_t));
	return (int64_t)s.size();
}

// one frame per window, seeds[i] -> payload bytes from xorshift-free LCG so Python can regenerate them
void ref_make_payload(uint64_t seed, uint8_t *out)
{
	uint64_t x = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
	for (int i = 0; i < kDataBytes; ++i) {
		x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
		out[i] = (uint8_t)((x * 0x2545F4914F6CDD1Dull) >> 56);
	}
}

// batch: n windows of `stride` sample frames each (zero padded), payload of window i from ref_make_payload(seed0+i);
// imp may be NULL; per-window awgn seed = imp->seed + i.  Threaded.
int ref_encode_batch_pcm16(int n, uint64_t seed0, int rate, int channels, int freq_off, const char *call_sign, int mode,
	const ref_impair_c *imp, int16_t *out, int64_t stride, int32_t *n_samples, uint8_t *payloads_out, int nthreads)
{
	std::atomic<int> next(0), err(0);
	auto work = [&]() {
		for (;;) {
			int i = next++;
			if (i >= n) break;
			uint8_t pl[kDataBytes];
			ref_make_payload(seed0 + i, pl);
			if (payloads_out) std::memcpy(payloads_out + (size_t)i * kDataBytes, pl, kDataBytes);
			ref_impair_c im;
			if (imp) { im = *imp; im.seed = imp->seed + i; }
			int16_t *dst = out + (size_t)i * stride * channels;
			std::memset(dst, 0, (size_t)stride * channels * sizeof(int16_t));
			int64_t got = ref_encode_pcm16(pl, 1, rate, channels, freq_off, call_sign, mode, imp ? &im : nullptr, dst, stride);
			if (got < 0) { err = 1; got = 0; }
			if (n_samples) n_samples[i] = (int32_t)got;
		}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, nthreads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
	return err ? -1 : 0;
}

// ---- receiver ------------------------------------------------------------------------------------------
static void fill_taps(ref_taps_c *t, const Taps &s)
{
	std::memset(t, 0, sizeof(*t));
	t->status = s.status; t->detections = s.detections; t->t_fire = s.t_fire; t->symbol_pos = s.symbol_pos; t->sc_pos = s.sc_pos;
	t->index_max = s.index_max; t->shift = s.shift; t->pos_err = s.pos_err; t->osd_unique = s.osd_unique; t->mode = s.mode;
	t->best_lane = s.best_lane; t->flips = s.flips; t->timing_max = s.timing_max; t->frac_cfo = s.frac_cfo; t->cfo_rad = s.cfo_rad;
	t->md = (int64_t)s.md; t->forks = s.forks; t->osd_visited = s.osd_visited;
	std::memcpy(t->soft, s.soft, 255); std::memcpy(t->hdr, s.hdr, 32); std::memcpy(t->call_sign, s.call_sign, 10);
	std::memcpy(t->metrics, s.metrics, sizeof(t->metrics));
	t->rows = (int)s.slope.size();
	t->cols = t->rows ? (int)(s.cons.size() / t->rows) : 0;
	for (int j = 0; j < t->rows && j < 128; ++j) { t->slope[j] = s.slope[j]; t->yint[j] = s.yint[j]; t->precision[j] = s.precision[j]; }
	for (size_t i = 0; i < s.cons.size() && i < 32400; ++i) {
		t->cons_raw[2 * i] = s.cons_raw[i].re; t->cons_raw[2 * i + 1] = s.cons_raw[i].im;
		t->cons[2 * i] = s.cons[i].re; t->cons[2 * i + 1] = s.cons[i].im;
	}
	for (size_t i = 0; i < s.llr.size() && i < 65536; ++i) t->llr[i] = s.llr[i];
}

// pcm: interleaved int16; converted like ReadWAV (v / 32767).  out: 5380 bytes, de-scrambled like decode.cc main()
// (zero-filled before de-scrambling on failure — the reference writes an uninitialised buffer there).
int ref_decode_pcm16(const int16_t *pcm, int64_t n_frames, int channels, int rate, int skip, int list_size, int r0_max,
	int osd_literal, uint8_t *out, ref_taps_c *taps)
{
	std::vector<float> f((size_t)n_frames * channels);
	for (size_t i = 0; i < f.size(); ++i) f[i] = float(pcm[i]) / 32767.f;
	Receiver rx(rate);
	RxOptions opt;
	opt.list_size = list_size; opt.r0_max = r0_max; opt.osd_literal = osd_literal != 0;
	uint8_t buf[kDataBytes];
	std::memset(buf, 0, sizeof(buf));
	int st = rx.run(buf, f.data(), (size_t)n_frames, channels, skip, opt);
	descramble(buf);
	if (out) std::memcpy(out, buf, kDataBytes);
	if (taps) fill_taps(taps, rx.taps);
	return st;
}

// front-end stage taps of one window: iq_out float2[n_frames + 1], timing_out float[n_frames + 1] (one entry per stream step).
// pcm16 != NULL: interleaved int16 converted like ReadWAV (v / 32767); else pcmf: interleaved floats as ReadWAV delivers them.
int64_t ref_front_taps(const int16_t *pcm16, const float *pcmf, int64_t n_frames, int channels, int rate, float *iq_out, float *timing_out)
{
	std::vector<float> f((size_t)n_frames * channels);
	for (size_t i = 0; i < f.size(); ++i) f[i] = pcm16 ? float(pcm16[i]) / 32767.f : pcmf[i];
	Receiver rx(rate);
	std::vector<cf> iq;
	std::vector<float> timing;
	rx.front_taps(f.data(), (size_t)n_frames, channels, iq, timing);
	for (size_t i = 0; i < iq.size(); ++i) { iq_out[2 * i] = iq[i].re; iq_out[2 * i + 1] = iq[i].im; timing_out[i] = timing[i]; }
	return (int64_t)iq.size();
}
// float samples as ReadWAV<float> delivers them (8 / 24 / 32-bit files): same receiver, no int16 conversion
int ref_decode_f32(const float *pcm, int64_t n_frames, int channels, int rate, int skip, uint8_t *out, ref_taps_c *taps)
{
	Receiver rx(rate);
	RxOptions opt;
	uint8_t buf[kDataBytes];
	std::memset(buf, 0, sizeof(buf));
	int st = rx.run(buf, pcm, (size_t)n_frames, channels, skip, opt);
	descramble(buf);
	if (out) std::memcpy(out, buf, kDataBytes);
	if (taps) fill_taps(taps, rx.taps);
	return st;
}

// threaded batch: windows of `stride` frames; status[i], payload_out[i*5380]
void ref_decode_batch_pcm16(const int16_t *pcm, int n, int64_t stride, const int32_t *n_samples, int channels, int rate, int skip,
	int list_size, uint8_t *payload_out, int32_t *status, int nthreads)
{
	std::atomic<int> next(0);
	auto work = [&]() {
		for (;;) {
			int i = next++;
			if (i >= n) break;
			status[i] = ref_decode_pcm16(pcm + (size_t)i * stride * channels, n_samples ? n_samples[i] : stride, channels, rate, skip,
				list_size, 1 << 16, 0, payload_out + (size_t)i * kDataBytes, nullptr);
		}
	};
	std::vector<std::thread> th;
	for (int t = 0; t < std::max(1, nthreads); ++t) th.emplace_back(work);
	for (auto &t : th) t.join();
}

// ---- primitives for known-answer tests and kernel-level parity --------------------------------------------
void ref_mls(int poly, int n, uint8_t *out) { MLS m(poly); for (int i = 0; i < n; ++i) out[i] = m(); }
uint32_t ref_crc16_u64(uint64_t v) { CRC<uint16_t> c(0xA8F4); return c.u64(v); }
uint32_t ref_crc32_bytes(const uint8_t *d, int n) { CRC<uint32_t> c(0xD419CC15); for (int i = 0; i < n; ++i) c.byte(d[i]); return c(); }
uint32_t ref_crc32_bits(const uint8_t *bits, int n) { CRC<uint32_t> c(0xD419CC15); for (int i = 0; i < n; ++i) c.bit(bits[i]); return c(); }
void ref_xorshift(int n, uint32_t *out) { Xorshift32 x; for (int i = 0; i < n; ++i) out[i] = x(); }
int64_t ref_base37(const char *s) { return base37_encode(s); }
void ref_frozen_table(int table, uint32_t *out) { std::memcpy(out, frozen_table(table), 2048 * 4); }
void ref_bch_generator(uint8_t *gen185) { BCH255_71 b; std::memcpy(gen185, b.gen, 185); }
void ref_bch_genmat(int8_t *genmat) { BCH255_71 b; b.matrix(genmat); }
void ref_bch_encode(const uint8_t *data71, uint8_t *parity184) { BCH255_71 b; b.encode_bits(data71, parity184); }
int ref_osd(const int8_t *soft, int literal, uint8_t *out32, int64_t *visited)
{
	static const BCH255_71 bch;
	int8_t genmat[255 * 71];
	bch.matrix(genmat);
	OSD255_71 osd;
	bool u = literal ? osd.decode_full(out32, soft, genmat) : osd.decode_pruned(out32, soft, genmat);
	if (visited) *visited = osd.visited;
	return u;
}
void ref_fft(int n, int sign, const float *in, float *out)
{
	FFT f(n, sign);
	f(reinterpret_cast<cf *>(out), reinterpret_cast<const cf *>(in));
}
void ref_theil_sen(const float *x, const float *y, int n, float *slope, float *yint)
{
	TheilSen t;
	t.compute(x, y, n);
	*slope = t.slope; *yint = t.yint;
}
void ref_payload_to_code(const uint8_t *payload_plain, int mode, uint8_t *code_bits)
{
	ModeParams mp;
	mode_params(mode, mp);
	uint8_t scr[kDataBytes];
	Xorshift32 prng;
	for (int i = 0; i < kDataBytes; ++i) scr[i] = payload_plain[i] ^ (uint8_t)prng();
	std::vector<uint8_t> code;
	Transmitter::payload_to_code(scr, mp, code);
	std::memcpy(code_bits, code.data(), code.size());
}
// full polar list decode of one codeword: llr[65536] -> lanes[L][65536] codeword bits (ascending metric), metrics[L];
// also CRC-selected payload (de-scrambled) like decode.cc:532-555.  returns best lane or -1.
int ref_polar_decode(const float *llr, int table, int list_size, int r0_max, uint8_t *lanes_out, float *metrics, uint8_t *payload, int32_t *flips)
{
	FrozenSet fs{frozen_table(table)};
	std::vector<std::vector<uint8_t>> lanes;
	float m[8] = {0};
	if (list_size == 8) { PolarListDecoder<8> d(16, fs.bits); d.r0_max = r0_max; d.decode(llr, lanes, m); }
	else { PolarListDecoder<4> d(16, fs.bits); d.r0_max = r0_max; d.decode(llr, lanes, m); }
	for (int k = 0; k < list_size; ++k) {
		if (lanes_out) std::memcpy(lanes_out + (size_t)k * 65536, lanes[k].data(), 65536);
		if (metrics) metrics[k] = m[k];
	}
	int mesg_bits = table ? 44096 : 43808;
	std::vector<uint8_t> mesg(mesg_bits);
	int best = -1;
	for (int k = 0; k < list_size && best < 0; ++k) {
		for (int i = 0, j = 0; i < 65536 && j < mesg_bits; ++i) if (!fs.frozen(i)) mesg[j++] = lanes[k][i];
		CRC<uint32_t> c(0xD419CC15);
		for (int i = 0; i < kCrcBits; ++i) c.bit(mesg[i]);
		if (c() == 0) best = k;
	}
	if (payload) {
		std::memset(payload, 0, kDataBytes);
		int fl = 0;
		if (best >= 0)
			for (int i = 0, j = 0; i < kDataBits; ++i, ++j) {
				while (fs.frozen(j)) ++j;
				fl += (llr[j] < 0.f) != (bool)mesg[i];
				set_le_bit(payload, i, mesg[i]);
			}
		descramble(payload);
		if (flips) *flips = best >= 0 ? fl : -1;
	}
	return best;
}
// list decoder on an arbitrary (small) code: n = 2^order LLRs, frozen = n/32 mask words (n >= 32) -> lanes[8][n] + metrics[8];
// for cross-checks against an independent textbook implementation (tests/test_oracle_independent.py)
void ref_polar_decode_any(int order, const uint32_t *frozen, const float *llr, int r0_max, uint8_t *lanes_out, float *metrics)
{
	PolarListDecoder<8> d(order, frozen);
	d.r0_max = r0_max;
	std::vector<std::vector<uint8_t>> lanes;
	float m[8] = {0};
	d.decode(llr, lanes, m);
	const int n = 1 << order;
	for (int k = 0; k < 8; ++k) { std::memcpy(lanes_out + (size_t)k * n, lanes[k].data(), n); metrics[k] = m[k]; }
}
int ref_taps_size() { return (int)sizeof(ref_taps_c); }

} // extern "C"
