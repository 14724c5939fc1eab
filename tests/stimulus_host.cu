// tests/stimulus_host.cu — TEST HELPER.  Runs the product's stimulus generator (modem_b200/csrc/stimulus.cuh: the same
// routines the kernels of stimulus.cu call, compiled for the host, one "thread") so that the CPU suite can check the
// transmitter arithmetic, the guard cross-fade gather and the impairment chain against the oracle without a GPU.
// Mirrors the orchestration of ofdmtx_encode_batch (stimulus.cu) for one window.
#include <cstddef>
#include <cstring>
#include "../modem_b200/csrc/stimulus.cuh"
#include "../modem_b200/csrc/tx_tables.h"
#include <vector>
#include <cstring>

using namespace ofdmrx;

template <int N>
static void symbols(const TxParams &p, const std::vector<uint32_t> &code, int frames)
{
	std::vector<cfx> b0(N), b1(N), car(kTxMaxCarriers), acc(kTxMaxCarriers);
	for (int s = 0; s < 3; ++s) {
		for (int c = 0; c < p.spec[s].count; ++c) car[c] = p.common_fdom[s * kTxMaxCarriers + c];
		tx_symbol_core<N>(car.data(), acc.data(), p.spec[s], s != kTxSymSc, b0.data(), b1.data(), p.tw_sym, p.tw_4n,
			p.tdom_common + (size_t)s * N, 0, 1);
	}
	for (int f = 0; f < frames; ++f)
		for (int row = 0; row < p.rows; ++row) {
			for (int c = 0; c < p.cols; ++c)
				car[c] = tx_data_carrier(code.data() + (size_t)f * kTxCodeWords, p.cols, p.mod_bits, row, c, p.common_fdom[kTxSymPilot * kTxMaxCarriers + c]);
			tx_symbol_core<N>(car.data(), acc.data(), TxSpec{p.code_off, 1, p.cols}, true, b0.data(), b1.data(), p.tw_sym, p.tw_4n,
				p.tdom + ((size_t)f * p.rows + row) * N, 0, 1);
		}
}

extern "C" int stimulus_host_code(const uint8_t *payload, int table, uint32_t *out)
{
	std::vector<uint32_t> tbl = make_frozen(kCodeOrder, table ? 64512 : 64800, kCrcBits);
	tbl.resize(4096);
	uint32_t acc = 0;
	for (int w = 0; w < 2048; ++w) { tbl[2048 + w] = acc; acc += 32 - __builtin_popcount(tbl[w]); }
	std::vector<uint32_t> scr(kDataBytes / 4, 0), mesg(kTxMesgWords), cw(kTxCodeWords);
	uint32_t y = 2463534242u;
	for (int i = 0; i < kDataBytes; ++i) { y ^= y << 13; y ^= y >> 17; y ^= y << 5; scr[i / 4] |= (uint32_t)(y & 255u) << (8 * (i % 4)); }
	uint32_t lut[256];
	crc32_table(0xD419CC15u, lut);
	tx_code_core(payload, scr.data(), lut, tbl.data(), tbl.data() + 2048, mesg.data(), cw.data(), out, 0, 1);
	return 0;
}

// one window of `fpw` frames; window_index only selects the noise stream (counter plane window_index of key seed)
extern "C" long long stimulus_host(int rate, int mode, int freq_off, long long call_sign, const uint8_t *payloads, int fpw,
	const TxImpair *imp, int window_index, int format, void *out, long long stride)
{
	if (!tx_check_args(rate, format == 0 ? 1 : 2, freq_off, mode, call_sign)) return -22;
	const ModeInfo mi = mode_info(mode);
	const int N = (1280 * rate) / 8000;
	TxParams p{};
	p.rate = rate; p.sym_len = N; p.guard_len = N / 8; p.pitch = N + N / 8;
	p.cols = mi.cols; p.mod_bits = mi.mod_bits; p.rows = mi.rows; p.cons_bits = mi.cons_bits; p.table = mi.table;
	p.frames_per_window = fpw;
	p.n_sym = 2 + fpw * (3 + mi.rows);
	p.len = tx_window_len(rate, mode, fpw);
	std::vector<float> common((size_t)3 * kTxMaxCarriers * 2);
	TxCarriers spec[3];
	tx_common_symbols(rate, mode, freq_off, call_sign, common.data(), spec);
	for (int i = 0; i < 3; ++i) p.spec[i] = TxSpec{spec[i].first, spec[i].step, spec[i].count};
	p.code_off = spec[0].first;
	std::vector<float> tw = twiddles(N, -1), tw4 = twiddles(4 * N, -1), ramp = tx_guard_ramp(N / 8);
	std::vector<cfx> tdom_common((size_t)3 * N), tdom((size_t)fpw * mi.rows * N);
	p.common_fdom = reinterpret_cast<const cfx *>(common.data());
	p.tw_sym = reinterpret_cast<const cfx *>(tw.data());
	p.tw_4n = reinterpret_cast<const cfx *>(tw4.data());
	p.ramp = ramp.data();
	p.tdom_common = tdom_common.data();
	p.tdom = tdom.data();
	std::vector<uint32_t> code((size_t)fpw * kTxCodeWords);
	for (int f = 0; f < fpw; ++f) stimulus_host_code(payloads + (size_t)f * kDataBytes, mi.table, code.data() + (size_t)f * kTxCodeWords);
#define SYMBOLS(R) symbols<Geo<R>::kSymLen>(p, code, fpw)
	OFDMRX_FOR_RATE(rate, SYMBOLS);
#undef SYMBOLS
	TxImpair im{};
	const bool has_imp = imp && (imp->multipath || imp->cfo_hz != 0.f || imp->sfo_ppm != 0.f || imp->awgn);
	if (has_imp) { std::memcpy(&im, imp, offsetof(TxImpair, window0)); im.window0 = (unsigned long long)window_index; } // the caller's struct is ofdmtx_impairments (no window0)
	const bool sfo = has_imp && im.sfo_ppm != 0.f;
	const long long nout = tx_resampled_len(p.len, sfo ? im.sfo_ppm : 0.f);
	if (stride < nout) return -22;
	std::vector<cfx> iq;
	if (sfo) {
		iq.resize(p.len);
		for (long long n = 0; n < p.len; ++n) iq[n] = tx_channel_sample(p, im, 0, n);
	}
	for (long long n = 0; n < stride; ++n) {
		cfx v = make_float2(0.f, 0.f);
		if (n < nout) {
			v = sfo ? tx_resample(iq.data(), p.len, im.sfo_ppm, n) : has_imp ? tx_channel_sample(p, im, 0, n) : tx_stream_sample(p, 0, n);
			if (has_imp && im.awgn) { const cfx z = tx_noise(im, 0, n); v.x += z.x; v.y += z.y; }
		}
		tx_store(out, format, n, v);
	}
	return nout;
}
