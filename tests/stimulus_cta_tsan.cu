// tests/stimulus_cta_tsan.cu — TEST HELPER.  The CTA-cooperative routines of the stimulus generator (stimulus.cuh: payload ->
// code bits, one OFDM symbol) and the Stockham FFT they share with the receiver (fft.cuh) executed by several HOST threads
// with OFDMRX_CTA_SYNC() mapped to a real barrier, under ThreadSanitizer: a missing __syncthreads() in the device code is a
// data race here.  Results are also compared with the one-thread run (bit-identical: the arithmetic per element is the same).
// Build: nvcc -std=c++20 -DOFDMRX_HOST_CTA -Xcompiler -fsanitize=thread,-ffp-contract=off ... (modem_b200/build.py: STIMTSAN)
#define OFDMRX_HOST_CTA 1
#include "../modem_b200/csrc/stimulus.cuh"
#include "../modem_b200/csrc/tx_tables.h"
#include <barrier>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

using namespace ofdmrx;

static std::unique_ptr<std::barrier<>> g_barrier;
namespace ofdmrx { void ofdmrx_host_cta_sync() { if (g_barrier) g_barrier->arrive_and_wait(); } }

static void run_cta(int nthr, const std::function<void(int, int)> &body)
{
	if (nthr == 1) { g_barrier.reset(); body(0, 1); return; }
	g_barrier = std::make_unique<std::barrier<>>(nthr);
	std::vector<std::thread> th;
	for (int t = 0; t < nthr; ++t) th.emplace_back(body, t, nthr);
	for (auto &t : th) t.join();
	g_barrier.reset();
}

template <int N>
static int check_symbol(int nthr, int rate, int mode)
{
	std::vector<float> common((size_t)3 * kTxMaxCarriers * 2);
	TxCarriers spec[3];
	tx_common_symbols(rate, mode, 2000, 1263905687425LL, common.data(), spec);
	std::vector<float> tw = twiddles(N, -1), tw4 = twiddles(4 * N, -1);
	int bad = 0;
	for (int s = 0; s < 3; ++s) {
		std::vector<cfx> out[2];
		for (int pass = 0; pass < 2; ++pass) {
			std::vector<cfx> b0(N), b1(N), car(kTxMaxCarriers), acc(kTxMaxCarriers);
			out[pass].assign(N, make_float2(0.f, 0.f));
			std::memcpy(car.data(), common.data() + (size_t)s * kTxMaxCarriers * 2, sizeof(cfx) * kTxMaxCarriers);
			TxSpec sp{spec[s].first, spec[s].step, spec[s].count};
			run_cta(pass ? nthr : 1, [&](int tid, int n) {
				tx_symbol_core<N>(car.data(), acc.data(), sp, s != kTxSymSc, b0.data(), b1.data(), reinterpret_cast<const cfx *>(tw.data()),
					reinterpret_cast<const cfx *>(tw4.data()), out[pass].data(), tid, n);
			});
		}
		bad += std::memcmp(out[0].data(), out[1].data(), sizeof(cfx) * N) != 0;
	}
	return bad;
}

extern "C" int stimulus_cta_check(int nthr)
{
	int bad = 0;
	// payload -> code bits
	{
		std::vector<uint32_t> tbl = make_frozen(kCodeOrder, 64800, kCrcBits);
		tbl.resize(4096);
		uint32_t acc = 0;
		for (int w = 0; w < 2048; ++w) { tbl[2048 + w] = acc; acc += 32 - __builtin_popcount(tbl[w]); }
		std::vector<uint32_t> scr(kDataBytes / 4, 0x5a5a1234u);
		uint32_t lut[256];
		crc32_table(0xD419CC15u, lut);
		std::vector<uint8_t> payload(kDataBytes);
		for (int i = 0; i < kDataBytes; ++i) payload[i] = (uint8_t)(i * 131 + 7);
		std::vector<uint32_t> out[2];
		for (int pass = 0; pass < 2; ++pass) {
			std::vector<uint32_t> mesg(kTxMesgWords), cw(kTxCodeWords);
			out[pass].assign(kTxCodeWords, 0);
			run_cta(pass ? nthr : 1, [&](int tid, int n) {
				tx_code_core(payload.data(), scr.data(), lut, tbl.data(), tbl.data() + 2048, mesg.data(), cw.data(), out[pass].data(), tid, n);
			});
		}
		bad += out[0] != out[1];
	}
	bad += check_symbol<1280>(nthr, 8000, 6);
	bad += check_symbol<7056>(nthr, 44100, 13);
	// the receiver's transform lengths through the same barrier
	{
		std::vector<float> tw = twiddles(640, -1);
		std::vector<cfx> a[2], b(640);
		for (int pass = 0; pass < 2; ++pass) {
			a[pass].resize(640);
			for (int i = 0; i < 640; ++i) a[pass][i] = make_float2((float)((i * 37) % 101) - 50.f, (float)((i * 11) % 53));
			cfx *res = nullptr;
			run_cta(pass ? nthr : 1, [&](int tid, int n) {
				cfx *r = fft_fwd<640>(a[pass].data(), b.data(), reinterpret_cast<const cfx *>(tw.data()), tid, n);
				if (tid == 0) res = r;
			});
			if (res != a[pass].data()) std::memcpy(a[pass].data(), res, sizeof(cfx) * 640);
		}
		bad += std::memcmp(a[0].data(), a[1].data(), sizeof(cfx) * 640) != 0;
	}
	return bad;
}

int main(int argc, char **argv)
{
	int nthr = argc > 1 ? std::atoi(argv[1]) : 6;
	int bad = stimulus_cta_check(nthr);
	std::printf("threads %d: %d mismatching results\n", nthr, bad);
	return bad != 0;
}
