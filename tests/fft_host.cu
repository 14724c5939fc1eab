// tests/fft_host.cu — TEST HELPER.  Runs the product's Stockham FFT plans (modem_b200/csrc/fft.cuh: the same pass and
// butterfly templates the kernels use, compiled for the host, one "thread") so that the CPU suite can check every length
// of the four sample rates — 640, 1280, 2560, 3528, 3840, 7056, 7680; radices 2, 3, 4, 5, 7 — against numpy.
#include "../modem_b200/csrc/fft.cuh"
#include <vector>
#include <cstring>

using namespace ofdmrx;

template <int N>
static void run(const float *in, float *out)
{
	std::vector<cfx> a(N), b(N), tw(N);
	std::vector<float> t = twiddles(N, -1);
	for (int i = 0; i < N; ++i) { a[i] = make_float2(in[2 * i], in[2 * i + 1]); tw[i] = make_float2(t[2 * i], t[2 * i + 1]); }
	const cfx *r = fft_fwd<N>(a.data(), b.data(), tw.data(), 0, 1);
	std::memcpy(out, r, sizeof(cfx) * N);
}

extern "C" int fft_host(int n, const float *in, float *out)
{
	switch (n) {
	case 640: run<640>(in, out); return 0;
	case 1280: run<1280>(in, out); return 0;
	case 2560: run<2560>(in, out); return 0;
	case 3528: run<3528>(in, out); return 0;
	case 3840: run<3840>(in, out); return 0;
	case 7056: run<7056>(in, out); return 0;
	case 7680: run<7680>(in, out); return 0;
	}
	return -1;
}
