// modem_b200/csrc/host/encode_main.cc — `encode OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT..`: the reference
// transmitter's command line (/root/reference/encode.cc:337-446) as a thin C++ host driver over the device-side stimulus
// generator of libofdmrx (include/ofdmtx.h).  Same argv rules and messages, "-" for stdout / stdin, one frame per INPUT
// file in one stream, 1 s of silence either side, BITS in {8, 16, 24, 32}.  There is no CPU fallback: without a B200 it
// exits 1.  (The PAPR report the reference prints on stderr is not produced.)
#include "ofdmrx.h"
#include "ofdmtx.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
	if (argc < 9) {
		std::cerr << "usage: " << argv[0] << " OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT.." << std::endl;
		return 1;
	}
	std::string output_name = argv[1];
	if (output_name == "-") output_name = "/dev/stdout";
	const int rate = std::atoi(argv[2]), bits = std::atoi(argv[3]), chan = std::atoi(argv[4]);
	const int freq_off = std::atoi(argv[5]), mode = std::atoi(argv[6]);
	if (mode < 6 || mode > 13) { std::cerr << "Unsupported operation mode." << std::endl; return 1; }
	const long long call_sign = ofdmtx_call_sign(argv[7]);
	if (call_sign <= 0 || call_sign >= 129961739795077LL) { std::cerr << "Unsupported call sign." << std::endl; return 1; }
	static const int bw_tab[8] = {2700, 2500, 2500, 2250, 3200, 2400, 2400, 1600}; // encode.cc:363-387
	const int bw = bw_tab[mode - 6];
	if ((chan == 1 && freq_off < bw / 2) || freq_off < bw / 2 - rate / 2 || freq_off > rate / 2 - bw / 2) {
		std::cerr << "Unsupported frequency offset." << std::endl;
		return 1;
	}
	if (freq_off % 50) { std::cerr << "Frequency offset must be divisible by 50." << std::endl; return 1; }
	if (rate != 8000 && rate != 16000 && rate != 44100 && rate != 48000) { std::cerr << "Unsupported sample rate." << std::endl; return 1; }
	if ((bits != 8 && bits != 16 && bits != 24 && bits != 32) || chan < 1 || chan > 2) { std::cerr << "Unsupported sample format." << std::endl; return 1; }
	const int count = argc - 8;
	std::vector<uint8_t> data((size_t)count * OFDMRX_PAYLOAD_BYTES);
	for (int j = 0; j < count; ++j) {
		std::string name = argv[j + 8];
		if (argc == 9 && name == "-") name = "/dev/stdin";
		std::ifstream in(name, std::ios::binary);
		if (in.bad()) { std::cerr << "Couldn't open file \"" << name << "\" for reading." << std::endl; return 1; }
		for (int i = 0; i < OFDMRX_PAYLOAD_BYTES; ++i) data[(size_t)j * OFDMRX_PAYLOAD_BYTES + i] = (uint8_t)in.get(); // encode.cc:414-415
	}
	const int64_t len = ofdmtx_window_samples(rate, mode, count);
	ofdmtx_t *h = nullptr;
	int rc = ofdmtx_create(&h, 0, rate, 1, count);
	if (rc) { std::cerr << "ofdmtx_create failed (" << rc << "): a B200 is required, there is no CPU path" << std::endl; return 1; }
	std::vector<float> iq((size_t)len * 2);
	// CHANNELS = 1 writes the real part of the same analytic stream (encode.cc:127-128: both channels are always produced)
	rc = ofdmtx_encode_batch(h, data.data(), OFDMRX_MEM_HOST, 1, mode, call_sign, freq_off, nullptr, iq.data(), OFDMRX_MEM_HOST,
		OFDMRX_FMT_F32_IQ, len, nullptr, nullptr);
	ofdmtx_destroy(h);
	if (rc) { std::cerr << "ofdmtx_encode_batch failed (" << rc << ")" << std::endl; return 1; }
	const int bytes = bits / 8;
	const size_t n = (size_t)len * chan;
	std::vector<uint8_t> o(44 + n * bytes);
	auto wr32 = [&](size_t p, uint32_t v) { for (int b = 0; b < 4; ++b) o[p + b] = (v >> (8 * b)) & 255; };
	auto wr16 = [&](size_t p, uint32_t v) { for (int b = 0; b < 2; ++b) o[p + b] = (v >> (8 * b)) & 255; };
	std::memcpy(&o[0], "RIFF", 4); wr32(4, (uint32_t)(36 + n * bytes)); std::memcpy(&o[8], "WAVEfmt ", 8);
	wr32(16, 16); wr16(20, 1); wr16(22, chan); wr32(24, rate); wr32(28, rate * chan * bytes);
	wr16(32, chan * bytes); wr16(34, bits); std::memcpy(&o[36], "data", 4); wr32(40, (uint32_t)(n * bytes));
	const float fac = (float)(std::ldexp(1.0, bits - 1) - 1.0); // DSP::WritePCM scale: 2^(bits-1) - 1 in fp32, 8-bit samples offset by 128
	for (size_t i = 0; i < n; ++i) {
		const float x = std::min(std::max(chan == 2 ? iq[i] : iq[2 * i], -1.f), 1.f);
		const int64_t v = std::min<int64_t>((int64_t)std::nearbyint(fac * x), 2147483647LL) + (bytes == 1 ? 128 : 0);
		for (int b = 0; b < bytes; ++b) o[44 + i * bytes + b] = (uint8_t)((v >> (8 * b)) & 255);
	}
	std::ofstream out(output_name, std::ios::binary | std::ios::trunc);
	if (out.bad()) { std::cerr << "Couldn't open file \"" << output_name << "\" for writing." << std::endl; return 1; }
	out.write(reinterpret_cast<const char *>(o.data()), o.size());
	return 0;
}
