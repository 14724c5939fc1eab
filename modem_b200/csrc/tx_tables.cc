// modem_b200/csrc/tx_tables.cc — see tx_tables.h.  Nothing here is copied from the reference: the symbol values are
// rebuilt from the MLS generators, the BCH generator rows and the CRC of host_tables.cc.
#include "tx_tables.h"
#include <cmath>
#include <cstring>

namespace ofdmrx {

long long base37_encode(const char *str)
{
	long long acc = 0;
	for (; *str; ++str) {
		const char c = *str;
		int v;
		if (c == ' ') v = 0;
		else if (c >= '0' && c <= '9') v = c - '0' + 1;
		else if (c >= 'a' && c <= 'z') v = c - 'a' + 11;
		else if (c >= 'A' && c <= 'Z') v = c - 'A' + 11;
		else return -1;
		acc = acc * 37 + v;
	}
	return acc;
}

int tx_band_width(int mode)
{
	switch (mode) {
	case 6: return 2700;
	case 7: case 8: return 2500;
	case 9: return 2250;
	case 10: return 3200;
	case 11: case 12: return 2400;
	case 13: return 1600;
	}
	return 0;
}

bool tx_check_args(int rate, int channels, int freq_off, int mode, long long call_sign)
{
	if (rate != 8000 && rate != 16000 && rate != 44100 && rate != 48000) return false;
	if (channels != 1 && channels != 2) return false;
	if (mode < 6 || mode > 13) return false;
	if (call_sign <= 0 || call_sign >= kCallSignLimit) return false;
	const int bw = tx_band_width(mode);
	if ((channels == 1 && freq_off < bw / 2) || freq_off < bw / 2 - rate / 2 || freq_off > rate / 2 - bw / 2) return false;
	return freq_off % 50 == 0;
}

long long tx_window_len(int rate, int mode, int frames)
{
	const int pitch = (1280 * rate) / 8000 + (1280 * rate) / 8000 / 8;
	return 2LL * rate + (2LL + (long long)frames * (3 + mode_info(mode).rows)) * pitch;
}

std::vector<float> tx_guard_ramp(int guard_len)
{
	std::vector<float> r(guard_len);
	const float pi = 3.14159265358979323846f;
	for (int i = 0; i < guard_len; ++i) {
		float x = float(i) / float(guard_len - 1);
		r[i] = 0.5f * (1.f - std::cos(pi * x));
	}
	return r;
}

void tx_common_symbols(int rate, int mode, int freq_off, long long call_sign, float *values, TxCarriers spec[3])
{
	const int sym_len = (1280 * rate) / 8000;
	const ModeInfo mi = mode_info(mode);
	const int offset = (freq_off * sym_len) / rate; // encode.cc:283
	std::memset(values, 0, sizeof(float) * 3 * 512 * 2);
	// pilot: MLS2 BPSK on the data carriers
	{
		std::vector<uint8_t> m = mls_bits(0b100101010001, mi.cols);
		const float fac = std::sqrt(float(sym_len) / float(mi.cols));
		for (int c = 0; c < mi.cols; ++c) values[2 * c] = fac * float(1 - 2 * (int)m[c]);
		spec[0] = TxCarriers{offset - mi.cols / 2, 1, mi.cols};
	}
	// Schmidl-Cox: every other carrier, reference carrier then MLS0 differentially along frequency
	{
		std::vector<uint8_t> m = mls_bits(0b10001001, 127);
		float *v = values + 2 * 512;
		float run = std::sqrt(float(2 * sym_len) / 127.f);
		v[0] = run;
		for (int i = 0; i < 127; ++i) { run *= float(1 - 2 * (int)m[i]); v[2 * (i + 1)] = run; }
		spec[1] = TxCarriers{offset - 127 + 1 - 2, 2, 128};
	}
	// metadata: 55 bits + CRC-16, BCH(255,71) parity, differential along frequency, then scrambled by MLS1
	{
		const uint64_t md = ((uint64_t)call_sign << 8) | (uint64_t)mode;
		uint8_t bits[71];
		for (int i = 0; i < 55; ++i) bits[i] = (md >> i) & 1;
		const uint16_t cs = crc16_u64(md << 9);
		for (int i = 0; i < 16; ++i) bits[55 + i] = (cs >> i) & 1;
		std::vector<uint32_t> rows = bch_generator_rows();
		uint32_t cw[8] = {};
		for (int i = 0; i < 71; ++i)
			if (bits[i]) for (int w = 0; w < 8; ++w) cw[w] ^= rows[(size_t)i * 8 + w];
		std::vector<uint8_t> m = mls_bits(0b100101011, 255);
		float *v = values + 2 * 2 * 512;
		float run = std::sqrt(float(sym_len) / 255.f);
		v[0] = run;
		for (int i = 0; i < 255; ++i) {
			const int bit = (cw[i / 32] >> (i % 32)) & 1;
			run *= float(1 - 2 * bit);
			v[2 * (i + 1)] = run * float(1 - 2 * (int)m[i]);
		}
		spec[2] = TxCarriers{offset - 255 / 2 - 1, 1, 256};
	}
}

} // namespace ofdmrx
