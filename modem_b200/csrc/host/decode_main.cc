// modem_b200/csrc/host/decode_main.cc — `decode OUTPUT INPUT [SKIP]`: the reference receiver's command line
// (/root/reference/decode.cc:559-620) as a thin C++ host driver over libofdmrx (include/ofdmrx.h).
// Same argv rules, "-" for stdin/stdout, same stderr lines (tests/test_cli_host.py diffs them against the reference's own
// main() on the CPU through a mock of the library), always writes 5380 bytes and exits 0 once the WAV opened.
// Extension: `decode --batch OUTPUT INPUT [SKIP]` treats INPUT as N back-to-back windows of 95200 frames (one
// single-frame recording each) and writes N x 5380 bytes.  There is no CPU fallback: without a B200 it exits 1.
#include "ofdmrx.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

namespace {

struct Wav {
	int rate = 0, channels = 0, bits = 0;
	std::vector<int16_t> pcm; // interleaved 16-bit samples (the device scales them like ReadWAV: v / 32767)
	std::vector<float> flt;   // any other depth: interleaved floats, scaled exactly like DSP::ReadWAV<float> (decode.cc:576)
	size_t count() const { return bits == 16 ? pcm.size() : flt.size(); }
};

bool parse_wav(const std::vector<uint8_t> &d, Wav &w)
{
	auto rd32 = [&](size_t o) { return (uint32_t)d[o] | (uint32_t)d[o + 1] << 8 | (uint32_t)d[o + 2] << 16 | (uint32_t)d[o + 3] << 24; };
	auto rd16 = [&](size_t o) { return (uint32_t)d[o] | (uint32_t)d[o + 1] << 8; };
	if (d.size() < 44 || std::memcmp(&d[0], "RIFF", 4) || std::memcmp(&d[8], "WAVE", 4)) return false;
	size_t o = 12;
	bool fmt = false;
	while (o + 8 <= d.size()) {
		uint32_t sz = rd32(o + 4);
		if (!std::memcmp(&d[o], "fmt ", 4)) {
			if (rd16(o + 8) != 1) return false;
			w.channels = rd16(o + 10); w.rate = rd32(o + 12); w.bits = rd16(o + 22);
			fmt = true;
		} else if (!std::memcmp(&d[o], "data", 4)) {
			if (!fmt || w.channels < 1) return false;
			size_t avail = d.size() - (o + 8), n = (sz == 0 || sz == 0xffffffffu) ? avail : std::min<size_t>(sz, avail);
			int bytes = w.bits / 8;
			if (bytes < 1 || bytes > 4) return false;
			size_t cnt = n / bytes / w.channels * w.channels;
			if (w.bits == 16) w.pcm.resize(cnt);
			else w.flt.resize(cnt);
			const float fac = (float)((1u << (w.bits - 1)) - 1u); // 2^(bits-1) - 1
			for (size_t i = 0; i < cnt; ++i) {
				const uint8_t *p = &d[o + 8 + i * bytes];
				int32_t v = 0;
				for (int b = 0; b < bytes; ++b) v |= (int32_t)p[b] << (8 * b);
				if (bytes > 1 && bytes < 4 && (v & (1 << (8 * bytes - 1)))) v |= ~((1 << (8 * bytes)) - 1);
				if (bytes == 1) v -= 128;
				if (w.bits == 16) w.pcm[i] = (int16_t)v;
				else w.flt[i] = (float)v / fac;
			}
			return true;
		}
		o += 8 + sz + (sz & 1);
	}
	return false;
}

void base37(char *str, long long val, int len)
{
	for (int i = len - 1; i >= 0; --i, val /= 37) str[i] = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"[val % 37];
}

} // namespace

int main(int argc, char **argv)
{
	// extension: --batch[=STRIDE] decodes N consecutive windows of STRIDE sample frames (default 95200: a single mode-6 frame
	// recording at 8 kHz) and writes N x 5380 bytes
	bool batch = argc > 1 && !std::strncmp(argv[1], "--batch", 7) && (argv[1][7] == 0 || argv[1][7] == '=');
	int64_t batch_stride = 95200;
	if (batch) {
		if (argv[1][7] == '=') batch_stride = std::atoll(argv[1] + 8);
		if (batch_stride < 1) { std::cerr << "usage: " << argv[0] << " [--batch[=STRIDE]] OUTPUT INPUT [SKIP]" << std::endl; return 1; }
		--argc; ++argv;
	}
	if (argc < 3 || argc > 4) {
		std::cerr << "usage: " << argv[0] << " OUTPUT INPUT [SKIP]" << std::endl;
		return 1;
	}
	std::string output_name = argv[1], input_name = argv[2];
	if (output_name == "-") output_name = "/dev/stdout";
	if (input_name == "-") input_name = "/dev/stdin";
	std::ifstream in(input_name, std::ios::binary);
	std::vector<uint8_t> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	Wav w;
	if (!parse_wav(raw, w)) { std::cerr << "Couldn't open file \"" << input_name << "\" for reading." << std::endl; return 1; }
	if (w.channels < 1 || w.channels > 2) {
		std::cerr << "Only real or analytic signal (one or two channels) supported." << std::endl;
		return 1;
	}
	int skip = argc > 3 ? std::atoi(argv[3]) : 0;
	if (skip < 0) skip = 0;
	if (w.rate != 8000 && w.rate != 16000 && w.rate != 44100 && w.rate != 48000) { // decode.cc:590-606
		std::cerr << "Unsupported sample rate." << std::endl;
		return 1;
	}
	const int64_t total = (int64_t)w.count() / w.channels;
	const int64_t stride = batch ? batch_stride : std::max<int64_t>(total, 1);
	const int n_frames = batch ? (int)(total / stride) : 1;
	if (n_frames < 1) { std::cerr << "input shorter than one window" << std::endl; return 1; }
	ofdmrx_t *h = nullptr;
	int rc = ofdmrx_create(&h, 0, w.rate, std::min(n_frames, 4096), (int)stride);
	if (rc) { std::cerr << "ofdmrx_create failed (" << rc << "): a B200 is required, there is no CPU path" << std::endl; return 1; }
	std::vector<uint8_t> out((size_t)n_frames * OFDMRX_PAYLOAD_BYTES);
	const bool is16 = w.bits == 16;
	if (is16 && w.pcm.size() < (size_t)n_frames * stride * w.channels) w.pcm.resize((size_t)n_frames * stride * w.channels, 0);
	if (!is16 && w.flt.size() < (size_t)n_frames * stride * w.channels) w.flt.resize((size_t)n_frames * stride * w.channels, 0.f);
	const void *samples = is16 ? (const void *)w.pcm.data() : (const void *)w.flt.data();
	const int format = w.channels == 1 ? (is16 ? OFDMRX_FMT_S16_MONO : OFDMRX_FMT_F32_MONO) : (is16 ? OFDMRX_FMT_S16_IQ : OFDMRX_FMT_F32_IQ);
	std::vector<int32_t> ns(n_frames, (int32_t)std::min<int64_t>(stride, total));
	// The reference prints the header diagnostics of EVERY detection it consumes on the way to the SKIP-th one
	// (decode.cc:390-448).  The library reports the last consumed detection of a call, so the driver asks for skip = 0, 1, ..
	// SKIP in turn (the walk is deterministic: call k ends on detection k) and prints each new detection once.
	std::vector<std::vector<ofdmrx_frame_status>> walk(skip + 1, std::vector<ofdmrx_frame_status>(n_frames));
	for (int k = 0; k <= skip && !rc; ++k) {
		rc = ofdmrx_decode_batch(h, samples, OFDMRX_MEM_HOST, format, n_frames, stride, ns.data(), k, out.data(), walk[k].data(), nullptr);
		bool any = false;
		for (int i = 0; i < n_frames; ++i) any |= walk[k][i].detections == k + 1;
		if (!any) break; // no window holds a detection k: larger skips end the same way (a huge SKIP costs one extra call)
	}
	if (rc) { ofdmrx_destroy(h); std::cerr << "ofdmrx_decode_batch failed (" << rc << ")" << std::endl; return 1; }
	for (int i = 0; i < n_frames; ++i)
		if (walk[0][i].det_overflow) // (not a reference message: the reference's walk is unbounded, decode.cc:390-448)
			std::cerr << "warning: window " << i << " holds more correlator detections than the library's list; later ones were not examined" << std::endl;
	const int sym_len = 1280 * w.rate / 8000, pitch = sym_len + sym_len / 8;
	for (int i = 0; i < n_frames; ++i) {
		if (batch) std::cerr << "window " << i << ":" << std::endl;
		const ofdmrx_frame_status *last = nullptr;
		for (int k = 0; k <= skip; ++k) {
			const ofdmrx_frame_status &s = walk[k][i];
			if (s.detections != k + 1) break; // the stream ended before detection k
			last = &s;
			std::cerr << "symbol pos: " << s.symbol_pos << std::endl;
			std::cerr << "coarse cfo: " << s.cfo_rad * ((float)w.rate / 6.28318530717958647692f) << " Hz " << std::endl;
			switch (s.status) {
			case OFDMRX_ST_OSD_FAIL: std::cerr << "OSD error." << std::endl; break;
			case OFDMRX_ST_HDR_CRC: std::cerr << "header CRC error." << std::endl; break;
			case OFDMRX_ST_BAD_MODE: case OFDMRX_ST_UNSUPPORTED_MODE: std::cerr << "operation mode " << s.mode << " unsupported." << std::endl; break;
			case OFDMRX_ST_BAD_CALL: std::cerr << "oper mode: " << s.mode << std::endl << "call sign unsupported." << std::endl; break;
			default: {
				char cs[10];
				base37(cs, (long long)((((uint64_t)s.md_hi << 32) | s.md_lo) >> 8), 9);
				cs[9] = 0;
				std::cerr << "oper mode: " << s.mode << std::endl << "call sign: " << cs << std::endl;
			}
			}
		}
		if (!last || last->detections != skip + 1 || (last->status != OFDMRX_ST_OK && last->status != OFDMRX_ST_PAYLOAD_CRC)) continue;
		const ofdmrx_frame_status &s = *last;
		static const int rows_of_mode[8] = {50, 54, 81, 90, 42, 56, 84, 126}; // cons_bits / mod_bits / cons_cols, decode.cc:302-374
		const int rows = rows_of_mode[s.mode - 6];
		std::cerr << "demod ";
		for (int j = 0; j < rows; ++j) std::cerr << ".";
		std::cerr << " done" << std::endl;
		// per-row phase line and Es/N0 (decode.cc:479-523) from the stage taps of the chunk this window was decoded in
		std::vector<float> ts((size_t)ofdmrx_tap_elems(h, OFDMRX_TAP_TS));
		if (n_frames <= 4096 && ofdmrx_get_taps(h, OFDMRX_TAP_TS, i, 1, ts.data(), ts.size() * sizeof(float)) == 0) {
			float sum_slope = 0, sum_yint = 0;
			for (int j = 0; j < rows; ++j) { sum_slope += ts[3 * j]; sum_yint += ts[3 * j + 1]; }
			// sfo_rad starts from an uninitialised member in the reference (decode.cc:210,500); zero is what a fresh heap gives
			const float sfo_rad = 0.f - (sum_slope / rows) * sym_len / float(pitch);
			const float cfo_rad = s.cfo_rad + (sum_yint / rows) / pitch;
			std::cerr << "coarse sfo: " << 1000000 * sfo_rad / 6.28318530717958647692f << " ppm" << std::endl;
			std::cerr << "finer cfo: " << cfo_rad * ((float)w.rate / 6.28318530717958647692f) << " Hz " << std::endl;
			std::cerr << "Es/N0 (dB):";
			for (int j = 0; j < rows; ++j) std::cerr << " " << 10.f * std::log10(ts[3 * j + 2]);
			std::cerr << std::endl;
		}
		if (s.status == OFDMRX_ST_OK) std::cerr << "bit flips: " << s.flips << std::endl;
		else std::cerr << "payload decoding error." << std::endl;
	}
	ofdmrx_destroy(h);
	std::ofstream of(output_name, std::ios::binary | std::ios::trunc);
	if (of.bad()) { std::cerr << "Couldn't open file \"" << output_name << "\" for writing." << std::endl; return 1; }
	of.write(reinterpret_cast<const char *>(out.data()), out.size());
	return 0;
}
