// oracle/decode_ref.cc — TEST INFRASTRUCTURE ONLY.  CPU receiver with the argv contract of the reference
// (/root/reference/decode.cc:559-620): decode OUTPUT INPUT [SKIP]; same stderr lines; always writes 5380 bytes.
// Extra env knobs (oracle only): REF_LIST=4|8, REF_OSD_LITERAL=1, REF_R0MAX=n.
#include "ref_modem.hh"
#include <fstream>
#include <iostream>
#include <iterator>
using namespace ref;

int main(int argc, char **argv)
{
	if (argc < 3 || argc > 4) {
		std::cerr << "usage: " << argv[0] << " OUTPUT INPUT [SKIP]" << std::endl;
		return 1;
	}
	std::string output_name = argv[1], input_name = argv[2];
	if (output_name == "-") output_name = "/dev/stdout";
	if (input_name == "-") input_name = "/dev/stdin";
	std::ifstream in(input_name, std::ios::binary);
	std::vector<uint8_t> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	WavData w;
	if (!wav_parse(raw.data(), raw.size(), w)) { std::cerr << "Couldn't parse WAV input." << std::endl; return 1; }
	if (w.channels < 1 || w.channels > 2) {
		std::cerr << "Only real or analytic signal (one or two channels) supported." << std::endl;
		return 1;
	}
	int skip = argc > 3 ? std::atoi(argv[3]) : 0;
	if (w.rate != 8000 && w.rate != 16000 && w.rate != 44100 && w.rate != 48000) { std::cerr << "Unsupported sample rate." << std::endl; return 1; }
	RxOptions opt;
	if (const char *e = std::getenv("REF_LIST")) opt.list_size = std::atoi(e);
	if (const char *e = std::getenv("REF_OSD_LITERAL")) opt.osd_literal = std::atoi(e) != 0;
	if (const char *e = std::getenv("REF_R0MAX")) opt.r0_max = std::atoi(e);
	Receiver rx(w.rate);
	uint8_t out[kDataBytes];
	std::memset(out, 0, sizeof(out));
	int st = rx.run(out, w.samples.data(), w.frames(), w.channels, skip, opt);
	const Taps &t = rx.taps;
	std::cerr << rx.header_log; // per consumed detection: symbol pos, coarse cfo, header outcome (decode.cc:400-446)
	if (st == ST_OK || st == ST_PAYLOAD_CRC) {
		std::cerr << "demod ";
		for (size_t j = 0; j < t.slope.size(); ++j) std::cerr << ".";
		std::cerr << " done" << std::endl;
		{ // decode.cc:480,491-503; sfo_rad is an uninitialised member there (decode.cc:210): zero, as a fresh heap gives
			float sum_slope = 0, sum_yint = 0;
			for (size_t j = 0; j < t.slope.size(); ++j) { sum_slope += t.slope[j]; sum_yint += t.yint[j]; }
			const int symbol_len = 1280 * w.rate / 8000, guard_len = symbol_len / 8, rows = (int)t.slope.size();
			const float sfo_rad = 0.f - (sum_slope / rows) * symbol_len / float(symbol_len + guard_len);
			const float cfo_rad = t.cfo_rad + (sum_yint / rows) / (symbol_len + guard_len);
			std::cerr << "coarse sfo: " << 1000000 * sfo_rad / kTwoPi << " ppm" << std::endl;
			std::cerr << "finer cfo: " << cfo_rad * (w.rate / kTwoPi) << " Hz " << std::endl;
		}
		std::cerr << "Es/N0 (dB):";
		for (float p : t.precision) std::cerr << " " << 10.f * std::log10(p);
		std::cerr << std::endl;
		if (st == ST_OK) std::cerr << "bit flips: " << t.flips << std::endl;
		else std::cerr << "payload decoding error." << std::endl;
	}
	descramble(out);
	std::ofstream of(output_name, std::ios::binary | std::ios::trunc);
	if (of.bad()) { std::cerr << "Couldn't open file \"" << output_name << "\" for writing." << std::endl; return 1; }
	of.write(reinterpret_cast<const char *>(out), kDataBytes);
	return 0;
}
