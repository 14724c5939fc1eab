// oracle/encode_ref.cc — TEST INFRASTRUCTURE ONLY.  Stimulus generator with the argv contract of the reference
// transmitter (/root/reference/encode.cc:337-445): encode OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT..
#include "ref_modem.hh"
#include <fstream>
#include <iostream>
using namespace ref;

int main(int argc, char **argv)
{
	if (argc < 9) {
		std::cerr << "usage: " << argv[0] << " OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT.." << std::endl;
		return 1;
	}
	std::string output_name = argv[1];
	if (output_name == "-") output_name = "/dev/stdout";
	int rate = std::atoi(argv[2]), bits = std::atoi(argv[3]), chan = std::atoi(argv[4]);
	int freq_off = std::atoi(argv[5]), mode = std::atoi(argv[6]);
	if (mode < 6 || mode > 13) { std::cerr << "Unsupported operation mode." << std::endl; return 1; }
	long long call_sign = base37_encode(argv[7]);
	if (call_sign <= 0 || call_sign >= kCallSignLimit) { std::cerr << "Unsupported call sign." << std::endl; return 1; }
	int bw = band_width(mode);
	if ((chan == 1 && freq_off < bw / 2) || freq_off < bw / 2 - rate / 2 || freq_off > rate / 2 - bw / 2) {
		std::cerr << "Unsupported frequency offset." << std::endl;
		return 1;
	}
	if (freq_off % 50) { std::cerr << "Frequency offset must be divisible by 50." << std::endl; return 1; }
	if (rate != 8000 && rate != 16000 && rate != 44100 && rate != 48000) { std::cerr << "Unsupported sample rate." << std::endl; return 1; }
	int count = argc - 8;
	std::vector<uint8_t> data((size_t)count * kDataBytes);
	for (int j = 0; j < count; ++j) {
		std::string name = argv[j + 8];
		if (argc == 9 && name == "-") name = "/dev/stdin";
		std::ifstream in(name, std::ios::binary);
		if (in.bad()) { std::cerr << "Couldn't open file \"" << name << "\" for reading." << std::endl; return 1; }
		for (int i = 0; i < kDataBytes; ++i) data[(size_t)j * kDataBytes + i] = (uint8_t)in.get(); // short file => 0xFF (get() == -1)
	}
	Transmitter tx(rate);
	std::vector<cf> s;
	tx.encode(s, data.data(), count, freq_off, call_sign, mode);
	std::vector<float> inter(s.size() * chan);
	for (size_t n = 0; n < s.size(); ++n) {
		inter[n * chan] = s[n].re;
		if (chan == 2) inter[n * chan + 1] = s[n].im;
	}
	std::vector<uint8_t> wav = wav_serialize(rate, bits, chan, inter);
	std::ofstream out(output_name, std::ios::binary | std::ios::trunc);
	out.write(reinterpret_cast<const char *>(wav.data()), wav.size());
	return 0;
}
