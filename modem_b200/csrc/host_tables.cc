// modem_b200/csrc/host_tables.cc — see host_tables.h.  Plain C++ (no CUDA) so the host logic is unit-testable on CPU.
#include "host_tables.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ofdmrx {

// Polar code construction of freezer.cc:14-32: evolve the erasure probability of a BEC through the polar
// transform in long double (left child 2p-p^2, right child p^2, natural index order) starting from
// exp(-10^((design_SNR+1.59175)/10)); the K + 2^M - N most reliable positions carry information.
static void evolve(std::vector<long double> &prob, long double pe, int i, int h)
{
	if (!h) { prob[i] = pe; return; }
	evolve(prob, pe * (2 - pe), i, h / 2);
	evolve(prob, pe * pe, i + h, h / 2);
}
std::vector<uint32_t> make_frozen(int order, int n_tx, int k_info)
{
	const int len = 1 << order;
	long double erasure = (long double)(n_tx - k_info) / n_tx;
	double design_snr = 10 * std::log10(-std::log(erasure));
	long double start = std::exp(-std::pow(10.0, (design_snr + 1.59175) / 10));
	std::vector<long double> prob(len);
	evolve(prob, start, 0, len / 2);
	const int keep = k_info + len - n_tx;
	std::vector<int> idx(len);
	for (int i = 0; i < len; ++i) idx[i] = i;
	std::nth_element(idx.begin(), idx.begin() + keep, idx.end(), [&](int a, int b) { return prob[a] < prob[b]; });
	std::vector<uint32_t> frozen(len / 32, 0);
	for (int i = keep; i < len; ++i) frozen[idx[i] / 32] |= 1u << (idx[i] % 32);
	return frozen;
}

static bool all_frozen(const std::vector<uint32_t> &fr, int index, int n)
{
	for (int w = index / 32; w < (index + n) / 32; ++w)
		if (fr[w] != 0xffffffffu) return false;
	return true;
}
static bool all_free(const std::vector<uint32_t> &fr, int index, int n)
{
	for (int w = index / 32; w < (index + n) / 32; ++w)
		if (fr[w] != 0u) return false;
	return true;
}
namespace {
struct SclGen {
	std::vector<uint32_t> ops;
	const std::vector<uint32_t> &fr;
	int max_depth;
	bool top, r1;
	// a node gets a rate-1 attempt (OP_R1) when it is all-free, sits above the 32-leaf words and is not the left child of a
	// rate-1 node (whose attempt has the same operands: same metrics, same minimum — it would fail again)
	bool tries(int level, int index, bool notry) const { return r1 && !notry && level >= 6 && level <= kSclR1MaxLevel && all_free(fr, index, 1 << level); }
	// number of F steps that can be chained below a node of `level` at `index` whose alpha has just been produced: the node must
	// be an internal node above the 32-leaf words (level >= 6), not a rate-0 node and not a node that starts with a rate-1 attempt
	int chain_len(int level, int index, int max_more, bool notry) const
	{
		int n = 0;
		while (n < max_more && level >= 6 && !all_frozen(fr, index, 1 << level)) {
			if (tries(level, index, notry)) break;
			notry = r1 && all_free(fr, index, 1 << level);
			++n; --level;
		}
		return n;
	}
	// skip_f: this node's own F step was already performed by a fused op of an ancestor (skip_f - 1 more follow)
	void gen(int level, int index, int skip_f, bool notry)
	{
		const int n = 1 << level;
		if (top && level > kSclTopLevel) { // virtual node: its children are produced from the channel values by TOP ops
			for (int half = 0; half < 2; ++half) {
				const int ci = index + half * (n / 2);
				if (level - 1 == kSclTopLevel) {
					const int more = 1; // the kernel's TOP op always produces the left child too (no rate-0 / rate-1 node up here)
					ops.push_back(scl_pack(OP_TOP, kSclTopLevel, ci, 1 + more));
					gen(kSclTopLevel, ci, more, false);
				} else {
					gen(level - 1, ci, 0, false);
				}
			}
			ops.push_back(scl_pack(OP_C, level, index));
			return;
		}
		if (all_frozen(fr, index, n)) { ops.push_back(scl_pack(OP_R0, level, index)); return; }
		if (level == 5) { ops.push_back(scl_pack(OP_WORD, level, index)); return; }
		const bool rate1 = r1 && all_free(fr, index, n);
		size_t patch = 0;
		if (tries(level, index, notry)) { // OP_R1 + the pc to continue at when the attempt succeeds
			ops.push_back(scl_pack(OP_R1, level, index));
			patch = ops.size();
			ops.push_back(0);
		}
		int pass_down = 0;
		if (skip_f > 0) pass_down = skip_f - 1;
		else {
			const int more = chain_len(level - 1, index, max_depth - 1, rate1);
			ops.push_back(scl_pack(OP_F, level, index, 1 + more));
			pass_down = more;
		}
		gen(level - 1, index, pass_down, rate1);
		const int more = chain_len(level - 1, index + n / 2, max_depth - 1, false);
		ops.push_back(scl_pack(OP_G, level, index, 1 + more));
		gen(level - 1, index + n / 2, more, false);
		ops.push_back(scl_pack(OP_C, level, index));
		if (patch) ops[patch] = (uint32_t)ops.size();
	}
};
} // namespace
std::vector<uint32_t> make_scl_schedule(const std::vector<uint32_t> &frozen, int order, int max_depth, bool top_ops, bool r1_ops)
{
	max_depth = std::max(1, std::min(max_depth, kSclMaxFuse));
	// TOP ops need no rate-0 node at or above the top level (true for both tables of the reference: the largest are R0-2048 and
	// R1-2048; rate-1 attempts stop at level kSclR1MaxLevel < kSclTopLevel)
	for (int i = 0; top_ops && i < (1 << order); i += 1 << kSclTopLevel)
		if (all_frozen(frozen, i, 1 << kSclTopLevel)) top_ops = false;
	SclGen g{{}, frozen, max_depth, top_ops && order > kSclTopLevel, r1_ops};
	g.gen(order, 0, 0, false);
	g.ops.push_back(scl_pack(OP_END, 0, 0));
	return g.ops;
}

std::vector<uint8_t> mls_bits(int poly, int n)
{
	unsigned hi = poly;
	hi |= hi >> 1; hi |= hi >> 2; hi |= hi >> 4; hi |= hi >> 8; hi |= hi >> 16;
	int test = (hi ^ (hi >> 1)) >> 1, reg = 1;
	std::vector<uint8_t> out(n);
	for (int i = 0; i < n; ++i) {
		bool fb = reg & test;
		reg <<= 1;
		reg ^= fb * poly;
		out[i] = fb;
	}
	return out;
}

// g(x) = product of the 24 minimal polynomials (decode.cc:379-384); row i = [e_i | (x^(254-i) mod g)] MSB first.
std::vector<uint32_t> bch_generator_rows()
{
	static const int polys[24] = {
		0b100011101, 0b101110111, 0b111110011, 0b101101001, 0b110111101, 0b111100111, 0b100101011, 0b111010111,
		0b000010011, 0b101100101, 0b110001011, 0b101100011, 0b100011011, 0b100111111, 0b110001101, 0b100101101,
		0b101011111, 0b111111001, 0b111000011, 0b100111001, 0b110101001, 0b000011111, 0b110000111, 0b110110001};
	std::vector<uint8_t> g(1, 1);
	for (int p : polys) {
		int deg = 31 - __builtin_clz((unsigned)p);
		std::vector<uint8_t> r(g.size() + deg, 0);
		for (size_t i = 0; i < g.size(); ++i)
			if (g[i])
				for (int b = 0; b <= deg; ++b)
					if ((p >> b) & 1) r[i + b] ^= 1;
		g.swap(r);
	}
	const int NP = 184, K = kHdrK;
	std::vector<uint32_t> rows((size_t)K * 8, 0);
	for (int i = 0; i < K; ++i) {
		// remainder of x^(NP + K-1-i) modulo g, by shifting a single 1 through the divider K-i times
		std::vector<uint8_t> reg(NP, 0);
		for (int s = i; s < K; ++s) {
			uint8_t fb = (s == i ? 1 : 0) ^ reg[NP - 1];
			for (int d = NP - 1; d > 0; --d) reg[d] = reg[d - 1] ^ (fb & g[d]);
			reg[0] = fb & g[0];
		}
		uint32_t *row = &rows[(size_t)i * 8];
		row[i / 32] |= 1u << (i % 32);
		for (int j = 0; j < NP; ++j)
			if (reg[NP - 1 - j]) row[(K + j) / 32] |= 1u << ((K + j) % 32);
	}
	return rows;
}

static float bessel_i0(float x)
{
	float sum = 1, val = 1;
	for (int n = 1; n < 35; ++n) { val *= x / float(2 * n); sum += val * val; }
	return sum;
}
static float kaiser(float a, int n, int N)
{
	const float pi = 3.14159265358979323846f;
	float t = float(2 * n) / float(N - 1) - 1.f;
	return bessel_i0(pi * a * std::sqrt(1.f - t * t)) / bessel_i0(pi * a);
}
std::vector<float> hilbert_coeffs(int taps, float *reco)
{
	const float pi = 3.14159265358979323846f;
	*reco = kaiser(2.f, (taps - 1) / 2, taps);
	std::vector<float> im;
	for (int i = 0; i < (taps - 1) / 4; ++i)
		im.push_back(kaiser(2.f, (2 * i + 1) + (taps - 1) / 2, taps) * 2.f / (float(2 * i + 1) * pi));
	return im;
}

std::vector<float> twiddles(int n, int sign)
{
	std::vector<float> t((size_t)2 * n);
	for (int k = 0; k < n; ++k) {
		double a = sign * 2.0 * M_PI * (double)k / (double)n;
		t[2 * k] = (float)std::cos(a);
		t[2 * k + 1] = (float)std::sin(a);
	}
	return t;
}

std::vector<float> mls0_kernel(int half)
{
	const int kHalf = half; // shadows the 8 kHz constant
	// template: +-1 at bins -63..63 of 640 (decode.cc:236-244); spectrum by direct DFT in double, conj, / 640
	std::vector<double> seq(kHalf, 0.0);
	std::vector<uint8_t> m = mls_bits(0b10001001, 127);
	const int off = (-127 + 1) / 2;
	for (int i = 0; i < 127; ++i) seq[(i + off + kHalf) % kHalf] = 1 - 2 * (int)m[i];
	std::vector<float> k((size_t)2 * kHalf);
	for (int b = 0; b < kHalf; ++b) {
		double re = 0, im = 0;
		for (int n = 0; n < kHalf; ++n) {
			if (seq[n] == 0.0) continue;
			double a = -2.0 * M_PI * (double)((long long)b * n % kHalf) / (double)kHalf;
			re += seq[n] * std::cos(a);
			im += seq[n] * std::sin(a);
		}
		k[2 * b] = (float)(re / kHalf);
		k[2 * b + 1] = (float)(-im / kHalf);
	}
	return k;
}

void crc32_table(uint32_t poly, uint32_t *lut)
{
	for (uint32_t j = 0; j < 256; ++j) {
		uint32_t t = j;
		for (int i = 0; i < 8; ++i) t = (t >> 1) ^ ((t & 1) * poly);
		lut[j] = t;
	}
}
std::vector<uint32_t> crc32_pieces(const std::vector<uint32_t> &frozen, int crc_bits)
{
	const int words = (int)frozen.size();
	std::vector<int> before(words + 1, 0);
	for (int w = 0; w < words; ++w) before[w + 1] = before[w] + 32 - __builtin_popcount(frozen[w]);
	std::vector<uint32_t> out(16 + 256, 0);
	for (int j = 0; j <= 8; ++j) { // piece j starts at the first word holding message bit j * crc_bits / 8 or a later one
		const int target = (int)((long long)j * crc_bits / 8);
		int w = 0;
		while (w < words && before[w] < target) ++w;
		out[j] = (uint32_t)w;
	}
	for (int j = 0; j < 8; ++j) {
		const int end_bits = std::min(crc_bits, before[out[j + 1]]); // message bits up to the end of piece j
		const int follow = crc_bits - std::min(crc_bits, std::max(end_bits, 0));
		for (int c = 0; c < 32; ++c) {
			uint32_t reg = 1u << c;
			for (int n = 0; n < follow; ++n) reg = (reg >> 1) ^ ((reg & 1u) ? 0xD419CC15u : 0u);
			out[16 + 32 * j + c] = reg;
		}
	}
	return out;
}
uint16_t crc16_u64(uint64_t v)
{
	uint16_t crc = 0;
	for (int b = 0; b < 64; ++b) {
		uint16_t bit = (v >> b) & 1;
		crc = (crc >> 1) ^ (((crc ^ bit) & 1) * 0xA8F4);
	}
	return crc;
}
void base37_decode(char *str, long long val, int len)
{
	static const char tab[] = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ";
	for (int i = len - 1; i >= 0; --i, val /= 37) str[i] = tab[val % 37];
}

} // namespace ofdmrx
