// oracle/shim/shim_code.hh — TEST INFRASTRUCTURE ONLY.
//
// API-compatible stand-ins for the headers of aicodix/code that /root/reference/decode.cc:29-35, encode.cc:11-25 and
// freezer.cc:12 include and that are ABSENT from this box (see shim_dsp.hh for the purpose and the caveat: the arithmetic
// behind each class is the oracle's restatement in ref_code.hh / ref_freezer.hh, wrapped one to one).
#pragma once
#include "../ref_code.hh"
#include "../ref_freezer.hh"
#include <cstdlib>
#include <initializer_list>
#include <vector>

// simd.hh: the reference only uses the lane array and the lane count (decode.cc:165-169,534-537,551)
template <typename TYPE, int WIDTH>
struct SIMD {
	static const int SIZE = WIDTH;
	typedef TYPE value_type;
	TYPE v[WIDTH];
};

namespace CODE {

// bitman.hh
static inline bool get_be_bit(const uint8_t *b, int i) { return ref::get_be_bit(b, i); }
static inline bool get_le_bit(const uint8_t *b, int i) { return ref::get_le_bit(b, i); }
static inline void set_be_bit(uint8_t *b, int i, bool v) { ref::set_be_bit(b, i, v); }
static inline void set_le_bit(uint8_t *b, int i, bool v) { ref::set_le_bit(b, i, v); }

// xorshift.hh, mls.hh
struct Xorshift32 {
	ref::Xorshift32 x;
	uint32_t operator()() { return x(); }
};
class MLS {
	ref::MLS m_;
public:
	explicit MLS(int poly, int reg = 1) : m_(poly, reg) {}
	bool operator()() { return m_(); }
};

// crc.hh: reflected, init 0, no final xor; wide integers are fed little-endian bytewise
template <typename TYPE>
class CRC {
	ref::CRC<TYPE> c_;
public:
	explicit CRC(TYPE poly, TYPE crc = 0) : c_(poly, crc) {}
	void reset(TYPE v = 0) { c_.reset(v); }
	TYPE operator()() { return c_(); }
	TYPE operator()(bool d) { return c_.bit(d); }
	TYPE operator()(uint8_t d) { return c_.byte(d); }
	TYPE operator()(uint64_t d) { return c_.u64(d); }
};

// bose_chaudhuri_hocquenghem_encoder.hh: generator = product of the minimal polynomials handed in; systematic,
// parity(x) = data(x) x^NP mod g(x), first bit = highest power, big-endian bit packing
template <int N, int K>
struct BchGen {
	static const int NP = N - K;
	uint8_t gen[NP + 1];
	explicit BchGen(std::initializer_list<int> minimal_polynomials)
	{
		std::vector<uint8_t> g(1, 1);
		for (int p : minimal_polynomials) {
			int deg = 31 - __builtin_clz((unsigned)p);
			std::vector<uint8_t> r(g.size() + deg, 0);
			for (size_t i = 0; i < g.size(); ++i)
				if (g[i])
					for (int b = 0; b <= deg; ++b)
						if ((p >> b) & 1) r[i + b] ^= 1;
			g.swap(r);
		}
		if ((int)g.size() != NP + 1) { std::fprintf(stderr, "BCH: generator degree %d, expected %d\n", (int)g.size() - 1, NP); std::abort(); }
		for (int d = 0; d <= NP; ++d) gen[d] = g[d];
	}
	void encode_bits(const uint8_t *data_bits, uint8_t *parity_bits) const
	{
		uint8_t reg[NP];
		std::memset(reg, 0, sizeof(reg));
		for (int i = 0; i < K; ++i) {
			uint8_t fb = data_bits[i] ^ reg[NP - 1];
			for (int d = NP - 1; d > 0; --d) reg[d] = reg[d - 1] ^ (fb & gen[d]);
			reg[0] = fb & gen[0];
		}
		for (int j = 0; j < NP; ++j) parity_bits[j] = reg[NP - 1 - j];
	}
};
template <int N, int K>
class BoseChaudhuriHocquenghemEncoder {
	BchGen<N, K> g_;
public:
	BoseChaudhuriHocquenghemEncoder(std::initializer_list<int> minimal_polynomials) : g_(minimal_polynomials) {}
	void operator()(const uint8_t *data, uint8_t *parity, int data_len = K)
	{
		uint8_t d[K], p[N - K];
		for (int i = 0; i < K; ++i) d[i] = i < data_len ? get_be_bit(data, i) : 0;
		g_.encode_bits(d, p);
		for (int j = 0; j < N - K; ++j) set_be_bit(parity, j, p[j]);
	}
};
template <int N, int K>
struct BoseChaudhuriHocquenghemGenerator {
	// rows [e_i | parity(e_i)], genmat[N * i + j] in {0, 1}
	static void matrix(int8_t *genmat, bool systematic, std::initializer_list<int> minimal_polynomials)
	{
		if (!systematic) { std::fprintf(stderr, "BCH generator: only the systematic form is restated\n"); std::abort(); }
		BchGen<N, K> g(minimal_polynomials);
		for (int i = 0; i < K; ++i) {
			uint8_t d[K], p[N - K];
			std::memset(d, 0, K);
			d[i] = 1;
			g.encode_bits(d, p);
			for (int j = 0; j < K; ++j) genmat[N * i + j] = d[j];
			for (int j = 0; j < N - K; ++j) genmat[N * i + K + j] = p[j];
		}
	}
};

// osd.hh: order-O reprocessing; REF_OSD_LITERAL=1 walks all candidates like the reference, the default is the oracle's
// exact-equivalent branch and bound (tests/test_oracle_kat.py::test_osd_pruned_equals_literal)
template <int N, int K, int O>
class OrderedStatisticsDecoder {
	static_assert(N == 255 && K == 71 && O == 4, "only the instance decode.cc:199 uses is restated");
	ref::OSD255_71 osd_;
public:
	bool operator()(uint8_t *hard, const int8_t *soft, const int8_t *genmat)
	{
		const char *lit = std::getenv("REF_OSD_LITERAL");
		return lit && lit[0] == '1' ? osd_.decode_full(hard, soft, genmat) : osd_.decode_pruned(hard, soft, genmat);
	}
};

// polar_helper.hh
template <typename TYPE>
struct PolarHelper {
	static TYPE quant(double v) { return (TYPE)v; }
};

// polar_encoder.hh: non-systematic transform on +-1 values, natural order (product = XOR); lanes are independent
template <typename TYPE>
struct PolarLanes {
	static void mul(TYPE &a, const TYPE &b) { a *= b; }
	static void one(TYPE &a) { a = 1; }
};
template <typename T, int W>
struct PolarLanes<SIMD<T, W>> {
	static void mul(SIMD<T, W> &a, const SIMD<T, W> &b) { for (int k = 0; k < W; ++k) a.v[k] *= b.v[k]; }
	static void one(SIMD<T, W> &a) { for (int k = 0; k < W; ++k) a.v[k] = 1; }
};
template <typename TYPE>
static inline void polar_butterflies(TYPE *c, int n)
{
	for (int h = 1; h < n; h *= 2)
		for (int i = 0; i < n; i += 2 * h)
			for (int j = i; j < i + h; ++j) PolarLanes<TYPE>::mul(c[j], c[j + h]);
}
template <typename TYPE>
struct PolarEncoder {
	void operator()(TYPE *codeword, const TYPE *message, const uint32_t *frozen, int level)
	{
		int n = 1 << level;
		for (int i = 0, j = 0; i < n; ++i)
			if ((frozen[i / 32] >> (i % 32)) & 1) PolarLanes<TYPE>::one(codeword[i]);
			else codeword[i] = message[j++];
		polar_butterflies(codeword, n);
	}
};
template <typename TYPE>
struct PolarSysEnc {
	void operator()(TYPE *codeword, const TYPE *message, const uint32_t *frozen, int level)
	{
		int n = 1 << level;
		PolarEncoder<TYPE>()(codeword, message, frozen, level);
		for (int i = 0; i < n; ++i)
			if ((frozen[i / 32] >> (i % 32)) & 1) PolarLanes<TYPE>::one(codeword[i]);
		polar_butterflies(codeword, n);
	}
};

// polar_list_decoder.hh: message[j].v[k] = u (+-1) of the j-th free index on list lane k, lane 0 = smallest metric
template <typename TYPE, int MAX_M>
class PolarListDecoder {
	static const int L = TYPE::SIZE;
public:
	void operator()(int64_t *, TYPE *message, const typename TYPE::value_type *codeword, const uint32_t *frozen, int level)
	{
		ref::PolarListDecoder<L> dec(level, frozen);
		if (const char *r0 = std::getenv("REF_R0MAX")) dec.r0_max = std::atoi(r0);
		std::vector<std::vector<uint8_t>> x;
		float metrics[L];
		dec.decode(codeword, x, metrics);
		int n = 1 << level;
		for (int k = 0; k < L; ++k) ref::polar_transform(x[k].data(), n); // partial sums back to the u domain
		for (int i = 0, j = 0; i < n; ++i)
			if (!((frozen[i / 32] >> (i % 32)) & 1)) {
				for (int k = 0; k < L; ++k) message[j].v[k] = x[k][i] ? -1.f : 1.f;
				++j;
			}
	}
};

// polar_freezer.hh: binary-erasure-channel evolution, the K most reliable indices stay free
template <int MAX_M>
struct PolarCodeConst0 {
	void operator()(uint32_t *frozen_bits, int level, int K, long double probability)
	{
		int len = 1 << level;
		std::vector<long double> prob(len);
		ref::bec_evolve(prob, probability, 0, len / 2);
		std::vector<int> idx(len);
		for (int i = 0; i < len; ++i) idx[i] = i;
		std::nth_element(idx.begin(), idx.begin() + K, idx.end(), [&](int a, int b) { return prob[a] < prob[b]; });
		for (int i = 0; i < len / 32; ++i) frozen_bits[i] = 0;
		for (int i = K; i < len; ++i) frozen_bits[idx[i] / 32] |= 1u << (idx[i] % 32);
	}
};

} // namespace CODE
