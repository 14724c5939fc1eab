// oracle/ref_code.hh — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Restatement of the coding primitives the reference pulls from the ABSENT, UNPINNED header
// library aicodix/code (`-I../code`, /root/reference/Makefile:2): MLS, CRC, xorshift, bit helpers,
// BCH(255,71) encoder + generator matrix, ordered-statistics decoder, polar systematic encoder,
// polar successive-cancellation list decoder; plus the reference's own psk.hh 8PSK rules.
// (recalled) = semantics remembered from the public sources, not read here: PARITY UNPINNED.
// Known-answer values from SURVEY.md Appendix B are asserted in tests/test_oracle_kat.py.
#pragma once
#include "ref_dsp.hh"
#include <cassert>

namespace ref {

// ---------------------------------------------------------------- bit helpers (bitman.hh, recalled)
static inline bool get_be_bit(const uint8_t *b, int i) { return (b[i / 8] >> (7 - i % 8)) & 1; }
static inline bool get_le_bit(const uint8_t *b, int i) { return (b[i / 8] >> (i % 8)) & 1; }
static inline void set_be_bit(uint8_t *b, int i, bool v) { b[i / 8] = (b[i / 8] & ~(1 << (7 - i % 8))) | ((int)v << (7 - i % 8)); }
static inline void set_le_bit(uint8_t *b, int i, bool v) { b[i / 8] = (b[i / 8] & ~(1 << (i % 8))) | ((int)v << (i % 8)); }

// ---------------------------------------------------------------- MLS (mls.hh, recalled; decode.cc:238,407, encode.cc:134,144,165)
class MLS {
	int poly_, test_, reg_;
	static int hibit(unsigned n) { n |= n >> 1; n |= n >> 2; n |= n >> 4; n |= n >> 8; n |= n >> 16; return n ^ (n >> 1); }
public:
	explicit MLS(int poly, int reg = 1) : poly_(poly), test_(hibit(poly) >> 1), reg_(reg) {}
	bool operator()()
	{
		bool fb = reg_ & test_;
		reg_ <<= 1;
		reg_ ^= fb * poly_;
		return fb;
	}
};

// ---------------------------------------------------------------- CRC (crc.hh, recalled; decode.cc:197-198,428-429,534-537)
// reflected, init 0, no final xor; wide integers are fed little-endian bytewise.
template <typename T>
class CRC {
	T lut_[256];
	T poly_, crc_;
	T step(T prev, bool data) const { T tmp = prev ^ (T)data; return (prev >> 1) ^ ((tmp & 1) * poly_); }
public:
	explicit CRC(T poly, T crc = 0) : poly_(poly), crc_(crc)
	{
		for (int j = 0; j < 256; ++j) {
			T tmp = j;
			for (int i = 8; i; --i) tmp = step(tmp, 0);
			lut_[j] = tmp;
		}
	}
	void reset(T v = 0) { crc_ = v; }
	T operator()() const { return crc_; }
	T bit(bool d) { return crc_ = step(crc_, d); }
	T byte(uint8_t d) { T tmp = crc_ ^ d; return crc_ = (crc_ >> 8) ^ lut_[tmp & 255]; }
	T u64(uint64_t d) { for (int i = 0; i < 8; ++i) byte((d >> (8 * i)) & 255); return crc_; }
};

// ---------------------------------------------------------------- Xorshift32 (xorshift.hh, recalled; decode.cc:613-615, encode.cc:417-419)
struct Xorshift32 {
	uint32_t y = 2463534242u;
	uint32_t operator()() { y ^= y << 13; y ^= y >> 17; y ^= y << 5; return y; }
};

// ---------------------------------------------------------------- base-37 call signs (decode.cc:155-159, encode.cc:320-335)
static inline long long base37_encode(const char *str)
{
	long long acc = 0;
	for (; *str; ++str) {
		char c = *str;
		acc *= 37;
		if (c >= '0' && c <= '9') acc += c - '0' + 1;
		else if (c >= 'a' && c <= 'z') acc += c - 'a' + 11;
		else if (c >= 'A' && c <= 'Z') acc += c - 'A' + 11;
		else if (c != ' ') return -1;
	}
	return acc;
}
static inline void base37_decode(char *str, long long val, int len)
{
	static const char tab[] = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ";
	for (int i = len - 1; i >= 0; --i, val /= 37) str[i] = tab[val % 37];
}

// ---------------------------------------------------------------- BCH(255,71) (bose_chaudhuri_hocquenghem_encoder.hh, recalled)
// g(x) = product of the 24 minimal polynomials at decode.cc:378-384 / encode.cc:272-278 (degree 184).
// Systematic: parity(x) = data(x) * x^184 mod g(x), first bit = highest power, big-endian bit packing.
struct BCH255_71 {
	static const int N = 255, K = 71, NP = 184;
	uint8_t gen[NP + 1]; // gen[d] = coefficient of x^d
	BCH255_71()
	{
		static const int polys[24] = {
			0b100011101, 0b101110111, 0b111110011, 0b101101001,
			0b110111101, 0b111100111, 0b100101011, 0b111010111,
			0b000010011, 0b101100101, 0b110001011, 0b101100011,
			0b100011011, 0b100111111, 0b110001101, 0b100101101,
			0b101011111, 0b111111001, 0b111000011, 0b100111001,
			0b110101001, 0b000011111, 0b110000111, 0b110110001};
		std::vector<uint8_t> g(1, 1);
		for (int p : polys) {
			int deg = 0;
			for (int b = 0; b < 16; ++b) if ((p >> b) & 1) deg = b;
			std::vector<uint8_t> r(g.size() + deg, 0);
			for (size_t i = 0; i < g.size(); ++i)
				if (g[i])
					for (int b = 0; b <= deg; ++b)
						if ((p >> b) & 1) r[i + b] ^= 1;
			g = r;
		}
		assert((int)g.size() == NP + 1);
		for (int d = 0; d <= NP; ++d) gen[d] = g[d];
	}
	// data: K bits (bit i = data_bits[i]); parity: NP bits
	void encode_bits(const uint8_t *data_bits, uint8_t *parity_bits) const
	{
		uint8_t reg[NP]; // reg[d] = coefficient x^d of running remainder
		std::memset(reg, 0, sizeof(reg));
		for (int i = 0; i < K; ++i) {
			uint8_t fb = data_bits[i] ^ reg[NP - 1];
			for (int d = NP - 1; d > 0; --d) reg[d] = reg[d - 1] ^ (fb & gen[d]);
			reg[0] = fb & gen[0];
		}
		for (int j = 0; j < NP; ++j) parity_bits[j] = reg[NP - 1 - j];
	}
	// systematic generator matrix rows [e_i | parity(e_i)], genmat[N*i + j] in {0,1}  (decode.cc:378)
	void matrix(int8_t *genmat) const
	{
		for (int i = 0; i < K; ++i) {
			uint8_t d[K], p[NP];
			std::memset(d, 0, K);
			d[i] = 1;
			encode_bits(d, p);
			for (int j = 0; j < K; ++j) genmat[N * i + j] = d[j];
			for (int j = 0; j < NP; ++j) genmat[N * i + K + j] = p[j];
		}
	}
};

// ---------------------------------------------------------------- OSD order 4 (osd.hh, recalled; decode.cc:199,417)
// soft[255] int8 -> most-reliable-basis reprocessing with up to 4 flips, metric sum (1-2c_i)*soft_i,
// returns unique = (best != runner-up); hard bits big-endian into out[32].
// Deviation pinned here: the reliability sort is STABLE (ties keep ascending position); the reference's
// std::sort tie order is implementation-defined.
struct OSD255_71 {
	static const int N = 255, K = 71, W = 256;
	int perm[W];
	int8_t softperm[W];
	uint8_t G[K][W];      // permuted, reduced generator
	uint8_t c0[W];        // order-0 codeword (permuted domain)
	long long visited = 0; // candidates evaluated by the last call (statistics)

	void prepare(const int8_t *soft, const int8_t *genmat)
	{
		int8_t rel[N];
		for (int i = 0; i < N; ++i) { perm[i] = i; rel[i] = (int8_t)std::abs((int)std::max<int8_t>(soft[i], -127)); }
		std::stable_sort(perm, perm + N, [&](int a, int b) { return rel[a] > rel[b]; });
		for (int j = 0; j < K; ++j) {
			for (int i = 0; i < N; ++i) G[j][i] = genmat[N * j + perm[i]];
			G[j][N] = 0;
		}
		// row echelon with column swaps (columns beyond K are pulled in when a pivot is missing)
		for (int k = 0; k < K; ++k) {
			for (int j = k; j < K; ++j)
				if (G[j][k]) { if (j != k) std::swap_ranges(G[j], G[j] + W, G[k]); break; }
			for (int j = k + 1; !G[k][k] && j < N; ++j)
				for (int h = k; h < K; ++h)
					if (G[h][j]) {
						std::swap(perm[k], perm[j]);
						for (int i = 0; i < K; ++i) std::swap(G[i][k], G[i][j]);
						if (h != k) std::swap_ranges(G[h], G[h] + W, G[k]);
						break;
					}
			assert(G[k][k]);
			for (int j = k + 1; j < K; ++j)
				if (G[j][k])
					for (int i = k; i < N; ++i) G[j][i] ^= G[k][i];
		}
		// back substitution -> [I | P]
		for (int k = K - 1; k; --k)
			for (int j = 0; j < k; ++j)
				if (G[j][k])
					for (int i = k; i < N; ++i) G[j][i] ^= G[k][i];
		for (int i = 0; i < N; ++i) softperm[i] = std::max<int8_t>(soft[perm[i]], -127);
		softperm[N] = 0;
		for (int i = 0; i < K; ++i) c0[i] = softperm[i] < 0;
		for (int i = K; i < W; ++i) {
			uint8_t b = 0;
			for (int j = 0; j < K; ++j) b ^= c0[j] & G[j][i];
			c0[i] = b;
		}
	}
	int metric(const uint8_t *c) const
	{
		int s = 0;
		for (int i = 0; i < W; ++i) s += (1 - 2 * (int)c[i]) * (int)softperm[i];
		return s;
	}
	void finish(uint8_t *out, const uint8_t *cand) const
	{
		std::memset(out, 0, 32);
		for (int i = 0; i < N; ++i) set_be_bit(out, perm[i], cand[i]);
	}
	// literal enumeration of all sum_{o<=4} C(71,o) = 1 031 347 candidates, in the reference's nesting order
	bool decode_full(uint8_t *out, const int8_t *soft, const int8_t *genmat)
	{
		prepare(soft, genmat);
		uint8_t cw[W], cand[W];
		std::memcpy(cw, c0, W);
		std::memcpy(cand, c0, W);
		int best = metric(cw), next = -1;
		visited = 1;
		auto flip = [&](int j) { for (int i = 0; i < W; ++i) cw[i] ^= G[j][i]; };
		auto update = [&]() {
			++visited;
			int m = metric(cw);
			if (m > best) { next = best; best = m; std::memcpy(cand, cw, W); }
			else if (m > next) next = m;
		};
		for (int a = 0; a < K; ++a) {
			flip(a); update();
			for (int b = a + 1; b < K; ++b) {
				flip(b); update();
				for (int c = b + 1; c < K; ++c) {
					flip(c); update();
					for (int d = c + 1; d < K; ++d) { flip(d); update(); flip(d); }
					flip(c);
				}
				flip(b);
			}
			flip(a);
		}
		finish(out, cand);
		return best != next;
	}
	// Exact-equivalent branch and bound.  With w_i = (1-2*c0_i)*softperm_i the metric of c0^e is
	// M0 - 2*D(e), D(e) = sum_{i in supp(e)} w_i; on the basis positions w_i = |soft| >= 0, so
	// D >= (sum of flipped basis w) + (sum of all negative parity w).  Sub-trees whose bound exceeds the
	// best D found so far can hold neither the winner nor a tie for it, so `best`, the winning codeword
	// and `unique` (= exactly one candidate attains the maximum and best != -1) equal decode_full's.
	bool decode_pruned(uint8_t *out, const int8_t *soft, const int8_t *genmat)
	{
		prepare(soft, genmat);
		int w[W];
		int m0 = 0, wneg = 0;
		for (int i = 0; i < W; ++i) { w[i] = (1 - 2 * (int)c0[i]) * (int)softperm[i]; m0 += w[i]; }
		for (int i = K; i < W; ++i) if (w[i] < 0) wneg += w[i];
		int bestD = 0, ties = 1; // order-0 candidate: D = 0
		int bsel[4] = {-1, -1, -1, -1};
		visited = 1;
		uint32_t rows[K][8];
		for (int j = 0; j < K; ++j) {
			std::memset(rows[j], 0, 32);
			for (int i = 0; i < W; ++i) if (G[j][i]) rows[j][i / 32] |= 1u << (i % 32);
		}
		auto dist = [&](const uint32_t *e) {
			int d = 0;
			for (int wd = 0; wd < 8; ++wd) { uint32_t x = e[wd]; while (x) { int b = __builtin_ctz(x); d += w[wd * 32 + b]; x &= x - 1; } }
			return d;
		};
		auto consider = [&](const uint32_t *e, int a, int b, int c, int d) {
			++visited;
			int D = dist(e);
			if (D < bestD) { bestD = D; ties = 1; bsel[0] = a; bsel[1] = b; bsel[2] = c; bsel[3] = d; }
			else if (D == bestD) ++ties;
		};
		uint32_t ea[8], eb[8], ec[8], ed[8];
		for (int a = 0; a < K; ++a) {
			if (w[a] + wneg > bestD) continue;
			for (int i = 0; i < 8; ++i) ea[i] = rows[a][i];
			consider(ea, a, -1, -1, -1);
			for (int b = a + 1; b < K; ++b) {
				if (w[a] + w[b] + wneg > bestD) continue;
				for (int i = 0; i < 8; ++i) eb[i] = ea[i] ^ rows[b][i];
				consider(eb, a, b, -1, -1);
				for (int c = b + 1; c < K; ++c) {
					if (w[a] + w[b] + w[c] + wneg > bestD) continue;
					for (int i = 0; i < 8; ++i) ec[i] = eb[i] ^ rows[c][i];
					consider(ec, a, b, c, -1);
					for (int d = c + 1; d < K; ++d) {
						if (w[a] + w[b] + w[c] + w[d] + wneg > bestD) continue;
						for (int i = 0; i < 8; ++i) ed[i] = ec[i] ^ rows[d][i];
						consider(ed, a, b, c, d);
					}
				}
			}
		}
		uint8_t cand[W];
		std::memcpy(cand, c0, W);
		for (int s = 0; s < 4; ++s)
			if (bsel[s] >= 0)
				for (int i = 0; i < W; ++i) cand[i] ^= G[bsel[s]][i];
		finish(out, cand);
		int best = m0 - 2 * bestD;
		return ties == 1 && best != -1;
	}
};

// ---------------------------------------------------------------- 8PSK (reference psk.hh:90-140)
struct PSK8 {
	static constexpr float cos_pi_8 = 0.92387953251128675613f;
	static constexpr float sin_pi_8 = 0.38268343236508977173f;
	static constexpr float rcp_sqrt_2 = 0.70710678118654752440f;
	static constexpr float DIST = 2 * sin_pi_8;
	// b[0..2] in {-1,+1}
	static void hard(float *b, cf c) // psk.hh:118-123
	{
		b[1] = c.re < 0.f ? -1.f : 1.f;
		b[2] = c.im < 0.f ? -1.f : 1.f;
		b[0] = std::abs(c.re) < std::abs(c.im) ? -1.f : 1.f;
	}
	static cf map(const float *b) // psk.hh:132-139
	{
		float real = cos_pi_8, imag = sin_pi_8;
		if (b[0] < 0.f) std::swap(real, imag);
		return cf(real * b[1], imag * b[2]);
	}
	static float quantize(float precision, float value) { value *= DIST * precision; return value; } // psk.hh:108-116, code_type float
	static void soft(float *b, cf c, float precision) // psk.hh:125-130
	{
		b[1] = quantize(precision, c.re);
		b[2] = quantize(precision, c.im);
		b[0] = quantize(precision, rcp_sqrt_2 * (std::abs(c.re) - std::abs(c.im)));
	}
};

// ---------------------------------------------------------------- polar code helpers
struct FrozenSet {
	const uint32_t *bits; // bit i of word i/32 set => index i frozen (polar_tables.hh)
	bool frozen(int i) const { return (bits[i / 32] >> (i % 32)) & 1; }
	bool all_frozen(int index, int n) const
	{
		if (n >= 32) {
			for (int w = index / 32; w < (index + n) / 32; ++w) if (bits[w] != 0xffffffffu) return false;
			return true;
		}
		uint32_t mask = ((1u << n) - 1u) << (index % 32);
		return (bits[index / 32] & mask) == mask;
	}
};

// x = u * F^{(x)n}, natural order, in place on bits (polar_encoder.hh PolarEncoder butterflies, recalled)
static inline void polar_transform(uint8_t *x, int n)
{
	for (int h = 1; h < n; h *= 2)
		for (int i = 0; i < n; i += 2 * h)
			for (int j = i; j < i + h; ++j) x[j] ^= x[j + h];
}
// PolarSysEnc (polar_encoder.hh, recalled; encode.cc:48,302): two-pass systematic encoding, frozen = 0
static inline void polar_sys_encode(uint8_t *code, const uint8_t *mesg, const FrozenSet &fs, int order)
{
	int n = 1 << order;
	for (int i = 0, j = 0; i < n; ++i) code[i] = fs.frozen(i) ? 0 : mesg[j++];
	polar_transform(code, n);
	for (int i = 0; i < n; ++i) if (fs.frozen(i)) code[i] = 0;
	polar_transform(code, n);
}

// ---------------------------------------------------------------- polar SCL decoder (polar_list_decoder.hh, recalled; decode.cc:201,530)
// Natural-order min-sum SC list decoding with L lanes (L = SIMD width of the reference: 8 with AVX2, else 4).
//   f(a,b) = sgn(a) sgn(b) min(|a|,|b|)          (PolarHelper::prod)
//   g(a,b,u) = u*a + b,  u in {+1,-1}            (PolarHelper::madd)
//   frozen leaf : u = +1, metric += |llr| if llr < 0
//   free leaf   : 2L forks (lane k, bit 0) / (lane k, bit 1); metric += |llr| on the branch that disagrees with
//                 sign(llr); keep the L smallest.  Initial metrics: lane 0 = 0, others 1000.
// Restatement choices where the reference is implementation-defined or only mathematically pinned:
//   * survivors are kept in (metric, fork index 2k+bit) ascending order — a full stable sort instead of
//     std::nth_element, so ties are deterministic;
//   * a maximal all-frozen sub-tree (any size up to r0_max) is handled as one rate-0 node: metric[k] +=
//     sum_i (alpha_i < 0 ? -alpha_i : 0) accumulated in index order, which equals the leaf-by-leaf min-sum
//     accumulation in exact arithmetic (rounding order differs).  r0_max = 1 gives the leaf-by-leaf order.
//   * final candidate order = ascending metric (stable).
template <int L>
class PolarListDecoder {
public:
	struct Map { uint8_t v[L]; };
	int order, n;
	FrozenSet fs;
	int r0_max = 1 << 16;
	std::vector<float> soft;   // [2n][L], level buffers in heap layout: node of size s reads soft[s+i], writes soft[s/2+i]
	std::vector<uint8_t> hard; // [n][L] partial sums as bits
	float metric[L];
	long long forks = 0;

	PolarListDecoder(int order_, const uint32_t *frozen) : order(order_), n(1 << order_), fs{frozen},
		soft((size_t)2 * n * L), hard((size_t)n * L) {}

	static inline float prod(float a, float b)
	{
		float m = std::min(std::abs(a), std::abs(b));
		bool neg = (a < 0.f) != (b < 0.f);
		if (a == 0.f || b == 0.f) return 0.f;
		return neg ? -m : m;
	}
	Map identity() const { Map m; for (int k = 0; k < L; ++k) m.v[k] = k; return m; }

	Map rate0(int index, int s)
	{
		for (int i = 0; i < s; ++i)
			for (int k = 0; k < L; ++k) {
				float a = soft[(size_t)(s + i) * L + k];
				if (a < 0.f) metric[k] -= a;
				hard[(size_t)(index + i) * L + k] = 0;
			}
		return identity();
	}
	Map leaf_free(int index)
	{
		++forks;
		float fork[2 * L];
		for (int k = 0; k < L; ++k) {
			float a = soft[(size_t)1 * L + k];
			fork[2 * k] = fork[2 * k + 1] = metric[k];
			if (a < 0.f) fork[2 * k] -= a;
			else fork[2 * k + 1] += a;
		}
		int perm[2 * L];
		for (int k = 0; k < 2 * L; ++k) perm[k] = k;
		std::stable_sort(perm, perm + 2 * L, [&](int a, int b) { return fork[a] < fork[b]; });
		Map m;
		for (int k = 0; k < L; ++k) {
			metric[k] = fork[perm[k]];
			m.v[k] = perm[k] >> 1;
			hard[(size_t)index * L + k] = perm[k] & 1;
		}
		return m;
	}
	Map node(int level, int index)
	{
		int s = 1 << level;
		if (s <= r0_max && fs.all_frozen(index, s)) return rate0(index, s);
		if (level == 0) return fs.frozen(index) ? rate0(index, 1) : leaf_free(index);
		int h = s / 2;
		for (int i = 0; i < h; ++i)
			for (int k = 0; k < L; ++k)
				soft[(size_t)(h + i) * L + k] = prod(soft[(size_t)(s + i) * L + k], soft[(size_t)(s + h + i) * L + k]);
		Map lm = node(level - 1, index);
		for (int i = 0; i < h; ++i)
			for (int k = 0; k < L; ++k) {
				float a = soft[(size_t)(s + i) * L + lm.v[k]], b = soft[(size_t)(s + h + i) * L + lm.v[k]];
				soft[(size_t)(h + i) * L + k] = hard[(size_t)(index + i) * L + k] ? b - a : b + a;
			}
		Map rm = node(level - 1, index + h);
		for (int i = 0; i < h; ++i) {
			uint8_t t[L];
			for (int k = 0; k < L; ++k) t[k] = hard[(size_t)(index + i) * L + rm.v[k]] ^ hard[(size_t)(index + h + i) * L + k];
			for (int k = 0; k < L; ++k) hard[(size_t)(index + i) * L + k] = t[k];
		}
		Map m;
		for (int k = 0; k < L; ++k) m.v[k] = lm.v[rm.v[k]];
		return m;
	}
	// codeword: n channel LLRs (positive = bit 0).  Output: x_lanes[lane][n] re-encoded codeword bits of the L
	// survivors in ascending-metric order (== systematic message at the non-frozen indices, decode.cc:254-261),
	// final metrics in the same order.
	void decode(const float *codeword, std::vector<std::vector<uint8_t>> &x_lanes, float *metrics_out)
	{
		metric[0] = 0;
		for (int k = 1; k < L; ++k) metric[k] = 1000;
		forks = 0;
		for (int i = 0; i < n; ++i)
			for (int k = 0; k < L; ++k) soft[(size_t)(n + i) * L + k] = codeword[i];
		node(order, 0);
		int perm[L];
		for (int k = 0; k < L; ++k) perm[k] = k;
		std::stable_sort(perm, perm + L, [&](int a, int b) { return metric[a] < metric[b]; });
		x_lanes.assign(L, std::vector<uint8_t>(n));
		for (int k = 0; k < L; ++k) {
			for (int i = 0; i < n; ++i) x_lanes[k][i] = hard[(size_t)i * L + perm[k]];
			metrics_out[k] = metric[perm[k]];
		}
	}
};

// ---------------------------------------------------------------- Theil–Sen (theil_sen.hh, recalled; decode.cc:195,488-494)
// slope = upper median (element count/2 after nth_element) of all pairwise slopes, intercept likewise.
struct TheilSen {
	float slope = 0, yint = 0;
	std::vector<float> tmp;
	void compute(const float *x, const float *y, int len)
	{
		tmp.clear();
		for (int i = 0; i < len; ++i)
			for (int j = i + 1; j < len; ++j)
				if (x[j] != x[i]) tmp.push_back((y[j] - y[i]) / (x[j] - x[i]));
		size_t c = tmp.size();
		std::nth_element(tmp.begin(), tmp.begin() + c / 2, tmp.end());
		slope = tmp[c / 2];
		tmp.clear();
		for (int i = 0; i < len; ++i) tmp.push_back(y[i] - slope * x[i]);
		c = tmp.size();
		std::nth_element(tmp.begin(), tmp.begin() + c / 2, tmp.end());
		yint = tmp[c / 2];
	}
	float operator()(float x) const { return yint + slope * x; }
};

} // namespace ref
