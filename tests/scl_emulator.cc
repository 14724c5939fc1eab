// tests/scl_emulator.cc — TEST HELPER.  A scalar, lane-array emulation of the CUDA list decoder's data flow
// (modem_b200/csrc/polar.cu): same op schedule (host_tables.cc), same bit-packed partial sums, same lane-map
// bookkeeping, same fp32 operation order — with warp shuffles replaced by array indexing.  The CPU test-suite
// checks it against the recursive oracle (oracle/ref_code.hh) so that algorithmic mistakes in the schedule /
// map algebra are caught without a GPU; the GPU tests then check the kernel against the same oracle.
#include "../modem_b200/csrc/host_tables.h"
#include <cmath>
#include <cstring>
#include <algorithm>

using namespace ofdmrx;

namespace {
const int L = 8;

inline float ff(float a, float b)
{
	uint32_t ua, ub;
	std::memcpy(&ua, &a, 4); std::memcpy(&ub, &b, 4);
	float m = std::min(std::fabs(a), std::fabs(b));
	uint32_t um;
	std::memcpy(&um, &m, 4);
	um |= (ua ^ ub) & 0x80000000u;
	std::memcpy(&m, &um, 4);
	return m;
}
inline float gg(float a, float b, uint32_t bit) { return bit ? b - a : b + a; }

struct Emu {
	std::vector<uint32_t> frozen, ops;
	std::vector<std::vector<float>> A; // A[l]: [2^l][L]
	std::vector<uint32_t> B;           // [2048][L]
	float metric[L];
	int ret[L];
	int lm[17][L];
	long long forks = 0, fast_hits = 0, keep_all = 0;

	void fork_leaf(const float *llr, uint32_t *bit_out)
	{
		// ranks of the 2L forks by (metric, fork index), survivors written in rank order
		float m0[L], m1[L];
		for (int t = 0; t < L; ++t) {
			float a = llr[t], pen = std::fabs(a);
			m0[t] = metric[t] + (a < 0.f ? pen : 0.f);
			m1[t] = metric[t] + (a < 0.f ? 0.f : pen);
		}
		{ // statistics: would the 'all keeps survive and lanes already sorted' shortcut apply?
			float maxK = metric[0], minF = 1e30f; bool sorted = true;
			for (int t = 0; t < L; ++t) { maxK = std::max(maxK, metric[t]); minF = std::min(minF, metric[t] + std::fabs(llr[t])); if (t && metric[t] < metric[t-1]) sorted = false; }
			if (maxK < minF) { ++keep_all; if (sorted) ++fast_hits; }
		}
		int src[L], bit[L];
		float nm[L];
		for (int t = 0; t < L; ++t) {
			int r0 = 0, r1 = 0;
			for (int j = 0; j < L; ++j) {
				r0 += (m0[j] < m0[t]) || (m0[j] == m0[t] && j < t);
				r0 += (m1[j] < m0[t]) || (m1[j] == m0[t] && j < t);
				r1 += (m0[j] < m1[t]) || (m0[j] == m1[t] && j <= t);
				r1 += (m1[j] < m1[t]) || (m1[j] == m1[t] && j < t);
			}
			if (r0 < L) { src[r0] = t; bit[r0] = 0; nm[r0] = m0[t]; }
			if (r1 < L) { src[r1] = t; bit[r1] = 1; nm[r1] = m1[t]; }
		}
		for (int t = 0; t < L; ++t) { metric[t] = nm[t]; ret[t] = src[t]; bit_out[t] = bit[t]; }
		++forks;
	}

	// in-block recursion over levels 4..0 (a 32-leaf word); mirrors polar.cu's blk_node<LVL,BASE>
	float a[5][16][L]; // a[l][k][lane], level-l buffer has 2^l entries
	uint32_t W[L];
	int lmb[6][L];
	const float *A5 = nullptr;
	uint32_t fmask = 0;

	// parent value k of the level-(lvl) buffer as seen by lane s
	float parent(int lvl, int k, int s) const { return lvl == 5 ? A5[(size_t)k * L + s] : a[lvl][k][s]; }

	void blk_node(int lvl, int base)
	{
		const int n = 1 << lvl;
		const uint32_t sub = n == 32 ? 0xffffffffu : (((1u << n) - 1u) << base);
		if (lvl < 5 && (fmask & sub) == sub) { // rate-0 node inside the word (size 1..16)
			for (int k = 0; k < n; ++k)
				for (int t = 0; t < L; ++t) {
					float v = a[lvl][k][t];
					if (v < 0.f) metric[t] -= v;
				}
			for (int t = 0; t < L; ++t) ret[t] = t;
			return;
		}
		if (lvl == 0) {
			float llr[L];
			uint32_t bit[L];
			for (int t = 0; t < L; ++t) llr[t] = a[0][0][t];
			fork_leaf(llr, bit);
			for (int t = 0; t < L; ++t) W[t] |= bit[t] << base;
			return;
		}
		const int h = n / 2;
		for (int k = 0; k < h; ++k)
			for (int t = 0; t < L; ++t) a[lvl - 1][k][t] = ff(parent(lvl, k, t), parent(lvl, k + h, t));
		blk_node(lvl - 1, base);
		for (int t = 0; t < L; ++t) lmb[lvl][t] = ret[t];
		{
			float tmp[16][L];
			for (int k = 0; k < h; ++k)
				for (int t = 0; t < L; ++t) {
					int s = ret[t];
					tmp[k][t] = gg(parent(lvl, k, s), parent(lvl, k + h, s), (W[t] >> (base + k)) & 1);
				}
			for (int k = 0; k < h; ++k)
				for (int t = 0; t < L; ++t) a[lvl - 1][k][t] = tmp[k][t];
		}
		blk_node(lvl - 1, base + h);
		{
			uint32_t maskL = ((1u << h) - 1u) << base;
			uint32_t Wn[L];
			int rn[L];
			for (int t = 0; t < L; ++t) {
				uint32_t Wl = W[ret[t]];
				Wn[t] = (W[t] & ~maskL) | ((Wl ^ (W[t] >> h)) & maskL);
				rn[t] = lmb[lvl][ret[t]];
			}
			for (int t = 0; t < L; ++t) { W[t] = Wn[t]; ret[t] = rn[t]; }
		}
	}
	void word_block(int index)
	{
		fmask = frozen[index / 32];
		A5 = A[5].data();
		for (int t = 0; t < L; ++t) W[t] = 0;
		blk_node(5, 0);
		for (int t = 0; t < L; ++t) B[(size_t)(index / 32) * L + t] = W[t];
	}

	void run(const float *llr)
	{
		A.assign(17, std::vector<float>());
		for (int l = 5; l <= 15; ++l) A[l].assign((size_t)(1 << l) * L, 0.f);
		B.assign((size_t)2048 * L, 0u);
		metric[0] = 0.f;
		for (int t = 1; t < L; ++t) metric[t] = 1000.f;
		for (int t = 0; t < L; ++t) ret[t] = t;
		forks = 0; fast_hits = 0; keep_all = 0;
		for (size_t pc = 0;; ++pc) {
			uint32_t w = ops[pc];
			uint32_t op = scl_op(w), l = scl_level(w), index = scl_index(w);
			if (op == OP_END) break;
			int n = 1 << l, h = n / 2;
			switch (op) {
			case OP_F:
			case OP_G: {
				if (op == OP_G) for (int t = 0; t < L; ++t) lm[l][t] = ret[t];
				for (int i = 0; i < h; ++i)
					for (int t = 0; t < L; ++t) {
						int s = op == OP_G ? ret[t] : t;
						float pa = l == 16 ? llr[i] : A[l][(size_t)i * L + s];
						float pb = l == 16 ? llr[i + h] : A[l][(size_t)(i + h) * L + s];
						if (op == OP_G) {
							uint32_t bit = (B[(size_t)((index + i) / 32) * L + t] >> ((index + i) % 32)) & 1;
							A[l - 1][(size_t)i * L + t] = gg(pa, pb, bit);
						} else {
							A[l - 1][(size_t)i * L + t] = ff(pa, pb);
						}
					}
				// fused F steps down the left spine of the child just produced
				for (uint32_t d = 1; d < scl_depth(w); ++d) {
					int ll = l - d, hh = 1 << (ll - 1);
					for (int i = 0; i < hh; ++i)
						for (int t = 0; t < L; ++t)
							A[ll - 1][(size_t)i * L + t] = ff(A[ll][(size_t)i * L + t], A[ll][(size_t)(i + hh) * L + t]);
				}
				break;
			}
			case OP_TOP: { // level-13 node j from the channel LLRs and the betas of its left-hand relatives
				const int j = index >> 13, j2 = (j >> 2) & 1, j1 = (j >> 1) & 1, j0 = j & 1;
				if (j > 0) { const int lv = 14 + __builtin_ctz(j); for (int t = 0; t < L; ++t) lm[lv][t] = ret[t]; }
				int s14[L], s15[L];
				for (int t = 0; t < L; ++t) { int u = j0 ? lm[14][t] : t; s14[t] = u; s15[t] = j1 ? lm[15][u] : u; }
				auto bit = [&](int base, int p, int lane) { return (B[(size_t)((base + p) / 32) * L + lane] >> ((base + p) % 32)) & 1u; };
				for (int i = 0; i < 8192; ++i)
					for (int t = 0; t < L; ++t) {
						float cc[8], x[4], y[2];
						for (int k = 0; k < 8; ++k) cc[k] = llr[i + 8192 * k];
						for (int m = 0; m < 4; ++m) x[m] = j2 ? gg(cc[m], cc[m + 4], bit(0, i + 8192 * m, s15[t])) : ff(cc[m], cc[m + 4]);
						for (int m = 0; m < 2; ++m) y[m] = j1 ? gg(x[m], x[m + 2], bit(j2 * 32768, i + 8192 * m, s14[t])) : ff(x[m], x[m + 2]);
						A[13][(size_t)i * L + t] = j0 ? gg(y[0], y[1], bit((j - 1) * 8192, i, t)) : ff(y[0], y[1]);
					}
				for (uint32_t d = 1; d < scl_depth(w); ++d) {
					int ll = 13 - (d - 1), hh = 1 << (ll - 1);
					for (int i = 0; i < hh; ++i)
						for (int t = 0; t < L; ++t)
							A[ll - 1][(size_t)i * L + t] = ff(A[ll][(size_t)i * L + t], A[ll][(size_t)(i + hh) * L + t]);
				}
				break;
			}
			case OP_WORD:
				word_block(index);
				break;
			case OP_R0:
				for (int i = 0; i < n; ++i)
					for (int t = 0; t < L; ++t) {
						float v = l == 16 ? llr[i] : A[l][(size_t)i * L + t];
						if (v < 0.f) metric[t] -= v;
					}
				for (int w2 = 0; w2 < n / 32; ++w2)
					for (int t = 0; t < L; ++t) B[(size_t)(index / 32 + w2) * L + t] = 0;
				for (int t = 0; t < L; ++t) ret[t] = t;
				break;
			case OP_C: {
				int hw = h / 32;
				for (int w2 = 0; w2 < hw; ++w2) {
					uint32_t nw[L];
					for (int t = 0; t < L; ++t)
						nw[t] = B[(size_t)(index / 32 + w2) * L + ret[t]] ^ B[(size_t)(index / 32 + hw + w2) * L + t];
					for (int t = 0; t < L; ++t) B[(size_t)(index / 32 + w2) * L + t] = nw[t];
				}
				int rn[L];
				for (int t = 0; t < L; ++t) rn[t] = lm[l][ret[t]];
				for (int t = 0; t < L; ++t) ret[t] = rn[t];
				break;
			}
			default: break;
			}
		}
	}
};
} // namespace

extern "C" {
// llr[65536] -> lanes[8][65536] codeword bits in ascending (metric, lane) order + metrics[8]
static void emu_polar_decode_table(int table, const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks)
{
	static Emu emu[2];
	Emu &e = emu[table ? 1 : 0];
	if (e.frozen.empty()) {
		e.frozen = make_frozen(kCodeOrder, table ? 64512 : kConsBits, kCrcBits);
		e.ops = make_scl_schedule(e.frozen, kCodeOrder);
	}
	e.run(llr);
	int perm[L];
	for (int t = 0; t < L; ++t) perm[t] = t;
	std::stable_sort(perm, perm + L, [&](int a, int b) { return e.metric[a] < e.metric[b]; });
	for (int k = 0; k < L; ++k) {
		metrics_out[k] = e.metric[perm[k]];
		for (int i = 0; i < kCodeLen; ++i) lanes_out[(size_t)k * kCodeLen + i] = (e.B[(size_t)(i / 32) * L + perm[k]] >> (i % 32)) & 1;
	}
	if (forks) { forks[0] = e.forks; forks[1] = e.fast_hits; forks[2] = e.keep_all; }
}
void emu_polar_decode(const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks) { emu_polar_decode_table(0, llr, lanes_out, metrics_out, forks); }
// table 1: the frozen set of modes 10..13 (decode.cc:342-343)
void emu_polar_decode_alt(const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks) { emu_polar_decode_table(1, llr, lanes_out, metrics_out, forks); }
void host_frozen(uint32_t *out) { auto f = make_frozen(kCodeOrder, kConsBits, kCrcBits); std::memcpy(out, f.data(), 2048 * 4); }
void host_frozen_alt(uint32_t *out) { auto f = make_frozen(kCodeOrder, 64512, kCrcBits); std::memcpy(out, f.data(), 2048 * 4); }
int host_schedule(uint32_t *out, int cap)
{
	auto f = make_frozen(kCodeOrder, kConsBits, kCrcBits);
	auto s = make_scl_schedule(f, kCodeOrder);
	if (out) std::memcpy(out, s.data(), std::min<size_t>(cap, s.size()) * 4);
	return (int)s.size();
}
void host_bch_rows(uint32_t *out) { auto r = bch_generator_rows(); std::memcpy(out, r.data(), r.size() * 4); }
void host_mls(int poly, int n, uint8_t *out) { auto m = mls_bits(poly, n); std::memcpy(out, m.data(), n); }
unsigned host_crc16(uint64_t v) { return crc16_u64(v); }
void host_hilbert(float *reco, float *im5) { auto c = hilbert_coeffs(kFilterLen, reco); std::memcpy(im5, c.data(), c.size() * 4); }
}
