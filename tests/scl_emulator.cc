// tests/scl_emulator.cc — TEST HELPER.  A scalar, lane-array emulation of the CUDA list decoder's data flow
// (modem_b200/csrc/polar.cu): same op schedule (host_tables.cc), same bit-packed partial sums, same lane-map
// bookkeeping, same fp32 operation order — with warp shuffles replaced by array indexing.  The CPU test-suite
// checks it against the recursive oracle (oracle/ref_code.hh) so that algorithmic mistakes in the schedule /
// map algebra are caught without a GPU; the GPU tests then check the kernel against the same oracle.
#include "../modem_b200/csrc/host_tables.h"
#include <cmath>
#include <cstring>
#include <cmath>
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

// Emulation of the kernel's data flow.  Differences from a plain lane-array SCL that the kernel relies on and that this file
// therefore spells out (and checks: every buffer slot nobody should read is poisoned with NaN / 0xDEADBEEF):
//   * path classes: lanes that hold the same path (the 0 / 1000 initial metrics make all eight lanes copies of one path
//     until a flip displaces a copy) form a class; only the class representative rep[t] = lowest lane of the class owns
//     alpha / beta storage ("slot"), everybody else reads it.  Classes change only in the ranked (slow) fork.
//   * rs[l][t]: the representative that lane t had when the data of level l it belongs to was written (alpha_l of the
//     current node until G(l), from then on the left child's beta until C(l)); reads go through the lane maps first
//     (which lane was I then?) and through rs second (whose slot held that lane's data then?).
//   * rate-1 attempts (OP_R1 above the words, and on all-free sub-blocks inside a word).
struct Emu {
	std::vector<uint32_t> frozen, ops;
	std::vector<std::vector<float>> A; // A[l]: [slot][2^l]
	std::vector<uint32_t> B;           // [2048][slot]
	float metric[L];
	int ret[L], rep[L];
	int lm[17][L], rs[17][L];
	long long forks = 0, fast_hits = 0, keep_all = 0;
	long long r1_try = 0, r1_pass = 0, r1w_try = 0, r1w_pass = 0, words = 0, elem_ops = 0, class_hist[L + 1] = {0};
	uint32_t fail_seed = 0; // != 0: attempts and leaf shortcuts are refused at random (as another codeword of the warp would force)
	bool refuse()
	{
		if (!fail_seed) return false;
		fail_seed ^= fail_seed << 13; fail_seed ^= fail_seed >> 17; fail_seed ^= fail_seed << 5;
		return (fail_seed & 3u) == 0u;
	}
	bool is_rep(int t) const { return rep[t] == t; }
	float *lvl(int l, int slot) { return A[l].data() + ((size_t)slot << l); }
	void poison(int l) { for (int t = 0; t < L; ++t) if (!is_rep(t)) std::fill(lvl(l, t), lvl(l, t) + (1 << l), NAN); }

	void fork_leaf(const float *llr, uint32_t *bit_out)
	{
		// ranks of the 2L forks by (metric, fork index), survivors written in rank order
		float m0[L], m1[L];
		for (int t = 0; t < L; ++t) {
			float a = llr[t], pen = std::fabs(a);
			m0[t] = metric[t] + (a < 0.f ? pen : 0.f);
			m1[t] = metric[t] + (a < 0.f ? 0.f : pen);
		}
		bool fast = !refuse();
		for (int t = 0; t < L; ++t) fast = fast && metric[L - 1] < metric[t] + std::fabs(llr[t]) && (t == 0 || metric[t - 1] <= metric[t]);
		++forks;
		{ int n = 0; for (int t = 0; t < L; ++t) n += is_rep(t); ++class_hist[n]; }
		if (fast) { // all keeps survive in place
			++fast_hits;
			for (int t = 0; t < L; ++t) { ret[t] = t; bit_out[t] = llr[t] < 0.f; }
			return;
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
		int key[L], nrep[L];
		for (int t = 0; t < L; ++t) key[t] = rep[src[t]] * 2 + bit[t];
		for (int t = 0; t < L; ++t) { int j = 0; while (key[j] != key[t]) ++j; nrep[t] = j; }
		for (int t = 0; t < L; ++t) { metric[t] = nm[t]; ret[t] = src[t]; bit_out[t] = bit[t]; rep[t] = nrep[t]; }
	}
	// rate-1 attempt on n alphas per lane (mn[t] = min |alpha|): true = the list cannot change inside the node
	bool rate1_ok(const float *mn)
	{
		bool ok = !refuse();
		for (int t = 0; t < L; ++t) ok = ok && metric[L - 1] < metric[t] + mn[t] && (t == 0 || metric[t - 1] <= metric[t]);
		return ok;
	}

	// in-block recursion over levels 4..0 (a 32-leaf word); mirrors polar.cu's blk_node<LVL,BASE>
	float a[5][16][L]; // a[l][k][lane], level-l buffer has 2^l entries (every lane computes its own copy: registers)
	uint32_t W[L];
	int lmb[6][L];
	int rs5[L];
	uint32_t fmask = 0;

	// parent value k of the level-(lvl) buffer as seen by lane s (at the start of the word for level 5)
	float parent(int lvl, int k, int s) { return lvl == 5 ? lvl5(rs5[s])[k] : a[lvl][k][s]; }
	float *lvl5(int slot) { return lvl(5, slot); }

	void blk_node(int lvl, int base)
	{
		const int n = 1 << lvl;
		const uint32_t sub = n == 32 ? 0xffffffffu : (((1u << n) - 1u) << base);
		if ((fmask & sub) == sub) { // rate-0 node inside the word (size 1..16; a whole frozen word is an OP_R0)
			for (int k = 0; k < n; ++k)
				for (int t = 0; t < L; ++t) {
					float v = parent(lvl, k, t);
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
		if ((fmask & sub) == 0u) { // rate-1 attempt inside the word
			float mn[L];
			for (int t = 0; t < L; ++t) { mn[t] = INFINITY; for (int k = 0; k < n; ++k) mn[t] = std::min(mn[t], std::fabs(parent(lvl, k, t))); }
			++r1w_try;
			if (rate1_ok(mn)) {
				++r1w_pass;
				for (int t = 0; t < L; ++t) {
					for (int k = 0; k < n; ++k) W[t] |= (uint32_t)(parent(lvl, k, t) < 0.f) << (base + k);
					ret[t] = t;
				}
				return;
			}
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
		++words;
		fmask = frozen[index / 32];
		for (int t = 0; t < L; ++t) { W[t] = 0; rs5[t] = rs[5][t]; }
		blk_node(5, 0);
		for (int t = 0; t < L; ++t) B[(size_t)(index / 32) * L + t] = is_rep(t) ? W[t] : 0xDEADBEEFu; // only the representatives' slots are written
	}
	// the d-1 fused F steps below the child (level l-1) just produced in the representatives' slots
	void chain(int l, uint32_t depth)
	{
		for (uint32_t d = 1; d < depth; ++d) {
			int ll = l - d, hh = 1 << (ll - 1);
			for (int r = 0; r < L; ++r) {
				if (!is_rep(r)) continue;
				float *P = lvl(ll, r), *D = lvl(ll - 1, r);
				for (int i = 0; i < hh; ++i) D[i] = ff(P[i], P[i + hh]);
				elem_ops += hh;
			}
			for (int t = 0; t < L; ++t) rs[ll - 1][t] = rep[t];
			poison(ll - 1);
		}
	}

	void run(const float *llr)
	{
		A.assign(17, std::vector<float>());
		for (int l = 5; l <= 15; ++l) A[l].assign((size_t)(1 << l) * L, NAN);
		B.assign((size_t)2048 * L, 0xDEADBEEFu);
		metric[0] = 0.f;
		for (int t = 1; t < L; ++t) metric[t] = 1000.f;
		for (int t = 0; t < L; ++t) { ret[t] = t; rep[t] = 0; }
		for (int l = 0; l < 17; ++l) for (int t = 0; t < L; ++t) { lm[l][t] = t; rs[l][t] = 0; }
		forks = fast_hits = keep_all = r1_try = r1_pass = r1w_try = r1w_pass = words = elem_ops = 0;
		for (int k = 0; k <= L; ++k) class_hist[k] = 0;
		for (size_t pc = 0;; ++pc) {
			uint32_t w = ops[pc];
			uint32_t op = scl_op(w), l = scl_level(w), index = scl_index(w);
			if (op == OP_END) break;
			int n = 1 << l, h = n / 2;
			switch (op) {
			case OP_F:
			case OP_G: {
				if (op == OP_G) for (int t = 0; t < L; ++t) lm[l][t] = ret[t];
				for (int r = 0; r < L; ++r) {
					if (!is_rep(r)) continue;
					// whose slot holds my level-l alphas?  F: mine (just written); G: that of the lane I was when the node started
					const int ps = op == OP_G ? rs[l][ret[r]] : rs[l][r];
					const float *P = l == 16 ? llr : lvl(l, ps);
					float *D = lvl(l - 1, r);
					for (int i = 0; i < h; ++i) {
						if (op == OP_G) {
							uint32_t bit = (B[(size_t)((index + i) / 32) * L + r] >> ((index + i) % 32)) & 1;
							D[i] = gg(P[i], P[i + h], bit);
						} else {
							D[i] = ff(P[i], P[i + h]);
						}
					}
					elem_ops += h;
				}
				for (int t = 0; t < L; ++t) { rs[l - 1][t] = rep[t]; if (op == OP_G) rs[l][t] = rep[t]; }
				poison(l - 1);
				chain(l, scl_depth(w));
				break;
			}
			case OP_TOP: { // level-13 node j from the channel LLRs and the betas of its left-hand relatives
				const int j = index >> 13, j2 = (j >> 2) & 1, j1 = (j >> 1) & 1, j0 = j & 1;
				if (j > 0) { const int lv = 14 + __builtin_ctz(j); for (int t = 0; t < L; ++t) { lm[lv][t] = ret[t]; rs[lv][t] = rep[t]; } }
				auto bit = [&](int base, int p, int slot) { return (B[(size_t)((base + p) / 32) * L + slot] >> ((base + p) % 32)) & 1u; };
				for (int r = 0; r < L; ++r) {
					if (!is_rep(r)) continue;
					const int u = j0 ? lm[14][r] : r;          // my lane when the level-14 relative completed
					const int v = j1 ? lm[15][u] : u;          // ... and when node (15, 0) completed
					const int s13 = r, s14 = rs[15][u], s15 = rs[16][v]; // slots of the three relatives' betas
					float *D = lvl(13, r);
					for (int i = 0; i < 8192; ++i) {
						float cc[8], x[4], y[2];
						for (int k = 0; k < 8; ++k) cc[k] = llr[i + 8192 * k];
						for (int m = 0; m < 4; ++m) x[m] = j2 ? gg(cc[m], cc[m + 4], bit(0, i + 8192 * m, s15)) : ff(cc[m], cc[m + 4]);
						for (int m = 0; m < 2; ++m) y[m] = j1 ? gg(x[m], x[m + 2], bit(j2 * 32768, i + 8192 * m, s14)) : ff(x[m], x[m + 2]);
						D[i] = j0 ? gg(y[0], y[1], bit((j - 1) * 8192, i, s13)) : ff(y[0], y[1]);
					}
					elem_ops += 8192;
				}
				for (int t = 0; t < L; ++t) rs[13][t] = rep[t];
				poison(13);
				chain(14, scl_depth(w));
				break;
			}
			case OP_WORD:
				word_block(index);
				break;
			case OP_R0:
				for (int t = 0; t < L; ++t) {
					const float *P = lvl(l, rs[l][t]);
					for (int i = 0; i < n; ++i)
						if (P[i] < 0.f) metric[t] -= P[i];
				}
				for (int w2 = 0; w2 < n / 32; ++w2)
					for (int t = 0; t < L; ++t) B[(size_t)(index / 32 + w2) * L + t] = is_rep(t) ? 0u : 0xDEADBEEFu;
				for (int t = 0; t < L; ++t) ret[t] = t;
				break;
			case OP_R1: {
				float mn[L];
				for (int t = 0; t < L; ++t) {
					const float *P = lvl(l, rs[l][t]);
					mn[t] = INFINITY;
					for (int i = 0; i < n; ++i) mn[t] = std::min(mn[t], std::fabs(P[i]));
				}
				++r1_try;
				if (rate1_ok(mn)) {
					++r1_pass;
					for (int w2 = 0; w2 < n / 32; ++w2)
						for (int t = 0; t < L; ++t) {
							uint32_t x = 0xDEADBEEFu;
							if (is_rep(t)) {
								const float *P = lvl(l, rs[l][t]) + 32 * w2;
								x = 0;
								for (int b = 0; b < 32; ++b) x |= (uint32_t)(P[b] < 0.f) << b;
							}
							B[(size_t)(index / 32 + w2) * L + t] = x;
						}
					for (int t = 0; t < L; ++t) ret[t] = t;
					pc = ops[pc + 1] - 1;
				} else {
					++pc; // skip the target word
				}
				break;
			}
			case OP_C: {
				int hw = h / 32;
				for (int w2 = 0; w2 < hw; ++w2) {
					uint32_t nw[L];
					for (int t = 0; t < L; ++t)
						nw[t] = is_rep(t) ? B[(size_t)(index / 32 + w2) * L + rs[l][ret[t]]] ^ B[(size_t)(index / 32 + hw + w2) * L + t] : 0xDEADBEEFu;
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

static uint32_t g_fail_seed = 0;
extern "C" {
// != 0: rate-1 attempts and leaf shortcuts are refused at random, as another codeword of the warp would force (both branches are exact)
void emu_set_fail_seed(uint32_t s) { g_fail_seed = s; }
// llr[65536] -> lanes[8][65536] codeword bits in ascending (metric, lane) order + metrics[8]
static void emu_polar_decode_table(int table, const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks)
{
	static Emu emu[2];
	Emu &e = emu[table ? 1 : 0];
	if (e.frozen.empty()) {
		e.frozen = make_frozen(kCodeOrder, table ? 64512 : kConsBits, kCrcBits);
		e.ops = make_scl_schedule(e.frozen, kCodeOrder, kSclMaxFuse, true, true);
	}
	e.fail_seed = g_fail_seed;
	e.run(llr);
	int perm[L];
	for (int t = 0; t < L; ++t) perm[t] = t;
	std::stable_sort(perm, perm + L, [&](int a, int b) { return e.metric[a] < e.metric[b]; });
	for (int k = 0; k < L; ++k) {
		metrics_out[k] = e.metric[perm[k]];
		for (int i = 0; i < kCodeLen; ++i) lanes_out[(size_t)k * kCodeLen + i] = (e.B[(size_t)(i / 32) * L + e.rep[perm[k]]] >> (i % 32)) & 1;
	}
	if (forks) {
		forks[0] = e.forks; forks[1] = e.fast_hits; forks[2] = e.r1_try; forks[3] = e.r1_pass; forks[4] = e.r1w_try; forks[5] = e.r1w_pass;
		forks[6] = e.words; forks[7] = e.elem_ops;
		for (int k = 1; k <= L; ++k) forks[7 + k] = e.class_hist[k];
	}
}
void emu_polar_decode(const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks) { emu_polar_decode_table(0, llr, lanes_out, metrics_out, forks); }
// table 1: the frozen set of modes 10..13 (decode.cc:342-343)
void emu_polar_decode_alt(const float *llr, uint8_t *lanes_out, float *metrics_out, long long *forks) { emu_polar_decode_table(1, llr, lanes_out, metrics_out, forks); }
void host_frozen(uint32_t *out) { auto f = make_frozen(kCodeOrder, kConsBits, kCrcBits); std::memcpy(out, f.data(), 2048 * 4); }
void host_frozen_alt(uint32_t *out) { auto f = make_frozen(kCodeOrder, 64512, kCrcBits); std::memcpy(out, f.data(), 2048 * 4); }
int host_schedule(uint32_t *out, int cap)
{
	auto f = make_frozen(kCodeOrder, kConsBits, kCrcBits);
	auto s = make_scl_schedule(f, kCodeOrder, kSclMaxFuse, true, true);
	if (out) std::memcpy(out, s.data(), std::min<size_t>(cap, s.size()) * 4);
	return (int)s.size();
}
// CRC pieces of the list decoder's epilogue (host_tables.cc: crc32_pieces): 16 + 256 words
void host_crc_pieces(int table, uint32_t *out)
{
	auto f = make_frozen(kCodeOrder, table ? 64512 : kConsBits, kCrcBits);
	auto p = crc32_pieces(f, kCrcBits);
	std::memcpy(out, p.data(), p.size() * 4);
}
void host_bch_rows(uint32_t *out) { auto r = bch_generator_rows(); std::memcpy(out, r.data(), r.size() * 4); }
void host_mls(int poly, int n, uint8_t *out) { auto m = mls_bits(poly, n); std::memcpy(out, m.data(), n); }
unsigned host_crc16(uint64_t v) { return crc16_u64(v); }
void host_hilbert(float *reco, float *im5) { auto c = hilbert_coeffs(kFilterLen, reco); std::memcpy(im5, c.data(), c.size() * 4); }
}
