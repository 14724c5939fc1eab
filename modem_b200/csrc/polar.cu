// modem_b200/csrc/polar.cu — batched CRC-aided polar successive-cancellation list decoding, N = 65536, L = 8.
//
// Replaces CODE::PolarListDecoder<SIMD<float,8>,16> + systematic() + the CRC-32 candidate scan of the reference
// receiver (/root/reference/decode.cc:201,530-555).  Design (B200-first, not a port of the SIMD recursion):
//   * one warp decodes four codewords in lock step along a precomputed op schedule (host_tables.cc: the frozen set is
//     fixed, so the tree walk is static and a warp never diverges on it).  The eight threads of a codeword play two roles:
//       - in the 32-leaf words (levels 4..0, registers) thread t IS list lane t: forks are ranked with 8-wide shuffles;
//       - above the words (levels 6..13, HBM/L2 scratch; level 5 in shared memory) the eight threads split the tree
//         POSITIONS of one path at a time (thread j takes quads j, j+8, ...), looping over the codeword's path classes.
//   * path classes: the reference starts its eight lanes as copies of one path (metrics 0, 1000, 1000, ...); copies stay
//     copies until a flipped decision displaces one, which on a clean channel never happens and on the README
//     impairment chain happens late.  Lanes holding the same path form a class, only the class representative (lowest
//     lane) owns alpha / beta storage and is computed; every lane keeps its own fp32 metric, so the result is the
//     reference's bit for bit.  Work and scratch traffic scale with the number of DISTINCT paths (1..8), not with L.
//   * rate-1 attempts (OP_R1 and inside the words): an all-free node whose forks provably cannot change the list —
//     lanes in metric order and largest metric < every lane's metric + min|alpha| — is decided by the signs of its alphas
//     in one pass; exact, because on a sign-following path the smallest |leaf LLR| of the node IS min|alpha| (the first
//     leaf's) and every other leaf's magnitude is a rounded sum that is not smaller.  Otherwise the node is walked as usual.
//   * levels 14..16 are never stored (TOP ops recompute level 13 from the lane-shared channel LLRs and a few beta bits).
//   * two code tables (modes 6..9 / 10..13): one op schedule each; the four codewords of a warp share a table.
//   * partial sums (beta) are bit-packed, 32 tree positions per word; at the root they ARE the re-encoded
//     codeword, whose non-frozen positions are the systematic message (decode.cc:254-261) — no message/map
//     trace-back is stored.
//   * fp32 operation order is identical to oracle/ref_code.hh PolarListDecoder (f = sign-min, g = b +- a,
//     rate-0 nodes summed in index order, forks ranked by (metric, 2*lane+bit)), so the result is bit-exact
//     against the oracle for the same LLRs.  tests/scl_emulator.cc is the scalar statement of this file's bookkeeping.
// No tensor cores: there is no dense contraction anywhere on this path.
#include "common.cuh"
#include "polar.cuh"

namespace ofdmrx {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float f_op(float a, float b)
{
	uint32_t s = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
	return __uint_as_float(__float_as_uint(fminf(fabsf(a), fabsf(b))) | s);
}
__device__ __forceinline__ float g_op(float a, float b, uint32_t bit)
{
	return __fadd_rn(b, __uint_as_float(__float_as_uint(a) ^ (bit << 31)));
}

struct SclCtx {
	float metric;
	int ret;          // lane map returned by the node that completed last: my path descends from lane `ret`
	int rep;          // representative (lowest lane) of my path class
	uint32_t W;       // partial sums of the current 32-leaf word
	uint32_t fmask;   // frozen mask of the current word
	int t, gbase;     // my list lane (0..7), warp lane of lane 0 of my codeword
};

// "no fork at the next free leaves can change the list": lanes in metric order and the largest metric below every lane's
// metric + mn (mn = the smallest |LLR| a free leaf can show on my path).  Warp-wide vote: the four codewords walk one op list.
__device__ __forceinline__ bool list_is_stable(const SclCtx &c, float mn)
{
	const float m7 = __shfl_sync(FULL, c.metric, c.gbase | 7);
	const float prev = __shfl_up_sync(FULL, c.metric, 1);
	const bool easy = m7 < __fadd_rn(c.metric, mn) && (c.t == 0 || prev <= c.metric);
	return __all_sync(FULL, easy);
}

struct ForkResult { float metric; int src_bit; int rep; };

// Free leaf: 2L forks, keep the L smallest by (metric, fork index); survivors land in rank order.  Survivors that extend
// the same class by the same bit are the same path: the new class representative is the lowest such lane.
// Not inlined: the 32-leaf word is fully unrolled around it and would otherwise carry 32 copies (I-cache).
__device__ __noinline__ ForkResult leaf_fork(float metric, float a, int t, int gbase, int rep)
{
	const float pen = fabsf(a);
	const float m0 = a < 0.f ? __fadd_rn(metric, pen) : metric; // decide 0
	const float m1 = a < 0.f ? metric : __fadd_rn(metric, pen); // decide 1
	float o0[8], o1[8];
	int r0 = 0, r1 = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		o0[j] = __shfl_sync(FULL, m0, gbase + j);
		o1[j] = __shfl_sync(FULL, m1, gbase + j);
		const bool lt = j < t, le = j <= t;
		r0 += (o0[j] < m0) || (o0[j] == m0 && lt);
		r0 += (o1[j] < m0) || (o1[j] == m0 && lt);
		r1 += (o0[j] < m1) || (o0[j] == m1 && le);
		r1 += (o1[j] < m1) || (o1[j] == m1 && lt);
	}
	const int packed = r0 | (r1 << 8);
	int sb = 0;
	float nm = 0.f;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const int pr = __shfl_sync(FULL, packed, gbase + j);
		if ((pr & 255) == t) { sb = j; nm = o0[j]; }
		if ((pr >> 8) == t) { sb = j | 8; nm = o1[j]; }
	}
	const int key = (__shfl_sync(FULL, rep, gbase + (sb & 7)) << 1) | (sb >> 3);
	int nrep = t;
#pragma unroll
	for (int j = 7; j >= 0; --j)
		if (__shfl_sync(FULL, key, gbase + j) == key) nrep = j;
	ForkResult r;
	r.metric = nm;
	r.src_bit = sb;
	r.rep = nrep;
	return r;
}

// ---- the 32-leaf word -------------------------------------------------------------------------------------------
// Code footprint matters more than instruction count here (the first fully unrolled version was 210 KB of SASS and
// spent 55 % of its stall cycles waiting for instruction fetch): the word is decoded by ONE non-inlined 8-leaf routine
// (levels 2..0 unrolled in registers, called four times) under ONE non-inlined 16-leaf routine (called twice).
struct Sub { float metric; uint32_t W; int ret; int rep; }; // result of a sub-tree: path metric, partial sums, lane map, class

// One node of an 8-leaf group, LVL = log2(size) <= 3, BASE = first leaf inside the group.  `a` = the node's alphas
// (registers); c.W / c.fmask hold the group's 8 local partial-sum / frozen bits.
template <int LVL, int BASE>
__device__ __forceinline__ void blk_node(SclCtx &c, const float *a)
{
	constexpr int N = 1 << LVL;
	constexpr uint32_t SUB = ((1u << N) - 1u) << BASE;
	if ((c.fmask & SUB) == SUB) { // rate-0 node (also the frozen leaf)
#pragma unroll
		for (int k = 0; k < N; ++k) {
			const float v = a[k];
			if (v < 0.f) c.metric = __fsub_rn(c.metric, v);
		}
		c.ret = c.t;
		return;
	}
	if constexpr (LVL == 0) {
		// Fast path (exact): if every "follow the sign" fork beats every "flip" fork and the lanes are already in
		// (metric, lane) order, the 8 survivors are the 8 keeps in place — no ranking, no permutation, classes unchanged.
		// (no 8-lane REDUX here: a __reduce_*_sync whose mask differs between the four codewords of the warp is compiled
		// into one serialised WARPSYNC.COLLECTIVE pass per group.)
		const float a0 = a[0];
		if (list_is_stable(c, fabsf(a0))) {
			c.ret = c.t;
			c.W |= (a0 < 0.f ? 1u : 0u) << BASE;
		} else {
			const ForkResult r = leaf_fork(c.metric, a0, c.t, c.gbase, c.rep);
			c.metric = r.metric;
			c.ret = r.src_bit & 7;
			c.rep = r.rep;
			c.W |= (uint32_t)(r.src_bit >> 3) << BASE;
		}
	} else {
		if ((c.fmask & SUB) == 0u) { // rate-1 attempt
			float mn = fabsf(a[0]);
#pragma unroll
			for (int k = 1; k < N; ++k) mn = fminf(mn, fabsf(a[k]));
			if (list_is_stable(c, mn)) {
#pragma unroll
				for (int k = 0; k < N; ++k) c.W |= (a[k] < 0.f ? 1u : 0u) << (BASE + k);
				c.ret = c.t;
				return;
			}
		}
		constexpr int H = N / 2;
		float ch[H];
#pragma unroll
		for (int k = 0; k < H; ++k) ch[k] = f_op(a[k], a[k + H]);
		blk_node<LVL - 1, BASE>(c, ch);
		const int lmap = c.ret;
		const int srcl = c.gbase + lmap;
#pragma unroll
		for (int k = 0; k < H; ++k) {
			const float pa = __shfl_sync(FULL, a[k], srcl), pb = __shfl_sync(FULL, a[k + H], srcl);
			ch[k] = g_op(pa, pb, (c.W >> (BASE + k)) & 1u);
		}
		blk_node<LVL - 1, BASE + H>(c, ch);
		constexpr uint32_t MASKL = ((1u << H) - 1u) << BASE;
		const int srcr = c.gbase + c.ret;
		const uint32_t Wl = __shfl_sync(FULL, c.W, srcr);
		c.W = (c.W & ~MASKL) | ((Wl ^ (c.W >> H)) & MASKL);
		c.ret = __shfl_sync(FULL, lmap, srcr);
	}
}

__device__ __noinline__ Sub leaf8(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
	uint32_t fmask8, float metric, int t, int gbase, int rep)
{
	SclCtx c;
	c.metric = metric; c.ret = t; c.rep = rep; c.W = 0; c.fmask = fmask8; c.t = t; c.gbase = gbase;
	const float a[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
	blk_node<3, 0>(c, a);
	Sub r;
	r.metric = c.metric; r.W = c.W; r.ret = c.ret; r.rep = c.rep;
	return r;
}

// 16 leaves: f -> left 8 -> g (parents through the lane map) -> right 8 -> combine
__device__ __noinline__ Sub node16(float4 x0, float4 x1, float4 x2, float4 x3, uint32_t fmask16, float metric, int t, int gbase, int rep)
{
	const float a[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
	Sub r;
	r.rep = rep;
	if (fmask16 == 0xffffu) {
#pragma unroll
		for (int k = 0; k < 16; ++k) if (a[k] < 0.f) metric = __fsub_rn(metric, a[k]);
		r.metric = metric; r.W = 0; r.ret = t;
		return r;
	}
	if (fmask16 == 0u) { // rate-1 attempt
		float mn = fabsf(a[0]);
#pragma unroll
		for (int k = 1; k < 16; ++k) mn = fminf(mn, fabsf(a[k]));
		SclCtx c;
		c.metric = metric; c.t = t; c.gbase = gbase;
		if (list_is_stable(c, mn)) {
			uint32_t W = 0;
#pragma unroll
			for (int k = 0; k < 16; ++k) W |= (a[k] < 0.f ? 1u : 0u) << k;
			r.metric = metric; r.W = W; r.ret = t;
			return r;
		}
	}
	float ch[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) ch[k] = f_op(a[k], a[k + 8]);
	const Sub l = leaf8(ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], ch[6], ch[7], fmask16 & 0xffu, metric, t, gbase, rep);
	const int srcl = gbase + l.ret;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const float pa = __shfl_sync(FULL, a[k], srcl), pb = __shfl_sync(FULL, a[k + 8], srcl);
		ch[k] = g_op(pa, pb, (l.W >> k) & 1u);
	}
	const Sub rr = leaf8(ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], ch[6], ch[7], fmask16 >> 8, l.metric, t, gbase, l.rep);
	const int srcr = gbase + rr.ret;
	const uint32_t Wl = __shfl_sync(FULL, l.W, srcr);
	r.metric = rr.metric;
	r.W = ((Wl ^ rr.W) & 0xffu) | (rr.W << 8);
	r.ret = __shfl_sync(FULL, l.ret, srcr);
	r.rep = rr.rep;
	return r;
}

// Alpha level 5 (the buffer every 32-leaf word reads twice and every F6/G6 writes) lives in shared memory: 4 KB per warp,
// [quad][codeword * 8 + slot] float4 with a row pitch of 33: the word's readers (8 lanes of a codeword, one quad, their
// slots) and the level-6 ops' writers (8 quads of one slot) both touch eight different 16-byte bank groups.
constexpr int kS5Pitch = 33;
constexpr int kS5Quads = 8 * kS5Pitch;
constexpr int kCStage = 2; // words per thread and round an OP_C stages in shared memory (2 KB per warp)

// the whole word: thread = list lane; `rs5` = the slot that holds my level-5 alphas (my class representative when they were written)
__device__ __forceinline__ void word32(SclCtx &c, const float4 *S5, int rs5)
{
	float4 v[8];
	const int own = c.gbase + rs5;
#pragma unroll
	for (int q = 0; q < 8; ++q) v[q] = S5[q * kS5Pitch + own];
	if (c.fmask == 0u) { // rate-1 attempt on the whole word
		float mn = fabsf(v[0].x);
#pragma unroll
		for (int q = 0; q < 8; ++q) mn = fminf(fminf(mn, fabsf(v[q].x)), fminf(fminf(fabsf(v[q].y), fabsf(v[q].z)), fabsf(v[q].w)));
		if (list_is_stable(c, mn)) {
			uint32_t W = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q)
				W |= ((v[q].x < 0.f ? 1u : 0u) | (v[q].y < 0.f ? 2u : 0u) | (v[q].z < 0.f ? 4u : 0u) | (v[q].w < 0.f ? 8u : 0u)) << (4 * q);
			c.W = W;
			c.ret = c.t;
			return;
		}
	}
	float4 x[4];
#pragma unroll
	for (int q = 0; q < 4; ++q) x[q] = make_float4(f_op(v[q].x, v[q + 4].x), f_op(v[q].y, v[q + 4].y), f_op(v[q].z, v[q + 4].z), f_op(v[q].w, v[q + 4].w));
	const Sub l = node16(x[0], x[1], x[2], x[3], c.fmask & 0xffffu, c.metric, c.t, c.gbase, c.rep);
	const int srcl = c.gbase + __shfl_sync(FULL, rs5, c.gbase + l.ret); // the slot of the lane I descend from
#pragma unroll
	for (int q = 0; q < 8; ++q) v[q] = S5[q * kS5Pitch + srcl];
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t wb = l.W >> (4 * q);
		x[q] = make_float4(g_op(v[q].x, v[q + 4].x, wb & 1u), g_op(v[q].y, v[q + 4].y, (wb >> 1) & 1u),
			g_op(v[q].z, v[q + 4].z, (wb >> 2) & 1u), g_op(v[q].w, v[q + 4].w, (wb >> 3) & 1u));
	}
	const Sub r = node16(x[0], x[1], x[2], x[3], c.fmask >> 16, l.metric, c.t, c.gbase, l.rep);
	const int srcr = c.gbase + r.ret;
	const uint32_t Wl = __shfl_sync(FULL, l.W, srcr);
	c.metric = r.metric;
	c.W = ((Wl ^ r.W) & 0xffffu) | (r.W << 16);
	c.ret = __shfl_sync(FULL, l.ret, srcr);
	c.rep = r.rep;
}

__device__ __forceinline__ float4 f_op4(float4 a, float4 b) { return make_float4(f_op(a.x, b.x), f_op(a.y, b.y), f_op(a.z, b.z), f_op(a.w, b.w)); }
__device__ __forceinline__ float4 g_op4(float4 a, float4 b, uint32_t bits)
{
	return make_float4(g_op(a.x, b.x, bits & 1u), g_op(a.y, b.y, (bits >> 1) & 1u), g_op(a.z, b.z, (bits >> 2) & 1u), g_op(a.w, b.w, (bits >> 3) & 1u));
}
__device__ __forceinline__ float r0_acc(float m, float4 v)
{
	if (v.x < 0.f) m = __fsub_rn(m, v.x);
	if (v.y < 0.f) m = __fsub_rn(m, v.y);
	if (v.z < 0.f) m = __fsub_rn(m, v.z);
	if (v.w < 0.f) m = __fsub_rn(m, v.w);
	return m;
}
__device__ __forceinline__ float min_abs4(float m, float4 v) { return fminf(fminf(m, fabsf(v.x)), fminf(fminf(fabsf(v.y), fabsf(v.z)), fabsf(v.w))); }
__device__ __forceinline__ uint32_t neg_bits4(float4 v) { return (v.x < 0.f ? 1u : 0u) | (v.y < 0.f ? 2u : 0u) | (v.z < 0.f ? 4u : 0u) | (v.w < 0.f ? 8u : 0u); }

// L2 residency: a warp's rows of the low levels are re-read within microseconds and fit the 126 MB L2 for all resident warps
// (clean frames: 15 KB per warp up to level 9); the channel values (read eight times per codeword, 2.6 GB per 10 000) and the
// rows of the top levels stream through once per use.  Loads and stores of levels >= kSclStreamLevel and all channel loads
// carry the evict-first / streaming hint so that they do not push the low levels out.
#ifndef OFDMRX_SCL_STREAM_LEVEL
#define OFDMRX_SCL_STREAM_LEVEL 11
#endif
constexpr int kSclStreamLevel = OFDMRX_SCL_STREAM_LEVEL; // 99 = no hints (A/B switch)
__device__ __forceinline__ float4 ld_row(const float4 *p, bool stream) { return stream ? __ldcs(p) : *p; }
__device__ __forceinline__ float4 ld_chan(const float4 *p) { return kSclStreamLevel < 99 ? __ldcs(p) : __ldg(p); }

// ---- storage of the levels above the words -------------------------------------------------------------------------
// level l (6..13) of codeword group `gbase` (= codeword * 8), slot s, quad q  ->  A[scl_off4(l) + ((gbase + s) << (l - 2)) + q]
__device__ __forceinline__ float4 *lvl_row(float4 *A, int l, int gslot) { return A + scl_off4(l) + ((size_t)gslot << (l - 2)); }
__device__ __forceinline__ void st_lvl(float4 *A, float4 *S5, int l, int gslot, int q, float4 v)
{
	if (l == 5) S5[q * kS5Pitch + gslot] = v;
	else if (l >= kSclStreamLevel) __stcs(&lvl_row(A, l, gslot)[q], v);
	else lvl_row(A, l, gslot)[q] = v;
}

// The big ops above the words are latency-bound streams (a warp holds four 128-bit loads per thread, the next turn's loads
// cannot be hoisted over this turn's stores): the lines of the turn after next are requested early — into L1 for the
// scratch rows (a few KB per warp), into L2 only for the TOP ops' channel values (8 KB per warp and turn).
#ifndef OFDMRX_SCL_PREFETCH
#define OFDMRX_SCL_PREFETCH 1 // bit 0: scratch rows into L1, bit 1: channel values into L2 (A/B switch)
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { if (OFDMRX_SCL_PREFETCH & 1) asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { if (OFDMRX_SCL_PREFETCH & 2) asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// 3-bit-per-level stacks (lane maps, slots) in one 64-bit register
__device__ __forceinline__ int stk_get(uint64_t s, int l) { return (int)((s >> (3 * l)) & 7ull); }
__device__ __forceinline__ uint64_t stk_set(uint64_t s, int l, int v) { return (s & ~(7ull << (3 * l))) | ((uint64_t)v << (3 * l)); }

// beta bits: [codeword * 8 + slot][2048 words] per warp — a path's partial sums are one contiguous 8 KB row, so the eight
// threads of a codeword read and write consecutive words of it
__device__ __forceinline__ uint32_t *beta_row(uint32_t *B, int gslot) { return B + ((size_t)gslot << 11); }

// F or G at level l, optionally fused with the F step that follows it down the left spine (host_tables.cc: depth field), for
// every path class of the codeword in turn.  The eight threads of the codeword split the positions: thread j takes the quad
// pairs q = j, j + 8, ... of the parent whose results meet again in the chained F step, so the intermediate level is
// produced in registers, written once (the later G needs it) and never re-read by an F.  Four 128-bit loads are in flight
// per thread and turn.  One instantiation serves F and G, chained or not (the kernel is short of instruction cache, not
// of issue slots).  `myps` = slot of the parent alphas of MY lane (meaningful on the representatives); iw = first beta word
// of the node.
__device__ __forceinline__ void fused_op(float4 *A, float4 *S5, const uint32_t *B, int iw, int l, bool is_g, bool chain, int gbase, int j, uint32_t repmask, int myps)
{
	const int hq = 1 << (l - 3), step = chain ? hq >> 1 : hq, second = chain ? step : 8;
	uint32_t m = repmask;
	while (__any_sync(FULL, m != 0u)) {
		const bool act = m != 0u;
		const int r = act ? __ffs(m) - 1 : 0;
		m &= m - 1u;
		const int ps = __shfl_sync(FULL, myps, gbase + r);
		if (act) {
			const float4 *P = lvl_row(A, l, gbase + ps);
			const uint32_t *Bw = beta_row(const_cast<uint32_t *>(B), gbase + r) + iw;
			// per turn two quad pairs: q and q + step when an F step is chained (its two operands), else q and q + 8
			const int stride = chain ? 8 : 16;
			for (int q0 = j; q0 < step; q0 += stride) {
				const int q1 = q0 + second;
				const bool two = chain || q1 < step;
				if (q0 + 2 * stride < step) { // (implies a second pair two turns ahead as well)
					const int qa = q0 + 2 * stride, qb = qa + second;
					prefetch_l1(&P[qa]); prefetch_l1(&P[qa + hq]); prefetch_l1(&P[qb]); prefetch_l1(&P[qb + hq]);
				}
				const bool strm = l >= kSclStreamLevel;
				const float4 pa0 = ld_row(&P[q0], strm), pb0 = ld_row(&P[q0 + hq], strm);
				float4 pa1 = pa0, pb1 = pb0;
				if (two) { pa1 = ld_row(&P[q1], strm); pb1 = ld_row(&P[q1 + hq], strm); }
				float4 v0, v1;
				if (is_g) {
					const uint32_t w0 = Bw[q0 >> 3], w1 = two ? Bw[q1 >> 3] : 0u;
					v0 = g_op4(pa0, pb0, (w0 >> (4 * j)) & 15u);
					v1 = g_op4(pa1, pb1, (w1 >> (4 * j)) & 15u);
				} else {
					v0 = f_op4(pa0, pb0);
					v1 = f_op4(pa1, pb1);
				}
				st_lvl(A, S5, l - 1, gbase + r, q0, v0);
				if (two) st_lvl(A, S5, l - 1, gbase + r, q1, v1);
				if (chain) st_lvl(A, S5, l - 2, gbase + r, q0, f_op4(v0, v1));
			}
		}
	}
}

// TOP(jn): alpha of the level-13 node jn (8192 positions) straight from the channel LLRs: three f/g steps per value whose
// operands are 8 lane-shared channel values and up to 7 partial-sum bits of the node's left-hand relatives at levels
// 15, 14 and 13 (read from the slots s15, s14, own of the lanes this path descended from).  Levels 16..14 are never stored
// — they were 2 x 2.1 MB of writes plus as much again in reads per codeword, all of it DRAM traffic.  The op also produces
// the left child at level 12 (the F step that always follows).
__device__ __forceinline__ void top_op(float4 *A, const float4 *C4, const uint32_t *B, int jn, int gbase, int j, uint32_t repmask, int mys14, int mys15)
{
	const bool j2 = jn & 4, j1 = jn & 2, j0 = jn & 1;
	uint32_t m = repmask;
	while (__any_sync(FULL, m != 0u)) {
		const bool act = m != 0u;
		const int r = act ? __ffs(m) - 1 : 0;
		m &= m - 1u;
		const int s14 = __shfl_sync(FULL, mys14, gbase + r), s15 = __shfl_sync(FULL, mys15, gbase + r);
		if (act) {
			const uint32_t *B15 = beta_row(const_cast<uint32_t *>(B), gbase + s15);                            // beta of node (15, 0): words 0..1023
			const uint32_t *B14 = beta_row(const_cast<uint32_t *>(B), gbase + s14) + (j2 ? 1024 : 0);          // beta of node (14, 2 j2)
			const uint32_t *B13 = beta_row(const_cast<uint32_t *>(B), gbase + r) + (j0 ? (jn - 1) * 256 : 0);  // beta of node (13, jn - 1)
			float4 *D13 = lvl_row(A, 13, gbase + r), *D12 = lvl_row(A, 12, gbase + r);
			const int sh = 4 * j;
#pragma unroll 1
			for (int q0 = j; q0 < 1024; q0 += 8) {
				float4 z[2];
				if (q0 + 16 < 1024) {
#pragma unroll
					for (int kk = 0; kk < 16; ++kk) prefetch_l2(&C4[q0 + 16 + 1024 * kk]);
				}
#pragma unroll
				for (int h = 0; h < 2; ++h) {
					const int q = q0 + 1024 * h, wq = q >> 3;
					float4 c[8];
#pragma unroll
					for (int kk = 0; kk < 8; ++kk) c[kk] = ld_chan(&C4[q + 2048 * kk]);
					float4 x[4], y[2];
					if (j2) {
#pragma unroll
						for (int mm = 0; mm < 4; ++mm) x[mm] = g_op4(c[mm], c[mm + 4], (B15[wq + 256 * mm] >> sh) & 15u);
					} else {
#pragma unroll
						for (int mm = 0; mm < 4; ++mm) x[mm] = f_op4(c[mm], c[mm + 4]);
					}
					if (j1) {
#pragma unroll
						for (int mm = 0; mm < 2; ++mm) y[mm] = g_op4(x[mm], x[mm + 2], (B14[wq + 256 * mm] >> sh) & 15u);
					} else {
#pragma unroll
						for (int mm = 0; mm < 2; ++mm) y[mm] = f_op4(x[mm], x[mm + 2]);
					}
					z[h] = j0 ? g_op4(y[0], y[1], (B13[wq] >> sh) & 15u) : f_op4(y[0], y[1]);
					if (13 >= kSclStreamLevel) __stcs(&D13[q], z[h]); else D13[q] = z[h];
				}
				if (12 >= kSclStreamLevel) __stcs(&D12[q0], f_op4(z[0], z[1])); else D12[q0] = f_op4(z[0], z[1]);
			}
		}
	}
}

__global__ void __launch_bounds__(kSclThreads, kSclCtasPerSm) k_polar_scl(SclParams p)
{
	const int lane32 = threadIdx.x & 31;
	const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	float4 *A = reinterpret_cast<float4 *>(p.A) + (size_t)warp_global * kSclWarpQuads;
	uint32_t *B = p.B + (size_t)warp_global * kSclWarpWords;
	__shared__ __align__(16) float4 S5[kS5Quads];
	__shared__ uint32_t crc_lut[256];
	__shared__ uint32_t SX[kCStage * 8 * 32]; // OP_C staging: [word of the thread][slot][thread]
	for (int i = lane32; i < 256; i += 32) { // CRC-32 0xD419CC15, reflected, one byte per step (decode.cc:198,534-537)
		uint32_t v = (uint32_t)i;
#pragma unroll
		for (int k = 0; k < 8; ++k) v = (v >> 1) ^ ((v & 1u) ? 0xD419CC15u : 0u);
		crc_lut[i] = v;
	}
	__syncwarp();
	SclCtx c;
	c.t = lane32 & 7;
	c.gbase = lane32 & ~7;
	const int j = c.t, gbase = c.gbase; // role above the words: position slice j of my codeword

	// groups of four codewords never mix code tables (the four walk one schedule in lock step): table 0's groups first
	const int n0 = p.n_cw_ptr ? p.n_cw_ptr[0] : p.n_cw[0], n1 = p.n_cw_ptr ? p.n_cw_ptr[1] : p.n_cw[1];
	const int g0 = (n0 + 3) >> 2, g1 = (n1 + 3) >> 2;
	for (;;) {
		int g4 = 0;
		if (lane32 == 0) g4 = atomicAdd(p.work, 1);
		g4 = __shfl_sync(FULL, g4, 0);
		if (g4 >= g0 + g1) break;
		const int tb = g4 < g0 ? 0 : 1;
		const int first = tb ? 4 * g0 : 0, n_cw = tb ? n1 : n0; // table 1's list starts at the next multiple of 4
		const int slot = (g4 - (tb ? g0 : 0)) * 4 + (lane32 >> 3);
		const bool active = slot < n_cw;
		const int pos = first + (active ? slot : n_cw - 1);
		const int frame = p.cw_list ? p.cw_list[pos] : pos;
		// (a ternary, not p.tbl[tb]: a run-time index into the kernel parameters makes the compiler copy them to local memory)
		const uint32_t *frozen = tb ? p.tbl[1] : p.tbl[0];
		const uint32_t *ops = frozen + kSclTblOps;
		const float *C = p.llr + (size_t)frame * kCodeLen;
		const float4 *C4 = reinterpret_cast<const float4 *>(C);
		c.metric = c.t == 0 ? 0.f : 1000.f;
		c.ret = c.t;
		c.rep = 0;                // eight copies of one path
		uint64_t lmstack = 0;     // [level]: lane map of the left child of the level's current node (kept by G / TOP for the C op)
		uint64_t rsstack = 0;     // [level]: my class representative when the level's data was written (alphas until the G, then
		                          // the left child's betas until the C): the slot to read them from

		for (int pc = 0;; ++pc) {
			const uint32_t opw = __ldg(&ops[pc]);
			const uint32_t op = opw & 7u, iw = (opw >> 8) & 0x3fffffu; // iw = first word of the node
			const int l = (int)((opw >> 3) & 31u);
			if (op == OP_END) break;
			const uint32_t repmask = (__ballot_sync(FULL, c.rep == c.t) >> gbase) & 255u; // my codeword's class representatives
			if (op == OP_F || op == OP_G) {
				const uint32_t depth = opw >> 30; // fused F steps that follow (0..1)
				int myps = stk_get(rsstack, l);
				if (op == OP_G) {
					lmstack = stk_set(lmstack, l, c.ret);
					myps = __shfl_sync(FULL, myps, gbase + c.ret); // the slot of the lane I was when this node's alphas were written
					rsstack = stk_set(rsstack, l, c.rep);          // from now on: where the left child's betas are (for the C op)
				}
				rsstack = stk_set(rsstack, l - 1, c.rep);
				if (depth) rsstack = stk_set(rsstack, l - 2, c.rep);
				fused_op(A, S5, B, (int)iw, l, op == OP_G, depth != 0u, gbase, j, repmask, myps); // chains of at most one F step (host_tables.h)
				__syncwarp();
			} else if (op == OP_TOP) {
				const int jn = (int)(iw >> 8); // node index at level 13
				if (jn > 0) { // the left-hand relative at level 13 + ctz(jn) has just completed: keep its lane map and slots (as a G would)
					const int lv = 14 + (__ffs(jn) - 1);
					lmstack = stk_set(lmstack, lv, c.ret);
					rsstack = stk_set(rsstack, lv, c.rep);
				}
				const int u = (jn & 1) ? stk_get(lmstack, 14) : c.t;                                        // my lane when the level-14 relative completed
				const int v = (jn & 2) ? __shfl_sync(FULL, stk_get(lmstack, 15), gbase + u) : u;              // ... and when node (15, 0) completed
				const int mys14 = __shfl_sync(FULL, stk_get(rsstack, 15), gbase + u);
				const int mys15 = __shfl_sync(FULL, stk_get(rsstack, 16), gbase + v);
				rsstack = stk_set(stk_set(rsstack, 13, c.rep), 12, c.rep);
				top_op(A, C4, B, jn, gbase, j, repmask, mys14, mys15); // always chained with the F step below (host_tables.cc)
				__syncwarp();
			} else if (op == OP_WORD) {
				c.fmask = __ldg(&frozen[iw]);
				c.W = 0;
				word32(c, S5, stk_get(rsstack, 5));
				if (c.rep == c.t) beta_row(B, lane32)[iw] = c.W;
				__syncwarp();
			} else if (op == OP_R0) {
				// thread = list lane again: the penalties are summed in index order into every lane's own metric
				const int nq = 1 << (l - 2);
				const int gslot = gbase + stk_get(rsstack, l);
				float mt = c.metric;
				if (l == 5) {
#pragma unroll
					for (int q = 0; q < 8; ++q) mt = r0_acc(mt, S5[q * kS5Pitch + gslot]);
				} else {
					const float4 *P = lvl_row(A, l, gslot);
					for (int q0 = 0; q0 < nq; q0 += 8) {
						float4 v[8];
#pragma unroll
						for (int uu = 0; uu < 8; ++uu) v[uu] = P[q0 + uu];
#pragma unroll
						for (int uu = 0; uu < 8; ++uu) mt = r0_acc(mt, v[uu]);
					}
				}
				c.metric = mt;
				if (c.rep == c.t) {
					uint32_t *Bz = beta_row(B, lane32) + iw;
					for (int w = 0; w < nq / 8; ++w) Bz[w] = 0u;
				}
				c.ret = c.t;
				__syncwarp();
			} else if (op == OP_R1) {
				// rate-1 attempt: per class one pass over the node's alphas (thread j takes the words j, j + 8, ...): minimum
				// magnitude and sign bits; the betas are stored before the verdict (nobody reads this node's words before they
				// are decoded, and a failed attempt is followed by the ops that decode them)
				const int nw = 1 << (l - 5);
				float mymin = 0.f;
				uint32_t m = repmask;
				while (__any_sync(FULL, m != 0u)) {
					const bool act = m != 0u;
					const int r = act ? __ffs(m) - 1 : 0;
					m &= m - 1u;
					float mn = __int_as_float(0x7f800000);
					if (act) {
						const float4 *P = lvl_row(A, l, gbase + r);
						for (int w = j; w < nw; w += 8) {
							float4 v[8];
#pragma unroll
							for (int q = 0; q < 8; ++q) v[q] = P[w * 8 + q];
							uint32_t x = 0;
#pragma unroll
							for (int q = 0; q < 8; ++q) { mn = min_abs4(mn, v[q]); x |= neg_bits4(v[q]) << (4 * q); }
							beta_row(B, gbase + r)[iw + w] = x;
						}
					}
#pragma unroll
					for (int d = 1; d < 8; d <<= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, d));
					if (act && c.rep == r) mymin = mn;
				}
				if (list_is_stable(c, mymin)) {
					c.ret = c.t;
					pc = (int)__ldg(&ops[pc + 1]) - 1;
				} else {
					++pc;
				}
				__syncwarp();
			} else { // OP_C
				const int hw = 1 << (l - 6);
				const int myL = __shfl_sync(FULL, stk_get(rsstack, l), gbase + c.ret); // the slot that holds the left half's betas of my path
				// every class: left half <- (left half of the slot it descends from) ^ (its right half).  A class may read a slot
				// that another class of the codeword overwrites in the same op: all reads of a word come before its writes
				// (staged in shared memory, one column per thread; thread j takes the words j, j + 8, ...).
				for (int wb = 0; wb < hw; wb += 8 * kCStage) { // (warp-uniform trip count: the class loop below votes)
					const int w0 = wb + j;
					uint32_t m = repmask;
					while (__any_sync(FULL, m != 0u)) {
						const bool act = m != 0u;
						const int r = act ? __ffs(m) - 1 : 0;
						m &= m - 1u;
						const int Lr = __shfl_sync(FULL, myL, gbase + r);
						if (act) {
							const uint32_t *Bl = beta_row(B, gbase + Lr) + iw, *Br = beta_row(B, gbase + r) + iw + hw;
#pragma unroll
							for (int k = 0; k < kCStage; ++k)
								if (w0 + 8 * k < hw) SX[(k * 8 + r) * 32 + lane32] = Bl[w0 + 8 * k] ^ Br[w0 + 8 * k];
						}
					}
					m = repmask;
					while (m) { // (no collective inside: this loop may run differently long in the four codewords)
						const int r = __ffs(m) - 1;
						m &= m - 1u;
						uint32_t *Bl = beta_row(B, gbase + r) + iw;
#pragma unroll
						for (int k = 0; k < kCStage; ++k)
							if (w0 + 8 * k < hw) Bl[w0 + 8 * k] = SX[(k * 8 + r) * 32 + lane32];
					}
				}
				__syncwarp();
				c.ret = __shfl_sync(FULL, stk_get(lmstack, l), gbase + c.ret);
			}
		}

		// ---- candidate order, CRC-32 (decode.cc:532-541), payload (decode.cc:546-554) ------------------------
		int rank = 0;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const float mk = __shfl_sync(FULL, c.metric, gbase + k);
			rank += (mk < c.metric) || (mk == c.metric && k < c.t);
		}
		// CRC-32 of every distinct path, once: the eight threads of the codeword take one eighth of the message bits each
		// (word ranges from the table), and the pieces are joined through the CRC's linearity — the register after piece j,
		// advanced over the bits that follow it (a 32 x 32 bit matrix per piece, host_tables.cc), XORed over the pieces.
		uint32_t crc = 0;
		{
			const uint32_t *ctab = frozen + kSclTblCrc;
			const int w_begin = (int)__ldg(&ctab[j]), w_end = (int)__ldg(&ctab[j + 1]);
			uint32_t m = (__ballot_sync(FULL, c.rep == c.t) >> gbase) & 255u;
			while (__any_sync(FULL, m != 0u)) {
				const bool act = m != 0u;
				const int r = act ? __ffs(m) - 1 : 0;
				m &= m - 1u;
				uint32_t part = 0;
				if (act) {
					const uint32_t *Brow = beta_row(B, gbase + r);
					uint32_t reg = 0;
					uint64_t acc = 0;
					int nacc = 0;
					for (int w = w_begin; w < w_end; ++w) {
						uint32_t fr = ~__ldg(&frozen[w]);
						if (!fr) continue;
						const int base = (int)__ldg(&frozen[kSclTblMsgOff + w]);
						const uint32_t x = Brow[w];
						uint32_t bits = x;
						int n = 32;
						if (fr != 0xffffffffu) {
							bits = 0; n = 0;
							while (fr) {
								const int b = __ffs(fr) - 1;
								fr &= fr - 1;
								bits |= ((x >> b) & 1u) << n;
								++n;
							}
						}
						if (base + n > kCrcBits) { n = kCrcBits - base; bits &= (1u << n) - 1u; } // (only in the last piece; n > 0 there)
						acc |= (uint64_t)bits << nacc;
						nacc += n;
						while (nacc >= 8) {
							reg = (reg >> 8) ^ crc_lut[(reg ^ (uint32_t)acc) & 255u];
							acc >>= 8;
							nacc -= 8;
						}
					}
					for (; nacc > 0; --nacc, acc >>= 1) reg = (reg >> 1) ^ (((reg ^ (uint32_t)acc) & 1u) ? 0xD419CC15u : 0u);
					const uint32_t *M = ctab + 16 + 32 * j;
					while (reg) {
						const int b = __ffs(reg) - 1;
						reg &= reg - 1u;
						part ^= __ldg(&M[b]);
					}
				}
#pragma unroll
				for (int d = 1; d < 8; d <<= 1) part ^= __shfl_xor_sync(FULL, part, d);
				if (act && c.rep == r) crc = part;
			}
		}
		const uint32_t *Bmine = beta_row(B, gbase + c.rep); // the root's betas = my path's re-encoded codeword
		const bool ok = crc == 0u;
		int key = ok ? rank : 64;
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) key = min(key, __shfl_xor_sync(FULL, key, d));
		const unsigned bal = __ballot_sync(FULL, ok && rank == key);
		const int win = __ffs((bal >> gbase) & 0xffu) - 1; // -1: no candidate passes the CRC
		const int wslot = __shfl_sync(FULL, c.rep, gbase + (win < 0 ? 0 : win));
		int flips = 0;
		if (active) {
			FrameState &st = p.st[frame];
			st.metrics[rank] = c.metric;
			if (p.xbits)
				for (int w = 0; w < kCodeLen / 32; ++w)
					p.xbits[((size_t)(first + slot) * 8 + rank) * (kCodeLen / 32) + w] = Bmine[w];
			if (win >= 0) {
				uint32_t *out = p.payload + (size_t)frame * (kDataBytes / 4);
				const uint32_t *Bwin = beta_row(B, gbase + wslot);
				for (int w = c.t; w < kCodeLen / 32; w += 8) {
					const int base = (int)__ldg(&frozen[kSclTblMsgOff + w]);
					if (base >= kDataBits) break;
					uint32_t fr = ~__ldg(&frozen[w]);
					if (!fr) continue;
					const uint32_t x = Bwin[w];
					uint32_t neg = 0; // channel hard decisions of the word (decode.cc:549-552)
#pragma unroll
					for (int q = 0; q < 8; ++q) neg |= neg_bits4(ld_chan(&C4[w * 8 + q])) << (4 * q);
					uint64_t mbits = x;
					int k = 32;
					if (fr != 0xffffffffu || base + 32 > kDataBits) {
						mbits = 0; k = 0;
						while (fr && base + k < kDataBits) {
							const int b = __ffs(fr) - 1;
							fr &= fr - 1;
							const uint32_t bit = (x >> b) & 1u;
							mbits |= (uint64_t)bit << k;
							flips += (int)(((neg >> b) & 1u) != bit);
							++k;
						}
					} else {
						flips += __popc(neg ^ x);
					}
					const uint64_t sh = mbits << (base & 31);
					const int wi = base >> 5;
					if ((uint32_t)sh) atomicXor(&out[wi], (uint32_t)sh);
					if ((uint32_t)(sh >> 32) && wi + 1 < kDataBytes / 4) atomicXor(&out[wi + 1], (uint32_t)(sh >> 32));
				}
			}
		}
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) flips += __shfl_xor_sync(FULL, flips, d);
		if (active && c.t == 0) {
			FrameState &st = p.st[frame];
			st.best_lane = win;
			st.flips = win >= 0 ? flips : -1;
			st.status = win >= 0 ? ST_OK : ST_PAYLOAD_CRC;
		}
		__syncwarp();
	}
}

// payload buffer <- scrambler sequence (decode.cc:613-615: out ^= xorshift); the decoder XORs the message in.  Also resets
// the list decoder's work counter.
__global__ void k_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames, int *work)
{
	const int per = kDataBytes / 4;
	if (work && blockIdx.x == 0 && threadIdx.x == 0) *work = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_frames * per; i += (size_t)gridDim.x * blockDim.x)
		payload[i] = scr_words[i % per];
}

} // namespace

int scl_resident_warps(int ctas_per_sm, int n_sm) { return ctas_per_sm * n_sm * (kSclThreads / 32); }

cudaError_t launch_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames, int *work, cudaStream_t s)
{
	k_payload_init<<<n_frames > 0 ? 592 : 1, 256, 0, s>>>(payload, scr_words, n_frames, work);
	return cudaGetLastError();
}

cudaError_t launch_polar_scl(const SclParams &p, int grid, cudaStream_t s)
{
	if (!p.n_cw_ptr && p.n_cw[0] + p.n_cw[1] <= 0) return cudaSuccess;
	k_polar_scl<<<grid, kSclThreads, 0, s>>>(p);
	return cudaGetLastError();
}

int scl_occupancy_ctas_per_sm()
{
	int n = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_polar_scl, kSclThreads, 0);
	return n;
}

} // namespace ofdmrx
