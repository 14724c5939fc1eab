// modem_b200/csrc/polar.cu — batched CRC-aided polar successive-cancellation list decoding, N = 65536, L = 8.
//
// Replaces CODE::PolarListDecoder<SIMD<float,8>,16> + systematic() + the CRC-32 candidate scan of the reference
// receiver (/root/reference/decode.cc:201,530-555).  Design (B200-first, not a port of the SIMD recursion):
//   * one list lane per THREAD, one codeword per 8 threads, four codewords per warp.  Every codeword walks the
//     same precomputed op schedule (host_tables.cc: the frozen set is fixed), so a warp never diverges and the
//     only cross-thread traffic is 8-wide shuffles (lane permutation after a fork, fork ranking).
//   * alpha (LLR) buffers of tree levels 6..13 live in a per-warp HBM/L2 scratch laid out [quad][warp lane][4]
//     (a quad = 4 consecutive tree positions) so that every warp access is 512 coalesced bytes of 128-bit loads,
//     also when a thread reads through the lane map; level 5 lives in shared memory; levels 14..16 are never stored
//     (TOP ops recompute level 13 from the lane-shared channel LLRs); levels 0..4 (a 32-leaf "word") are fully
//     unrolled and live in registers.
//   * two code tables (modes 6..9 / 10..13): one op schedule each; the four codewords of a warp share a table.
//   * partial sums (beta) are bit-packed, 32 tree positions per word; at the root they ARE the re-encoded
//     codeword, whose non-frozen positions are the systematic message (decode.cc:254-261) — no message/map
//     trace-back is stored.
//   * fp32 operation order is identical to oracle/ref_code.hh PolarListDecoder (f = sign-min, g = b +- a,
//     rate-0 nodes summed in index order, forks ranked by (metric, 2*lane+bit)), so the result is bit-exact
//     against the oracle for the same LLRs.
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
	uint32_t W;       // partial sums of the current 32-leaf word
	uint32_t fmask;   // frozen mask of the current word
	int t, gbase;     // my list lane (0..7), warp lane of lane 0 of my codeword
	const float4 *A5; // level-5 alpha buffer of this warp ([8 quads][32 lanes])
};

struct ForkResult { float metric; int src_bit; };

// Free leaf: 2L forks, keep the L smallest by (metric, fork index); survivors land in rank order.
// Not inlined: the 32-leaf word is fully unrolled around it and would otherwise carry 32 copies (I-cache).
__device__ __noinline__ ForkResult leaf_fork(float metric, float a, int t, int gbase)
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
	ForkResult r;
	r.metric = nm;
	r.src_bit = sb;
	return r;
}

// ---- the 32-leaf word -------------------------------------------------------------------------------------------
// Code footprint matters more than instruction count here (the first fully unrolled version was 210 KB of SASS and
// spent 55 % of its stall cycles waiting for instruction fetch): the word is decoded by ONE non-inlined 8-leaf routine
// (levels 2..0 unrolled in registers, called four times) under ONE non-inlined 16-leaf routine (called twice).
struct Sub { float metric; uint32_t W; int ret; int pad; }; // result of a sub-tree: path metric, partial sums, lane map

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
		// (metric, lane) order, the 8 survivors are the 8 keeps in place — no ranking, no permutation.  Metrics are
		// non-negative, so their bit patterns order like unsigned integers (REDUX instead of shuffle trees).
		const float a0 = a[0];
		// (no 8-lane REDUX here: a __reduce_*_sync whose mask differs between the four codewords of the warp is compiled
		// into one serialised WARPSYNC.COLLECTIVE pass per group.)  If the lanes are in order, the largest keep metric is
		// lane 7's, and "it is below every flip metric" can be tested per lane.
		const float m7 = __shfl_sync(FULL, c.metric, c.gbase | 7);
		const float prev = __shfl_up_sync(FULL, c.metric, 1);
		const bool easy = m7 < __fadd_rn(c.metric, fabsf(a0)) && (c.t == 0 || prev <= c.metric);
		if (__all_sync(FULL, easy)) {
			c.ret = c.t;
			c.W |= (a0 < 0.f ? 1u : 0u) << BASE;
		} else {
			const ForkResult r = leaf_fork(c.metric, a0, c.t, c.gbase);
			c.metric = r.metric;
			c.ret = r.src_bit & 7;
			c.W |= (uint32_t)(r.src_bit >> 3) << BASE;
		}
	} else {
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
	uint32_t fmask8, float metric, int t, int gbase)
{
	SclCtx c;
	c.metric = metric; c.ret = t; c.W = 0; c.fmask = fmask8; c.t = t; c.gbase = gbase; c.A5 = nullptr;
	const float a[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
	blk_node<3, 0>(c, a);
	Sub r;
	r.metric = c.metric; r.W = c.W; r.ret = c.ret; r.pad = 0;
	return r;
}

// 16 leaves: f -> left 8 -> g (parents through the lane map) -> right 8 -> combine
__device__ __noinline__ Sub node16(float4 x0, float4 x1, float4 x2, float4 x3, uint32_t fmask16, float metric, int t, int gbase)
{
	const float a[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
	Sub r;
	r.pad = 0;
	if (fmask16 == 0xffffu) {
#pragma unroll
		for (int k = 0; k < 16; ++k) if (a[k] < 0.f) metric = __fsub_rn(metric, a[k]);
		r.metric = metric; r.W = 0; r.ret = t;
		return r;
	}
	float ch[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) ch[k] = f_op(a[k], a[k + 8]);
	const Sub l = leaf8(ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], ch[6], ch[7], fmask16 & 0xffu, metric, t, gbase);
	const int srcl = gbase + l.ret;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const float pa = __shfl_sync(FULL, a[k], srcl), pb = __shfl_sync(FULL, a[k + 8], srcl);
		ch[k] = g_op(pa, pb, (l.W >> k) & 1u);
	}
	const Sub rr = leaf8(ch[0], ch[1], ch[2], ch[3], ch[4], ch[5], ch[6], ch[7], fmask16 >> 8, l.metric, t, gbase);
	const int srcr = gbase + rr.ret;
	const uint32_t Wl = __shfl_sync(FULL, l.W, srcr);
	r.metric = rr.metric;
	r.W = ((Wl ^ rr.W) & 0xffu) | (rr.W << 8);
	r.ret = __shfl_sync(FULL, l.ret, srcr);
	return r;
}

// the whole word: alphas of the level-5 node are in the scratch buffer (8 quads per lane)
__device__ __forceinline__ void word32(SclCtx &c)
{
	float4 v[8];
	const int own = c.gbase + c.t;
#pragma unroll
	for (int q = 0; q < 8; ++q) v[q] = c.A5[q * 32 + own];
	float4 x[4];
#pragma unroll
	for (int q = 0; q < 4; ++q) x[q] = make_float4(f_op(v[q].x, v[q + 4].x), f_op(v[q].y, v[q + 4].y), f_op(v[q].z, v[q + 4].z), f_op(v[q].w, v[q + 4].w));
	const Sub l = node16(x[0], x[1], x[2], x[3], c.fmask & 0xffffu, c.metric, c.t, c.gbase);
	const int srcl = c.gbase + l.ret;
#pragma unroll
	for (int q = 0; q < 8; ++q) v[q] = c.A5[q * 32 + srcl];
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t wb = l.W >> (4 * q);
		x[q] = make_float4(g_op(v[q].x, v[q + 4].x, wb & 1u), g_op(v[q].y, v[q + 4].y, (wb >> 1) & 1u),
			g_op(v[q].z, v[q + 4].z, (wb >> 2) & 1u), g_op(v[q].w, v[q + 4].w, (wb >> 3) & 1u));
	}
	const Sub r = node16(x[0], x[1], x[2], x[3], c.fmask >> 16, l.metric, c.t, c.gbase);
	const int srcr = c.gbase + r.ret;
	const uint32_t Wl = __shfl_sync(FULL, l.W, srcr);
	c.metric = r.metric;
	c.W = ((Wl ^ r.W) & 0xffffu) | (r.W << 16);
	c.ret = __shfl_sync(FULL, l.ret, srcr);
}

// L2 cache policy for the stores of the TOP ops (levels 13 and 12 stream through: evict-first keeps them from pushing
// the small, frequently re-read levels out of the 126 MB L2); the policy is a runtime operand.  The F/G ops use plain
// generic accesses (their lowest level lives in shared memory; hints measured no gain there).
__device__ __forceinline__ uint64_t l2_policy(bool stream)
{
	uint64_t p;
	if (stream) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ void st_pol(float4 *ptr, float4 v, uint64_t pol)
{
	asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
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

// Alpha level 5 (the buffer every 32-leaf word reads twice and every F6/G6 writes) lives in shared memory: 4 KB per warp,
// [quad][lane] float4 like the global levels; accesses below use generic addressing.  Measured per 10 000 codewords:
// none 70.2 ms, level 5 68.2 ms, levels 5 and 6 85.7 ms (204 KB of shared memory leave almost no L1 for the rest).
#ifndef OFDMRX_SCL_SMEM_LEVELS
#define OFDMRX_SCL_SMEM_LEVELS 1
#endif
constexpr int kSclSmemLevels = OFDMRX_SCL_SMEM_LEVELS; // 0: none, 1: level 5, 2: levels 5 and 6
constexpr int kSclSmemBytes = (kSclSmemLevels >= 1 ? 8 : 0) * 32 * 16 + (kSclSmemLevels >= 2 ? 16 : 0) * 32 * 16;
__device__ __forceinline__ float4 *lvl_ptr(float4 *A, float4 *S, int l)
{
	if (kSclSmemLevels >= 1 && l == 5) return S;
	if (kSclSmemLevels >= 2 && l == 6) return S + 8 * 32;
	return A + scl_off4(l);
}

#ifndef OFDMRX_SCL_PAIRS
#define OFDMRX_SCL_PAIRS 4
#endif
constexpr int kSclPairsInFlight = OFDMRX_SCL_PAIRS; // quad pairs loaded per thread before the first store (x2 128-bit loads in flight)

// F or G at level l fused with the D-1 F steps that follow it down the left spine (host_tables.cc: depth field).
// One iteration takes the 2^(D-1) quad pairs of the parent whose results meet again in the chained F steps, so the
// intermediate levels are produced in registers, written once (the later G needs them) and never re-read by an F.
// 8 x 128-bit loads are in flight per thread for every D (D <= kSclMaxFuse = 2: deeper chains cost registers and
// instruction-cache footprint and measured slower).
template <int D, bool IS_G>
__device__ __forceinline__ void fused_op(float4 *A, float4 *S, const float4 *C4, const uint32_t *Bw, int l, int src, int lane32)
{
	constexpr int M = 1 << (D - 1), U = kSclPairsInFlight / M;
	const int hq = 1 << (l - 3), step = hq >> (D - 1);
	const float4 *P = lvl_ptr(A, S, l > 15 ? 15 : l);
	float4 *D1 = lvl_ptr(A, S, l - 1);
	float4 *D2 = lvl_ptr(A, S, D >= 2 ? l - 2 : l - 1);
	const bool root = l == 16;
	for (int q0 = 0; q0 < step; q0 += 8) {
		uint32_t bw[M];
		if constexpr (IS_G) {
#pragma unroll
			for (int m = 0; m < M; ++m) bw[m] = Bw[((q0 + m * step) >> 3) * 32];
		}
#pragma unroll 1
		for (int k = 0; k < 8; k += U) {
			float4 pa[U][M], pb[U][M];
#pragma unroll
			for (int u = 0; u < U; ++u)
#pragma unroll
				for (int m = 0; m < M; ++m) {
					const int q = q0 + k + u + m * step;
					if (root) { pa[u][m] = __ldg(&C4[q]); pb[u][m] = __ldg(&C4[q + hq]); }
					else { pa[u][m] = P[q * 32 + src]; pb[u][m] = P[(q + hq) * 32 + src]; }
				}
#pragma unroll
			for (int u = 0; u < U; ++u) {
				float4 v1[M];
#pragma unroll
				for (int m = 0; m < M; ++m) {
					const int q = q0 + k + u + m * step;
					if constexpr (IS_G) v1[m] = g_op4(pa[u][m], pb[u][m], (bw[m] >> (4 * (k + u))) & 15u);
					else v1[m] = f_op4(pa[u][m], pb[u][m]);
					D1[q * 32 + lane32] = v1[m];
				}
				if constexpr (D >= 2) {
					float4 v2[M / 2];
#pragma unroll
					for (int m = 0; m < M / 2; ++m) {
						v2[m] = f_op4(v1[m], v1[m + M / 2]);
						D2[(q0 + k + u + m * step) * 32 + lane32] = v2[m];
					}
				}
			}
		}
	}
}

// TOP(j): alpha of the level-13 node j (8192 positions) straight from the channel LLRs: three f/g steps per value whose
// operands are 8 lane-shared channel values and up to 7 partial-sum bits of the node's left-hand relatives at levels
// 15, 14 and 13 (read from the lanes this path descends from: s15, s14, own).  Levels 16..14 are never stored — they
// were 2 x 2.1 MB of writes plus as much again in reads per codeword, all of it DRAM traffic.  D = 2 also produces
// the left child at level 12 (the F step that always follows).
template <int D>
__device__ __forceinline__ void top_op(float4 *A, const float4 *C4, const uint32_t *B, int j, int s15, int s14, int lane32, int stream_level)
{
	const bool j2 = j & 4, j1 = j & 2, j0 = j & 1;
	const uint32_t *B15 = B + s15;                                           // beta of node (15, 0): words 0..1023
	const uint32_t *B14 = B + (size_t)(j2 ? 1024 : 0) * 32 + s14;            // beta of node (14, 2 j2)
	const uint32_t *B13 = B + (size_t)(j0 ? (j - 1) * 256 : 0) * 32 + lane32; // beta of node (13, j - 1)
	float4 *D13 = A + scl_off4(13), *D12 = A + scl_off4(12);
	const uint64_t p13 = l2_policy(13 >= stream_level), p12 = l2_policy(12 >= stream_level);
	constexpr int NQ = D == 2 ? 1024 : 2048;
	for (int q0 = 0; q0 < NQ; q0 += 8) {
		uint32_t w15[D][4], w14[D][2], w13[D];
#pragma unroll
		for (int h = 0; h < D; ++h) {
			const int wq = (q0 + 1024 * h) >> 3;
#pragma unroll
			for (int m = 0; m < 4; ++m) w15[h][m] = j2 ? B15[(size_t)(wq + 256 * m) * 32] : 0u;
#pragma unroll
			for (int m = 0; m < 2; ++m) w14[h][m] = j1 ? B14[(size_t)(wq + 256 * m) * 32] : 0u;
			w13[h] = j0 ? B13[(size_t)wq * 32] : 0u;
		}
#pragma unroll 1
		for (int k = 0; k < 8; ++k) {
			const int sh = 4 * k;
			float4 z[D];
#pragma unroll
			for (int h = 0; h < D; ++h) {
				float4 c[8];
#pragma unroll
				for (int kk = 0; kk < 8; ++kk) c[kk] = __ldg(&C4[q0 + k + 1024 * h + 2048 * kk]);
				float4 x[4], y[2];
				if (j2) {
#pragma unroll
					for (int m = 0; m < 4; ++m) x[m] = g_op4(c[m], c[m + 4], (w15[h][m] >> sh) & 15u);
				} else {
#pragma unroll
					for (int m = 0; m < 4; ++m) x[m] = f_op4(c[m], c[m + 4]);
				}
				if (j1) {
#pragma unroll
					for (int m = 0; m < 2; ++m) y[m] = g_op4(x[m], x[m + 2], (w14[h][m] >> sh) & 15u);
				} else {
#pragma unroll
					for (int m = 0; m < 2; ++m) y[m] = f_op4(x[m], x[m + 2]);
				}
				z[h] = j0 ? g_op4(y[0], y[1], (w13[h] >> sh) & 15u) : f_op4(y[0], y[1]);
				st_pol(&D13[(size_t)(q0 + k + 1024 * h) * 32 + lane32], z[h], p13);
			}
			if constexpr (D == 2) st_pol(&D12[(size_t)(q0 + k) * 32 + lane32], f_op4(z[0], z[1]), p12);
		}
	}
}

// Upper-level ops work on quads (float4 = 4 consecutive tree positions of one lane); loads of a batch of U quads are
// issued before anything is stored so that U*2 128-bit loads are in flight per thread (the stores may alias the loads
// as far as the compiler knows, so the batching has to be explicit).
constexpr int kU = 4;

__global__ void __launch_bounds__(kSclThreads, kSclCtasPerSm) k_polar_scl(SclParams p)
{
	const int lane32 = threadIdx.x & 31;
	const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int n_warps = (gridDim.x * blockDim.x) >> 5;
	float4 *A = reinterpret_cast<float4 *>(p.A + (size_t)warp_global * p.a_stride);
	uint32_t *B = p.B + (size_t)warp_global * kSclWarpWords;
	SclCtx c;
	c.t = lane32 & 7;
	c.gbase = lane32 & ~7;
	extern __shared__ __align__(16) float4 scl_smem[];
	float4 *S = scl_smem;
	c.A5 = lvl_ptr(A, S, 5);

	// groups of four codewords never mix code tables (the four walk one schedule in lock step): table 0's groups first
	const int n0 = p.n_cw_ptr ? p.n_cw_ptr[0] : p.n_cw[0], n1 = p.n_cw_ptr ? p.n_cw_ptr[1] : p.n_cw[1];
	const int g0 = (n0 + 3) >> 2, g1 = (n1 + 3) >> 2;
	for (int g4 = warp_global; g4 < g0 + g1; g4 += n_warps) {
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
		uint64_t lmstack = 0;

		for (int pc = 0;; ++pc) {
			const uint32_t opw = __ldg(&ops[pc]);
			const uint32_t op = opw & 7u, l = (opw >> 3) & 31u, iw = (opw >> 8) & 0x3fffffu; // iw = first word of the node
			if (op == OP_END) break;
			const int hq = 1 << (l - 3);               // quads per half node
			const float4 *P = lvl_ptr(A, S, l > 15 ? 15 : (int)l); // parent level (valid for l <= 15)
			if (op == OP_F || op == OP_G) {
				const uint32_t depth = opw >> 30; // fused F steps that follow (0..2)
				if (op == OP_G) lmstack = (lmstack & ~(7ull << (3 * l))) | ((uint64_t)c.ret << (3 * l));
				const int src = op == OP_G ? c.gbase + c.ret : lane32;
				const uint32_t *Bw = B + (size_t)iw * 32 + lane32;
				// chains of at most kSclMaxFuse - 1 F steps (host_tables.h); deeper fusion measured slower (registers) and
				// every instantiation costs instruction-cache footprint, which this kernel is short of
				if (op == OP_F) {
					if (depth == 1) fused_op<2, false>(A, S, C4, Bw, l, src, lane32);
					else fused_op<1, false>(A, S, C4, Bw, l, src, lane32);
				} else {
					if (depth == 1) fused_op<2, true>(A, S, C4, Bw, l, src, lane32);
					else fused_op<1, true>(A, S, C4, Bw, l, src, lane32);
				}
				__syncwarp();
			} else if (op == OP_TOP) {
				const int j = (int)(iw >> 8); // node index at level 13
				if (j > 0) { // the left-hand relative at level 13 + ctz(j) has just completed: keep its lane map (as a G would)
					const int lv = 14 + (__ffs(j) - 1);
					lmstack = (lmstack & ~(7ull << (3 * lv))) | ((uint64_t)c.ret << (3 * lv));
				}
				const int lm14 = (int)((lmstack >> 42) & 7ull), lm15 = (int)((lmstack >> 45) & 7ull);
				const int u = (j & 1) ? lm14 : c.t;                  // my lane when node (14, 2 j2) completed
				const int a15 = __shfl_sync(FULL, lm15, c.gbase + u); // ... and when node (15, 0) completed
				const int s14 = c.gbase + u, s15 = c.gbase + ((j & 2) ? a15 : u);
				top_op<2>(A, C4, B, j, s15, s14, lane32, p.stream_level); // always chained with the F step below (host_tables.cc)
				__syncwarp();
			} else if (op == OP_WORD) {
				c.fmask = __ldg(&frozen[iw]);
				word32(c);
				B[(size_t)iw * 32 + lane32] = c.W;
				__syncwarp();
			} else if (op == OP_R0) {
				const int nq = 2 * hq;
				float m = c.metric;
				for (int q0 = 0; q0 < nq; q0 += 2 * kU) {
					float4 v[2 * kU];
					if (l == 16) {
#pragma unroll
						for (int u = 0; u < 2 * kU; ++u) v[u] = __ldg(&C4[q0 + u]);
					} else {
#pragma unroll
						for (int u = 0; u < 2 * kU; ++u) v[u] = P[(q0 + u) * 32 + lane32];
					}
#pragma unroll
					for (int u = 0; u < 2 * kU; ++u) m = r0_acc(m, v[u]);
				}
				c.metric = m;
				for (int w = 0; w < nq / 8; ++w) B[(size_t)(iw + w) * 32 + lane32] = 0u;
				c.ret = c.t;
				__syncwarp();
			} else { // OP_C
				const int hw = hq >> 3;
				const int src = c.gbase + c.ret;
				for (int w0 = 0; w0 < hw; w0 += 8) {
					uint32_t x[8];
#pragma unroll
					for (int k = 0; k < 8; ++k)
						if (w0 + k < hw) x[k] = B[(size_t)(iw + w0 + k) * 32 + src] ^ B[(size_t)(iw + hw + w0 + k) * 32 + lane32];
					__syncwarp();
#pragma unroll
					for (int k = 0; k < 8; ++k)
						if (w0 + k < hw) B[(size_t)(iw + w0 + k) * 32 + lane32] = x[k];
					__syncwarp();
				}
				const int lm = (int)((lmstack >> (3 * l)) & 7ull);
				c.ret = __shfl_sync(FULL, lm, src);
			}
		}

		// ---- candidate order, CRC-32 (decode.cc:532-541), payload (decode.cc:546-554) ------------------------
		int rank = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const float mj = __shfl_sync(FULL, c.metric, c.gbase + j);
			rank += (mj < c.metric) || (mj == c.metric && j < c.t);
		}
		uint32_t crc = 0;
		{
			int cnt = 0;
			for (int w = 0; w < kCodeLen / 32 && cnt < kCrcBits; ++w) {
				const uint32_t x = B[(size_t)w * 32 + lane32];
				uint32_t fr = ~__ldg(&frozen[w]);
				while (fr && cnt < kCrcBits) {
					const int b = __ffs(fr) - 1;
					fr &= fr - 1;
					const uint32_t bit = (x >> b) & 1u;
					crc = (crc >> 1) ^ (((crc ^ bit) & 1u) ? 0xD419CC15u : 0u);
					++cnt;
				}
			}
		}
		const bool ok = crc == 0u;
		int key = ok ? rank : 64;
#pragma unroll
		for (int d = 1; d < 8; d <<= 1) key = min(key, __shfl_xor_sync(FULL, key, d));
		const unsigned bal = __ballot_sync(FULL, ok && rank == key);
		const int win = __ffs((bal >> c.gbase) & 0xffu) - 1; // -1: no candidate passes the CRC
		int flips = 0;
		if (active) {
			FrameState &st = p.st[frame];
			st.metrics[rank] = c.metric;
			if (p.xbits)
				for (int w = 0; w < kCodeLen / 32; ++w)
					p.xbits[((size_t)(first + slot) * 8 + rank) * (kCodeLen / 32) + w] = B[(size_t)w * 32 + lane32];
			if (win >= 0) {
				uint32_t *out = p.payload + (size_t)frame * (kDataBytes / 4);
				for (int w = c.t; w < kCodeLen / 32; w += 8) {
					const int base = (int)__ldg(&frozen[kSclTblMsgOff + w]);
					if (base >= kDataBits) break;
					const uint32_t x = B[(size_t)w * 32 + c.gbase + win];
					uint32_t fr = ~__ldg(&frozen[w]);
					uint64_t m = 0;
					int k = 0;
					while (fr && base + k < kDataBits) {
						const int b = __ffs(fr) - 1;
						fr &= fr - 1;
						const uint32_t bit = (x >> b) & 1u;
						m |= (uint64_t)bit << k;
						flips += (int)((C[w * 32 + b] < 0.f) != (bit != 0u));
						++k;
					}
					const uint64_t sh = m << (base & 31);
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

// payload buffer <- scrambler sequence (decode.cc:613-615: out ^= xorshift); the decoder XORs the message in.
__global__ void k_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames)
{
	const int per = kDataBytes / 4;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_frames * per; i += (size_t)gridDim.x * blockDim.x)
		payload[i] = scr_words[i % per];
}

} // namespace

int scl_resident_warps(int ctas_per_sm, int n_sm) { return ctas_per_sm * n_sm * (kSclThreads / 32); }

cudaError_t launch_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames, cudaStream_t s)
{
	if (n_frames <= 0) return cudaSuccess;
	k_payload_init<<<592, 256, 0, s>>>(payload, scr_words, n_frames);
	return cudaGetLastError();
}

cudaError_t launch_polar_scl(const SclParams &p, int grid, cudaStream_t s)
{
	if (!p.n_cw_ptr && p.n_cw[0] + p.n_cw[1] <= 0) return cudaSuccess;
	k_polar_scl<<<grid, kSclThreads, kSclSmemBytes, s>>>(p);
	return cudaGetLastError();
}

int scl_occupancy_ctas_per_sm()
{
	int n = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_polar_scl, kSclThreads, kSclSmemBytes);
	return n;
}

} // namespace ofdmrx
