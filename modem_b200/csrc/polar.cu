// modem_b200/csrc/polar.cu — batched CRC-aided polar successive-cancellation list decoding, N = 65536, L = 8.
//
// Replaces CODE::PolarListDecoder<SIMD<float,8>,16> + systematic() + the CRC-32 candidate scan of the reference
// receiver (/root/reference/decode.cc:201,530-555).  Design (B200-first, not a port of the SIMD recursion):
//   * one list lane per THREAD, one codeword per 8 threads, four codewords per warp.  Every codeword walks the
//     same precomputed op schedule (host_tables.cc: the frozen set is fixed), so a warp never diverges and the
//     only cross-thread traffic is 8-wide shuffles (lane permutation after a fork, fork ranking).
//   * alpha (LLR) buffers of tree levels 5..15 live in a per-warp HBM/L2 scratch laid out [element][warp lane]
//     so that every warp access is one coalesced 128-byte line, also when a thread reads through the lane map;
//     levels 0..4 (a 32-leaf "word") are fully unrolled and live in registers.
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
	const float *A5;  // level-5 alpha buffer of this warp ([32][32])
};

// Free leaf: 2L forks, keep the L smallest by (metric, fork index); survivors land in rank order.
__device__ __forceinline__ void leaf_fork(SclCtx &c, float a, int pos)
{
	const float pen = fabsf(a);
	const float m0 = a < 0.f ? __fadd_rn(c.metric, pen) : c.metric; // decide 0
	const float m1 = a < 0.f ? c.metric : __fadd_rn(c.metric, pen); // decide 1
	float o0[8], o1[8];
	int r0 = 0, r1 = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		o0[j] = __shfl_sync(FULL, m0, c.gbase + j);
		o1[j] = __shfl_sync(FULL, m1, c.gbase + j);
		const bool lt = j < c.t, le = j <= c.t;
		r0 += (o0[j] < m0) || (o0[j] == m0 && lt);
		r0 += (o1[j] < m0) || (o1[j] == m0 && lt);
		r1 += (o0[j] < m1) || (o0[j] == m1 && le);
		r1 += (o1[j] < m1) || (o1[j] == m1 && lt);
	}
	const int packed = r0 | (r1 << 8);
	int src = 0, bit = 0;
	float nm = 0.f;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const int pr = __shfl_sync(FULL, packed, c.gbase + j);
		if ((pr & 255) == c.t) { src = j; bit = 0; nm = o0[j]; }
		if ((pr >> 8) == c.t) { src = j; bit = 1; nm = o1[j]; }
	}
	c.metric = nm;
	c.ret = src;
	c.W |= (uint32_t)bit << pos;
}

// One node of the 32-leaf word, LVL = log2(size), BASE = first leaf.  `a` = this node's alpha values in
// registers (unused for LVL 5, whose alphas are in the level-5 scratch buffer).
template <int LVL, int BASE>
__device__ __forceinline__ void blk_node(SclCtx &c, const float *a)
{
	constexpr int N = 1 << LVL;
	if constexpr (LVL < 5) {
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
	}
	if constexpr (LVL == 0) {
		leaf_fork(c, a[0], BASE);
	} else {
		constexpr int H = N / 2;
		float ch[H];
		const int own = c.gbase + c.t;
#pragma unroll
		for (int k = 0; k < H; ++k) {
			float pa, pb;
			if constexpr (LVL == 5) { pa = c.A5[k * 32 + own]; pb = c.A5[(k + H) * 32 + own]; }
			else { pa = a[k]; pb = a[k + H]; }
			ch[k] = f_op(pa, pb);
		}
		blk_node<LVL - 1, BASE>(c, ch);
		const int lmap = c.ret;
		const int srcl = c.gbase + lmap;
#pragma unroll
		for (int k = 0; k < H; ++k) {
			float pa, pb;
			if constexpr (LVL == 5) { pa = c.A5[k * 32 + srcl]; pb = c.A5[(k + H) * 32 + srcl]; }
			else { pa = __shfl_sync(FULL, a[k], srcl); pb = __shfl_sync(FULL, a[k + H], srcl); }
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

__global__ void __launch_bounds__(kSclThreads) k_polar_scl(SclParams p)
{
	const int lane32 = threadIdx.x & 31;
	const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int n_warps = (gridDim.x * blockDim.x) >> 5;
	float *A = p.A + (size_t)warp_global * kSclWarpFloats;
	uint32_t *B = p.B + (size_t)warp_global * kSclWarpWords;
	SclCtx c;
	c.t = lane32 & 7;
	c.gbase = lane32 & ~7;
	c.A5 = A + scl_off(5);

	const int n_cw = p.n_cw_ptr ? *p.n_cw_ptr : p.n_cw;
	for (int g4 = warp_global; g4 * 4 < n_cw; g4 += n_warps) {
		const int slot = g4 * 4 + (lane32 >> 3);
		const bool active = slot < n_cw;
		const int frame = p.cw_list ? p.cw_list[active ? slot : n_cw - 1] : (active ? slot : n_cw - 1);
		const float *C = p.llr + (size_t)frame * kCodeLen;
		c.metric = c.t == 0 ? 0.f : 1000.f;
		c.ret = c.t;
		uint64_t lmstack = 0;

		for (int pc = 0;; ++pc) {
			const uint32_t opw = __ldg(&p.ops[pc]);
			const uint32_t op = opw & 7u, l = (opw >> 3) & 31u, iw = opw >> 8; // iw = first word of the node
			if (op == OP_END) break;
			const int h = 1 << (l - 1);
			const float *P = A + scl_off(l);   // parent level (valid for l <= 15)
			float *D = A + scl_off(l - 1);
			if (op == OP_F) {
				if (l == 16) {
#pragma unroll 4
					for (int i = 0; i < h; ++i) D[i * 32 + lane32] = f_op(C[i], C[i + h]);
				} else {
#pragma unroll 4
					for (int i = 0; i < h; ++i) D[i * 32 + lane32] = f_op(P[i * 32 + lane32], P[(i + h) * 32 + lane32]);
				}
				__syncwarp();
			} else if (op == OP_G) {
				lmstack = (lmstack & ~(7ull << (3 * l))) | ((uint64_t)c.ret << (3 * l));
				const int src = c.gbase + c.ret;
				const uint32_t *Bw = B + (size_t)iw * 32 + lane32;
				for (int i0 = 0; i0 < h; i0 += 32) {
					const uint32_t bw = Bw[(i0 >> 5) * 32];
					if (l == 16) {
#pragma unroll 8
						for (int k = 0; k < 32; ++k) {
							const int i = i0 + k;
							D[i * 32 + lane32] = g_op(C[i], C[i + h], (bw >> k) & 1u);
						}
					} else {
#pragma unroll 8
						for (int k = 0; k < 32; ++k) {
							const int i = i0 + k;
							D[i * 32 + lane32] = g_op(P[i * 32 + src], P[(i + h) * 32 + src], (bw >> k) & 1u);
						}
					}
				}
				__syncwarp();
			} else if (op == OP_WORD) {
				c.fmask = __ldg(&p.frozen[iw]);
				c.W = 0;
				blk_node<5, 0>(c, nullptr);
				B[(size_t)iw * 32 + lane32] = c.W;
				__syncwarp();
			} else if (op == OP_R0) {
				const int n = 2 * h;
				float m = c.metric;
				if (l == 16) {
					for (int i = 0; i < n; ++i) { const float v = C[i]; if (v < 0.f) m = __fsub_rn(m, v); }
				} else {
#pragma unroll 4
					for (int i = 0; i < n; ++i) { const float v = P[i * 32 + lane32]; if (v < 0.f) m = __fsub_rn(m, v); }
				}
				c.metric = m;
				for (int w = 0; w < n / 32; ++w) B[(size_t)(iw + w) * 32 + lane32] = 0u;
				c.ret = c.t;
				__syncwarp();
			} else { // OP_C
				const int hw = h >> 5;
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
				uint32_t fr = ~__ldg(&p.frozen[w]);
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
					p.xbits[((size_t)slot * 8 + rank) * (kCodeLen / 32) + w] = B[(size_t)w * 32 + lane32];
			if (win >= 0) {
				uint32_t *out = p.payload + (size_t)frame * (kDataBytes / 4);
				for (int w = c.t; w < kCodeLen / 32; w += 8) {
					const int base = (int)__ldg(&p.msg_off[w]);
					if (base >= kDataBits) break;
					const uint32_t x = B[(size_t)w * 32 + c.gbase + win];
					uint32_t fr = ~__ldg(&p.frozen[w]);
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
	if (!p.n_cw_ptr && p.n_cw <= 0) return cudaSuccess;
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
