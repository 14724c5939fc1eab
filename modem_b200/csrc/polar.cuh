// modem_b200/csrc/polar.cuh — interface of the list decoder kernel (polar.cu).
#pragma once
#include "common.cuh"

namespace ofdmrx {

constexpr int kSclThreads = 32;                        // one warp = 4 codewords per CTA (no CTA-level cooperation is needed)
constexpr int kSclCtasPerSm = 17;                      // 17 x 148 = 2516 warps: 10 000 codewords (BASELINE configs[1]) fit in one pass; needs <= 120 registers
constexpr size_t kSclWarpFloats = (size_t)(65536 - 32) * 32; // alpha levels 5..15, [element][warp lane]
constexpr size_t kSclWarpWords = (size_t)2048 * 32;          // beta bits, [word][warp lane]
__host__ __device__ constexpr size_t scl_off(int l) { return (size_t)((1 << l) - 32) * 32; }
// same offset in float4 units: level l holds 2^l/4 quads per lane, laid out [quad][warp lane]
__host__ __device__ constexpr size_t scl_off4(int l) { return (size_t)((1 << l) - 32) * 8; }

constexpr int kSclTblMsgOff = 2048, kSclTblOps = 4096; // word offsets inside SclParams::tbl

struct SclParams {
	const float *llr;        // [frames][65536] channel LLRs after lengthen() (decode.cc:529)
	const int *cw_list;      // frames to decode: the header-ok frames of code table 0 (modes 6..9), then from the next
	                         // multiple of 4 on those of table 1 (modes 10..13); nullptr = identity (one table)
	int n_cw[2];             // codewords per code table, or read from n_cw_ptr (device, 2 ints) when that is set
	const int *n_cw_ptr;
	float *A;                // scratch: resident warps x a_stride floats
	size_t a_stride;         // kSclWarpFloats, or scl_off(14) when the schedules use TOP ops (levels 14, 15 never stored)
	uint32_t *B;             // scratch: resident warps x kSclWarpWords
	const uint32_t *tbl[2];  // per code table, one array: frozen set (2048 words), number of non-frozen indices before each
	                         // word (2048), op schedule (host_tables.cc) — one base pointer keeps the kernel's registers down
	uint32_t *payload;       // [frames][1345] words pre-filled with the scrambler sequence
	FrameState *st;
	int stream_level;        // alpha levels >= this use the L2 evict-first policy (17 = none)
	uint32_t *xbits;         // optional [n_cw][8][2048]: all candidates' codeword bits in rank order (tests)
};

int scl_resident_warps(int ctas_per_sm, int n_sm);
int scl_occupancy_ctas_per_sm();
cudaError_t launch_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames, cudaStream_t s);
cudaError_t launch_polar_scl(const SclParams &p, int grid, cudaStream_t s);

} // namespace ofdmrx
