// modem_b200/csrc/polar.cuh — interface of the list decoder kernel (polar.cu).
#pragma once
#include "common.cuh"

namespace ofdmrx {

constexpr int kSclThreads = 32;                        // one warp = 4 codewords per CTA (no CTA-level cooperation is needed)
#ifndef OFDMRX_SCL_CTAS
#define OFDMRX_SCL_CTAS 17
#endif
constexpr int kSclCtasPerSm = OFDMRX_SCL_CTAS;                      // 17 x 148 = 2516 warps: 10 000 codewords (BASELINE configs[1]) fit in one pass; needs <= 120 registers
// Per-warp scratch.  Alpha levels 6..13 back to back, level l as [codeword 0..3][slot 0..7][quad 0..2^(l-2)) float4 (a quad =
// 4 consecutive tree positions of one path; a slot = the storage of one path class, polar.cu); level 5 lives in shared
// memory, levels 14..16 are never materialised (TOP ops), levels 0..4 live in registers.
__host__ __device__ constexpr size_t scl_off4(int l) { return (size_t)32 * ((1u << (l - 2)) - 16u); } // float4 units
constexpr size_t kSclWarpQuads = scl_off4(14);               // 130 560 float4 = 2.09 MB
constexpr size_t kSclWarpFloats = kSclWarpQuads * 4;
constexpr size_t kSclWarpWords = (size_t)2048 * 32;          // beta bits, [codeword * 8 + slot][word]

// word offsets inside SclParams::tbl: frozen set | message bits before each word | CRC pieces (9 word bounds, pad, 8 x 32 matrix
// columns: the CRC register advanced over the message bits that follow piece j) | op schedule
constexpr int kSclTblMsgOff = 2048, kSclTblCrc = 4096, kSclTblOps = 4096 + 16 + 256;

struct SclParams {
	const float *llr;        // [frames][65536] channel LLRs after lengthen() (decode.cc:529)
	const int *cw_list;      // frames to decode: the header-ok frames of code table 0 (modes 6..9), then from the next
	                         // multiple of 4 on those of table 1 (modes 10..13); nullptr = identity (one table)
	int n_cw[2];             // codewords per code table, or read from n_cw_ptr (device, 2 ints) when that is set
	const int *n_cw_ptr;
	float *A;                // scratch: resident warps x kSclWarpFloats
	uint32_t *B;             // scratch: resident warps x kSclWarpWords
	const uint32_t *tbl[2];  // per code table, one array (layout above) — one base pointer keeps the kernel's registers down
	uint32_t *payload;       // [frames][1345] words pre-filled with the scrambler sequence
	FrameState *st;
	int *work;               // device counter (zeroed before the launch): next group of four codewords to hand out
	uint32_t *xbits;         // optional [n_cw][8][2048]: all candidates' codeword bits in rank order (tests)
};

int scl_resident_warps(int ctas_per_sm, int n_sm);
int scl_occupancy_ctas_per_sm();
cudaError_t launch_payload_init(uint32_t *payload, const uint32_t *scr_words, int n_frames, int *work, cudaStream_t s);
cudaError_t launch_polar_scl(const SclParams &p, int grid, cudaStream_t s);

} // namespace ofdmrx
