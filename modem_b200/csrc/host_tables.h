// modem_b200/csrc/host_tables.h — host-side constants of the mode-6 / 8 kHz receive path.
//
// Everything here is computed at handle creation from first principles (no table is copied from the
// reference): frozen set (construction of /root/reference/freezer.cc:14-32), the successive-cancellation
// node schedule derived from it, MLS sequences (decode.cc:184,187,238,407), the BCH(255,71) generator
// (decode.cc:378-384), Hilbert coefficients (decode.cc:172,193), FFT twiddles, CRC tables (decode.cc:197-198).
#pragma once
// run CALL(RATE) for the sample rate `rate` (one of the four the reference instantiates, decode.cc:590-606)
#define OFDMRX_FOR_RATE(rate, CALL) \
	switch (rate) { case 16000: CALL(16000); break; case 44100: CALL(44100); break; case 48000: CALL(48000); break; default: CALL(8000); break; }
#include <cstdint>
#include <vector>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace ofdmrx {

// ---- sample-rate geometry (decode.cc:171-173,188-189,196,590-606; SchmidlCox template arguments decode.cc:41-42) ---
// RATE in {8000, 16000, 44100, 48000}: symbol lengths 1280, 2560, 7056 (= 2^4 3^2 7^2), 7680 (= 2^9 3 5).
template <int RATE>
struct Geo {
	static constexpr int kRate = RATE;
	static constexpr int kSymLen = (1280 * RATE) / 8000, kGuardLen = kSymLen / 8, kPitch = kSymLen + kGuardLen, kHalf = kSymLen / 2;
	static constexpr int kBufferLen = 6 * kPitch;               // 8640 at 8 kHz
	static constexpr int kSearchPos = kBufferLen - 4 * kPitch;  // 2880
	static constexpr int kMatchLen = kGuardLen | 1, kMatchDel = (kMatchLen - 1) / 2; // 161, 80
	static constexpr int kFilterLen = (((21 * RATE) / 8000) & ~3) | 1; // 21, 41, 113, 125
};
constexpr int kRate = Geo<8000>::kRate;
constexpr int kSymLen = Geo<8000>::kSymLen, kGuardLen = Geo<8000>::kGuardLen, kPitch = Geo<8000>::kPitch, kHalf = Geo<8000>::kHalf;
constexpr int kBufferLen = Geo<8000>::kBufferLen, kSearchPos = Geo<8000>::kSearchPos;
constexpr int kMatchLen = Geo<8000>::kMatchLen, kMatchDel = Geo<8000>::kMatchDel;
constexpr int kFilterLen = Geo<8000>::kFilterLen;
constexpr int kMaxHilbertCoeffs = 31; // (125 - 1) / 4
constexpr int kConsCols = 432, kConsRows = 50, kModBits = 3, kConsCnt = 21600, kConsBits = 64800;
constexpr int kCodeOrder = 16, kCodeLen = 65536, kMesgBits = 43808, kDataBits = 43040, kCrcBits = 43072, kDataBytes = 5380;
constexpr int kHdrBits = 255, kHdrK = 71;

// ---- operation modes 6..13 (decode.cc:302-374): carriers per symbol, bits per carrier, transmitted code bits ----------
// rows = cons_bits / mod_bits / cols symbols follow the pilot; modes 6..9 use the frozen set of (64800, 43072), modes
// 10..13 that of (64512, 43072); payload (43040 bits) and CRC span (43072 bits) are the same for all.
struct ModeInfo { int cols, mod_bits, cons_bits, rows, table; };
__host__ __device__ constexpr ModeInfo mode_info(int mode)
{
	return mode == 6 ? ModeInfo{432, 3, 64800, 50, 0} : mode == 7 ? ModeInfo{400, 3, 64800, 54, 0}
		: mode == 8 ? ModeInfo{400, 2, 64800, 81, 0} : mode == 9 ? ModeInfo{360, 2, 64800, 90, 0}
		: mode == 10 ? ModeInfo{512, 3, 64512, 42, 1} : mode == 11 ? ModeInfo{384, 3, 64512, 56, 1}
		: mode == 12 ? ModeInfo{384, 2, 64512, 84, 1} : mode == 13 ? ModeInfo{256, 2, 64512, 126, 1} : ModeInfo{0, 0, 0, 0, 0};
}
constexpr int kMaxCols = 512, kMaxRows = 126, kMaxCons = 32400; // largest carrier count, row count, constellation count
constexpr long long kCallSignLimit = 129961739795077LL;

// ---- frozen set ---------------------------------------------------------------------------------------------
// bit i of word i/32 set => index i frozen.  make_frozen(16, 64800, 43072) == reference frozen_64800_43072.
std::vector<uint32_t> make_frozen(int order, int n_tx, int k_info);

// ---- SCL schedule -------------------------------------------------------------------------------------------
// The decoder walks the polar tree above 32-leaf "word" blocks with a precomputed op list (identical for every
// codeword, so a warp never diverges on it):
//   F(l, idx)    alpha_{l-1}[i]   = f(alpha_l[i], alpha_l[i+h])                h = 2^(l-1)
//   G(l, idx)    alpha_{l-1}[i]   = g(alpha_l[i], alpha_l[i+h], beta[idx+i])   (parent read through the left child's lane map)
//   WORD(idx)    decode the 32 leaves idx..idx+31 (frozen mask word idx/32)
//   R0(l, idx)   maximal all-frozen node: metric += sum of negative alpha_l, beta = 0
//   C(l, idx)    beta[idx..idx+h) = perm(beta[idx..idx+h)) ^ beta[idx+h..idx+2h), compose lane maps
//   TOP(13, idx) alpha_13 of the level-13 node at idx recomputed straight from the channel LLRs and the partial sums of its
//                left-hand relatives (levels 16..14 are then never materialised: their alphas are functions of the
//                lane-shared channel values and a few beta bits)
//   R1(l, idx)   rate-1 attempt on a maximal all-free node (6 <= l <= kSclR1MaxLevel), followed by one extra word = the pc to continue at
//                when it succeeds: if the lanes are in metric order and every lane's  metric + min|alpha_l|  exceeds the
//                largest metric, no fork inside the node can change the list, so beta = sign bits of alpha_l, metrics and
//                lanes stay as they are (exact: the smallest |leaf LLR| of a sign-following path IS min|alpha_l|, every
//                other leaf's is a rounded sum >= it).  Otherwise the ops of the sub-tree follow as usual; the left spine
//                of a rate-1 node carries no further attempts (same metrics, same minimum).
enum SclOp : uint32_t { OP_F = 0, OP_G = 1, OP_WORD = 2, OP_R0 = 3, OP_C = 4, OP_TOP = 5, OP_R1 = 6, OP_END = 7 };
constexpr int kSclTopLevel = 13;
constexpr int kSclR1MaxLevel = 12;
constexpr int kSclMaxFuse = 2; // longest op chain the kernel instantiates: the op itself + one F step
// F and G carry a fusion depth d = 1..3 in bits 30..31 (stored as d-1): the op also performs the d-1 F steps that always
// follow it on the way down the left spine (F(l-1), F(l-2)), so those levels are produced in registers and written once.
static inline uint32_t scl_pack(uint32_t op, uint32_t level, uint32_t index, uint32_t depth = 1) { return op | (level << 3) | ((index / 32) << 8) | ((depth - 1) << 30); }
static inline uint32_t scl_op(uint32_t w) { return w & 7; }
static inline uint32_t scl_level(uint32_t w) { return (w >> 3) & 31; }
static inline uint32_t scl_index(uint32_t w) { return ((w >> 8) & 0x3fffffu) * 32; }
static inline uint32_t scl_depth(uint32_t w) { return (w >> 30) + 1; }
std::vector<uint32_t> make_scl_schedule(const std::vector<uint32_t> &frozen, int order, int max_depth = kSclMaxFuse, bool top_ops = true, bool r1_ops = true);

// ---- misc sequences / codes -----------------------------------------------------------------------------------
std::vector<uint8_t> mls_bits(int poly, int n);              // first n outputs of the Galois LFSR (reg = 1)
std::vector<uint32_t> bch_generator_rows();                  // 71 rows x 8 words, bit j of row i = G[i][j], systematic [I|P]
std::vector<float> hilbert_coeffs(int taps, float *reco);    // imag-branch coefficients for odd offsets 1,3,..
std::vector<float> twiddles(int n, int sign);                // n complex values exp(sign*2*pi*j*k/n) as (re,im) pairs
std::vector<float> mls0_kernel(int half = kHalf);             // conj(FFT_half(template))/half, `half` complex values (decode.cc:76-83,236-244)
void crc32_table(uint32_t poly, uint32_t *lut256);
// CRC-32 (0xD419CC15, reflected, init 0) of the first crc_bits message bits in 8 pieces of about equal length: out[0..8] = word
// bounds of the pieces (message bits = the non-frozen positions), out[16 + 32 j + c] = register bit c advanced over the message
// bits that follow piece j.  The list decoder XORs the advanced piece registers (decode.cc:534-537 computes the same CRC bit by bit).
std::vector<uint32_t> crc32_pieces(const std::vector<uint32_t> &frozen, int crc_bits);
uint16_t crc16_u64(uint64_t v);                              // CRC-16 0xA8F4 over the 8 LE bytes (decode.cc:428-429)
void base37_decode(char *str, long long val, int len);       // decode.cc:155-159

} // namespace ofdmrx
