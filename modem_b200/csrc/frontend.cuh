// modem_b200/csrc/frontend.cuh — interface of frontend.cu / acquire.cu / demod.cu launchers.
#pragma once
#include "common.cuh"

namespace ofdmrx {

// Schmidl-Cox detections per window: the list is sized from the window length at handle creation (det_cap = one per symbol
// pitch + 8, at least 16; decode.cc:390-448 has no bound).  A window that still produces more trigger edges than fit is
// flagged (FrameState::det_overflow) instead of being silently truncated.
constexpr int kDetOverflowBit = 1 << 30; // in det_count[]: the edge list of the window overflowed

struct Detection {
	int32_t t_fall;    // stream index of the falling-edge step (decode.cc:94)
	int32_t t_max;     // stream index of the first strict maximum of the timing metric inside the segment
	float timing_max;
	int32_t index_max; // decode.cc:99-105
};

struct FrontendConsts {
	float dc_a, dc_b;  // BlockDC for 2*(1280+160) samples (decode.cc:386)
	float reco, imco[kMaxHilbertCoeffs]; // Hilbert<21>: 5 coefficients ... Hilbert<125>: 31
};

struct AcquireConsts {
	const cfx *tw1280, *tw640; // forward twiddles exp(-2 pi j k / N) for N = symbol_len and symbol_len / 2 (names: 8 kHz)
	const cfx *kern640;        // conj(FFT_half(MLS0 template)) / half (decode.cc:76-83)
	const uint8_t *mls1;       // 255 scrambler bits (decode.cc:407)
	const uint32_t *bch_rows;  // 71 x 8 words, systematic generator (decode.cc:378-384)
};

// format: OFDMRX_FMT_* (0 int16 real, 1 int16 I/Q, 2 float2 I/Q, 3 float real)
cudaError_t launch_frontend(int rate, int format, const void *samples, int64_t stride, const int32_t *n_samples, int n_default, int n_frames,
	cfx *iq, int64_t iq_stride, int iq_len, const FrontendConsts &fc, cudaStream_t s);
// masks: [n_frames][2][mask_words] trigger comparisons (v > high | v < low), one bit per stream step; the timing values are stored only
// for tiles that can lie inside a (rise, fall) segment unless write_all is set (stage taps)
cudaError_t launch_sync_metric(int rate, const cfx *iq, int64_t iq_stride, int iq_len, const int32_t *n_samples, int n_default, int n_max, int n_frames,
	float *timing, int64_t timing_stride, uint32_t *masks, int mask_words, int write_all, cudaStream_t s);
// det: [n_frames][det_cap]; edges: scratch [n_frames][2 * det_cap + 2]
cudaError_t launch_sync_detect(int rate, const float *timing, int64_t timing_stride, const uint32_t *masks, int mask_words, const int32_t *n_samples,
	int n_default, int n_frames, Detection *det, int32_t *det_count, int det_cap, int32_t *edges, cudaStream_t s);
cudaError_t launch_acquire(int rate, const cfx *iq, int64_t iq_stride, int iq_len, const Detection *det, const int32_t *det_count, int det_cap, int skip,
	int n_frames, FrameState *st, int8_t *soft_out, const AcquireConsts &ac, cudaStream_t s);
// three kernels: FFT + differential demodulation (cons_raw, phase errors yph), Theil-Sen per row (ts[row] = slope, yint,
// precision), soft demapping (llr; cons = derotated constellation, optional)
cudaError_t launch_demod(int rate, const cfx *iq, int64_t iq_stride, int iq_len, FrameState *st, int n_frames, const cfx *tw1280,
	cfx *cons_raw, float *yph, cfx *cons, float *ts, float *llr, int n_sm, cudaStream_t s, cudaEvent_t ev_fft_done = nullptr, cudaEvent_t ev_ts_done = nullptr);
cudaError_t launch_theil_sen_rows(const float *yph, int n_rows, int cols, float *ts, int n_sm, cudaStream_t s);
cudaError_t launch_compact(const FrameState *st, int n_frames, int *cw_list, int *n_cw, cudaStream_t s);

} // namespace ofdmrx
