/* include/ofdmrx.h — C-ABI of libofdmrx.so: batched B200 (sm_100a) receive path for aicodix/modem frames (modes 6..13 at 8000, 16000, 44100 and 48000 Hz).
 *
 * The reference has NO plugin / FFI interface: its only stable contract is the command line
 * `decode OUTPUT INPUT [SKIP]` with WAV in / 5380 payload bytes out (/root/reference/decode.cc:559-620).  This
 * header is the boundary underneath that contract: the `decode` host driver (modem_b200/csrc/host/decode_main.cc),
 * the Python mirror (modem_b200/__init__.py) and any foreign binding call exactly these symbols.
 * Each entry point names the reference code it stands in for.
 *
 * Conventions: plain pointers and sizes only; return 0 on success, negative on error (never throws); the caller owns
 * every I/O buffer, the handle owns all device scratch; one handle per (device, host thread) — calls on one handle
 * are not re-entrant; there is NO CPU fallback: creation fails if no sm_100-class GPU is present.
 */
#ifndef OFDMRX_H
#define OFDMRX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OFDMRX_PAYLOAD_BYTES 5380 /* 43040 bits, decode.cc:174,587 */
#define OFDMRX_CODE_LEN 65536     /* polar code length, decode.cc:308 */

/* sample formats: what DSP::ReadWAV<float> would deliver (decode.cc:294-301,576-581) */
#define OFDMRX_FMT_S16_MONO 0 /* 1 channel, real: DC blocker + Hilbert are applied (decode.cc:298-299) */
#define OFDMRX_FMT_S16_IQ 1   /* 2 channels, analytic I/Q, used as is */
#define OFDMRX_FMT_F32_IQ 2   /* float2 I/Q already scaled to [-1,1]: what ReadWAV delivers for 2 channels at any bit depth */
#define OFDMRX_FMT_F32_MONO 3 /* float real samples in [-1,1] (8 / 24 / 32-bit WAVs at the reference's own precision); DC blocker + Hilbert applied */

#define OFDMRX_MEM_HOST 0   /* samples/payload/status pointers are host memory (pinned memory makes the copies async) */
#define OFDMRX_MEM_DEVICE 1 /* pointers are device memory on the handle's GPU */

/* per-window outcome — what decode.cc only prints on stderr (decode.cc:400-401,419,430,435,438,440,446,543,555) */
#define OFDMRX_ST_OK 0               /* payload CRC-32 matched */
#define OFDMRX_ST_NO_SYNC 1          /* no accepted Schmidl-Cox detection (decode.cc:392-396) */
#define OFDMRX_ST_OSD_FAIL 2         /* "OSD error." (decode.cc:417-421) */
#define OFDMRX_ST_HDR_CRC 3          /* "header CRC error." (decode.cc:428-432) */
#define OFDMRX_ST_BAD_MODE 4         /* "operation mode N unsupported." (decode.cc:434-437) */
#define OFDMRX_ST_BAD_CALL 5         /* "call sign unsupported." (decode.cc:439-442) */
#define OFDMRX_ST_PAYLOAD_CRC 6      /* "payload decoding error." (decode.cc:542-545) */
#define OFDMRX_ST_UNSUPPORTED_MODE 7 /* (not produced any more: modes 6..13 are all decoded on the GPU) */

typedef struct ofdmrx_handle ofdmrx_t;

typedef struct ofdmrx_frame_status {
	int32_t status;      /* OFDMRX_ST_* */
	int32_t detections;  /* accepted correlator detections consumed (SKIP semantics, decode.cc:390-448) */
	int32_t t_fire;      /* stream index of the sample whose arrival fired the detection */
	int32_t symbol_pos;  /* "symbol pos" (decode.cc:398,400) */
	int32_t sc_pos;      /* absolute stream index of the Schmidl-Cox symbol body */
	int32_t index_max, shift, pos_err; /* decode.cc:103-105,127-146 */
	float timing_max, frac_cfo;
	float cfo_rad;       /* "coarse cfo" = cfo_rad * rate / 2pi (decode.cc:148-150,401) */
	int32_t osd_unique;
	int32_t mode;        /* "oper mode" (decode.cc:433,438) */
	uint32_t md_lo, md_hi; /* 55-bit metadata word: (call_sign << 8) | mode (decode.cc:422-424) */
	int32_t best_lane;   /* list candidate (ascending metric) whose CRC matched, -1 if none (decode.cc:532-541) */
	int32_t flips;       /* "bit flips" (decode.cc:546-555) */
	float metrics[8];    /* final path metrics, ascending */
	int32_t osd_visited;
	int32_t ts_sweeps;   /* pair sweeps the Theil-Sen search took, summed over the rows (rows = the minimum) */
	int32_t det_overflow; /* 1: the window produced more correlator trigger edges than its detection list holds (sized at
	                       * creation: one per symbol pitch of max_samples_per_frame + 8); detections beyond it were not examined */
} ofdmrx_frame_status;

/* stages whose outputs can be read back for parity tests (ofdmrx_get_taps); layouts are per window */
#define OFDMRX_TAP_IQ 0       /* float2[iq_len]           analytic stream after next_sample() (decode.cc:294-301) */
#define OFDMRX_TAP_TIMING 1   /* float[iq_len]            box-161 timing metric per stream step (decode.cc:90) */
#define OFDMRX_TAP_SOFT 2     /* int8[256]                header soft bits (decode.cc:410-416) */
/* constellation taps hold rows x cols values of the window's mode (mode 6: 50 x 432) at the front of 32400 slots */
#define OFDMRX_TAP_CONS_RAW 3 /* float2[32400]            cons after demod_or_erase (decode.cc:475) */
#define OFDMRX_TAP_CONS 4     /* float2[32400]            cons after Theil-Sen derotation (decode.cc:494) */
#define OFDMRX_TAP_TS 5       /* float[126*3]             slope, yint, precision per row (decode.cc:488-492,517) */
#define OFDMRX_TAP_LLR 6      /* float[65536]             code[] after lengthen() (decode.cc:529) */
#define OFDMRX_TAP_PHASE 7    /* float[32400]             decision-directed phase errors fed to Theil-Sen (decode.cc:483-486) */

/* Replaces: `new Decoder<float, Complex<float>, 8000>` set-up work (decode.cc:375-387,590-606): constant tables,
 * BCH generator, correlator kernel — plus device scratch for up to max_frames windows of max_samples sample frames
 * processed at a time (larger batches are chunked internally).  rate_hz: 8000, 16000, 44100 or 48000 — the four rates the
 * reference instantiates (decode.cc:590-606); anything else returns -22 ("Unsupported sample rate."). */
int ofdmrx_create(ofdmrx_t **h, int device, int rate_hz, int max_frames, int max_samples_per_frame);
void ofdmrx_destroy(ofdmrx_t *h);

/* Tunables / test switches: "keep_taps" (0/1, default 0), "scl_ctas_per_sm" (resident list-decoder warps per SM, 1..occupancy,
 * before the first decode; default: the occupancy limit, 21), "polar_table" (0: modes 6..9, 1: modes 10..13 — code table used by ofdmrx_polar_decode). */
int ofdmrx_set_option(ofdmrx_t *h, const char *key, int value);

/* Replaces: one `decode OUTPUT INPUT [SKIP]` invocation per window (decode.cc:375-556 + the de-scrambling of
 * decode.cc:613-615), for n_frames independent windows.  Window i starts at samples + i*frame_stride_samples sample
 * frames and holds n_samples[i] (host array, or NULL = frame_stride_samples) frames.  payload_out: n_frames x 5380
 * bytes, de-scrambled; a failed window yields the de-scrambled all-zero buffer (the reference writes an
 * uninitialised one, decode.cc:588).  status_out may be NULL.  `stream` is a cudaStream_t (NULL = default stream);
 * with OFDMRX_MEM_DEVICE the call only enqueues work, with OFDMRX_MEM_HOST it returns after the results are in
 * host memory. */
int ofdmrx_decode_batch(ofdmrx_t *h, const void *samples, int mem_kind, int sample_format, int n_frames,
	int64_t frame_stride_samples, const int32_t *n_samples, int skip, uint8_t *payload_out,
	ofdmrx_frame_status *status_out, void *stream);

/* Replaces: polardec() + systematic() + CRC-32 scan + bit output (decode.cc:530-555) for n host-resident LLR vectors
 * of 65536 floats.  xbits (optional): n*8*2048 words, every candidate's re-encoded codeword in ascending-metric order. */
int ofdmrx_polar_decode(ofdmrx_t *h, const float *llr, int n, uint8_t *payload_out, ofdmrx_frame_status *status_out,
	uint32_t *xbits);

/* Replaces: DSP::TheilSenEstimator<float,512>::compute over x = i - cols/2 (decode.cc:452,484,488) for n_rows
 * host-resident rows of `cols` (<= 512) phase values; out3 receives (slope, yint, pair sweeps the search took) per row. */
int ofdmrx_theil_sen(ofdmrx_t *h, const float *y, int n_rows, int cols, float *out3);

/* Copies stage outputs of the LAST decode_batch chunk (needs option keep_taps=1 for CONS and TIMING) to host memory. */
int ofdmrx_get_taps(ofdmrx_t *h, int stage, int frame_first, int frame_count, void *dst, size_t bytes);
/* elements per window of a tap (in units of the tap's element type) */
int64_t ofdmrx_tap_elems(ofdmrx_t *h, int stage);

/* Measured FP32 FMA throughput of the device (TFLOP/s, a dependent-chain-free FMA kernel over all SMs): the roofline
 * denominator SURVEY.md 8(d) names for the list decoder. */
int ofdmrx_measure_fp32(ofdmrx_t *h, float *tflops);
/* kernels launched by the last ofdmrx_decode_batch / ofdmrx_polar_decode call (bench.py's gpu_launches) */
int ofdmrx_last_launches(ofdmrx_t *h);
/* CUDA-event durations (ms, on the launching stream) of the stages of the LAST chunk processed by decode_batch:
 * ms[0] frontend, [1] timing metric, [2] detection, [3] acquire (fine sync + header), [4] demod (FFT/Theil-Sen/LLR),
 * [5] compaction + payload init, [6] polar list decoder; with n >= 10 also the three kernels of [4]: [7] FFT + differential
 * demodulation, [8] Theil-Sen, [9] soft demapping.  Returns the number of windows in that chunk (<0 on error). */
int ofdmrx_stage_times(ofdmrx_t *h, float *ms, int n);
/* device-side copies of the constant tables, for tests: which = 0 frozen set (2048 u32) / 1 SCL schedule of modes 6..9,
 * 2 / 3 the same for modes 10..13 */
int ofdmrx_get_table(ofdmrx_t *h, int which, void *dst, size_t bytes);
const char *ofdmrx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OFDMRX_H */
