/* include/ofdmtx.h — C-ABI of the device-side stimulus generator in libofdmrx.so.
 *
 * Stands in for the reference's transmitter command line
 *     encode OUTPUT RATE BITS CHANNELS OFFSET MODE CALLSIGN INPUT..        (/root/reference/encode.cc:337-446)
 * piped through the `disorders` tools of README.md:49, for n independent windows at once: it exists so that the large
 * gates of the receive path (10^4 … 10^6 windows) do not wait for a CPU encoder.  Same conventions as ofdmrx.h: plain
 * pointers and sizes, 0 / negative return codes, caller-owned I/O buffers, one handle per (device, host thread), no CPU
 * fallback.
 */
#ifndef OFDMTX_H
#define OFDMTX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ofdmtx_handle ofdmtx_t;

/* The impairment chain in README.md:49 order.  aicodix/disorders is not part of the reference repository; these are the
 * definitions the CPU oracle uses (oracle/ref_modem.hh: apply_impairments): a fixed 4-tap sparse complex FIR, a complex
 * mixer, 33-tap Kaiser-windowed-sinc resampling by (1 + ppm 1e-6), complex Gaussian noise of total variance
 * 10^(awgn_db / 10).  Window i of a call draws its noise from the Philox-4x32-10 stream with key `seed` and counter (sample index, i). */
typedef struct ofdmtx_impairments {
	int32_t multipath; /* 0 / 1 */
	float cfo_hz;
	float sfo_ppm;
	int32_t awgn;      /* 0 / 1 */
	float awgn_db;
	uint64_t seed;
} ofdmtx_impairments;

/* Replaces: `new Encoder<float, Complex<float>, RATE>` set-up (encode.cc:272-287,424-440): twiddles, code tables, CRC
 * table — plus device scratch for max_windows windows of frames_per_window frames each, processed at a time (larger
 * batches are chunked).  rate_hz: 8000, 16000, 44100 or 48000 (encode.cc:424-440), else -22. */
int ofdmtx_create(ofdmtx_t **h, int device, int rate_hz, int max_windows, int frames_per_window);
void ofdmtx_destroy(ofdmtx_t *h);

/* encode.cc:319-335: base-37 value of a call sign, -1 if it holds a character outside " 0-9A-Za-z" */
int64_t ofdmtx_call_sign(const char *str);
/* sample frames of one window: 1 s of silence, leading pilot, frames_per_window x (Schmidl-Cox, metadata, pilot, data
 * rows), one zero symbol, 1 s of silence (encode.cc:288-313,423,441); mode 6 at 8 kHz, one frame: 95200 */
int64_t ofdmtx_window_samples(int rate_hz, int mode, int frames_per_window);

/* Replaces: one `encode - RATE 16 CHANNELS OFFSET MODE CALLSIGN INPUT.. | multipath | cfo | sfo | awgn` pipeline per
 * window (encode.cc:271-317,403-441; README.md:49).  payloads: n_windows x frames_per_window x 5380 plain bytes (host or
 * device, payload_mem = OFDMRX_MEM_*); samples_out: n_windows windows at a pitch of frame_stride_samples sample frames
 * in sample_format (OFDMRX_FMT_S16_MONO: real part only, as CHANNELS = 1 writes it; OFDMRX_FMT_S16_IQ;
 * OFDMRX_FMT_F32_IQ: the un-quantised analytic stream), zero-filled behind the window; imp may be NULL (clean channel);
 * n_samples_out (host, may be NULL) receives the sample frames each window holds (ofdmtx_window_samples divided by
 * 1 + sfo_ppm 1e-6).  Argument rules as the reference command line enforces them (mode 6..13, call sign in (0, 37^9),
 * offset a multiple of 50 Hz inside the band limits of the mode, encode.cc:345-397): -22 otherwise, and -22 if
 * frame_stride_samples is shorter than the window. */
int ofdmtx_encode_batch(ofdmtx_t *h, const uint8_t *payloads, int payload_mem, int n_windows, int mode, int64_t call_sign,
	int freq_off_hz, const ofdmtx_impairments *imp, void *samples_out, int mem_kind, int sample_format,
	int64_t frame_stride_samples, int32_t *n_samples_out, void *stream);

/* Test tap: the transmitted code bits (2048 words per frame, bit i of word i/32 = code bit i; the first cons_bits are on
 * the air) of the LAST chunk encoded. */
int ofdmtx_get_code(ofdmtx_t *h, int frame_first, int frame_count, uint32_t *dst);
/* kernels launched by the last ofdmtx_encode_batch call */
int ofdmtx_last_launches(ofdmtx_t *h);

#ifdef __cplusplus
}
#endif
#endif /* OFDMTX_H */
