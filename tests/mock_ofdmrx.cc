// tests/mock_ofdmrx.cc — TEST HELPER.  A CPU stand-in for libofdmrx.so's C-ABI (include/ofdmrx.h) built on the oracle, so that
// the product's host drivers (modem_b200/csrc/host/decode_main.cc) can be exercised without a GPU: argv handling, WAV parsing,
// the SKIP walk and every stderr line are then diffed against the reference's own main() (oracle/_ref/decode) by
// tests/test_cli_host.py.  Never shipped, never linked into the product: the real library has no CPU path.
#include "../include/ofdmrx.h"
#include "../include/ofdmtx.h"
#include "../oracle/ref_modem.hh"
#include <cstring>
#include <vector>

struct ofdmtx_handle { int rate, fpw; };

struct ofdmrx_handle {
	int rate, max_frames;
	std::vector<std::vector<float>> ts; // per window of the last call: slope, yint, precision per row
};

extern "C" {

const char *ofdmrx_version(void) { return "mock over the CPU oracle (tests only)"; }
int ofdmrx_create(ofdmrx_t **h, int, int rate_hz, int max_frames, int)
{
	if (rate_hz != 8000 && rate_hz != 16000 && rate_hz != 44100 && rate_hz != 48000) return -22;
	*h = new ofdmrx_handle{rate_hz, max_frames, {}};
	return 0;
}
void ofdmrx_destroy(ofdmrx_t *h) { delete h; }
int ofdmrx_decode_batch(ofdmrx_t *h, const void *samples, int mem_kind, int format, int n_frames, int64_t stride, const int32_t *n_samples,
	int skip, uint8_t *payload_out, ofdmrx_frame_status *st, void *)
{
	if (mem_kind != OFDMRX_MEM_HOST || format < 0 || format > OFDMRX_FMT_F32_MONO) return -22;
	const int ch = (format == OFDMRX_FMT_S16_MONO || format == OFDMRX_FMT_F32_MONO) ? 1 : 2;
	const bool is_float = format == OFDMRX_FMT_F32_IQ || format == OFDMRX_FMT_F32_MONO;
	h->ts.assign(n_frames, std::vector<float>(126 * 3, 0.f));
	for (int i = 0; i < n_frames; ++i) {
		const int16_t *pcm = (const int16_t *)samples + (size_t)i * stride * ch;
		const float *flt = (const float *)samples + (size_t)i * stride * ch;
		const int64_t n = n_samples ? n_samples[i] : stride;
		std::vector<float> f((size_t)n * ch);
		for (size_t k = 0; k < f.size(); ++k) f[k] = is_float ? flt[k] : float(pcm[k]) / 32767.f;
		ref::Receiver rx(h->rate);
		uint8_t buf[ref::kDataBytes];
		std::memset(buf, 0, sizeof(buf));
		int status = rx.run(buf, f.data(), (size_t)n, ch, skip);
		ref::descramble(buf);
		std::memcpy(payload_out + (size_t)i * ref::kDataBytes, buf, ref::kDataBytes);
		const ref::Taps &t = rx.taps;
		if (st) {
			ofdmrx_frame_status &s = st[i];
			std::memset(&s, 0, sizeof(s));
			s.status = status; s.detections = t.detections; s.t_fire = t.t_fire; s.symbol_pos = t.symbol_pos; s.sc_pos = t.sc_pos;
			s.index_max = t.index_max; s.shift = t.shift; s.pos_err = t.pos_err; s.timing_max = t.timing_max; s.frac_cfo = t.frac_cfo;
			s.cfo_rad = t.cfo_rad; s.osd_unique = t.osd_unique; s.mode = t.mode; s.md_lo = (uint32_t)t.md; s.md_hi = (uint32_t)(t.md >> 32);
			s.best_lane = t.best_lane; s.flips = t.flips;
		}
		for (size_t j = 0; j < t.slope.size() && j < 126; ++j) {
			h->ts[i][3 * j] = t.slope[j]; h->ts[i][3 * j + 1] = t.yint[j]; h->ts[i][3 * j + 2] = t.precision[j];
		}
	}
	return 0;
}
int ofdmrx_measure_fp32(ofdmrx_t *, float *) { return -38; }
int64_t ofdmrx_tap_elems(ofdmrx_t *, int stage) { return stage == OFDMRX_TAP_TS ? 126 * 3 : -22; }
int ofdmrx_get_taps(ofdmrx_t *h, int stage, int first, int count, void *dst, size_t bytes)
{
	if (stage != OFDMRX_TAP_TS || first < 0 || first + count > (int)h->ts.size() || bytes < (size_t)count * 126 * 3 * 4) return -22;
	for (int i = 0; i < count; ++i) std::memcpy((float *)dst + (size_t)i * 126 * 3, h->ts[first + i].data(), 126 * 3 * 4);
	return 0;
}

// entry points the Python mirror binds at load time but the host-driver tests never call
int ofdmrx_set_option(ofdmrx_t *, const char *, int) { return 0; }
int ofdmrx_polar_decode(ofdmrx_t *, const float *, int, uint8_t *, ofdmrx_frame_status *, uint32_t *) { return -38; }
int ofdmrx_theil_sen(ofdmrx_t *, const float *, int, int, float *) { return -38; }
int ofdmrx_last_launches(ofdmrx_t *) { return 0; }
int ofdmrx_stage_times(ofdmrx_t *, float *, int) { return -38; }
int ofdmrx_get_table(ofdmrx_t *, int, void *, size_t) { return -38; }

// ---- include/ofdmtx.h over the oracle's transmitter: clean streams only, float I/Q or 16-bit output to host memory
int ofdmtx_create(ofdmtx_t **h, int, int rate_hz, int, int frames_per_window)
{
	if (rate_hz != 8000 && rate_hz != 16000 && rate_hz != 44100 && rate_hz != 48000) return -22;
	*h = new ofdmtx_handle{rate_hz, frames_per_window};
	return 0;
}
void ofdmtx_destroy(ofdmtx_t *h) { delete h; }
int64_t ofdmtx_call_sign(const char *str) { return ref::base37_encode(str); }
int64_t ofdmtx_window_samples(int rate_hz, int mode, int frames_per_window)
{
	ref::ModeParams mp{};
	if (!ref::mode_params(mode, mp)) return -22;
	const int pitch = 1440 * rate_hz / 8000;
	return 2LL * rate_hz + (2LL + (long long)frames_per_window * (3 + mp.cons_rows())) * pitch;
}
int ofdmtx_encode_batch(ofdmtx_t *h, const uint8_t *payloads, int, int n_windows, int mode, int64_t call_sign, int freq_off_hz,
	const ofdmtx_impairments *imp, void *samples_out, int mem_kind, int format, int64_t stride, int32_t *n_samples_out, void *)
{
	if (mem_kind != OFDMRX_MEM_HOST) return -22;
	if (!ref::Transmitter::check_args(h->rate, format == OFDMRX_FMT_S16_MONO ? 1 : 2, freq_off_hz, mode, call_sign)) return -22;
	for (int i = 0; i < n_windows; ++i) {
		ref::Transmitter tx(h->rate);
		std::vector<ref::cf> s;
		if (!tx.encode(s, payloads + (size_t)i * h->fpw * ref::kDataBytes, h->fpw, freq_off_hz, call_sign, mode)) return -22;
		if (imp) { // the oracle's chain (its own mt19937 noise: the mock is about plumbing, not about the Philox stream)
			ref::Impair im;
			im.multipath = imp->multipath != 0; im.cfo_hz = imp->cfo_hz; im.sfo_ppm = imp->sfo_ppm; im.awgn = imp->awgn != 0;
			im.awgn_db = imp->awgn_db; im.seed = imp->seed + (uint64_t)i;
			ref::apply_impairments(s, h->rate, im);
		}
		if ((int64_t)s.size() > stride) return -22;
		for (int64_t n = (int64_t)s.size(); n < stride; ++n) { // zero-fill behind the window
			if (format == OFDMRX_FMT_F32_IQ) { ((float *)samples_out)[2 * ((size_t)i * stride + n)] = 0.f; ((float *)samples_out)[2 * ((size_t)i * stride + n) + 1] = 0.f; }
			else if (format == OFDMRX_FMT_S16_IQ) { ((int16_t *)samples_out)[2 * ((size_t)i * stride + n)] = 0; ((int16_t *)samples_out)[2 * ((size_t)i * stride + n) + 1] = 0; }
			else ((int16_t *)samples_out)[(size_t)i * stride + n] = 0;
		}
		for (size_t n = 0; n < s.size(); ++n) {
			if (format == OFDMRX_FMT_F32_IQ) { ((float *)samples_out)[2 * ((size_t)i * stride + n)] = s[n].re; ((float *)samples_out)[2 * ((size_t)i * stride + n) + 1] = s[n].im; }
			else if (format == OFDMRX_FMT_S16_IQ) { ((int16_t *)samples_out)[2 * ((size_t)i * stride + n)] = ref::quantize16(s[n].re); ((int16_t *)samples_out)[2 * ((size_t)i * stride + n) + 1] = ref::quantize16(s[n].im); }
			else ((int16_t *)samples_out)[(size_t)i * stride + n] = ref::quantize16(s[n].re);
		}
		if (n_samples_out) n_samples_out[i] = (int32_t)s.size();
	}
	return 0;
}
int ofdmtx_get_code(ofdmtx_t *, int, int, uint32_t *) { return -38; }
int ofdmtx_last_launches(ofdmtx_t *) { return 0; }

} // extern "C"
