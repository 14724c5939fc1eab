// modem_b200/csrc/ofdmrx.cu — C-ABI (include/ofdmrx.h) and per-chunk orchestration of the receive kernels.
//
// Pipeline per chunk of windows (all on one stream, no host synchronisation between stages):
//   K0 frontend (int16 -> analytic IQ)            frontend.cu     decode.cc:294-301
//   K1a timing metric, K1b detection list         frontend.cu     decode.cc:84-108
//   K2 fine sync + header (OSD, CRC-16)           acquire.cu      decode.cc:110-151, 398-447
//   K3/K4 demod + Theil-Sen + LLRs                demod.cu        decode.cc:456-529
//   compaction of header-ok windows -> K5 SCL     polar.cu        decode.cc:530-555, 613-615
#include "common.cuh"
#include "frontend.cuh"
#include "polar.cuh"
#include "../../include/ofdmrx.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <new>

using namespace ofdmrx;

static_assert(sizeof(ofdmrx_frame_status) == sizeof(FrameState), "ABI status struct must mirror the device struct");

struct ofdmrx_handle {
	int device = 0, n_sm = 0;
	int rate = 8000;    // 8000, 16000, 44100 or 48000 (decode.cc:590-606)
	int max_frames = 0, max_samples = 0, iq_len = 0;
	bool keep_taps = false;
	int scl_ctas_per_sm = 0, scl_grid = 0, scl_warps = 0;
	int launches = 0;
	// constants
	uint32_t *d_tbl[2] = {}, *d_scr = nullptr, *d_bch = nullptr; // [code table]: frozen set | message offsets | SCL schedule
	int polar_table = 0; // table used by ofdmrx_polar_decode (option "polar_table")
	uint8_t *d_mls1 = nullptr;
	cfx *d_tw1280 = nullptr, *d_tw640 = nullptr, *d_kern = nullptr;
	FrontendConsts fc;
	std::vector<uint32_t> h_frozen[2], h_ops[2];
	// per-chunk scratch
	void *d_in = nullptr; size_t in_bytes = 0;
	int32_t *d_nsamp = nullptr;
	cfx *d_iq = nullptr;
	float *d_timing = nullptr;
	Detection *d_det = nullptr;
	int32_t *d_detcnt = nullptr, *d_edges = nullptr;
	uint32_t *d_masks = nullptr; int mask_words = 0;
	int det_cap = 16;
	FrameState *d_st = nullptr;
	int8_t *d_soft = nullptr;
	cfx *d_cons_raw = nullptr, *d_cons = nullptr;
	float *d_ts = nullptr, *d_llr = nullptr, *d_y = nullptr;
	int *d_cwlist = nullptr, *d_ncw = nullptr, *d_work = nullptr;
	uint32_t *d_payload = nullptr;
	float *d_A = nullptr; uint32_t *d_B = nullptr;
	uint32_t *d_xbits = nullptr; size_t xbits_frames = 0;
	int last_chunk_frames = 0;
	cudaEvent_t ev[10] = {};
	bool ev_valid = false;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_slice[16] = {}, ev_in_free = nullptr;
	// sub-chunk pipeline: the list decoder of sub-chunk k runs on scl_stream while the front stages of sub-chunk k+1 run on the
	// caller's stream (the decoder is latency-bound and leaves most issue slots free; the front stages are not)
	static constexpr int kMaxSub = 8;
	cudaStream_t scl_stream = nullptr;
	cudaEvent_t ev_front[kMaxSub] = {}, ev_scl_done = nullptr;
	int sub_chunks = 0; // option "sub_chunks": 0 / 1 = one list-decoder launch per chunk (default), 2..8 = pipelined sub-chunks
};

namespace {

template <typename T>
int dev_alloc(T **p, size_t count)
{
	OFDMRX_CUDA_TRY(cudaMalloc((void **)p, count * sizeof(T)));
	return 0;
}
template <typename T>
int dev_upload(T **p, const void *src, size_t count)
{
	OFDMRX_CUDA_TRY(cudaMalloc((void **)p, count * sizeof(T)));
	OFDMRX_CUDA_TRY(cudaMemcpy(*p, src, count * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

int ensure_scl_scratch(ofdmrx_handle *h)
{
	if (h->d_A) return 0;
	int occ = scl_occupancy_ctas_per_sm();
	if (occ < 1) occ = 1;
	// as many one-warp CTAs as the SM holds (21 at 96 registers): the kernel is compiled for at least kSclCtasPerSm = 17, which one
	// pass over 10 000 codewords needs; the extra residency lets batches of up to 12 400 codewords finish in one pass as well
	int want = h->scl_ctas_per_sm > 0 ? h->scl_ctas_per_sm : occ;
	if (const char *e = std::getenv("OFDMRX_SCL_CTAS_PER_SM")) want = std::atoi(e);
	if (want > occ) want = occ;
	if (want < 1) want = 1;
	h->scl_ctas_per_sm = want;
	h->scl_grid = want * h->n_sm;
	// never allocate scratch for more warps than a full chunk can occupy
	int need_warps = (h->max_frames + 3) / 4;
	int warps = h->scl_grid * (kSclThreads / 32);
	while (h->scl_grid > 1 && (h->scl_grid - 1) * (kSclThreads / 32) >= need_warps) { --h->scl_grid; }
	warps = h->scl_grid * (kSclThreads / 32);
	h->scl_warps = warps;
	if (std::getenv("OFDMRX_DEBUG")) std::fprintf(stderr, "ofdmrx: list decoder occupancy %d CTAs per SM, grid %d one-warp CTAs\n", occ, h->scl_grid);
	// per warp: alpha levels 6..13 back to back (2.1 MB; polar.cuh) + the beta words
	if (int r = dev_alloc(&h->d_A, (size_t)warps * kSclWarpFloats)) return r;
	if (int r = dev_alloc(&h->d_B, (size_t)warps * kSclWarpWords)) return r;
	return 0;
}

} // namespace

extern "C" {

const char *ofdmrx_version(void) { return "ofdmrx 0.3 (sm_100a; modes 6-13 @ 8/16/44.1/48 kHz; SCL L=8)"; }

int ofdmrx_create(ofdmrx_t **out, int device, int rate_hz, int max_frames, int max_samples)
{
	if (!out || (rate_hz != 8000 && rate_hz != 16000 && rate_hz != 44100 && rate_hz != 48000) || max_frames < 1 || max_samples < 1) return -22;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
		std::fprintf(stderr, "ofdmrx: no CUDA device %d (there is no CPU fallback)\n", device);
		return -19;
	}
	OFDMRX_CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	OFDMRX_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		std::fprintf(stderr, "ofdmrx: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
		return -19;
	}
	ofdmrx_handle *h = new (std::nothrow) ofdmrx_handle;
	if (!h) return -12;
	h->device = device;
	h->rate = rate_hz;
	h->n_sm = prop.multiProcessorCount;
	h->max_frames = max_frames;
	h->max_samples = max_samples;
	h->iq_len = ((max_samples + 1 + 127) / 128) * 128; // stream steps t = 0..n, padded
	// ---- constant tables
	int fuse = kSclMaxFuse; // F/G chain fusion depth of the SCL schedule (1 = none); OFDMRX_SCL_FUSE overrides for A/B runs
	if (const char *e = std::getenv("OFDMRX_SCL_FUSE")) fuse = std::max(1, std::min(kSclMaxFuse, std::atoi(e)));
	bool r1 = true; // rate-1 attempts (OP_R1); OFDMRX_SCL_R1=0 walks every node (A/B runs)
	if (const char *e = std::getenv("OFDMRX_SCL_R1")) r1 = std::atoi(e) != 0;
	bool sched_ok = true;
	std::vector<uint32_t> msg_off[2];
	for (int tb = 0; tb < 2; ++tb) { // code tables of modes 6..9 and 10..13 (decode.cc:310-311,342-343)
		h->h_frozen[tb] = make_frozen(kCodeOrder, tb ? 64512 : 64800, kCrcBits);
		h->h_ops[tb] = make_scl_schedule(h->h_frozen[tb], kCodeOrder, fuse, true, r1);
		msg_off[tb].resize(2048);
		uint32_t acc = 0;
		for (int w = 0; w < 2048; ++w) { msg_off[tb][w] = acc; acc += 32 - __builtin_popcount(h->h_frozen[tb][w]); }
		bool has_top = false; // the kernel never stores levels 14..16: the generator must have been able to use TOP ops
		for (uint32_t op : h->h_ops[tb]) has_top |= scl_op(op) == OP_TOP;
		if (!has_top) sched_ok = false;
		// lengthen() writes 9000 at code[cons_bits..65535] (demod.cu): that equals decode.cc:245-253 only while those indices are
		// all non-frozen; the sets are recomputed here in long double, so check instead of assuming
		const int cons_bits = tb ? 64512 : 64800;
		for (int i = cons_bits; i < kCodeLen; ++i)
			if ((h->h_frozen[tb][i / 32] >> (i % 32)) & 1u) sched_ok = false;
	}
	if (!sched_ok) {
		std::fprintf(stderr, "ofdmrx: the frozen sets computed on this host do not have the structure the kernels rely on\n");
		delete h;
		return -5;
	}
	std::vector<uint32_t> scr(kDataBytes / 4, 0);
	{
		uint32_t y = 2463534242u; // CODE::Xorshift32 (decode.cc:613-615)
		for (int i = 0; i < kDataBytes; ++i) {
			y ^= y << 13; y ^= y >> 17; y ^= y << 5;
			scr[i / 4] |= (uint32_t)(y & 255u) << (8 * (i % 4));
		}
	}
	std::vector<uint32_t> bch = bch_generator_rows();
	std::vector<uint8_t> mls1 = mls_bits(0b100101011, 255);
	const int sym_len = (1280 * rate_hz) / 8000, half = sym_len / 2, pitch = sym_len + sym_len / 8;
	const int filter_len = (((21 * rate_hz) / 8000) & ~3) | 1;
	std::vector<float> tw1280 = twiddles(sym_len, -1), tw640 = twiddles(half, -1), kern = mls0_kernel(half);
	float reco;
	std::vector<float> imco = hilbert_coeffs(filter_len, &reco);
	h->fc.dc_a = float(2 * pitch - 1) / float(2 * pitch);
	h->fc.dc_b = (1.f + h->fc.dc_a) / 2.f;
	h->fc.reco = reco;
	for (int i = 0; i < kMaxHilbertCoeffs; ++i) h->fc.imco[i] = i < (int)imco.size() ? imco[i] : 0.f;
	int r = 0;
	for (int tb = 0; tb < 2; ++tb) {
		std::vector<uint32_t> tbl(h->h_frozen[tb]);
		tbl.insert(tbl.end(), msg_off[tb].begin(), msg_off[tb].end());
		const std::vector<uint32_t> pieces = crc32_pieces(h->h_frozen[tb], kCrcBits);
		tbl.insert(tbl.end(), pieces.begin(), pieces.end());
		tbl.insert(tbl.end(), h->h_ops[tb].begin(), h->h_ops[tb].end());
		if (!r) r = dev_upload(&h->d_tbl[tb], tbl.data(), tbl.size());
	}
	if (!r) r = dev_upload(&h->d_scr, scr.data(), scr.size());
	if (!r) r = dev_upload(&h->d_bch, bch.data(), bch.size());
	if (!r) r = dev_upload(&h->d_mls1, mls1.data(), mls1.size());
	if (!r) r = dev_upload(&h->d_tw1280, tw1280.data(), (size_t)sym_len);
	if (!r) r = dev_upload(&h->d_tw640, tw640.data(), (size_t)half);
	if (!r) r = dev_upload(&h->d_kern, kern.data(), (size_t)half);
	// ---- per-chunk scratch
	const size_t F = (size_t)max_frames;
	h->in_bytes = F * (size_t)max_samples * sizeof(cfx); // large enough for any supported input format
	if (!r) r = dev_alloc((char **)&h->d_in, F * (size_t)max_samples * 4); // int16 IQ is the widest staged host format (float2 is staged in two halves? no: see decode)
	if (!r) r = dev_alloc(&h->d_nsamp, F);
	if (!r) r = dev_alloc(&h->d_iq, F * (size_t)h->iq_len);
	if (!r) r = dev_alloc(&h->d_timing, F * (size_t)h->iq_len);
	h->mask_words = h->iq_len / 32 + 256; // one bit per stream step, rounded up to whole correlator tiles (<= 8192 steps each)
	if (!r) r = dev_alloc(&h->d_masks, F * 2 * (size_t)h->mask_words);
	h->det_cap = std::max(16, max_samples / pitch + 8); // decode.cc:390-448 walks detections without bound: one per symbol pitch is generous
	if (!r) r = dev_alloc(&h->d_det, F * h->det_cap);
	if (!r) r = dev_alloc(&h->d_edges, F * (2 * (size_t)h->det_cap + 2));
	if (!r) r = dev_alloc(&h->d_detcnt, F);
	if (!r) r = dev_alloc(&h->d_st, F);
	if (!r) r = dev_alloc(&h->d_soft, F * 256);
	if (!r) r = dev_alloc(&h->d_llr, F * (size_t)kCodeLen);
	if (!r) r = dev_alloc(&h->d_cons_raw, F * kMaxCons);
	if (!r) r = dev_alloc(&h->d_y, F * kMaxCons);
	if (!r) r = dev_alloc(&h->d_ts, F * kMaxRows * 3);
	if (!r) r = dev_alloc(&h->d_cwlist, F + 4 * ofdmrx_handle::kMaxSub);
	if (!r) r = dev_alloc(&h->d_ncw, (size_t)2 * ofdmrx_handle::kMaxSub);
	if (!r) r = dev_alloc(&h->d_work, (size_t)ofdmrx_handle::kMaxSub);
	if (!r) r = dev_alloc(&h->d_payload, F * (size_t)(kDataBytes / 4));
	for (int i = 0; i < 10 && !r; ++i) if (cudaEventCreate(&h->ev[i]) != cudaSuccess) r = -12;
	for (int i = 0; i < 16 && !r; ++i) if (cudaEventCreateWithFlags(&h->ev_slice[i], cudaEventDisableTiming) != cudaSuccess) r = -12;
	if (!r && cudaEventCreateWithFlags(&h->ev_in_free, cudaEventDisableTiming) != cudaSuccess) r = -12;
	if (!r && cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) r = -12;
	if (!r && cudaStreamCreateWithFlags(&h->scl_stream, cudaStreamNonBlocking) != cudaSuccess) r = -12;
	for (int i = 0; i < ofdmrx_handle::kMaxSub && !r; ++i) if (cudaEventCreateWithFlags(&h->ev_front[i], cudaEventDisableTiming) != cudaSuccess) r = -12;
	if (!r && cudaEventCreateWithFlags(&h->ev_scl_done, cudaEventDisableTiming) != cudaSuccess) r = -12;
	if (r) { ofdmrx_destroy(h); return r; }
	h->in_bytes = F * (size_t)max_samples * 4;
	*out = h;
	return 0;
}

void ofdmrx_destroy(ofdmrx_t *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	void *ptrs[] = {h->d_tbl[0], h->d_tbl[1], h->d_scr, h->d_bch, h->d_mls1, h->d_tw1280, h->d_tw640, h->d_kern, h->d_in,
		h->d_nsamp, h->d_iq, h->d_timing, h->d_det, h->d_detcnt, h->d_edges, h->d_masks, h->d_st, h->d_soft, h->d_cons_raw, h->d_cons, h->d_ts, h->d_llr, h->d_y,
		h->d_cwlist, h->d_ncw, h->d_work, h->d_payload, h->d_A, h->d_B, h->d_xbits};
	for (void *p : ptrs) if (p) cudaFree(p);
	for (int i = 0; i < 10; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
	for (int i = 0; i < 16; ++i) if (h->ev_slice[i]) cudaEventDestroy(h->ev_slice[i]);
	if (h->ev_in_free) cudaEventDestroy(h->ev_in_free);
	if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
	if (h->scl_stream) cudaStreamDestroy(h->scl_stream);
	for (int i = 0; i < ofdmrx_handle::kMaxSub; ++i) if (h->ev_front[i]) cudaEventDestroy(h->ev_front[i]);
	if (h->ev_scl_done) cudaEventDestroy(h->ev_scl_done);
	delete h;
}

int ofdmrx_set_option(ofdmrx_t *h, const char *key, int value)
{
	if (!h || !key) return -22;
	if (!std::strcmp(key, "keep_taps")) {
		h->keep_taps = value != 0;
		if (h->keep_taps && !h->d_cons) {
			cudaSetDevice(h->device);
			const size_t F = (size_t)h->max_frames;
			return dev_alloc(&h->d_cons, F * kMaxCons);
		}
		return 0;
	}
	if (!std::strcmp(key, "polar_table")) {
		if (value < 0 || value > 1) return -22;
		h->polar_table = value;
		return 0;
	}
	if (!std::strcmp(key, "sub_chunks")) {
		if (value < 0 || value > ofdmrx_handle::kMaxSub) return -22;
		h->sub_chunks = value;
		return 0;
	}
	if (!std::strcmp(key, "scl_ctas_per_sm")) {
		if (h->d_A) return -16; // scratch already sized
		h->scl_ctas_per_sm = value;
		return 0;
	}
	return -22;
}

int ofdmrx_last_launches(ofdmrx_t *h) { return h ? h->launches : -22; }

namespace {
// roofline denominator of the list decoder (FP32 CUDA-core peak): 16 independent FMA chains per thread
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters)
{
	float a[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) a[k] = (float)(threadIdx.x + k);
	const float m = 0.999f, c = 0.001f * (float)blockIdx.x;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], m, c);
	}
	float s = 0.f;
#pragma unroll
	for (int k = 0; k < 16; ++k) s += a[k];
	if (s == 12345.678f) out[0] = s; // never true: keeps the chains alive
}
} // namespace

int ofdmrx_measure_fp32(ofdmrx_t *h, float *tflops)
{
	if (!h || !tflops) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	cudaEvent_t e0, e1;
	OFDMRX_CUDA_TRY(cudaEventCreate(&e0));
	OFDMRX_CUDA_TRY(cudaEventCreate(&e1));
	const int grid = h->n_sm * 8, iters = 8192;
	float best = 0.f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0, nullptr);
		k_fp32_peak<<<grid, 256>>>((float *)h->d_ncw, iters);
		cudaEventRecord(e1, nullptr);
		OFDMRX_CUDA_TRY(cudaEventSynchronize(e1));
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		const float tf = 2.f * 16.f * (float)iters * 256.f * (float)grid / (ms * 1e-3f) / 1e12f;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	*tflops = best;
	return 0;
}

int ofdmrx_stage_times(ofdmrx_t *h, float *ms, int n)
{
	if (!h || !ms || n < 7) return -22;
	if (!h->ev_valid) return -61;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	OFDMRX_CUDA_TRY(cudaEventSynchronize(h->ev[7]));
	for (int i = 0; i < 7; ++i) OFDMRX_CUDA_TRY(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
	if (n >= 10) { // the three kernels of the demod stage: FFT + differential demodulation, Theil-Sen, soft demapping
		OFDMRX_CUDA_TRY(cudaEventElapsedTime(&ms[7], h->ev[4], h->ev[8]));
		OFDMRX_CUDA_TRY(cudaEventElapsedTime(&ms[8], h->ev[8], h->ev[9]));
		OFDMRX_CUDA_TRY(cudaEventElapsedTime(&ms[9], h->ev[9], h->ev[5]));
	}
	return h->last_chunk_frames;
}

int ofdmrx_get_table(ofdmrx_t *h, int which, void *dst, size_t bytes)
{
	if (!h || !dst) return -22;
	cudaSetDevice(h->device);
	if (which < 0 || which > 3) return -22;
	const int tb = which >> 1;
	const void *src = (which & 1) == 0 ? (const void *)h->d_tbl[tb] : (const void *)(h->d_tbl[tb] + kSclTblOps);
	const size_t have = (which & 1) == 0 ? 2048 * 4 : h->h_ops[tb].size() * 4;
	if (bytes > have) bytes = have;
	OFDMRX_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
	return (int)(have / 4);
}

// front stages (K0..K4) for windows [f0, f0+nf) of the current chunk; d_samples points at window f0
static int run_front(ofdmrx_handle *h, const void *d_samples, int format, int f0, int nf, int64_t stride, const int32_t *d_ns, int n_default,
	int n_max, int skip, cudaStream_t s, bool record)
{
	const size_t L = (size_t)h->iq_len;
	cfx *iq = h->d_iq + (size_t)f0 * L;
	float *timing = h->d_timing + (size_t)f0 * L;
	const int32_t *ns = d_ns ? d_ns + f0 : nullptr;
	if (record) cudaEventRecord(h->ev[0], s);
	OFDMRX_CUDA_TRY(launch_frontend(h->rate, format, d_samples, stride, ns, n_default, nf, iq, h->iq_len, h->iq_len, h->fc, s));
	if (record) cudaEventRecord(h->ev[1], s);
	uint32_t *masks = h->d_masks + (size_t)f0 * 2 * h->mask_words;
	OFDMRX_CUDA_TRY(launch_sync_metric(h->rate, iq, h->iq_len, h->iq_len, ns, n_default, n_max, nf, timing, h->iq_len, masks, h->mask_words, h->keep_taps ? 1 : 0, s));
	if (record) cudaEventRecord(h->ev[2], s);
	OFDMRX_CUDA_TRY(launch_sync_detect(h->rate, timing, h->iq_len, masks, h->mask_words, ns, n_default, nf, h->d_det + (size_t)f0 * h->det_cap, h->d_detcnt + f0, h->det_cap,
		h->d_edges + (size_t)f0 * (2 * h->det_cap + 2), s));
	if (record) cudaEventRecord(h->ev[3], s);
	AcquireConsts ac{h->d_tw1280, h->d_tw640, h->d_kern, h->d_mls1, h->d_bch};
	OFDMRX_CUDA_TRY(launch_acquire(h->rate, iq, h->iq_len, h->iq_len, h->d_det + (size_t)f0 * h->det_cap, h->d_detcnt + f0, h->det_cap, skip, nf, h->d_st + f0,
		h->d_soft + (size_t)f0 * 256, ac, s));
	if (record) cudaEventRecord(h->ev[4], s);
	OFDMRX_CUDA_TRY(launch_demod(h->rate, iq, h->iq_len, h->iq_len, h->d_st + f0, nf, h->d_tw1280, h->d_cons_raw + (size_t)f0 * kMaxCons,
		h->d_y + (size_t)f0 * kMaxCons, h->keep_taps ? h->d_cons + (size_t)f0 * kMaxCons : nullptr, h->d_ts + (size_t)f0 * kMaxRows * 3,
		h->d_llr + (size_t)f0 * kCodeLen, h->n_sm, s, record ? h->ev[8] : nullptr, record ? h->ev[9] : nullptr));
	if (record) cudaEventRecord(h->ev[5], s);
	h->launches += 7;
	return 0;
}

// compaction of the header-ok windows + list decoding of the windows [f0, f0 + nf) of the chunk (sub-chunk `sub`), on stream s
static int run_scl(ofdmrx_handle *h, int f0, int nf, int sub, cudaStream_t s, bool record)
{
	int *cwlist = h->d_cwlist + f0 + 4 * sub, *ncw = h->d_ncw + 2 * sub, *work = h->d_work + sub;
	uint32_t *payload = h->d_payload + (size_t)f0 * (kDataBytes / 4);
	OFDMRX_CUDA_TRY(launch_compact(h->d_st + f0, nf, cwlist, ncw, s));
	OFDMRX_CUDA_TRY(launch_payload_init(payload, h->d_scr, nf, work, s));
	if (int r = ensure_scl_scratch(h)) return r;
	if (record) cudaEventRecord(h->ev[6], s);
	SclParams p{};
	p.llr = h->d_llr + (size_t)f0 * kCodeLen; p.cw_list = cwlist; p.n_cw_ptr = ncw; p.A = h->d_A; p.B = h->d_B;
	p.tbl[0] = h->d_tbl[0]; p.tbl[1] = h->d_tbl[1]; p.work = work;
	p.payload = payload; p.st = h->d_st + f0; p.xbits = nullptr;
	OFDMRX_CUDA_TRY(launch_polar_scl(p, std::min(h->scl_grid, (nf + 3) / 4 + 1), s));
	if (record) cudaEventRecord(h->ev[7], s);
	h->launches += 3;
	return 0;
}

int ofdmrx_decode_batch(ofdmrx_t *h, const void *samples, int mem_kind, int format, int n_frames, int64_t stride,
	const int32_t *n_samples, int skip, uint8_t *payload_out, ofdmrx_frame_status *status_out, void *stream)
{
	if (!h || !samples || n_frames < 0 || stride < 1 || format < 0 || format > 3 || skip < 0 || !payload_out) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	cudaStream_t s = (cudaStream_t)stream;
	h->launches = 0;
	const size_t frame_bytes = (size_t)stride * (format == OFDMRX_FMT_S16_MONO ? 2 : format == OFDMRX_FMT_F32_IQ ? 8 : 4);
	for (int f0 = 0; f0 < n_frames; f0 += h->max_frames) {
		const int nf = std::min(h->max_frames, n_frames - f0);
		const char *src = (const char *)samples + (size_t)f0 * frame_bytes;
		const int32_t *h_ns = n_samples ? n_samples + f0 : nullptr;
		int n_default = (int)std::min<int64_t>(stride, h->max_samples), n_max = n_default;
		const int32_t *d_ns = nullptr;
		if (h_ns) {
			n_max = 0;
			for (int i = 0; i < nf; ++i) {
				if (h_ns[i] < 0 || h_ns[i] > h->max_samples || h_ns[i] > stride) return -22;
				n_max = std::max(n_max, (int)h_ns[i]);
			}
			OFDMRX_CUDA_TRY(cudaMemcpyAsync(h->d_nsamp, h_ns, (size_t)nf * 4, cudaMemcpyHostToDevice, s));
			d_ns = h->d_nsamp;
		}
		// sub-chunks: the list decoder of one runs on its own stream beside the front stages of the next
		// (off unless asked for — measured on 10 000 clean windows: 1 sub-chunk 54.6 ms, 2: 57.5, 4: 70.4, 8: 98.0: a list-decoder launch
		// over a quarter of the codewords takes as long as one over all of them (its warps are latency-bound, the launch ends
		// with its slowest group), and the front kernels' shared memory keeps most of its CTAs from co-residing)
		int subs = h->sub_chunks > 0 ? h->sub_chunks : 1;
		subs = std::min(subs, ofdmrx_handle::kMaxSub);
		const bool piped = subs > 1;
		auto scl_after_front = [&](int a, int m, int sub, bool last) -> int {
			if (!piped) return run_scl(h, a, m, 0, s, true);
			OFDMRX_CUDA_TRY(cudaEventRecord(h->ev_front[sub], s));
			OFDMRX_CUDA_TRY(cudaStreamWaitEvent(h->scl_stream, h->ev_front[sub], 0));
			if (int r = run_scl(h, a, m, sub, h->scl_stream, false)) return r;
			if (last) {
				OFDMRX_CUDA_TRY(cudaEventRecord(h->ev_scl_done, h->scl_stream));
				OFDMRX_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_scl_done, 0));
			}
			return 0;
		};
		if (mem_kind == OFDMRX_MEM_HOST) {
			// host windows: the H2D copy of slice k+1 runs on the copy stream while slice k goes through the front stages
			if ((size_t)nf * frame_bytes > h->in_bytes) { // float2 windows need twice the staging the handle starts with
				OFDMRX_CUDA_TRY(cudaStreamSynchronize(s));
				OFDMRX_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
				if (h->d_in) cudaFree(h->d_in);
				h->d_in = nullptr; h->in_bytes = 0;
				if (int r = dev_alloc((char **)&h->d_in, (size_t)h->max_frames * (size_t)h->max_samples * 8)) return r;
				h->in_bytes = (size_t)h->max_frames * (size_t)h->max_samples * 8;
			}
			// (only the first slice's copy is exposed: many small slices keep that short)
			const int slices = nf >= 8192 ? 16 : nf >= 4096 ? 8 : nf >= 1024 ? 4 : 1;
			const int per = (nf + slices - 1) / slices;
			const int per_sub = std::max(1, slices / subs); // slices per sub-chunk
			OFDMRX_CUDA_TRY(cudaEventRecord(h->ev_in_free, s)); // earlier work on `s` may still read d_in
			OFDMRX_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_in_free, 0));
			for (int k = 0, a = 0; a < nf; ++k, a += per) {
				const int m = std::min(per, nf - a);
				OFDMRX_CUDA_TRY(cudaMemcpyAsync((char *)h->d_in + (size_t)a * frame_bytes, src + (size_t)a * frame_bytes, (size_t)m * frame_bytes,
					cudaMemcpyHostToDevice, h->copy_stream));
				OFDMRX_CUDA_TRY(cudaEventRecord(h->ev_slice[k], h->copy_stream));
			}
			int sub_first = 0, sub = 0;
			for (int k = 0, a = 0; a < nf; ++k, a += per) {
				const int m = std::min(per, nf - a);
				const bool last = a + m >= nf;
				OFDMRX_CUDA_TRY(cudaStreamWaitEvent(s, h->ev_slice[k], 0));
				if (int r = run_front(h, (char *)h->d_in + (size_t)a * frame_bytes, format, a, m, stride, d_ns, n_default, n_max, skip, s, last && !piped)) return r;
				if (!piped) continue;
				if (last || ((k + 1) % per_sub == 0 && sub < subs - 1)) {
					if (int r = scl_after_front(sub_first, a + m - sub_first, sub, last)) return r;
					sub_first = a + m; ++sub;
				}
			}
			if (!piped) { if (int r = run_scl(h, 0, nf, 0, s, true)) return r; }
		} else {
			const int per = (nf + subs - 1) / subs;
			for (int sub = 0, a = 0; a < nf; ++sub, a += per) {
				const int m = std::min(per, nf - a);
				if (int r = run_front(h, src + (size_t)a * frame_bytes, format, a, m, stride, d_ns, n_default, n_max, skip, s, !piped)) return r;
				if (int r = scl_after_front(a, m, sub, a + m >= nf)) return r;
			}
		}
		h->ev_valid = !piped;
		h->last_chunk_frames = nf;
		const cudaMemcpyKind k = mem_kind == OFDMRX_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
		OFDMRX_CUDA_TRY(cudaMemcpyAsync(payload_out + (size_t)f0 * kDataBytes, h->d_payload, (size_t)nf * kDataBytes, k, s));
		if (status_out) OFDMRX_CUDA_TRY(cudaMemcpyAsync(status_out + f0, h->d_st, (size_t)nf * sizeof(FrameState), k, s));
	}
	if (mem_kind == OFDMRX_MEM_HOST) OFDMRX_CUDA_TRY(cudaStreamSynchronize(s));
	return 0;
}

int ofdmrx_polar_decode(ofdmrx_t *h, const float *llr, int n, uint8_t *payload_out, ofdmrx_frame_status *status_out, uint32_t *xbits)
{
	if (!h || !llr || n < 0 || !payload_out) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	h->launches = 0;
	if (int r = ensure_scl_scratch(h)) return r;
	cudaStream_t s = nullptr;
	for (int f0 = 0; f0 < n; f0 += h->max_frames) {
		const int nf = std::min(h->max_frames, n - f0);
		OFDMRX_CUDA_TRY(cudaMemcpyAsync(h->d_llr, llr + (size_t)f0 * kCodeLen, (size_t)nf * kCodeLen * 4, cudaMemcpyHostToDevice, s));
		OFDMRX_CUDA_TRY(cudaMemsetAsync(h->d_st, 0, (size_t)nf * sizeof(FrameState), s));
		OFDMRX_CUDA_TRY(launch_payload_init(h->d_payload, h->d_scr, nf, h->d_work, s));
		if (xbits && h->xbits_frames < (size_t)nf) {
			if (h->d_xbits) cudaFree(h->d_xbits);
			h->d_xbits = nullptr;
			if (int r = dev_alloc(&h->d_xbits, (size_t)nf * 8 * 2048)) return r;
			h->xbits_frames = nf;
		}
		SclParams p{};
		p.llr = h->d_llr; p.cw_list = nullptr; p.n_cw_ptr = nullptr; p.A = h->d_A; p.B = h->d_B;
		p.n_cw[h->polar_table] = nf; p.n_cw[1 - h->polar_table] = 0;
		p.tbl[0] = h->d_tbl[0]; p.tbl[1] = h->d_tbl[1]; p.work = h->d_work;
		p.payload = h->d_payload; p.st = h->d_st;
		p.xbits = xbits ? h->d_xbits : nullptr;
		OFDMRX_CUDA_TRY(launch_polar_scl(p, h->scl_grid, s));
		h->launches += 2;
		OFDMRX_CUDA_TRY(cudaMemcpyAsync(payload_out + (size_t)f0 * kDataBytes, h->d_payload, (size_t)nf * kDataBytes, cudaMemcpyDeviceToHost, s));
		if (status_out) OFDMRX_CUDA_TRY(cudaMemcpyAsync(status_out + f0, h->d_st, (size_t)nf * sizeof(FrameState), cudaMemcpyDeviceToHost, s));
		if (xbits) OFDMRX_CUDA_TRY(cudaMemcpyAsync(xbits + (size_t)f0 * 8 * 2048, h->d_xbits, (size_t)nf * 8 * 2048 * 4, cudaMemcpyDeviceToHost, s));
		OFDMRX_CUDA_TRY(cudaStreamSynchronize(s));
	}
	return 0;
}

int ofdmrx_theil_sen(ofdmrx_t *h, const float *y, int n_rows, int cols, float *out3)
{
	if (!h || !y || !out3 || n_rows < 0 || cols < 8 || cols > kMaxCols) return -22;
	if ((size_t)n_rows * cols > (size_t)h->max_frames * kMaxCons || n_rows > h->max_frames * kMaxRows) return -27;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	cudaStream_t s = nullptr;
	OFDMRX_CUDA_TRY(cudaMemcpyAsync(h->d_y, y, (size_t)n_rows * cols * 4, cudaMemcpyHostToDevice, s));
	OFDMRX_CUDA_TRY(cudaMemsetAsync(h->d_ts, 0, (size_t)n_rows * 3 * 4, s));
	OFDMRX_CUDA_TRY(launch_theil_sen_rows(h->d_y, n_rows, cols, h->d_ts, h->n_sm, s));
	OFDMRX_CUDA_TRY(cudaMemcpyAsync(out3, h->d_ts, (size_t)n_rows * 3 * 4, cudaMemcpyDeviceToHost, s));
	OFDMRX_CUDA_TRY(cudaStreamSynchronize(s));
	h->launches = 1;
	return 0;
}

int64_t ofdmrx_tap_elems(ofdmrx_t *h, int stage)
{
	if (!h) return -22;
	switch (stage) {
	case OFDMRX_TAP_IQ: case OFDMRX_TAP_TIMING: return h->iq_len;
	case OFDMRX_TAP_SOFT: return 256;
	case OFDMRX_TAP_CONS_RAW: case OFDMRX_TAP_CONS: return kMaxCons;
	case OFDMRX_TAP_TS: return kMaxRows * 3;
	case OFDMRX_TAP_LLR: return kCodeLen;
	case OFDMRX_TAP_PHASE: return kMaxCons;
	}
	return -22;
}

int ofdmrx_get_taps(ofdmrx_t *h, int stage, int first, int count, void *dst, size_t bytes)
{
	if (!h || !dst || first < 0 || count < 0 || first + count > h->max_frames) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	const void *src = nullptr;
	size_t esz = 0;
	switch (stage) {
	case OFDMRX_TAP_IQ: src = h->d_iq; esz = sizeof(cfx); break;
	case OFDMRX_TAP_TIMING: src = h->d_timing; esz = 4; break;
	case OFDMRX_TAP_SOFT: src = h->d_soft; esz = 1; break;
	case OFDMRX_TAP_CONS_RAW: src = h->d_cons_raw; esz = sizeof(cfx); break;
	case OFDMRX_TAP_CONS: src = h->d_cons; esz = sizeof(cfx); break;
	case OFDMRX_TAP_TS: src = h->d_ts; esz = 4; break;
	case OFDMRX_TAP_LLR: src = h->d_llr; esz = 4; break;
	case OFDMRX_TAP_PHASE: src = h->d_y; esz = 4; break;
	default: return -22;
	}
	if (!src || (stage == OFDMRX_TAP_TIMING && !h->keep_taps)) return -61; // keep_taps was off (the timing metric is then stored only around detections)
	const size_t per = (size_t)ofdmrx_tap_elems(h, stage) * esz;
	if (bytes < per * count) return -27;
	OFDMRX_CUDA_TRY(cudaDeviceSynchronize());
	OFDMRX_CUDA_TRY(cudaMemcpy(dst, (const char *)src + per * first, per * count, cudaMemcpyDeviceToHost));
	return 0;
}

} // extern "C"
