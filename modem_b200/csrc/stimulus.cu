// modem_b200/csrc/stimulus.cu — kernels and C-ABI (include/ofdmtx.h) of the device-side stimulus generator: the
// transmitter of /root/reference/encode.cc and the re-specified impairment chain, batched.
// The arithmetic lives in stimulus.cuh (shared with the host-compiled test harness); this file is launch geometry,
// shared-memory carving and scratch management.
//
//   T0 k_tx_code      one CTA per frame      payload -> scramble -> CRC-32 -> systematic polar encode (bit-packed)   encode.cc:293-303
//   T1 k_tx_symbol    one CTA per symbol     prefix product of PSK points -> PAPR clip -> inverse FFT                encode.cc:80-109,304-309
//   T2 k_tx_stream_a  one thread per sample  guard cross-fade gather -> multipath -> CFO [-> noise -> store]         encode.cc:110-131
//   T3 k_tx_stream_b  (only with SFO)        windowed-sinc resampling -> noise -> store
#include "stimulus.cuh"
#include "tx_tables.h"
#include "../../include/ofdmrx.h"
#include "../../include/ofdmtx.h"
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

using namespace ofdmrx;

namespace {

constexpr int kTxCodeThreads = 256;
constexpr int kTxStreamThreads = 256;

__global__ void __launch_bounds__(kTxCodeThreads) k_tx_code(const uint8_t *payloads, const uint32_t *scr, const uint32_t *lut,
	const uint32_t *tbl, uint32_t *code)
{
	__shared__ uint32_t mesg[kTxMesgWords], cw[kTxCodeWords], slut[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) slut[i] = lut[i];
	__syncthreads();
	const size_t f = blockIdx.x;
	tx_code_core(payloads + f * kDataBytes, scr, slut, tbl, tbl + 2048, mesg, cw, code + f * kTxCodeWords, threadIdx.x, blockDim.x);
}

template <int N> constexpr int tx_symbol_threads() { return N / 4 <= 1024 ? ((N / 4 + 31) / 32) * 32 : 1024; }
template <int N> constexpr size_t tx_symbol_smem() { return sizeof(cfx) * (2 * (size_t)N + 2 * kTxMaxCarriers); }

// blockIdx.x: common != 0: one of the three frame-constant symbols; else frame * rows + row
template <int N>
__global__ void __launch_bounds__(tx_symbol_threads<N>()) k_tx_symbol(TxParams p, const uint32_t *code, int common)
{
	extern __shared__ float4 tx_smem[];
	cfx *b0 = reinterpret_cast<cfx *>(tx_smem), *b1 = b0 + N, *car = b1 + N, *acc = car + kTxMaxCarriers;
	TxSpec sp;
	bool papr = true;
	cfx *out;
	if (common) {
		const int s = blockIdx.x;
		sp = p.spec[s];
		for (int c = threadIdx.x; c < sp.count; c += blockDim.x) car[c] = p.common_fdom[s * kTxMaxCarriers + c];
		papr = s != kTxSymSc; // encode.cc:153: symbol(false)
		out = p.tdom_common + (size_t)s * N;
	} else {
		const size_t f = blockIdx.x / p.rows;
		const int row = blockIdx.x - (int)f * p.rows;
		sp = TxSpec{p.code_off, 1, p.cols};
		const uint32_t *cw = code + f * kTxCodeWords;
		for (int c = threadIdx.x; c < p.cols; c += blockDim.x)
			car[c] = tx_data_carrier(cw, p.cols, p.mod_bits, row, c, p.common_fdom[kTxSymPilot * kTxMaxCarriers + c]);
		out = p.tdom + (f * p.rows + row) * N;
	}
	__syncthreads();
	tx_symbol_core<N>(car, acc, sp, papr, b0, b1, p.tw_sym, p.tw_4n, out, threadIdx.x, blockDim.x);
}

// one thread per sample frame of the output pitch; iq_a != nullptr: store the channel output for the resampler instead
__global__ void __launch_bounds__(kTxStreamThreads) k_tx_stream_a(TxParams p, TxImpair im, int has_imp, long long stride, int tiles,
	cfx *iq_a, void *out, int format)
{
	const long long w = blockIdx.x / tiles;
	const long long n = (long long)(blockIdx.x - w * tiles) * kTxStreamThreads + threadIdx.x;
	if (iq_a) {
		if (n < p.len) iq_a[w * p.len + n] = tx_channel_sample(p, im, w, n);
		return;
	}
	if (n >= stride) return;
	cfx v = make_float2(0.f, 0.f);
	if (n < p.len) {
		v = has_imp ? tx_channel_sample(p, im, w, n) : tx_stream_sample(p, w, n);
		if (has_imp && im.awgn) { const cfx z = tx_noise(im, w, n); v.x += z.x; v.y += z.y; }
	}
	tx_store((char *)out + (size_t)w * stride * (format == 0 ? 2 : format == 1 ? 4 : 8), format, n, v);
}

__global__ void __launch_bounds__(kTxStreamThreads) k_tx_stream_b(TxImpair im, long long len, long long nout, long long stride, int tiles,
	const cfx *iq_a, void *out, int format)
{
	const long long w = blockIdx.x / tiles;
	const long long n = (long long)(blockIdx.x - w * tiles) * kTxStreamThreads + threadIdx.x;
	if (n >= stride) return;
	cfx v = make_float2(0.f, 0.f);
	if (n < nout) {
		v = tx_resample(iq_a + w * len, len, im.sfo_ppm, n);
		if (im.awgn) { const cfx z = tx_noise(im, w, n); v.x += z.x; v.y += z.y; }
	}
	tx_store((char *)out + (size_t)w * stride * (format == 0 ? 2 : format == 1 ? 4 : 8), format, n, v);
}

template <int N>
cudaError_t launch_tx_symbol(const TxParams &p, const uint32_t *code, int n_blocks, int common, cudaStream_t s)
{
	static DeviceOnce once;
	if (cudaError_t e = set_dynamic_smem_once(once, k_tx_symbol<N>, (int)tx_symbol_smem<N>())) return e;
	k_tx_symbol<N><<<n_blocks, tx_symbol_threads<N>(), tx_symbol_smem<N>(), s>>>(p, code, common);
	return cudaGetLastError();
}

} // namespace

struct ofdmtx_handle {
	int device = 0, rate = 8000, max_windows = 0, fpw = 1;
	int launches = 0;
	uint32_t *d_tbl[2] = {}, *d_scr = nullptr, *d_lut = nullptr;
	cfx *d_tw = nullptr, *d_tw4 = nullptr, *d_common_fdom = nullptr, *d_tdom_common = nullptr;
	float *d_ramp = nullptr;
	uint8_t *d_payload = nullptr;
	uint32_t *d_code = nullptr;
	// grow-only scratch sized by the mode / stride of the calls seen so far
	cfx *d_tdom = nullptr; size_t tdom_elems = 0;
	cfx *d_iq = nullptr; size_t iq_elems = 0;
	void *d_out = nullptr; size_t out_bytes = 0;
	int last_chunk_frames = 0;
};

namespace {

template <typename T>
int tx_upload(T **p, const void *src, size_t count)
{
	OFDMRX_CUDA_TRY(cudaMalloc((void **)p, count * sizeof(T)));
	OFDMRX_CUDA_TRY(cudaMemcpy(*p, src, count * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}
template <typename T>
int tx_grow(T **p, size_t *have, size_t want)
{
	if (*have >= want) return 0;
	if (*p) { OFDMRX_CUDA_TRY(cudaFree(*p)); *p = nullptr; *have = 0; }
	OFDMRX_CUDA_TRY(cudaMalloc((void **)p, want * sizeof(T)));
	*have = want;
	return 0;
}

} // namespace

extern "C" {

int64_t ofdmtx_call_sign(const char *str) { return str ? base37_encode(str) : -1; }

int64_t ofdmtx_window_samples(int rate_hz, int mode, int frames_per_window)
{
	if ((rate_hz != 8000 && rate_hz != 16000 && rate_hz != 44100 && rate_hz != 48000) || mode < 6 || mode > 13 || frames_per_window < 1) return -22;
	return tx_window_len(rate_hz, mode, frames_per_window);
}

int ofdmtx_create(ofdmtx_t **out, int device, int rate_hz, int max_windows, int frames_per_window)
{
	if (!out || (rate_hz != 8000 && rate_hz != 16000 && rate_hz != 44100 && rate_hz != 48000) || max_windows < 1 || frames_per_window < 1) return -22;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
		std::fprintf(stderr, "ofdmtx: no CUDA device %d (there is no CPU fallback)\n", device);
		return -19;
	}
	OFDMRX_CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	OFDMRX_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		std::fprintf(stderr, "ofdmtx: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
		return -19;
	}
	ofdmtx_handle *h = new (std::nothrow) ofdmtx_handle;
	if (!h) return -12;
	h->device = device;
	h->rate = rate_hz;
	h->max_windows = max_windows;
	h->fpw = frames_per_window;
	const int sym_len = (1280 * rate_hz) / 8000;
	int r = 0;
	for (int tb = 0; tb < 2 && !r; ++tb) { // frozen set | free positions before each word, as the receiver lays them out
		std::vector<uint32_t> tbl = make_frozen(kCodeOrder, tb ? 64512 : 64800, kCrcBits);
		tbl.resize(4096);
		uint32_t acc = 0;
		for (int w = 0; w < 2048; ++w) { tbl[2048 + w] = acc; acc += 32 - __builtin_popcount(tbl[w]); }
		r = tx_upload(&h->d_tbl[tb], tbl.data(), tbl.size());
	}
	std::vector<uint32_t> scr(kDataBytes / 4, 0);
	{
		uint32_t y = 2463534242u; // CODE::Xorshift32 (encode.cc:417-419)
		for (int i = 0; i < kDataBytes; ++i) {
			y ^= y << 13; y ^= y >> 17; y ^= y << 5;
			scr[i / 4] |= (uint32_t)(y & 255u) << (8 * (i % 4));
		}
	}
	uint32_t lut[256];
	crc32_table(0xD419CC15u, lut); // encode.cc:273
	std::vector<float> tw = twiddles(sym_len, -1), tw4 = twiddles(4 * sym_len, -1), ramp = tx_guard_ramp(sym_len / 8);
	if (!r) r = tx_upload(&h->d_scr, scr.data(), scr.size());
	if (!r) r = tx_upload(&h->d_lut, lut, (size_t)256);
	if (!r) r = tx_upload(&h->d_tw, tw.data(), (size_t)sym_len);
	if (!r) r = tx_upload(&h->d_tw4, tw4.data(), (size_t)4 * sym_len);
	if (!r) r = tx_upload(&h->d_ramp, ramp.data(), ramp.size());
	const size_t frames = (size_t)max_windows * frames_per_window;
	if (!r && cudaMalloc((void **)&h->d_common_fdom, sizeof(cfx) * 3 * kTxMaxCarriers) != cudaSuccess) r = -12;
	if (!r && cudaMalloc((void **)&h->d_tdom_common, sizeof(cfx) * 3 * sym_len) != cudaSuccess) r = -12;
	if (!r && cudaMalloc((void **)&h->d_payload, frames * kDataBytes) != cudaSuccess) r = -12;
	if (!r && cudaMalloc((void **)&h->d_code, frames * kTxCodeWords * sizeof(uint32_t)) != cudaSuccess) r = -12;
	if (r) { ofdmtx_destroy(h); return r; }
	*out = h;
	return 0;
}

void ofdmtx_destroy(ofdmtx_t *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	void *ptrs[] = {h->d_tbl[0], h->d_tbl[1], h->d_scr, h->d_lut, h->d_tw, h->d_tw4, h->d_common_fdom, h->d_tdom_common, h->d_ramp,
		h->d_payload, h->d_code, h->d_tdom, h->d_iq, h->d_out};
	for (void *p : ptrs) if (p) cudaFree(p);
	delete h;
}

int ofdmtx_last_launches(ofdmtx_t *h) { return h ? h->launches : -22; }

int ofdmtx_get_code(ofdmtx_t *h, int frame_first, int frame_count, uint32_t *dst)
{
	if (!h || !dst || frame_first < 0 || frame_count < 0 || frame_first + frame_count > h->last_chunk_frames) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	OFDMRX_CUDA_TRY(cudaMemcpy(dst, h->d_code + (size_t)frame_first * kTxCodeWords, (size_t)frame_count * kTxCodeWords * 4, cudaMemcpyDeviceToHost));
	return 0;
}

int ofdmtx_encode_batch(ofdmtx_t *h, const uint8_t *payloads, int payload_mem, int n_windows, int mode, int64_t call_sign,
	int freq_off_hz, const ofdmtx_impairments *imp, void *samples_out, int mem_kind, int format, int64_t stride,
	int32_t *n_samples_out, void *stream)
{
	if (!h || !payloads || !samples_out || n_windows < 0 || format < 0 || format > 2) return -22;
	if (!tx_check_args(h->rate, format == OFDMRX_FMT_S16_MONO ? 1 : 2, freq_off_hz, mode, call_sign)) return -22;
	OFDMRX_CUDA_TRY(cudaSetDevice(h->device));
	cudaStream_t s = (cudaStream_t)stream;
	h->launches = 0;
	const ModeInfo mi = mode_info(mode);
	const int N = (1280 * h->rate) / 8000;
	TxImpair im{};
	const bool has_imp = imp && (imp->multipath || imp->cfo_hz != 0.f || imp->sfo_ppm != 0.f || imp->awgn);
	if (has_imp) { im.multipath = imp->multipath; im.cfo_hz = imp->cfo_hz; im.sfo_ppm = imp->sfo_ppm; im.awgn = imp->awgn; im.awgn_db = imp->awgn_db; im.seed = imp->seed; }
	const bool sfo = has_imp && im.sfo_ppm != 0.f;
	TxParams p{};
	p.rate = h->rate; p.sym_len = N; p.guard_len = N / 8; p.pitch = N + N / 8;
	p.cols = mi.cols; p.mod_bits = mi.mod_bits; p.rows = mi.rows; p.cons_bits = mi.cons_bits; p.table = mi.table;
	p.frames_per_window = h->fpw;
	p.n_sym = 2 + h->fpw * (3 + mi.rows);
	p.len = tx_window_len(h->rate, mode, h->fpw);
	const long long nout = tx_resampled_len(p.len, sfo ? im.sfo_ppm : 0.f);
	if (stride < nout) return -22;
	// frame-constant symbols of this call
	std::vector<float> common((size_t)3 * kTxMaxCarriers * 2);
	TxCarriers spec[3];
	tx_common_symbols(h->rate, mode, freq_off_hz, call_sign, common.data(), spec);
	for (int i = 0; i < 3; ++i) p.spec[i] = TxSpec{spec[i].first, spec[i].step, spec[i].count};
	p.code_off = spec[0].first;
	OFDMRX_CUDA_TRY(cudaMemcpyAsync(h->d_common_fdom, common.data(), common.size() * sizeof(float), cudaMemcpyHostToDevice, s));
	OFDMRX_CUDA_TRY(cudaStreamSynchronize(s)); // `common` is pageable and dies with this scope
	const int chunk = std::min(n_windows, h->max_windows);
	const size_t frames_chunk = (size_t)chunk * h->fpw;
	if (int r = tx_grow(&h->d_tdom, &h->tdom_elems, frames_chunk * mi.rows * N)) return r;
	if (sfo) if (int r = tx_grow(&h->d_iq, &h->iq_elems, (size_t)chunk * p.len)) return r;
	const size_t sample_bytes = format == 0 ? 2 : format == 1 ? 4 : 8;
	if (mem_kind == OFDMRX_MEM_HOST) {
		size_t have = h->out_bytes;
		if (int r = tx_grow((char **)&h->d_out, &have, (size_t)chunk * stride * sample_bytes)) return r;
		h->out_bytes = have;
	}
	p.common_fdom = h->d_common_fdom; p.tw_sym = h->d_tw; p.tw_4n = h->d_tw4; p.ramp = h->d_ramp;
	p.tdom_common = h->d_tdom_common; p.tdom = h->d_tdom;
	cudaError_t e = cudaSuccess;
#define OFDMTX_SYMBOLS(R) e = launch_tx_symbol<Geo<R>::kSymLen>(p, h->d_code, n_blocks, common_flag, s)
	{
		const int n_blocks = 3, common_flag = 1;
		OFDMRX_FOR_RATE(h->rate, OFDMTX_SYMBOLS);
		OFDMRX_CUDA_TRY(e);
		++h->launches;
	}
	const int tiles = (int)((stride + kTxStreamThreads - 1) / kTxStreamThreads), tiles_a = (int)((p.len + kTxStreamThreads - 1) / kTxStreamThreads);
	for (int w0 = 0; w0 < n_windows; w0 += chunk) {
		const int nw = std::min(chunk, n_windows - w0);
		const size_t nf = (size_t)nw * h->fpw;
		const uint8_t *src = payloads + (size_t)w0 * h->fpw * kDataBytes;
		const uint8_t *d_pay = src;
		if (payload_mem == OFDMRX_MEM_HOST) {
			OFDMRX_CUDA_TRY(cudaMemcpyAsync(h->d_payload, src, nf * kDataBytes, cudaMemcpyHostToDevice, s));
			d_pay = h->d_payload;
		}
		k_tx_code<<<(unsigned)nf, kTxCodeThreads, 0, s>>>(d_pay, h->d_scr, h->d_lut, h->d_tbl[mi.table], h->d_code);
		OFDMRX_CUDA_TRY(cudaGetLastError());
		{
			const int n_blocks = (int)(nf * mi.rows), common_flag = 0;
			OFDMRX_FOR_RATE(h->rate, OFDMTX_SYMBOLS);
			OFDMRX_CUDA_TRY(e);
		}
		char *dst = mem_kind == OFDMRX_MEM_HOST ? (char *)h->d_out : (char *)samples_out + (size_t)w0 * stride * sample_bytes;
		TxImpair imw = im;
		imw.window0 = (unsigned long long)w0; // the window index inside the kernels is chunk-relative
		if (sfo) {
			k_tx_stream_a<<<(unsigned)(nw * tiles_a), kTxStreamThreads, 0, s>>>(p, imw, 1, stride, tiles_a, h->d_iq, nullptr, format);
			OFDMRX_CUDA_TRY(cudaGetLastError());
			k_tx_stream_b<<<(unsigned)(nw * tiles), kTxStreamThreads, 0, s>>>(imw, p.len, nout, stride, tiles, h->d_iq, dst, format);
			OFDMRX_CUDA_TRY(cudaGetLastError());
			h->launches += 4;
		} else {
			k_tx_stream_a<<<(unsigned)(nw * tiles), kTxStreamThreads, 0, s>>>(p, imw, has_imp ? 1 : 0, stride, tiles, nullptr, dst, format);
			OFDMRX_CUDA_TRY(cudaGetLastError());
			h->launches += 3;
		}
		if (mem_kind == OFDMRX_MEM_HOST) {
			OFDMRX_CUDA_TRY(cudaMemcpyAsync((char *)samples_out + (size_t)w0 * stride * sample_bytes, h->d_out, (size_t)nw * stride * sample_bytes,
				cudaMemcpyDeviceToHost, s));
			OFDMRX_CUDA_TRY(cudaStreamSynchronize(s)); // the staging buffer is reused by the next chunk
		}
		h->last_chunk_frames = (int)nf;
	}
#undef OFDMTX_SYMBOLS
	if (n_samples_out) for (int i = 0; i < n_windows; ++i) n_samples_out[i] = (int32_t)nout;
	if (payload_mem == OFDMRX_MEM_HOST && mem_kind != OFDMRX_MEM_HOST) OFDMRX_CUDA_TRY(cudaStreamSynchronize(s)); // caller may reuse `payloads`
	return 0;
}

} // extern "C"
