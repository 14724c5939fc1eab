// modem_b200/csrc/common.cuh — shared device helpers (sm_100a only; no fallback paths).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include "host_tables.h"

namespace ofdmrx {

#define OFDMRX_CUDA_TRY(expr)                                                                        \
	do {                                                                                             \
		cudaError_t e__ = (expr);                                                                    \
		if (e__ != cudaSuccess) {                                                                    \
			std::fprintf(stderr, "ofdmrx: %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return -(int)e__ - 1000;                                                                 \
		}                                                                                            \
	} while (0)

// Function attributes (dynamic shared-memory limits) are per device.  Launchers set them on first use of a device, once,
// under std::call_once: a second host thread (another handle on the same device) blocks until the attribute is in place
// instead of launching ahead of it, and the attribute call's error is kept and returned to every caller.
struct DeviceOnce {
	std::once_flag flag[64];
	cudaError_t err[64];
};
// max_carveout: also ask for the largest shared-memory carve-out.  Without the hint the driver picks the L1 / shared split of
// a launch from the kernel's needs AND from what the SM was last configured for; a kernel whose CTAs per SM are set by shared
// memory (k_theil_sen: six CTAs need 222 KB) then runs with fewer resident CTAs after some predecessors than after others.
template <typename Kernel>
inline cudaError_t set_dynamic_smem_once(DeviceOnce &once, Kernel kernel, int bytes, bool max_carveout = false)
{
	int d = 0;
	cudaGetDevice(&d);
	d &= 63;
	std::call_once(once.flag[d], [&] {
		once.err[d] = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		if (once.err[d] == cudaSuccess && max_carveout)
			once.err[d] = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
	});
	return once.err[d];
}

typedef float2 cfx;
__host__ __device__ __forceinline__ cfx cmul(cfx a, cfx b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ cfx cmulc(cfx a, cfx b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); } // a * conj(b)
__host__ __device__ __forceinline__ cfx cadd(cfx a, cfx b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cfx csub(cfx a, cfx b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float cnorm(cfx a) { return a.x * a.x + a.y * a.y; }
// decode.cc:62-70 / 227-235: differential demodulation with erasure
__device__ __forceinline__ cfx demod_or_erase(cfx curr, cfx prev)
{
	float n = cnorm(prev);
	if (!(n > 0.f)) return make_float2(0.f, 0.f);
	cfx q = cmulc(curr, prev);
	q.x = __fdiv_rn(q.x, n);
	q.y = __fdiv_rn(q.y, n);
	if (!(cnorm(q) <= 4.f)) return make_float2(0.f, 0.f);
	return q;
}
// exp(j * 2*pi * frac(turns)) with the phase kept in double so that sample indices ~1e5 do not cost precision
__device__ __forceinline__ cfx phasor_turns(double turns)
{
	double fr = turns - rint(turns);
	float s, c;
	sincospif(2.f * (float)fr, &s, &c);
	return make_float2(c, s);
}

// per-frame status written by the device (mirrors ofdmrx_frame_status in include/ofdmrx.h)
struct FrameState {
	int32_t status;       // OFDMRX_ST_*
	int32_t detections;
	int32_t t_fire, symbol_pos, sc_pos, index_max, shift, pos_err;
	float timing_max, frac_cfo, cfo_rad;
	int32_t osd_unique, mode;
	uint32_t md_lo, md_hi;
	int32_t best_lane, flips;
	float metrics[8];
	int32_t osd_visited;
	int32_t ts_sweeps;    // pair sweeps the Theil-Sen search took, summed over the window's rows
	int32_t det_overflow; // 1: more trigger edges than the window's detection list holds (later detections were not examined)
};
static_assert(sizeof(FrameState) == 112, "FrameState layout");

enum { ST_OK = 0, ST_NO_SYNC = 1, ST_OSD_FAIL = 2, ST_HDR_CRC = 3, ST_BAD_MODE = 4, ST_BAD_CALL = 5, ST_PAYLOAD_CRC = 6, ST_UNSUPPORTED_MODE = 7 };

} // namespace ofdmrx
