// modem_b200/csrc/tx_tables.h — host-side constants of the stimulus generator (stimulus.cuh): argument rules of the
// reference's `encode` command line and the frequency-domain values of the three frame-constant symbols.
#pragma once
#include "host_tables.h"

namespace ofdmrx {

struct TxCarriers { int first, step, count; }; // occupied carriers: signed index first + step * c

// encode.cc:319-335: base-37 call sign (" 0-9A-Z", case-insensitive), -1 on a character outside the alphabet
long long base37_encode(const char *str);
// encode.cc:345-397: mode 6..13, call sign in (0, 37^9), offset within the band limits of the mode and a multiple of 50 Hz
bool tx_check_args(int rate, int channels, int freq_off, int mode, long long call_sign);
int tx_band_width(int mode); // encode.cc:363-387
// Occupied-carrier values (re, im pairs, 512 slots each) of [0] the pilot block (encode.cc:132-141), [1] the Schmidl-Cox
// symbol (:142-154) and [2] the metadata symbol for md = (call_sign << 8) | mode (:155-179).
void tx_common_symbols(int rate, int mode, int freq_off, long long call_sign, float *values /*3 * 512 * 2*/, TxCarriers spec[3]);
// raised-cosine cross-fade weights of the guard interval (encode.cc:110-112)
std::vector<float> tx_guard_ramp(int guard_len);
// sample frames of a window holding `frames` back-to-back frames: 1 s + pilot + frames * (3 + rows) symbols + zero symbol + 1 s
long long tx_window_len(int rate, int mode, int frames);

} // namespace ofdmrx
