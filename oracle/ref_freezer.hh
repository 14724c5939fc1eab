// oracle/ref_freezer.hh — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Restatement of the frozen-set construction of /root/reference/freezer.cc:14-32 (CODE::PolarCodeConst0 from
// the absent aicodix/code, recalled): binary-erasure-channel evolution in long double, natural index order
// (left child p(2-p) at index i, right child p^2 at index i+h), the K' most reliable indices stay free.
// GOLDEN PIN: for (N,K) = (64800,43072) and (64512,43072) this reproduces BOTH tables of
// /root/reference/polar_tables.hh bit for bit (tests/golden/polar_tables.sha256, tests/test_oracle_kat.py).
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>

namespace ref {

static inline void bec_evolve(std::vector<long double> &prob, long double pe, int i, int h)
{
	if (h) {
		bec_evolve(prob, pe * (2 - pe), i, h / 2);
		bec_evolve(prob, pe * pe, i + h, h / 2);
	} else {
		prob[i] = pe;
	}
}

// order M, transmitted bits N (after shortening), information bits K (payload + CRC)  — freezer.cc:15-26
static inline std::vector<uint32_t> make_frozen_table(int M, int N, int K)
{
	int len = 1 << M;
	long double erasure_probability = (long double)(N - K) / N;
	double design_SNR = 10 * std::log10(-std::log(erasure_probability));
	double better_SNR = design_SNR + 1.59175;
	long double better_probability = std::exp(-std::pow(10.0, better_SNR / 10));
	std::vector<long double> prob(len);
	bec_evolve(prob, better_probability, 0, len / 2);
	int keep = K + len - N;
	std::vector<int> idx(len);
	for (int i = 0; i < len; ++i) idx[i] = i;
	std::nth_element(idx.begin(), idx.begin() + keep, idx.end(), [&](int a, int b) { return prob[a] < prob[b]; });
	std::vector<uint32_t> frozen(len / 32, 0);
	for (int i = keep; i < len; ++i) frozen[idx[i] / 32] |= 1u << (idx[i] % 32);
	return frozen;
}

} // namespace ref
