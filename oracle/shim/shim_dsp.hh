// oracle/shim/shim_dsp.hh — TEST INFRASTRUCTURE ONLY.
//
// API-compatible stand-ins for the headers of aicodix/dsp that /root/reference/decode.cc:13-28 and encode.cc:11-19 include
// and that are ABSENT from this box (unvendored, unpinned; /root/reference/Makefile:2 `-I../dsp`).  They let the reference's
// OWN translation units be compiled where they lie (oracle/Makefile target `_ref`) so that the oracle's restatement of
// decode.cc / encode.cc — control flow, index arithmetic, constants — can be checked against the real thing.
// The ARITHMETIC behind each class is the oracle's restatement (ref_dsp.hh, "recalled"), wrapped one to one: a test that
// compares oracle/_ref/decode with oracle/build/decode_ref therefore pins the restated decode.cc, NOT the third-party
// primitives — those stay "parity unpinned" (DESIGN.md §1).
#pragma once
#include "../ref_dsp.hh"
#include "../ref_code.hh"
#include <fstream>
#include <initializer_list>
#include <iterator>
#include <memory>
#include <type_traits>

namespace DSP {

template <typename T>
struct Const {
	static constexpr T Pi() { return T(3.14159265358979323846L); }
	static constexpr T TwoPi() { return T(6.28318530717958647692L); }
};

// complex.hh: same operation order as ref::cf (ref_dsp.hh)
template <typename T>
class Complex {
	T re, im;
public:
	typedef T value_type;
	constexpr Complex() : re(0), im(0) {}
	constexpr Complex(T r) : re(r), im(0) {}
	constexpr Complex(T r, T i) : re(r), im(i) {}
	constexpr T real() const { return re; }
	constexpr T imag() const { return im; }
	void real(T r) { re = r; }
	void imag(T i) { im = i; }
	Complex &operator+=(Complex a) { return *this = Complex(re + a.re, im + a.im); }
	Complex &operator-=(Complex a) { return *this = Complex(re - a.re, im - a.im); }
	Complex &operator*=(Complex a) { return *this = Complex(re * a.re - im * a.im, re * a.im + im * a.re); }
	Complex &operator*=(T a) { return *this = Complex(a * re, a * im); }
	Complex &operator/=(T a) { return *this = Complex(re / a, im / a); }
	Complex &operator/=(Complex a)
	{
		Complex n = *this;
		n *= Complex(a.re, -a.im);
		return *this = n /= (a.re * a.re + a.im * a.im);
	}
};
template <typename T> Complex<T> operator+(Complex<T> a, Complex<T> b) { return a += b; }
template <typename T> Complex<T> operator-(Complex<T> a, Complex<T> b) { return a -= b; }
template <typename T> Complex<T> operator-(Complex<T> a) { return Complex<T>(-a.real(), -a.imag()); }
template <typename T> Complex<T> operator*(Complex<T> a, Complex<T> b) { return a *= b; }
template <typename T> Complex<T> operator*(T a, Complex<T> b) { return b *= a; }
template <typename T> Complex<T> operator*(Complex<T> b, T a) { return b *= a; }
template <typename T> Complex<T> operator/(Complex<T> a, T b) { return a /= b; }
template <typename T> Complex<T> operator/(Complex<T> a, Complex<T> b) { return a /= b; }
template <typename T> Complex<T> conj(Complex<T> a) { return Complex<T>(a.real(), -a.imag()); }
template <typename T> T norm(Complex<T> a) { return a.real() * a.real() + a.imag() * a.imag(); }
template <typename T> T arg(Complex<T> a) { return std::atan2(a.imag(), a.real()); }
template <typename T> T abs(Complex<T> a) { return std::sqrt(norm(a)); }
template <typename T> Complex<T> polar(T r, T th) { return Complex<T>(r * std::cos(th), r * std::sin(th)); }

static inline ref::cf to_ref(Complex<float> a) { return ref::cf(a.real(), a.imag()); }
static inline Complex<float> from_ref(ref::cf a) { return Complex<float>(a.re, a.im); }

// utils.hh / decibel.hh
template <typename A, typename B> A lerp(A a, A b, B x) { return (B(1) - x) * a + x * b; }
template <typename T> T decibel(T v) { return T(10) * std::log10(v); }

// fft.hh: unnormalised, X[k] = sum x[n] exp(SIGN j 2 pi n k / N)
template <int N, typename TYPE, int SIGN>
class FastFourierTransform {
	static_assert(sizeof(TYPE) == sizeof(ref::cf), "fp32 complex only");
	ref::FFT fft_;
public:
	FastFourierTransform() : fft_(N, SIGN) {}
	void operator()(TYPE *out, const TYPE *in)
	{
		std::vector<ref::cf> a(N), b(N);
		for (int i = 0; i < N; ++i) a[i] = to_ref(in[i]);
		fft_(b.data(), a.data());
		for (int i = 0; i < N; ++i) out[i] = from_ref(b[i]);
	}
};

// sma.hh: sliding sum over the last NUM inputs, re-summed through a tree (no drift); NORM = divide by NUM
template <typename TYPE, typename VALUE, int NUM, bool NORM>
class SMA4 {
	ref::SlidingSum<TYPE> sum_;
public:
	SMA4() : sum_(NUM) {}
	TYPE operator()(TYPE in)
	{
		TYPE s = sum_(in);
		return NORM ? s / VALUE(NUM) : s;
	}
};

// delay.hh
template <typename TYPE, int NUM>
class Delay {
	ref::Delay<TYPE> d_;
public:
	Delay() : d_(NUM) {}
	TYPE operator()(TYPE in) { return d_(in); }
};

// trigger.hh
template <typename TYPE>
class SchmittTrigger {
	ref::SchmittTrigger t_;
public:
	SchmittTrigger(TYPE low, TYPE high) : t_(low, high) {}
	bool operator()(TYPE in) { return t_(in); }
};
class FallingEdgeTrigger {
	ref::FallingEdge f_;
public:
	bool operator()(bool in) { return f_(in); }
};

// bip_buffer.hh: the last NUM inputs, contiguous, oldest first
template <typename TYPE, int NUM>
class BipBuffer {
	std::vector<TYPE> ring_;
	int pos_ = 0;
public:
	BipBuffer() : ring_(2 * NUM) {}
	const TYPE *operator()(TYPE in)
	{
		ring_[pos_] = in;
		ring_[pos_ + NUM] = in;
		pos_ = (pos_ + 1) % NUM;
		return &ring_[pos_];
	}
};

// theil_sen.hh
template <typename TYPE, int MAX>
class TheilSenEstimator {
	ref::TheilSen ts_;
public:
	void compute(const TYPE *x, const TYPE *y, int len) { ts_.compute(x, y, len); }
	TYPE slope() const { return ts_.slope; }
	TYPE yint() const { return ts_.yint; }
	TYPE operator()(TYPE x) const { return ts_(x); }
};

// blockdc.hh
template <typename TYPE, typename VALUE>
class BlockDC {
	ref::BlockDC b_;
public:
	void samples(int s) { b_.samples(s); }
	TYPE operator()(TYPE in) { return b_(in); }
};

// hilbert.hh
template <typename TYPE, int TAPS>
class Hilbert {
	ref::Hilbert h_;
public:
	Hilbert() : h_(TAPS) {}
	TYPE operator()(typename TYPE::value_type in) { return from_ref(h_(in)); }
};

// phasor.hh
template <typename TYPE>
class Phasor {
	ref::Phasor p_;
public:
	void omega(typename TYPE::value_type v) { p_.omega(v); }
	TYPE operator()() { return from_ref(p_()); }
};

// pcm.hh / wav.hh
template <typename TYPE>
struct ReadPCM {
	virtual ~ReadPCM() = default;
	virtual bool good() = 0;
	virtual void read(TYPE *, int, int = -1) = 0;
	virtual void skip(int) = 0;
	virtual int rate() = 0;
	virtual int channels() = 0;
};
template <typename TYPE>
struct WritePCM {
	virtual ~WritePCM() = default;
	virtual void write(const TYPE *, int, int = -1) = 0;
	virtual void silence(int) = 0;
	virtual int rate() = 0;
	virtual int channels() = 0;
};
template <typename TYPE>
class ReadWAV : public ReadPCM<TYPE> {
	ref::WavData w_;
	size_t pos_ = 0;
	bool good_ = false;
public:
	explicit ReadWAV(const char *name)
	{
		std::ifstream in(name, std::ios::binary);
		std::vector<uint8_t> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
		good_ = ref::wav_parse(raw.data(), raw.size(), w_);
	}
	bool good() override { return good_; }
	void read(TYPE *buf, int num, int stride = -1) override
	{
		if (stride < 0) stride = w_.channels;
		for (int n = 0; n < num; ++n) {
			if (pos_ < w_.frames()) {
				for (int c = 0; c < w_.channels; ++c) buf[n * stride + c] = w_.samples[pos_ * w_.channels + c];
				++pos_;
			} else { // past the end: zeros and a failed stream, as the oracle defines it (the reference leaves it open)
				for (int c = 0; c < w_.channels; ++c) buf[n * stride + c] = 0;
				good_ = false;
			}
		}
	}
	void skip(int num) override { pos_ += num; }
	int rate() override { return w_.rate; }
	int channels() override { return w_.channels; }
};
template <typename TYPE>
class WriteWAV : public WritePCM<TYPE> {
	std::string name_;
	int rate_, bits_, channels_;
	std::vector<float> samples_;
public:
	WriteWAV(const char *name, int rate, int bits, int channels) : name_(name), rate_(rate), bits_(bits), channels_(channels) {}
	~WriteWAV()
	{
		std::vector<uint8_t> wav = ref::wav_serialize(rate_, bits_, channels_, samples_);
		std::ofstream out(name_, std::ios::binary | std::ios::trunc);
		out.write(reinterpret_cast<const char *>(wav.data()), wav.size());
	}
	void write(const TYPE *buf, int num, int stride = -1) override
	{
		if (stride < 0) stride = channels_;
		for (int n = 0; n < num; ++n)
			for (int c = 0; c < channels_; ++c) samples_.push_back(buf[n * stride + c]);
	}
	void silence(int num) override { samples_.insert(samples_.end(), (size_t)num * channels_, 0.f); }
	int rate() override { return rate_; }
	int channels() override { return channels_; }
};

} // namespace DSP
