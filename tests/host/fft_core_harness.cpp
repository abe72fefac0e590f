// TEST INFRASTRUCTURE: drives the product's FFT core (spectraldns_b200/csrc/fft_core.cuh) on the CPU.
// The P = N/E "threads" of a line run in lockstep: every stage is executed for all threads before the exchange
// (scatter for all, then gather for all), which is what the barriers of the kernels enforce.  For every compiled
// transform length, every elements-per-thread choice the kernels can make, both directions and the three element
// types (float2, double2, float2x2 = two columns per thread) the result is compared with a long-double DFT.
// Output: one line per case "type N E dir relerr".
#define SDNS_HOST_SHIM
#include "../../spectraldns_b200/csrc/fft_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <array>
#include <complex>
using namespace sdns;

template <typename V> struct Acc;
template <> struct Acc<float2>   { static constexpr int cols = 1; static void set(float2& v, int, double re, double im) { v.x = (float)re; v.y = (float)im; }
                                   static std::complex<double> get(const float2& v, int) { return {v.x, v.y}; } static const char* name() { return "float2"; } };
template <> struct Acc<double2>  { static constexpr int cols = 1; static void set(double2& v, int, double re, double im) { v.x = re; v.y = im; }
                                   static std::complex<double> get(const double2& v, int) { return {v.x, v.y}; } static const char* name() { return "double2"; } };
template <> struct Acc<float2x2> { static constexpr int cols = 2;
                                   static void set(float2x2& v, int c, double re, double im) { (c ? v.b : v.a) = make_float2((float)re, (float)im); }
                                   static std::complex<double> get(const float2x2& v, int c) { const float2& w = c ? v.b : v.a; return {w.x, w.y}; }
                                   static const char* name() { return "float2x2"; } };

struct Map { int operator()(int i) const { return i; } };

template <typename T, int N, int E, int DIR, int Ns, typename V, typename W>
void host_stages(std::vector<std::array<V, E>>& X, const W* tw, std::vector<V>& sm) {
    constexpr int P = N / E;
    if constexpr (Ns < N) {
        constexpr int R = pick_radix(N / Ns, E);
        static_assert(R > 0, "no radix");
        for (int t = 0; t < P; ++t) fft_stage<T, N, E, DIR, Ns, R>(*reinterpret_cast<V(*)[E]>(X[t].data()), t, tw);
        if constexpr (Ns * R < N) {
            Map map;
            for (int t = 0; t < P; ++t) fft_scatter<T, N, E, Ns, R>(*reinterpret_cast<V(*)[E]>(X[t].data()), t, sm.data(), map);
            for (int t = 0; t < P; ++t) fft_gather<T, N, E>(*reinterpret_cast<V(*)[E]>(X[t].data()), t, sm.data(), map);
            host_stages<T, N, E, DIR, Ns * R>(X, tw, sm);
        }
    }
}

template <typename V, int N, int E, int DIR>
void run_case() {
    typedef typename Elt<V>::T T;
    typedef typename Elt<V>::C W;
    constexpr int P = N / E;
    constexpr int C = Acc<V>::cols;
    std::vector<W> tw(N);
    for (int j = 0; j < N; ++j) { long double a = -2.0L * 3.14159265358979323846264338327950288L * j / N; tw[j].x = (T)cosl(a); tw[j].y = (T)sinl(a); }
    std::vector<std::complex<double>> in(C * N);
    srand(1234 + N + 7 * E);
    for (auto& z : in) z = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
    std::vector<std::array<V, E>> X(P);
    for (int t = 0; t < P; ++t) for (int q = 0; q < E; ++q) for (int c = 0; c < C; ++c)
        Acc<V>::set(X[t][q], c, in[c * N + t + q * P].real(), in[c * N + t + q * P].imag());
    std::vector<V> sm(N);
    host_stages<T, N, E, DIR, 1>(X, tw.data(), sm);
    long double err = 0, nrm = 0;
    std::vector<std::complex<long double>> wl(N);
    for (int m = 0; m < N; ++m) { long double a = DIR * 2.0L * 3.14159265358979323846264338327950288L * m / N; wl[m] = {cosl(a), sinl(a)}; }
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < N; ++k) {
            std::complex<long double> s = 0;
            for (int j = 0; j < N; ++j)
                s += std::complex<long double>(in[c * N + j].real(), in[c * N + j].imag()) * wl[(long long)j * k % N];
            std::complex<double> g = Acc<V>::get(X[k % P][k / P], c);
            err += std::norm(std::complex<long double>(g.real(), g.imag()) - s); nrm += std::norm(s);
        }
    printf("%s %d %d %d %.3Le\n", Acc<V>::name(), N, E, DIR, sqrtl(err / nrm));
}

template <typename V, int N, int E>
void maybe() {
    if constexpr (N % E == 0 && N / E >= 1 && N / E <= 128 && plan_ok(N, E)) { run_case<V, N, E, -1>(); run_case<V, N, E, +1>(); }
}
template <typename V, int N>
void all_E() { maybe<V, N, 4>(); maybe<V, N, 6>(); maybe<V, N, 8>(); maybe<V, N, 12>(); maybe<V, N, 16>(); maybe<V, N, 24>(); maybe<V, N, 32>(); }

int main(int argc, char** argv) {
    const int part = argc > 1 ? atoi(argv[1]) : 0;      // 0: float2, 1: double2, 2: float2x2
#define X(N) if (part == 0) all_E<float2, N>(); if (part == 1) all_E<double2, N>(); if (part == 2) all_E<float2x2, N>();
    X(8) X(12) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
    X(24) X(48) X(96) X(192) X(384) X(768) X(1536) X(3072)
#undef X
    // plans the stage rule admits for the reference demo's default grid (60, padded 90): information only
    if (part == 1) { maybe<double2, 60, 30>(); maybe<double2, 90, 30>(); }
    return 0;
}
