// The host C library's pow() tables, found in the libm mapped into this process, checked, and mirrored on the device (hostpow.cuh).
#include "hostpow.cuh"

#include "common.cuh"

#include <link.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <mutex>
#include <random>
#include <vector>

namespace msm {
namespace {

struct Segment { const unsigned char* p; size_t n; };

int collect_libm(struct dl_phdr_info* info, size_t, void* data) {
    auto* out = static_cast<std::vector<Segment>*>(data);
    const char* name = info->dlpi_name ? info->dlpi_name : "";
    if (!strstr(name, "libm.so") && !strstr(name, "libm-")) return 0;
    for (int i = 0; i < info->dlpi_phnum; ++i) {
        const ElfW(Phdr)& ph = info->dlpi_phdr[i];
        if (ph.p_type != PT_LOAD || !(ph.p_flags & PF_R) || (ph.p_flags & PF_X)) continue;   // read-only data
        out->push_back({reinterpret_cast<const unsigned char*>(info->dlpi_addr + ph.p_vaddr), (size_t)ph.p_memsz});
    }
    return 0;
}

const unsigned char* find_u64s(const Segment& s, const unsigned long long* pat, int n, size_t from = 0) {
    if (s.n < (size_t)n * 8) return nullptr;
    const size_t base = (reinterpret_cast<uintptr_t>(s.p) + 7) & ~(uintptr_t)7;
    for (size_t off = base - reinterpret_cast<uintptr_t>(s.p) + from; off + (size_t)n * 8 <= s.n; off += 8)
        if (!memcmp(s.p + off, pat, (size_t)n * 8)) return s.p + off;
    return nullptr;
}

struct HostTables {
    PowTables t{};
    std::vector<double> logtab;                 // [128][4]
    std::vector<unsigned long long> exptab;     // [256]
    bool ok = false;
    std::string why;
};

// log: {ln2hi, ln2lo} = the two published split constants of ln 2 open pow_log_data, followed by A[0] = -0.5, six more coefficients and
// 128 entries {invc, pad, logc, logctail}. exp: {128 / ln2, 0x1.8p52} open exp_data, followed by -ln2/128 split in two and the
// coefficients C2..C5; the table 2^(i/128) starts (after version-dependent members) with {0, bits(1.0)}.
bool locate(HostTables& H) {
    std::vector<Segment> segs;
    dl_iterate_phdr(collect_libm, &segs);
    if (segs.empty()) { H.why = "libm is not mapped as a shared object"; return false; }
    const unsigned long long log_sig[3] = {0x3fe62e42fefa3800ULL, 0x3d2ef35793c76730ULL, 0xbfe0000000000000ULL};
    const unsigned long long exp_sig[2] = {0x40671547652b82feULL, 0x4338000000000000ULL};
    const unsigned char *lg = nullptr, *ex = nullptr, *et = nullptr;
    for (const Segment& s : segs) {
        if (!lg) {
            const unsigned char* p = find_u64s(s, log_sig, 3);
            if (p && (size_t)(p - s.p) + 0x48 + 128 * 32 <= s.n) lg = p;
        }
        if (!ex) {
            const unsigned char* p = find_u64s(s, exp_sig, 2);
            if (p) {
                const unsigned long long tab_sig[2] = {0ULL, 0x3ff0000000000000ULL};
                // the table follows the header within a few hundred bytes; its second pair must be 2^(1/128) (top bits 0x3fef...)
                for (size_t from = (size_t)(p - s.p) + 0x40; from < (size_t)(p - s.p) + 0x400 && !et;) {
                    const unsigned char* q = find_u64s(s, tab_sig, 2, from);
                    if (!q || (size_t)(q - s.p) + 256 * 8 > s.n) break;
                    unsigned long long third;
                    memcpy(&third, q + 24, 8);
                    if ((third >> 44) == 0x3feffULL) et = q;
                    from = (size_t)(q - s.p) + 8;
                }
                if (et) ex = p;
            }
        }
    }
    if (!lg || !ex || !et) { H.why = "pow tables not found in the mapped libm"; return false; }
    const double* L = reinterpret_cast<const double*>(lg);
    H.t.ln2hi = L[0]; H.t.ln2lo = L[1];
    for (int i = 0; i < 7; ++i) H.t.A[i] = L[2 + i];
    H.logtab.assign(L + 9, L + 9 + 128 * 4);
    const double* E = reinterpret_cast<const double*>(ex);
    H.t.invln2N = E[0]; H.t.shift = E[1]; H.t.negln2hiN = E[2]; H.t.negln2loN = E[3];
    for (int i = 0; i < 4; ++i) H.t.C[i] = E[4 + i];
    H.exptab.assign(reinterpret_cast<const unsigned long long*>(et), reinterpret_cast<const unsigned long long*>(et) + 256);
    H.t.logtab = H.logtab.data();
    H.t.exptab = H.exptab.data();
    return true;
}

bool same(double a, double b) { return hp_bits(a) == hp_bits(b) || (a != a && b != b); }

// std::pow through a volatile pointer: the compiler must call the library, not fold or specialise (pow(x, 2.0) -> x * x)
double (*volatile lib_pow)(double, double) = static_cast<double (*)(double, double)>(std::pow);

bool self_test(HostTables& H) {
    std::mt19937_64 rng(20261018);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    auto rnd_bits = [&] { return hp_double(rng()); };
    const double ys[] = {2.0, 1.0, 0.5, 3.0, 1.5, 1.3, -1.0, -2.0, 0.1, 4.0, 1e-3, 2.5, 7.0, -0.5, 100.0, 1e-70, 1e70, 1023.5, -1074.2, 0.0, -0.0,
                         std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(), std::numeric_limits<double>::quiet_NaN()};
    const double xs[] = {0.0, -0.0, 1.0, -1.0, 2.0, 0.5, -2.0, -0.5, 1e-310, -1e-310, 4.9e-324, 1.7976931348623157e308, 2.2250738585072014e-308,
                         std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(), std::numeric_limits<double>::quiet_NaN(),
                         1.0 + 0x1p-52, 1.0 - 0x1p-53, 3.0, -3.0, 1e-5, 1e5};
    long bad = 0, n = 0;
    auto check = [&](double x, double y) {
        ++n;
        const double a = lib_pow(x, y), b = host_pow(x, y, H.t);
        if (!same(a, b)) {
            if (++bad <= 3 && std::getenv("MSMGPU_DEBUG"))
                std::fprintf(stderr, "[msmgpu] host_pow(%a, %a) = %a, libm says %a\n", x, y, b, a);
        }
    };
    for (double x : xs)
        for (double y : ys) check(x, y);
    for (int i = 0; i < 600000; ++i) {
        // the arguments of the strain energy: stretch / area ratios near 1, small energies, the exponents of the shipped configs
        const double x = 1.0 + std::ldexp(u01(rng), -(int)(rng() % 30)) * ((rng() & 1) ? 1.0 : -0.5);
        check(x, ys[rng() % 6]);
        check(std::ldexp(u01(rng), -(int)(rng() % 60)), ys[rng() % 6]);
        check(std::exp(40.0 * (u01(rng) - 0.5)), 8.0 * (u01(rng) - 0.5));
    }
    for (int i = 0; i < 100000; ++i) {
        check(rnd_bits(), rnd_bits());                                           // anything, incl. NaNs, negatives, huge exponents
        check(std::fabs(rnd_bits()), 2400.0 * (u01(rng) - 0.5));                 // over / underflow and the subnormal range
        check(-std::ldexp(1.0 + u01(rng), (int)(rng() % 40) - 20), (double)((int)(rng() % 41) - 20));   // negative base, integer exponent
    }
    if (std::getenv("MSMGPU_DEBUG")) std::fprintf(stderr, "[msmgpu] pow self-test: %ld of %ld arguments differ from the host library\n", bad, n);
    if (bad) { H.why = std::to_string(bad) + " of " + std::to_string(n) + " self-test arguments differ from the host pow()"; return false; }
    return true;
}

HostTables& host_tables() {
    static HostTables H;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* off = std::getenv("MSMGPU_DEVICE_POW");
        if (off && off[0] == '0') { H.why = "disabled by MSMGPU_DEVICE_POW=0"; return; }
        H.ok = locate(H);
        if (H.ok && std::getenv("MSMGPU_POW_SELFTEST_CORRUPT")) H.logtab[4 * 70 + 3] = -H.logtab[4 * 70 + 3];   // tests: the self-test must notice
        H.ok = H.ok && self_test(H);
        if (!H.ok && std::getenv("MSMGPU_DEBUG")) std::fprintf(stderr, "[msmgpu] device pow disabled: %s\n", H.why.c_str());
    });
    return H;
}

struct PerDevice {
    DevicePow dp;
    double* d_log = nullptr;
    unsigned long long* d_exp = nullptr;
};
std::mutex g_mu;
std::map<int, PerDevice> g_dev;

} // namespace

const PowTables* host_pow_tables() {
    HostTables& H = host_tables();
    return H.ok ? &H.t : nullptr;
}

const DevicePow& device_pow(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_dev.find(device);
    if (it != g_dev.end()) return it->second.dp;
    PerDevice& pd = g_dev[device];
    HostTables& H = host_tables();
    if (!H.ok) return pd.dp;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    bool ok = cudaMalloc((void**)&pd.d_log, H.logtab.size() * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void**)&pd.d_exp, H.exptab.size() * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemcpy(pd.d_log, H.logtab.data(), H.logtab.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(pd.d_exp, H.exptab.data(), H.exptab.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice) == cudaSuccess;
    cudaSetDevice(cur);
    if (!ok) { cudaGetLastError(); return pd.dp; }
    pd.dp.t = H.t;
    pd.dp.t.logtab = pd.d_log;
    pd.dp.t.exptab = pd.d_exp;
    pd.dp.enabled = true;
    return pd.dp;
}

// ---- test / tool entry: pow on the device for n argument pairs (tests compare it with the host's std::pow)
__global__ void k_host_pow(int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out, PowTables T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = host_pow(x[i], y[i], T);
}

} // namespace msm

using namespace msm;

extern "C" {

// 1: the device evaluates pow() with the host library's tables (self-test passed), 0: the host finishes the costs
int msmgpu_device_pow_enabled(void) { return host_pow_tables() != nullptr; }

msmgpu_status msmgpu_debug_device_pow(msmgpu_ctx* ctx, int n, const double* x, const double* y, double* out) {
    if (!ctx || n <= 0 || !x || !y || !out) return fail(MSMGPU_ERR_INVALID, "debug_device_pow: bad arguments");
    MSM_CUDA(cudaSetDevice(ctx->device));
    const DevicePow& dp = device_pow(ctx->device);
    if (!dp.enabled) return fail(MSMGPU_ERR_INVALID, "device pow is not enabled on this host (tables of the C library not found or self-test failed)");
    cudaStream_t s = ctx->stream;
    DevBuf<double> dx, dy, dout;
    MSM_CUDA(dx.alloc(n, s));
    MSM_CUDA(dy.alloc(n, s));
    MSM_CUDA(dout.alloc(n, s));
    MSM_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    MSM_CUDA(cudaMemcpyAsync(dy.p, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    k_host_pow<<<(n + 255) / 256, 256, 0, s>>>(n, dx.p, dy.p, dout.p, dp.t);
    MSM_LAUNCH_CHECK();
    MSM_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MSM_CUDA(cudaStreamSynchronize(s));
    return MSMGPU_OK;
}

} // extern "C"
