// tools/sincos_exhaustive.cpp -- TEST TOOL. Compiles playrho_b200/csrc/pr_math.cuh for the host (-DPR_HOST_EMUL, -ffp-contract=off:
// the same IEEE operations the device executes under --fmad=false) and compares pr::glibc_sincosf -- the restatement of glibc
// 2.39's sincosf that every UnitVec::Get(angle) of the device path goes through (third-party arithmetic) -- with
// this image's libm sincosf over ALL 2^32 float bit patterns. NaN results compare equal as a class (payloads are not used by
// the path); everything else must be bit-identical.
//   g++ -O2 -ffp-contract=off -DPR_HOST_EMUL -std=c++17 -pthread -o /tmp/sincos_exhaustive tools/sincos_exhaustive.cpp && /tmp/sincos_exhaustive
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../playrho_b200/csrc/pr_math.cuh"
namespace pr { thread_local unsigned emul_tid = 0, emul_nthreads = 1; }

int main(int argc, char** argv)
{
    const unsigned long long stride = argc > 1 ? std::max(1ull, std::strtoull(argv[1], nullptr, 10)) : 1ull; // 1 = every float
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<unsigned long long> bad{0}, checked{0};
    std::atomic<uint32_t> firstBad{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) {
        th.emplace_back([&, t]() {
            unsigned long long myBad = 0, my = 0;
            for (unsigned long long u = t * stride; u < (1ull << 32); u += T * stride) {
                const uint32_t bits = static_cast<uint32_t>(u);
                float x; std::memcpy(&x, &bits, 4);
                float s0, c0, s1, c1;
                pr::glibc_sincosf(x, &s0, &c0);
                ::sincosf(x, &s1, &c1);
                ++my;
                const bool nanOk = (std::isnan(s0) && std::isnan(s1) && std::isnan(c0) && std::isnan(c1));
                if (!nanOk && (std::memcmp(&s0, &s1, 4) != 0 || std::memcmp(&c0, &c1, 4) != 0)) { if (!myBad) firstBad = bits; ++myBad; }
            }
            bad += myBad; checked += my;
        });
    }
    for (auto& x : th) x.join();
    std::printf("glibc_sincosf restatement (pr_math.cuh, host build) vs libm sincosf: %llu inputs checked, %llu mismatches", checked.load(), bad.load());
    if (bad.load()) std::printf(" (e.g. bits 0x%08x)", firstBad.load());
    std::printf("\n");
    return bad.load() ? 1 : 0;
}
