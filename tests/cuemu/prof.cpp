// TEST INFRASTRUCTURE ONLY -- instruction-pointer sampler for the emulated library (SIGPROF, 1 kHz): with CUEMU_REPLAY the
// kernels are skipped and the recorded device->host copies are served instead, so the deterministic host glue of the session
// layer can be re-run at full speed and profiled on a machine without a GPU (tools/host_profile.py).
#include <signal.h>
#include <sys/time.h>
#include <ucontext.h>
#include <dlfcn.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>

static unsigned long *g_ips = nullptr;
static std::atomic<long> g_n{0};
static const long kCap = 4000000;

static void on_sigprof(int, siginfo_t *, void *uc) {
    const long i = g_n.fetch_add(1);
    if (i < kCap) g_ips[i] = ((ucontext_t *)uc)->uc_mcontext.gregs[REG_RIP];
}

extern "C" void cuemu_prof_start(int hz) {
    if (!g_ips) g_ips = (unsigned long *)std::malloc(kCap * sizeof(unsigned long));
    g_n = 0;
    struct sigaction sa {};
    sa.sa_sigaction = on_sigprof;
    sa.sa_flags = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, nullptr);
    struct itimerval it;
    it.it_interval.tv_sec = 0; it.it_interval.tv_usec = 1000000 / (hz > 0 ? hz : 1000);
    it.it_value = it.it_interval;
    setitimer(ITIMER_PROF, &it, nullptr);
}

// writes the load address of this library followed by one sampled instruction pointer per line
extern "C" void cuemu_prof_dump(const char *path) {
    struct itimerval it {};
    setitimer(ITIMER_PROF, &it, nullptr);
    Dl_info info;
    dladdr((void *)&cuemu_prof_dump, &info);
    FILE *f = std::fopen(path, "w");
    if (!f) return;
    long n = g_n.load();
    if (n > kCap) n = kCap;
    std::fprintf(f, "%p\n", info.dli_fbase);
    for (long i = 0; i < n; ++i) std::fprintf(f, "%lx\n", g_ips[i]);
    std::fclose(f);
}
