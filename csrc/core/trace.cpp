#include "trace.h"

#include <execinfo.h>
#include <nvtx3/nvToolsExt.h>
#include <signal.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <initializer_list>

namespace istore {

bool nvtx_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("ISTORE_NVTX");
        return e && e[0] && e[0] != '0';
    }();
    return on;
}

void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

namespace {
void on_fatal(int sig) {
    // async-signal-safe: write(2) and backtrace_symbols_fd only
    const char head[] = "\n[infini] fatal signal, backtrace:\n";
    (void)!write(STDERR_FILENO, head, sizeof(head) - 1);
    void* frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, STDERR_FILENO);
    signal(sig, SIG_DFL);
    raise(sig);
}
}  // namespace

void install_crash_handler() {
    void* warm[1];
    backtrace(warm, 1);  // loads libgcc now, not inside the handler
    struct sigaction sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.sa_handler = on_fatal;
    sigemptyset(&sa.sa_mask);
    sa.sa_flags = SA_RESETHAND;
    for (int sig : {SIGSEGV, SIGBUS, SIGFPE, SIGABRT}) sigaction(sig, &sa, nullptr);
}

}  // namespace istore
