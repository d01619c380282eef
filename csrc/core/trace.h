// Tracing and crash diagnostics.
//
// The reference has no structured tracing (wall-clock log lines only) and, for
// crashes, a signal handler that prints a boost::stacktrace (src/utils.cpp:115-122).  Here:
//   * NVTX ranges around the data-plane calls (visible in Nsight Systems / ncu), enabled with
//     ISTORE_NVTX=1 so that the hot path pays nothing otherwise;
//   * install_crash_handler(): backtrace of the faulting thread on SIGSEGV/SIGBUS/SIGFPE/
//     SIGABRT, then the default action.  Installed by the server entry point only - a
//     library must not take over the host application's signal handling.
#pragma once

namespace istore {

bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();

struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) : on(nvtx_enabled()) {
        if (on) nvtx_push(name);
    }
    ~NvtxRange() {
        if (on) nvtx_pop();
    }
};

void install_crash_handler();

}  // namespace istore
