#include "log.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <mutex>

namespace istore {
namespace {
std::atomic<int> g_level{static_cast<int>(LogLevel::kWarning)};
std::mutex g_mu;
const char* kNames[] = {"debug", "info", "warning", "error"};
const char* kColors[] = {"\033[36m", "\033[32m", "\033[33m", "\033[31m"};

int parse_level(const std::string& s) {
    if (s == "debug") return 0;
    if (s == "info") return 1;
    if (s == "warning" || s == "warn") return 2;
    if (s == "error") return 3;
    if (s == "off") return 4;
    return -1;
}
}  // namespace

bool set_log_level(const std::string& level) {
    const int l = parse_level(level);
    if (l < 0) return false;
    g_level.store(l, std::memory_order_relaxed);
    return true;
}

LogLevel log_level() { return static_cast<LogLevel>(g_level.load(std::memory_order_relaxed)); }

void log_write(LogLevel lvl, const char* file, int line, const char* fmt, ...) {
    char body[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(body, sizeof(body), fmt, ap);
    va_end(ap);

    using namespace std::chrono;
    const auto now = system_clock::now();
    const std::time_t t = system_clock::to_time_t(now);
    const int ms = int(duration_cast<milliseconds>(now.time_since_epoch()).count() % 1000);
    std::tm tm{};
    localtime_r(&t, &tm);
    const int li = static_cast<int>(lvl);
    const bool tty = false;  // colour only helps interactive use; keep logs grep-able
    const char* base = std::strrchr(file, '/');
    base = base ? base + 1 : file;

    std::lock_guard<std::mutex> lk(g_mu);
    if (li >= static_cast<int>(LogLevel::kWarning) && file[0])
        std::fprintf(stderr, "[%02d:%02d:%02d.%03d] [infini] [%s%s%s] [%s:%d] %s\n", tm.tm_hour,
                     tm.tm_min, tm.tm_sec, ms, tty ? kColors[li] : "", kNames[li],
                     tty ? "\033[0m" : "", base, line, body);
    else
        std::fprintf(stderr, "[%02d:%02d:%02d.%03d] [infini] [%s] %s\n", tm.tm_hour, tm.tm_min,
                     tm.tm_sec, ms, kNames[li], body);
}

void log_msg(const std::string& level, const std::string& msg) {
    int l = parse_level(level);
    if (l < 0 || l > 3) l = 1;
    if (l >= g_level.load(std::memory_order_relaxed))
        log_write(static_cast<LogLevel>(l), "", 0, "%s", msg.c_str());
}

}  // namespace istore
