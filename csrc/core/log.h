// Levelled console logger (role of the reference's spdlog wrapper, src/log.h:9-32).
// Levels and their string names are the reference's: debug, info, warning, error.
#pragma once

#include <cstdarg>
#include <string>

namespace istore {

enum class LogLevel : int { kDebug = 0, kInfo = 1, kWarning = 2, kError = 3, kOff = 4 };

// Returns false when `level` is not one of debug|info|warning|error|off.
bool set_log_level(const std::string& level);
LogLevel log_level();
void log_msg(const std::string& level, const std::string& msg);
void log_write(LogLevel lvl, const char* file, int line, const char* fmt, ...)
    __attribute__((format(printf, 4, 5)));

}  // namespace istore

#define IS_LOG(lvl, ...)                                                  \
    do {                                                                  \
        if (static_cast<int>(lvl) >= static_cast<int>(::istore::log_level())) \
            ::istore::log_write(lvl, __FILE__, __LINE__, __VA_ARGS__);    \
    } while (0)
#define LOG_DEBUG(...) IS_LOG(::istore::LogLevel::kDebug, __VA_ARGS__)
#define LOG_INFO(...) IS_LOG(::istore::LogLevel::kInfo, __VA_ARGS__)
#define LOG_WARN(...) IS_LOG(::istore::LogLevel::kWarning, __VA_ARGS__)
#define LOG_ERROR(...) IS_LOG(::istore::LogLevel::kError, __VA_ARGS__)
