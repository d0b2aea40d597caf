// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for fmt 11.2.0 (cmake/fmt.cmake). The
// reference uses fmt only to print per-phase wall-clock lines ("mc  calc", "mc  sort",
// "norm est", "oct  upd", "total", "sub fin") and a MortonCode
// formatter. This shim keeps the call sites compiling unmodified, prints nothing, and
// records the last value printed under each tag so the harness can report the phases.
#pragma once
#include <cstring>
#include <string>
#include <string_view>
#include <type_traits>

namespace chad_ref_shim {
struct PhaseLog {
    static constexpr int kMax = 8;
    char tag[kMax][16];
    double last[kMax];
    double sum[kMax];
    int n = 0;
    void record(std::string_view fmt_str, double v) {
        size_t e = fmt_str.find('{');
        if (e == std::string_view::npos) return;
        while (e > 0 && fmt_str[e - 1] == ' ') e--;
        if (e == 0 || e >= 16) return;
        for (int i = 0; i < n; i++)
            if (std::strlen(tag[i]) == e && std::memcmp(tag[i], fmt_str.data(), e) == 0) {
                last[i] = v;
                sum[i] += v;
                return;
            }
        if (n == kMax) return;
        std::memcpy(tag[n], fmt_str.data(), e);
        tag[n][e] = 0;
        last[n] = v;
        sum[n] = v;
        n++;
    }
};
inline PhaseLog& phase_log() {
    static PhaseLog log;
    return log;
}
}  // namespace chad_ref_shim

namespace fmt {
template <typename... Args>
inline void println(std::string_view fmt_str, const Args&... args) {
    if constexpr (sizeof...(Args) == 1) {
        const auto& first = (args, ...);
        if constexpr (std::is_same_v<std::decay_t<decltype(first)>, double>) chad_ref_shim::phase_log().record(fmt_str, first);
    }
}
struct format_context {
    using iterator = char*;
};
template <typename T, typename Enable = void>
struct formatter;
template <>
struct formatter<std::string> {
    auto format(const std::string&, format_context&) const -> format_context::iterator { return nullptr; }
};
}  // namespace fmt
