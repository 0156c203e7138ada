// Runtime of gtest-lite (see include/gtest/gtest.h).  Supports --gtest_filter=<substring>[-<substring>].
#include <cstdarg>
#include <cstdio>
#include <gtest/gtest.h>

namespace testing {

std::vector<TestInfo>& registry() { static std::vector<TestInfo> r; return r; }
int& currentFailures() { static int f = 0; return f; }
static std::string g_filter, g_neg_filter;

void InitGoogleTest(int* argc, char** argv)
{
    int w = 1;
    for (int i = 1; i < *argc; ++i) {
        if (std::strncmp(argv[i], "--gtest_filter=", 15) == 0) {
            std::string f = argv[i] + 15;
            const size_t dash = f.find('-');
            g_filter = f.substr(0, dash);
            if (dash != std::string::npos) g_neg_filter = f.substr(dash + 1);
            if (g_filter == "*") g_filter.clear();
        } else {
            argv[w++] = argv[i];
        }
    }
    *argc = w;
}

int RunAllTests()
{
    int failed = 0, ran = 0;
    for (auto& t : registry()) {
        const std::string full = std::string(t.suite) + "." + t.name;
        if (!g_filter.empty() && full.find(g_filter) == std::string::npos) continue;
        if (!g_neg_filter.empty() && full.find(g_neg_filter) != std::string::npos) continue;
        std::printf("[ RUN      ] %s\n", full.c_str());
        std::fflush(stdout);
        currentFailures() = 0;
        t.fn();
        ++ran;
        if (currentFailures()) { ++failed; std::printf("[  FAILED  ] %s (%d failures)\n", full.c_str(), currentFailures()); }
        else std::printf("[       OK ] %s\n", full.c_str());
    }
    std::printf("[==========] %d tests ran, %d failed.\n", ran, failed);
    return failed ? 1 : 0;
}

namespace internal {
enum GTestColor { COLOR_DEFAULT, COLOR_RED, COLOR_GREEN, COLOR_YELLOW };
void ColoredPrintf(GTestColor, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    std::vprintf(fmt, ap);
    va_end(ap);
}
}  // namespace internal
}  // namespace testing
