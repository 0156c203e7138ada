// gtest-lite: the subset of the GoogleTest API that the reference's stereoDNN/tests/tests_main.cpp uses, so that the
// UNCHANGED reference test suite can be compiled against libnvstereo_inference.so in a container without GoogleTest.
// Test infrastructure only (tools/dropin); not part of the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

struct TestInfo { const char* suite; const char* name; std::function<void()> fn; };
std::vector<TestInfo>& registry();
int& currentFailures();
void InitGoogleTest(int* argc, char** argv);
int RunAllTests();

struct Registrar {
    Registrar(const char* s, const char* n, std::function<void()> f) { registry().push_back({s, n, std::move(f)}); }
};

// Collects the user's streamed message; reports on destruction when the check failed.
class Result {
public:
    Result(bool ok, const char* file, int line, std::string what) : ok_(ok)
    {
        if (!ok_) head_ << file << ":" << line << ": Failure\n" << what << "\n";
    }
    Result(const Result& o) : ok_(o.ok_) { head_ << o.head_.str(); }
    ~Result()
    {
        if (!ok_) {
            if (currentFailures()++ < 20) std::cerr << head_.str() << msg_.str() << std::endl;
        }
    }
    template <typename T> Result& operator<<(const T& v) { if (!ok_) msg_ << v; return *this; }
    bool ok() const { return ok_; }
private:
    bool ok_;
    std::ostringstream head_, msg_;
};

template <typename A, typename B> std::string describe(const char* op, const char* ea, const char* eb, const A& a, const B& b)
{
    std::ostringstream s;
    s << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: " << a << " vs " << b;
    return s.str();
}
inline std::string describe(const char* op, const char* ea, const char* eb, std::nullptr_t, std::nullptr_t)
{
    return std::string("Expected: (") + ea + ") " + op + " (" + eb + ")";
}
template <typename A> std::string describe(const char* op, const char* ea, const char* eb, const A&, std::nullptr_t)
{
    return std::string("Expected: (") + ea + ") " + op + " (" + eb + ")";
}

// 4-ULP comparison, as EXPECT_FLOAT_EQ.
inline bool floatAlmostEqual(float a, float b)
{
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u + 1 : u | 0x80000000u; };
    const uint32_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}

}  // namespace testing

#define GTL_CAT_(a, b) a##b
#define GTL_CAT(a, b) GTL_CAT_(a, b)
#define TEST(suite, name)                                                                         \
    static void GTL_CAT(suite, GTL_CAT(_, GTL_CAT(name, _body)))();                               \
    static ::testing::Registrar GTL_CAT(suite, GTL_CAT(_, GTL_CAT(name, _reg)))(#suite, #name,    \
                                                           GTL_CAT(suite, GTL_CAT(_, GTL_CAT(name, _body)))); \
    static void GTL_CAT(suite, GTL_CAT(_, GTL_CAT(name, _body)))()

#define GTL_CMP(a, b, op, opname) ::testing::Result((a)op(b), __FILE__, __LINE__, ::testing::describe(opname, #a, #b, (a), (b)))
#define EXPECT_EQ(a, b) GTL_CMP(a, b, ==, "==")
#define EXPECT_NE(a, b) GTL_CMP(a, b, !=, "!=")
#define EXPECT_GT(a, b) GTL_CMP(a, b, >, ">")
#define EXPECT_GE(a, b) GTL_CMP(a, b, >=, ">=")
#define EXPECT_LT(a, b) GTL_CMP(a, b, <, "<")
#define EXPECT_LE(a, b) GTL_CMP(a, b, <=, "<=")
#define EXPECT_TRUE(c) ::testing::Result(static_cast<bool>(c), __FILE__, __LINE__, "Expected true: " #c)
#define EXPECT_FALSE(c) ::testing::Result(!static_cast<bool>(c), __FILE__, __LINE__, "Expected false: " #c)
#define EXPECT_FLOAT_EQ(a, b) ::testing::Result(::testing::floatAlmostEqual((a), (b)), __FILE__, __LINE__, ::testing::describe("~=", #a, #b, (a), (b)))
#define EXPECT_NEAR(a, b, tol) ::testing::Result(std::fabs((double)(a) - (double)(b)) <= (tol), __FILE__, __LINE__, ::testing::describe("near", #a, #b, (a), (b)))
#define ADD_FAILURE() ::testing::Result(false, __FILE__, __LINE__, "Failed")
// ASSERT_*: abort the current test function on failure.  The message stream of a failed assertion is discarded.
#define GTL_ASSERT(res) if (::testing::Result gtl_r_ = (res); gtl_r_.ok()) ; else return (void)(gtl_r_)
#define ASSERT_EQ(a, b) GTL_ASSERT(EXPECT_EQ(a, b))
#define ASSERT_NE(a, b) GTL_ASSERT(EXPECT_NE(a, b))
#define ASSERT_TRUE(c) GTL_ASSERT(EXPECT_TRUE(c))
#define RUN_ALL_TESTS() ::testing::RunAllTests()
