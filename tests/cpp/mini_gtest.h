// mini_gtest.h -- a very small GoogleTest look-alike (GoogleTest is not installed in this image) so that the
// host-layer tests can keep the shape of the reference's gtest programs (tests/*.cpp of the reference).
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace testing {
class Test {
 public:
  virtual ~Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};
struct Registry {
  struct Entry { std::string name; std::function<Test *()> make; };
  static std::vector<Entry> &tests() { static std::vector<Entry> t; return t; }
  static int &failures() { static int f = 0; return f; }
  static bool &current_failed() { static bool f = false; return f; }
};
struct Registrar {
  Registrar(const char *suite, const char *name, std::function<Test *()> make) {
    Registry::tests().push_back({std::string(suite) + "." + name, make});
  }
};
struct FatalFailure {};
inline void fail(const char *file, int line, const std::string &msg) {
  std::printf("%s:%d: Failure\n  %s\n", file, line, msg.c_str());
  Registry::current_failed() = true;
  throw FatalFailure();
}
inline int RunAllTests(const char *filter = nullptr) {
  int ran = 0;
  for (auto &e : Registry::tests()) {
    if (filter && e.name.find(filter) == std::string::npos) continue;
    std::printf("[ RUN      ] %s\n", e.name.c_str());
    std::fflush(stdout);
    Registry::current_failed() = false;
    Test *t = e.make();
    try {
      t->SetUp();
      t->TestBody();
    } catch (FatalFailure &) {
    } catch (std::exception &ex) {
      std::printf("  unexpected exception: %s\n", ex.what());
      Registry::current_failed() = true;
    }
    try { t->TearDown(); } catch (...) {}
    delete t;
    ran++;
    if (Registry::current_failed()) { Registry::failures()++; std::printf("[  FAILED  ] %s\n", e.name.c_str()); }
    else std::printf("[       OK ] %s\n", e.name.c_str());
    std::fflush(stdout);
  }
  std::printf("[==========] %d tests ran, %d failed.\n", ran, Registry::failures());
  return Registry::failures() == 0 ? 0 : 1;
}
}  // namespace testing

#define MG_CAT_(a, b) a##b
#define MG_CAT(a, b) MG_CAT_(a, b)
#define MG_TEST_(suite, name, base)                                                              \
  class MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test))) : public base {                            \
   public:                                                                                       \
    void TestBody() override;                                                                    \
  };                                                                                             \
  static ::testing::Registrar MG_CAT(reg_, MG_CAT(suite, MG_CAT(_, name)))(                      \
      #suite, #name, []() -> ::testing::Test * { return new MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test)))(); }); \
  void MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test)))::TestBody()
#define TEST(suite, name) MG_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) MG_TEST_(fixture, name, fixture)

#define MG_STR(x) std::to_string(x)
#define ASSERT_TRUE(c) do { if (!(c)) ::testing::fail(__FILE__, __LINE__, "expected true: " #c); } while (0)
#define ASSERT_FALSE(c) do { if ((c)) ::testing::fail(__FILE__, __LINE__, "expected false: " #c); } while (0)
#define ASSERT_EQ(a, b) do { auto va_ = (a); auto vb_ = (b); if (!(va_ == vb_)) ::testing::fail(__FILE__, __LINE__, std::string(#a " == " #b " failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
#define ASSERT_NE(a, b) do { auto va_ = (a); auto vb_ = (b); if (va_ == vb_) ::testing::fail(__FILE__, __LINE__, #a " != " #b " failed"); } while (0)
#define ASSERT_LE(a, b) do { auto va_ = (a); auto vb_ = (b); if (!(va_ <= vb_)) ::testing::fail(__FILE__, __LINE__, std::string(#a " <= " #b " failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
#define ASSERT_GE(a, b) do { auto va_ = (a); auto vb_ = (b); if (!(va_ >= vb_)) ::testing::fail(__FILE__, __LINE__, std::string(#a " >= " #b " failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
#define ASSERT_LT(a, b) do { auto va_ = (a); auto vb_ = (b); if (!(va_ < vb_)) ::testing::fail(__FILE__, __LINE__, std::string(#a " < " #b " failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
#define ASSERT_NEAR(a, b, tol) do { double va_ = (a), vb_ = (b); if (!(std::fabs(va_ - vb_) <= (tol))) ::testing::fail(__FILE__, __LINE__, std::string(#a " near " #b " failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
// 4-ULP equality like gtest's ASSERT_DOUBLE_EQ
#define ASSERT_DOUBLE_EQ(a, b) do { double va_ = (a), vb_ = (b); double sc_ = std::fmax(std::fabs(va_), std::fabs(vb_)); if (!(va_ == vb_ || std::fabs(va_ - vb_) <= 4.0 * 2.220446049250313e-16 * sc_)) ::testing::fail(__FILE__, __LINE__, std::string(#a " == " #b " (double) failed: ") + MG_STR(va_) + " vs " + MG_STR(vb_)); } while (0)
#define ASSERT_THROW(stmt, ex) do { bool caught_ = false; try { stmt; } catch (ex &) { caught_ = true; } catch (...) {} if (!caught_) ::testing::fail(__FILE__, __LINE__, "expected exception " #ex " from: " #stmt); } while (0)
#define EXPECT_TRUE ASSERT_TRUE
#define EXPECT_EQ ASSERT_EQ
#define EXPECT_LE ASSERT_LE
