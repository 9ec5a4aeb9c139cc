// Stand-in for googletest, TEST INFRASTRUCTURE ONLY.
// Lets the reference's *_test.cc files compile unmodified so that their
// file-local circuit builders / witness fillers can be reused by
// oracle/ref_build/*.cc.  Tests are compiled but never registered or run.
#ifndef ORACLE_STUB_GTEST_H_
#define ORACLE_STUB_GTEST_H_
#include <iostream>
namespace testing {
struct Sink {
  template <class T> Sink& operator<<(const T&) { return *this; }
};
class Test {
 public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
};
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing
#define ORACLE_CAT_(a, b) a##b
#define ORACLE_CAT(a, b) ORACLE_CAT_(a, b)
#define TEST(suite, name) \
  [[maybe_unused]] static void ORACLE_CAT(suite##_##name##_, __LINE__)()
#define ORACLE_CHECK(c) if (c) {} else ::testing::Sink()
#define EXPECT_TRUE(c) ORACLE_CHECK((c))
#define EXPECT_FALSE(c) ORACLE_CHECK(!(c))
#define EXPECT_EQ(a, b) ORACLE_CHECK((a) == (b))
#define EXPECT_NE(a, b) ORACLE_CHECK(!((a) == (b)))
#define EXPECT_LT(a, b) ORACLE_CHECK((a) < (b))
#define EXPECT_LE(a, b) ORACLE_CHECK((a) <= (b))
#define EXPECT_GT(a, b) ORACLE_CHECK((a) > (b))
#define EXPECT_GE(a, b) ORACLE_CHECK((a) >= (b))
#define ASSERT_TRUE EXPECT_TRUE
#define ASSERT_FALSE EXPECT_FALSE
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_NE EXPECT_NE
#define EXPECT_DEATH(stmt, re) ORACLE_CHECK(true)
inline int RUN_ALL_TESTS() { return 0; }
#endif
