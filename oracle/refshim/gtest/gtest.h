// gtest/gtest.h — STAND-IN for googletest, just enough to compile and run the reference's own
// tests/tstMesh.cpp, tests/tstProblemManager.cpp and tests/tstBoundaryConditions.cpp unmodified
// against the Kokkos / Cajita stand-ins of this directory.  TEST INFRASTRUCTURE ONLY.
// Supports: ::testing::Test (SetUp / TearDown), ::testing::Types, TYPED_TEST_SUITE, TYPED_TEST,
// EXPECT_EQ / ASSERT_EQ / EXPECT_TRUE / ASSERT_TRUE, InitGoogleTest, RUN_ALL_TESTS.
#ifndef CFREF_SHIM_GTEST_H
#define CFREF_SHIM_GTEST_H

#include <cstdio>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace testing
{

class Test
{
  public:
    virtual ~Test() = default;
    void Run()
    {
        SetUp();
        TestBody();
        TearDown();
    }

  protected:
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};

template <class... Ts>
struct Types
{
};

namespace internal
{
struct Case
{
    std::string name;
    std::function<void()> run;
};
inline std::vector<Case>& cases()
{
    static std::vector<Case> c;
    return c;
}
inline int& failures_in_current_test()
{
    static int f = 0;
    return f;
}

template <template <class> class TestT, class T>
void register_one( const char* suite, const char* name, int index )
{
    cases().push_back( Case{ std::string( suite ) + "/" + std::to_string( index ) + "." + name, []() {
                                TestT<T> t;
                                t.Run();
                            } } );
}
template <template <class> class TestT, class TypeList>
struct RegisterTyped;
template <template <class> class TestT, class... Ts>
struct RegisterTyped<TestT, Types<Ts...>>
{
    static bool run( const char* suite, const char* name )
    {
        int index = 0;
        ( register_one<TestT, Ts>( suite, name, index++ ), ... );
        return true;
    }
};

template <class A, class B>
bool check_eq( const A& a, const B& b, const char* ea, const char* eb, const char* file, int line )
{
    if ( a == b )
        return true;
    std::cerr << file << ":" << line << ": Failure\nExpected equality of these values:\n  " << ea << "\n    Which is: "
              << a << "\n  " << eb << "\n    Which is: " << b << "\n";
    ++failures_in_current_test();
    return false;
}
inline bool check_true( bool v, const char* e, const char* file, int line )
{
    if ( v )
        return true;
    std::cerr << file << ":" << line << ": Failure\nValue of: " << e << "\n  Actual: false\nExpected: true\n";
    ++failures_in_current_test();
    return false;
}
} // namespace internal

inline void InitGoogleTest( int*, char** ) {}

} // namespace testing

inline int RUN_ALL_TESTS()
{
    int failed = 0;
    const auto& cs = ::testing::internal::cases();
    std::printf( "[==========] Running %zu tests.\n", cs.size() );
    for ( const auto& c : cs )
    {
        std::printf( "[ RUN      ] %s\n", c.name.c_str() );
        ::testing::internal::failures_in_current_test() = 0;
        c.run();
        if ( ::testing::internal::failures_in_current_test() )
        {
            ++failed;
            std::printf( "[  FAILED  ] %s\n", c.name.c_str() );
        }
        else
            std::printf( "[       OK ] %s\n", c.name.c_str() );
    }
    std::printf( "[==========] %zu tests ran.\n[  PASSED  ] %zu tests.\n", cs.size(), cs.size() - failed );
    if ( failed )
        std::printf( "[  FAILED  ] %d tests.\n", failed );
    return failed ? 1 : 0;
}

#define TYPED_TEST_SUITE( Fixture, TypeList ) typedef TypeList gtest_type_params_##Fixture##_

#define TYPED_TEST( Fixture, Name )                                                                \
    template <class gtest_TypeParam_>                                                              \
    class Fixture##_##Name##_Test : public Fixture<gtest_TypeParam_>                               \
    {                                                                                              \
      protected:                                                                                   \
        typedef Fixture<gtest_TypeParam_> TestFixture;                                             \
        typedef gtest_TypeParam_ TypeParam;                                                        \
        void TestBody() override;                                                                  \
    };                                                                                             \
    static bool gtest_registered_##Fixture##_##Name##_ =                                           \
        ::testing::internal::RegisterTyped<Fixture##_##Name##_Test,                                \
                                           gtest_type_params_##Fixture##_>::run( #Fixture, #Name ); \
    template <class gtest_TypeParam_>                                                              \
    void Fixture##_##Name##_Test<gtest_TypeParam_>::TestBody()

#define EXPECT_EQ( a, b ) ::testing::internal::check_eq( ( a ), ( b ), #a, #b, __FILE__, __LINE__ )
#define ASSERT_EQ( a, b )                                                                          \
    if ( !::testing::internal::check_eq( ( a ), ( b ), #a, #b, __FILE__, __LINE__ ) )              \
    return
#define EXPECT_TRUE( a ) ::testing::internal::check_true( static_cast<bool>( a ), #a, __FILE__, __LINE__ )
#define ASSERT_TRUE( a )                                                                           \
    if ( !::testing::internal::check_true( static_cast<bool>( a ), #a, __FILE__, __LINE__ ) )      \
    return

#endif
