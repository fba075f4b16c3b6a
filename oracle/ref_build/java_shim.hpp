// java_shim.hpp — just enough Java for g++ to compile, as C++, the pieces of the reference's OWN
// Java sources that sit on the hot path (build_ref.sh extracts them by line range from
// /root/reference/.../{SWASA,ScielabProcessor,ImageManipulation}.java and applies purely
// syntactic rewrites: `float[] x` -> JArr<float> x, `new float[n]` -> JArr<float>(n),
// `this.` -> `this->`, access modifiers dropped).  TEST INFRASTRUCTURE.
//
// What the shim supplies is what those sources take from outside the repository:
//   Math.pow/exp/tanh/sqrt   -> glibc in double (the JDK's are fdlibm-derived; both < 1 ulp)
//   icy.util.Random          -> un-vendored icy.jar; pinned to the java.util.Random LCG of the JDK
//                               specification with an explicit seed
//   System.arraycopy, Double.MAX_VALUE, arrays with reference semantics and a .length field
// Java float arithmetic is IEEE binary32 with a rounding after every operation: compile with
// -ffp-contract=off (x86-64 SSE, FLT_EVAL_METHOD 0).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <memory>
#include <vector>

typedef bool boolean;

template <class T>
struct JArr {  // a Java array: a reference to shared storage
    std::shared_ptr<std::vector<T>> p;
    int length;
    JArr() : length(0) {}
    explicit JArr(int n) : p(std::make_shared<std::vector<T>>((size_t)n)), length(n) {}
    JArr(std::initializer_list<T> l) : p(std::make_shared<std::vector<T>>(l)), length((int)l.size()) {}
    T& operator[](int i) const { return (*p)[(size_t)i]; }
    T* begin() const { return p ? p->data() : nullptr; }
    T* end() const { return p ? p->data() + length : nullptr; }
};

struct JavaMath {
    double PI = 3.14159265358979323846;  // java.lang.Math.PI
    double pow(double a, double b) const { return ::pow(a, b); }
    double exp(double a) const { return ::exp(a); }
    double tanh(double a) const { return ::tanh(a); }
    double sqrt(double a) const { return ::sqrt(a); }
    double log(double a) const { return ::log(a); }
    double atan(double a) const { return ::atan(a); }
    double ceil(double a) const { return ::ceil(a); }
    long long round(double a) const { return (long long)::floor(a + 0.5); }  // Math.round(double): floor(x + 1/2)
    float signum(float a) const { return a > 0.0f ? 1.0f : a < 0.0f ? -1.0f : a; }
    float abs(float a) const { return ::fabsf(a); }
    double abs(double a) const { return ::fabs(a); }
    int abs(int a) const { return a < 0 ? -a : a; }
};
static const JavaMath Math;

struct JavaDouble { double MAX_VALUE = DBL_MAX; };
static const JavaDouble Double;

// icy.type.collection.array.Array1DUtil.fill (un-vendored icy.jar): sets every element
struct IcyArray1DUtil {
    template <class T>
    void fill(const JArr<T>& a, double v) const { for (int i = 0; i < a.length; ++i) a[i] = (T)v; }
};
static const IcyArray1DUtil Array1DUtil;

// the java.util.List<Integer> that ScielabProcessor.java:149-153 builds with a stream
struct JIntList {
    std::vector<int> v;
    int size() const { return (int)v.size(); }
    int get(int i) const { return v[(size_t)i]; }
};

struct JavaSystem {
    template <class T>
    void arraycopy(const JArr<T>& src, int srcPos, const JArr<T>& dst, int dstPos, int len) const {
        for (int i = 0; i < len; ++i) dst[dstPos + i] = src[srcPos + i];
    }
};
static const JavaSystem System;

// java.util.Random (JDK specification): 48-bit LCG
struct JavaUtilRandom {
    uint64_t state = 0;
    void setSeed(int64_t seed) { state = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
    int32_t next(int bits) {
        state = (state * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)((int64_t)state >> (48 - bits));
    }
    float nextFloat() { return (float)next(24) / (float)(1 << 24); }
    double nextDouble() { return (double)(((int64_t)next(26) << 27) + next(27)) * 0x1.0p-53; }
};
extern thread_local JavaUtilRandom g_java_random;
extern thread_local long g_java_random_draws;

// `icy.util.Random.nextFloat()` parses in C++ as member accesses on these objects
struct IcyRandom {
    float nextFloat() const { ++g_java_random_draws; return g_java_random.nextFloat(); }
    double nextDouble() const { ++g_java_random_draws; return g_java_random.nextDouble(); }
};
struct IcyUtil { IcyRandom Random; };
struct IcyRoot { IcyUtil util; };
static const IcyRoot icy;

// the plugin object SWASA keeps a reference to (only isStopFlag() is reached from the loop,
// HybridQuantization.java:311-319)
extern thread_local const volatile int* g_ref_stop;
struct HybridQuantization {
    boolean isStopFlag() const { return g_ref_stop && *g_ref_stop; }
};
