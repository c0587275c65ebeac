/*
 * glsl_shim.hpp -- TEST INFRASTRUCTURE (oracle/_ref build only, never linked into the product).
 *
 * Lets the REFERENCE'S OWN compute shaders (/root/reference/shaders/*.comp.glsl, read where they lie by
 * oracle/ref/translate.py) be compiled by g++ and executed on the host, so that the CPU oracle can be pinned against
 * the reference's real code instead of a restatement.  Nothing in this file restates reference logic: it is a small
 * GLSL run-time -- vector/matrix types with swizzles, the built-in functions the path uses, buffer references as host
 * pointers, atomics, and a SIMT emulator that runs every invocation of a workgroup as a cooperative fiber so that
 * barrier() and the subgroup operations (ballot, shuffle, scans, partition, ...) behave as on a 32-wide GPU.
 *
 * SIMT model.  Invocations run one at a time until they reach a collective (a subgroup operation or barrier()).  When
 * every live invocation of the workgroup is parked, ONE collective is resolved: of the lowest-numbered subgroup that has
 * lanes parked at a subgroup operation, the one with the smallest call-site id (= earliest in the shader text); only
 * lanes parked at the same call site take part -- the structured-reconvergence behaviour the shaders are written for.
 * barrier() is resolved when nothing else is left.  Global atomics are plain read-modify-writes (one host thread), so a
 * dispatch executes ONE valid serialisation of the reference's nondeterministic append order: workgroups ascending,
 * subgroups ascending between barriers, lanes ascending between collectives -- which is the canonical order the
 * oracle and the CUDA path define, so whole buffers can be compared byte for byte.
 *
 * Implementation-defined GLSL behaviour is DEFINED exactly as in the oracle and the kernels:
 * round() = ties-to-even, texture() = software bilinear/repeat on float texels, textureLod() on the far HiZ = max of the
 * 2x2 footprint at the nearest mip with clamp-to-edge.  ceil(log2(x)) is whatever libm gives (the oracle is exact).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

// cooperative context switch: six callee-saved registers + stack pointer (x86-64 SysV); ucontext elsewhere
#if defined(__x86_64__)
#define GLSL_FAST_SWITCH 1
extern "C" void glsl_ctx_switch(void** saveSp, void* loadSp);
#ifdef GLSL_SHIM_IMPLEMENT_SWITCH
asm(".text\n.globl glsl_ctx_switch\n.hidden glsl_ctx_switch\n.type glsl_ctx_switch,@function\nglsl_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size glsl_ctx_switch,.-glsl_ctx_switch\n");
#endif
#else
#define GLSL_FAST_SWITCH 0
#include <ucontext.h>
#endif

namespace glsl {

typedef uint32_t uint;

// --------------------------------------------------------------------------------------------------------------------
// vectors
// --------------------------------------------------------------------------------------------------------------------
template <class R, class S, int N>
struct Swz  // writable swizzle view
{
  S* p[N];
  operator R() const
  {
    R r;
    for(int i = 0; i < N; i++)
      r[i] = *p[i];
    return r;
  }
  Swz& operator=(const R& r)
  {
    for(int i = 0; i < N; i++)
      *p[i] = r[i];
    return *this;
  }
  Swz& operator=(const Swz& o)
  {
    R r = o;
    return *this = r;
  }
#define GLSL_SWZ_OP(op)                                                                                                \
  Swz& operator op##=(const R& r)                                                                                      \
  {                                                                                                                    \
    R a = *this;                                                                                                       \
    return *this = a op r;                                                                                             \
  }                                                                                                                    \
  Swz& operator op##=(S s)                                                                                             \
  {                                                                                                                    \
    R a = *this;                                                                                                       \
    return *this = a op s;                                                                                             \
  }
  GLSL_SWZ_OP(+) GLSL_SWZ_OP(-) GLSL_SWZ_OP(*) GLSL_SWZ_OP(/)
#undef GLSL_SWZ_OP
};

#define GLSL_ARITH(T) template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>>

#define GLSL_VEC2(V, S)                                                                                                \
  struct V                                                                                                             \
  {                                                                                                                    \
    union { S x; S r; };                                                                                               \
    union { S y; S g; };                                                                                               \
    V() = default;                                                                                                     \
    GLSL_ARITH(T) explicit V(T s) : x(S(s)), y(S(s)) {}                                                                \
    GLSL_ARITH2 V(A a, B b) : x(S(a)), y(S(b)) {}                                                                      \
    template <class O, class = decltype(O().x), class = decltype(O().y)> explicit V(const O& o) : x(S(o.x)), y(S(o.y)) {} \
    template <class R2, class S2> explicit V(const Swz<R2, S2, 2>& s) : x(S(*s.p[0])), y(S(*s.p[1])) {}                 \
    S& operator[](int i) { return (&x)[i]; }                                                                           \
    const S& operator[](int i) const { return (&x)[i]; }                                                               \
    Swz<V, S, 2> xy() { return {{&x, &y}}; }                                                                           \
    V xy() const { return *this; }                                                                                     \
    Swz<V, S, 2> yx() { return {{&y, &x}}; }                                                                           \
  };
#define GLSL_ARITH2 template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>

GLSL_VEC2(vec2, float)
GLSL_VEC2(uvec2, uint)
GLSL_VEC2(ivec2, int)
GLSL_VEC2(bvec2, bool)

#define GLSL_VEC3(V, V2, S)                                                                                            \
  struct V                                                                                                             \
  {                                                                                                                    \
    union { S x; S r; };                                                                                               \
    union { S y; S g; };                                                                                               \
    union { S z; S b; };                                                                                               \
    V() = default;                                                                                                     \
    GLSL_ARITH(T) explicit V(T s) : x(S(s)), y(S(s)), z(S(s)) {}                                                       \
    template <class A, class B, class C,                                                                               \
              class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value>> \
    V(A a, B b_, C c) : x(S(a)), y(S(b_)), z(S(c)) {}                                                                  \
    GLSL_ARITH(T) V(const V2& a, T c) : x(a.x), y(a.y), z(S(c)) {}                                                     \
    template <class O, class = decltype(O().x), class = decltype(O().z)> explicit V(const O& o) : x(S(o.x)), y(S(o.y)), z(S(o.z)) {} \
    template <class R2, class S2> explicit V(const Swz<R2, S2, 3>& s) : x(S(*s.p[0])), y(S(*s.p[1])), z(S(*s.p[2])) {}  \
    S& operator[](int i) { return (&x)[i]; }                                                                           \
    const S& operator[](int i) const { return (&x)[i]; }                                                               \
    Swz<V2, S, 2> xy() { return {{&x, &y}}; }                                                                          \
    Swz<V, S, 3> xyz() { return {{&x, &y, &z}}; }                                                                      \
    Swz<V, S, 3> yzx() { return {{&y, &z, &x}}; }                                                                      \
    Swz<V, S, 3> zxy() { return {{&z, &x, &y}}; }                                                                      \
    Swz<V, S, 3> xzy() { return {{&x, &z, &y}}; }                                                                      \
    Swz<V, S, 3> yxz() { return {{&y, &x, &z}}; }                                                                      \
    Swz<V, S, 3> zyx() { return {{&z, &y, &x}}; }                                                                      \
  };

GLSL_VEC3(vec3, vec2, float)
GLSL_VEC3(uvec3, uvec2, uint)
GLSL_VEC3(ivec3, ivec2, int)
GLSL_VEC3(bvec3, bvec2, bool)

#define GLSL_VEC4(V, V3, V2, S)                                                                                        \
  struct V                                                                                                             \
  {                                                                                                                    \
    union { S x; S r; };                                                                                               \
    union { S y; S g; };                                                                                               \
    union { S z; S b; };                                                                                               \
    union { S w; S a; };                                                                                               \
    V() = default;                                                                                                     \
    GLSL_ARITH(T) explicit V(T s) : x(S(s)), y(S(s)), z(S(s)), w(S(s)) {}                                              \
    template <class A, class B, class C, class D,                                                                      \
              class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value && std::is_arithmetic<D>::value>> \
    V(A a_, B b_, C c, D d) : x(S(a_)), y(S(b_)), z(S(c)), w(S(d)) {}                                                  \
    GLSL_ARITH(T) V(const V3& v, T d) : x(v.x), y(v.y), z(v.z), w(S(d)) {}                                             \
    V(const V2& p, const V2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}                                                    \
    S& operator[](int i) { return (&x)[i]; }                                                                           \
    const S& operator[](int i) const { return (&x)[i]; }                                                               \
    Swz<V2, S, 2> xy() { return {{&x, &y}}; }                                                                          \
    Swz<V2, S, 2> zw() { return {{&z, &w}}; }                                                                          \
    Swz<V3, S, 3> xyz() { return {{&x, &y, &z}}; }                                                                     \
  };

GLSL_VEC4(vec4, vec3, vec2, float)
GLSL_VEC4(uvec4, uvec3, uvec2, uint)
GLSL_VEC4(ivec4, ivec3, ivec2, int)
typedef uvec4 bvec4;  // a bool in a uniform / push-constant block occupies 32 bits

struct u8vec4
{
  uint8_t x, y, z, w;
  uvec3   xyz() const { return uvec3(uint(x), uint(y), uint(z)); }
};

// component-wise operators (concrete overloads, so that swizzle views convert implicitly)
#define GLSL_BINOP(V, S, N, op)                                                                                        \
  inline V operator op(const V& a, const V& b) { V r; for(int i = 0; i < N; i++) r[i] = a[i] op b[i]; return r; }      \
  inline V operator op(const V& a, S b) { V r; for(int i = 0; i < N; i++) r[i] = a[i] op b; return r; }                \
  inline V operator op(S a, const V& b) { V r; for(int i = 0; i < N; i++) r[i] = a op b[i]; return r; }                \
  inline V& operator op##=(V& a, const V& b) { a = a op b; return a; }                                                 \
  inline V& operator op##=(V& a, S b) { a = a op b; return a; }
#define GLSL_FOPS(V, N) GLSL_BINOP(V, float, N, +) GLSL_BINOP(V, float, N, -) GLSL_BINOP(V, float, N, *) GLSL_BINOP(V, float, N, /) \
  inline V operator-(const V& a) { V r; for(int i = 0; i < N; i++) r[i] = -a[i]; return r; }
#define GLSL_UOPS(V, S, N) GLSL_BINOP(V, S, N, +) GLSL_BINOP(V, S, N, -) GLSL_BINOP(V, S, N, *) GLSL_BINOP(V, S, N, /) \
  GLSL_BINOP(V, S, N, |) GLSL_BINOP(V, S, N, &) GLSL_BINOP(V, S, N, ^) GLSL_BINOP(V, S, N, >>) GLSL_BINOP(V, S, N, <<)
GLSL_FOPS(vec2, 2) GLSL_FOPS(vec3, 3) GLSL_FOPS(vec4, 4)
GLSL_UOPS(uvec2, uint, 2) GLSL_UOPS(uvec3, uint, 3) GLSL_UOPS(uvec4, uint, 4)
GLSL_UOPS(ivec2, int, 2) GLSL_UOPS(ivec3, int, 3)

#define GLSL_CMP(V, N)                                                                                                 \
  inline bool operator==(const V& a, const V& b) { for(int i = 0; i < N; i++) if(!(a[i] == b[i])) return false; return true; } \
  inline bool operator!=(const V& a, const V& b) { return !(a == b); }
GLSL_CMP(vec2, 2) GLSL_CMP(vec3, 3) GLSL_CMP(vec4, 4) GLSL_CMP(uvec2, 2) GLSL_CMP(uvec3, 3) GLSL_CMP(uvec4, 4)

// --------------------------------------------------------------------------------------------------------------------
// scalar built-ins (names hide the C library's inside namespace glsl)
// --------------------------------------------------------------------------------------------------------------------
template <class A, class B> using common_t = std::common_type_t<A, B>;
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
constexpr common_t<A, B> min(A a, B b) { return common_t<A, B>(b) < common_t<A, B>(a) ? common_t<A, B>(b) : common_t<A, B>(a); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
constexpr common_t<A, B> max(A a, B b) { return common_t<A, B>(a) < common_t<A, B>(b) ? common_t<A, B>(b) : common_t<A, B>(a); }
template <class A, class B, class C, class = std::enable_if_t<std::is_arithmetic<A>::value>>
constexpr A clamp(A v, B lo, C hi) { return min(max(v, A(lo)), A(hi)); }

inline float abs(float a) { return std::fabs(a); }
inline int   abs(int a) { return a < 0 ? -a : a; }
inline float sqrt(float a) { return std::sqrt(a); }
inline float floor(float a) { return std::floor(a); }
inline float ceil(float a) { return std::ceil(a); }
inline float round(float a) { return std::nearbyintf(a); }  // DEFINED: ties to even
inline float log2(float a) { return std::log2(a); }
inline float sin(float a) { return std::sin(a); }
inline float cos(float a) { return std::cos(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float fract(float a) { return a - std::floor(a); }
inline int   bitCount(uint v) { return __builtin_popcount(v); }
inline int   findMSB(uint v) { return v ? 31 - __builtin_clz(v) : -1; }
inline int   findLSB(uint v) { return v ? __builtin_ctz(v) : -1; }

#define GLSL_MAP2(fn, V, N) inline V fn(const V& a, const V& b) { V r; for(int i = 0; i < N; i++) r[i] = fn(a[i], b[i]); return r; }
#define GLSL_MAP1(fn, V, N) inline V fn(const V& a) { V r; for(int i = 0; i < N; i++) r[i] = fn(a[i]); return r; }
GLSL_MAP2(min, ivec2, 2) GLSL_MAP2(max, ivec2, 2) GLSL_MAP2(min, vec2, 2) GLSL_MAP2(min, vec3, 3) GLSL_MAP2(min, vec4, 4) GLSL_MAP2(min, uvec2, 2) GLSL_MAP2(min, uvec3, 3) GLSL_MAP2(min, uvec4, 4)
GLSL_MAP2(max, vec2, 2) GLSL_MAP2(max, vec3, 3) GLSL_MAP2(max, vec4, 4) GLSL_MAP2(max, uvec2, 2) GLSL_MAP2(max, uvec3, 3) GLSL_MAP2(max, uvec4, 4)
GLSL_MAP1(round, vec2, 2) GLSL_MAP1(round, vec3, 3) GLSL_MAP1(round, vec4, 4)
GLSL_MAP1(abs, vec2, 2) GLSL_MAP1(abs, vec3, 3) GLSL_MAP1(abs, vec4, 4)
GLSL_MAP1(floor, vec2, 2) GLSL_MAP1(floor, vec3, 3)
inline uvec3 min(const uvec3& a, uint b) { return min(a, uvec3(b)); }
inline uvec3 max(const uvec3& a, uint b) { return max(a, uvec3(b)); }
inline vec3  min(const vec3& a, float b) { return min(a, vec3(b)); }
inline vec3  max(const vec3& a, float b) { return max(a, vec3(b)); }
#define GLSL_CLAMP(V) inline V clamp(const V& v, const V& lo, const V& hi) { return min(max(v, lo), hi); }
GLSL_CLAMP(vec2) GLSL_CLAMP(vec3) GLSL_CLAMP(vec4) GLSL_CLAMP(uvec3)
inline vec2 clamp(const vec2& v, float lo, float hi) { return clamp(v, vec2(lo), vec2(hi)); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return clamp(v, vec3(lo), vec3(hi)); }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec2& a) { return std::sqrt(dot(a, a)); }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec3  normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3  cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline vec3  mix(const vec3& a, const vec3& b, const bvec3& s) { return vec3(s.x ? b.x : a.x, s.y ? b.y : a.y, s.z ? b.z : a.z); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline bvec2 greaterThan(const vec2& a, const vec2& b) { return bvec2(a.x > b.x, a.y > b.y); }
inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline bool  any(const bvec2& a) { return a.x || a.y; }
inline bool  any(const bvec3& a) { return a.x || a.y || a.z; }
inline bool  all(const bvec2& a) { return a.x && a.y; }
inline bool  all(const bvec3& a) { return a.x && a.y && a.z; }

inline uint64_t packUint2x32(const uvec2& v) { return uint64_t(v.x) | (uint64_t(v.y) << 32); }
inline uvec2    unpackUint2x32(uint64_t v) { return uvec2(uint(v & 0xFFFFFFFFull), uint(v >> 32)); }
inline u8vec4   unpack8(uint v) { return u8vec4{uint8_t(v), uint8_t(v >> 8), uint8_t(v >> 16), uint8_t(v >> 24)}; }
template <class T> inline T nonuniformEXT(T v) { return v; }

// --------------------------------------------------------------------------------------------------------------------
// matrices (column major, m[c] is column c)
// --------------------------------------------------------------------------------------------------------------------
struct mat4
{
  vec4 c[4];
  vec4&       operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
struct mat3
{
  vec3 c[3];
  mat3() = default;
  explicit mat3(const mat4& m) { for(int i = 0; i < 3; i++) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z); }
  vec3&       operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for(int i = 0; i < 4; i++) r.c[i] = a * b.c[i]; return r; }
inline vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
inline mat3 transpose(const mat3& m)
{
  mat3 r;
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
      r.c[i][j] = m.c[j][i];
  return r;
}
inline mat3 inverse(const mat3& m)
{
  vec3  r0 = cross(m.c[1], m.c[2]), r1 = cross(m.c[2], m.c[0]), r2 = cross(m.c[0], m.c[1]);
  float inv = 1.0f / dot(m.c[0], r0);
  mat3  r;
  r.c[0] = vec3(r0.x, r1.x, r2.x) * inv;
  r.c[1] = vec3(r0.y, r1.y, r2.y) * inv;
  r.c[2] = vec3(r0.z, r1.z, r2.z) * inv;
  return r;
}

// --------------------------------------------------------------------------------------------------------------------
// textures: handles owned by the harness
// --------------------------------------------------------------------------------------------------------------------
struct Texture2D
{
  uint32_t     width = 0, height = 0, mips = 1;
  const float* texels = nullptr;  // mip chain packed, level l has max(1, width>>l)^2 texels (HiZ is square)
};
typedef const Texture2D* sampler2D;

// DEFINED sampler for displacement textures: LOD 0, bilinear, repeat, float texels, fp32 weights
inline vec4 texture(sampler2D t, const vec2& uv)
{
  float x = uv.x * float(t->width) - 0.5f, y = uv.y * float(t->height) - 0.5f;
  float fx = std::floor(x), fy = std::floor(y);
  float ax = x - fx, ay = y - fy;
  int   w = int(t->width), h = int(t->height);
  int   x0 = int(fx) % w, y0 = int(fy) % h;
  if(x0 < 0) x0 += w;
  if(y0 < 0) y0 += h;
  int   x1 = x0 + 1 == w ? 0 : x0 + 1, y1 = y0 + 1 == h ? 0 : y0 + 1;
  float t00 = t->texels[size_t(y0) * w + x0], t10 = t->texels[size_t(y0) * w + x1];
  float t01 = t->texels[size_t(y1) * w + x0], t11 = t->texels[size_t(y1) * w + x1];
  float top = t00 + (t10 - t00) * ax, bot = t01 + (t11 - t01) * ax;
  return vec4(top + (bot - top) * ay, 0.0f, 0.0f, 1.0f);
}
// DEFINED sampler for the far HiZ: max of the bilinear footprint, nearest mip, clamp to edge (src/nvhiz_vk.cpp:83-115)
inline vec4 textureLod(sampler2D t, const vec2& uv, float lod)
{
  int level = 0;
  if(lod > 0.0f)
    level = min(int(lod), int(t->mips) - 1);
  uint32_t size = max(1u, t->width >> level);
  size_t   base = 0;
  for(int l = 0; l < level; l++)
  {
    size_t s = max(1u, t->width >> l);
    base += s * s;
  }
  float x = uv.x * float(size) - 0.5f, y = uv.y * float(size) - 0.5f;
  int   x0 = int(std::floor(x)), y0 = int(std::floor(y)), x1 = x0 + 1, y1 = y0 + 1;
  auto  cl = [&](int i) { return min(max(i, 0), int(size) - 1); };
  x0 = cl(x0); x1 = cl(x1); y0 = cl(y0); y1 = cl(y1);
  const float* p = t->texels + base;
  float a = p[size_t(y0) * size + x0], b = p[size_t(y0) * size + x1], d = p[size_t(y1) * size + x0], e = p[size_t(y1) * size + x1];
  return vec4(max(max(a, b), max(d, e)), 0.0f, 0.0f, 1.0f);
}

// texelFetchOffset: exact texel of one mip level.  A fetch outside the level is undefined in GLSL (robust buffer access
// returns 0); DEFINED as 0 here, as in the oracle and the kernels (nvhiz-update).
inline vec4 texelFetchOffset(sampler2D t, const ivec2& coord, int lod, const ivec2& offset)
{
  uint32_t w = t->mips > 1 || lod > 0 ? max(1u, t->width >> lod) : t->width, h = t->mips > 1 || lod > 0 ? max(1u, t->height >> lod) : t->height;
  size_t   base = 0;
  for(int l = 0; l < lod; l++)
    base += size_t(max(1u, t->width >> l)) * max(1u, t->height >> l);
  int x = coord.x + offset.x, y = coord.y + offset.y;
  float v = (x >= 0 && y >= 0 && uint32_t(x) < w && uint32_t(y) < h) ? t->texels[base + size_t(y) * w + x] : 0.0f;
  return vec4(v, 0.0f, 0.0f, 1.0f);
}
struct Image2D
{
  uint32_t width = 0, height = 0;
  float*   texels = nullptr;
};
typedef const Image2D* image2D;
inline void imageStore(image2D img, const ivec2& c, const vec4& v)  // writes outside the image are discarded
{
  if(img && c.x >= 0 && c.y >= 0 && uint32_t(c.x) < img->width && uint32_t(c.y) < img->height)
    img->texels[size_t(c.y) * img->width + c.x] = v.x;
}

// --------------------------------------------------------------------------------------------------------------------
// atomics (single host thread: plain read-modify-write)
// --------------------------------------------------------------------------------------------------------------------
template <class T, class V> inline T atomicAdd(T& mem, V v) { T old = mem; mem = T(old + T(v)); return old; }
template <class T, class V> inline T atomicMax(T& mem, V v) { T old = mem; if(T(v) > old) mem = T(v); return old; }
template <class T, class V> inline T atomicMin(T& mem, V v) { T old = mem; if(T(v) < old) mem = T(v); return old; }
template <class T, class V> inline T atomicOr(T& mem, V v) { T old = mem; mem = T(old | T(v)); return old; }
template <class T, class V> inline T atomicExchange(T& mem, V v) { T old = mem; mem = T(v); return old; }
template <class T> inline T atomicLoad(T& mem, int, int, int) { return mem; }
template <class T, class V> inline void atomicStore(T& mem, V v, int, int, int) { mem = T(v); }
inline void memoryBarrierShared() {}
inline void memoryBarrierBuffer() {}
template <class... A> inline void memoryBarrier(A...) {}
const int gl_ScopeDevice = 1, gl_ScopeSubgroup = 3, gl_SemanticsAcquire = 2, gl_SemanticsRelease = 4, gl_SemanticsAcquireRelease = 8,
          gl_StorageSemanticsBuffer = 0x40, gl_StorageSemanticsShared = 0x100;

// --------------------------------------------------------------------------------------------------------------------
// SIMT emulator
// --------------------------------------------------------------------------------------------------------------------
struct Fiber
{
#if GLSL_FAST_SWITCH
  void* sp;
#else
  ucontext_t ctx;
#endif
  uint       lane, subgroup;
  uvec3      localID, workGroupID, globalID;
  int        state;  // 0 runnable, 1 parked, 2 done
  int        site;
  bool       workgroupScope;
  const void* in;
  size_t     inSize;
  uint       groupMask;   // lanes of my subgroup resolved together with me
  const unsigned char* groupVals;
};

struct Simt
{
  static constexpr size_t kStack = 256 * 1024;
  static constexpr size_t kSlot  = 64;
  std::vector<Fiber>         fibers;
  std::vector<unsigned char> stacks;
  std::vector<unsigned char> scratch;  // [subgroup][lane][kSlot]
#if GLSL_FAST_SWITCH
  void* schedSp = nullptr;
  void  toScheduler(Fiber* f) { glsl_ctx_switch(&f->sp, schedSp); }
  void  toFiber(Fiber* f) { glsl_ctx_switch(&schedSp, f->sp); }
  void  initFiber(Fiber& f, unsigned char* stack)
  {
    uintptr_t top  = (reinterpret_cast<uintptr_t>(stack) + kStack) & ~uintptr_t(15);
    void**    slot = reinterpret_cast<void**>(top);
    slot[-1]       = nullptr;                                  // return address of the trampoline's imaginary caller
    slot[-2]       = reinterpret_cast<void*>(&Simt::trampoline);  // where the first switch "returns" to
    for(int i = 3; i <= 8; i++)
      slot[-i] = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = slot - 8;
  }
#else
  ucontext_t sched;
  void       toScheduler(Fiber* f) { swapcontext(&f->ctx, &sched); }
  void       toFiber(Fiber* f) { swapcontext(&sched, &f->ctx); }
  void       initFiber(Fiber& f, unsigned char* stack)
  {
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp   = stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link          = nullptr;
    makecontext(&f.ctx, &Simt::trampoline, 0);
  }
#endif
  Fiber*                     cur = nullptr;
  void (*entry)()              = nullptr;
  uint64_t collectives = 0, divergentGroups = 0;

  static Simt& get()
  {
    static Simt s;
    return s;
  }
  static void trampoline()
  {
    Simt& s = get();
    s.entry();
    s.cur->state = 2;
    s.toScheduler(s.cur);
    abort();  // a finished invocation is never resumed
  }
  void park(int site, bool wg, const void* in, size_t size)
  {
    if(size > kSlot)
    {
      fprintf(stderr, "glsl_shim: collective payload of %zu bytes\n", size);
      abort();
    }
    Fiber* f          = cur;
    f->state          = 1;
    f->site           = site;
    f->workgroupScope = wg;
    f->in             = in;
    f->inSize         = size;
    toScheduler(f);
  }
  void resolveGroup(uint first, uint last, int site, uint subgroupSize)
  {
    uint mask = 0, live = 0;
    for(uint i = first; i < last; i++)
    {
      Fiber& f = fibers[i];
      if(f.state != 2)
        live |= 1u << (i - first);
      if(f.state == 1 && f.site == site)
      {
        mask |= 1u << (i - first);
        memcpy(scratch.data() + size_t(i) * kSlot, f.in, f.inSize);
      }
    }
    if(!mask)
      return;
    collectives++;
    if(mask != live)
      divergentGroups++;
    for(uint i = first; i < last; i++)
      if(mask & (1u << (i - first)))
      {
        Fiber& f    = fibers[i];
        f.groupMask = mask;
        f.groupVals = scratch.data() + size_t(first) * kSlot;
        f.state     = 0;
      }
  }
  // run one workgroup of localSize invocations to completion
  void runWorkgroup(void (*fn)(), uint localSize, uint wgX, uint subgroupSize = 32, uint localSizeX = 0, uint wgY = 0)
  {
    entry = fn;
    if(!localSizeX)
      localSizeX = localSize;
    if(fibers.size() < localSize)
    {
      fibers.resize(localSize);
      stacks.resize(size_t(localSize) * kStack);
      scratch.resize(size_t(localSize) * kSlot);
    }
    for(uint i = 0; i < localSize; i++)
    {
      Fiber& f = fibers[i];
      initFiber(f, stacks.data() + size_t(i) * kStack);
      f.lane        = i % subgroupSize;
      f.subgroup    = i / subgroupSize;
      f.localID     = uvec3(i % localSizeX, i / localSizeX, 0);
      f.workGroupID = uvec3(wgX, wgY, 0);
      f.globalID    = uvec3(wgX * localSizeX + i % localSizeX, wgY * (localSize / localSizeX) + i / localSizeX, 0);
      f.state       = 0;
    }
    for(;;)
    {
      bool anyLive = false;
      for(uint i = 0; i < localSize; i++)
      {
        Fiber& f = fibers[i];
        if(f.state == 0)
        {
          cur = &f;
          toFiber(&f);
        }
        anyLive |= f.state != 2;
      }
      if(!anyLive)
        break;
      // every live invocation is parked.  Subgroup collectives first, lowest subgroup first (a subgroup runs ahead until it
      // needs the workgroup), and inside a subgroup the earliest call site; barriers when nothing else is left.
      uint numSubgroups = (localSize + subgroupSize - 1) / subgroupSize;
      bool resolved     = false;
      for(uint sg = 0; sg < numSubgroups && !resolved; sg++)
      {
        uint first = sg * subgroupSize, last = min(localSize, first + subgroupSize);
        int  site  = 0x7fffffff;
        for(uint i = first; i < last; i++)
          if(fibers[i].state == 1 && !fibers[i].workgroupScope && fibers[i].site < site)
            site = fibers[i].site;
        if(site == 0x7fffffff)
          continue;
        resolveGroup(first, last, site, subgroupSize);
        resolved = true;
      }
      if(!resolved)
      {
        int site = 0x7fffffff;
        for(uint i = 0; i < localSize; i++)
          if(fibers[i].state == 1 && fibers[i].site < site)
            site = fibers[i].site;
        for(uint sg = 0; sg < numSubgroups; sg++)
          resolveGroup(sg * subgroupSize, min(localSize, (sg + 1) * subgroupSize), site, subgroupSize);
      }
    }
    cur = nullptr;
  }
};

struct Group
{
  uint                 mask;
  const unsigned char* vals;
  template <class T> T val(uint lane) const
  {
    T v;
    memcpy(&v, vals + size_t(lane) * Simt::kSlot, sizeof(T));
    return v;
  }
};
template <class T> inline Group rendezvous(int site, const T& v, bool wg = false)
{
  Simt& s = Simt::get();
  s.park(site, wg, &v, sizeof(T));
  return Group{s.cur->groupMask, s.cur->groupVals};
}
inline uint simt_lane() { return Simt::get().cur->lane; }

inline void  barrier_(int site) { int d = 0; rendezvous(site, d, true); }
inline uvec4 subgroupBallot_(int site, bool v)
{
  Group g = rendezvous(site, v);
  uint  m = 0;
  for(uint l = 0; l < 32; l++)
    if((g.mask >> l & 1) && g.val<bool>(l))
      m |= 1u << l;
  return uvec4(m, 0u, 0u, 0u);
}
inline bool subgroupAny_(int site, bool v) { return subgroupBallot_(site, v).x != 0; }
inline bool subgroupAll_(int site, bool v)
{
  Group g = rendezvous(site, v);
  for(uint l = 0; l < 32; l++)
    if((g.mask >> l & 1) && !g.val<bool>(l))
      return false;
  return true;
}
inline bool subgroupElect_(int site)
{
  int   d = 0;
  Group g = rendezvous(site, d);
  return uint(__builtin_ctz(g.mask)) == simt_lane();
}
template <class T> inline T subgroupBroadcastFirst_(int site, T v)
{
  Group g = rendezvous(site, v);
  return g.val<T>(uint(__builtin_ctz(g.mask)));
}
template <class T, class I> inline T subgroupShuffle_(int site, T v, I src)
{
  Group g = rendezvous(site, v);
  uint  l = uint(src) & 31u;
  return (g.mask >> l & 1) ? g.val<T>(l) : v;  // reading an inactive lane is undefined in GLSL
}
template <class T, class I> inline T subgroupBroadcast_(int site, T v, I src) { return subgroupShuffle_(site, v, src); }
template <class T, class F> inline T subgroupReduce(int site, T v, F f, bool inclusive, bool exclusive)
{
  Group g    = rendezvous(site, v);
  uint  me   = simt_lane();
  bool  have = false;
  T     acc{};
  for(uint l = 0; l < 32; l++)
  {
    if(!(g.mask >> l & 1))
      continue;
    if(exclusive && l >= me)
      break;
    if(inclusive && l > me)
      break;
    T x  = g.val<T>(l);
    acc  = have ? f(acc, x) : x;
    have = true;
  }
  return acc;  // exclusive scan of the first lane: identity T{} (only used with add)
}
template <class T> inline T subgroupAdd_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return T(a + b); }, false, false); }
template <class T> inline T subgroupInclusiveAdd_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return T(a + b); }, true, false); }
template <class T> inline T subgroupExclusiveAdd_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return T(a + b); }, false, true); }
template <class T> inline T subgroupMax_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return a < b ? b : a; }, false, false); }
template <class T> inline T subgroupMin_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return b < a ? b : a; }, false, false); }
template <class T> inline T subgroupOr_(int site, T v) { return subgroupReduce(site, v, [](T a, T b) { return T(a | b); }, false, false); }
template <class T> inline uvec4 subgroupPartitionNV_(int site, T v)
{
  Group g = rendezvous(site, v);
  uint  m = 0;
  for(uint l = 0; l < 32; l++)
    if((g.mask >> l & 1) && g.val<T>(l) == v)
      m |= 1u << l;
  return uvec4(m, 0u, 0u, 0u);
}
// ballot helpers are not collectives
inline uint subgroupBallotBitCount(const uvec4& b) { return uint(__builtin_popcount(b.x)); }
inline uint subgroupBallotExclusiveBitCount(const uvec4& b) { return uint(__builtin_popcount(b.x & ((1u << simt_lane()) - 1u))); }
inline uint subgroupBallotInclusiveBitCount(const uvec4& b) { return uint(__builtin_popcount(b.x & ((2u << simt_lane()) - 1u))); }
inline uint subgroupBallotFindMSB(const uvec4& b) { return uint(findMSB(b.x)); }
inline uint subgroupBallotFindLSB(const uvec4& b) { return uint(findLSB(b.x)); }

}  // namespace glsl

#define gl_SubgroupInvocationID (::glsl::Simt::get().cur->lane)
#define gl_SubgroupID (::glsl::Simt::get().cur->subgroup)
#define gl_LocalInvocationID (::glsl::Simt::get().cur->localID)
#define gl_WorkGroupID (::glsl::Simt::get().cur->workGroupID)
#define gl_GlobalInvocationID (::glsl::Simt::get().cur->globalID)
#define gl_SubgroupSize 32u
#define gl_SubgroupLeMask (::glsl::uvec4((2u << ::glsl::Simt::get().cur->lane) - 1u, 0u, 0u, 0u))
#define gl_SubgroupLtMask (::glsl::uvec4((1u << ::glsl::Simt::get().cur->lane) - 1u, 0u, 0u, 0u))
