// axr mini-glm: the subset of the glm vector/matrix API that AxiomR's hot path touches.
//
// Why this exists: the reference builds against `external/glm`, which is git-ignored, not a
// submodule and not version-pinned (reference external/CMakeLists.txt:12-14). This header restates
// the *published scalar definitions* of glm 0.9.9.x / 1.0.x (default packed types => scalar code even
// with GLM_FORCE_AVX, reference include/math.hpp:9-12) operation by operation, in the same evaluation
// order, so that binary32 results are reproducible when compiled with -ffp-contract=off.
// It is shared by the host-side C++ mirror (axiomr_b200/host) and by the shimmed build of the
// unmodified reference sources (oracle/_ref). It is NOT a copy of glm: only the functions listed in
// SURVEY.md §8(c) are present.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <algorithm>

namespace glm {

typedef int length_t;

// ---------------------------------------------------------------- vec2
template <typename T>
struct tvec2 {
	union { T x, r, s; };
	union { T y, g, t; };
	tvec2() = default;
	constexpr explicit tvec2(T v) : x(v), y(v) {}
	constexpr tvec2(T a, T b) : x(a), y(b) {}
	template <typename U>
	constexpr explicit tvec2(const tvec2<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
	static constexpr length_t length() { return 2; }
	T& operator[](length_t i) { return i == 0 ? x : y; }
	const T& operator[](length_t i) const { return i == 0 ? x : y; }
	tvec2& operator+=(const tvec2& o) { x += o.x; y += o.y; return *this; }
	tvec2& operator-=(const tvec2& o) { x -= o.x; y -= o.y; return *this; }
	tvec2& operator*=(T v) { x *= v; y *= v; return *this; }
	tvec2& operator/=(T v) { x /= v; y /= v; return *this; }
};
template <typename T> constexpr tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a) { return tvec2<T>(-a.x, -a.y); }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <typename T> constexpr tvec2<T> operator*(T s, const tvec2<T>& a) { return tvec2<T>(s * a.x, s * a.y); }
template <typename T> constexpr tvec2<T> operator/(const tvec2<T>& a, T s) { return tvec2<T>(a.x / s, a.y / s); }
template <typename T> constexpr tvec2<T> operator/(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x / b.x, a.y / b.y); }
template <typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> constexpr bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }

// ---------------------------------------------------------------- vec3 / vec4
struct vec4;
struct vec3 {
	union { float x, r, s; };
	union { float y, g, t; };
	union { float z, b, p; };
	vec3() = default;
	constexpr explicit vec3(float v) : x(v), y(v), z(v) {}
	constexpr vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
	constexpr vec3(const tvec2<float>& v, float c) : x(v.x), y(v.y), z(c) {}
	constexpr explicit vec3(const vec4& v);
	static constexpr length_t length() { return 3; }  // component count, as in glm (reference src/mesh.cpp:283 relies on it)
	float& operator[](length_t i) { return i == 0 ? x : (i == 1 ? y : z); }
	const float& operator[](length_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
	vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	vec3& operator*=(float v) { x *= v; y *= v; z *= v; return *this; }
	vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
	vec3& operator/=(float v) { x /= v; y /= v; z /= v; return *this; }
};
struct vec4 {
	union { float x, r, s; };
	union { float y, g, t; };
	union { float z, b, p; };
	union { float w, a, q; };
	vec4() = default;
	constexpr explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
	constexpr vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
	constexpr vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
	static constexpr length_t length() { return 4; }
	float& operator[](length_t i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
	const float& operator[](length_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
	vec4& operator+=(const vec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
	vec4& operator*=(float v) { x *= v; y *= v; z *= v; w *= v; return *this; }
};
constexpr vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

constexpr vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
constexpr vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
constexpr vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
constexpr vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
constexpr vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
constexpr vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
constexpr vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
constexpr vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
constexpr vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
constexpr vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
constexpr bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
constexpr bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }

constexpr vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
constexpr vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
constexpr vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
constexpr vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
constexpr vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
constexpr vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
constexpr vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
constexpr bool operator==(const vec4& a, const vec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
constexpr bool operator!=(const vec4& a, const vec4& b) { return !(a == b); }

typedef tvec2<float> vec2;
typedef tvec2<int> ivec2;
typedef tvec2<unsigned int> uvec2;

// ---------------------------------------------------------------- scalar helpers
template <typename T> constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }
template <typename T> constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
constexpr float radians(float deg) { return deg * static_cast<float>(0.01745329251994329576923690768489); }
constexpr float degrees(float rad) { return rad * static_cast<float>(57.295779513082320876798154814105); }
inline float abs(float v) { return std::fabs(v); }
inline float asin(float v) { return std::asin(v); }
inline float acos(float v) { return std::acos(v); }
inline float sqrt(float v) { return std::sqrt(v); }
inline float inversesqrt(float v) { return 1.0f / std::sqrt(v); }
constexpr float min(float a, float b) { return (b < a) ? b : a; }
constexpr float max(float a, float b) { return (a < b) ? b : a; }
constexpr float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
constexpr float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float pow(float b, float e) { return std::pow(b, e); }

// ---------------------------------------------------------------- geometric (func_geometric.inl order)
constexpr float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
constexpr float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
constexpr float dot(const vec4& a, const vec4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
constexpr float length2(const vec3& v) { return dot(v, v); }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }
constexpr vec3 cross(const vec3& x, const vec3& y) {
	return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
constexpr vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }
constexpr vec2 mix(const vec2& x, const vec2& y, float a) { return x * (1.0f - a) + y * a; }
constexpr vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
constexpr vec4 mix(const vec4& x, const vec4& y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 pow(const vec3& b, const vec3& e) { return vec3(std::pow(b.x, e.x), std::pow(b.y, e.y), std::pow(b.z, e.z)); }
constexpr vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }

// ---------------------------------------------------------------- mat3 / mat4 (column-major)
struct mat4;
struct mat3 {
	vec3 value[3];
	mat3() = default;
	constexpr explicit mat3(float d) : value{vec3(d, 0, 0), vec3(0, d, 0), vec3(0, 0, d)} {}
	constexpr mat3(const vec3& a, const vec3& b, const vec3& c) : value{a, b, c} {}
	constexpr explicit mat3(const mat4& m);
	vec3& operator[](length_t i) { return value[i]; }
	constexpr const vec3& operator[](length_t i) const { return value[i]; }
};
struct mat4 {
	vec4 value[4];
	mat4() = default;
	constexpr explicit mat4(float d) : value{vec4(d, 0, 0, 0), vec4(0, d, 0, 0), vec4(0, 0, d, 0), vec4(0, 0, 0, d)} {}
	constexpr mat4(const vec4& a, const vec4& b, const vec4& c, const vec4& d) : value{a, b, c, d} {}
	constexpr explicit mat4(const mat3& m)
		: value{vec4(m[0], 0), vec4(m[1], 0), vec4(m[2], 0), vec4(0, 0, 0, 1)} {}
	vec4& operator[](length_t i) { return value[i]; }
	constexpr const vec4& operator[](length_t i) const { return value[i]; }
};
constexpr mat3::mat3(const mat4& m) : value{vec3(m[0]), vec3(m[1]), vec3(m[2])} {}

constexpr mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
constexpr mat3 operator*(float s, const mat3& m) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
constexpr mat3 operator+(const mat3& a, const mat3& b) { return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
constexpr vec3 operator*(const mat3& m, const vec3& v) {
	return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
	            m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
	            m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
constexpr mat3 transpose(const mat3& m) {
	return mat3(vec3(m[0][0], m[1][0], m[2][0]), vec3(m[0][1], m[1][1], m[2][1]), vec3(m[0][2], m[1][2], m[2][2]));
}
constexpr mat4 operator*(const mat4& m, float s) { return mat4(m[0] * s, m[1] * s, m[2] * s, m[3] * s); }
// type_mat4x4.inl: (m0*x + m1*y) + (m2*z + m3*w)
constexpr vec4 operator*(const mat4& m, const vec4& v) {
	return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w);
}
// type_mat4x4.inl: ((a0*b.x + a1*b.y) + a2*b.z) + a3*b.w per column
constexpr mat4 operator*(const mat4& a, const mat4& b) {
	return mat4(a[0] * b[0].x + a[1] * b[0].y + a[2] * b[0].z + a[3] * b[0].w,
	            a[0] * b[1].x + a[1] * b[1].y + a[2] * b[1].z + a[3] * b[1].w,
	            a[0] * b[2].x + a[1] * b[2].y + a[2] * b[2].z + a[3] * b[2].w,
	            a[0] * b[3].x + a[1] * b[3].y + a[2] * b[3].z + a[3] * b[3].w);
}
constexpr mat4 transpose(const mat4& m) {
	return mat4(vec4(m[0][0], m[1][0], m[2][0], m[3][0]), vec4(m[0][1], m[1][1], m[2][1], m[3][1]),
	            vec4(m[0][2], m[1][2], m[2][2], m[3][2]), vec4(m[0][3], m[1][3], m[2][3], m[3][3]));
}
// func_matrix.inl compute_inverse<4,4>: cofactor expansion, determinant from first column dot.
inline mat4 inverse(const mat4& m) {
	float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
	vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
	vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
	vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
	vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
	vec4 Fac5(Coef20, Coef20, Coef22, Coef23);
	vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
	vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
	vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
	vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
	vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
	vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
	vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
	vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
	vec4 SignA(+1, -1, +1, -1);
	vec4 SignB(-1, +1, -1, +1);
	mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
	vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
	vec4 Dot0(m[0] * Row0);
	float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
	float OneOverDeterminant = 1.0f / Dot1;
	return Inverse * OneOverDeterminant;
}
// gtc/matrix_transform
inline mat4 translate(const mat4& m, const vec3& v) {
	mat4 r(m);
	r[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
	return r;
}
inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
	float const a = angle;
	float const c = std::cos(a);
	float const s = std::sin(a);
	vec3 axis(normalize(v));
	vec3 temp((1.0f - c) * axis);
	mat4 Rotate(0.0f);
	Rotate[0][0] = c + temp[0] * axis[0];
	Rotate[0][1] = temp[0] * axis[1] + s * axis[2];
	Rotate[0][2] = temp[0] * axis[2] - s * axis[1];
	Rotate[1][0] = temp[1] * axis[0] - s * axis[2];
	Rotate[1][1] = c + temp[1] * axis[1];
	Rotate[1][2] = temp[1] * axis[2] + s * axis[0];
	Rotate[2][0] = temp[2] * axis[0] + s * axis[1];
	Rotate[2][1] = temp[2] * axis[1] - s * axis[0];
	Rotate[2][2] = c + temp[2] * axis[2];
	mat4 Result(0.0f);
	Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
	Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
	Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
	Result[3] = m[3];
	return Result;
}
// perspectiveRH_NO: right-handed, clip z in [-1,1] (glm default)
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
	float const tanHalfFovy = std::tan(fovy / 2.0f);
	mat4 Result(0.0f);
	Result[0][0] = 1.0f / (aspect * tanHalfFovy);
	Result[1][1] = 1.0f / (tanHalfFovy);
	Result[2][2] = -(zFar + zNear) / (zFar - zNear);
	Result[2][3] = -1.0f;
	Result[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
	return Result;
}

// ---------------------------------------------------------------- quaternion (gtc/quaternion, gtx/quaternion)
struct quat {
	float x, y, z, w;
	quat() = default;
	constexpr quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
	constexpr quat(float s, const vec3& v) : x(v.x), y(v.y), z(v.z), w(s) {}
};
constexpr float dot(const quat& a, const quat& b) { return (a.w * b.w + a.x * b.x) + (a.y * b.y + a.z * b.z); }
inline float length(const quat& q) { return std::sqrt(dot(q, q)); }
inline quat normalize(const quat& q) {
	float len = length(q);
	if (len <= 0.0f) return quat(1, 0, 0, 0);
	float oneOverLen = 1.0f / len;
	return quat(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
}
constexpr quat conjugate(const quat& q) { return quat(q.w, -q.x, -q.y, -q.z); }
constexpr quat operator/(const quat& q, float s) { return quat(q.w / s, q.x / s, q.y / s, q.z / s); }
constexpr quat inverse(const quat& q) { return conjugate(q) / dot(q, q); }
constexpr quat operator*(const quat& p, const quat& q) {
	return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z,
	            p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
	            p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z,
	            p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
constexpr vec3 operator*(const quat& q, const vec3& v) {
	vec3 const QuatVector(q.x, q.y, q.z);
	vec3 const uv(cross(QuatVector, v));
	vec3 const uuv(cross(QuatVector, uv));
	return v + ((uv * q.w) + uuv) * 2.0f;
}
inline quat angleAxis(float angle, const vec3& v) {
	float const s = std::sin(angle * 0.5f);
	return quat(std::cos(angle * 0.5f), v * s);
}
constexpr mat3 mat3_cast(const quat& q) {
	float qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z);
	float qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
	float qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
	return mat3(vec3(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy)),
	            vec3(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx)),
	            vec3(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy)));
}
constexpr mat4 mat4_cast(const quat& q) { return mat4(mat3_cast(q)); }
inline quat rotation(const vec3& orig, const vec3& dest) {
	float cosTheta = dot(orig, dest);
	vec3 rotationAxis;
	if (cosTheta >= 1.0f - epsilon<float>()) return quat(1, 0, 0, 0);
	if (cosTheta < -1.0f + epsilon<float>()) {
		rotationAxis = cross(vec3(0, 0, 1), orig);
		if (length2(rotationAxis) < epsilon<float>()) rotationAxis = cross(vec3(1, 0, 0), orig);
		rotationAxis = normalize(rotationAxis);
		return angleAxis(pi<float>(), rotationAxis);
	}
	rotationAxis = cross(orig, dest);
	float s = std::sqrt((1.0f + cosTheta) * 2.0f);
	float invs = 1.0f / s;
	return quat(s * 0.5f, rotationAxis.x * invs, rotationAxis.y * invs, rotationAxis.z * invs);
}

// gtc/type_ptr
inline const float* value_ptr(const mat4& m) { return &m.value[0].x; }
inline const float* value_ptr(const mat3& m) { return &m.value[0].x; }
inline const float* value_ptr(const vec3& v) { return &v.x; }
inline const float* value_ptr(const vec4& v) { return &v.x; }

}  // namespace glm
